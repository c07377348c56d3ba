// Per-particle Dense on the 5th-generation tensor cores (tcgen05.mma, accumulator in tensor memory): the "tensor cores for the
// dense per-particle MLP between conv layers" of the north star (tf.keras.layers.Dense at models/pbf_model.py:140-152,
// models/hrnet.py:63-66; layer-by-layer mode -- in the fused step the Dense rows ride inside the conv kernels).
//
//   out[n, :] = g(x[n, :]) @ W[cin, cout] + b,     float32 in, float32 out, 3xTF32 inside:
//   x = x_hi + x_lo, W = W_hi + W_lo with *_hi = the upper 19 bits (what kind::tf32 reads of a float32 word), *_lo the exact
//   remainder;  D = x_hi W_hi + x_lo W_hi + x_hi W_lo  accumulated in float32 by the tensor core (the dropped x_lo W_lo term
//   is 2^-22 relative): float32-level parity with the SIMT kernel k_dense, which the tests hold it to.
//
// One CTA (128 threads) owns tiles of 128 rows: thread t stages row t (relu, hi / lo split) into the canonical K-major
// no-swizzle UMMA operand layout -- 8-row x 16-byte core matrices, [k chunk][row group][8][4 floats], so the 32 threads of a
// warp write 512 contiguous bytes per chunk -- one elected thread issues 3 MMAs (M = 128, N = cout rounded up to 16, K = 8)
// per k-step and commits them to an mbarrier; every warp then reads its 32 lanes of the accumulator with tcgen05.ld, adds the
// bias and stores the row.  HBM bound (4 (cin + cout) bytes per row): several CTAs per SM keep enough rows in flight.
#include "cconv_common.cuh"

namespace dmcf {

namespace umma {

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): 16-byte units; LBO = distance between the
// two 16-byte K chunks of one MMA, SBO = distance between 8-row groups; version 1 (Blackwell), layout type 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a lost MMA must end as an error, not as a hung GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

}  // namespace umma

// kp = cin rounded up to 8 (K of the MMAs), np = cout rounded up to 16 (N), tmem_cols = power of two >= max(np, 32).
// Shared memory: A_hi | A_lo ([kp/4][16][8][4] floats each) | B_hi | B_lo ([kp/4][np/8][8][4] floats each, B[n][k] = W[k][n]).
__global__ void __launch_bounds__(128) k_dense_umma(const float* __restrict__ x, int64_t n, int cin, int64_t x_stride,
                                                      const float* __restrict__ w, const float* __restrict__ b, int cout,
                                                      int relu_input, float* __restrict__ out, int64_t out_stride, int kp, int np,
                                                      int tmem_cols) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_words = kp * 128, b_words = kp * np;
    float* a_hi = smem;
    float* a_lo = a_hi + a_words;
    float* b_hi = a_lo + a_words;
    float* b_lo = b_hi + b_words;
    const uint32_t bar_addr = tma::smem_u32(&bar);
    if (warp == 0) umma::tmem_alloc(tma::smem_u32(&tmem_base_slot), (uint32_t)tmem_cols);
    if (tid == 0) {
        tma::mbar_init(bar_addr, 1);
        tma::mbar_fence_init();
    }
    // B = W^T, split once per CTA: element (nn, k) of core matrix (k / 4, nn / 8)
    for (int i = tid; i < kp * np; i += 128) {
        const int k = i / np, nn = i % np;
        const float v = (k < cin && nn < cout) ? __ldg(w + (int64_t)k * cout + nn) : 0.0f;
        const float hi = umma::tf32_hi(v);
        const int idx = (((k >> 2) * (np >> 3) + (nn >> 3)) * 8 + (nn & 7)) * 4 + (k & 3);
        b_hi[idx] = hi;
        b_lo[idx] = v - hi;
    }
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = umma::instr_desc_tf32(128, np);
    const uint32_t a_lbo = 16 * 128, b_lbo = (uint32_t)(np >> 3) * 128;  // bytes between the K chunks of an operand
    const bool vec = (cin & 3) == 0 && (x_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool vec_out = (out_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int64_t n_tiles = (n + 127) / 128;
    uint32_t parity = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- stage row `tid` of the tile: relu, hi / lo split, UMMA layout ----
        const int64_t r = tile * 128 + tid;
        const bool r_ok = r < n;
        const float* xr = x + r * x_stride;
        for (int q = 0; q < (kp >> 2); ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r_ok) {
                const int k = q * 4;
                if (vec && k + 3 < cin) {
                    v = __ldg(reinterpret_cast<const float4*>(xr + k));
                } else {
                    if (k < cin) v.x = __ldg(xr + k);
                    if (k + 1 < cin) v.y = __ldg(xr + k + 1);
                    if (k + 2 < cin) v.z = __ldg(xr + k + 2);
                    if (k + 3 < cin) v.w = __ldg(xr + k + 3);
                }
                if (relu_input) {
                    v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
                }
            }
            const float4 hi = make_float4(umma::tf32_hi(v.x), umma::tf32_hi(v.y), umma::tf32_hi(v.z), umma::tf32_hi(v.w));
            const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
            const int idx = ((q * 16 + (tid >> 3)) * 8 + (tid & 7)) * 4;
            *reinterpret_cast<float4*>(a_hi + idx) = hi;
            *reinterpret_cast<float4*>(a_lo + idx) = lo;
        }
        umma::fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        umma::fence_before();
        __syncthreads();
        // ---- one thread issues the MMAs of the tile ----
        if (tid == 0) {
            umma::fence_after();
            const uint32_t ah = tma::smem_u32(a_hi), al = tma::smem_u32(a_lo), bh = tma::smem_u32(b_hi), bl = tma::smem_u32(b_lo);
            for (int s = 0; s < (kp >> 3); ++s) {
                const uint32_t ao = (uint32_t)s * 2 * a_lbo, bo = (uint32_t)s * 2 * b_lbo;
                const uint64_t d_ah = umma::smem_desc(ah + ao, a_lbo, 128), d_al = umma::smem_desc(al + ao, a_lbo, 128);
                const uint64_t d_bh = umma::smem_desc(bh + bo, b_lbo, 128), d_bl = umma::smem_desc(bl + bo, b_lbo, 128);
                umma::mma_tf32(tmem_d, d_ah, d_bh, idesc, s > 0);
                umma::mma_tf32(tmem_d, d_al, d_bh, idesc, 1);
                umma::mma_tf32(tmem_d, d_ah, d_bl, idesc, 1);
            }
            umma::commit(bar_addr);  // arrives when the MMAs above have completed (implies fence::before_thread_sync)
        }
        // ---- epilogue: warp w owns accumulator lanes 32 w .. 32 w + 31 = rows of the tile ----
        if (!umma::mbar_wait_bounded(bar_addr, parity)) __trap();  // the launch fails loudly instead of hanging the GPU
        parity ^= 1;
        umma::fence_after();
        const uint32_t t_row = tmem_d + ((uint32_t)(warp * 32) << 16);
        float* orow = out + r * out_stride;
        for (int c0 = 0; c0 < cout; c0 += 8) {
            float v[8];
            umma::tmem_ld8(t_row + (uint32_t)c0, v);  // warp-collective: every lane takes part, rows beyond n are not stored
            if (r_ok) {
                if (b) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (c0 + i < cout) v[i] += __ldg(b + c0 + i);
                }
                if (vec_out && c0 + 8 <= cout) {
                    *reinterpret_cast<float4*>(orow + c0) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(orow + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (c0 + i < cout) orow[c0 + i] = v[i];
                }
            }
        }
        umma::fence_before();
        __syncthreads();  // accumulator and operand tiles are free for the next tile
        umma::fence_after();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
}

// Tries the tensor-core kernel; *handled = false means "not eligible" (the caller runs k_dense).
int launch_dense_umma(const float* x, int64_t n, int cin, int64_t x_stride, const float* w, const float* b, int cout, int relu_input,
                      float* out, int64_t out_stride, cudaStream_t st, bool* handled) {
    *handled = false;
    if (n < 4096 || cin > 128 || cout > 128) return DMCF_OK;
    const int kp = (cin + 7) & ~7, np = (cout + 15) & ~15;
    int tmem_cols = 32;
    while (tmem_cols < np) tmem_cols <<= 1;
    const size_t smem = (size_t)(2 * kp * 128 + 2 * kp * np) * sizeof(float);
    if (smem > 160 * 1024) return DMCF_OK;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_dense_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_dense_umma)");
        attr_set = true;
    }
    *handled = true;
    // CTAs per SM: shared memory, and tensor memory (512 columns per SM: an allocation that does not fit would wait for ever)
    int per_sm = (int)((227 * 1024) / (smem + 2048));
    if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (n + 127) / 128;
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    k_dense_umma<<<(unsigned)blocks, 128, smem, st>>>(x, n, cin, x_stride, w, b, cout, relu_input, out, out_stride, kp, np, tmem_cols);
    DMCF_LAUNCH_CHECK("k_dense_umma");
    return DMCF_OK;
}


// ---- measured experiment (dmcf_umma_probe): the tensor-core shape a conv's patch x filter product would have -------------------
// D[64 x N] (+)= A_s[64 x 16] . B_s[N x 16]^T, kind::f16, for s = 0 .. ks-1 and `n_tiles` tiles per CTA: the A operand
// (filter rows: [F_hi ; F_lo] of 32 output channels, 2 KB per k-step) STREAMS from global memory / L2 through a ring of
// `stages` bulk copies (cp.async.bulk completing on mbarriers, slots released by tcgen05.commit), the B operand (the patch
// tile: N points x K, hi and lo halves) is resident in shared memory.  One thread issues the copies, one the MMAs; the other
// warps only read the accumulator of the last tile (which lets the host check the numerics and the M = 64 lane layout).
// Timed from the host over all 148 SMs streaming the same filter: the L2 -> shared-memory stream is what bounds it.
namespace umma {
__host__ __device__ constexpr uint32_t instr_desc_f16(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
}  // namespace umma

__global__ void __launch_bounds__(128, 1) k_umma_probe(const uint16_t* __restrict__ a_stream, const uint16_t* __restrict__ b_tile,
                                                        int ks, int n, int n_tiles, int stages, int passes,
                                                        float* __restrict__ d_out, long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char psm[];
    __shared__ __align__(8) uint64_t bars[2 * 64 + 1];  // full[stages], empty[stages], done
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ng = n >> 3;                           // row groups of the B operand
    const uint32_t b_lbo = (uint32_t)ng * 128 + 16;  // K-chunk stride, padded by 16 bytes (bank spread of the patch stores)
    unsigned char* ring = psm;                       // [stages][2048]
    unsigned char* btile = psm + (size_t)stages * 2048;  // [passes][ks * 2 chunks][b_lbo]
    const size_t b_bytes = (size_t)ks * 2 * b_lbo;
    if (warp == 0) umma::tmem_alloc(tma::smem_u32(&tmem_base_slot), 32);
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) {
            tma::mbar_init(tma::smem_u32(&bars[i]), 1);
            tma::mbar_init(tma::smem_u32(&bars[64 + i]), 1);
        }
        tma::mbar_init(tma::smem_u32(&bars[128]), 1);
        tma::mbar_fence_init();
    }
    for (size_t i = tid; i < (size_t)passes * b_bytes / 4; i += 128)
        reinterpret_cast<uint32_t*>(btile)[i] = __ldg(reinterpret_cast<const uint32_t*>(b_tile) + i);
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    if (warp == 0 && lane == 0) {  // producer: the filter stream
        uint32_t it = 0;
        long long t_wait = 0, t_copy = 0;
        for (int t = 0; t < n_tiles; ++t)
            for (int s = 0; s < ks; ++s, ++it) {
                const int st = it % stages;
                const uint32_t ph = (it / stages) & 1;
                const long long c0 = clock64();
                if (it >= (uint32_t)stages && !umma::mbar_wait_bounded(tma::smem_u32(&bars[64 + st]), ph ^ 1)) __trap();
                const long long c1 = clock64();
                tma::mbar_expect_tx(tma::smem_u32(&bars[st]), 2048);
                tma::bulk_g2s(tma::smem_u32(ring + (size_t)st * 2048), a_stream + (size_t)s * 1024, 2048, tma::smem_u32(&bars[st]));
                const long long c2 = clock64();
                t_wait += c1 - c0; t_copy += c2 - c1;
            }
        if (stats && blockIdx.x == 0) { stats[0] = t_wait; stats[1] = t_copy; }
    } else if (warp == 1 && lane == 0) {  // MMA issuer
        long long t_wait = 0, t_mma = 0, t_commit = 0;
        const uint32_t idesc = umma::instr_desc_f16(64, n);
        uint32_t it = 0;
        for (int t = 0; t < n_tiles; ++t) {
            for (int s = 0; s < ks; ++s, ++it) {
                const int st = it % stages;
                const uint32_t ph = (it / stages) & 1;
                const long long c0 = clock64();
                if (!umma::mbar_wait_bounded(tma::smem_u32(&bars[st]), ph)) __trap();
                umma::fence_after();
                const long long c1 = clock64();
                const uint64_t da = umma::smem_desc(tma::smem_u32(ring + (size_t)st * 2048), 1024, 128);
                for (int q = 0; q < passes; ++q) {
                    const uint64_t db = umma::smem_desc(tma::smem_u32(btile + q * b_bytes + (size_t)s * 2 * b_lbo), b_lbo, 128);
                    umma::mma_f16(tmem_d, da, db, idesc, (s > 0 || q > 0) ? 1u : 0u);
                }
                const long long c2 = clock64();
                umma::commit(tma::smem_u32(&bars[64 + st]));  // the slot is free once these MMAs have read it
                const long long c3 = clock64();
                t_wait += c1 - c0; t_mma += c2 - c1; t_commit += c3 - c2;
            }
        }
        umma::commit(tma::smem_u32(&bars[128]));
        if (stats && blockIdx.x == 0) { stats[2] = t_wait; stats[3] = t_mma; stats[4] = t_commit; }
    }
    // the idle lanes of the two role warps park HERE: a lane spinning in mbarrier.try_wait suspends its whole warp for the wait's
    // time limit and would stall the issuing lane of the same warp on every iteration (measured: 1.7 us per k-step)
    __syncwarp();
    if (!umma::mbar_wait_bounded(tma::smem_u32(&bars[128]), 0)) __trap();
    umma::fence_after();
    __syncwarp();
    if (d_out && blockIdx.x == 0) {  // all 128 accumulator lanes x n columns, raw (the host maps rows to lanes)
        for (int c0 = 0; c0 < n; c0 += 8) {
            float v[8];
            umma::tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int i = 0; i < 8; ++i) d_out[(size_t)tid * n + c0 + i] = v[i];
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, 32);
}

}  // namespace dmcf

extern "C" int dmcf_umma_probe(const void* a_stream, const void* b_tile, int32_t ks, int32_t n, int32_t n_tiles, int32_t stages,
                               int32_t passes, int32_t n_ctas, float* d_out, long long* stats, void* stream) {
    using namespace dmcf;
    DMCF_REQUIRE(a_stream && b_tile && ks >= 1 && n >= 8 && n <= 32 && (n & 7) == 0 && n_tiles >= 1 && stages >= 1 && stages <= 64 &&
                     passes >= 1 && passes <= 2 && n_ctas >= 1,
                 "umma_probe: bad arguments");
    const size_t smem = (size_t)stages * 2048 + (size_t)passes * ks * 2 * ((n >> 3) * 128 + 16);
    DMCF_REQUIRE(smem <= 220 * 1024, "umma_probe: %zu bytes of shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_umma_probe)");
        attr_set = true;
    }
    k_umma_probe<<<(unsigned)n_ctas, 128, smem, (cudaStream_t)stream>>>((const uint16_t*)a_stream, (const uint16_t*)b_tile, ks, n,
                                                                         n_tiles, stages, passes, d_out, stats);
    DMCF_LAUNCH_CHECK("k_umma_probe");
    return DMCF_OK;
}
