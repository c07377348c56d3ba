// Shared pieces of the continuous_conv kernels: launch parameters and the patch x filter product + epilogue.
#pragma once
#include "cconv_geom.cuh"

namespace dmcf {

struct ConvParams {
    GeomParams gp;
    int cin, cout;
    int normalize, window;
    float window_fac, r2;
    int relu_input;
    float feat_scale;
    int ascc, skip_self, nbr_lo, nbr_hi, dense_cin, accumulate;
    int kc_conv, kc, kc_pad;  // patch columns: conv part, conv+dense, padded to 4
    int debug_wrap_w;         // timing experiment switches (dmcf_set_kernel_options bits 8+), never set in production
    int filter_antisym;       // desc flag: filters[rev(cell)] == -filters[cell]
    int use_zsplit;           // option bit 2: run 4x4x4 wide layers as two z-half launches (2 CTAs/SM)
    int lean_cta_per_tile;    // option bit 14: the tensor-core k_cconv_lean launches one CTA per tile instead of persistent CTAs (A/B)
    int lean_tc_16;           // option bit 16: tensor-core phase 2 on a 16-point / 16-warp tile instead of 24 points / 12 warps (A/B)
    int no_lean_tc;           // option bit 15: k_cconv_lean keeps the FFMA2 phase 2 (A/B switch for the tensor-core phase 2)
    int no_multipair;         // option bit 5: k_cconv_lean keeps the one-pair-per-step walk for narrow inputs (A/B switch)
    int cip, cp;              // pow2 lane groupings for input / output channels
    int blk_ca, blk_na, blk_nb;  // block-diagonal promise (dmcf_conv_desc::block_cin / block_cout), 0 = none
    const float* filters;
    const float* out_pos;
    const float* inp_pos;
    const float* inp_feat;
    int64_t inp_stride;
    int64_t n_out, n_inp;
    const int32_t* n_out_dev;  // optional device-side out point count (n_out is then the capacity the grid was sized for)
    const float* inp_importance;
    const int32_t* nbr_index;
    const int64_t* row_splits;
    const float* nbr_importance;
    const float* bias;
    const float* dense_inp;
    int64_t dense_stride;
    const float* residual;
    int64_t residual_stride;
    float* out;
    int64_t out_stride;
    // optional precomputed pair records (dmcf_cconv_prepare): 9 arrays of n_pairs words
    // {row, i0, i1, wx0, wx1, wy0, wy1, wz0*a, wz1*a}; row < 0 marks a dropped pair
    const float* records;
    int64_t n_pairs;
    // dmcf_cconv_patches: stop after phase 1 and write the patch rows [n_out, kc_conv] here (filters is not read)
    float* patch_out;
    int64_t patch_stride;
};

static constexpr int kRecordFields = 9;

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier -----------------------------------------------
// Used to stage a kernel's RESIDENT filter (k_cconv_direct, k_cconv_apatch: up to ~100 KB that stay in shared memory for the
// lifetime of a persistent CTA) with a few bulk copies issued by one thread instead of a grid-stride loop of LDG + STS through
// the registers of every thread.
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "LAB_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// Whole-CTA helper: `n_words` floats global -> shared.  TMA path when the block is a multiple of 16 bytes and both ends are
// 16-byte aligned (every shipped filter), else the plain loop.  Ends with the data visible to every thread of the CTA.
__device__ __forceinline__ void stage_block(float* dst, const float* src, int n_words, uint64_t* bar, int n_threads) {
    const bool bulk = n_words > 0 && (n_words & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0 && n_words < (1 << 18);
    if (!bulk) {
        for (int i = threadIdx.x; i < n_words; i += n_threads) dst[i] = __ldg(src + i);
        __syncthreads();
        return;
    }
    const uint32_t b = smem_u32(bar);
    if (threadIdx.x == 0) {
        mbar_init(b, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)n_words * 4u;
        mbar_expect_tx(b, bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
            bulk_g2s(smem_u32(dst) + off, reinterpret_cast<const char*>(src) + off, n, b);
        }
    }
    mbar_wait(b, 0);
}
}  // namespace tma

// Number of out points this launch really has: the device-side count when the caller runs capacity-sized buffers.
__device__ __forceinline__ int64_t conv_n_out(const ConvParams& p) {
    if (p.n_out_dev == nullptr) return p.n_out;
    const int64_t n = (int64_t)__ldg(p.n_out_dev);
    return n < p.n_out ? (n < 0 ? 0 : n) : p.n_out;
}

// One neighbour pair of an out point, ready for the scatter: feature row, corner cells and per-axis corner weights
// with the pair's importance (window * neighbour importance * input importance) folded into the z weights.
struct PairRec {
    int row;      // input feature/position row, -1 = dropped (sub-range filter, skip-self, out of the list)
    PairGeom g;
    float norm;   // contribution to the normaliser (sum of neighbour importances / neighbour count)
};

// Evaluates pair `n` of the CSR list for the out point at (ox,oy,oz); `idx` = nbr_index[n], loaded by the caller (a kernel
// that evaluates the geometry itself fetches the indices of the NEXT chunk while it works on this one, so that the dependent
// position gather is the only exposed round trip).
__device__ __forceinline__ PairRec eval_pair_idx(const ConvParams& p, int64_t n, bool in_range, int idx, float ox, float oy, float oz) {
    PairRec r;
    r.row = -1;
    r.norm = 0.0f;
    r.g.i0 = r.g.i1 = 0;
    r.g.wx0 = r.g.wx1 = r.g.wy0 = r.g.wy1 = r.g.wz0 = r.g.wz1 = 0.0f;
    if (!in_range) return r;
    const bool filter_nbr = p.nbr_hi > p.nbr_lo;
    if (filter_nbr && (idx < p.nbr_lo || idx >= p.nbr_hi)) return r;
    const int row = filter_nbr ? idx - p.nbr_lo : idx;
    const float dx = __ldg(p.inp_pos + 3 * (int64_t)row) - ox;
    const float dy = __ldg(p.inp_pos + 3 * (int64_t)row + 1) - oy;
    const float dz = __ldg(p.inp_pos + 3 * (int64_t)row + 2) - oz;
    if (p.skip_self && dx == 0.0f && dy == 0.0f && dz == 0.0f) return r;
    float a = 1.0f;
    if (p.nbr_importance) {
        a = __ldg(p.nbr_importance + n);
    } else if (p.window != DMCF_WIN_NONE) {
        const float q = __fdiv_rn(dist2_exact(dx, dy, dz), p.r2);
        a = window_value(p.window, p.window_fac, q);
    }
    r.norm = (p.nbr_importance || p.window != DMCF_WIN_NONE) ? a : 1.0f;
    if (p.inp_importance) a *= __ldg(p.inp_importance + row);
    r.g = pair_geometry(p.gp, dx, dy, dz);
    r.g.wz0 *= a;
    r.g.wz1 *= a;
    r.row = row;
    return r;
}

__device__ __forceinline__ PairRec eval_pair(const ConvParams& p, int64_t n, bool in_range, float ox, float oy, float oz) {
    return eval_pair_idx(p, n, in_range, in_range ? __ldg(p.nbr_index + n) : 0, ox, oy, oz);
}

// Same pair from the precomputed record arrays (normaliser not stored: prepare refuses `normalize`).
__device__ __forceinline__ PairRec load_pair(const ConvParams& p, int64_t n, bool in_range) {
    PairRec r;
    r.row = -1;
    r.norm = 0.0f;
    r.g.i0 = r.g.i1 = 0;
    r.g.wx0 = r.g.wx1 = r.g.wy0 = r.g.wy1 = r.g.wz0 = r.g.wz1 = 0.0f;
    if (!in_range) return r;
    const float* f = p.records + n;
    const int64_t P = p.n_pairs;
    // all nine fields in flight together (no early return on a dropped pair: that would serialise two round trips)
    r.row = __float_as_int(__ldg(f));
    r.g.i0 = __float_as_int(__ldg(f + P));
    r.g.i1 = __float_as_int(__ldg(f + 2 * P));
    r.g.wx0 = __ldg(f + 3 * P); r.g.wx1 = __ldg(f + 4 * P);
    r.g.wy0 = __ldg(f + 5 * P); r.g.wy1 = __ldg(f + 6 * P);
    r.g.wz0 = __ldg(f + 7 * P); r.g.wz1 = __ldg(f + 8 * P);
    if (r.row < 0) {  // dropped pair: same all-zero record as eval_pair returns
        r.g.i0 = r.g.i1 = 0;
        r.g.wx0 = r.g.wx1 = r.g.wy0 = r.g.wy1 = r.g.wz0 = r.g.wz1 = 0.0f;
    }
    return r;
}

__device__ __forceinline__ PairRec pair_record(const ConvParams& p, int64_t n, bool in_range, float ox, float oy, float oz) {
    return p.records ? load_pair(p, n, in_range) : eval_pair(p, n, in_range, ox, oy, oz);
}

// ---- phase 2 + epilogue: [MT x kc] x [kc x cout], split-K over warps, lane = (k sub-slice, output channel) ----------
// `red` may alias `patch` when RED_ALIASES_PATCH (only valid for cout <= 32: the patch is dead once every warp has
// left the k loop, which the extra barrier guarantees).
template <int MT, int NW, bool RED_ALIASES_PATCH>
__device__ __forceinline__ void cconv_phase2(const ConvParams& p, const float* patch, float* red, const float* norm,
                                             int64_t tile_base, int64_t n_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cp = p.cp, ks = lane / cp, cl = lane % cp, n_ks = 32 / cp;
    const int kq_total = p.kc_pad / 4;
    const int kq_stride = NW * n_ks;
    for (int cb = 0; cb < p.cout; cb += 32) {
        const int co = cb + cl;
        const bool co_ok = co < p.cout;
        float acc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
        auto load_w = [&](int kq, float& w0, float& w1, float& w2, float& w3) {
            const int k = kq * 4;
            const float* wrow = p.filters + (int64_t)k * p.cout + co;
            w0 = (co_ok && k + 0 < p.kc) ? __ldg(wrow) : 0.0f;
            w1 = (co_ok && k + 1 < p.kc) ? __ldg(wrow + p.cout) : 0.0f;
            w2 = (co_ok && k + 2 < p.kc) ? __ldg(wrow + 2 * p.cout) : 0.0f;
            w3 = (co_ok && k + 3 < p.kc) ? __ldg(wrow + 3 * p.cout) : 0.0f;
        };
        int kq = warp * n_ks + ks;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
        if (kq < kq_total) load_w(kq, w0, w1, w2, w3);
        while (kq < kq_total) {
            // register double buffer: the next filter rows travel from L2 while this quad is consumed
            const int kn = kq + kq_stride;
            float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
            if (kn < kq_total) load_w(kn, n0, n1, n2, n3);
            const float* prow = patch + kq * 4;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float4 pv = *reinterpret_cast<const float4*>(prow + (size_t)m * p.kc_pad);
                acc[m] = fmaf(pv.x, w0, acc[m]);
                acc[m] = fmaf(pv.y, w1, acc[m]);
                acc[m] = fmaf(pv.z, w2, acc[m]);
                acc[m] = fmaf(pv.w, w3, acc[m]);
            }
            w0 = n0; w1 = n1; w2 = n2; w3 = n3;
            kq = kn;
        }
        if (RED_ALIASES_PATCH) __syncthreads();
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float v = acc[m];
            for (int off = cp; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (ks == 0) red[((size_t)warp * MT + m) * cp + cl] = v;
        }
        __syncthreads();
        for (int t = tid; t < MT * cp; t += NW * 32) {
            const int m = t / cp, c = t % cp;
            const int64_t o = tile_base + m;
            const int oc = cb + c;
            if (o < n_out && oc < p.cout) {
                float v = 0.0f;
#pragma unroll
                for (int w = 0; w < NW; ++w) v += red[((size_t)w * MT + m) * cp + c];
                if (p.normalize) {
                    const float nv = norm[m];
                    if (nv != 0.0f) v /= nv;
                }
                if (p.bias) v += __ldg(p.bias + oc);
                if (p.residual) v += __ldg(p.residual + o * p.residual_stride + oc);
                float* dst = p.out + o * p.out_stride + oc;
                if (p.accumulate) v += *dst;
                *dst = v;
            }
        }
        __syncthreads();
    }
}

// ---- phase 2, register-blocked variant used by k_cconv_wide ------------------------------------------------------------
// Patch tile stored k-quad major: word((k, m)) = ((k >> 2) * (MT + 1) + m) * 4 + (k & 3)  (the +1 keeps the phase-1
// stores conflict free), so the MT float4s of one k-quad sit at compile-time offsets: no address arithmetic in the
// inner loop.  Lane = (point group pg = lane / 8, channel quad cq = lane % 8): each thread owns MT/4 points x 4 output
// channels, filter rows arrive as LDG.128 (register double buffered), 16 FFMA per LDS.128.
// Needs cout % 4 == 0 and a 16-byte aligned filter.  Split-K over warps, partial sums meet in `red` ([NW][MT][32]).
template <int MT>
__device__ __forceinline__ int patchq_index(int m, int k) {
    return ((k >> 2) * (MT + 1) + m) * 4 + (k & 3);
}

template <int MT, int NW, bool RED_ALIASES_PATCH>
__device__ __forceinline__ void cconv_phase2_v2(const ConvParams& p, const float* patchq, float* red, const float* norm,
                                                int64_t tile_base, int64_t n_out) {
    static_assert(MT % 4 == 0, "MT must be a multiple of 4");
    constexpr int R = MT / 4, MTP = MT + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = lane >> 3, cq = lane & 7;
    const int kq_total = p.kc_pad / 4;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int cb = 0; cb < p.cout; cb += 32) {
        const int co0 = cb + cq * 4;
        const bool co_ok = co0 < p.cout;
        // accumulators as float2 pairs: Blackwell's packed FFMA2 (fma.rn.f32x2) does two FMAs per issued instruction
        float2 acc[R][2];
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i][0] = acc[i][1] = make_float2(0.0f, 0.0f);
        // filter rows as float4 (cout % 4 == 0, 16-byte aligned); whole k-quads below kq_full need no guards
        const int row4 = p.cout >> 2;                      // float4s per filter row
        const int kq_full = co_ok ? (p.kc >> 2) : 0;        // k-quads with all 4 rows present (0 disables the lane)
        const float4* wbase = reinterpret_cast<const float4*>(p.filters) + (co0 >> 2);
        const int dbg_mask = 0x7fffffff;
        auto load_w = [&](int kq, float4 (&w)[4]) {
            if (kq < kq_full) {
                const float4* wp = wbase + (size_t)(kq & dbg_mask) * 4 * row4;
                w[0] = __ldg(wp); w[1] = __ldg(wp + row4); w[2] = __ldg(wp + 2 * row4); w[3] = __ldg(wp + 3 * row4);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w[j] = (co_ok && kq * 4 + j < p.kc) ? __ldg(wbase + (size_t)(kq * 4 + j) * row4) : zero4;
            }
        };
        int kq = warp;
        float4 w[4] = {zero4, zero4, zero4, zero4};
        if (kq < kq_total) load_w(kq, w);
        while (kq < kq_total) {
            const int kn = kq + NW;
            float4 wn[4] = {zero4, zero4, zero4, zero4};
            if (kn < kq_total) load_w(kn, wn);
            const float4* pq = reinterpret_cast<const float4*>(patchq) + (size_t)kq * MTP + pg * R;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float4 pv = pq[i];
                const float2 px = make_float2(pv.x, pv.x), py = make_float2(pv.y, pv.y), pz = make_float2(pv.z, pv.z),
                             pw = make_float2(pv.w, pv.w);
                acc[i][0] = __ffma2_rn(px, make_float2(w[0].x, w[0].y), acc[i][0]);
                acc[i][1] = __ffma2_rn(px, make_float2(w[0].z, w[0].w), acc[i][1]);
                acc[i][0] = __ffma2_rn(py, make_float2(w[1].x, w[1].y), acc[i][0]);
                acc[i][1] = __ffma2_rn(py, make_float2(w[1].z, w[1].w), acc[i][1]);
                acc[i][0] = __ffma2_rn(pz, make_float2(w[2].x, w[2].y), acc[i][0]);
                acc[i][1] = __ffma2_rn(pz, make_float2(w[2].z, w[2].w), acc[i][1]);
                acc[i][0] = __ffma2_rn(pw, make_float2(w[3].x, w[3].y), acc[i][0]);
                acc[i][1] = __ffma2_rn(pw, make_float2(w[3].z, w[3].w), acc[i][1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = wn[j];
            kq = kn;
        }
        if (RED_ALIASES_PATCH) __syncthreads();
#pragma unroll
        for (int i = 0; i < R; ++i)
            *reinterpret_cast<float4*>(red + ((size_t)warp * MT + pg * R + i) * 32 + cq * 4) =
                make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
        __syncthreads();
        for (int t = tid; t < MT * 32; t += NW * 32) {
            const int m = t >> 5, c = t & 31;
            const int64_t o = tile_base + m;
            const int oc = cb + c;
            if (o < n_out && oc < p.cout) {
                float v = 0.0f;
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) v += red[((size_t)w2 * MT + m) * 32 + c];
                if (p.normalize) {
                    const float nv = norm[m];
                    if (nv != 0.0f) v /= nv;
                }
                if (p.bias) v += __ldg(p.bias + oc);
                if (p.residual) v += __ldg(p.residual + o * p.residual_stride + oc);
                float* dst = p.out + o * p.out_stride + oc;
                if (p.accumulate) v += *dst;
                *dst = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace dmcf
