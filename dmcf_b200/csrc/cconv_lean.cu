// continuous_conv forward, "lean" register-patch kernel: the production path for the wide layers of a DMCF step
// (compile-time filter grids 4x4x4 / 1x8x8 / 1x8x1, linear interpolation, cin <= 32, cout <= 32, cout % 4 == 0).
//
// Same algorithm as k_cconv_wide (cconv_wide.cu: one warp per out point, lane = input channel, the whole trilinear
// patch of the point in registers, then patch x filter with split-K over the warps), rebuilt around what the round-1
// measurements showed to bound that kernel (profiles/README.md):
//   phase 1 was LATENCY bound: the register queue of gathered feature rows was shifted every pair, and a MOV of a
//   register with a load in flight waits for the load, so every pair paid an L2 round trip (~500 cycles per pair and
//   warp).  Here
//     * the feature rows travel global -> shared memory with cp.async (LDGSTS) into a per-warp ring, four pairs ahead,
//       and no register is tied to a load in flight;
//     * NO indirect branch: the pairs of a chunk are ordered by the base cell of their corner block (done by
//       dmcf_cconv_prepare, or by a warp bitonic sort here when the geometry is evaluated in-kernel) and the walk is a
//       merge over the compile-time list of base cells -- `while (b == c) { 8 FFMA on static registers }`, cells that do
//       not occur in the chunk skipped by a uniform bit test.  (nvcc lowers a 27-way switch to a compare tree plus
//       three-entry jump tables: ~10 instructions and two dependent branches per pair.)  The order is verified when a
//       chunk is built (one shuffle + vote) and restored by the in-kernel sort if a caller's records are not ordered;
//     * feature rows are addressed with 32-bit byte offsets computed once per chunk; per-pair metadata
//       {offset of pair j+4, base cell of pair j+1} is one LDS.64; chunks are null padded: no per-pair bounds checks;
//     * the raw pair records of the NEXT chunk (also across the warp's points) are loaded before the walk.
//   phase 2 was bound by the shared-memory -> register path (128 B/clk/SM = one word per lane and cycle for the whole
//   SM): with 24 outputs per thread every FMA pair needed ~0.46 operand words.  Here
//     * thread tile 6 points x 16 output channels (96 accumulators), the four quarter-warps take the four k of a
//       k-quad: 22 operand words per 96 FMA (0.23), which puts the FFMA2 pipe and the operand path in balance;
//       the quarter-warp partial sums meet in a two-step shuffle reduce-scatter, the warps' in shared memory;
//     * the filter streams L2 -> shared memory through a per-warp cp.async ring, two k-quads ahead
//       (one LDGSTS per k-quad per warp, no registers held by the prefetch).
// Round 2: for cout == 32 phase 2 runs on the tensor cores through the warp-level path (mma.sync, 3xTF32 split in registers;
// template parameter TC below).  tcgen05 was measured and ruled out for this shape (dense_umma.cu).
#include <cuda_pipeline_primitives.h>

#include "cconv_walk.cuh"

namespace dmcf {


// SL = 1: the register-patch walk above.  SL = 4 / 8 (cin <= 8 / 4): the multi-pair phase 1 of cconv_walk.cuh
// (lean::point_patch_mp; relu / scale are run-time flags there, so those instances use RELU = FX = false).
// TC = true (round 2, cout == 32, cin % 8 == 0, dense_cin % 8 == 0): phase 2 on the tensor cores through the warp-level path
// (mma.sync m16n8k8 tf32, SASS HMMA.1688: 511 MAC per clock and SM measured, scripts/mma_sync_probe.cu, against 128 for FFMA2) as
// 3xTF32 -- operands are split into hi / lo IN REGISTERS after the fragment loads, so the split costs no shared memory (what
// rules tcgen05 out here): D[co][pt] += F_hi P_hi + F_lo P_hi + F_hi P_lo, float32 accumulators.  A warp
// owns k-OCTETS w, w + NW, ... and the whole 32 x 24 output tile: 2 x 3 MMA tiles, 14 conflict-free LDS (8 filter words from a
// swizzled 1 KB ring slot, 6 patch words) and 18 HMMA per octet instead of 2 x (28 wavefronts + 48 FFMA2).  The fused Dense rows
// are not tile columns in this mode (the tile holds kc_conv columns, which makes room for two octet slots per warp): their B
// fragments are read from the out points' own feature rows in global memory by the (at most one per warp) octet that needs them.
template <int KZ, int KY, int KX, int MT, int NW, bool RELU, bool FX, int SL, int TC = 0>  // TC = filter octet slots per warp, 0 = FFMA2
__global__ void __launch_bounds__(NW * 32, 1) k_cconv_lean(const ConvParams p) {
    using G = FilterGrid<KZ, KY, KX>;
    constexpr int K = G::K;
    constexpr int PPW = MT / NW;  // points per warp
    static_assert(MT % NW == 0 && MT % 8 == 0, "tile shape");
    constexpr int MTP = MT + 1;
    extern __shared__ __align__(1024) float smem[];
    // [NW gather rings of 512 B][patch tile, k-quad major [kc_pad/4][MT+1][4]; later [NW][MT][32]][NW record blocks][norm]
    // TC: [NW scratch blocks of 2 KB = gather ring + record block, later two filter octet slots][patch tile of kc_conv columns][norm]
    const int tile_cols = TC ? p.kc_conv : p.kc_pad;
    constexpr int TCS = TC > 0 ? TC : 2;              // filter octet slots per warp (tensor-core phase 2)
    constexpr int TCW = TCS * 256;                   // words of a warp's scratch block in that mode
    static_assert(!TC || TCW >= lean::kScratchWords, "the octet slots reuse the phase-1 scratch");
    float* rings = smem;
    float* patch = rings + (size_t)NW * (TC ? TCW : lean::kGatherSlots * 32);
    const size_t tile_words = (size_t)(tile_cols / 4) * MTP * 4, red_words = (size_t)NW * MT * 32;
    float* recs = patch + (tile_words > red_words ? tile_words : red_words);
    float* norm = TC ? recs : recs + (size_t)NW * lean::kRecWords;  // [MT]
    float* accs = norm + MT;                                         // SL > 1: [NW][K][32] per-warp slot patches

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_out = conv_n_out(p);
    const int64_t n_tiles = (n_out + MT - 1) / MT;  // capacity-sized launch: tiles beyond the device-side count do not exist
    float* wring = TC ? rings + (size_t)warp * TCW : rings + (size_t)warp * lean::kGatherSlots * 32;
    float* wrec = TC ? wring + lean::kGatherSlots * 32 : recs + (size_t)warp * lean::kRecWords;
    const bool lane_ci = lane < p.cin;
    lean::WarpCtx cx;
    cx.init(wring, wrec, p, lane);

    // The grid may be smaller than the number of tiles (persistent CTAs, tensor-core variant): a CTA then walks tiles blockIdx.x,
    // blockIdx.x + gridDim.x, ... and the first point of this warp in the NEXT tile is fetched during the current one -- its row
    // bounds and position at the start of phase 1, its first chunk of pair records at the end of phase 1 (in flight during phase
    // 2) -- so a tile starts with its operands at hand instead of two dependent round trips after a CTA launch.
    // (The FFMA2 variants are launched with one CTA per tile and carry nothing across tiles: their 96 accumulators leave no
    // registers for it -- measured: 7.08 -> 7.20 ms with the carried values spilled.)
    int64_t tile = blockIdx.x;
    int64_t t_rs = 0, t_re = 0;
    float t_ox = 0.f, t_oy = 0.f, t_oz = 0.f;
    PairRec t_cur;
    if constexpr (TC) {
        const int64_t o0 = tile * MT + warp;
        if (tile < n_tiles && o0 < n_out) {
            t_rs = p.row_splits[o0]; t_re = p.row_splits[o0 + 1];
            t_ox = __ldg(p.out_pos + 3 * o0); t_oy = __ldg(p.out_pos + 3 * o0 + 1); t_oz = __ldg(p.out_pos + 3 * o0 + 2);
        }
        t_cur = pair_record(p, t_rs + lane, t_rs + lane < t_re, t_ox, t_oy, t_oz);
    }
    if (tile >= n_tiles) return;
#pragma unroll 1
    do {  // one pass for the FFMA2 variants (compile-time false loop condition), a tile loop for the tensor-core variant
    const int64_t tile_base = tile * MT;

    // ================= phase 1: patch rows of this warp's points =================
    int64_t o = tile_base + warp;
    bool o_ok = o < n_out;
    int64_t rs = 0, re = 0;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    PairRec cur;
    bool t_ok = false;
    if constexpr (TC) {
        rs = t_rs; re = t_re; ox = t_ox; oy = t_oy; oz = t_oz;
        cur = t_cur;
        const int64_t tile_n = tile + gridDim.x;
        const int64_t o_t = tile_n * MT + warp;
        t_ok = tile_n < n_tiles && o_t < n_out;
        t_rs = 0; t_re = 0;
        if (t_ok) {
            t_rs = p.row_splits[o_t]; t_re = p.row_splits[o_t + 1];
            t_ox = __ldg(p.out_pos + 3 * o_t); t_oy = __ldg(p.out_pos + 3 * o_t + 1); t_oz = __ldg(p.out_pos + 3 * o_t + 2);
        }
    } else {
        if (o_ok) {
            rs = p.row_splits[o]; re = p.row_splits[o + 1];
            ox = __ldg(p.out_pos + 3 * o); oy = __ldg(p.out_pos + 3 * o + 1); oz = __ldg(p.out_pos + 3 * o + 2);
        }
        cur = pair_record(p, rs + lane, rs + lane < re, ox, oy, oz);
    }
    if constexpr (SL > 1) {
        float* a = accs + (size_t)warp * K * 32 + lane;
#pragma unroll 8
        for (int c = 0; c < K; ++c) a[c * 32] = 0.0f;
        __syncwarp();
    }

#pragma unroll 1
    for (int i = 0; i < PPW; ++i) {
        const int m = warp + i * NW;
        // the warp's next point
        const int64_t o_n = o + NW;
        const bool n_ok = (i + 1 < PPW) && o_n < n_out;
        int64_t rs_n = 0, re_n = 0;
        float ox_n = 0.f, oy_n = 0.f, oz_n = 0.f;
        if (n_ok) {
            rs_n = p.row_splits[o_n]; re_n = p.row_splits[o_n + 1];
            ox_n = __ldg(p.out_pos + 3 * o_n); oy_n = __ldg(p.out_pos + 3 * o_n + 1); oz_n = __ldg(p.out_pos + 3 * o_n + 2);
        }
        // first chunk of the next point: in flight during this whole point (a short last chunk would not cover it)
        const PairRec first_n = pair_record(p, rs_n + lane, n_ok && rs_n + lane < re_n, ox_n, oy_n, oz_n);
        float norm_acc;
        if constexpr (SL > 1) {
            // lanes = SL pair slots x 32/SL channels; the patch row is written by point_patch_mp
            const int ch = lane % (32 / SL);
            float fc = 0.0f;
            if (p.ascc && o_ok && ch < p.cin) {
                fc = __ldg(p.inp_feat + o * p.inp_stride + ch);
                if (p.relu_input) fc = fmaxf(fc, 0.0f);
                fc *= p.feat_scale;
            }
            norm_acc = lean::point_patch_mp<G, SL, MT>(p, lane, cur, rs, re, ox, oy, oz, fc, p.ascc || p.feat_scale != 1.0f,
                                                        accs + (size_t)warp * K * 32, patch, m);
        } else {
            float acc[K];
#pragma unroll
            for (int c = 0; c < K; ++c) acc[c] = 0.0f;
            float fc = 0.0f;  // centre feature of the antisymmetric layer (out point o == input row o)
            if (p.ascc && o_ok && lane_ci) {
                fc = __ldg(p.inp_feat + o * p.inp_stride + lane);
                if (RELU) fc = fmaxf(fc, 0.0f);
                fc *= p.feat_scale;
            }
            norm_acc = lean::point_patch<lean::FullPatch<KZ, KY, KX>, RELU, FX>(p, cx, cur, rs, re, ox, oy, oz, fc, acc);
            // ---- patch row -> shared memory (lane = channel: conflict free) ----
            if (o_ok && lane_ci) {
                if ((p.cin & 3) == 0) {  // k = c*cin + lane: the k-quad advances by cin/4 per cell -> one running pointer
                    float* pp = patch + patchq_index<MT>(m, lane);
                    // the shipped widths get compile-time offsets: 64 STS with immediates instead of 64 x (address + STS)
                    if (p.cin == 32) lean::store_patch_row<32 * MTP, K>(pp, acc);
                    else if (p.cin == 24) lean::store_patch_row<24 * MTP, K>(pp, acc);
                    else if (p.cin == 16) lean::store_patch_row<16 * MTP, K>(pp, acc);
                    else {
                        const int step = p.cin * MTP;
#pragma unroll
                        for (int c = 0; c < K; ++c) pp[c * step] = acc[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < K; ++c) patch[patchq_index<MT>(m, c * p.cin + lane)] = acc[c];
                }
            }
        }
        // ---- Dense columns, padding ----
        if (!o_ok) {
            for (int k = lane; k < tile_cols; k += 32) patch[patchq_index<MT>(m, k)] = 0.0f;
        } else if (TC) {
            if (p.normalize) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) norm_acc += __shfl_xor_sync(0xffffffffu, norm_acc, off);
                if (lane == 0) norm[m] = norm_acc;
            }
        } else {
            if (p.dense_cin > 0) {
                const float* drow = p.dense_inp + o * p.dense_stride;
                for (int ci = lane; ci < p.dense_cin; ci += 32) {
                    float f = __ldg(drow + ci);
                    if (p.relu_input) f = fmaxf(f, 0.0f);
                    patch[patchq_index<MT>(m, p.kc_conv + ci)] = f;
                }
            }
            for (int k = p.kc + lane; k < p.kc_pad; k += 32) patch[patchq_index<MT>(m, k)] = 0.0f;
            if (p.normalize) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) norm_acc += __shfl_xor_sync(0xffffffffu, norm_acc, off);
                if (lane == 0) norm[m] = norm_acc;
            }
        }
        o = o_n; o_ok = n_ok; rs = rs_n; re = re_n; ox = ox_n; oy = oy_n; oz = oz_n;
        cur = first_n;
    }
    if constexpr (TC) t_cur = pair_record(p, t_rs + lane, t_ok && t_rs + lane < t_re, t_ox, t_oy, t_oz);  // in flight during phase 2
    lean::cp_wait<0>();

    if constexpr (TC) {
        // ================= phase 2 on the tensor cores: D[co][pt] = sum_k F[k][co] P[pt][k], split-K over the warps =================
        static_assert(MT % 8 == 0 && SL == 1, "tensor-core phase 2: whole 8-point MMA tiles, register-patch phase 1");
        constexpr int NT = MT / 8;  // MMA tiles along the points
        const int g = lane >> 2, t = lane & 3;
        const int n_oct_conv = p.kc_conv >> 3, n_oct = (p.kc_conv + p.dense_cin) >> 3;
        const int n_it = n_oct > warp ? (n_oct - warp + NW - 1) / NW : 0;
        // lane L moves float4 L and L + 32 of an octet (8 filter rows x 32 channels, contiguous); row r is rotated by 8 (r % 4)
        // words so that the fragment loads below (lanes = 8 channels x 4 rows) hit 32 different banks
        uint32_t dst_off[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int f4 = lane + 32 * h, row = f4 >> 3, c4 = f4 & 7;
            dst_off[h] = (uint32_t)(row * 32 + ((c4 * 4 + 8 * (row & 3)) & 31)) * 4u;
        }
        const float* fsrc = p.filters + (size_t)warp * 256 + lane * 4;
        auto issue = [&](int it, float* sl) {
            const uint32_t base = (uint32_t)__cvta_generic_to_shared(sl);
            const int pred = it < n_it;
            asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q cp.async.cg.shared.global [%0], [%2], 16; "
                         "@q cp.async.cg.shared.global [%1], [%2 + 512], 16; }"
                         ::"r"(base + dst_off[0]), "r"(base + dst_off[1]), "l"(fsrc), "r"(pred));
            lean::cp_commit();
            fsrc += (size_t)NW * 256;
        };
        __syncwarp();  // this warp's phase-1 scratch is dead
#pragma unroll
        for (int i = 0; i < TCS; ++i) issue(i, wring + i * 256);
        // B fragments of this warp's Dense octet (at most one: dense_cin <= 8 NW): the out points' own feature rows
        float bd[NT][2];
        {
            const int first_d = n_oct_conv + ((warp - n_oct_conv % NW) + NW) % NW;  // first octet >= n_oct_conv of this warp
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                bd[n][0] = bd[n][1] = 0.0f;
                const int64_t oo = tile_base + 8 * n + g;
                if (first_d < n_oct && oo < n_out) {
                    const float* drow = p.dense_inp + oo * p.dense_stride + 8 * (first_d - n_oct_conv) + t;
                    bd[n][0] = __ldg(drow); bd[n][1] = __ldg(drow + 4);
                    if (p.relu_input) { bd[n][0] = fmaxf(bd[n][0], 0.0f); bd[n][1] = fmaxf(bd[n][1], 0.0f); }
                }
            }
        }
        __syncthreads();  // patch tile complete
        float c[2][NT][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) c[m][n][i] = 0.0f;
        // fragment word offsets inside a slot: rows t and t + 4 share the rotation 8 t
        int aoff[2][2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            aoff[m][0] = t * 32 + ((16 * m + g + 8 * t) & 31);
            aoff[m][1] = t * 32 + ((16 * m + 8 + g + 8 * t) & 31);
        }
        const float* pb = patch + (size_t)(2 * warp) * MTP * 4 + g * 4 + t;  // word t of point g of the octet's first k-quad
#pragma unroll 1
        for (int it = 0; it < n_it; ++it) {
            float* sl = wring + (it % TCS) * 256;
            lean::cp_wait<TCS - 1>();
            __syncwarp();  // every lane's part of this slot has landed
            uint32_t ah[2][4], al[2][4], bh[NT][2], bl[NT][2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float a[4] = {sl[aoff[m][0]], sl[aoff[m][1]], sl[aoff[m][0] + 128], sl[aoff[m][1] + 128]};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ah[m][i] = __float_as_uint(a[i]) & 0xffffe000u;
                    al[m][i] = __float_as_uint(a[i] - __uint_as_float(ah[m][i]));
                }
            }
            const bool dense_oct = warp + it * NW >= n_oct_conv;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                float b0, b1;
                if (dense_oct) {
                    b0 = bd[n][0]; b1 = bd[n][1];
                } else {
                    b0 = pb[n * 32]; b1 = pb[n * 32 + MTP * 4];
                }
                bh[n][0] = __float_as_uint(b0) & 0xffffe000u; bl[n][0] = __float_as_uint(b0 - __uint_as_float(bh[n][0]));
                bh[n][1] = __float_as_uint(b1) & 0xffffe000u; bl[n][1] = __float_as_uint(b1 - __uint_as_float(bh[n][1]));
            }
            __syncwarp();  // every lane has read this slot
            issue(it + TCS, sl);
            pb += (size_t)2 * NW * MTP * 4;
#define DMCF_MMA_TF32(C, A, B)                                                                                              \
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                 : "+f"(C[0]), "+f"(C[1]), "+f"(C[2]), "+f"(C[3])                                                           \
                 : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]))
            // three passes over the six accumulator tiles: consecutive MMAs on one accumulator are six instructions apart
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) DMCF_MMA_TF32(c[m][n], al[m], bh[n]);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) DMCF_MMA_TF32(c[m][n], ah[m], bl[n]);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) DMCF_MMA_TF32(c[m][n], ah[m], bh[n]);
#undef DMCF_MMA_TF32
        }
        lean::cp_wait<0>();
        __syncthreads();  // the patch is dead: partial sums reuse its storage
        float* red = patch;  // [NW][MT][32]
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                float* r = red + ((size_t)warp * MT + 8 * n + 2 * t) * 32 + 16 * m + g;
                r[0] = c[m][n][0]; r[32] = c[m][n][1]; r[8] = c[m][n][2]; r[40] = c[m][n][3];
            }
        __syncthreads();
        for (int tt = tid; tt < MT * 32; tt += NW * 32) {
            const int m = tt >> 5, cc2 = tt & 31;
            const int64_t oo = tile_base + m;
            if (oo < n_out && cc2 < p.cout) {
                float v = 0.0f;
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) v += red[((size_t)w2 * MT + m) * 32 + cc2];
                if (p.normalize) {
                    const float nv = norm[m];
                    if (nv != 0.0f) v /= nv;
                }
                if (p.bias) v += __ldg(p.bias + cc2);
                if (p.residual) v += __ldg(p.residual + oo * p.residual_stride + cc2);
                float* dst = p.out + oo * p.out_stride + cc2;
                if (p.accumulate) v += *dst;
                *dst = v;
            }
        }
        __syncthreads();  // partial sums read: the storage is the next tile's patch, the ring slots its scratch
    } else {
    // ================= phase 2: [MT x kc] x [kc x cout], split-K over the warps and over the quarter-warps =================
    // lane = (q = lane / 8: k = 4*kq + q, pr = (lane / 2) % 4: points 6*pr..6*pr+5, cc = lane % 2: channels 16*cc..16*cc+15)
    static_assert(MT == 24, "phase 2 thread tile assumes 24 points per CTA");
    const int kq_total = p.kc_pad / 4;
    const int cout = p.cout;
    // filter ring [kFilterSlots][4 rows][cout] in this warp's phase-1 scratch (its phase 1 is complete, nobody else
    // touches it): slot 0 = the gather ring, slots 1.. = the record block
    float* gring = rings + (size_t)warp * lean::kGatherSlots * 32;
    float* const slot0 = gring, * const slot1 = wrec, * const slot2 = wrec + 128;
    const int n_it = kq_total > warp ? (kq_total - warp + NW - 1) / NW : 0;
    // lane L moves the L-th float4 of a k-quad (4 consecutive filter rows = cout float4s).  Rows >= kc exist only in the
    // last k-quad; they are read from row kc-1 instead (their patch columns are zero).
    const bool cp_lane = lane < cout;
    const int cp_row = (lane * 4) / cout;
    const float* fsrc = p.filters + (size_t)warp * 4 * cout + lane * 4;  // this lane's float4 of the next k-quad to fetch
    const size_t fstep = (size_t)NW * 4 * cout;
    const int it_last = (kq_total - 1) / NW;
    const int last_fix = ((kq_total - 1) % NW == warp && (kq_total - 1) * 4 + cp_row > p.kc - 1)
                             ? ((kq_total - 1) * 4 + cp_row - (p.kc - 1)) * cout : 0;
    // (A per-warp TMA bulk copy per k-quad -- cp.async.bulk completing on an mbarrier -- was built and measured here: 4.28 ms
    // against 3.88 ms per 32->32 launch over 531 k points, profiles/README.md: with 512-byte slots and two k-quads of lookahead
    // the higher latency of the bulk path is not covered, and there is no shared memory left for deeper / larger stages.)
    auto issue = [&](int it, float* slot) {  // fetch k-quad warp + it*NW into `slot`
        const float* src = fsrc - (it == it_last ? last_fix : 0);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + lane * 4);
        const int pred = (it < n_it) && cp_lane;
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.cg.shared.global [%0], [%1], 16; }"
                     ::"r"(dst), "l"(src), "r"(pred));
        lean::cp_commit();
        fsrc += fstep;
    };
    __syncwarp();
    if (!p.patch_out) { issue(0, slot0); issue(1, slot1); }
    __syncthreads();  // patch tile complete
    if (p.patch_out) {
        // dmcf_cconv_patches: the patch tile goes to global memory (row o = [kc_conv] floats), consecutive threads take
        // consecutive k-quads of one point: conflict-free LDS.128, coalesced 16-byte stores
        const float4* p4 = reinterpret_cast<const float4*>(patch);
        const int kq_conv = (p.kc_conv + 3) / 4;
        const bool vec = (p.kc_conv & 3) == 0 && (p.patch_stride & 3) == 0 && ((uintptr_t)p.patch_out & 15) == 0;
        for (int idx = tid; idx < MT * kq_conv; idx += NW * 32) {
            const int m = idx / kq_conv, kq = idx - m * kq_conv;
            const int64_t oo = tile_base + m;
            if (oo >= n_out) continue;
            const float4 v = p4[(size_t)kq * MTP + m];
            float* dst = p.patch_out + oo * p.patch_stride + 4 * kq;
            if (vec) {
                *reinterpret_cast<float4*>(dst) = v;
            } else {
                const float e[4] = {v.x, v.y, v.z, v.w};
                for (int j = 0; j < 4 && 4 * kq + j < p.kc_conv; ++j) dst[j] = e[j];
            }
        }
        return;
    }
    if (p.debug_wrap_w & 2) {  // timing experiment: phase 1 only
        lean::cp_wait<0>();
        return;
    }

    const int q = lane >> 3, pr = (lane >> 1) & 3, cc = lane & 1;
    // this thread's four float4s of filter row q in each ring slot (channel offsets clamped: lanes beyond cout recompute
    // the last quad)
    const float4* fw0[4]; const float4* fw1[4]; const float4* fw2[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int off = q * cout + min(cc * 16 + 4 * h, cout - 4);
        fw0[h] = reinterpret_cast<const float4*>(slot0 + off);
        fw1[h] = reinterpret_cast<const float4*>(slot1 + off);
        fw2[h] = reinterpret_cast<const float4*>(slot2 + off);
    }
    float2 acc2[6][8];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int h = 0; h < 8; ++h) acc2[i][h] = make_float2(0.0f, 0.0f);
    const float* pw = patch + ((size_t)warp * MTP + pr * 6) * 4 + q;  // word q of point 6*pr of k-quad `warp`
    // one k-quad: operands of slot FR, refill of the slot consumed in the previous step (WS) with k-quad it+2
#define DMCF_LEAN_STEP(FR, WS)                                                                                   \
    {                                                                                                            \
        lean::cp_wait<1>();                                                                                      \
        __syncwarp(); /* every lane's part of this slot landed; every lane is done with the previous slot */    \
        float4 w[4];                                                                                             \
        _Pragma("unroll") for (int h = 0; h < 4; ++h) w[h] = *FR[h];                                              \
        float pv[6];                                                                                             \
        _Pragma("unroll") for (int i = 0; i < 6; ++i) pv[i] = pw[i * 4];                                          \
        issue(it + 2, WS);                                                                                       \
        pw += (size_t)NW * MTP * 4;                                                                              \
        _Pragma("unroll") for (int i = 0; i < 6; ++i) {                                                          \
            const float2 pp = make_float2(pv[i], pv[i]);                                                         \
            _Pragma("unroll") for (int h = 0; h < 4; ++h) {                                                      \
                acc2[i][2 * h] = __ffma2_rn(pp, make_float2(w[h].x, w[h].y), acc2[i][2 * h]);                    \
                acc2[i][2 * h + 1] = __ffma2_rn(pp, make_float2(w[h].z, w[h].w), acc2[i][2 * h + 1]);            \
            }                                                                                                    \
        }                                                                                                        \
        ++it;                                                                                                    \
    }
    {
        int it = 0;
#pragma unroll 1
        while (it + 3 <= n_it) {
            DMCF_LEAN_STEP(fw0, slot2)
            DMCF_LEAN_STEP(fw1, slot0)
            DMCF_LEAN_STEP(fw2, slot1)
        }
        if (it < n_it) DMCF_LEAN_STEP(fw0, slot2)
        if (it < n_it) DMCF_LEAN_STEP(fw1, slot0)
    }
#undef DMCF_LEAN_STEP
    lean::cp_wait<0>();
    // ---- quarter-warp partial sums: reduce-scatter over lanes ^16 (keeps 3 of the 6 points) and ^8 (8 of the 16 channels)
    float2 a1[3][8];
    {
        const bool hi = (q & 2) != 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const float2 keep = hi ? acc2[i + 3][h] : acc2[i][h];
                const float2 send = hi ? acc2[i][h] : acc2[i + 3][h];
                a1[i][h].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 16);
                a1[i][h].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 16);
            }
    }
    float2 a2[3][4];
    {
        const bool hi = (q & 1) != 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const float2 keep = hi ? a1[i][h + 4] : a1[i][h];
                const float2 send = hi ? a1[i][h] : a1[i][h + 4];
                a2[i][h].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 8);
                a2[i][h].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 8);
            }
    }
    __syncthreads();  // the patch is dead: partial sums reuse its storage
    float* red = patch;  // [NW][MT][32]
    {
        // this thread now owns points 6*pr + 3*(q>>1) + i, channels 16*cc + 8*(q&1) + 0..7
        const int m0 = pr * 6 + 3 * (q >> 1), c0 = cc * 16 + 8 * (q & 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float* r = red + ((size_t)warp * MT + m0 + i) * 32 + c0;
            if (c0 < cout) *reinterpret_cast<float4*>(r) = make_float4(a2[i][0].x, a2[i][0].y, a2[i][1].x, a2[i][1].y);
            if (c0 + 4 < cout) *reinterpret_cast<float4*>(r + 4) = make_float4(a2[i][2].x, a2[i][2].y, a2[i][3].x, a2[i][3].y);
        }
    }
    __syncthreads();
    for (int t = tid; t < MT * 32; t += NW * 32) {
        const int m = t >> 5, c = t & 31;
        const int64_t oo = tile_base + m;
        if (oo < n_out && c < cout) {
            float v = 0.0f;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) v += red[((size_t)w2 * MT + m) * 32 + c];
            if (p.normalize) {
                const float nv = norm[m];
                if (nv != 0.0f) v /= nv;
            }
            if (p.bias) v += __ldg(p.bias + c);
            if (p.residual) v += __ldg(p.residual + oo * p.residual_stride + c);
            float* dst = p.out + oo * p.out_stride + c;
            if (p.accumulate) v += *dst;
            *dst = v;
        }
    }
    }  // FFMA2 phase 2
    } while (TC && (tile += gridDim.x) < n_tiles);
}

static size_t lean_tc_smem_bytes(int mt, int nw, int kc_conv, int slots) {
    const size_t tile = (size_t)(kc_conv / 4) * (mt + 1) * 4, red = (size_t)nw * mt * 32;
    return ((tile > red ? tile : red) + (size_t)nw * slots * 256 + mt) * sizeof(float);
}

// Tensor-core phase 2 (mma.sync 3xTF32), two tile shapes: the FFMA2 kernel's 24 points x 12 warps (default), or 16 points x
// 16 warps at 128 registers (option bit 16; one point per warp in phase 1, four octet slots per warp -- the 24 accumulators of
// this phase 2 leave the registers for it).  Measured on the C4 32->32 layer: 6.52 ms against 6.70 ms (FFMA2 phase 2: 7.08 ms):
// sixteen warps do not make phase 1 faster, and the filter is streamed 1.5 x as often.
template <int KZ, int KY, int KX, int MT, int NW, int SLOTS>
static int launch_lean_tc(const ConvParams& p, cudaStream_t st) {
    const bool fx = p.ascc || p.feat_scale != 1.0f;
    void (*kerns_tc[2][2])(const ConvParams) = {
        {k_cconv_lean<KZ, KY, KX, MT, NW, false, false, 1, SLOTS>, k_cconv_lean<KZ, KY, KX, MT, NW, false, true, 1, SLOTS>},
        {k_cconv_lean<KZ, KY, KX, MT, NW, true, false, 1, SLOTS>, k_cconv_lean<KZ, KY, KX, MT, NW, true, true, 1, SLOTS>}};
    static bool tc_attr_set = false;
    if (!tc_attr_set) {
        for (int i = 0; i < 4; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns_tc[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_lean, tensor-core phase 2)");
        }
        tc_attr_set = true;
    }
    int64_t tiles = ceil_div(p.n_out, MT);
    if (!p.lean_cta_per_tile && tiles > 148) tiles = 148;  // persistent CTAs, one per SM (option bit 14: one CTA per tile)
    kerns_tc[p.relu_input ? 1 : 0][fx ? 1 : 0]<<<(unsigned)tiles, NW * 32, lean_tc_smem_bytes(MT, NW, p.kc_conv, SLOTS), st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_lean (tensor-core phase 2)");
    return DMCF_OK;
}

static size_t lean_smem_bytes(int mt, int nw, int kc_pad, int slot_patch_cells = 0) {
    const size_t tile = (size_t)(kc_pad / 4) * (mt + 1) * 4, red = (size_t)nw * mt * 32;
    return ((tile > red ? tile : red) + (size_t)nw * lean::kScratchWords + mt + (size_t)nw * slot_patch_cells * 32) * sizeof(float);
}

template <int KZ, int KY, int KX>
static int launch_lean_grid(const ConvParams& p, cudaStream_t st, bool* handled) {
    constexpr int MT = 24, NW = 12;
    *handled = false;
    if (lean_smem_bytes(MT, NW, p.kc_pad) > 227 * 1024) return DMCF_OK;
    static bool attr_set = false;
    // [relu on the input][feature scale and/or the antisymmetric centre term]
    void (*kerns[2][2])(const ConvParams) = {
        {k_cconv_lean<KZ, KY, KX, MT, NW, false, false, 1>, k_cconv_lean<KZ, KY, KX, MT, NW, false, true, 1>},
        {k_cconv_lean<KZ, KY, KX, MT, NW, true, false, 1>, k_cconv_lean<KZ, KY, KX, MT, NW, true, true, 1>}};
    // multi-pair phase 1 for narrow inputs: [0] cin <= 8 (4 pair slots), [1] cin <= 4 (8 pair slots).  Measured on the
    // Liquid3d cross-scale convs (26 M pairs): cin = 4: 1.95 -> 1.02 ms, cin = 8: 2.07 -> 1.59 ms; two slots (cin <= 16) were
    // slower than the single-pair walk (2.15 -> 2.60 ms: a step costs about two walk iterations) and are not instantiated.
    void (*kerns_mp[2])(const ConvParams) = {k_cconv_lean<KZ, KY, KX, MT, NW, false, false, 4>,
                                             k_cconv_lean<KZ, KY, KX, MT, NW, false, false, 8>};
    if (!attr_set) {
        for (int i = 0; i < 4; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_lean)");
        }
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns_mp[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_lean, multi-pair)");
        }
        attr_set = true;
    }
    *handled = true;
    const int64_t tiles = ceil_div(p.n_out, MT);
    constexpr int K = KZ * KY * KX;
    if (!p.no_multipair && p.cin <= 8 && lean_smem_bytes(MT, NW, p.kc_pad, K) <= 227 * 1024) {
        kerns_mp[p.cin <= 4 ? 1 : 0]<<<(unsigned)tiles, NW * 32, lean_smem_bytes(MT, NW, p.kc_pad, K), st>>>(p);
        DMCF_LAUNCH_CHECK("k_cconv_lean (multi-pair)");
        return DMCF_OK;
    }
    const bool fx = p.ascc || p.feat_scale != 1.0f;
    // tensor-core phase 2 (mma.sync 3xTF32): 32 output channels, whole k-octets, room for two octet slots per warp
    if (!p.no_lean_tc && !p.patch_out && p.cout == 32 && p.cin % 8 == 0 && p.dense_cin % 8 == 0 && p.dense_cin <= 8 * NW) {
        if (p.lean_tc_16 && lean_tc_smem_bytes(16, 16, p.kc_conv, 4) <= 227 * 1024) return launch_lean_tc<KZ, KY, KX, 16, 16, 4>(p, st);
        // two octet slots per warp (four, where the 154 KB tile of 24 input channels leaves room, measured slower: 5.66 vs 5.58 ms)
        if (lean_tc_smem_bytes(MT, NW, p.kc_conv, 2) <= 227 * 1024) return launch_lean_tc<KZ, KY, KX, MT, NW, 2>(p, st);
    }
    kerns[p.relu_input ? 1 : 0][fx ? 1 : 0]<<<(unsigned)tiles, NW * 32, lean_smem_bytes(MT, NW, p.kc_pad), st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_lean");
    return DMCF_OK;
}

// Tries the lean kernel; *handled = false means "not eligible" (the caller falls back to k_cconv_wide / k_cconv_tile).
int launch_cconv_lean(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p.gp.interp != DMCF_INTERP_LINEAR || p.cin > 32) return DMCF_OK;
    if (p.cout % 4 != 0 || p.cout > 32 || ((uintptr_t)p.filters & 15) != 0) return DMCF_OK;
    // feature rows are addressed with 32-bit byte offsets
    if ((p.n_inp > 0 ? p.n_inp : 1) * p.inp_stride * 4 >= ((int64_t)1 << 31)) return DMCF_OK;
    if (p.gp.kz == 4 && p.gp.ky == 4 && p.gp.kx == 4) return launch_lean_grid<4, 4, 4>(p, st, handled);
    if (p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 8) return launch_lean_grid<1, 8, 8>(p, st, handled);
    if (p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 1) return launch_lean_grid<1, 8, 1>(p, st, handled);
    return DMCF_OK;
}

}  // namespace dmcf
