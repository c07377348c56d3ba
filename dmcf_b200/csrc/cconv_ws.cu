// continuous_conv forward, WARP-SPECIALISED register-patch kernel (round 2): the two phases of k_cconv_lean run CONCURRENTLY.
//
// STATUS: a measured experiment, NOT the production path (dmcf_set_kernel_options bit 7 enables it; the parity suite runs it).
// Measured on the C4 layers (1.06 M out points, profiles/README.md): 32->32 14.5 ms against 7.08 ms for k_cconv_lean, 24->32
// 11.4 ms against 6.24 ms.  Why, from the two timing switches (option bits 8 / 11): with the consumers' k loop switched off the
// producers alone need 5.0 ms (11 working warps at 112 registers are slower per point than lean's 12 at 168: 7.9 us against 5.2 us);
// with the producers switched off the consumers alone need 13.4 ms, and that time follows the DEPTH of their filter ring (5 slots
// 6.9 ms, 4 slots 10.3 ms on the 24->32 layer), not the arithmetic: streaming the 266 KB filter once per 11 points needs
// ~53 GB/s per SM, i.e. ~35 KB in flight at L2 latency, and next to two 90 KB half tiles there are 18 KB left for the rings
// (k_cconv_lean amortises every filter row over 24 points and spreads the stream over 12 warps).  Overlapping the phases costs
// more in filter streaming than it wins; cluster multicast of the filter would halve the stream but needs the CTAs of a pair in
// lock step.  Kept as evidence and as a starting point.
//
// k_cconv_lean (cconv_lean.cu) owns the SM with one 24-point patch tile and moves all 12 warps through phase 1 (pair walk,
// latency bound: 51 % issue utilisation, no FFMA2) and then phase 2 (patch x filter, FFMA2 / shared-memory operand bound, no
// L2 traffic) in lock step: neither phase can use what the other leaves idle (profiles/r1g_cconv_lean_source.md: 3.1 ms
// each for the 32->32 layer).  Here a persistent CTA of 16 warps splits the roles:
//   * 12 PRODUCER warps (3 warpgroups, `setmaxnreg.dec` to 112 registers): phase 1 exactly as in k_cconv_lean -- one warp per
//     out point, lane = input channel, the point's whole trilinear patch in 64 registers, cp.async gather ring, merge walk over
//     the base cells (cconv_walk.cuh) -- writing the finished patch rows into one of TWO half tiles of 11 points;
//   * 4 CONSUMER warps (one warpgroup, `setmaxnreg.inc` to 168): phase 2 on the other half tile: split-K over the four warps
//     and over the eight k of a step inside a warp (lane = k x point group x channel half), thread tile 6 points x 16 channels
//     (96 accumulators, 22 operand words per 96 FMA), filter rows streamed L2 -> shared memory through a per-warp cp.async ring,
//     packed FFMA2; then the cross-lane / cross-warp reduction and the epilogue (normalise, bias, fused Dense, residual, store).
//   Half tiles are handed over with mbarriers (full: 12 producer arrivals, empty: 4 consumer arrivals); the producers' gathers and
//   the consumers' FFMA2 stream overlap, so a layer costs max(phase 1, phase 2) plus the hand-over instead of their sum.
// Why 11 points: a half tile is stored k-quad major, [(k/4)][point][k%4], so that consecutive k-quads of a lane's column are
// 11 float4 = 44 words apart -- 12 (mod 32) banks, which makes the producers' 64 patch stores per point conflict free WITHOUT
// the padding column the 24-point tile needs; two half tiles of the 32-channel layer are 2 x 90 KB, which leaves room for the
// producers' scratch (12 x 2 KB), the consumers' filter rings (4 x 3 KB) and their partial sums.  The twelfth producer warp has
// no point (it only keeps the barrier counts; warpgroups are the granularity of setmaxnreg).  The filter is streamed once per
// 11 points instead of once per 24: 2.2x the L2 -> shared-memory traffic of k_cconv_lean (L2 runs at 8.5 % of its peak there).
// The fused Dense rows ride in the tile as one more "cell" like in k_cconv_lean (a first version added x_o . Wd in the epilogue
// with loads from L2 -- the SM has no L1 left next to 223 KB of shared memory -- and cost 15 us per tile).
// Same arithmetic per point as k_cconv_lean except for the summation order of the split-K partial sums (float32 rounding).
#include <cuda_pipeline_primitives.h>

#include "cconv_walk.cuh"

namespace dmcf {

namespace ws {
static constexpr int MT = 11;    // points per half tile
static constexpr int NPW = 12;   // producer warps (warp MT..NPW-1 have no point)
static constexpr int NCW = 4;    // consumer warps
static constexpr int kRowWords = 36;       // ring slot row: cout <= 32 words + 4 of padding -- with a 32-word stride the eight rows a
                                           // warp reads at once (one per k lane) sit in the same banks: 8-way conflicts on every LDS.128
static constexpr int kSlotWords = 8 * kRowWords;  // filter ring slot: 8 rows
static constexpr int kSlots = 4;
static constexpr int kRedRows = 12;        // partial-sum rows per consumer warp (row 11 is the padding point of the thread tile)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// word index of (point m, patch column k) in a half tile
__device__ __forceinline__ int tile_index(int m, int k) { return (((k >> 2) * MT) + m) * 4 + (k & 3); }
}  // namespace ws

template <int KZ, int KY, int KX, bool RELU, bool FX>
__global__ void __launch_bounds__((ws::NPW + ws::NCW) * 32, 1) k_cconv_ws(const ConvParams p) {
    using G = FilterGrid<KZ, KY, KX>;
    constexpr int K = G::K;
    constexpr int MT = ws::MT, NPW = ws::NPW, NCW = ws::NCW;
    extern __shared__ __align__(1024) float smem[];
    // [NPW gather rings of 512 B][NPW record blocks][2 half tiles (+ 4 words)][NCW filter rings][NCW partial sums][norm][barriers]
    float* rings = smem;
    float* recs = rings + (size_t)NPW * lean::kGatherSlots * 32;
    float* tiles = recs + (size_t)NPW * lean::kRecWords;
    const int kq_total = p.kc >> 2;                            // conv rows + fused Dense rows; kc % 8 == 0 (launcher)
    const size_t tile_words = (size_t)kq_total * MT * 4 + 4;   // + one float4: the thread tile's padding point reads past the last row
    float* frings = tiles + 2 * tile_words;
    float* red = frings + (size_t)NCW * ws::kSlots * ws::kSlotWords;
    float* norm = red + (size_t)NCW * ws::kRedRows * 32;       // [2][12]
    uint64_t* bars = reinterpret_cast<uint64_t*>(norm + 2 * 12);  // full[2], empty[2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_out = conv_n_out(p);
    const int64_t n_tiles = (n_out + MT - 1) / MT;
    const uint32_t bar_full0 = tma::smem_u32(bars), bar_empty0 = tma::smem_u32(bars + 2);
    if (tid == 0) {
        tma::mbar_init(bar_full0, NPW); tma::mbar_init(bar_full0 + 8, NPW);
        tma::mbar_init(bar_empty0, NCW); tma::mbar_init(bar_empty0 + 8, NCW);
        tma::mbar_fence_init();
    }
    __syncthreads();

    if (warp < NPW) {
        // =============================== producers: phase 1 ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        float* wrec = recs + (size_t)warp * lean::kRecWords;
        const bool lane_ci = lane < p.cin;
        const bool has_point = warp < MT;
        lean::WarpCtx cx;
        cx.init(rings + (size_t)warp * lean::kGatherSlots * 32, wrec, p, lane);
        int64_t t = blockIdx.x;
        int64_t o = t * MT + warp;
        bool o_ok = has_point && t < n_tiles && o < n_out;
        int64_t rs = 0, re = 0;
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (o_ok) {
            rs = p.row_splits[o]; re = p.row_splits[o + 1];
            ox = __ldg(p.out_pos + 3 * o); oy = __ldg(p.out_pos + 3 * o + 1); oz = __ldg(p.out_pos + 3 * o + 2);
        }
        PairRec cur = pair_record(p, rs + lane, rs + lane < re, ox, oy, oz);
#pragma unroll 1
        for (int i = 0; t < n_tiles; t += gridDim.x, ++i) {
            const int b = i & 1;
            // this warp's point of the CTA's next tile: its first chunk of records is in flight during this whole point
            const int64_t t_n = t + gridDim.x;
            const int64_t o_n = t_n * MT + warp;
            const bool n_ok = has_point && t_n < n_tiles && o_n < n_out;
            int64_t rs_n = 0, re_n = 0;
            float ox_n = 0.f, oy_n = 0.f, oz_n = 0.f;
            if (n_ok) {
                rs_n = p.row_splits[o_n]; re_n = p.row_splits[o_n + 1];
                ox_n = __ldg(p.out_pos + 3 * o_n); oy_n = __ldg(p.out_pos + 3 * o_n + 1); oz_n = __ldg(p.out_pos + 3 * o_n + 2);
            }
            const PairRec first_n = pair_record(p, rs_n + lane, n_ok && rs_n + lane < re_n, ox_n, oy_n, oz_n);
            float acc[K];
#pragma unroll
            for (int c = 0; c < K; ++c) acc[c] = 0.0f;
            float norm_acc = 0.0f;
            if (o_ok && !(p.debug_wrap_w & 8)) {  // (timing experiment, option bit 11: no phase 1 -> consumer-bound time)
                float fc = 0.0f;  // centre feature of the antisymmetric layer (out point o == input row o)
                if (p.ascc && lane_ci) {
                    fc = __ldg(p.inp_feat + o * p.inp_stride + lane);
                    if (RELU) fc = fmaxf(fc, 0.0f);
                    fc *= p.feat_scale;
                }
                norm_acc = lean::point_patch<lean::FullPatch<KZ, KY, KX>, RELU, FX>(p, cx, cur, rs, re, ox, oy, oz, fc, acc);
            }
            // ---- the half tile must have been drained by the consumers (its previous use) before the row goes in ----
            tma::mbar_wait(b ? bar_empty0 + 8 : bar_empty0, (uint32_t)(((i >> 1) & 1) ^ 1));
            if (o_ok) {
                float* tile = tiles + (size_t)b * tile_words;
                if (lane_ci) {
                    if ((p.cin & 3) == 0) {  // k = c*cin + lane: the k-quad advances by cin/4 per cell -> one running pointer
                        float* pp = tile + ws::tile_index(warp, lane);
                        const int step = p.cin * MT;
#pragma unroll
                        for (int c = 0; c < K; ++c) pp[c * step] = acc[c];
                    } else {
#pragma unroll
                        for (int c = 0; c < K; ++c) tile[ws::tile_index(warp, c * p.cin + lane)] = acc[c];
                    }
                }
                if (p.dense_cin > 0) {  // fused Dense: the (relu'd, unscaled) centre features are one more "cell" of the patch
                    const float* drow = p.dense_inp + o * p.dense_stride;
                    for (int ci = lane; ci < p.dense_cin; ci += 32) {
                        float f = __ldg(drow + ci);
                        if (p.relu_input) f = fmaxf(f, 0.0f);
                        tile[ws::tile_index(warp, p.kc_conv + ci)] = f;
                    }
                }
                if (p.normalize) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) norm_acc += __shfl_xor_sync(0xffffffffu, norm_acc, off);
                    if (lane == 0) norm[b * 12 + warp] = norm_acc;
                }
            }
            __syncwarp();
            if (lane == 0) ws::mbar_arrive(b ? bar_full0 + 8 : bar_full0);
            o = o_n; o_ok = n_ok; rs = rs_n; re = re_n; ox = ox_n; oy = oy_n; oz = oz_n;
            cur = first_n;
        }
        lean::cp_wait<0>();
        return;
    }

    // =============================== consumers: phase 2 + epilogue ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int wc = warp - NPW;
    const int ctid = tid - NPW * 32;
    const int cout = p.cout;
    const int groups = p.kc >> 3;                               // 8-row steps over the whole filter (conv + Dense rows)
    const int n_it = groups > wc ? (groups - wc + NCW - 1) / NCW : 0;
    float* ring = frings + (size_t)wc * ws::kSlots * ws::kSlotWords;
    const int f4_per_step = 2 * cout, c4 = cout >> 2;           // float4s of 8 filter rows / of one row
    // lane = (q = lane / 4: k = 8 g + q, pr = (lane / 2) % 2: points 6 pr .. 6 pr + 5, cc = lane % 2: channels 16 cc .. 16 cc + 15)
    const int q = lane >> 2, pr = (lane >> 1) & 1, cc = lane & 1;
    int fw_off[4];  // this lane's four float4s of filter row q inside a ring slot (lanes beyond cout recompute the last quad)
#pragma unroll
    for (int h = 0; h < 4; ++h) fw_off[h] = q * ws::kRowWords + min(cc * 16 + 4 * h, cout - 4);
    // this lane's patch words: k-quad 2 g + q / 4, word q % 4, points 6 pr + i
    const int pw_off = (((2 * wc + (q >> 2)) * MT) + pr * 6) * 4 + (q & 3);
    constexpr int pw_step = 2 * NCW * MT * 4;                   // g += NCW  ->  k-quad += 2 NCW
#pragma unroll 1
    for (int64_t t = blockIdx.x, i = 0; t < n_tiles; t += gridDim.x, ++i) {
        const int b = (int)(i & 1);
        const float* fsrc = p.filters + (size_t)wc * 8 * cout;  // this warp's next 8 rows to fetch
        const size_t fstep = (size_t)NCW * 8 * cout;
        auto issue = [&](int it, float* slot) {
            if (it < n_it) {
                for (int f = lane; f < f4_per_step; f += 32) {  // float4 f of the 8 x cout block: row f / (cout / 4), quad f % (cout / 4)
                    const int row = f / c4, col = f - row * c4;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + row * ws::kRowWords + col * 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(fsrc + f * 4));
                }
            }
            lean::cp_commit();
            fsrc += fstep;
        };
        int fetch_slot = 0;
#pragma unroll
        for (int pre = 0; pre < ws::kSlots - 1; ++pre) {
            issue(pre, ring + fetch_slot * ws::kSlotWords);
            ++fetch_slot;
        }
        tma::mbar_wait(b ? bar_full0 + 8 : bar_full0, (uint32_t)((i >> 1) & 1));
        float2 acc2[6][8];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int h = 0; h < 8; ++h) acc2[a][h] = make_float2(0.0f, 0.0f);
        const float* pw = tiles + (size_t)b * tile_words + pw_off;
        int use_slot = 0;  // fetch_slot == kSlots - 1: the slot consumed in the previous step
        const int n_run = (p.debug_wrap_w & 1) ? 0 : n_it;  // (timing experiment, option bit 8: no phase 2 -> producer-bound time)
#pragma unroll 1
        for (int it = 0; it < n_run; ++it) {
            lean::cp_wait<ws::kSlots - 2>();
            __syncwarp();  // every lane's part of this slot landed; every lane is done with the slot consumed one step ago
            const float* sl = ring + use_slot * ws::kSlotWords;
            float4 w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) w[h] = *reinterpret_cast<const float4*>(sl + fw_off[h]);
            float pv[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) pv[a] = pw[a * 4];
            issue(it + ws::kSlots - 1, ring + fetch_slot * ws::kSlotWords);
            fetch_slot = fetch_slot + 1 == ws::kSlots ? 0 : fetch_slot + 1;
            use_slot = use_slot + 1 == ws::kSlots ? 0 : use_slot + 1;
            pw += pw_step;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const float2 pp = make_float2(pv[a], pv[a]);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    acc2[a][2 * h] = __ffma2_rn(pp, make_float2(w[h].x, w[h].y), acc2[a][2 * h]);
                    acc2[a][2 * h + 1] = __ffma2_rn(pp, make_float2(w[h].z, w[h].w), acc2[a][2 * h + 1]);
                }
            }
        }
        lean::cp_wait<0>();
        // the half tile is not read any more (its normalisers are taken along in registers): hand it back to the producers now,
        // they refill it while this tile's partial sums are reduced and stored
        float nvals[3] = {0.f, 0.f, 0.f};
        if (p.normalize) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int idx = ctid + j * NCW * 32;
                if (idx < MT * 32) nvals[j] = norm[b * 12 + (idx >> 5)];
            }
        }
        __syncwarp();
        if (lane == 0) ws::mbar_arrive(b ? bar_empty0 + 8 : bar_empty0);
        // ---- partial sums over the eight k lanes: reduce-scatter over lane bits 4 (3 of the 6 points), 3 (8 of the 16
        //      channels) and 2 (4 of those 8) ----
        float2 a1[3][8];
        {
            const bool hi = (q & 4) != 0;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float2 keep = hi ? acc2[a + 3][h] : acc2[a][h];
                    const float2 send = hi ? acc2[a][h] : acc2[a + 3][h];
                    a1[a][h].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 16);
                    a1[a][h].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 16);
                }
        }
        float2 a2[3][4];
        {
            const bool hi = (q & 2) != 0;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float2 keep = hi ? a1[a][h + 4] : a1[a][h];
                    const float2 send = hi ? a1[a][h] : a1[a][h + 4];
                    a2[a][h].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 8);
                    a2[a][h].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 8);
                }
        }
        float2 a3[3][2];
        {
            const bool hi = (q & 1) != 0;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 keep = hi ? a2[a][h + 2] : a2[a][h];
                    const float2 send = hi ? a2[a][h] : a2[a][h + 2];
                    a3[a][h].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 4);
                    a3[a][h].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 4);
                }
        }
        {
            // this thread owns points 6 pr + 3 (q / 4) + a, channels 16 cc + 8 ((q / 2) % 2) + 4 (q % 2) + 0..3
            const int m0 = pr * 6 + 3 * (q >> 2), c0 = cc * 16 + 8 * ((q >> 1) & 1) + 4 * (q & 1);
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (c0 < cout)
                    *reinterpret_cast<float4*>(red + ((size_t)wc * ws::kRedRows + m0 + a) * 32 + c0) =
                        make_float4(a3[a][0].x, a3[a][0].y, a3[a][1].x, a3[a][1].y);
        }
        ws::consumer_sync();  // partial sums of the four warps are in place
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int idx = ctid + j * NCW * 32;
            if (idx >= MT * 32) break;
            const int m = idx >> 5, c = idx & 31;
            const int64_t oo = t * MT + m;
            if (oo < n_out && c < cout) {
                float v = 0.0f;
#pragma unroll
                for (int w2 = 0; w2 < NCW; ++w2) v += red[((size_t)w2 * ws::kRedRows + m) * 32 + c];
                if (p.normalize && nvals[j] != 0.0f) v /= nvals[j];
                if (p.bias) v += __ldg(p.bias + c);
                if (p.residual) v += __ldg(p.residual + oo * p.residual_stride + c);
                float* dst = p.out + oo * p.out_stride + c;
                if (p.accumulate) v += *dst;
                *dst = v;
            }
        }
        ws::consumer_sync();  // `red` has been consumed
    }
}

static size_t ws_smem_bytes(int kc) {
    const size_t tile_words = (size_t)(kc / 4) * ws::MT * 4 + 4;
    const size_t words = (size_t)ws::NPW * lean::kScratchWords + 2 * tile_words + (size_t)ws::NCW * ws::kSlots * ws::kSlotWords +
                         (size_t)ws::NCW * ws::kRedRows * 32 + 2 * 12;
    return words * sizeof(float) + 4 * sizeof(uint64_t);
}

template <int KZ, int KY, int KX>
static int launch_ws_grid(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    const size_t smem = ws_smem_bytes(p.kc);
    if (smem > 227 * 1024) return DMCF_OK;
    static bool attr_set = false;
    // [relu on the input][feature scale and/or the antisymmetric centre term]
    void (*kerns[2][2])(const ConvParams) = {{k_cconv_ws<KZ, KY, KX, false, false>, k_cconv_ws<KZ, KY, KX, false, true>},
                                             {k_cconv_ws<KZ, KY, KX, true, false>, k_cconv_ws<KZ, KY, KX, true, true>}};
    if (!attr_set) {
        for (int i = 0; i < 4; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_ws)");
        }
        attr_set = true;
    }
    *handled = true;
    const int64_t tiles = ceil_div(p.n_out, ws::MT);
    const unsigned blocks = (unsigned)(tiles < 148 ? tiles : 148);  // persistent: one CTA per SM, tiles dealt round robin
    const bool fx = p.ascc || p.feat_scale != 1.0f;
    kerns[p.relu_input ? 1 : 0][fx ? 1 : 0]<<<blocks, (ws::NPW + ws::NCW) * 32, smem, st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_ws");
    return DMCF_OK;
}

// Tries the warp-specialised kernel; *handled = false means "not eligible" (the caller falls back to k_cconv_lean).
int launch_cconv_ws(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p.gp.interp != DMCF_INTERP_LINEAR || p.cin > 32 || p.cin <= 8) return DMCF_OK;  // narrow inputs: multi-pair phase 1 of lean
    if (p.cout % 4 != 0 || p.cout > 32 || ((uintptr_t)p.filters & 15) != 0 || p.patch_out) return DMCF_OK;
    if ((p.kc & 7) != 0 || (p.kc >> 3) < ws::NCW) return DMCF_OK;  // 8-row steps over conv + fused Dense rows
    if ((p.n_inp > 0 ? p.n_inp : 1) * p.inp_stride * 4 >= ((int64_t)1 << 31)) return DMCF_OK;  // 32-bit gather offsets
    if (p.gp.kz == 4 && p.gp.ky == 4 && p.gp.kx == 4) return launch_ws_grid<4, 4, 4>(p, st, handled);
    if (p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 8) return launch_ws_grid<1, 8, 8>(p, st, handled);
    if (p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 1) return launch_ws_grid<1, 8, 1>(p, st, handled);
    return DMCF_OK;
}

}  // namespace dmcf
