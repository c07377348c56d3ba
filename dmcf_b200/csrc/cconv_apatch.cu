// continuous_conv forward for layers with <= 4 output channels whose (folded) patch fits the registers of a warp.
//
// First use -- the 2-D antisymmetric output layer of SymNet (utils/convolutions.py:242-254, 410-458;
// models/sym_net.py:39-67; configs/WBC-SPH.yml): 2 output channels, a 1x8x8 filter that the layer builds by mirroring and
// negating its stored half, F[rev(cell)] = -F[cell].
//
// k_cconv_direct evaluates every pair against the filter in shared memory: 8 corners x cout filter words per lane and
// pair, ~37 shared-memory wavefronts per pair -- it is bound by the shared-memory -> register path (128 B/clk/SM).
// The antisymmetry halves the patch instead:
//     out_o = sum_cell patch_o[cell] . F[cell] = sum_{cell < K/2} (patch_o[cell] - patch_o[K-1-cell]) . F[cell],
// and the folded half patch of a point (32 cells) lives in the REGISTERS of one warp (lane = input channel), so
//   phase 1  is the merge walk of k_cconv_lean (cconv_walk.cuh) with a scatter that adds corners of the upper half and
//            subtracts corners of the lower half: ~13 wavefronts per pair, no filter access at all;
//   phase 2  stays in the warp: each lane multiplies its folded patch values with its column of the resident half
//            filter (K/2 x cout conflict-free LDS per point instead of 8 x cout per PAIR), one shuffle reduction over
//            the channels per out point.  No patch tile in shared memory, no CTA barrier, persistent warps.
// Measured (400^2 particles, 10.5 pairs per point): 0.29 ms against 0.32 ms for k_cconv_direct.
// The 3-D layer (6x6x6, 125 base cells, 108 folded cells) was measured too and is NOT routed here: one walk body per base
// cell is 68 KB of code and thrashes the instruction cache (26 ms against 5.0 ms for k_cconv_direct); a single body with a
// per-pair 125-way switch is 8.0 ms (the compare tree nvcc builds costs several dependent branches per pair).
//
// Second use -- any 4x4x4 / 1x8x8 / 1x8x1 layer with <= 4 output channels (lean::FullPatch: the whole patch in registers,
// the same in-warp phase 2 against the whole resident filter): the 4-channel coarse scales of the multi-scale nets.
#include "cconv_walk.cuh"

namespace dmcf {

template <class S, int COUT, int NW, bool RELU>
__global__ void __launch_bounds__(NW * 32, 1) k_cconv_apatch(const ConvParams p) {
    constexpr int NACC = S::NACC;
    extern __shared__ __align__(1024) float smem[];
    // [NW gather rings of 512 B][NW record blocks][(half) filter [NACC][cin][COUT]][Dense kernel [dense_cin][COUT]]
    float* rings = smem;
    float* recs = rings + (size_t)NW * lean::kGatherSlots * 32;
    float* fh = recs + (size_t)NW * lean::kRecWords;
    const int half_words = NACC * p.cin * COUT;
    float* fd = fh + ((half_words + 3) & ~3);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint64_t filter_bar;
    for (int i = tid; i < p.dense_cin * COUT; i += NW * 32) fd[i] = __ldg(p.filters + (size_t)p.kc_conv * COUT + i);
    tma::stage_block(fh, p.filters, half_words, &filter_bar, NW * 32);  // the resident (half) filter: TMA bulk copies

    const bool lane_ci = lane < p.cin;
    lean::WarpCtx cx;
    cx.init(rings + (size_t)warp * lean::kGatherSlots * 32, recs + (size_t)warp * lean::kRecWords, p, lane);
    const float* fl = fh + (size_t)(lane_ci ? lane : 0) * COUT;  // this lane's column: + t * cin * COUT
    const int cell_stride = p.cin * COUT;

    const int64_t stride = (int64_t)gridDim.x * NW;
    const int64_t n_out = conv_n_out(p);
    int64_t o = (int64_t)blockIdx.x * NW + warp;
    bool o_ok = o < n_out;
    int64_t rs = 0, re = 0;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (o_ok) {
        rs = p.row_splits[o]; re = p.row_splits[o + 1];
        ox = __ldg(p.out_pos + 3 * o); oy = __ldg(p.out_pos + 3 * o + 1); oz = __ldg(p.out_pos + 3 * o + 2);
    }
    PairRec cur = pair_record(p, rs + lane, o_ok && rs + lane < re, ox, oy, oz);
#pragma unroll 1
    while (o_ok) {
        // the warp's next point; its first chunk of records is in flight during this whole point
        const int64_t o_n = o + stride;
        const bool n_ok = o_n < n_out;
        int64_t rs_n = 0, re_n = 0;
        float ox_n = 0.f, oy_n = 0.f, oz_n = 0.f;
        if (n_ok) {
            rs_n = p.row_splits[o_n]; re_n = p.row_splits[o_n + 1];
            ox_n = __ldg(p.out_pos + 3 * o_n); oy_n = __ldg(p.out_pos + 3 * o_n + 1); oz_n = __ldg(p.out_pos + 3 * o_n + 2);
        }
        const PairRec first_n = pair_record(p, rs_n + lane, n_ok && rs_n + lane < re_n, ox_n, oy_n, oz_n);
        float acc[NACC];
#pragma unroll
        for (int c = 0; c < NACC; ++c) acc[c] = 0.0f;
        float fc = 0.0f;  // centre feature of the antisymmetric layer (out point o == input row o)
        if (p.ascc && lane_ci) {
            fc = __ldg(p.inp_feat + o * p.inp_stride + lane);
            if (RELU) fc = fmaxf(fc, 0.0f);
            fc *= p.feat_scale;
        }
        lean::point_patch<S, RELU, true>(p, cx, cur, rs, re, ox, oy, oz, fc, acc);
        // ---- phase 2 in registers: folded patch x this lane's column of the half filter ----
        float out[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) out[co] = 0.0f;
        if (lane_ci) {
#pragma unroll
            for (int t = 0; t < NACC; ++t) {
                const float* fp = fl + t * cell_stride;
                if constexpr (COUT == 2) {
                    const float2 w = *reinterpret_cast<const float2*>(fp);
                    out[0] = fmaf(acc[t], w.x, out[0]); out[1] = fmaf(acc[t], w.y, out[1]);
                } else if constexpr (COUT == 4) {
                    const float4 w = *reinterpret_cast<const float4*>(fp);
                    out[0] = fmaf(acc[t], w.x, out[0]); out[1] = fmaf(acc[t], w.y, out[1]);
                    out[2] = fmaf(acc[t], w.z, out[2]); out[3] = fmaf(acc[t], w.w, out[3]);
                } else {
#pragma unroll
                    for (int co = 0; co < COUT; ++co) out[co] = fmaf(acc[t], fp[co], out[co]);
                }
            }
        }
        // fused Dense on the (relu'd, unscaled) centre features
        if (p.dense_cin > 0) {
            for (int ci = lane; ci < p.dense_cin; ci += 32) {
                float f = __ldg(p.dense_inp + o * p.dense_stride + ci);
                if (p.relu_input) f = fmaxf(f, 0.0f);
#pragma unroll
                for (int co = 0; co < COUT; ++co) out[co] = fmaf(f, fd[ci * COUT + co], out[co]);
            }
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float v = out[co];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            out[co] = v;
        }
        if (lane < COUT) {
            float v = out[0];
#pragma unroll
            for (int co = 1; co < COUT; ++co)
                if (lane == co) v = out[co];
            if (p.bias) v += __ldg(p.bias + lane);
            if (p.residual) v += __ldg(p.residual + o * p.residual_stride + lane);
            float* dst = p.out + o * p.out_stride + lane;
            if (p.accumulate) v += *dst;
            *dst = v;
        }
        o = o_n; o_ok = n_ok; rs = rs_n; re = re_n; ox = ox_n; oy = oy_n; oz = oz_n;
        cur = first_n;
    }
    lean::cp_wait<0>();
}

template <class S, int COUT>
static int launch_apatch(const ConvParams& p, cudaStream_t st, bool* handled) {
    constexpr int NW = S::NACC > 32 ? 12 : 16;
    constexpr int NACC = S::NACC;
    const size_t words = (size_t)NW * lean::kScratchWords + (((size_t)NACC * p.cin * COUT + 3) & ~(size_t)3) + (size_t)p.dense_cin * COUT;
    *handled = false;
    if (words * sizeof(float) > 200 * 1024) return DMCF_OK;
    static bool attr_set = false;
    void (*kerns[2])(const ConvParams) = {k_cconv_apatch<S, COUT, NW, false>, k_cconv_apatch<S, COUT, NW, true>};
    if (!attr_set) {
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_apatch)");
        }
        attr_set = true;
    }
    *handled = true;
    int64_t blocks = ceil_div(p.n_out, NW);
    if (blocks > 148) blocks = 148;  // persistent: one CTA per SM, warps stride over the out points
    kerns[p.relu_input ? 1 : 0]<<<(unsigned)blocks, NW * 32, words * sizeof(float), st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_apatch");
    return DMCF_OK;
}

template <int KZ, int KY, int KX>
static int launch_rpatch_grid(const ConvParams& p, cudaStream_t st, bool* handled) {
    using S = lean::FullPatch<KZ, KY, KX>;
    switch (p.cout) {
        case 1: return launch_apatch<S, 1>(p, st, handled);
        case 2: return launch_apatch<S, 2>(p, st, handled);
        case 3: return launch_apatch<S, 3>(p, st, handled);
        case 4: return launch_apatch<S, 4>(p, st, handled);
        default: return DMCF_OK;
    }
}

// Tries the register-patch kernels for layers with <= 4 output channels; *handled = false means "not eligible".
//   antisymmetric 1x8x8 filter (descriptor promise) -> folded half patch;
//   4x4x4 / 1x8x8 / 1x8x1 filters                   -> full patch in registers (the multi-scale nets' 4-channel scales:
//   their fine -> coarse convs see ~1600 pairs per out point, where the walk's 28 instructions per pair beat
//   k_cconv_direct's ~110: 4.6 -> 2.0 ms for the 24->4 conv of a Liquid3d step).
int launch_cconv_apatch(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p.normalize || p.gp.interp != DMCF_INTERP_LINEAR || p.cin > 32 || p.cout > 4) return DMCF_OK;
    if ((p.n_inp > 0 ? p.n_inp : 1) * p.inp_stride * 4 >= ((int64_t)1 << 31)) return DMCF_OK;  // 32-bit gather offsets
    const bool g444 = p.gp.kz == 4 && p.gp.ky == 4 && p.gp.kx == 4, g188 = p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 8;
    const bool g181 = p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 1;
    if (p.filter_antisym && g188 && p.cout == 2) return launch_apatch<lean::AntiPatch<1, 8, 8>, 2>(p, st, handled);
    if (g444) return launch_rpatch_grid<4, 4, 4>(p, st, handled);
    if (g188) return launch_rpatch_grid<1, 8, 8>(p, st, handled);
    if (g181) return launch_rpatch_grid<1, 8, 1>(p, st, handled);
    return DMCF_OK;
}

}  // namespace dmcf
