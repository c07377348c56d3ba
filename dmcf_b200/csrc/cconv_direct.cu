// continuous_conv forward, direct variant for layers with very few output channels (cout <= 4): the antisymmetric
// output layer of SymNet (32 -> 2/3, 6x6x6 or 1x8x8 filter) and the 1-channel heads.
//
// With so few outputs the patch x filter product of the other kernels is the wrong shape (a 216-cell x 32-channel
// patch per point, ~28 KB, to produce 3 numbers).  Instead the whole filter (<= ~100 KB) stays RESIDENT in shared
// memory for the lifetime of a persistent CTA and every pair is evaluated directly:
//     out[o][co] += sum_corners w_c * g(f_n)[ci] * F[cell_c][ci][co],      lane = input channel,
// three conflict-free LDS + three FFMA per corner, no read-modify-write, no second phase -- one warp-shuffle
// reduction over the channels per out point.  Pair geometry is evaluated lane-parallel for 32 neighbours and parked
// as records in a per-warp scratch exactly like k_cconv_wide.
#include "cconv_common.cuh"

namespace dmcf {

static constexpr int kDirWarps = 16;
static constexpr int kDirRecWords = 12;  // {row, base offset, dx | dy << 16, dz (filter words), w0..w3, w4..w7}

template <int COUT>
__global__ void __launch_bounds__(kDirWarps * 32, 1) k_cconv_direct(const ConvParams p, int n_filter_words) {
    extern __shared__ __align__(16) float smem[];
    float* filt = smem;                                               // [kc][COUT] (conv rows, then Dense rows)
    float* scratch = filt + ((n_filter_words + 3) & ~3);             // [kDirWarps][32][kDirRecWords]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint64_t filter_bar;
    tma::stage_block(filt, p.filters, n_filter_words, &filter_bar, kDirWarps * 32);  // the resident filter: TMA bulk copies
    float* rec = scratch + (size_t)warp * 32 * kDirRecWords;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int kx = p.gp.kx, kyx = p.gp.ky * p.gp.kx;
    const int cell_stride = p.cin * COUT;

    const int64_t n_warps = (int64_t)gridDim.x * kDirWarps;
    const int64_t n_out = conv_n_out(p);
    // the next point's row bounds and position are fetched while this point is worked on, its first neighbour indices once
    // the bounds have arrived (end of this point): a point starts with its operands in flight instead of three dependent trips
    int64_t o = (int64_t)blockIdx.x * kDirWarps + warp;
    int64_t rs_n = 0, re_n = 0;
    float ox_n = 0.f, oy_n = 0.f, oz_n = 0.f;
    int idx_first = 0;
    if (o < n_out) {
        rs_n = p.row_splits[o]; re_n = p.row_splits[o + 1];
        ox_n = __ldg(p.out_pos + 3 * o); oy_n = __ldg(p.out_pos + 3 * o + 1); oz_n = __ldg(p.out_pos + 3 * o + 2);
        if (!p.records && rs_n + lane < re_n) idx_first = __ldg(p.nbr_index + rs_n + lane);
    }
    for (; o < n_out; o += n_warps) {
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
        const float ox = ox_n, oy = oy_n, oz = oz_n;
        const int64_t rs = rs_n, re = re_n;
        const int64_t o2 = o + n_warps;
        if (o2 < n_out) {
            rs_n = p.row_splits[o2]; re_n = p.row_splits[o2 + 1];
            ox_n = __ldg(p.out_pos + 3 * o2); oy_n = __ldg(p.out_pos + 3 * o2 + 1); oz_n = __ldg(p.out_pos + 3 * o2 + 2);
        }
        float norm_acc = 0.0f;
        // without precomputed records the neighbour indices run one chunk ahead of the geometry that gathers their positions
        int idx_next = idx_first;
        for (int64_t c0 = rs; c0 < re; c0 += 32) {
            const int64_t n = c0 + lane;
            const int idx_cur = idx_next;
            if (!p.records && n + 32 < re) idx_next = __ldg(p.nbr_index + n + 32);
            const PairRec pr = p.records ? load_pair(p, n, n < re) : eval_pair_idx(p, n, n < re, idx_cur, ox, oy, oz);
            const int row = pr.row;
            norm_acc += pr.norm;
            // Parked record: {row, base cell offset, dx | dy << 16, dz} in filter WORDS (cell stride folded in once per pair by the
            // lane that owns the pair instead of once per corner by every lane), then the eight corner weights.
            int base = 0, dxy = 0, dzo = 0;
            float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
            if (row >= 0) {
                const PairGeom& g = pr.g;
                const int x0 = g.i0 & 0xff, y0 = (g.i0 >> 8) & 0xff, z0 = (g.i0 >> 16) & 0xff;
                const int x1 = g.i1 & 0xff, y1 = (g.i1 >> 8) & 0xff, z1 = (g.i1 >> 16) & 0xff;
                base = (z0 * kyx + y0 * kx + x0) * cell_stride;
                dxy = ((x1 - x0) * cell_stride) | (((y1 - y0) * kx * cell_stride) << 16);  // each difference is 0 or 1
                dzo = (z1 - z0) * kyx * cell_stride;
                wa = make_float4(g.wx0 * g.wy0 * g.wz0, g.wx1 * g.wy0 * g.wz0, g.wx0 * g.wy1 * g.wz0, g.wx1 * g.wy1 * g.wz0);
                wb = make_float4(g.wx0 * g.wy0 * g.wz1, g.wx1 * g.wy0 * g.wz1, g.wx0 * g.wy1 * g.wz1, g.wx1 * g.wy1 * g.wz1);
            }
            const unsigned active = __ballot_sync(0xffffffffu, row >= 0);
            const int cnt = __popc(active);
            __syncwarp();
            if (row >= 0) {
                float* r = rec + __popc(active & lt_mask) * kDirRecWords;
                *reinterpret_cast<int4*>(r) = make_int4(row, base, dxy, dzo);
                *reinterpret_cast<float4*>(r + 4) = wa;
                *reinterpret_cast<float4*>(r + 8) = wb;
            }
            if (lane < 3 && cnt + lane < ((cnt + 3) & ~3)) {  // pad the list to whole groups of four: no tail tests in the pair loop
                float* r = rec + (cnt + lane) * kDirRecWords;
                *reinterpret_cast<int4*>(r) = make_int4(-1, 0, 0, 0);
                *reinterpret_cast<float4*>(r + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(r + 8) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            for (int cb0 = 0; cb0 < p.cin; cb0 += 32) {
                const int ci = cb0 + lane;
                const bool ci_ok = ci < p.cin;
                const int cic = ci_ok ? ci : p.cin - 1;  // lanes beyond cin read a valid address and contribute f = 0
                float fc = 0.0f;
                if (p.ascc && ci_ok) {
                    fc = __ldg(p.inp_feat + o * p.inp_stride + ci);
                    if (p.relu_input) fc = fmaxf(fc, 0.0f);
                    fc *= p.feat_scale;
                }
                const float* fl = filt + (size_t)cic * COUT;  // + cell * cin * COUT
                for (int j = 0; j < cnt; j += 4) {
                    int4 hd[4];
                    float fv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        hd[u] = *reinterpret_cast<const int4*>(rec + (j + u) * kDirRecWords);
                        fv[u] = hd[u].x >= 0 ? __ldg(p.inp_feat + (int64_t)hd[u].x * p.inp_stride + cic) : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 wa2 = *reinterpret_cast<const float4*>(rec + (j + u) * kDirRecWords + 4);
                        const float4 wb2 = *reinterpret_cast<const float4*>(rec + (j + u) * kDirRecWords + 8);
                        float f = fv[u];
                        if (p.relu_input) f = fmaxf(f, 0.0f);
                        f = fmaf(f, p.feat_scale, fc);
                        if (!ci_ok || hd[u].x < 0) f = 0.0f;
                        const int dxo = hd[u].z & 0xffff, dyo = hd[u].z >> 16, dzo2 = hd[u].w;
                        const float* fb = fl + hd[u].y;
                        const float w[8] = {wa2.x, wa2.y, wa2.z, wa2.w, wb2.x, wb2.y, wb2.z, wb2.w};
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float* fp = fb + ((c & 1) ? dxo : 0) + ((c & 2) ? dyo : 0) + ((c & 4) ? dzo2 : 0);
                            const float wf = w[c] * f;
#pragma unroll
                            for (int co = 0; co < COUT; ++co) acc[co] = fmaf(wf, fp[co], acc[co]);
                        }
                    }
                }
            }
        }
        if (!p.records && o2 < n_out && rs_n + lane < re_n) idx_first = __ldg(p.nbr_index + rs_n + lane);
        if (p.normalize) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) norm_acc += __shfl_xor_sync(0xffffffffu, norm_acc, off);
            if (norm_acc != 0.0f) {
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[co] /= norm_acc;
            }
        }
        // fused Dense on the (relu'd, unscaled) centre features: rows kc_conv.. of the filter
        if (p.dense_cin > 0) {
            for (int ci = lane; ci < p.dense_cin; ci += 32) {
                float f = __ldg(p.dense_inp + o * p.dense_stride + ci);
                if (p.relu_input) f = fmaxf(f, 0.0f);
                const float* fp = filt + (size_t)(p.kc_conv + ci) * COUT;
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[co] = fmaf(f, fp[co], acc[co]);
            }
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float v = acc[co];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == co && co < p.cout) {
                if (p.bias) v += __ldg(p.bias + co);
                if (p.residual) v += __ldg(p.residual + o * p.residual_stride + co);
                float* dst = p.out + o * p.out_stride + co;
                if (p.accumulate) v += *dst;
                *dst = v;
            }
        }
    }
}

template <int COUT>
static int launch_direct(const ConvParams& p, cudaStream_t st) {
    const int n_words = p.kc * COUT;
    const size_t smem = ((size_t)((n_words + 3) & ~3) + (size_t)kDirWarps * 32 * kDirRecWords) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_cconv_direct<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_direct)");
        attr_set = true;
    }
    const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
    int64_t blocks = ceil_div(p.n_out, kDirWarps);
    if (blocks > 148 * ctas_per_sm) blocks = 148 * ctas_per_sm;  // persistent: the filter is staged once per CTA
    k_cconv_direct<COUT><<<(unsigned)blocks, kDirWarps * 32, smem, st>>>(p, n_words);
    DMCF_LAUNCH_CHECK("k_cconv_direct");
    return DMCF_OK;
}

// Tries the direct kernel; *handled = false means "not eligible".
int launch_cconv_direct(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p.cout > 4) return DMCF_OK;
    const size_t smem = ((size_t)p.kc * p.cout + 4 + (size_t)kDirWarps * 32 * kDirRecWords) * sizeof(float);
    if (smem > 200 * 1024) return DMCF_OK;
    if ((int64_t)p.gp.kx * p.cin * p.cout >= 32768) return DMCF_OK;  // corner offsets travel as 16-bit fields of the parked record
    *handled = true;
    switch (p.cout) {
        case 1: return launch_direct<1>(p, st);
        case 2: return launch_direct<2>(p, st);
        case 3: return launch_direct<3>(p, st);
        default: return launch_direct<4>(p, st);
    }
}

}  // namespace dmcf
