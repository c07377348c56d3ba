// continuous_conv forward, fused with the layer glue DMCF wraps around it (window, relu on the input, the
// antisymmetric "centre feature" term, per-particle Dense, bias, residual, add-merge).
// Replaces open3d.ml.tf.ops.continuous_conv as called at utils/convolutions.py:431/454/1054 of the reference.
//
// Kernel `k_cconv_tile` (generic over kernel size, channels, mappings):
//   a CTA owns a tile of MT consecutive out points.
//   phase 1  each warp builds the trilinear "patch" B[o][cell][ci] = sum_n a_n w_cell(n) g(f_n)[ci] of its points
//            in shared memory (geometry evaluated lane-parallel over 32 neighbours, then broadcast by shuffles;
//            lanes own input channels so the shared-memory read-modify-writes are conflict free);
//   phase 2  the CTA multiplies the [MT x KC] patch tile with the [KC x Cout] filter (+ appended Dense kernel):
//            split-K over warps, lane = output channel, MT accumulators per thread, the filter is read exactly
//            once per tile from L2 and every element is reused MT times; partial sums meet in shared memory;
//   epilogue normalise, bias, residual, store / accumulate.
// No patch matrix ever goes to HBM (the reference materialises [N, K*Cin] and runs a separate SGEMM).
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
    int cip, cp;              // pow2 lane groupings for input / output channels
    const float* filters;
    const float* out_pos;
    const float* inp_pos;
    const float* inp_feat;
    int64_t inp_stride;
    int64_t n_out, n_inp;
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
};

static constexpr int kConvWarps = 8;

template <int MT>
__global__ void __launch_bounds__(kConvWarps * 32, 1) k_cconv_tile(const ConvParams p) {
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                                    // [MT][kc_pad]
    float* red = patch + (size_t)MT * p.kc_pad;             // [kConvWarps][MT][cp]
    float* norm = red + (size_t)kConvWarps * MT * p.cp;     // [MT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile_base = (int64_t)blockIdx.x * MT;

    // ---- phase 0: clear the patch tile ------------------------------------------------------------------
    {
        float4* p4 = reinterpret_cast<float4*>(patch);
        const int n4 = MT * p.kc_pad / 4;
        for (int i = tid; i < n4; i += kConvWarps * 32) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    // ---- phase 1: patch build, one warp per out point -----------------------------------------------------
    const int cip = p.cip, cg = lane / cip, ci0 = lane % cip, n_cg = 32 / cip;
    const bool filter_nbr = p.nbr_hi > p.nbr_lo;
    for (int m = warp; m < MT; m += kConvWarps) {
        const int64_t o = tile_base + m;
        if (o >= p.n_out) break;
        float* prow = patch + (size_t)m * p.kc_pad;
        const float ox = __ldg(p.out_pos + 3 * o), oy = __ldg(p.out_pos + 3 * o + 1), oz = __ldg(p.out_pos + 3 * o + 2);
        const int64_t rs = p.row_splits[o], re = p.row_splits[o + 1];
        float norm_acc = 0.0f;
        for (int64_t c0 = rs; c0 < re; c0 += 32) {
            // lane-parallel geometry for up to 32 neighbours
            const int64_t n = c0 + lane;
            int row = -1;
            PairGeom g;
            g.i0 = g.i1 = 0;
            g.wx0 = g.wx1 = g.wy0 = g.wy1 = g.wz0 = g.wz1 = 0.0f;
            if (n < re) {
                const int idx = __ldg(p.nbr_index + n);
                bool keep = !filter_nbr || (idx >= p.nbr_lo && idx < p.nbr_hi);
                const int prow_idx = filter_nbr ? idx - p.nbr_lo : idx;
                if (keep) {
                    const float dx = __ldg(p.inp_pos + 3 * (int64_t)prow_idx) - ox;
                    const float dy = __ldg(p.inp_pos + 3 * (int64_t)prow_idx + 1) - oy;
                    const float dz = __ldg(p.inp_pos + 3 * (int64_t)prow_idx + 2) - oz;
                    if (p.skip_self && dx == 0.0f && dy == 0.0f && dz == 0.0f) keep = false;
                    if (keep) {
                        float a = 1.0f;
                        if (p.nbr_importance) {
                            a = __ldg(p.nbr_importance + n);
                        } else if (p.window != DMCF_WIN_NONE) {
                            const float q = __fdiv_rn(dist2_exact(dx, dy, dz), p.r2);
                            a = window_value(p.window, p.window_fac, q);
                        }
                        norm_acc += (p.nbr_importance || p.window != DMCF_WIN_NONE) ? a : 1.0f;
                        if (p.inp_importance) a *= __ldg(p.inp_importance + prow_idx);
                        g = pair_geometry(p.gp, dx, dy, dz);
                        g.wz0 *= a;
                        g.wz1 *= a;
                        row = prow_idx;
                    }
                }
            }
            const unsigned active = __ballot_sync(0xffffffffu, row >= 0);
            // broadcast every kept pair; lanes own (corner group, input channel)
            unsigned todo = active;
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int r_row = __shfl_sync(0xffffffffu, row, src);
                const int r_i0 = __shfl_sync(0xffffffffu, g.i0, src);
                const int r_i1 = __shfl_sync(0xffffffffu, g.i1, src);
                const float wx0 = __shfl_sync(0xffffffffu, g.wx0, src), wx1 = __shfl_sync(0xffffffffu, g.wx1, src);
                const float wy0 = __shfl_sync(0xffffffffu, g.wy0, src), wy1 = __shfl_sync(0xffffffffu, g.wy1, src);
                const float wz0 = __shfl_sync(0xffffffffu, g.wz0, src), wz1 = __shfl_sync(0xffffffffu, g.wz1, src);
                const float* frow = p.inp_feat + (int64_t)r_row * p.inp_stride;
                const float* crow = p.inp_feat + o * p.inp_stride;  // centre row (ascc: out set == inp set)
                for (int ci = ci0; ci < p.cin; ci += cip) {
                    float f = __ldg(frow + ci);
                    if (p.relu_input) f = fmaxf(f, 0.0f);
                    f *= p.feat_scale;
                    if (p.ascc) {
                        float fc = __ldg(crow + ci);
                        if (p.relu_input) fc = fmaxf(fc, 0.0f);
                        f += fc * p.feat_scale;
                    }
                    for (int c = cg; c < 8; c += n_cg) {
                        const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
                        const float w = (bx ? wx1 : wx0) * (by ? wy1 : wy0) * (bz ? wz1 : wz0);
                        if (w != 0.0f) {
                            const int sel = (bx ? r_i1 : r_i0) & 0xff;
                            const int sely = ((by ? r_i1 : r_i0) >> 8) & 0xff;
                            const int selz = ((bz ? r_i1 : r_i0) >> 16) & 0xff;
                            const int cell = (selz * p.gp.ky + sely) * p.gp.kx + sel;
                            prow[cell * p.cin + ci] += w * f;
                        }
                    }
                }
                __syncwarp();
            }
        }
        // fused Dense input: relu'd (unscaled) centre features appended as an extra "cell"
        if (p.dense_cin > 0) {
            const float* drow = p.dense_inp + o * p.dense_stride;
            for (int ci = lane; ci < p.dense_cin; ci += 32) {
                float f = __ldg(drow + ci);
                if (p.relu_input) f = fmaxf(f, 0.0f);
                prow[p.kc_conv + ci] = f;
            }
        }
        if (p.normalize) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) norm_acc += __shfl_xor_sync(0xffffffffu, norm_acc, off);
            if (lane == 0) norm[m] = norm_acc;
        }
    }
    __syncthreads();

    // ---- phase 2: [MT x kc] x [kc x cout], split-K over warps, lane = (k sub-slice, output channel) -------------
    const int cp = p.cp, ks = lane / cp, cl = lane % cp, n_ks = 32 / cp;
    const int kq_total = p.kc_pad / 4;
    for (int cb = 0; cb < p.cout; cb += 32) {
        const int co = cb + cl;
        const bool co_ok = co < p.cout;
        float acc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
        for (int kq = warp * n_ks + ks; kq < kq_total; kq += kConvWarps * n_ks) {
            const int k = kq * 4;
            const float* wrow = p.filters + (int64_t)k * p.cout + co;
            const float w0 = (co_ok && k + 0 < p.kc) ? __ldg(wrow) : 0.0f;
            const float w1 = (co_ok && k + 1 < p.kc) ? __ldg(wrow + p.cout) : 0.0f;
            const float w2 = (co_ok && k + 2 < p.kc) ? __ldg(wrow + 2 * p.cout) : 0.0f;
            const float w3 = (co_ok && k + 3 < p.kc) ? __ldg(wrow + 3 * p.cout) : 0.0f;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float4 pv = *reinterpret_cast<const float4*>(patch + (size_t)m * p.kc_pad + k);
                acc[m] = fmaf(pv.x, w0, acc[m]);
                acc[m] = fmaf(pv.y, w1, acc[m]);
                acc[m] = fmaf(pv.z, w2, acc[m]);
                acc[m] = fmaf(pv.w, w3, acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float v = acc[m];
            for (int off = cp; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (ks == 0) red[((size_t)warp * MT + m) * cp + cl] = v;
        }
        __syncthreads();
        for (int t = tid; t < MT * cp; t += kConvWarps * 32) {
            const int m = t / cp, c = t % cp;
            const int64_t o = tile_base + m;
            const int oc = cb + c;
            if (o < p.n_out && oc < p.cout) {
                float v = 0.0f;
#pragma unroll
                for (int w = 0; w < kConvWarps; ++w) v += red[((size_t)w * MT + m) * cp + c];
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

static int next_pow2(int v) {
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

static size_t conv_smem_bytes(int mt, int kc_pad, int cp) {
    return ((size_t)mt * kc_pad + (size_t)kConvWarps * mt * cp + mt) * sizeof(float);
}

template <int MT>
static int launch_cconv(const ConvParams& p, size_t smem, cudaStream_t st) {
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_cconv_tile<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_tile)");
        attr_set = true;
    }
    const int64_t tiles = ceil_div(p.n_out, MT);
    k_cconv_tile<MT><<<(unsigned)tiles, kConvWarps * 32, smem, st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_tile");
    return DMCF_OK;
}

}  // namespace dmcf

using namespace dmcf;

extern "C" int dmcf_cconv_forward(const dmcf_conv_desc* d, const float* filters, const float* out_positions, int64_t n_out,
                                  const float* inp_positions, const float* inp_features, int64_t inp_stride, int64_t n_inp,
                                  const float* inp_importance, const int32_t* neighbors_index,
                                  const int64_t* neighbors_row_splits, const float* neighbors_importance, const float* bias,
                                  const float* dense_inp, int64_t dense_stride, const float* residual, int64_t residual_stride,
                                  float* out, int64_t out_stride, void* stream) {
    DMCF_REQUIRE(d != nullptr, "cconv: desc is NULL");
    DMCF_REQUIRE(d->kernel_size[0] >= 1 && d->kernel_size[1] >= 1 && d->kernel_size[2] >= 1 && d->kernel_size[0] <= 255 &&
                     d->kernel_size[1] <= 255 && d->kernel_size[2] <= 255,
                 "cconv: kernel_size must be in [1,255]");
    DMCF_REQUIRE(d->cin >= 1 && d->cout >= 1, "cconv: cin/cout must be positive");
    DMCF_REQUIRE(d->mapping >= 0 && d->mapping <= 2, "cconv: unknown coordinate_mapping %d", d->mapping);
    DMCF_REQUIRE(d->interpolation >= 0 && d->interpolation <= 2, "cconv: unknown interpolation %d", d->interpolation);
    DMCF_REQUIRE(d->window >= 0 && d->window <= 5, "cconv: unknown window %d", d->window);
    DMCF_REQUIRE(d->extent > 0.0f, "cconv: extent must be positive");
    DMCF_REQUIRE(n_out >= 0 && n_inp >= 0, "cconv: negative point count");
    DMCF_REQUIRE(!(d->normalize && d->dense_cin > 0), "cconv: normalize cannot be combined with a fused Dense");
    DMCF_REQUIRE(!(d->ascc && d->nbr_hi > d->nbr_lo), "cconv: ascc needs the full neighbour set");
    DMCF_REQUIRE(!(d->ascc && n_out != n_inp), "cconv: ascc needs out set == inp set");
    DMCF_REQUIRE(d->dense_cin >= 0 && (d->dense_cin == 0 || dense_inp), "cconv: dense_inp is NULL");
    if (n_out == 0) return DMCF_OK;
    DMCF_REQUIRE(filters && out_positions && neighbors_row_splits && out, "cconv: NULL buffer");
    DMCF_REQUIRE(n_inp == 0 || (inp_positions && inp_features && neighbors_index), "cconv: NULL input buffer");
    DMCF_REQUIRE(inp_stride >= d->cin && out_stride >= d->cout, "cconv: row stride smaller than channel count");

    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.gp.kz = d->kernel_size[0]; p.gp.ky = d->kernel_size[1]; p.gp.kx = d->kernel_size[2];
    p.gp.mapping = d->mapping; p.gp.interp = d->interpolation; p.gp.align_corners = d->align_corners;
    p.gp.inv_extent = 1.0f / d->extent;
    p.gp.offx = d->offset[0]; p.gp.offy = d->offset[1]; p.gp.offz = d->offset[2];
    p.cin = d->cin; p.cout = d->cout;
    p.normalize = d->normalize; p.window = d->window; p.window_fac = d->window_fac;
    {
        const float r = 0.5f * d->extent;
        p.r2 = r * r;
    }
    p.relu_input = d->relu_input; p.feat_scale = d->feat_scale;
    p.ascc = d->ascc; p.skip_self = d->skip_self; p.nbr_lo = d->nbr_lo; p.nbr_hi = d->nbr_hi;
    p.dense_cin = d->dense_cin; p.accumulate = d->accumulate;
    const int64_t cells = (int64_t)p.gp.kx * p.gp.ky * p.gp.kz;
    DMCF_REQUIRE(cells * d->cin + d->dense_cin < (1 << 24), "cconv: filter too large");
    p.kc_conv = (int)(cells * d->cin);
    p.kc = p.kc_conv + d->dense_cin;
    p.kc_pad = (p.kc + 3) / 4 * 4;
    p.cip = next_pow2(d->cin < 32 ? d->cin : 32);
    p.cp = next_pow2(d->cout < 32 ? d->cout : 32);
    p.filters = filters; p.out_pos = out_positions; p.inp_pos = inp_positions; p.inp_feat = inp_features;
    p.inp_stride = inp_stride; p.n_out = n_out; p.n_inp = n_inp; p.inp_importance = inp_importance;
    p.nbr_index = neighbors_index; p.row_splits = neighbors_row_splits; p.nbr_importance = neighbors_importance;
    p.bias = bias; p.dense_inp = dense_inp; p.dense_stride = dense_stride;
    p.residual = residual; p.residual_stride = residual_stride; p.out = out; p.out_stride = out_stride;

    const size_t limit = 227 * 1024;
    cudaStream_t st = (cudaStream_t)stream;
    // largest tile that fits (bigger tile = fewer passes over the filter); small patches prefer several CTAs per SM
    if (conv_smem_bytes(32, p.kc_pad, p.cp) <= limit / 2) return launch_cconv<32>(p, conv_smem_bytes(32, p.kc_pad, p.cp), st);
    if (conv_smem_bytes(32, p.kc_pad, p.cp) <= limit) return launch_cconv<32>(p, conv_smem_bytes(32, p.kc_pad, p.cp), st);
    if (conv_smem_bytes(24, p.kc_pad, p.cp) <= limit) return launch_cconv<24>(p, conv_smem_bytes(24, p.kc_pad, p.cp), st);
    if (conv_smem_bytes(16, p.kc_pad, p.cp) <= limit) return launch_cconv<16>(p, conv_smem_bytes(16, p.kc_pad, p.cp), st);
    if (conv_smem_bytes(8, p.kc_pad, p.cp) <= limit) return launch_cconv<8>(p, conv_smem_bytes(8, p.kc_pad, p.cp), st);
    return set_error(DMCF_ERR_UNSUPPORTED, "cconv: filter %dx%dx%dx%d needs %zu B of shared memory per 8 points (> %zu)",
                     p.gp.kz, p.gp.ky, p.gp.kx, d->cin, conv_smem_bytes(8, p.kc_pad, p.cp), limit);
}
