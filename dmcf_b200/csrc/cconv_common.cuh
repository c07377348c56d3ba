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

// ---- phase 2 + epilogue: [MT x kc] x [kc x cout], split-K over warps, lane = (k sub-slice, output channel) ----------
// `red` may alias `patch` when RED_ALIASES_PATCH (only valid for cout <= 32: the patch is dead once every warp has
// left the k loop, which the extra barrier guarantees).
template <int MT, int NW, bool RED_ALIASES_PATCH>
__device__ __forceinline__ void cconv_phase2(const ConvParams& p, const float* patch, float* red, const float* norm,
                                             int64_t tile_base) {
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
            if (o < p.n_out && oc < p.cout) {
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

}  // namespace dmcf
