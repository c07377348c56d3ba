// continuous_conv forward, register-patch variant for filters with a compile-time grid (4x4x4, 1x8x8, 1x8x1),
// linear interpolation and <= 32 input channels -- the wide layers that dominate a DMCF step.
//
// Same tiling as k_cconv_tile (cconv.cu) but phase 1 keeps the whole trilinear patch of a point in REGISTERS:
// one warp per out point, lane = input channel, acc[cell] for every filter cell.  The pair geometry is evaluated
// lane-parallel for 32 neighbours and parked as compact records {row, base cell, 8 corner weights} in a per-warp
// shared-memory scratch; the warp then walks the records (broadcast LDS.128), gathers the neighbour's feature row
// (coalesced, four rows in flight) and a warp-uniform switch on the base cell turns the 8 corner updates into FFMAs
// on statically indexed registers: no shared-memory read-modify-write and ~3x fewer instructions per pair than the
// generic kernel.  The finished patch row is written once to shared memory for the shared phase 2 (patch x filter).
#include "cconv_scatter.cuh"

namespace dmcf {

static constexpr int kRecWords = 12;  // {row, base, pad, pad, w0..w3, w4..w7}: 48 B, 16 B aligned, conflict-free STS.128

// p.filters / p.kc_conv / p.kc / p.kc_pad describe THIS instance's slice of the filter (planes [ZLO, ZLO+NZ) followed by
// the Dense rows if this launch carries them).
template <int KZ, int KY, int KX, int ZLO, int NZ, int MT, int NW, int MINB, bool RED_ALIAS>
__global__ void __launch_bounds__(NW * 32, MINB) k_cconv_wide(const ConvParams p) {
    using G = FilterGrid<KZ, KY, KX>;
    constexpr int KL = NZ * KY * KX;  // filter cells of this instance
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;  // k-quad major [kc_pad/4][MT+1][4] (also [NW][MT][32] partial sums if RED_ALIAS)
    const size_t tile_words = (size_t)(p.kc_pad / 4) * (MT + 1) * 4, red_words = (size_t)NW * MT * 32;
    const size_t patch_words = (RED_ALIAS && red_words > tile_words) ? red_words : tile_words;
    float* scratch = patch + patch_words;                      // [NW][32][kRecWords]
    float* norm = scratch + (size_t)NW * 32 * kRecWords;       // [MT]
    float* red = RED_ALIAS ? patch : norm + MT;                // [NW][MT][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile_base = (int64_t)blockIdx.x * MT;
    const int64_t n_out = conv_n_out(p);
    if (tile_base >= n_out) return;
    float* rec = scratch + (size_t)warp * 32 * kRecWords;
    const bool lane_ci = lane < p.cin;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int m = warp; m < MT; m += NW) {
        const int64_t o = tile_base + m;
        if (o >= n_out) {  // keep unused rows finite (they are multiplied, never stored)
            for (int k = lane; k < p.kc_pad; k += 32) patch[patchq_index<MT>(m, k)] = 0.0f;
            continue;
        }
        float acc[KL];
#pragma unroll
        for (int c = 0; c < KL; ++c) acc[c] = 0.0f;
        const float ox = __ldg(p.out_pos + 3 * o), oy = __ldg(p.out_pos + 3 * o + 1), oz = __ldg(p.out_pos + 3 * o + 2);
        const int64_t rs = p.row_splits[o], re = p.row_splits[o + 1];
        float fc = 0.0f;  // centre feature of the antisymmetric layer (out set == inp set)
        if (p.ascc && lane_ci) {
            fc = __ldg(p.inp_feat + o * p.inp_stride + lane);
            if (p.relu_input) fc = fmaxf(fc, 0.0f);
            fc *= p.feat_scale;
        }
        float norm_acc = 0.0f;
        for (int64_t c0 = rs; c0 < re; c0 += 32) {
            // ---- lane-parallel geometry of up to 32 neighbours -> compact records in the warp's scratch ----
            const int64_t n = c0 + lane;
            const PairRec pr = pair_record(p, n, n < re, ox, oy, oz);
            int row2 = pr.row;
            norm_acc += pr.norm;
            int b = 0;
            float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
            if (row2 >= 0) {
                int bx, by, bz;
                float xl, xh, yl, yh, zl, zh;
                base_axis(KX, pr.g.i0 & 0xff, pr.g.wx0, pr.g.wx1, bx, xl, xh);
                base_axis(KY, (pr.g.i0 >> 8) & 0xff, pr.g.wy0, pr.g.wy1, by, yl, yh);
                base_axis(KZ, (pr.g.i0 >> 16) & 0xff, pr.g.wz0, pr.g.wz1, bz, zl, zh);
                if (NZ < KZ && (bz + 1 < ZLO || bz >= ZLO + NZ)) row2 = -1;  // touches none of this instance's planes
                b = (bz * G::NBY + by) * G::NBX + bx;
                wa = make_float4(xl * yl * zl, xh * yl * zl, xl * yh * zl, xh * yh * zl);
                wb = make_float4(xl * yl * zh, xh * yl * zh, xl * yh * zh, xh * yh * zh);
            }
            const int row = row2;
            const unsigned active = __ballot_sync(0xffffffffu, row >= 0);
            const int cnt = __popc(active);
            __syncwarp();  // previous chunk's records fully consumed
            if (row >= 0) {
                float* r = rec + __popc(active & lt_mask) * kRecWords;
                *reinterpret_cast<int4*>(r) = make_int4(row, b, 0, 0);
                *reinterpret_cast<float4*>(r + 4) = wa;
                *reinterpret_cast<float4*>(r + 8) = wb;
            }
            __syncwarp();
            // ---- walk the records: ONE copy of the scatter switch (small code: the 27 cases stay in the instruction
            // cache; with records sorted by base cell consecutive pairs take neighbouring cases), feature rows of the
            // next PF pairs already in flight ----
            auto gather = [&](int j) -> float {
                return (j < cnt && lane_ci) ? __ldg(p.inp_feat + (int64_t)__float_as_int(rec[j * kRecWords]) * p.inp_stride + lane)
                                            : 0.0f;
            };
            float fq0 = gather(0), fq1 = gather(1), fq2 = gather(2), fq3 = gather(3);
#pragma unroll 1
            for (int j = 0; j < cnt; ++j) {
                float f = fq0;
                fq0 = fq1; fq1 = fq2; fq2 = fq3;
                fq3 = gather(j + 4);
                const int b = __float_as_int(rec[j * kRecWords + 1]);
                const float4 wa2 = *reinterpret_cast<const float4*>(rec + j * kRecWords + 4);
                const float4 wb2 = *reinterpret_cast<const float4*>(rec + j * kRecWords + 8);
                if (p.relu_input) f = fmaxf(f, 0.0f);
                f = fmaf(f, p.feat_scale, fc);
                scatter_switch<KZ, KY, KX, ZLO, NZ>(b, acc, wa2, wb2, f);
            }
        }
        // ---- patch row -> shared memory (lane = channel: conflict-free), Dense columns, padding ----
        if (lane_ci) {
            if ((p.cin & 3) == 0) {  // k = c*cin + lane: the k-quad advances by cin/4 per cell -> one running pointer
                float* pp = patch + patchq_index<MT>(m, lane);
                const int step = p.cin * (MT + 1);
#pragma unroll
                for (int c = 0; c < KL; ++c) pp[c * step] = acc[c];
            } else {
#pragma unroll
                for (int c = 0; c < KL; ++c) patch[patchq_index<MT>(m, c * p.cin + lane)] = acc[c];
            }
        }
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
    __syncthreads();
    if (p.debug_wrap_w & 2) return;  // timing experiment: phase 1 only
    cconv_phase2_v2<MT, NW, RED_ALIAS>(p, patch, red, norm, tile_base, n_out);
}

static size_t wide_smem_bytes(int mt, int nw, int kc_pad, int cp, bool alias) {
    (void)cp;
    size_t patch = (size_t)(kc_pad / 4) * (mt + 1) * 4, red = (size_t)nw * mt * 32;
    size_t words = (size_t)nw * 32 * kRecWords + mt + (alias ? (patch > red ? patch : red) : patch + red);
    return words * sizeof(float);
}

template <int KZ, int KY, int KX, int ZLO, int NZ, int MT, int NW, int MINB, bool ALIAS>
static int launch_wide(const ConvParams& p, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = k_cconv_wide<KZ, KY, KX, ZLO, NZ, MT, NW, MINB, ALIAS>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_wide)");
        attr_set = true;
    }
    const int64_t tiles = ceil_div(p.n_out, MT);
    kern<<<(unsigned)tiles, NW * 32, wide_smem_bytes(MT, NW, p.kc_pad, p.cp, ALIAS), st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_wide");
    return DMCF_OK;
}

// whole filter in one launch (one CTA per SM)
template <int KZ, int KY, int KX>
static int launch_wide_grid(const ConvParams& p, cudaStream_t st, bool* handled) {
    const size_t limit = 227 * 1024;
    *handled = true;
    if (p.cout <= 32) {  // partial sums may reuse the patch tile
        if (wide_smem_bytes(32, 16, p.kc_pad, p.cp, true) <= limit) return launch_wide<KZ, KY, KX, 0, KZ, 32, 16, 1, true>(p, st);
        if (wide_smem_bytes(24, 12, p.kc_pad, p.cp, true) <= limit) return launch_wide<KZ, KY, KX, 0, KZ, 24, 12, 1, true>(p, st);
        if (wide_smem_bytes(20, 16, p.kc_pad, p.cp, true) <= limit) return launch_wide<KZ, KY, KX, 0, KZ, 20, 16, 1, true>(p, st);
        if (wide_smem_bytes(16, 16, p.kc_pad, p.cp, true) <= limit) return launch_wide<KZ, KY, KX, 0, KZ, 16, 16, 1, true>(p, st);
    } else {
        if (wide_smem_bytes(24, 12, p.kc_pad, p.cp, false) <= limit) return launch_wide<KZ, KY, KX, 0, KZ, 24, 12, 1, false>(p, st);
        if (wide_smem_bytes(16, 16, p.kc_pad, p.cp, false) <= limit) return launch_wide<KZ, KY, KX, 0, KZ, 16, 16, 1, false>(p, st);
    }
    *handled = false;
    return DMCF_OK;
}

// 4x4x4 filters whose patch tile would own the SM: two launches over the z-plane halves {0,1} and {2,3}.  Half the patch
// per point -> two CTAs (24 warps) per SM, which is what hides the feature-gather and shared-memory latencies; the pairs
// straddling the halves are visited by both launches.  The second launch carries Dense / bias / residual and accumulates.
static int launch_wide_444_split(const ConvParams& p, cudaStream_t st, bool* handled) {
    constexpr int MT = 20, NW = 12;
    *handled = false;
    if (p.cout > 32 || p.normalize) return DMCF_OK;
    ConvParams lo = p, hi = p;
    const int half_rows = 2 * 16 * p.cin;
    lo.kc_conv = lo.kc = half_rows;
    lo.kc_pad = (lo.kc + 3) / 4 * 4;
    lo.dense_cin = 0; lo.dense_inp = nullptr; lo.bias = nullptr; lo.residual = nullptr;
    hi.filters = p.filters + (size_t)half_rows * p.cout;
    hi.kc_conv = half_rows;
    hi.kc = half_rows + p.dense_cin;
    hi.kc_pad = (hi.kc + 3) / 4 * 4;
    hi.accumulate = 1;
    if (((uintptr_t)hi.filters & 15) != 0) return DMCF_OK;
    if (2 * wide_smem_bytes(MT, NW, hi.kc_pad, p.cp, true) > 227 * 1024) return DMCF_OK;
    *handled = true;
    int rc = launch_wide<4, 4, 4, 0, 2, MT, NW, 2, true>(lo, st);
    if (rc) return rc;
    return launch_wide<4, 4, 4, 2, 2, MT, NW, 2, true>(hi, st);
}

// Tries the register-patch kernel; *handled = false means "not eligible, use the generic kernel".
int launch_cconv_wide(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p.gp.interp != DMCF_INTERP_LINEAR || p.cin > 32) return DMCF_OK;
    if (p.cout % 4 != 0 || ((uintptr_t)p.filters & 15) != 0) return DMCF_OK;  // phase 2 reads filter rows as float4
    if (p.gp.kz == 4 && p.gp.ky == 4 && p.gp.kx == 4) {
        if (p.use_zsplit && wide_smem_bytes(32, 16, p.kc_pad, p.cp, true) > 113 * 1024) {
            int rc = launch_wide_444_split(p, st, handled);
            if (rc || *handled) return rc;
        }
        return launch_wide_grid<4, 4, 4>(p, st, handled);
    }
    if (p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 8) return launch_wide_grid<1, 8, 8>(p, st, handled);
    if (p.gp.kz == 1 && p.gp.ky == 8 && p.gp.kx == 1) return launch_wide_grid<1, 8, 1>(p, st, handled);
    return DMCF_OK;
}

}  // namespace dmcf
