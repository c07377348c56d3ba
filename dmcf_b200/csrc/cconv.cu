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
#include <cstdlib>

#include "cconv_common.cuh"

namespace dmcf {

template <int MT, int NW, int CIP>
__global__ void __launch_bounds__(NW * 32, 1) k_cconv_tile(const ConvParams p) {
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                            // [MT][kc_pad]
    float* red = patch + (size_t)MT * p.kc_pad;     // [NW][MT][cp]
    float* norm = red + (size_t)NW * MT * p.cp;     // [MT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile_base = (int64_t)blockIdx.x * MT;
    const int64_t n_out = conv_n_out(p);
    if (tile_base >= n_out) return;

    // ---- phase 0: clear the patch tile ------------------------------------------------------------------
    {
        float4* p4 = reinterpret_cast<float4*>(patch);
        const int n4 = MT * p.kc_pad / 4;
        for (int i = tid; i < n4; i += NW * 32) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    // ---- phase 1: patch build, one warp per out point -----------------------------------------------------
    // lanes = (corner group, input channel): CIP = pow2 >= min(cin,32) channels, CG = 32/CIP corners in parallel
    constexpr int cip = CIP, n_cg = (32 / CIP > 8) ? 8 : 32 / CIP;  // at most the 8 corners; extra lanes idle
    const int cg = lane / CIP, ci0 = lane % CIP;
    const bool lane_ok = cg < n_cg;
    for (int m = warp; m < MT; m += NW) {
        const int64_t o = tile_base + m;
        if (o >= n_out) break;
        float* prow = patch + (size_t)m * p.kc_pad;
        const float ox = __ldg(p.out_pos + 3 * o), oy = __ldg(p.out_pos + 3 * o + 1), oz = __ldg(p.out_pos + 3 * o + 2);
        const int64_t rs = p.row_splits[o], re = p.row_splits[o + 1];
        const float* crow = p.inp_feat + o * p.inp_stride;  // centre row (ascc: out set == inp set)
        float norm_acc = 0.0f;
        for (int64_t c0 = rs; c0 < re; c0 += 32) {
            // lane-parallel geometry for up to 32 neighbours
            const int64_t n = c0 + lane;
            const PairRec pr = pair_record(p, n, n < re, ox, oy, oz);
            const int row = pr.row;
            const PairGeom g = pr.g;
            norm_acc += pr.norm;
            // broadcast the kept pairs four at a time (four independent feature gathers in flight per lane);
            // lanes own (corner group, input channel)
            unsigned todo = __ballot_sync(0xffffffffu, row >= 0);
            while (todo) {
                constexpr int U = 4;
                int r_row[U], r_i0[U], r_i1[U];
                float wx0[U], wx1[U], wy0[U], wy1[U], wz0[U], wz1[U];
                int first = __ffs(todo) - 1;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool valid = todo != 0;
                    const int src = valid ? __ffs(todo) - 1 : first;
                    if (valid) todo &= todo - 1;
                    r_row[u] = __shfl_sync(0xffffffffu, row, src);
                    r_i0[u] = __shfl_sync(0xffffffffu, g.i0, src);
                    r_i1[u] = __shfl_sync(0xffffffffu, g.i1, src);
                    wx0[u] = __shfl_sync(0xffffffffu, g.wx0, src);
                    wx1[u] = __shfl_sync(0xffffffffu, g.wx1, src);
                    wy0[u] = __shfl_sync(0xffffffffu, g.wy0, src);
                    wy1[u] = __shfl_sync(0xffffffffu, g.wy1, src);
                    const float z0 = __shfl_sync(0xffffffffu, g.wz0, src), z1 = __shfl_sync(0xffffffffu, g.wz1, src);
                    wz0[u] = valid ? z0 : 0.0f;
                    wz1[u] = valid ? z1 : 0.0f;
                }
                for (int cb0 = 0; cb0 < p.cin; cb0 += cip) {  // warp-uniform trip count (__syncwarp inside)
                    const int ci = cb0 + ci0;
                    const bool ci_ok = lane_ok && ci < p.cin;
                    float f[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) f[u] = ci_ok ? __ldg(p.inp_feat + (int64_t)r_row[u] * p.inp_stride + ci) : 0.0f;
                    float fc = 0.0f;
                    if (p.ascc && ci_ok) {
                        fc = __ldg(crow + ci);
                        if (p.relu_input) fc = fmaxf(fc, 0.0f);
                        fc *= p.feat_scale;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        float fv = f[u];
                        if (p.relu_input) fv = fmaxf(fv, 0.0f);
                        fv = fv * p.feat_scale + fc;
                        const int x0 = r_i0[u] & 0xff, x1 = r_i1[u] & 0xff;
                        const int zy00 = (((r_i0[u] >> 16) & 0xff) * p.gp.ky + ((r_i0[u] >> 8) & 0xff)) * p.gp.kx;
                        const int zy01 = (((r_i0[u] >> 16) & 0xff) * p.gp.ky + ((r_i1[u] >> 8) & 0xff)) * p.gp.kx;
                        const int zy10 = (((r_i1[u] >> 16) & 0xff) * p.gp.ky + ((r_i0[u] >> 8) & 0xff)) * p.gp.kx;
                        const int zy11 = (((r_i1[u] >> 16) & 0xff) * p.gp.ky + ((r_i1[u] >> 8) & 0xff)) * p.gp.kx;
                        float* pc = prow + ci;
#pragma unroll
                        for (int j = 0; j < 8 / n_cg; ++j) {
                            const int c = j * n_cg + cg;  // compile-time when CIP == 32
                            const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
                            const float w = (bx ? wx1[u] : wx0[u]) * (by ? wy1[u] : wy0[u]) * (bz ? wz1[u] : wz0[u]);
                            if (ci_ok && w != 0.0f) {
                                const int zy = bz ? (by ? zy11 : zy10) : (by ? zy01 : zy00);
                                const int cell = zy + (bx ? x1 : x0);
                                pc[cell * p.cin] += w * fv;
                            }
                        }
                        // with several corner groups per channel two pairs may hit one address from different lanes
                        if (n_cg > 1) __syncwarp();
                    }
                }
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
    if (p.patch_out) {  // dmcf_cconv_patches: export the patch rows instead of multiplying them with the filter
        for (int idx = tid; idx < MT * p.kc_conv; idx += NW * 32) {
            const int m = idx / p.kc_conv, k = idx - m * p.kc_conv;
            const int64_t o = tile_base + m;
            if (o < n_out) p.patch_out[o * p.patch_stride + k] = patch[(size_t)m * p.kc_pad + k];
        }
        return;
    }

    cconv_phase2<MT, NW, false>(p, patch, red, norm, tile_base, n_out);
}

static int next_pow2(int v) {
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

static size_t conv_smem_bytes(int mt, int nw, int kc_pad, int cp) {
    return ((size_t)mt * kc_pad + (size_t)nw * mt * cp + mt) * sizeof(float);
}

template <int MT, int NW, int CIP>
static int launch_cconv_cip(const ConvParams& p, cudaStream_t st) {
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_cconv_tile<MT, NW, CIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_tile)");
        attr_set = true;
    }
    const int64_t tiles = ceil_div(p.n_out, MT);
    k_cconv_tile<MT, NW, CIP><<<(unsigned)tiles, NW * 32, conv_smem_bytes(MT, NW, p.kc_pad, p.cp), st>>>(p);
    DMCF_LAUNCH_CHECK("k_cconv_tile");
    return DMCF_OK;
}

template <int MT, int NW>
static int launch_cconv(const ConvParams& p, cudaStream_t st) {
    switch (p.cip) {
        case 1: return launch_cconv_cip<MT, NW, 1>(p, st);
        case 2: return launch_cconv_cip<MT, NW, 2>(p, st);
        case 4: return launch_cconv_cip<MT, NW, 4>(p, st);
        case 8: return launch_cconv_cip<MT, NW, 8>(p, st);
        case 16: return launch_cconv_cip<MT, NW, 16>(p, st);
        default: return launch_cconv_cip<MT, NW, 32>(p, st);
    }
}

int launch_cconv_ws(const ConvParams& p, cudaStream_t st, bool* handled);      // cconv_ws.cu
int launch_cconv_lean(const ConvParams& p, cudaStream_t st, bool* handled);    // cconv_lean.cu
int launch_cconv_apatch(const ConvParams& p, cudaStream_t st, bool* handled);  // cconv_apatch.cu
int launch_cconv_wide(const ConvParams& p, cudaStream_t st, bool* handled);    // cconv_wide.cu
int launch_cconv_direct(const ConvParams& p, cudaStream_t st, bool* handled);  // cconv_direct.cu
int launch_cconv_narrow(const ConvParams& p, cudaStream_t st, bool* handled);  // cconv_narrow.cu
std::atomic<int> g_kernel_options{3};



static int fill_params(const dmcf_conv_desc* d, const float* filters, const float* out_positions, int64_t n_out,
                       const float* inp_positions, const float* inp_features, int64_t inp_stride, int64_t n_inp,
                       const float* inp_importance, const int32_t* neighbors_index, const int64_t* neighbors_row_splits,
                       const float* neighbors_importance, ConvParams* pp) {
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
    DMCF_REQUIRE(!(d->ascc && n_out > n_inp), "cconv: ascc needs out point o == input row o (out set a prefix of the inp set)");
    ConvParams& p = *pp;
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
    p.dense_cin = d->dense_cin; p.accumulate = d->accumulate; p.filter_antisym = d->filter_antisym;
    p.n_out_dev = d->n_out_dev;
    if (d->block_cin > 0) {
        DMCF_REQUIRE(d->block_cin < d->cin && d->block_cout[0] >= 1 && d->block_cout[1] >= 1 &&
                         d->block_cout[0] + d->block_cout[1] <= d->cout,
                     "cconv: block promise (block_cin %d, block_cout %d + %d) does not fit cin %d / cout %d", d->block_cin,
                     d->block_cout[0], d->block_cout[1], d->cin, d->cout);
        p.blk_ca = d->block_cin; p.blk_na = d->block_cout[0]; p.blk_nb = d->block_cout[1];
    }
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
    return DMCF_OK;
}

// one warp per out point, lanes over its neighbours: coalesced writes of the 9 record arrays
__global__ void __launch_bounds__(256) k_cconv_prepare(const ConvParams p, float* __restrict__ records) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t P = p.n_pairs;
    const int64_t n_out = conv_n_out(p);
    for (int64_t o = warp0; o < n_out; o += n_warps) {
        const float ox = __ldg(p.out_pos + 3 * o), oy = __ldg(p.out_pos + 3 * o + 1), oz = __ldg(p.out_pos + 3 * o + 2);
        const int64_t rs = p.row_splits[o], re = p.row_splits[o + 1];
        for (int64_t c0 = rs; c0 < re; c0 += 32) {
            const int64_t n = c0 + lane;
            PairRec r = eval_pair(p, n, n < re, ox, oy, oz);
            // Order the chunk by the corner block's base cell (dropped pairs last): the conv kernels walk the records
            // in order, so consecutive pairs land in neighbouring cases of their scatter switch (instruction-cache
            // locality).  Bitonic sort of (key, lane) over the warp, then one gather of the 9 fields.
            // (key = linear index of the block's base cell, the corner-0 cell clamped to <= fs-2 per axis: exactly the
            // case index of the register-patch kernels, so their merge walk consumes a chunk in one sweep)
            unsigned cellkey = 0xffffffu;
            if (r.row >= 0) {
                const int nbx = max(p.gp.kx - 1, 1), nby = max(p.gp.ky - 1, 1);
                const int bx = min(r.g.i0 & 0xff, nbx - 1), by = min((r.g.i0 >> 8) & 0xff, nby - 1);
                const int bz = min((r.g.i0 >> 16) & 0xff, max(p.gp.kz - 1, 1) - 1);
                cellkey = (unsigned)((bz * nby + by) * nbx + bx);
            }
            unsigned key = (cellkey << 5) | (unsigned)lane;
            int src;
            if (re - c0 <= 8) {
                // short chunk (the tail of a 33..40-pair row is the common case): rank of this lane's key among the
                // first eight lanes instead of the full 32-lane bitonic network
                int rank = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) rank += __shfl_sync(0xffffffffu, key, j) < key ? 1 : 0;
                // lane `rank` must pull from this lane: invert the permutation (lanes >= 8 hold no pairs and stay put)
                src = lane;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (__shfl_sync(0xffffffffu, rank, j) == lane && lane < 8) src = j;
            } else {
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const unsigned other = __shfl_xor_sync(0xffffffffu, key, j);
                        const bool up = ((lane & k) == 0);
                        const bool lower = ((lane & j) == 0);
                        const unsigned mn = min(key, other), mx = max(key, other);
                        key = (up == lower) ? mn : mx;
                    }
                }
                src = key & 31;
            }
            r.row = __shfl_sync(0xffffffffu, r.row, src);
            r.g.i0 = __shfl_sync(0xffffffffu, r.g.i0, src);
            r.g.i1 = __shfl_sync(0xffffffffu, r.g.i1, src);
            r.g.wx0 = __shfl_sync(0xffffffffu, r.g.wx0, src); r.g.wx1 = __shfl_sync(0xffffffffu, r.g.wx1, src);
            r.g.wy0 = __shfl_sync(0xffffffffu, r.g.wy0, src); r.g.wy1 = __shfl_sync(0xffffffffu, r.g.wy1, src);
            r.g.wz0 = __shfl_sync(0xffffffffu, r.g.wz0, src); r.g.wz1 = __shfl_sync(0xffffffffu, r.g.wz1, src);
            if (n < re) {  // slots of the chunk beyond the row end hold dropped pairs (sorted last)
                float* f = records + n;
                f[0] = __int_as_float(r.row);
                f[P] = __int_as_float(r.g.i0);
                f[2 * P] = __int_as_float(r.g.i1);
                f[3 * P] = r.g.wx0; f[4 * P] = r.g.wx1;
                f[5 * P] = r.g.wy0; f[6 * P] = r.g.wy1;
                f[7 * P] = r.g.wz0; f[8 * P] = r.g.wz1;
            }
        }
    }
}

}  // namespace dmcf

using namespace dmcf;

extern "C" size_t dmcf_cconv_records_bytes(int64_t n_pairs) {
    return (size_t)(n_pairs > 0 ? n_pairs : 0) * kRecordFields * sizeof(float);
}

extern "C" int dmcf_cconv_prepare(const dmcf_conv_desc* d, const float* out_positions, int64_t n_out, const float* inp_positions,
                                  int64_t n_inp, const float* inp_importance, const int32_t* neighbors_index,
                                  const int64_t* neighbors_row_splits, const float* neighbors_importance, int64_t n_pairs,
                                  float* records, void* stream) {
    ConvParams p;
    int rc = fill_params(d, nullptr, out_positions, n_out, inp_positions, nullptr, 0, n_inp, inp_importance, neighbors_index,
                         neighbors_row_splits, neighbors_importance, &p);
    if (rc) return rc;
    DMCF_REQUIRE(!d->normalize, "cconv_prepare: records do not carry the normaliser; call cconv_forward without records");
    DMCF_REQUIRE(n_pairs >= 0, "cconv_prepare: negative pair count");
    if (n_out == 0 || n_pairs == 0) return DMCF_OK;
    DMCF_REQUIRE(out_positions && inp_positions && neighbors_index && neighbors_row_splits && records, "cconv_prepare: NULL buffer");
    p.n_pairs = n_pairs;
    int64_t blocks = ceil_div(n_out, 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_cconv_prepare<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, records);
    DMCF_LAUNCH_CHECK("k_cconv_prepare");
    return DMCF_OK;
}

extern "C" int dmcf_cconv_forward(const dmcf_conv_desc* d, const float* filters, const float* out_positions, int64_t n_out,
                                  const float* inp_positions, const float* inp_features, int64_t inp_stride, int64_t n_inp,
                                  const float* inp_importance, const int32_t* neighbors_index,
                                  const int64_t* neighbors_row_splits, const float* neighbors_importance, const float* bias,
                                  const float* dense_inp, int64_t dense_stride, const float* residual, int64_t residual_stride,
                                  float* out, int64_t out_stride, const float* pair_records, int64_t n_pairs, void* stream) {
    ConvParams p;
    int rc = fill_params(d, filters, out_positions, n_out, inp_positions, inp_features, inp_stride, n_inp, inp_importance,
                         neighbors_index, neighbors_row_splits, neighbors_importance, &p);
    if (rc) return rc;
    DMCF_REQUIRE(d->dense_cin >= 0 && (d->dense_cin == 0 || dense_inp), "cconv: dense_inp is NULL");
    DMCF_REQUIRE(!(pair_records && d->normalize), "cconv: pair records cannot be combined with normalize");
    if (n_out == 0) return DMCF_OK;
    DMCF_REQUIRE(filters && out_positions && neighbors_row_splits && out, "cconv: NULL buffer");
    DMCF_REQUIRE(n_inp == 0 || (inp_features && (pair_records || (inp_positions && neighbors_index))), "cconv: NULL input buffer");
    DMCF_REQUIRE(inp_stride >= d->cin && out_stride >= d->cout, "cconv: row stride smaller than channel count");
    p.bias = bias; p.dense_inp = dense_inp; p.dense_stride = dense_stride;
    p.residual = residual; p.residual_stride = residual_stride; p.out = out; p.out_stride = out_stride;
    p.records = pair_records; p.n_pairs = n_pairs;

    const size_t limit = 227 * 1024;
    cudaStream_t st = (cudaStream_t)stream;
    // specialised kernels (dmcf_set_kernel_options(0) forces the generic kernel; the parity tests run both)
    const int options = g_kernel_options.load(std::memory_order_relaxed);
    p.debug_wrap_w = (options >> 8) & 15;
    p.use_zsplit = (options >> 2) & 1;
    p.no_multipair = (options >> 5) & 1;
    p.no_lean_tc = (options >> 15) & 1;
    p.lean_tc_16 = (options >> 16) & 1;
    p.lean_cta_per_tile = (options >> 14) & 1;
    if ((options & 2) && !(options & 16)) {  // folded half-patch kernel for the antisymmetric output layer
        bool handled = false;
        rc = launch_cconv_apatch(p, st, &handled);
        if (rc || handled) return rc;
    }
    if (options & 2) {  // resident-filter direct kernel for cout <= 4
        bool handled = false;
        rc = launch_cconv_direct(p, st, &handled);
        if (rc || handled) return rc;
    }
    if ((options & 2) && !(options & 4096)) {  // direct kernel for narrow (block diagonal) layers: the input conv of every net
        bool handled = false;
        rc = launch_cconv_narrow(p, st, &handled);
        if (rc || handled) return rc;
    }
    if ((options & 1) && !(options & 8) && (options & 128)) {  // warp-specialised register-patch kernel (measured experiment, off by
                                                               // default: slower than k_cconv_lean, see cconv_ws.cu)
        bool handled = false;
        rc = launch_cconv_ws(p, st, &handled);
        if (rc || handled) return rc;
    }
    if ((options & 1) && !(options & 8)) {  // lean register-patch kernel (production path of the wide layers)
        bool handled = false;
        rc = launch_cconv_lean(p, st, &handled);
        if (rc || handled) return rc;
    }
    if (options & 1) {
        bool handled = false;
        rc = launch_cconv_wide(p, st, &handled);
        if (rc || handled) return rc;
    }
    // Largest tile that fits (bigger tile = fewer passes over the filter).  Small patches leave room for several
    // 8-warp CTAs per SM; a patch tile that owns the SM runs 16 warps to hide the gather / filter latency.
    if (conv_smem_bytes(32, 8, p.kc_pad, p.cp) <= limit / 2) return launch_cconv<32, 8>(p, st);
    if (conv_smem_bytes(32, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<32, 16>(p, st);
    if (conv_smem_bytes(24, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<24, 16>(p, st);
    if (conv_smem_bytes(16, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<16, 16>(p, st);
    if (conv_smem_bytes(8, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<8, 16>(p, st);
    return set_error(DMCF_ERR_UNSUPPORTED, "cconv: filter %dx%dx%dx%d needs %zu B of shared memory per 8 points (> %zu)",
                     p.gp.kz, p.gp.ky, p.gp.kx, d->cin, conv_smem_bytes(8, 16, p.kc_pad, p.cp), limit);
}

// Patch rows only: the trilinear "patch" of every out point, B[o][cell*cin + ci] = sum_n a_n w_cell(n,o) g(f_n)[ci], i.e. the
// matrix whose product with the flattened filter is the conv output.  This is what the gradient w.r.t. the filter needs
// (dW = B^T dOut, a plain GEMM), see dmcf_b200/autograd.py.
extern "C" int dmcf_cconv_patches(const dmcf_conv_desc* d, const float* out_positions, int64_t n_out, const float* inp_positions,
                                  const float* inp_features, int64_t inp_stride, int64_t n_inp, const float* inp_importance,
                                  const int32_t* neighbors_index, const int64_t* neighbors_row_splits,
                                  const float* neighbors_importance, const float* pair_records, int64_t n_pairs,
                                  float* patches, int64_t patch_stride, void* stream) {
    ConvParams p;
    int rc = fill_params(d, nullptr, out_positions, n_out, inp_positions, inp_features, inp_stride, n_inp, inp_importance,
                         neighbors_index, neighbors_row_splits, neighbors_importance, &p);
    if (rc) return rc;
    DMCF_REQUIRE(d->dense_cin == 0 && !d->normalize && !d->ascc, "cconv_patches: dense / normalize / ascc are not supported");
    if (n_out == 0) return DMCF_OK;
    DMCF_REQUIRE(out_positions && neighbors_row_splits && patches, "cconv_patches: NULL buffer");
    DMCF_REQUIRE(n_inp == 0 || (inp_features && (pair_records || (inp_positions && neighbors_index))), "cconv_patches: NULL input buffer");
    DMCF_REQUIRE(inp_stride >= d->cin && patch_stride >= p.kc_conv, "cconv_patches: row stride smaller than the row");
    p.records = pair_records; p.n_pairs = n_pairs;
    p.patch_out = patches; p.patch_stride = patch_stride;
    p.cout = 4;  // unused by phase 1; keeps the launch heuristics (shared-memory sizes) in their usual range
    p.cp = 4;
    cudaStream_t st = (cudaStream_t)stream;
    const int options = g_kernel_options.load(std::memory_order_relaxed);
    if ((options & 1) && !(options & 8)) {
        bool handled = false;
        rc = launch_cconv_lean(p, st, &handled);
        if (rc || handled) return rc;
    }
    const size_t limit = 227 * 1024;
    if (conv_smem_bytes(32, 8, p.kc_pad, p.cp) <= limit / 2) return launch_cconv<32, 8>(p, st);
    if (conv_smem_bytes(32, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<32, 16>(p, st);
    if (conv_smem_bytes(24, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<24, 16>(p, st);
    if (conv_smem_bytes(16, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<16, 16>(p, st);
    if (conv_smem_bytes(8, 16, p.kc_pad, p.cp) <= limit) return launch_cconv<8, 16>(p, st);
    return set_error(DMCF_ERR_UNSUPPORTED, "cconv_patches: filter %dx%dx%dx%d needs too much shared memory", p.gp.kz, p.gp.ky,
                     p.gp.kx, d->cin);
}

extern "C" int dmcf_set_kernel_options(int options) {
    return g_kernel_options.exchange(options, std::memory_order_relaxed);
}
