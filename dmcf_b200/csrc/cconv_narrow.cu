// continuous_conv forward, direct variant for NARROW layers: at most 4 input channels and 8 output channels per filter block.
//
// This is the input layer of every DMCF net (models/pbf_model.py:375-411): fluid_convs ([1, v] -> ch), obs_convs ([1, n] -> ch)
// and the two Dense layers, which the fused step expresses as ONE conv over zero-padded [fluid | box] features with a block
// diagonal filter (dmcf_b200/models.py::_input_weights).  The register-patch kernel (k_cconv_lean, multi-pair phase 1) spent
// 4.3 ms on it at 1.06 M points / 32 M pairs: with ~30 pairs per point the per-point costs of the patch route dominate (slot
// reduction, a patch x filter product over K * 8 columns of which half are structural zeros), 1 990 warp instructions per point.
// With so few channels the direct form is cheaper than any patch:
//     out[o][co] += sum_corners w_c * sum_ci g(f_n)[ci] * F[cell_c][ci][co]          (8 x 5 FMA per pair and lane)
// lane = (pair slot, output channel): four pairs per step, each on a quarter warp whose 8 lanes own the 8 output channels of the
// pair's block; the filter is resident in shared memory as float4 over the input channels, [block][cell][co] -> one LDS.128 per
// corner, and the 8 lanes of a quarter warp read 128 contiguous bytes (conflict free by construction).  No patch, no
// read-modify-write, no second phase, no CTA barrier after the filter is staged.
// The block structure is a PROMISE of the caller (dmcf_conv_desc::block_cin / block_cout): input channels [0, ca) feed outputs
// [0, na) only, channels [ca, cin) feed outputs [na, na + nb) only, and every input row is zero in one of the two channel
// groups -- the row's group is read off the data (any non-zero in [ca, cin)), so ghost rows in any order work.  Without the
// promise the kernel takes plain layers with cin <= 4 and cout <= 8.
#include "cconv_common.cuh"

namespace dmcf {

namespace narrow {
static constexpr int kWarps = 8;
static constexpr int kRecWords = 12;  // {row, base cell byte offset, dx | dy << 16 (byte offsets), dz, w000 .. w111}
static constexpr int kCols = 8;       // output channels per block = lanes of a quarter warp
}  // namespace narrow

struct NarrowSet {  // one step's operands of this lane's pair slot
    int4 hd;
    float4 fa, fb;
};

template <bool BLOCKS, bool VEC>
__global__ void __launch_bounds__(narrow::kWarps * 32, 4) k_cconv_narrow(const ConvParams p, int K, int ca, int na, int nb) {
    using namespace narrow;
    extern __shared__ __align__(16) float smem[];
    const int n_blk = BLOCKS ? 2 : 1;
    float4* filt = reinterpret_cast<float4*>(smem);                       // [n_blk][K][kCols] float4 over the block's input channels
    float* dense = smem + (size_t)n_blk * K * kCols * 4;                  // [dense_cin][cout]
    float* scratch = dense + ((p.dense_cin * p.cout + 3) & ~3);           // [kWarps][32][kRecWords]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane >> 3, col = lane & 7;
    const int cb = p.cin - ca;
    // stage the filter, transposed to float4 over the input channels of a block (zero padded)
    for (int i = tid; i < n_blk * K * kCols; i += kWarps * 32) {
        const int c = i & 7, cell = (i >> 3) % K, t = (i >> 3) / K;
        const int c_in0 = t ? ca : 0, c_n = t ? cb : ca, co = (t ? na : 0) + c;
        const bool co_ok = c < (t ? nb : na);
        float v[4];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
            v[ci] = (co_ok && ci < c_n) ? __ldg(p.filters + ((int64_t)cell * p.cin + c_in0 + ci) * p.cout + co) : 0.0f;
        filt[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int i = tid; i < p.dense_cin * p.cout; i += kWarps * 32) dense[i] = __ldg(p.filters + (int64_t)p.kc_conv * p.cout + i);
    __syncthreads();

    float* rec = scratch + (size_t)warp * 32 * kRecWords;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int kx = p.gp.kx, kyx = p.gp.ky * p.gp.kx;
    const char* fbase = reinterpret_cast<const char*>(filt) + col * 16;
    const int blk_bytes = K * kCols * 16;

    const int64_t n_warps = (int64_t)gridDim.x * kWarps;
    const int64_t n_out = conv_n_out(p);
    for (int64_t o = (int64_t)blockIdx.x * kWarps + warp; o < n_out; o += n_warps) {
        const float ox = __ldg(p.out_pos + 3 * o), oy = __ldg(p.out_pos + 3 * o + 1), oz = __ldg(p.out_pos + 3 * o + 2);
        const int64_t rs = p.row_splits[o], re = p.row_splits[o + 1];
        float acc_a = 0.0f, acc_b = 0.0f;
        for (int64_t c0 = rs; c0 < re; c0 += 32) {
            const int64_t n = c0 + lane;
            const PairRec pr = pair_record(p, n, n < re, ox, oy, oz);
            const unsigned active = __ballot_sync(0xffffffffu, pr.row >= 0);
            const int cnt = __popc(active);
            __syncwarp();  // the previous chunk's records are no longer read
            if (pr.row >= 0) {
                const PairGeom& g = pr.g;
                const int x0 = g.i0 & 0xff, y0 = (g.i0 >> 8) & 0xff, z0 = (g.i0 >> 16) & 0xff;
                const int x1 = g.i1 & 0xff, y1 = (g.i1 >> 8) & 0xff, z1 = (g.i1 >> 16) & 0xff;
                const int base = (z0 * kyx + y0 * kx + x0) * (kCols * 16);
                const int dx = (x1 - x0) * (kCols * 16), dy = (y1 - y0) * kx * (kCols * 16), dz = (z1 - z0) * kyx * (kCols * 16);
                float* r = rec + __popc(active & lt_mask) * kRecWords;
                *reinterpret_cast<int4*>(r) = make_int4(pr.row, base, dx | (dy << 16), dz);
                *reinterpret_cast<float4*>(r + 4) =
                    make_float4(g.wx0 * g.wy0 * g.wz0, g.wx1 * g.wy0 * g.wz0, g.wx0 * g.wy1 * g.wz0, g.wx1 * g.wy1 * g.wz0);
                *reinterpret_cast<float4*>(r + 8) =
                    make_float4(g.wx0 * g.wy0 * g.wz1, g.wx1 * g.wy0 * g.wz1, g.wx0 * g.wy1 * g.wz1, g.wx1 * g.wy1 * g.wz1);
            }
            __syncwarp();
            // operands of step j for this lane's slot: the record header and the neighbour's feature row (both channel groups)
            auto fetch = [&](NarrowSet& s, int j) {
                const int r = j + slot;
                s.hd = make_int4(-1, 0, 0, 0);
                s.fa = make_float4(0.f, 0.f, 0.f, 0.f);
                s.fb = s.fa;
                if (r < cnt) {
                    s.hd = *reinterpret_cast<const int4*>(rec + r * kRecWords);
                    const float* row = p.inp_feat + (int64_t)s.hd.x * p.inp_stride;
                    if (VEC) {
                        s.fa = __ldg(reinterpret_cast<const float4*>(row));
                        if (BLOCKS) s.fb = __ldg(reinterpret_cast<const float4*>(row + 4));
                    } else {
                        s.fa.x = __ldg(row);
                        if (ca > 1) s.fa.y = __ldg(row + 1);
                        if (ca > 2) s.fa.z = __ldg(row + 2);
                        if (ca > 3) s.fa.w = __ldg(row + 3);
                        if (BLOCKS) {
                            s.fb.x = __ldg(row + ca);
                            if (cb > 1) s.fb.y = __ldg(row + ca + 1);
                            if (cb > 2) s.fb.z = __ldg(row + ca + 2);
                            if (cb > 3) s.fb.w = __ldg(row + ca + 3);
                        }
                    }
                }
            };
            auto apply = [&](const NarrowSet& s, int j) {
                const int r = j + slot;
                if (r >= cnt) return;
                const float4 wa = *reinterpret_cast<const float4*>(rec + r * kRecWords + 4);
                const float4 wb = *reinterpret_cast<const float4*>(rec + r * kRecWords + 8);
                bool is_b = false;
                float4 f = s.fa;
                if (BLOCKS) {
                    is_b = (s.fb.x != 0.0f) | (s.fb.y != 0.0f) | (s.fb.z != 0.0f) | (s.fb.w != 0.0f);
                    if (is_b) f = s.fb;
                }
                if (p.relu_input) {
                    f.x = fmaxf(f.x, 0.0f); f.y = fmaxf(f.y, 0.0f); f.z = fmaxf(f.z, 0.0f); f.w = fmaxf(f.w, 0.0f);
                }
                f.x *= p.feat_scale; f.y *= p.feat_scale; f.z *= p.feat_scale; f.w *= p.feat_scale;
                const char* fp = fbase + s.hd.y + (is_b ? blk_bytes : 0);
                const int dx = s.hd.z & 0xffff, dy = s.hd.z >> 16, dz = s.hd.w;
                const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                float sum = 0.0f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 F = *reinterpret_cast<const float4*>(fp + ((c & 1) ? dx : 0) + ((c & 2) ? dy : 0) + ((c & 4) ? dz : 0));
                    float t = f.x * F.x;
                    t = fmaf(f.y, F.y, t);
                    t = fmaf(f.z, F.z, t);
                    t = fmaf(f.w, F.w, t);
                    sum = fmaf(w[c], t, sum);
                }
                if (is_b) acc_b += sum; else acc_a += sum;
            };
            // two static operand sets: the rows of step j + 4 are in flight while step j is applied, and no register is moved
            // behind a load in flight
            NarrowSet s0, s1;
            fetch(s0, 0);
            for (int j = 0; j < cnt; j += 8) {
                fetch(s1, j + 4);
                apply(s0, j);
                fetch(s0, j + 8);
                apply(s1, j + 4);
            }
        }
        // the four pair slots of every output channel
        acc_a += __shfl_xor_sync(0xffffffffu, acc_a, 8);
        acc_a += __shfl_xor_sync(0xffffffffu, acc_a, 16);
        if (BLOCKS) {
            acc_b += __shfl_xor_sync(0xffffffffu, acc_b, 8);
            acc_b += __shfl_xor_sync(0xffffffffu, acc_b, 16);
        }
        // lane = output column: conv blocks, then fused Dense on the centre row, bias, residual
        const int src = lane < na ? lane : lane - na;
        const float va = __shfl_sync(0xffffffffu, acc_a, src & 7);
        const float vb = BLOCKS ? __shfl_sync(0xffffffffu, acc_b, src & 7) : 0.0f;
        float v = lane < na ? va : ((BLOCKS && lane < na + nb) ? vb : 0.0f);
        if (lane < p.cout) {
            for (int ci = 0; ci < p.dense_cin; ++ci) {
                float f = __ldg(p.dense_inp + o * p.dense_stride + ci);
                if (p.relu_input) f = fmaxf(f, 0.0f);
                v = fmaf(f, dense[ci * p.cout + lane], v);
            }
            if (p.bias) v += __ldg(p.bias + lane);
            if (p.residual) v += __ldg(p.residual + o * p.residual_stride + lane);
            float* dst = p.out + o * p.out_stride + lane;
            if (p.accumulate) v += *dst;
            *dst = v;
        }
    }
}

template <bool BLOCKS, bool VEC>
static int launch_narrow(const ConvParams& p, int K, int ca, int na, int nb, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_cconv_narrow<BLOCKS, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_cconv_narrow)");
        attr_set = true;
    }
    int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int64_t blocks = ceil_div(p.n_out, narrow::kWarps);
    if (blocks > 148 * ctas_per_sm) blocks = 148 * ctas_per_sm;  // persistent: the filter is staged once per CTA
    k_cconv_narrow<BLOCKS, VEC><<<(unsigned)blocks, narrow::kWarps * 32, smem, st>>>(p, K, ca, na, nb);
    DMCF_LAUNCH_CHECK("k_cconv_narrow");
    return DMCF_OK;
}

// Tries the narrow direct kernel; *handled = false means "not eligible".
int launch_cconv_narrow(const ConvParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p.ascc || p.normalize || p.patch_out || p.cout > 32 || p.dense_cin > 32 || p.nbr_hi > p.nbr_lo) return DMCF_OK;
    const bool blocks = p.blk_ca > 0;
    int ca, na, nb;
    if (blocks) {
        ca = p.blk_ca; na = p.blk_na; nb = p.blk_nb;
        const int cb = p.cin - ca;
        if (ca > 4 || cb < 1 || cb > 4 || na < 1 || nb < 1 || na > narrow::kCols || nb > narrow::kCols || na + nb > p.cout)
            return DMCF_OK;
    } else {
        if (p.cin > 4 || p.cout > narrow::kCols) return DMCF_OK;
        // a plain narrow layer pays 8 x 4 x 8 FMA per pair here and 8 x cin per pair (plus a per-point product) on the
        // register-patch route: direct evaluation wins while a point has few pairs
        if (p.n_pairs > 0 && p.n_pairs > 96 * p.n_out) return DMCF_OK;
        ca = p.cin; na = p.cout; nb = 0;
    }
    const int K = p.gp.kx * p.gp.ky * p.gp.kz;
    // byte offsets of a corner inside a block travel as 16-bit fields
    if ((int64_t)K * narrow::kCols * 16 >= 65536) return DMCF_OK;
    const size_t smem = ((size_t)(blocks ? 2 : 1) * K * narrow::kCols * 4 + ((p.dense_cin * p.cout + 3) & ~3) +
                         (size_t)narrow::kWarps * 32 * narrow::kRecWords) * sizeof(float);
    if (smem > 100 * 1024) return DMCF_OK;
    *handled = true;
    const bool vec = (blocks ? (ca == 4 && p.cin == 8) : p.cin == 4) && (p.inp_stride % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.inp_feat) & 15) == 0);
    if (blocks) return vec ? launch_narrow<true, true>(p, K, ca, na, nb, smem, st) : launch_narrow<true, false>(p, K, ca, na, nb, smem, st);
    return vec ? launch_narrow<false, true>(p, K, ca, na, nb, smem, st) : launch_narrow<false, false>(p, K, ca, na, nb, smem, st);
}

}  // namespace dmcf
