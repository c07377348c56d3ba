// Point-set ops of the reference's in-repo CUDA extensions, rebuilt for sm_100a:
//   farthest-point sampling   utils/tools/sampling.cu:125-190   (multi-scale sub-sampling with `voxel_size: null`,
//                                                                 utils/tools/losses.py:236-247, 274-282)
//   approx-match / match-cost utils/tools/tf_approxmatch.cu:27-160, 300-412 and the CPU twins tf_approxmatch.cpp:51-196
//                                                                (EMD metric of run_valid, pipelines/simulator.py:247-249)
//   nn-distance               utils/tools/nn_distance.cpp:47-70, nn_distance.cu
//
// What is different from the reference kernels (which run ONE thread block per batch item, i.e. one SM of 148 for the
// batch-of-one calls DMCF makes):
//   * FPS: a thread-block CLUSTER per batch item; each CTA keeps its slice of the points and their running minimum
//     distance in shared memory, the per-iteration arg-max travels through distributed shared memory and one
//     cluster barrier.  The winner is chosen with the reference's tie order (strided 512-thread scan + pairwise tree:
//     smaller k mod 512 first, then smaller k), so the sampled indices are the reference's.
//   * approx-match: the three passes of a level are three grid-wide launches, one thread per row / column with the
//     inner sum in the reference's sequential order (deterministic, same rounding sequence as the reference kernel),
//     the other point set staged through shared memory; an optional fused mode accumulates the match cost without ever
//     materialising the n x m match matrix.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dmcf {

// ------------------------------------------------------------------------------------------------------------
// farthest-point sampling
// ------------------------------------------------------------------------------------------------------------
static constexpr int kFpsThreads = 1024;
static constexpr int kFpsRefThreads = 512;  // the reference's block size: defines the tie order

// squared distance exactly as nvcc contracts the reference's expression
// (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1):  fma(dz,dz, fma(dx,dx, dy*dy))
__device__ __forceinline__ float fps_dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    return v;
}

// One cluster of CL CTAs per batch item.  Dynamic shared memory: float4 cache[cap] (x, y, z, min distance) of the
// first `cap` points of this CTA's slice; the rest of the slice (if any) uses global memory (`temp` for the distances).
template <int CL>
__global__ void __launch_bounds__(kFpsThreads, 1) k_fps(const float* __restrict__ points, int n, int m, float* __restrict__ temp,
                                                          int* __restrict__ idx_out, int cap) {
    extern __shared__ __align__(16) float4 cache[];
    __shared__ unsigned long long wbest[kFpsThreads / 32];
    __shared__ unsigned long long slot[2][CL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank = 0;
    if constexpr (CL > 1) rank = cg::this_cluster().block_rank();
    const int batch = blockIdx.x / CL;
    points += (size_t)batch * n * 3;
    temp += (size_t)batch * n;
    idx_out += (size_t)batch * m;
    const int per = (n + CL - 1) / CL;
    const int lo = min(n, (int)rank * per), hi = min(n, lo + per);
    const unsigned tie_cnt = (unsigned)((n + kFpsRefThreads - 1) / kFpsRefThreads);

    for (int k = lo + tid; k < hi; k += kFpsThreads) {
        if (k - lo < cap) cache[k - lo] = make_float4(points[3 * k], points[3 * k + 1], points[3 * k + 2], 1e38f);
        else temp[k] = 1e38f;
    }
    int old = 0;
    if (rank == 0 && tid == 0 && m > 0) idx_out[0] = 0;
    // every CTA of the cluster must be running before a sibling writes into its shared memory (first write: round j = 1)
    if constexpr (CL > 1) cg::this_cluster().sync(); else __syncthreads();
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(points + 3 * old), y1 = __ldg(points + 3 * old + 1), z1 = __ldg(points + 3 * old + 2);
        unsigned long long best = 0ull;
        for (int k = lo + tid; k < hi; k += kFpsThreads) {
            float x2, y2, z2, td;
            const bool cached = k - lo < cap;
            if (cached) {
                const float4 v = cache[k - lo];
                x2 = v.x; y2 = v.y; z2 = v.z; td = v.w;
            } else {
                x2 = points[3 * k]; y2 = points[3 * k + 1]; z2 = points[3 * k + 2]; td = temp[k];
            }
            const float d = fps_dist2(x2 - x1, y2 - y1, z2 - z1);
            const float d2 = fminf(d, td);
            if (d2 != td) {
                if (cached) cache[k - lo].w = d2; else temp[k] = d2;
            }
            // larger distance first; ties: smaller (k mod 512), then smaller k (the reference's scan + tree order)
            const unsigned tie = ((unsigned)k & (kFpsRefThreads - 1)) * tie_cnt + ((unsigned)k / kFpsRefThreads);
            const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(0xffffffffu - tie);
            best = key > best ? key : best;
        }
        best = warp_max_u64(best);
        if (lane == 0) wbest[warp] = best;
        __syncthreads();
        if (warp == 0) {
            unsigned long long b = warp_max_u64(wbest[lane]);
            if constexpr (CL > 1) {
                cg::cluster_group cluster = cg::this_cluster();
                if (lane < CL) *cluster.map_shared_rank(&slot[j & 1][rank], lane) = b;
            } else {
                if (lane == 0) slot[j & 1][0] = b;
            }
        }
        if constexpr (CL > 1) cg::this_cluster().sync(); else __syncthreads();
        unsigned long long g = slot[j & 1][0];
#pragma unroll
        for (int c = 1; c < CL; ++c) g = slot[j & 1][c] > g ? slot[j & 1][c] : g;
        if (g == 0ull) {
            old = 0;  // no point at all (n == 0 is rejected by the launcher): the reference's besti = 0
        } else {
            const unsigned tie = 0xffffffffu - (unsigned)(g & 0xffffffffull);
            old = (int)((tie % tie_cnt) * kFpsRefThreads + tie / tie_cnt);
        }
        if (rank == 0 && tid == 0) idx_out[j] = old;
    }
    // no CTA may exit while a sibling can still write into its shared memory
    if constexpr (CL > 1) cg::this_cluster().sync();
}

template <int CL>
static int launch_fps(const float* points, int b, int n, int m, float* temp, int* idx_out, cudaStream_t st) {
    const int per = (n + CL - 1) / CL;
    const int cap_max = (200 * 1024) / 16;
    const int cap = per < cap_max ? per : cap_max;
    const size_t smem = (size_t)cap * 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_fps<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_fps)");
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * CL), 1, 1);
    cfg.blockDim = dim3(kFpsThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_fps<CL>, points, n, m, temp, idx_out, cap);
    if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(k_fps)");
    DMCF_LAUNCH_CHECK("k_fps");
    return DMCF_OK;
}

// ------------------------------------------------------------------------------------------------------------
// approx-match (the auction-like soft assignment of the EMD metric)
// ------------------------------------------------------------------------------------------------------------
static constexpr int kAmTile = 1024;

// squared distance as nvcc contracts the reference's expression (same shape as in FPS)
__device__ __forceinline__ float am_dist2(float x1, float y1, float z1, float x2, float y2, float z2) {
    return fps_dist2(x2 - x1, y2 - y1, z2 - z1);
}

// PASS 1 (rows k of set 1):  ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level d2) remainR[l])
// PASS 3 (rows k of set 1):  w = exp(level d2) ratioL[k] ratioR[l];  match[l][k] += w;  remainL[k] = max(0, remainL[k] - sum_l w)
//                            (FUSED: cost_rows[k] += sum_l sqrt(d2) w instead of / in addition to the match update)
template <int PASS>
__global__ void __launch_bounds__(256) k_am_rows(const float* __restrict__ xyz1, int n, const float* __restrict__ xyz2, int m,
                                                   float level, float* __restrict__ remainL, const float* __restrict__ other,
                                                   float* __restrict__ ratioL, float* __restrict__ match, int64_t match_ld,
                                                   float* __restrict__ cost_rows) {
    __shared__ float4 buf[kAmTile];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (k < n) { x1 = xyz1[3 * k]; y1 = xyz1[3 * k + 1]; z1 = xyz1[3 * k + 2]; }
    float suml = PASS == 1 ? 1e-9f : 0.0f;
    float cost = 0.0f;
    const float rl = (PASS == 3 && k < n) ? ratioL[k] : 0.0f;
    for (int l0 = 0; l0 < m; l0 += kAmTile) {
        const int lend = min(m, l0 + kAmTile) - l0;
        for (int l = threadIdx.x; l < lend; l += blockDim.x)
            buf[l] = make_float4(xyz2[3 * (l0 + l)], xyz2[3 * (l0 + l) + 1], xyz2[3 * (l0 + l) + 2], other[l0 + l]);
        __syncthreads();
        if (k < n) {
            for (int l = 0; l < lend; ++l) {
                const float4 q = buf[l];
                const float dsq = am_dist2(x1, y1, z1, q.x, q.y, q.z);
                if (PASS == 1) {
                    const float w = __expf(level * dsq) * q.w;
                    suml += w;
                } else {
                    const float w = __expf(level * dsq) * rl * q.w;
                    if (match) match[(int64_t)(l0 + l) * match_ld + k] += w;
                    if (cost_rows) cost += sqrtf(dsq) * w;
                    suml += w;
                }
            }
        }
        __syncthreads();
    }
    if (k < n) {
        if (PASS == 1) {
            ratioL[k] = remainL[k] / suml;
        } else {
            remainL[k] = fmaxf(0.0f, remainL[k] - suml);
            if (cost_rows) cost_rows[k] += cost;
        }
    }
}

// PASS 2 (columns l of set 2): sumr = remainR[l] sum_k exp(level d2) ratioL[k];
//   ratioR[l] = min(remainR[l] / (sumr + 1e-9), 1) remainR[l];  remainR[l] = max(0, remainR[l] - sumr)
__global__ void __launch_bounds__(256) k_am_cols(const float* __restrict__ xyz1, int n, const float* __restrict__ xyz2, int m,
                                                   float level, const float* __restrict__ ratioL, float* __restrict__ remainR,
                                                   float* __restrict__ ratioR) {
    __shared__ float4 buf[kAmTile];
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    float x2 = 0.f, y2 = 0.f, z2 = 0.f;
    if (l < m) { x2 = xyz2[3 * l]; y2 = xyz2[3 * l + 1]; z2 = xyz2[3 * l + 2]; }
    float sumr = 0.0f;
    for (int k0 = 0; k0 < n; k0 += kAmTile) {
        const int kend = min(n, k0 + kAmTile) - k0;
        for (int k = threadIdx.x; k < kend; k += blockDim.x)
            buf[k] = make_float4(xyz1[3 * (k0 + k)], xyz1[3 * (k0 + k) + 1], xyz1[3 * (k0 + k) + 2], ratioL[k0 + k]);
        __syncthreads();
        if (l < m) {
            for (int k = 0; k < kend; ++k) {
                const float4 q = buf[k];
                const float w = __expf(level * am_dist2(q.x, q.y, q.z, x2, y2, z2)) * q.w;
                sumr += w;
            }
        }
        __syncthreads();
    }
    if (l < m) {
        const float rr = remainR[l];
        sumr *= rr;
        const float consumption = fminf(rr / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * rr;
        remainR[l] = fmaxf(0.0f, rr - sumr);
    }
}

__global__ void k_fill_f32(float* __restrict__ a, int64_t na, float va, float* __restrict__ b, int64_t nb, float vb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) a[i] = va;
    if (i < nb) b[i] = vb;
}

// deterministic sum of a float array into one float (double accumulation, fixed order): one block
__global__ void __launch_bounds__(1024) k_sum_f32(const float* __restrict__ v, int64_t n, const double* __restrict__ vd, int64_t nd,
                                                    float* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)v[i];
    for (int64_t i = threadIdx.x; i < nd; i += 1024) s += vd[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)sh[0];
}

// match cost (tf_approxmatch.cpp:177-196): sum_k sum_l sqrt(d2) match[l][k]; per-block partial sums in double
static constexpr int kMcChunk = 256;  // columns l per block
__global__ void __launch_bounds__(256) k_match_cost(const float* __restrict__ xyz1, int n, const float* __restrict__ xyz2, int m,
                                                      const float* __restrict__ match, int64_t match_ld, double* __restrict__ partial) {
    __shared__ float4 buf[kMcChunk];
    __shared__ double sh[256];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int l0 = blockIdx.y * kMcChunk;
    const int lend = min(m, l0 + kMcChunk) - l0;
    for (int l = threadIdx.x; l < lend; l += blockDim.x)
        buf[l] = make_float4(xyz2[3 * (l0 + l)], xyz2[3 * (l0 + l) + 1], xyz2[3 * (l0 + l) + 2], 0.0f);
    __syncthreads();
    double s = 0.0;
    if (k < n) {
        const float x1 = xyz1[3 * k], y1 = xyz1[3 * k + 1], z1 = xyz1[3 * k + 2];
        float acc = 0.0f;
        for (int l = 0; l < lend; ++l) {
            const float4 q = buf[l];
            acc += sqrtf(am_dist2(x1, y1, z1, q.x, q.y, q.z)) * match[(int64_t)(l0 + l) * match_ld + k];
        }
        s = (double)acc;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = sh[0];
}

// gradient of the match cost w.r.t. both point sets (tf_approxmatch.cpp:198-232, the match is a constant):
//   g = match[l][k] (x2_l - x1_k) / max(|x2_l - x1_k|, 1e-20);  grad1[k] = -sum_l g;  grad2[l] = sum_k g
// grad1: one thread per k (match reads coalesced over k); grad2: one warp per l, lanes over k, shuffle reduction.
__global__ void __launch_bounds__(256) k_match_cost_grad1(const float* __restrict__ xyz1, int n, const float* __restrict__ xyz2, int m,
                                                            const float* __restrict__ match, int64_t match_ld, float* __restrict__ grad1) {
    __shared__ float buf[3 * kAmTile];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (k < n) { x1 = xyz1[3 * k]; y1 = xyz1[3 * k + 1]; z1 = xyz1[3 * k + 2]; }
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int l0 = 0; l0 < m; l0 += kAmTile) {
        const int lend = min(m, l0 + kAmTile) - l0;
        for (int t = threadIdx.x; t < 3 * lend; t += blockDim.x) buf[t] = xyz2[(size_t)3 * l0 + t];
        __syncthreads();
        if (k < n) {
            for (int l = 0; l < lend; ++l) {
                const float dx = buf[3 * l] - x1, dy = buf[3 * l + 1] - y1, dz = buf[3 * l + 2] - z1;
                const float d = fmaxf(sqrtf(fps_dist2(dx, dy, dz)), 1e-20f);
                const float w = match[(int64_t)(l0 + l) * match_ld + k];
                gx -= w * (dx / d); gy -= w * (dy / d); gz -= w * (dz / d);
            }
        }
        __syncthreads();
    }
    if (k < n) { grad1[3 * k] = gx; grad1[3 * k + 1] = gy; grad1[3 * k + 2] = gz; }
}

__global__ void __launch_bounds__(256) k_match_cost_grad2(const float* __restrict__ xyz1, int n, const float* __restrict__ xyz2, int m,
                                                            const float* __restrict__ match, int64_t match_ld, float* __restrict__ grad2) {
    const int lane = threadIdx.x & 31;
    const int l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (l >= m) return;
    const float x2 = xyz2[3 * l], y2 = xyz2[3 * l + 1], z2 = xyz2[3 * l + 2];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int k = lane; k < n; k += 32) {
        const float dx = x2 - xyz1[3 * k], dy = y2 - xyz1[3 * k + 1], dz = z2 - xyz1[3 * k + 2];
        const float d = fmaxf(sqrtf(fps_dist2(dx, dy, dz)), 1e-20f);
        const float w = match[(int64_t)l * match_ld + k];
        sx += w * (dx / d); sy += w * (dy / d); sz += w * (dz / d);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, off);
        sy += __shfl_xor_sync(0xffffffffu, sy, off);
        sz += __shfl_xor_sync(0xffffffffu, sz, off);
    }
    if (lane == 0) { grad2[3 * l] = sx; grad2[3 * l + 1] = sy; grad2[3 * l + 2] = sz; }
}

// nearest neighbour in set 2 of every point of set 1 (utils/tools/nn_distance.cpp:47-70): squared distance
// (x*x + y*y) + z*z without contraction, first minimum wins
__global__ void __launch_bounds__(256) k_nn_distance(const float* __restrict__ xyz1, int n, const float* __restrict__ xyz2, int m,
                                                       float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ float buf[3 * kAmTile];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (j < n) { x1 = xyz1[3 * j]; y1 = xyz1[3 * j + 1]; z1 = xyz1[3 * j + 2]; }
    float best = 0.0f;
    int besti = 0;
    for (int k0 = 0; k0 < m; k0 += kAmTile) {
        const int kend = min(m, k0 + kAmTile) - k0;
        for (int t = threadIdx.x; t < 3 * kend; t += blockDim.x) buf[t] = xyz2[(size_t)3 * k0 + t];
        __syncthreads();
        if (j < n) {
            for (int k = 0; k < kend; ++k) {
                const float d = dist2_exact(buf[3 * k] - x1, buf[3 * k + 1] - y1, buf[3 * k + 2] - z1);
                if ((k0 + k == 0) || d < best) { best = d; besti = k0 + k; }
            }
        }
        __syncthreads();
    }
    if (j < n) { dist[j] = best; idx[j] = besti; }
}

static int am_threads(int rows) { return rows >= 148 * 256 ? 256 : (rows >= 148 * 128 ? 128 : 64); }

}  // namespace dmcf

using namespace dmcf;

extern "C" int dmcf_farthest_point_sample(const float* points, int32_t b, int32_t n, int32_t m, float* temp, int32_t* idx_out,
                                          int32_t cluster_size, void* stream) {
    DMCF_REQUIRE(b >= 0 && n >= 1 && m >= 0 && m <= n, "fps: need n >= 1 and 0 <= m <= n (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || m == 0) return DMCF_OK;
    DMCF_REQUIRE(points && temp && idx_out, "fps: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    if (cluster_size == 0) cluster_size = n > 4096 ? 8 : 1;
    switch (cluster_size) {
        case 1: return launch_fps<1>(points, b, n, m, temp, idx_out, st);
        case 2: return launch_fps<2>(points, b, n, m, temp, idx_out, st);
        case 4: return launch_fps<4>(points, b, n, m, temp, idx_out, st);
        case 8: return launch_fps<8>(points, b, n, m, temp, idx_out, st);
        default: return set_error(DMCF_ERR_INVALID, "fps: cluster_size must be 0 (auto), 1, 2, 4 or 8");
    }
}

extern "C" size_t dmcf_approx_match_workspace_bytes(int32_t n, int32_t m) {
    return (size_t)(3 * (int64_t)n + 2 * (int64_t)m + 4) * sizeof(float);
}

extern "C" int dmcf_approx_match(const float* xyz1, int32_t n, const float* xyz2, int32_t m, int32_t first_level,
                                 float* match, int64_t match_ld, float* cost_out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    DMCF_REQUIRE(n >= 1 && m >= 1, "approx_match: empty point set (n=%d m=%d)", n, m);
    DMCF_REQUIRE(first_level >= -1 && first_level <= 10, "approx_match: first_level %d out of range", first_level);
    DMCF_REQUIRE(xyz1 && xyz2 && workspace && (match || cost_out), "approx_match: NULL buffer");
    DMCF_REQUIRE(!match || match_ld >= n, "approx_match: match row stride smaller than n");
    DMCF_REQUIRE(workspace_bytes >= dmcf_approx_match_workspace_bytes(n, m), "approx_match: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* remainL = (float*)workspace;
    float* ratioL = remainL + n;
    float* cost_rows = ratioL + n;
    float* remainR = cost_rows + n;
    float* ratioR = remainR + m;
    // capacities (tf_approxmatch.cu:30-36): the smaller set may be matched max/min (integer division) times
    const float multiL = n >= m ? 1.0f : (float)(m / n), multiR = n >= m ? (float)(n / m) : 1.0f;
    const int64_t big = n > m ? n : m;
    k_fill_f32<<<(unsigned)ceil_div(big, 256), 256, 0, st>>>(remainL, n, multiL, remainR, m, multiR);
    DMCF_LAUNCH_CHECK("k_fill_f32");
    if (cost_out) {
        k_fill_f32<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(cost_rows, n, 0.0f, nullptr, 0, 0.0f);
        DMCF_LAUNCH_CHECK("k_fill_f32");
    }
    if (match) {
        cudaError_t e = cudaMemset2DAsync(match, (size_t)match_ld * 4, 0, (size_t)n * 4, (size_t)m, st);
        if (e != cudaSuccess) return check_cuda(e, "cudaMemset2DAsync(match)");
    }
    const int tr = am_threads(n), tc = am_threads(m);
    for (int j = first_level; j >= -2; --j) {
        const float level = j == -2 ? 0.0f : -powf(4.0f, (float)j);
        k_am_rows<1><<<(unsigned)ceil_div(n, tr), tr, 0, st>>>(xyz1, n, xyz2, m, level, remainL, remainR, ratioL, nullptr, 0, nullptr);
        DMCF_LAUNCH_CHECK("k_am_rows<1>");
        k_am_cols<<<(unsigned)ceil_div(m, tc), tc, 0, st>>>(xyz1, n, xyz2, m, level, ratioL, remainR, ratioR);
        DMCF_LAUNCH_CHECK("k_am_cols");
        k_am_rows<3><<<(unsigned)ceil_div(n, tr), tr, 0, st>>>(xyz1, n, xyz2, m, level, remainL, ratioR, ratioL, match, match_ld,
                                                                cost_out ? cost_rows : nullptr);
        DMCF_LAUNCH_CHECK("k_am_rows<3>");
    }
    if (cost_out) {
        k_sum_f32<<<1, 1024, 0, st>>>(cost_rows, n, nullptr, 0, cost_out);
        DMCF_LAUNCH_CHECK("k_sum_f32");
    }
    return DMCF_OK;
}

extern "C" size_t dmcf_match_cost_workspace_bytes(int32_t n, int32_t m) {
    return (size_t)(ceil_div(n, 256) * ceil_div(m, kMcChunk)) * sizeof(double);
}

extern "C" int dmcf_match_cost(const float* xyz1, int32_t n, const float* xyz2, int32_t m, const float* match, int64_t match_ld,
                               float* cost_out, void* workspace, size_t workspace_bytes, void* stream) {
    DMCF_REQUIRE(n >= 1 && m >= 1, "match_cost: empty point set (n=%d m=%d)", n, m);
    DMCF_REQUIRE(xyz1 && xyz2 && match && cost_out && workspace, "match_cost: NULL buffer");
    DMCF_REQUIRE(match_ld >= n, "match_cost: match row stride smaller than n");
    DMCF_REQUIRE(workspace_bytes >= dmcf_match_cost_workspace_bytes(n, m), "match_cost: workspace too small");
    DMCF_REQUIRE(((uintptr_t)workspace & 7) == 0, "match_cost: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)ceil_div(n, 256), (unsigned)ceil_div(m, kMcChunk));
    DMCF_REQUIRE(grid.y <= 65535, "match_cost: m too large");
    k_match_cost<<<grid, 256, 0, st>>>(xyz1, n, xyz2, m, match, match_ld, (double*)workspace);
    DMCF_LAUNCH_CHECK("k_match_cost");
    k_sum_f32<<<1, 1024, 0, st>>>(nullptr, 0, (const double*)workspace, (int64_t)grid.x * grid.y, cost_out);
    DMCF_LAUNCH_CHECK("k_sum_f32");
    return DMCF_OK;
}

extern "C" int dmcf_match_cost_grad(const float* xyz1, int32_t n, const float* xyz2, int32_t m, const float* match, int64_t match_ld,
                                    float* grad1, float* grad2, void* stream) {
    DMCF_REQUIRE(n >= 1 && m >= 1, "match_cost_grad: empty point set (n=%d m=%d)", n, m);
    DMCF_REQUIRE(xyz1 && xyz2 && match && grad1 && grad2, "match_cost_grad: NULL buffer");
    DMCF_REQUIRE(match_ld >= n, "match_cost_grad: match row stride smaller than n");
    cudaStream_t st = (cudaStream_t)stream;
    k_match_cost_grad1<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(xyz1, n, xyz2, m, match, match_ld, grad1);
    DMCF_LAUNCH_CHECK("k_match_cost_grad1");
    k_match_cost_grad2<<<(unsigned)ceil_div(m, 8), 256, 0, st>>>(xyz1, n, xyz2, m, match, match_ld, grad2);
    DMCF_LAUNCH_CHECK("k_match_cost_grad2");
    return DMCF_OK;
}

extern "C" int dmcf_nn_distance(const float* xyz1, int32_t n, const float* xyz2, int32_t m, float* dist, int32_t* idx, void* stream) {
    DMCF_REQUIRE(n >= 0 && m >= 1, "nn_distance: need m >= 1 (n=%d m=%d)", n, m);
    if (n == 0) return DMCF_OK;
    DMCF_REQUIRE(xyz1 && xyz2 && dist && idx, "nn_distance: NULL buffer");
    const int t = am_threads(n);
    k_nn_distance<<<(unsigned)ceil_div(n, t), t, 0, (cudaStream_t)stream>>>(xyz1, n, xyz2, m, dist, idx);
    DMCF_LAUNCH_CHECK("k_nn_distance");
    return DMCF_OK;
}
