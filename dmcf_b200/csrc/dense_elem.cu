// Per-particle Dense layer and the elementwise pieces of one simulator step.
// Reference: tf.keras.layers.Dense at models/pbf_model.py:140-152, models/hrnet.py:63-66;
// integrate / correct at models/pbf_model.py:234-250, 466-487.  All HBM-bound.
#include "common.cuh"

namespace dmcf {

// out[n, co] = sum_ci g(x[n, ci]) W[ci, co] + b[co]; W (<= 96x96 floats) is staged in shared memory once per CTA,
// a warp handles one particle per iteration with lanes over output channels.
__global__ void __launch_bounds__(256) k_dense(const float* __restrict__ x, int64_t n, int cin, int64_t x_stride,
                                                 const float* __restrict__ w, const float* __restrict__ b, int cout, int relu_input,
                                                 float* __restrict__ out, int64_t out_stride) {
    extern __shared__ float sw[];  // [cin*cout] + per-warp x rows [8][cin]
    float* sx = sw + cin * cout;
    for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* myx = sx + warp * cin;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; r < n; r += n_warps) {
        for (int ci = lane; ci < cin; ci += 32) {
            float v = x[r * x_stride + ci];
            if (relu_input) v = fmaxf(v, 0.0f);
            myx[ci] = v;
        }
        __syncwarp();
        for (int co = lane; co < cout; co += 32) {
            float acc = b ? b[co] : 0.0f;
            for (int ci = 0; ci < cin; ++ci) acc = fmaf(myx[ci], sw[ci * cout + co], acc);
            out[r * out_stride + co] = acc;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_integrate(const float* __restrict__ pos, const float* __restrict__ vel,
                                                     const float* __restrict__ acc, float gx, float gy, float gz, float dt, int64_t n3,
                                                     float* __restrict__ pos2, float* __restrict__ vel2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const int c = (int)(i % 3);
    const float a = acc ? acc[i] : (c == 0 ? gx : (c == 1 ? gy : gz));
    // separately rounded multiply / add like the reference's elementwise TF ops (no FMA contraction)
    const float v2 = __fadd_rn(vel[i], __fmul_rn(dt, a));  // models/pbf_model.py:237-238
    vel2[i] = v2;
    pos2[i] = __fadd_rn(pos[i], __fmul_rn(dt, v2));  // :239
}

__global__ void __launch_bounds__(256) k_correct(const float* __restrict__ pos, const float* __restrict__ pos2,
                                                   const float* __restrict__ net, int64_t net_stride, int net_c, float sx, float sy,
                                                   float sz, float dt, int64_t n3, float* __restrict__ pos_new,
                                                   float* __restrict__ vel_new) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const int64_t r = i / 3;
    const int c = (int)(i % 3);
    // channel expansion of models/pbf_model.py:466-469: 1 -> repeat, 2 -> [a, b, a]
    const int src = (net_c == 1) ? 0 : ((net_c == 2 && c == 2) ? 0 : c);
    const float s = c == 0 ? sx : (c == 1 ? sy : sz);
    const float pn = __fadd_rn(pos2[i], __fmul_rn(s, net[r * net_stride + src]));  // :474, :248
    pos_new[i] = pn;
    vel_new[i] = __fdiv_rn(__fsub_rn(pn, pos[i]), dt);  // :249
}

}  // namespace dmcf

namespace dmcf {
int launch_dense_umma(const float* x, int64_t n, int cin, int64_t x_stride, const float* w, const float* b, int cout, int relu_input,
                      float* out, int64_t out_stride, cudaStream_t st, bool* handled);  // dense_umma.cu
extern std::atomic<int> g_kernel_options;                                               // cconv.cu
}  // namespace dmcf

using namespace dmcf;

extern "C" int dmcf_dense_forward(const float* x, int64_t n, int32_t cin, int64_t x_stride, const float* w, const float* b,
                                  int32_t cout, int32_t relu_input, float* out, int64_t out_stride, void* stream) {
    DMCF_REQUIRE(cin >= 1 && cout >= 1 && n >= 0, "dense: bad shape");
    DMCF_REQUIRE((size_t)(cin * cout + 8 * cin) * 4 <= 96 * 1024, "dense: kernel %dx%d too large", cin, cout);
    if (n == 0) return DMCF_OK;
    DMCF_REQUIRE(x && w && out, "dense: NULL buffer");
    DMCF_REQUIRE(x_stride >= cin && out_stride >= cout, "dense: row stride smaller than channel count");
    if (!(g_kernel_options.load(std::memory_order_relaxed) & 8192)) {  // tensor-core kernel (tcgen05, 3xTF32) for large batches
        bool handled = false;
        const int rc = launch_dense_umma(x, n, cin, x_stride, w, b, cout, relu_input, out, out_stride, (cudaStream_t)stream, &handled);
        if (rc || handled) return rc;
    }
    const size_t smem = (size_t)(cin * cout + 8 * cin) * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_dense)");
        attr_set = true;
    }
    int64_t blocks = ceil_div(n, 8 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_dense<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(x, n, cin, x_stride, w, b, cout, relu_input, out, out_stride);
    DMCF_LAUNCH_CHECK("k_dense");
    return DMCF_OK;
}

extern "C" int dmcf_integrate(const float* pos, const float* vel, const float* acc, const float* g, float dt, int64_t n,
                              float* pos2, float* vel2, void* stream) {
    DMCF_REQUIRE(n >= 0, "integrate: negative n");
    if (n == 0) return DMCF_OK;
    DMCF_REQUIRE(pos && vel && pos2 && vel2 && (acc || g), "integrate: NULL buffer");
    const float gx = g ? g[0] : 0.f, gy = g ? g[1] : 0.f, gz = g ? g[2] : 0.f;
    k_integrate<<<(unsigned)ceil_div(3 * n, 256), 256, 0, (cudaStream_t)stream>>>(pos, vel, acc, gx, gy, gz, dt, 3 * n, pos2, vel2);
    DMCF_LAUNCH_CHECK("k_integrate");
    return DMCF_OK;
}

extern "C" int dmcf_correct(const float* pos, const float* pos2, const float* net, int64_t net_stride, int32_t net_c,
                            const float* s, float dt, int64_t n, float* pos_new, float* vel_new, void* stream) {
    DMCF_REQUIRE(n >= 0 && net_c >= 1 && net_c <= 3, "correct: bad shape");
    if (n == 0) return DMCF_OK;
    DMCF_REQUIRE(pos && pos2 && net && s && pos_new && vel_new, "correct: NULL buffer");
    DMCF_REQUIRE(dt != 0.0f, "correct: dt is zero");
    k_correct<<<(unsigned)ceil_div(3 * n, 256), 256, 0, (cudaStream_t)stream>>>(pos, pos2, net, net_stride, net_c, s[0], s[1], s[2], dt,
                                                                                  3 * n, pos_new, vel_new);
    DMCF_LAUNCH_CHECK("k_correct");
    return DMCF_OK;
}
