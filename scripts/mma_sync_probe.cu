// Throughput of the LEGACY tensor path (mma.sync m16n8k8 tf32, SASS HMMA) on sm_100a: MACs per clock and SM for W warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/mma_sync_probe.cu -o scripts/bin/mma_sync_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k4(int iters, float* out, long long* cyc) {  // m16n8k4
    float c[6][4] = {};
    unsigned a[2][2], b[3];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) a[i][j] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i + j);
    for (int i = 0; i < 3; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 3; ++n)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[m * 3 + n][0]), "+f"(c[m * 3 + n][1]), "+f"(c[m * 3 + n][2]), "+f"(c[m * 3 + n][3])
                             : "r"(a[m][0]), "r"(a[m][1]), "r"(b[n]));
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k(int iters, float* out, long long* cyc) {
    float c[6][4] = {};
    unsigned a[2][4], b[3][2];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) a[i][j] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i + j);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) b[i][j] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i + j);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 3; ++n)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[m * 3 + n][0]), "+f"(c[m * 3 + n][1]), "+f"(c[m * 3 + n][2]), "+f"(c[m * 3 + n][3])
                             : "r"(a[m][0]), "r"(a[m][1]), "r"(a[m][2]), "r"(a[m][3]), "r"(b[n][0]), "r"(b[n][1]));
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 12, 16, 32}) {
        const int iters = 20000;
        k<<<148, warps * 32>>>(iters, out, cyc);
        cudaDeviceSynchronize();
        k<<<148, warps * 32>>>(iters, out, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        printf("m16n8k8 warps/SM %2d: %.1f cycles per 6 MMAs per warp, %.0f tf32 MAC/clk/SM\n", warps, c / iters, warps * 6.0 * 1024 * iters / c);
        k4<<<148, warps * 32>>>(iters, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        printf("m16n8k4 warps/SM %2d: %.1f cycles per 6 MMAs per warp, %.0f tf32 MAC/clk/SM\n", warps, c / iters, warps * 6.0 * 512 * iters / c);
    }
    return 0;
}
