// Shared helpers for libdmcf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "dmcf_b200.h"

namespace dmcf {

// thread-local error message, read through dmcf_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return DMCF_OK;
    return set_error(DMCF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define DMCF_LAUNCH_CHECK(name)                                  \
    do {                                                         \
        ::dmcf::g_launches.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t _e = cudaGetLastError();                     \
        if (_e != cudaSuccess) return ::dmcf::check_cuda(_e, name); \
    } while (0)

#define DMCF_REQUIRE(cond, ...)                                            \
    do {                                                                   \
        if (!(cond)) return ::dmcf::set_error(DMCF_ERR_INVALID, __VA_ARGS__); \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Squared distance in the operation order the oracle pins: (dx*dx + dy*dy) + dz*dz, every
// operation rounded to nearest float32, never contracted into an FMA.
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Cell coordinate of a position along one axis: monotone in x, clamped into the grid.
__device__ __forceinline__ int cell_coord(float x, float origin, float inv_cell, int dim) {
    float f = floorf(__fmul_rn(__fsub_rn(x, origin), inv_cell));
    f = fminf(fmaxf(f, 0.0f), (float)(dim - 1));
    return (int)f;
}

}  // namespace dmcf
