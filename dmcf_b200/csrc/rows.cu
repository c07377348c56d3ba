// Row movers for capacity-sized point sets: the glue of the sync-free slab step (halo packing, ghost rows appended behind the
// owned rows, migration).  One launch replaces the arange / compare / select / index_copy chain a tensor library needs when
// the destination offset and the row count live in device memory.
#include "common.cuh"

namespace dmcf {

struct AppendParams {
    float* dst;
    int64_t dst_stride, dst_capacity, base_host;
    const int32_t* base_dev;
    const float* src;
    int64_t src_stride, n_src;
    const int32_t* src_count_dev;
    const int64_t* src_index;
    int width;
    int32_t* new_count_dev;
    int32_t* overflow;
};

// thread = (row i, column group): dst[base + i][:] = src[index ? index[i] : i][:]
template <int VEC>
__global__ void __launch_bounds__(256) k_rows_append(const AppendParams p) {
    const int64_t base = p.base_host + (p.base_dev ? (int64_t)__ldg(p.base_dev) : 0);
    int64_t count = p.n_src;
    if (p.src_count_dev) {
        const int64_t c = (int64_t)__ldg(p.src_count_dev);
        count = c < count ? (c < 0 ? 0 : c) : count;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (p.new_count_dev) *p.new_count_dev = (int32_t)(base + count < p.dst_capacity ? base + count : p.dst_capacity);
        if (p.overflow && base + count > p.dst_capacity) *p.overflow = 1;
    }
    const int groups = p.width / VEC;
    const int64_t total = count * groups;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / groups;
        const int g = (int)(t - i * groups);
        const int64_t d = base + i;
        if (d >= p.dst_capacity) continue;
        const int64_t r = p.src_index ? p.src_index[i] : i;
        if (VEC == 4) {
            *reinterpret_cast<float4*>(p.dst + d * p.dst_stride + 4 * g) =
                __ldg(reinterpret_cast<const float4*>(p.src + r * p.src_stride + 4 * g));
        } else {
            p.dst[d * p.dst_stride + g] = __ldg(p.src + r * p.src_stride + g);
        }
    }
}

}  // namespace dmcf

using namespace dmcf;

extern "C" int dmcf_rows_append(float* dst, int64_t dst_stride, int64_t dst_capacity, int64_t base_host, const int32_t* base_dev,
                                const float* src, int64_t src_stride, int64_t n_src, const int32_t* src_count_dev,
                                const int64_t* src_index, int32_t width, int32_t* new_count_dev, int32_t* overflow_flag,
                                void* stream) {
    DMCF_REQUIRE(width >= 1 && n_src >= 0 && dst_capacity >= 0 && base_host >= 0, "rows_append: bad shape");
    DMCF_REQUIRE(dst_stride >= width && src_stride >= width, "rows_append: row stride smaller than the row");
    DMCF_REQUIRE(dst && (n_src == 0 || src), "rows_append: NULL buffer");
    AppendParams p{dst, dst_stride, dst_capacity, base_host, base_dev, src, src_stride, n_src, src_count_dev, src_index,
                   width, new_count_dev, overflow_flag};
    const bool vec = (width % 4 == 0) && (dst_stride % 4 == 0) && (src_stride % 4 == 0) && (((uintptr_t)dst | (uintptr_t)src) & 15) == 0;
    const int64_t work = (n_src > 0 ? n_src : 1) * (vec ? width / 4 : width);
    int64_t blocks = ceil_div(work, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (vec)
        k_rows_append<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else
        k_rows_append<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    DMCF_LAUNCH_CHECK("k_rows_append");
    return DMCF_OK;
}
