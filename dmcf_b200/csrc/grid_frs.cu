// Cell list build + fixed-radius search (count / fill) + exclusive scans.
// Replaces open3d build_spatial_hash_table / fixed_radius_search as called from
// utils/convolutions.py:354-358 of the reference.  HBM-bound integer/byte work: coalesced float4 candidate
// reads from the cell-major copy of the points, warp-ballot compaction for ordered, coalesced CSR writes.
#include "common.cuh"

namespace dmcf {

static constexpr int kScanThreads = 512;
static constexpr int kScanItems = 4;
static constexpr int kScanTile = kScanThreads * kScanItems;

// ---------------------------------------------------------------------------------------------------------
// exclusive scan (int32 in -> T out), three small kernels; out has n+1 entries.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_local(const int32_t* __restrict__ in, int64_t n,
                                                               T* __restrict__ out, int64_t* __restrict__ tile_sums) {
    __shared__ int64_t warp_sums[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int64_t local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        local += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = (lane < kScanThreads / 32) ? warp_sums[lane] : 0;
        int64_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    int64_t run = warp_sums[warp] + incl - local;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = (T)run;
        run += v[i];
    }
}

// single block: exclusive scan of the tile sums in place
__global__ void __launch_bounds__(1024) k_scan_tiles(int64_t* __restrict__ tile_sums, int64_t n_tiles) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t start = 0; start < n_tiles; start += 1024) {
        int64_t i = start + threadIdx.x;
        int64_t v = (i < n_tiles) ? tile_sums[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sums[lane];
            int64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t excl = carry + warp_sums[warp] + incl - v;
        if (i < n_tiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_add(T* __restrict__ out, int64_t n,
                                                             const int64_t* __restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    const int64_t off = tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) out[base + i] = (T)(out[base + i] + off);
}

// total: out[n] = out[n-1] + in[n-1]
template <typename T>
__global__ void k_scan_total(const int32_t* __restrict__ in, int64_t n, T* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[n] = (n > 0) ? (T)(out[n - 1] + in[n - 1]) : (T)0;
}

template <typename T>
static int exclusive_scan(const int32_t* in, int64_t n, T* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    DMCF_REQUIRE(n >= 0, "scan: negative length");
    const int64_t n_tiles = ceil_div(n, kScanTile);
    if ((size_t)(n_tiles > 0 ? n_tiles : 1) * sizeof(int64_t) > ws_bytes)
        return set_error(DMCF_ERR_WORKSPACE, "scan: workspace %zu < %zu", ws_bytes, (size_t)n_tiles * 8);
    int64_t* tile_sums = (int64_t*)ws;
    if (n_tiles > 0) {
        k_scan_local<T><<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, n, out, tile_sums);
        DMCF_LAUNCH_CHECK("k_scan_local");
        if (n_tiles > 1) {
            k_scan_tiles<<<1, 1024, 0, st>>>(tile_sums, n_tiles);
            DMCF_LAUNCH_CHECK("k_scan_tiles");
            k_scan_add<T><<<(unsigned)n_tiles, kScanThreads, 0, st>>>(out, n, tile_sums);
            DMCF_LAUNCH_CHECK("k_scan_add");
        }
    }
    k_scan_total<T><<<1, 32, 0, st>>>(in, n, out);
    DMCF_LAUNCH_CHECK("k_scan_total");
    return DMCF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// cell list
// ---------------------------------------------------------------------------------------------------------
struct GridView {
    float ox, oy, oz, inv_cell;
    int nx, ny, nz, n_points;
    const int32_t* n_points_dev;  // optional device-side count (n_points is then the capacity)
    const float* points;          // the array the grid was built from (set by dmcf_grid_build)
    const int32_t* cell_start;
    const int32_t* sorted_index;
    const float4* sorted_pos;
};

__device__ __forceinline__ int cell_of_point(const GridView& g, float x, float y, float z) {
    const int cx = cell_coord(x, g.ox, g.inv_cell, g.nx);
    const int cy = cell_coord(y, g.oy, g.inv_cell, g.ny);
    const int cz = cell_coord(z, g.oz, g.inv_cell, g.nz);
    return (cz * g.ny + cy) * g.nx + cx;
}

__device__ __forceinline__ int grid_n_points(const GridView& g) {
    if (g.n_points_dev == nullptr) return g.n_points;
    const int n = __ldg(g.n_points_dev);
    return n < g.n_points ? (n < 0 ? 0 : n) : g.n_points;
}

__global__ void __launch_bounds__(256) k_cell_hist(GridView g, const float* __restrict__ pts, int32_t* __restrict__ cell_of,
                                                     int32_t* __restrict__ cell_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= grid_n_points(g)) return;
    const int c = cell_of_point(g, pts[3 * (int64_t)i], pts[3 * (int64_t)i + 1], pts[3 * (int64_t)i + 2]);
    cell_of[i] = c;
    atomicAdd(&cell_count[c], 1);
}

__global__ void __launch_bounds__(256) k_cell_scatter(GridView g, const int32_t* __restrict__ cell_of,
                                                        const int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_fill,
                                                        int32_t* __restrict__ sorted_index) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= grid_n_points(g)) return;
    const int c = cell_of[i];
    const int slot = atomicAdd(&cell_fill[c], 1);
    sorted_index[cell_start[c] + slot] = i;
}

// One WARP per cell: order the ids of the cell ascending so that the layout (and every neighbour row built from it) is
// deterministic and independent of atomic arrival order.  Up to 32 ids: bitonic network on registers; up to 1024 (the
// coarse cells of the multi-scale nets hold ~500 points): bitonic sort in the warp's shared-memory buffer; beyond that
// (clamped border cells of degenerate inputs) one lane runs a heap sort in place.
static constexpr int kCellSortWarps = 8, kCellSortMax = 1024;

__global__ void __launch_bounds__(kCellSortWarps * 32) k_cell_sort(int64_t n_cells, const int32_t* __restrict__ cell_start,
                                                                    int32_t* __restrict__ sorted_index) {
    __shared__ int32_t buf[kCellSortWarps][kCellSortMax];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * kCellSortWarps + warp;
    if (c >= n_cells) return;
    const int s = cell_start[c], e = cell_start[c + 1];
    const int m = e - s;
    if (m < 2) return;
    int32_t* a = sorted_index + s;
    if (m <= 32) {
        int32_t v = lane < m ? a[lane] : 0x7fffffff;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int32_t other = __shfl_xor_sync(0xffffffffu, v, j);
                const bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
                v = (up == lower) ? min(v, other) : max(v, other);
            }
        }
        if (lane < m) a[lane] = v;
        return;
    }
    if (m <= kCellSortMax) {
        int32_t* b = buf[warp];
        int P = 64;
        while (P < m) P <<= 1;
        for (int i = lane; i < P; i += 32) b[i] = i < m ? a[i] : 0x7fffffff;
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < P / 2; t += 32) {
                    // t-th compare-exchange of this step: partner indices i < l = i ^ j
                    const int i = 2 * t - (t & (j - 1));
                    const int l = i + j;
                    const int32_t x = b[i], y = b[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { b[i] = y; b[l] = x; }
                }
                __syncwarp();
            }
        }
        for (int i = lane; i < m; i += 32) a[i] = b[i];
        return;
    }
    if (lane != 0) return;
    // heap sort for crowded cells (clamped border cells, degenerate inputs)
    for (int start = m / 2 - 1; start >= 0; --start) {
        int root = start;
        const int32_t v = a[root];
        while (true) {
            int child = 2 * root + 1;
            if (child >= m) break;
            if (child + 1 < m && a[child + 1] > a[child]) ++child;
            if (a[child] <= v) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    }
    for (int end = m - 1; end > 0; --end) {
        const int32_t v = a[end];
        a[end] = a[0];
        int root = 0;
        while (true) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child + 1] > a[child]) ++child;
            if (a[child] <= v) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    }
}

__global__ void __launch_bounds__(256) k_cell_gather_pos(GridView g, const float* __restrict__ pts,
                                                           const int32_t* __restrict__ sorted_index, float4* __restrict__ sorted_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= grid_n_points(g)) return;
    const int id = sorted_index[i];
    sorted_pos[i] = make_float4(pts[3 * (int64_t)id], pts[3 * (int64_t)id + 1], pts[3 * (int64_t)id + 2], __int_as_float(id));
}

// ---------------------------------------------------------------------------------------------------------
// radius queries: one warp per query, lanes stride over the contiguous candidate run of each (z,y) row.
// ---------------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256) k_frs(GridView g, const float* __restrict__ queries, int64_t n_queries_cap,
                                               const int32_t* __restrict__ n_queries_dev, float radius, float thr,
                                               int ignore_query_point, const int64_t* __restrict__ row_splits, int64_t capacity,
                                               int32_t* __restrict__ counts, int32_t* __restrict__ nbr_index,
                                               float* __restrict__ nbr_dist, int32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int64_t n_queries = n_queries_cap;
    if (n_queries_dev != nullptr) {
        const int64_t nd = (int64_t)__ldg(n_queries_dev);
        n_queries = nd < n_queries_cap ? (nd < 0 ? 0 : nd) : n_queries_cap;
        if (!FILL)  // rows beyond the device-side count are empty (the scan runs over the whole capacity)
            for (int64_t q = n_queries + warp0 * 32 + lane; q < n_queries_cap; q += n_warps * 32) counts[q] = 0;
    }
    for (int64_t q = warp0; q < n_queries; q += n_warps) {
        const float qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
        const float rp = radius * 1.0001f;
        const float px = rp + fabsf(qx) * 1e-6f, py = rp + fabsf(qy) * 1e-6f, pz = rp + fabsf(qz) * 1e-6f;
        const int x0 = cell_coord(qx - px, g.ox, g.inv_cell, g.nx), x1 = cell_coord(qx + px, g.ox, g.inv_cell, g.nx);
        const int y0 = cell_coord(qy - py, g.oy, g.inv_cell, g.ny), y1 = cell_coord(qy + py, g.oy, g.inv_cell, g.ny);
        const int z0 = cell_coord(qz - pz, g.oz, g.inv_cell, g.nz), z1 = cell_coord(qz + pz, g.oz, g.inv_cell, g.nz);
        int count = 0;
        int64_t base = 0;
        if (FILL) base = row_splits[q];
        // candidate runs of the (z,y) cell rows: lane r fetches the bounds of row r (up to 32 rows: 3x3 when the cell edge
        // is the radius), so the loads are in flight together instead of one dependent pair per row
        const int ny_rows = y1 - y0 + 1, n_rows = (z1 - z0 + 1) * ny_rows;
        int rs_lane = 0, re_lane = 0;
        if (lane < n_rows && n_rows <= 32) {
            const int z = z0 + lane / ny_rows, y = y0 + lane % ny_rows;
            const int64_t crow = ((int64_t)z * g.ny + y) * g.nx;
            rs_lane = __ldg(g.cell_start + crow + x0);
            re_lane = __ldg(g.cell_start + crow + x1 + 1);
        }
        for (int r = 0; r < n_rows; ++r) {
            {
                int s, e;
                if (n_rows <= 32) {
                    s = __shfl_sync(0xffffffffu, rs_lane, r);
                    e = __shfl_sync(0xffffffffu, re_lane, r);
                } else {
                    const int z = z0 + r / ny_rows, y = y0 + r % ny_rows;
                    const int64_t crow = ((int64_t)z * g.ny + y) * g.nx;
                    s = __ldg(g.cell_start + crow + x0);
                    e = __ldg(g.cell_start + crow + x1 + 1);
                }
                for (int i0 = s; i0 < e; i0 += 32) {
                    const int i = i0 + lane;
                    bool hit = false;
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    float d2 = 0.f;
                    if (i < e) {
                        p = __ldg(g.sorted_pos + i);
                        d2 = dist2_exact(p.x - qx, p.y - qy, p.z - qz);
                        hit = d2 <= thr;
                        if (ignore_query_point && p.x == qx && p.y == qy && p.z == qz) hit = false;
                    }
                    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
                    if (FILL) {
                        if (hit) {
                            const int64_t pos = base + count + __popc(ballot & lt_mask);
                            if (pos < capacity) {
                                nbr_index[pos] = __float_as_int(p.w);
                                if (nbr_dist) nbr_dist[pos] = d2;
                            } else if (overflow) {
                                *overflow = 1;
                            }
                        }
                    }
                    count += __popc(ballot);
                }
            }
        }
        if (!FILL && lane == 0) counts[q] = count;
    }
}

// ---------------------------------------------------------------------------------------------------------
// radius queries, cell-centric variant for the common case "the queries are a prefix of the point set the grid was built from"
// (every same-set search of a DMCF step: [fluid | boundary] rows out, the same rows plus ghost rows in).
// k_frs above is bound by L2 traffic: every query warp streams its own ~3.4 KB of candidates (9 cell rows) although the ~8
// queries of a cell share them.  Here ONE WARP OWNS A CELL: its points ARE its queries (position and id come with the
// cell-major float4 list, no gather), the candidate rows of the cell's neighbourhood are loaded once, 32 candidates per batch
// (one per lane), and every batch is tested against all queries of the cell (query position broadcast by shuffles, hits
// compacted by ballot): 8x less candidate traffic, same row order (cell rows ascending, ascending position in the row).
// The neighbourhood is the union of the per-query cell ranges computed exactly like k_frs does; the extra candidates a query
// sees that way are rejected by the same exact distance test.
// ---------------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256) k_frs_cell(GridView g, int64_t n_queries_cap, const int32_t* __restrict__ n_queries_dev,
                                                    float radius, float thr, int ignore_query_point,
                                                    const int64_t* __restrict__ row_splits, int64_t capacity,
                                                    int32_t* __restrict__ counts, int32_t* __restrict__ nbr_index,
                                                    float* __restrict__ nbr_dist, int32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int64_t n_queries = n_queries_cap;
    if (n_queries_dev != nullptr) {
        const int64_t nd = (int64_t)__ldg(n_queries_dev);
        n_queries = nd < n_queries_cap ? (nd < 0 ? 0 : nd) : n_queries_cap;
    }
    // rows that are no query of any cell (beyond the device-side count, or beyond the points the grid holds) count zero
    if (!FILL) {
        const int n_pts = grid_n_points(g);
        const int64_t first_empty = n_queries < n_pts ? n_queries : (int64_t)n_pts;
        for (int64_t q = first_empty + warp0 * 32 + lane; q < n_queries_cap; q += n_warps * 32) counts[q] = 0;
    }
    const int64_t n_cells = (int64_t)g.nx * g.ny * g.nz;
    const float rp = radius * 1.0001f;
    for (int64_t cell = warp0; cell < n_cells; cell += n_warps) {
        const int cs = __ldg(g.cell_start + cell), ce = __ldg(g.cell_start + cell + 1);
        for (int qb = cs; qb < ce; qb += 32) {
            // ---- this lane's query (a point of the cell) ----
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            bool is_q = false;
            if (qb + lane < ce) {
                q = __ldg(g.sorted_pos + qb + lane);
                is_q = (int64_t)__float_as_int(q.w) < n_queries;
            }
            const unsigned qmask = __ballot_sync(0xffffffffu, is_q);
            if (qmask == 0u) continue;
            const int qid = __float_as_int(q.w);
            int x0 = 0x7fffffff, x1 = -1, y0 = 0x7fffffff, y1 = -1, z0 = 0x7fffffff, z1 = -1;
            if (is_q) {
                const float px = rp + fabsf(q.x) * 1e-6f, py = rp + fabsf(q.y) * 1e-6f, pz = rp + fabsf(q.z) * 1e-6f;
                x0 = cell_coord(q.x - px, g.ox, g.inv_cell, g.nx); x1 = cell_coord(q.x + px, g.ox, g.inv_cell, g.nx);
                y0 = cell_coord(q.y - py, g.oy, g.inv_cell, g.ny); y1 = cell_coord(q.y + py, g.oy, g.inv_cell, g.ny);
                z0 = cell_coord(q.z - pz, g.oz, g.inv_cell, g.nz); z1 = cell_coord(q.z + pz, g.oz, g.inv_cell, g.nz);
            }
            const int X0 = __reduce_min_sync(0xffffffffu, x0), X1 = __reduce_max_sync(0xffffffffu, x1);
            const int Y0 = __reduce_min_sync(0xffffffffu, y0), Y1 = __reduce_max_sync(0xffffffffu, y1);
            const int Z0 = __reduce_min_sync(0xffffffffu, z0), Z1 = __reduce_max_sync(0xffffffffu, z1);
            int count = 0;
            int64_t base = 0;
            if (FILL && is_q) base = row_splits[qid];
            for (int z = Z0; z <= Z1; ++z) {
                for (int y = Y0; y <= Y1; ++y) {
                    const int64_t crow = ((int64_t)z * g.ny + y) * g.nx;
                    const int s = __ldg(g.cell_start + crow + X0), e = __ldg(g.cell_start + crow + X1 + 1);
                    for (int i0 = s; i0 < e; i0 += 32) {
                        const bool c_ok = i0 + lane < e;
                        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c_ok) c = __ldg(g.sorted_pos + i0 + lane);
                        unsigned todo = qmask;
                        while (todo) {  // warp uniform
                            const int j = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const float qx = __shfl_sync(0xffffffffu, q.x, j), qy = __shfl_sync(0xffffffffu, q.y, j);
                            const float qz = __shfl_sync(0xffffffffu, q.z, j);
                            const float d2 = dist2_exact(c.x - qx, c.y - qy, c.z - qz);
                            bool hit = c_ok && d2 <= thr;
                            if (ignore_query_point && c.x == qx && c.y == qy && c.z == qz) hit = false;
                            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
                            if (ballot == 0u) continue;
                            if (FILL) {
                                const int64_t pos0 = __shfl_sync(0xffffffffu, base + count, j);
                                if (hit) {
                                    const int64_t pos = pos0 + __popc(ballot & lt_mask);
                                    if (pos < capacity) {
                                        nbr_index[pos] = __float_as_int(c.w);
                                        if (nbr_dist) nbr_dist[pos] = d2;
                                    } else if (overflow) {
                                        *overflow = 1;
                                    }
                                }
                            }
                            if (lane == j) count += __popc(ballot);
                        }
                    }
                }
            }
            if (!FILL && is_q) counts[qid] = count;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// grid_pos (utils/tools/losses.py:136-181)
// ---------------------------------------------------------------------------------------------------------
struct LatticeParams {
    float vx, vy, vz;     // voxel pitch
    float cx, cy, cz;     // centre (0 when not centralised)
    const float* c_dev;   // centre in device memory instead (sync-free steps); overrides cx, cy, cz
    float hx, hy, hz;     // hysteresis per axis (0 on inactive axes)
    int ax, ay, az;       // axis active (voxel >= 1e-5)
    int lox, loy, loz, dx, dy, dz;
    int centralize;
};

__device__ __forceinline__ int lattice_coord(float p, float c, float v, float h) {
    // floor(pos / max(v,1e-5) -/+ hyst) in float32, each operation rounded (utils/tools/losses.py:142-150)
    const float vm = fmaxf(v, 1e-5f);
    return (int)floorf(__fadd_rn(__fdiv_rn(__fsub_rn(p, c), vm), h));
}

__global__ void __launch_bounds__(256) k_grid_pos_mark(const float* __restrict__ pos, int64_t n, const int32_t* __restrict__ n_dev,
                                                         LatticeParams L, int32_t* __restrict__ flags, int32_t* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev != nullptr && i >= (int64_t)__ldg(n_dev))) return;
    if (L.c_dev != nullptr) { L.cx = __ldg(L.c_dev); L.cy = __ldg(L.c_dev + 1); L.cz = __ldg(L.c_dev + 2); }
    const float px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
        const float s = sgn ? 1.0f : -1.0f;
        const int bx = lattice_coord(px, L.cx, L.vx, s * L.hx) - L.lox;
        const int by = lattice_coord(py, L.cy, L.vy, s * L.hy) - L.loy;
        const int bz = lattice_coord(pz, L.cz, L.vz, s * L.hz) - L.loz;
        for (int oz = 0; oz <= L.az; ++oz)
            for (int oy = 0; oy <= L.ay; ++oy)
                for (int ox = 0; ox <= L.ax; ++ox) {
                    const int x = bx + ox, y = by + oy, z = bz + oz;
                    if (x >= 0 && x < L.dx && y >= 0 && y < L.dy && z >= 0 && z < L.dz)
                        flags[((int64_t)z * L.dy + y) * L.dx + x] = 1;
                    else if (overflow != nullptr)
                        *overflow = 1;  // the particle left the planned lattice bounds: the caller re-plans
                }
    }
}

__global__ void __launch_bounds__(256) k_grid_pos_emit(const int32_t* __restrict__ flags, const int32_t* __restrict__ offsets,
                                                         int64_t n_cells, LatticeParams L, float* __restrict__ out, int64_t capacity,
                                                         int32_t* __restrict__ overflow) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || !flags[c]) return;
    if (capacity >= 0 && (int64_t)offsets[c] >= capacity) {
        if (overflow != nullptr) *overflow = 1;
        return;
    }
    if (L.c_dev != nullptr) { L.cx = __ldg(L.c_dev); L.cy = __ldg(L.c_dev + 1); L.cz = __ldg(L.c_dev + 2); }
    const int x = (int)(c % L.dx) + L.lox;
    const int y = (int)((c / L.dx) % L.dy) + L.loy;
    const int z = (int)(c / ((int64_t)L.dx * L.dy)) + L.loz;
    const int64_t o = offsets[c];
    // gpos * voxel + (center | voxel/2), multiply and add rounded separately (:176-179)
    const float ax = L.centralize ? L.cx : L.vx / 2.0f;
    const float ay = L.centralize ? L.cy : L.vy / 2.0f;
    const float az = L.centralize ? L.cz : L.vz / 2.0f;
    out[3 * o] = __fadd_rn(__fmul_rn((float)x, L.vx), ax);
    out[3 * o + 1] = __fadd_rn(__fmul_rn((float)y, L.vy), ay);
    out[3 * o + 2] = __fadd_rn(__fmul_rn((float)z, L.vz), az);
}

static int make_lattice(const float* v, const float* c, const float* c_dev, float hyst, const int32_t* lo, const int32_t* dims,
                        LatticeParams* L) {
    DMCF_REQUIRE(v && lo && dims, "grid_pos: NULL parameter");
    DMCF_REQUIRE(!(c && c_dev), "grid_pos: give the centre in host OR device memory");
    L->c_dev = c_dev;
    DMCF_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "grid_pos: dims must be positive");
    DMCF_REQUIRE((int64_t)dims[0] * dims[1] * dims[2] < ((int64_t)1 << 31), "grid_pos: lattice too large");
    L->vx = v[0]; L->vy = v[1]; L->vz = v[2];
    L->centralize = c != nullptr || c_dev != nullptr;
    L->cx = c ? c[0] : 0.f; L->cy = c ? c[1] : 0.f; L->cz = c ? c[2] : 0.f;
    L->ax = v[0] >= 1e-5f; L->ay = v[1] >= 1e-5f; L->az = v[2] >= 1e-5f;
    L->hx = L->ax ? hyst : 0.f; L->hy = L->ay ? hyst : 0.f; L->hz = L->az ? hyst : 0.f;
    L->lox = lo[0]; L->loy = lo[1]; L->loz = lo[2];
    L->dx = dims[0]; L->dy = dims[1]; L->dz = dims[2];
    return DMCF_OK;
}

static int make_view(const dmcf_grid* grid, GridView* v) {
    DMCF_REQUIRE(grid != nullptr, "grid is NULL");
    DMCF_REQUIRE(grid->dims[0] > 0 && grid->dims[1] > 0 && grid->dims[2] > 0, "grid dims must be positive");
    DMCF_REQUIRE((int64_t)grid->dims[0] * grid->dims[1] * grid->dims[2] < (int64_t)1 << 31, "grid has too many cells");
    DMCF_REQUIRE(grid->inv_cell > 0.0f && grid->inv_cell == grid->inv_cell, "grid inv_cell must be positive");
    DMCF_REQUIRE(grid->n_points >= 0, "grid n_points negative");
    DMCF_REQUIRE(grid->cell_start && (grid->n_points == 0 || (grid->sorted_index && grid->sorted_pos)), "grid buffers are NULL");
    DMCF_REQUIRE(((uintptr_t)grid->sorted_pos & 15) == 0, "grid sorted_pos must be 16-byte aligned");
    v->ox = grid->origin[0]; v->oy = grid->origin[1]; v->oz = grid->origin[2];
    v->inv_cell = grid->inv_cell;
    v->nx = grid->dims[0]; v->ny = grid->dims[1]; v->nz = grid->dims[2];
    v->n_points = grid->n_points;
    v->n_points_dev = grid->n_points_dev;
    v->points = grid->points;
    v->cell_start = grid->cell_start;
    v->sorted_index = grid->sorted_index;
    v->sorted_pos = (const float4*)grid->sorted_pos;
    return DMCF_OK;
}

}  // namespace dmcf

using namespace dmcf;

extern "C" size_t dmcf_scan_workspace_bytes(int64_t n) {
    int64_t t = ceil_div(n > 0 ? n : 1, kScanTile);
    return align_up((size_t)t * sizeof(int64_t), 256);
}

extern "C" int dmcf_exclusive_scan_i32_i64(const int32_t* in, int64_t n, int64_t* out, void* ws, size_t ws_bytes, void* stream) {
    DMCF_REQUIRE(out && (n == 0 || in) && ws, "scan: NULL buffer");
    return exclusive_scan<int64_t>(in, n, out, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dmcf_exclusive_scan_i32_i32(const int32_t* in, int64_t n, int32_t* out, void* ws, size_t ws_bytes, void* stream) {
    DMCF_REQUIRE(out && (n == 0 || in) && ws, "scan: NULL buffer");
    return exclusive_scan<int32_t>(in, n, out, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t dmcf_grid_workspace_bytes(int64_t n_points, int64_t n_cells) {
    // cell_of[n] + cell_count[n_cells] + scan tiles
    return align_up((size_t)(n_points > 0 ? n_points : 1) * 4, 256) + align_up((size_t)(n_cells > 0 ? n_cells : 1) * 4, 256) +
           dmcf_scan_workspace_bytes(n_cells);
}

extern "C" int dmcf_grid_build(const float* points, dmcf_grid* grid, void* workspace, size_t workspace_bytes, void* stream) {
    GridView g;
    int rc = make_view(grid, &g);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = grid->n_points;
    const int64_t n_cells = (int64_t)g.nx * g.ny * g.nz;
    grid->points = points;
    DMCF_REQUIRE(n == 0 || points, "grid_build: points is NULL");
    DMCF_REQUIRE(workspace != nullptr, "grid_build: workspace is NULL");
    if (workspace_bytes < dmcf_grid_workspace_bytes(n, n_cells))
        return set_error(DMCF_ERR_WORKSPACE, "grid_build: workspace %zu < %zu", workspace_bytes, dmcf_grid_workspace_bytes(n, n_cells));
    char* ws = (char*)workspace;
    int32_t* cell_of = (int32_t*)ws;
    ws += align_up((size_t)(n > 0 ? n : 1) * 4, 256);
    int32_t* cell_count = (int32_t*)ws;
    ws += align_up((size_t)n_cells * 4, 256);
    void* scan_ws = ws;
    rc = check_cuda(cudaMemsetAsync(cell_count, 0, (size_t)n_cells * 4, st), "memset cell_count");
    if (rc) return rc;
    if (n > 0) {
        k_cell_hist<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(g, points, cell_of, cell_count);
        DMCF_LAUNCH_CHECK("k_cell_hist");
    }
    rc = exclusive_scan<int32_t>(cell_count, n_cells, grid->cell_start, scan_ws, dmcf_scan_workspace_bytes(n_cells), st);
    if (rc) return rc;
    if (n > 0) {
        rc = check_cuda(cudaMemsetAsync(cell_count, 0, (size_t)n_cells * 4, st), "memset cell_fill");
        if (rc) return rc;
        k_cell_scatter<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(g, cell_of, grid->cell_start, cell_count, grid->sorted_index);
        DMCF_LAUNCH_CHECK("k_cell_scatter");
        k_cell_sort<<<(unsigned)ceil_div(n_cells, kCellSortWarps), kCellSortWarps * 32, 0, st>>>(n_cells, grid->cell_start, grid->sorted_index);
        DMCF_LAUNCH_CHECK("k_cell_sort");
        k_cell_gather_pos<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(g, points, grid->sorted_index, (float4*)grid->sorted_pos);
        DMCF_LAUNCH_CHECK("k_cell_gather_pos");
    }
    return DMCF_OK;
}

namespace dmcf {
extern std::atomic<int> g_kernel_options;  // cconv.cu; bit 6 here: keep the query-centric k_frs for prefix searches (A/B switch)
}

// queries == the points the grid was built from (a prefix of them): the cell-centric kernel applies
// ... and pays when a cell holds few points: the kernel saves candidate TRAFFIC (one load per cell instead of one per query) but
// spends more instructions per (query, candidate batch) test, and a crowded cell is one warp's serial work.  Measured: 8 points
// per cell (C4, r = 2 spacings) count + fill 1.51 -> 1.10 ms; ~64 per cell (the coarse scales of Liquid3d) 3.8 -> 5.4 ms.
static bool frs_prefix_case(const dmcf_grid* grid, const float* queries, int64_t n_queries) {
    const int64_t n_cells = (int64_t)grid->dims[0] * grid->dims[1] * grid->dims[2];
    // points per OCCUPIED cell as the caller estimates it from the points' bounding box (the grid itself may be padded), else
    // points per cell of the grid
    const float occupancy = grid->mean_occupancy > 0.0f ? grid->mean_occupancy : (float)grid->n_points / (float)(n_cells > 0 ? n_cells : 1);
    return grid->points != nullptr && queries == grid->points && n_queries <= (int64_t)grid->n_points && occupancy <= 12.0f &&
           !(g_kernel_options.load(std::memory_order_relaxed) & 64);
}

static unsigned frs_cell_blocks(const GridView& g) {
    const int64_t want = ceil_div((int64_t)g.nx * g.ny * g.nz, 8);
    const int64_t cap = 148 * 32;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

static unsigned frs_blocks(int64_t n_queries) {
    int64_t want = ceil_div(n_queries, 8);  // 8 warps per block
    int64_t cap = 148 * 32;                 // persistent-ish: a multiple of the SM count
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

extern "C" int dmcf_frs_count(const dmcf_grid* grid, const float* queries, int64_t n_queries, const int32_t* n_queries_dev, float radius,
                              int ignore_query_point, int32_t* counts, void* stream) {
    GridView g;
    int rc = make_view(grid, &g);
    if (rc) return rc;
    DMCF_REQUIRE(n_queries >= 0 && (n_queries == 0 || (queries && counts)), "frs_count: NULL buffer");
    DMCF_REQUIRE(radius >= 0.0f, "frs_count: negative radius");
    if (n_queries == 0) return DMCF_OK;
    if (frs_prefix_case(grid, queries, n_queries)) {
        k_frs_cell<false><<<frs_cell_blocks(g), 256, 0, (cudaStream_t)stream>>>(g, n_queries, n_queries_dev, radius, radius * radius,
                                                                                   ignore_query_point, nullptr, 0, counts, nullptr, nullptr, nullptr);
        DMCF_LAUNCH_CHECK("k_frs_cell<count>");
        return DMCF_OK;
    }
    k_frs<false><<<frs_blocks(n_queries), 256, 0, (cudaStream_t)stream>>>(g, queries, n_queries, n_queries_dev, radius, radius * radius,
                                                                            ignore_query_point, nullptr, 0, counts, nullptr, nullptr, nullptr);
    DMCF_LAUNCH_CHECK("k_frs<count>");
    return DMCF_OK;
}

extern "C" int dmcf_frs_fill(const dmcf_grid* grid, const float* queries, int64_t n_queries, const int32_t* n_queries_dev, float radius,
                             int ignore_query_point, const int64_t* row_splits, int64_t capacity,
                             int32_t* neighbors_index, float* neighbors_distance, int32_t* overflow_flag, void* stream) {
    GridView g;
    int rc = make_view(grid, &g);
    if (rc) return rc;
    DMCF_REQUIRE(n_queries >= 0 && (n_queries == 0 || (queries && row_splits)), "frs_fill: NULL buffer");
    DMCF_REQUIRE(capacity >= 0 && (capacity == 0 || neighbors_index), "frs_fill: NULL neighbors_index");
    if (n_queries == 0) return DMCF_OK;
    if (frs_prefix_case(grid, queries, n_queries)) {
        k_frs_cell<true><<<frs_cell_blocks(g), 256, 0, (cudaStream_t)stream>>>(g, n_queries, n_queries_dev, radius, radius * radius,
                                                                                  ignore_query_point, row_splits, capacity, nullptr,
                                                                                  neighbors_index, neighbors_distance, overflow_flag);
        DMCF_LAUNCH_CHECK("k_frs_cell<fill>");
        return DMCF_OK;
    }
    k_frs<true><<<frs_blocks(n_queries), 256, 0, (cudaStream_t)stream>>>(g, queries, n_queries, n_queries_dev, radius, radius * radius,
                                                                           ignore_query_point, row_splits, capacity, nullptr,
                                                                           neighbors_index, neighbors_distance, overflow_flag);
    DMCF_LAUNCH_CHECK("k_frs<fill>");
    return DMCF_OK;
}

extern "C" int dmcf_grid_pos_mark(const float* pos, int64_t n, const int32_t* n_dev, const float* voxel, const float* center,
                                  const float* center_dev, float hyst, const int32_t* lo, const int32_t* dims, int32_t* flags,
                                  int32_t* overflow_flag, void* stream) {
    LatticeParams L;
    int rc = make_lattice(voxel, center, center_dev, hyst, lo, dims, &L);
    if (rc) return rc;
    DMCF_REQUIRE(n >= 0 && flags && (n == 0 || pos), "grid_pos_mark: NULL buffer");
    if (n == 0) return DMCF_OK;
    k_grid_pos_mark<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(pos, n, n_dev, L, flags, overflow_flag);
    DMCF_LAUNCH_CHECK("k_grid_pos_mark");
    return DMCF_OK;
}

extern "C" int dmcf_grid_pos_emit(const int32_t* flags, const int32_t* offsets, const float* voxel, const float* center,
                                  const float* center_dev, const int32_t* lo, const int32_t* dims, float* out, int64_t capacity,
                                  int32_t* overflow_flag, void* stream) {
    LatticeParams L;
    int rc = make_lattice(voxel, center, center_dev, 0.f, lo, dims, &L);
    if (rc) return rc;
    DMCF_REQUIRE(flags && offsets && out, "grid_pos_emit: NULL buffer");
    const int64_t n_cells = (int64_t)L.dx * L.dy * L.dz;
    k_grid_pos_emit<<<(unsigned)ceil_div(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(flags, offsets, n_cells, L, out, capacity,
                                                                                          overflow_flag);
    DMCF_LAUNCH_CHECK("k_grid_pos_emit");
    return DMCF_OK;
}
