/* O32 -- plain C / OpenMP float32 restatement of the two native ops on DMCF's hot path (TEST INFRASTRUCTURE AND
 * TIMED CPU BASELINE ONLY; never linked into or called from the product library).
 *
 * PARITY UNPINNED: restates the published CPU algorithms of open3d 0.15.2 (pip wheel pinned by the reference's
 * requirements.txt:2, not vendored, not installable here):
 *   - FixedRadiusSearch: spatial hash with cell edge 2r, hash (x*73856093 ^ y*19349663 ^ z*83492791) mod table,
 *     count / prefix / fill build, queries visit the bins of the <= 8 cells touched by q +- r   
 *   - ContinuousConv CPU: per out point accumulate the trilinear patch B[cell][ci] += a s w_c f, then
 *     out = B(K*Cin) x filter(K*Cin, Cout), blocks of out points in parallel                     
 * Call sites in the reference: utils/convolutions.py:354-358 (search), :431/:454/:1054 (conv).
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int o32_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void o32_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* d2 = (dx*dx + dy*dy) + dz*dz, every op rounded to float (compile with -ffp-contract=off) */
static inline float dist2(const float* p, const float* q) {
    volatile float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    volatile float s = xx + yy;
    return s + zz;
}

static inline int64_t cell_of(float v, float inv_cell) { return (int64_t)floorf(v * inv_cell); }

static inline uint64_t hash_cell(int64_t x, int64_t y, int64_t z, uint64_t table) {
    return (((uint64_t)x * 73856093ull) ^ ((uint64_t)y * 19349663ull) ^ ((uint64_t)z * 83492791ull)) % table;
}

typedef struct {
    uint64_t table;
    float inv_cell;
    int64_t* bin_start; /* [table+1] */
    int32_t* bin_index; /* [n] point ids ascending inside a bin */
} o32_hash;

static void hash_build(const float* pts, int64_t n, float radius, o32_hash* h) {
    h->table = (uint64_t)(n / 64 > 1 ? n / 64 : 1); /* hash_table_size_factor = 1/64 */
    if (h->table < 1024 && n > 1024) h->table = 1024;
    /* cell edge a hair above 2r so that the widened +-r corners below never straddle three cells */
    h->inv_cell = 1.0f / (2.02f * radius);
    h->bin_start = (int64_t*)calloc(h->table + 1, sizeof(int64_t));
    h->bin_index = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    uint64_t* bin_of = (uint64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(uint64_t));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        bin_of[i] = hash_cell(cell_of(pts[3 * i], h->inv_cell), cell_of(pts[3 * i + 1], h->inv_cell),
                              cell_of(pts[3 * i + 2], h->inv_cell), h->table);
    for (int64_t i = 0; i < n; ++i) h->bin_start[bin_of[i] + 1]++;
    for (uint64_t b = 0; b < h->table; ++b) h->bin_start[b + 1] += h->bin_start[b];
    int64_t* fill = (int64_t*)malloc(h->table * sizeof(int64_t));
    memcpy(fill, h->bin_start, h->table * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) h->bin_index[fill[bin_of[i]]++] = (int32_t)i;
    free(fill);
    free(bin_of);
}

static void hash_free(o32_hash* h) {
    free(h->bin_start);
    free(h->bin_index);
}

/* visits every point of the (deduplicated) bins touched by q +- r; returns count, optionally writes */
static inline int64_t query_one(const o32_hash* h, const float* pts, const float* q, float radius, float thr, int ignore,
                                int32_t* idx_out, float* dist_out) {
    uint64_t bins[8];
    int nb = 0;
    for (int dz = -1; dz <= 1; dz += 2)
        for (int dy = -1; dy <= 1; dy += 2)
            for (int dx = -1; dx <= 1; dx += 2) {
                /* slightly widened corner so float rounding of the cell coordinate cannot miss a neighbour */
                const float pad = radius * 1.0001f;
                const float cx = q[0] + dx * (pad + fabsf(q[0]) * 1e-6f), cy = q[1] + dy * (pad + fabsf(q[1]) * 1e-6f),
                            cz = q[2] + dz * (pad + fabsf(q[2]) * 1e-6f);
                const uint64_t b = hash_cell(cell_of(cx, h->inv_cell), cell_of(cy, h->inv_cell), cell_of(cz, h->inv_cell), h->table);
                int seen = 0;
                for (int k = 0; k < nb; ++k) seen |= bins[k] == b;
                if (!seen) bins[nb++] = b;
            }
    /* ascending bin order => deterministic row order */
    for (int i = 1; i < nb; ++i) {
        uint64_t v = bins[i];
        int j = i - 1;
        while (j >= 0 && bins[j] > v) {
            bins[j + 1] = bins[j];
            --j;
        }
        bins[j + 1] = v;
    }
    int64_t count = 0;
    for (int k = 0; k < nb; ++k)
        for (int64_t s = h->bin_start[bins[k]]; s < h->bin_start[bins[k] + 1]; ++s) {
            const int32_t i = h->bin_index[s];
            const float* p = pts + 3 * (int64_t)i;
            const float d2 = dist2(p, q);
            if (d2 <= thr) {
                if (ignore && p[0] == q[0] && p[1] == q[1] && p[2] == q[2]) continue;
                if (idx_out) {
                    idx_out[count] = i;
                    if (dist_out) dist_out[count] = d2;
                }
                ++count;
            }
        }
    return count;
}

/* Phase 1: row_splits[nq+1].  Returns an opaque handle to pass to o32_frs_fill (which frees it). */
void* o32_frs_count(const float* pts, int64_t n, const float* queries, int64_t nq, float radius, int ignore, int64_t* row_splits) {
    o32_hash* h = (o32_hash*)malloc(sizeof(o32_hash));
    hash_build(pts, n, radius, h);
    volatile float thr_v = radius * radius;
    const float thr = thr_v;
    row_splits[0] = 0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t q = 0; q < nq; ++q) row_splits[q + 1] = query_one(h, pts, queries + 3 * q, radius, thr, ignore, NULL, NULL);
    for (int64_t q = 0; q < nq; ++q) row_splits[q + 1] += row_splits[q];
    return h;
}

void o32_frs_fill(void* handle, const float* pts, const float* queries, int64_t nq, float radius, int ignore,
                  const int64_t* row_splits, int32_t* index, float* dist) {
    o32_hash* h = (o32_hash*)handle;
    volatile float thr_v = radius * radius;
    const float thr = thr_v;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t q = 0; q < nq; ++q)
        query_one(h, pts, queries + 3 * q, radius, thr, ignore, index + row_splits[q], dist ? dist + row_splits[q] : NULL);
    hash_free(h);
    free(h);
}

/* ------------------------------------------------------------------------------------------------------ */
static inline void map_to_cube(int mapping, float inv_extent, float* px, float* py, float* pz) {
    float x = *px, y = *py, z = *pz;
    if (mapping == 0) {
        x *= inv_extent; y *= inv_extent; z *= inv_extent;
    } else {
        const float s2 = 2.0f * inv_extent;
        x *= s2; y *= s2; z *= s2;
        if (mapping == 1) {
            const float rad = sqrtf(x * x + y * y + z * z);
            const float amax = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
            if (amax < 1e-8f) {
                x = y = z = 0.0f;
            } else {
                const float s = 0.5f * rad / amax;
                x *= s; y *= s; z *= s;
            }
        } else {
            const float xy2 = x * x + y * y;
            const float sq = xy2 + z * z;
            const float n = sqrtf(sq);
            if (sq < 1e-12f) {
                x = y = z = 0.0f;
            } else if (1.25f * z * z > xy2) {
                const float s = sqrtf(3.0f * n / (n + fabsf(z)));
                x *= s; y *= s;
                z = copysignf(n, z);
            } else {
                const float s = n / sqrtf(xy2);
                x *= s; y *= s;
                z *= 1.5f;
            }
            const float sq2 = x * x + y * y;
            const float m = sqrtf(sq2);
            const float four_over_pi = 1.27323954473516f;
            if (sq2 < 1e-12f) {
                x = y = 0.0f;
            } else if (fabsf(y) <= fabsf(x)) {
                const float t = copysignf(m, x);
                y = t * four_over_pi * atanf(y / x);
                x = t;
            } else {
                const float t = copysignf(m, y);
                x = t * four_over_pi * atanf(x / y);
                y = t;
            }
            x *= 0.5f; y *= 0.5f; z *= 0.5f;
        }
    }
    *px = x; *py = y; *pz = z;
}

static inline void axis_w(int interp, float g, int fs, int* i0, int* i1, float* w0, float* w1) {
    if (interp == 0) {
        int a0 = (int)g;
        a0 = a0 < 0 ? 0 : (a0 > fs - 1 ? fs - 1 : a0);
        int a1 = a0 + 1 > fs - 1 ? fs - 1 : a0 + 1;
        float a = g - (float)a0;
        a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);
        *i0 = a0; *i1 = a1; *w0 = 1.0f - a; *w1 = a;
    } else if (interp == 1) {
        const float f = floorf(g);
        const float a = g - f;
        const int j0 = (int)f, j1 = (int)f + 1;
        *w0 = (j0 >= 0 && j0 <= fs - 1) ? 1.0f - a : 0.0f;
        *w1 = (j1 >= 0 && j1 <= fs - 1) ? a : 0.0f;
        *i0 = j0 < 0 ? 0 : (j0 > fs - 1 ? fs - 1 : j0);
        *i1 = j1 < 0 ? 0 : (j1 > fs - 1 ? fs - 1 : j1);
    } else {
        int a0 = (int)floorf(g + 0.5f);
        a0 = a0 < 0 ? 0 : (a0 > fs - 1 ? fs - 1 : a0);
        *i0 = *i1 = a0; *w0 = 1.0f; *w1 = 0.0f;
    }
}

/* mapping: 0 identity, 1 ball_to_cube_radial, 2 ball_to_cube_volume_preserving; interpolation: 0 linear,
 * 1 linear_border, 2 nearest_neighbor */
void o32_continuous_conv(const float* filters, int kz, int ky, int kx, int cin, int cout, const float* out_pos, int64_t n_out,
                         float extent, const float* offset, const float* inp_pos, const float* inp_feat,
                         const float* inp_importance, const int32_t* nbr_index, const float* nbr_importance,
                         const int64_t* row_splits, int align_corners, int mapping, int normalize, int interpolation, float* out) {
    const int K = kz * ky * kx;
    const int64_t KC = (int64_t)K * cin;
    const float inv_extent = 1.0f / extent;
    const float ox = offset ? offset[0] : 0.0f, oy = offset ? offset[1] : 0.0f, oz = offset ? offset[2] : 0.0f;
    enum { BLOCK = 32 };
#pragma omp parallel
    {
        float* B = (float*)malloc((size_t)BLOCK * KC * sizeof(float));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b0 = 0; b0 < n_out; b0 += BLOCK) {
            const int nb = (int)(n_out - b0 < BLOCK ? n_out - b0 : BLOCK);
            memset(B, 0, (size_t)nb * KC * sizeof(float));
            float norm[BLOCK];
            for (int m = 0; m < nb; ++m) {
                const int64_t o = b0 + m;
                float* Bm = B + (size_t)m * KC;
                float nrm = 0.0f;
                for (int64_t n = row_splits[o]; n < row_splits[o + 1]; ++n) {
                    const int32_t idx = nbr_index[n];
                    float x = inp_pos[3 * (int64_t)idx] - out_pos[3 * o];
                    float y = inp_pos[3 * (int64_t)idx + 1] - out_pos[3 * o + 1];
                    float z = inp_pos[3 * (int64_t)idx + 2] - out_pos[3 * o + 2];
                    const float a = nbr_importance ? nbr_importance[n] : 1.0f;
                    nrm += a;
                    const float w_pair = a * (inp_importance ? inp_importance[idx] : 1.0f);
                    map_to_cube(mapping, inv_extent, &x, &y, &z);
                    float gx, gy, gz;
                    if (align_corners) {
                        gx = (x + 0.5f) * (float)(kx - 1) + ox;
                        gy = (y + 0.5f) * (float)(ky - 1) + oy;
                        gz = (z + 0.5f) * (float)(kz - 1) + oz;
                    } else {
                        gx = (x + 0.5f) * (float)kx - 0.5f + ox;
                        gy = (y + 0.5f) * (float)ky - 0.5f + oy;
                        gz = (z + 0.5f) * (float)kz - 0.5f + oz;
                    }
                    int x0, x1, y0, y1, z0, z1;
                    float wx[2], wy[2], wz[2];
                    axis_w(interpolation, gx, kx, &x0, &x1, &wx[0], &wx[1]);
                    axis_w(interpolation, gy, ky, &y0, &y1, &wy[0], &wy[1]);
                    axis_w(interpolation, gz, kz, &z0, &z1, &wz[0], &wz[1]);
                    const int xi[2] = {x0, x1}, yi[2] = {y0, y1}, zi[2] = {z0, z1};
                    const float* f = inp_feat + (int64_t)idx * cin;
                    for (int c = 0; c < (interpolation == 2 ? 1 : 8); ++c) {
                        const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
                        const float w = wx[bx] * wy[by] * wz[bz] * w_pair;
                        if (w == 0.0f) continue;
                        float* dst = Bm + (size_t)((zi[bz] * ky + yi[by]) * kx + xi[bx]) * cin;
                        for (int ci = 0; ci < cin; ++ci) dst[ci] += w * f[ci];
                    }
                }
                norm[m] = nrm;
            }
            /* out[nb x cout] = B[nb x KC] * filters[KC x cout] */
#ifdef O32_TIMED
            /* timed build (libo32_timed.so, bench.py's CPU arm only): the same product as a register-blocked micro-kernel,
             * 4 out points x 16 channels of accumulators, FMA contraction allowed (-ffp-contract=fast) -- Open3D hands this
             * product to a BLAS-class GEMM, so the baseline is not handicapped by a scalar read-modify-write loop.  Results
             * differ from the parity build by float32 rounding only (tests/test_oracle_cpu.py). */
            for (int m0 = 0; m0 < nb; m0 += 4) {
                const int mr = nb - m0 < 4 ? nb - m0 : 4;
                const float* Br[4];
                for (int r = 0; r < 4; ++r) Br[r] = B + (size_t)(m0 + (r < mr ? r : 0)) * KC;
                for (int cb = 0; cb < cout; cb += 16) {
                    const int cw = cout - cb < 16 ? cout - cb : 16;
                    float acc[4][16];
                    if (cw == 16) {
                        /* GCC vector extension: 8 accumulators of 8 floats stay in ymm registers */
                        typedef float v8f __attribute__((vector_size(32), aligned(4)));
                        v8f a00 = {0}, a01 = {0}, a10 = {0}, a11 = {0}, a20 = {0}, a21 = {0}, a30 = {0}, a31 = {0};
                        const float* restrict wp = filters + cb;
                        for (int64_t k = 0; k < KC; ++k, wp += cout) {
                            const v8f w0 = *(const v8f*)wp, w1 = *(const v8f*)(wp + 8);
                            const float s0 = Br[0][k], s1 = Br[1][k], s2 = Br[2][k], s3 = Br[3][k];
                            const v8f v0 = {s0, s0, s0, s0, s0, s0, s0, s0}, v1 = {s1, s1, s1, s1, s1, s1, s1, s1};
                            const v8f v2 = {s2, s2, s2, s2, s2, s2, s2, s2}, v3 = {s3, s3, s3, s3, s3, s3, s3, s3};
                            a00 += v0 * w0; a01 += v0 * w1; a10 += v1 * w0; a11 += v1 * w1;
                            a20 += v2 * w0; a21 += v2 * w1; a30 += v3 * w0; a31 += v3 * w1;
                        }
                        *(v8f*)&acc[0][0] = a00; *(v8f*)&acc[0][8] = a01; *(v8f*)&acc[1][0] = a10; *(v8f*)&acc[1][8] = a11;
                        *(v8f*)&acc[2][0] = a20; *(v8f*)&acc[2][8] = a21; *(v8f*)&acc[3][0] = a30; *(v8f*)&acc[3][8] = a31;
                    } else if (cw == 8) {
                        typedef float v8f __attribute__((vector_size(32), aligned(4)));
                        v8f a0 = {0}, a1 = {0}, a2 = {0}, a3 = {0};
                        const float* restrict wp = filters + cb;
                        for (int64_t k = 0; k < KC; ++k, wp += cout) {
                            const v8f w0 = *(const v8f*)wp;
                            const float s0 = Br[0][k], s1 = Br[1][k], s2 = Br[2][k], s3 = Br[3][k];
                            a0 += (v8f){s0, s0, s0, s0, s0, s0, s0, s0} * w0; a1 += (v8f){s1, s1, s1, s1, s1, s1, s1, s1} * w0;
                            a2 += (v8f){s2, s2, s2, s2, s2, s2, s2, s2} * w0; a3 += (v8f){s3, s3, s3, s3, s3, s3, s3, s3} * w0;
                        }
                        *(v8f*)&acc[0][0] = a0; *(v8f*)&acc[1][0] = a1; *(v8f*)&acc[2][0] = a2; *(v8f*)&acc[3][0] = a3;
                    } else {
                        for (int r = 0; r < 4; ++r)
                            for (int c = 0; c < 16; ++c) acc[r][c] = 0.0f;
                        for (int64_t k = 0; k < KC; ++k) {
                            const float* restrict wrow = filters + k * cout + cb;
                            for (int r = 0; r < 4; ++r) {
                                const float v = Br[r][k];
                                for (int c = 0; c < cw; ++c) acc[r][c] += v * wrow[c];
                            }
                        }
                    }
                    for (int r = 0; r < mr; ++r)
                        for (int c = 0; c < cw; ++c) out[(b0 + m0 + r) * cout + cb + c] = acc[r][c];
                }
            }
            if (normalize)
                for (int m = 0; m < nb; ++m) {
                    float* dst = out + (b0 + m) * cout;
                    const float nv = nbr_importance ? norm[m] : (float)(row_splits[b0 + m + 1] - row_splits[b0 + m]);
                    if (nv != 0.0f)
                        for (int co = 0; co < cout; ++co) dst[co] /= nv;
                }
#else
            for (int m = 0; m < nb; ++m) {
                float* dst = out + (b0 + m) * cout;
                for (int co = 0; co < cout; ++co) dst[co] = 0.0f;
                const float* Bm = B + (size_t)m * KC;
                for (int64_t k = 0; k < KC; ++k) {
                    const float v = Bm[k];
                    if (v == 0.0f) continue;
                    const float* wrow = filters + k * cout;
                    for (int co = 0; co < cout; ++co) dst[co] += v * wrow[co];
                }
                if (normalize) {
                    const float nv = nbr_importance ? norm[m] : (float)(row_splits[b0 + m + 1] - row_splits[b0 + m]);
                    if (nv != 0.0f)
                        for (int co = 0; co < cout; ++co) dst[co] /= nv;
                }
            }
#endif
        }
        free(B);
    }
}

/* out[n x cout] = x[n x cin] * w[cin x cout] + b */
void o32_dense(const float* x, int64_t n, int cin, const float* w, const float* b, int cout, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        float* dst = out + r * cout;
        for (int co = 0; co < cout; ++co) dst[co] = b ? b[co] : 0.0f;
        for (int ci = 0; ci < cin; ++ci) {
            const float v = x[r * cin + ci];
            const float* wrow = w + (int64_t)ci * cout;
            for (int co = 0; co < cout; ++co) dst[co] += v * wrow[co];
        }
    }
}
