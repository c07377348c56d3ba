/* dmcf_b200.h -- C ABI of libdmcf_b200.so: the B200 (sm_100a) replacement for the native ops DMCF's per-step
 * particle hot path calls through TensorFlow's custom-op ABI.
 *
 * Reference interfaces replaced (paths relative to the reference checkout, tum-pbs/DMCF):
 *   - open3d.ml.tf.layers.FixedRadiusSearch -> ops build_spatial_hash_table + fixed_radius_search
 *         constructed at utils/convolutions.py:207-210, called at utils/convolutions.py:354-358,
 *         utils/tools/losses.py:296-298, 339-341
 *   - open3d.ml.tf.ops.continuous_conv     called at utils/convolutions.py:431, 454, 1054 with the kwargs
 *         assembled at utils/convolutions.py:414-429
 *   - open3d.ml.tf.ops.reduce_subarrays_sum called at models/pbf_model.py:450-453
 *   - tf.keras.layers.Dense (per-particle) models/pbf_model.py:140-152, models/hrnet.py:63-66
 *   - grid_pos (tf.unique based lattice sampling) utils/tools/losses.py:136-181
 *   - integrate / correct elementwise steps  models/pbf_model.py:234-250, 466-487
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns every buffer
 *     (the library never allocates or frees device memory);
 *   - every entry point takes the CUDA stream to enqueue on (a cudaStream_t passed as void*), never
 *     synchronises the device and is CUDA-graph capturable unless stated otherwise;
 *   - DEVICE-SIDE COUNTS (sync-free, CUDA-graph capturable steps): a point set may be a capacity-sized buffer whose number of
 *     valid rows lives in device memory (dmcf_grid::n_points_dev, n_queries_dev, dmcf_conv_desc::n_out_dev; NULL = the host
 *     count is exact).  The host count is then the CAPACITY; kernels are launched for it and ignore rows >= the device count.
 *     Data-dependent OUTPUT sizes (neighbour pairs, lattice points) are bounded by a caller-given capacity; exceeding it sets
 *     the caller's overflow flag (device int32, read once at the end of the step) instead of writing out of bounds;
 *   - return value 0 = success; otherwise an error code, message via dmcf_last_error() (thread local);
 *   - positions are float32 [n,3] row-major; features float32 row-major with an explicit row stride
 *     (in floats); neighbour lists are CSR: int32 index [P], int64 row_splits [n_out+1].
 */
#ifndef DMCF_B200_H
#define DMCF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMCF_B200_VERSION 105

enum dmcf_status {
    DMCF_OK = 0,
    DMCF_ERR_INVALID = 1,   /* bad argument (the reference raises InvalidArgument from OP_REQUIRES) */
    DMCF_ERR_UNSUPPORTED = 2,
    DMCF_ERR_WORKSPACE = 3, /* workspace too small */
    DMCF_ERR_CUDA = 4       /* a CUDA runtime call failed */
};

enum dmcf_mapping { DMCF_MAP_IDENTITY = 0, DMCF_MAP_BALL_TO_CUBE_RADIAL = 1, DMCF_MAP_BALL_TO_CUBE_VOLUME_PRESERVING = 2 };
enum dmcf_interp { DMCF_INTERP_LINEAR = 0, DMCF_INTERP_LINEAR_BORDER = 1, DMCF_INTERP_NEAREST = 2 };
/* window functions of utils/tools/losses.py:8-44 on q = d^2/r^2 */
enum dmcf_window { DMCF_WIN_NONE = 0, DMCF_WIN_POLY6 = 1, DMCF_WIN_CUBIC = 2, DMCF_WIN_LINEAR = 3, DMCF_WIN_PEAK = 4, DMCF_WIN_CUBIC_GRAD = 5 };

int dmcf_version(void);
const char* dmcf_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches claim) */
int64_t dmcf_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * Cell list ("spatial hash table" of the reference: open3d build_spatial_hash_table).
 * A uniform grid anchored at `origin` with cubic cells of edge 1/inv_cell and dims[] cells per axis;
 * points outside are clamped into the border cells, so any origin/dims is *correct* (tight ones are fast).
 * After dmcf_grid_build:  cell_start[c]..cell_start[c+1] delimit cell c (linear id (z*dims[1]+y)*dims[0]+x)
 * in sorted_index (original point ids, ascending inside a cell => deterministic) and sorted_pos
 * (x,y,z,bitcast(original id)) laid out cell-major for coalesced candidate reads.
 * ------------------------------------------------------------------------------------------------- */
typedef struct dmcf_grid {
    float origin[3];
    float inv_cell;
    int32_t dims[3];
    int32_t n_points;
    int32_t* cell_start;   /* [dims[0]*dims[1]*dims[2] + 1] */
    int32_t* sorted_index; /* [n_points] */
    float* sorted_pos;     /* [n_points,4], 16-byte aligned */
    const int32_t* n_points_dev; /* optional device-side count: only rows [0, *n_points_dev) of `points` are inserted (n_points is
                                    then the capacity; sorted_index / sorted_pos entries beyond the count are not written) */
    const float* points;   /* written by dmcf_grid_build: the array the grid was built from.  A search whose `queries` pointer
                              equals it (queries = a prefix of the points) takes the cell-centric kernel k_frs_cell */
    float mean_occupancy;  /* optional hint: points per occupied cell (0 = unknown, the library then uses n_points / cells of the
                              grid, which underestimates it for padded grids); k_frs_cell is used up to 12 points per cell */
} dmcf_grid;

size_t dmcf_grid_workspace_bytes(int64_t n_points, int64_t n_cells);
int dmcf_grid_build(const float* points, dmcf_grid* grid, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * fixed_radius_search (utils/convolutions.py:354-358): L2 metric, inclusive d^2 <= r*r evaluated in float32
 * as (dx*dx + dy*dy) + dz*dz without FMA contraction; ignore_query_point drops points whose position equals
 * the query position.  Two phases because P is data dependent:
 *   dmcf_frs_count  -> counts[n_queries] (this IS reduce_subarrays_sum of ones, models/pbf_model.py:450-453)
 *   dmcf_exclusive_scan_i32_i64 -> row_splits[n_queries+1]   (row_splits[n_queries] = P)
 *   dmcf_frs_fill   -> neighbors_index[P] (original point ids), neighbors_distance[P] (squared; may be NULL)
 * Row order: cells ascending (z,y,x), ascending point id inside a cell.
 * ------------------------------------------------------------------------------------------------- */
int dmcf_frs_count(const dmcf_grid* grid, const float* queries, int64_t n_queries, const int32_t* n_queries_dev, float radius,
                   int ignore_query_point, int32_t* counts, void* stream);
/* `capacity` bounds the pairs written; a row that would pass it sets *overflow_flag = 1 (if given) and is truncated.  With
 * n_queries_dev the rows >= *n_queries_dev get count 0 / are skipped. */
int dmcf_frs_fill(const dmcf_grid* grid, const float* queries, int64_t n_queries, const int32_t* n_queries_dev, float radius,
                  int ignore_query_point, const int64_t* row_splits, int64_t capacity,
                  int32_t* neighbors_index, float* neighbors_distance, int32_t* overflow_flag, void* stream);

size_t dmcf_scan_workspace_bytes(int64_t n);
/* out[0]=0, out[i]=sum(in[0..i)), out has n+1 entries */
int dmcf_exclusive_scan_i32_i64(const int32_t* in, int64_t n, int64_t* out, void* workspace, size_t workspace_bytes, void* stream);
int dmcf_exclusive_scan_i32_i32(const int32_t* in, int64_t n, int32_t* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * continuous_conv forward (utils/convolutions.py:414-431) with the layer glue DMCF wraps around it fused in:
 *   out[o,:] (+)= sum_n a_n s_n W(x_n - y_o)^T g(f_n)  [/ sum_n a_n]  + [g'(c_o) Wd] + bias + residual[o,:]
 *   g(f) = feat_scale * (relu_input ? max(f,0) : f)  (+ g(f_o) if `ascc`: the fused form of the
 *   antisymmetric layer's second pass, utils/convolutions.py:433-458)
 *   a_n  = neighbors_importance[n] if given, else window(d^2/r^2) (utils/convolutions.py:359-379), else 1
 *   W(.) = trilinear lookup in `filters` [kz,ky,kx,cin,cout] after the coordinate mapping
 *   g'(c_o) Wd = fused per-particle Dense on dense_inp (models/hrnet.py:94-96): filters then holds
 *   (kz*ky*kx*cin + dense_cin) rows, the Dense kernel [dense_cin,cout] appended after the conv filter.
 * ------------------------------------------------------------------------------------------------- */
typedef struct dmcf_conv_desc {
    int32_t kernel_size[3]; /* kz, ky, kx of the EFFECTIVE filter (after antisymmetric mirroring) */
    int32_t cin, cout;
    int32_t mapping;        /* enum dmcf_mapping */
    int32_t interpolation;  /* enum dmcf_interp */
    int32_t align_corners;
    int32_t normalize;
    int32_t window;         /* enum dmcf_window, used only when neighbors_importance == NULL */
    float window_fac;
    float extent;           /* filter diameter = 2*radius (scalar extents only) */
    float offset[3];        /* x,y,z */
    int32_t relu_input;
    float feat_scale;
    int32_t ascc;           /* add the centre feature to every neighbour feature: out point o must be input row o
                               (out set == inp set, or a prefix of it when ghost rows are appended) */
    int32_t skip_self;      /* drop neighbours whose position equals the out position (lets one CSR that
                               contains self serve ignore_query_point layers) */
    int32_t nbr_lo, nbr_hi; /* keep only neighbours with nbr_lo <= index < nbr_hi; feature row = index - nbr_lo.
                               nbr_hi <= nbr_lo means "all" */
    int32_t dense_cin;      /* 0 = no fused Dense */
    int32_t accumulate;     /* out += result instead of out = result */
    int32_t filter_antisym; /* the caller guarantees filters[kz-1-z][ky-1-y][kx-1-x] == -filters[z][y][x] (bit exact), as
                               the antisymmetric layer builds its effective kernel (utils/convolutions.py:410-412):
                               allows the folded half-patch kernel k_cconv_apatch.  0 is always safe */
    const int32_t* n_out_dev; /* optional device-side number of out points (<= n_out, which is then the capacity the kernels
                                 are launched for); rows >= *n_out_dev of `out` are not written */
    int32_t block_cin;      /* > 0: the caller guarantees a BLOCK DIAGONAL layer, as the fused input layer of the DMCF nets is built
                               (fluid_convs + obs_convs over zero-padded [fluid | box] features, models/pbf_model.py:375-411):
                               input channels [0, block_cin) reach outputs [0, block_cout[0]) only, input channels
                               [block_cin, cin) reach outputs [block_cout[0], block_cout[0] + block_cout[1]) only (every other
                               conv filter entry is zero; fused Dense rows are not restricted), and every input feature row is
                               all zero in one of the two channel groups.  Allows the narrow direct kernel k_cconv_narrow.
                               0 is always safe */
    int32_t block_cout[2];
} dmcf_conv_desc;

int dmcf_cconv_forward(const dmcf_conv_desc* desc, const float* filters,
                       const float* out_positions, int64_t n_out,
                       const float* inp_positions, const float* inp_features, int64_t inp_stride, int64_t n_inp,
                       const float* inp_importance,
                       const int32_t* neighbors_index, const int64_t* neighbors_row_splits,
                       const float* neighbors_importance,
                       const float* bias, const float* dense_inp, int64_t dense_stride,
                       const float* residual, int64_t residual_stride,
                       float* out, int64_t out_stride,
                       const float* pair_records, int64_t n_pairs, void* stream);

/* Optional: evaluate the per-pair geometry (coordinate mapping, trilinear corner cells/weights, window / importances,
 * sub-range and skip-self filtering) ONCE for a neighbour list and reuse it for every conv that shares the list and the
 * geometry fields of `desc` (kernel_size, mapping, interpolation, align_corners, extent, offset, window*, skip_self,
 * nbr_lo/hi) -- e.g. the input conv and the three CConv layers of one DMCF step.  `records` holds
 * dmcf_cconv_records_bytes(n_pairs) bytes (9 float arrays of n_pairs).  Pass it as `pair_records` to
 * dmcf_cconv_forward; positions / neighbors_index / importances are then not read.  Not available with `normalize`. */
size_t dmcf_cconv_records_bytes(int64_t n_pairs);
int dmcf_cconv_prepare(const dmcf_conv_desc* desc, const float* out_positions, int64_t n_out,
                       const float* inp_positions, int64_t n_inp, const float* inp_importance,
                       const int32_t* neighbors_index, const int64_t* neighbors_row_splits,
                       const float* neighbors_importance, int64_t n_pairs, float* records, void* stream);

/* The "patch" rows of a continuous_conv (no reference counterpart as a separate op: Open3D materialises the same matrix
 * inside continuous_conv / continuous_conv_backprop_filter): patches[o, cell*cin + ci] = sum_n a_n w_cell(n,o) g(f_n)[ci] with
 * a_n, g as in dmcf_cconv_forward (window, relu_input, feat_scale, skip_self, nbr range).  out = patches @ filters, so the
 * gradient of a conv w.r.t. its filter is patches^T @ d_out (dmcf_b200/autograd.py).  patch_stride >= kz*ky*kx*cin floats.
 * desc->cout is ignored; dense_cin, normalize and ascc must be 0. */
int dmcf_cconv_patches(const dmcf_conv_desc* desc, const float* out_positions, int64_t n_out,
                       const float* inp_positions, const float* inp_features, int64_t inp_stride, int64_t n_inp,
                       const float* inp_importance, const int32_t* neighbors_index, const int64_t* neighbors_row_splits,
                       const float* neighbors_importance, const float* pair_records, int64_t n_pairs,
                       float* patches, int64_t patch_stride, void* stream);

/* Kernel selection bit mask (default 3): bit 0 = register-patch kernels for compile-time filter grids (k_cconv_lean;
 * k_cconv_wide where the lean kernel is not eligible), bit 1 = resident-filter direct kernel for cout <= 4
 * (k_cconv_direct) and the folded half-patch kernel for antisymmetric filters (k_cconv_apatch), bit 2 = run 4x4x4 layers of
 * the legacy k_cconv_wide as two z-half launches, bit 3 = use the legacy k_cconv_wide instead of k_cconv_lean, bit 4 = do
 * not use k_cconv_apatch, bit 5 = k_cconv_lean keeps the one-pair-per-step walk for inputs with <= 8 channels instead of the
 * multi-pair phase 1, bit 6 = searches whose queries are a prefix of the grid's points keep the query-centric k_frs instead of
 * the cell-centric k_frs_cell, bit 7 = the wide layers run the warp-specialised k_cconv_ws (producer / consumer warps on
 * double-buffered half tiles; a measured experiment, slower than k_cconv_lean), bit 12 = do not use the narrow direct kernel
 * k_cconv_narrow, bit 13 = dmcf_dense_forward keeps the SIMT kernel k_dense instead of the tensor-core kernel k_dense_umma,
 * bit 15 = k_cconv_lean keeps its FFMA2 patch x filter product instead of the tensor-core one (mma.sync, 3xTF32),
 * bit 14 = the tensor-core k_cconv_lean launches one CTA per tile instead of persistent CTAs, bit 16 = its product runs on
 * 16-point / 16-warp tiles instead of 24 points / 12 warps (bits 3-7 and 12-16 are kept for A/B measurements);
 * 0 forces the generic kernel.  Returns the previous
 * mask.  Results agree to float32 rounding. */
int dmcf_set_kernel_options(int options);

/* ---------------------------------------------------------------------------------------------------
 * per-particle Dense:  out[n,:] = (relu_input ? max(x,0) : x) @ W[cin,cout] + b   (Keras layout)
 * ------------------------------------------------------------------------------------------------- */
int dmcf_dense_forward(const float* x, int64_t n, int32_t cin, int64_t x_stride, const float* w, const float* b,
                       int32_t cout, int32_t relu_input, float* out, int64_t out_stride, void* stream);

/* Measured experiment, not a product path (profiles/README.md, "tensor cores for the conv"): the tcgen05 shape a conv's
 * patch x filter product would have.  Every CTA runs n_tiles x ks k-steps of  D[64 x n] += A_s[64 x 16] B_s[n x 16]^T
 * (kind::f16, `passes` B operands per step) with the A operand (2 KB per k-step, UMMA K-major no-swizzle core-matrix order)
 * streamed from `a_stream` through a ring of `stages` cp.async.bulk copies and the B operand (`passes` x ks x 2 chunks of
 * (n / 8) x 128 + 16 bytes) resident in shared memory; d_out (may be NULL) receives the 128 x n accumulator lanes of CTA 0,
 * stats (may be NULL) five cycle counters of CTA 0: producer {wait for a free slot, issue the copy}, MMA thread {wait for the
 * data, issue the MMAs, commit}. */
int dmcf_umma_probe(const void* a_stream, const void* b_tile, int32_t ks, int32_t n, int32_t n_tiles, int32_t stages,
                    int32_t passes, int32_t n_ctas, float* d_out, long long* stats, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * elementwise step pieces (models/pbf_model.py:234-250, 466-487)
 *   integrate: vel2 = vel + dt*acc (acc == NULL -> gravity vector), pos2 = pos + dt*vel2
 *   correct:   pos' = pos2 + out_scale * net[:, map(c)];  vel' = (pos' - pos)/dt     net has net_c in {1,2,3}
 *              channels (1 -> repeated, 2 -> [a,b,a], models/pbf_model.py:466-469)
 * ------------------------------------------------------------------------------------------------- */
int dmcf_integrate(const float* pos, const float* vel, const float* acc, const float* gravity_host3, float dt,
                   int64_t n, float* pos2, float* vel2, void* stream);
int dmcf_correct(const float* pos, const float* pos2, const float* net, int64_t net_stride, int32_t net_c,
                 const float* out_scale_host3, float dt, int64_t n, float* pos_new, float* vel_new, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * grid_pos (utils/tools/losses.py:136-181): lattice points of pitch `voxel` touched by any particle, with the
 * +-hyst hysteresis and the {0,1} corner offsets on active axes (voxel >= 1e-5).  Float32 arithmetic in the
 * reference's operation order.  The caller supplies the integer lattice bounds lo/dims (computable from the
 * position min/max because every step is monotone) and a zeroed flags[dims0*dims1*dims2] array (x fastest):
 *   dmcf_grid_pos_mark -> flags;  dmcf_exclusive_scan_i32_i32(flags) -> offsets (offsets[n_cells] = count);
 *   dmcf_grid_pos_emit -> out[count,3] in ascending linear voxel id (the reference's tf.unique order is
 *   first-occurrence; consumers are order independent).
 * `center_host3` / `center_dev3` (the same 3 floats in host or in device memory; at most one of them) both NULL = not
 * centralised (points sit at voxel centres g*v + v/2, else g*v + center).  Sync-free use: lo/dims come from a plan instead of
 * this step's min/max; a particle that falls outside them sets *overflow_flag (mark), as does a lattice of more than
 * `capacity` points (emit, which then writes only the first `capacity`; capacity < 0 = unbounded); n_dev = optional device-side
 * particle count.
 * ------------------------------------------------------------------------------------------------- */
int dmcf_grid_pos_mark(const float* pos, int64_t n, const int32_t* n_dev, const float* voxel_host3, const float* center_host3,
                       const float* center_dev3, float hyst, const int32_t* lo_host3, const int32_t* dims_host3, int32_t* flags,
                       int32_t* overflow_flag, void* stream);
int dmcf_grid_pos_emit(const int32_t* flags, const int32_t* offsets, const float* voxel_host3, const float* center_host3,
                       const float* center_dev3, const int32_t* lo_host3, const int32_t* dims_host3, float* out,
                       int64_t capacity, int32_t* overflow_flag, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Row mover for capacity-sized point sets (no reference counterpart: TensorFlow's dynamic shapes hide this; here it is the glue
 * of the sync-free multi-GPU step -- halo packing, ghost rows appended behind the owned rows, migration):
 *   dst[(base + i) * dst_stride + c] = src[row(i) * src_stride + c],   i < count, c < width
 *   base = base_host + (base_dev ? *base_dev : 0);  count = src_count_dev ? min(*src_count_dev, n_src) : n_src;
 *   row(i) = src_index ? src_index[i] : i  (gather and append in one pass);  *new_count_dev (optional) = base + count.
 * Rows that would pass dst_capacity are dropped and set *overflow_flag.
 * ------------------------------------------------------------------------------------------------- */
int dmcf_rows_append(float* dst, int64_t dst_stride, int64_t dst_capacity, int64_t base_host, const int32_t* base_dev,
                     const float* src, int64_t src_stride, int64_t n_src, const int32_t* src_count_dev,
                     const int64_t* src_index, int32_t width, int32_t* new_count_dev, int32_t* overflow_flag, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Point-set ops of the reference's in-repo CUDA extensions.
 *
 * dmcf_farthest_point_sample  replaces op FarthestPointSample (utils/tools/sampling.cpp:49-61,115-148, kernel
 *   utils/tools/sampling.cu:125-190; called at utils/tools/losses.py:241, 277-279).  points [b,n,3] -> idx_out [b,m]
 *   int32, first index 0, then repeatedly the point farthest from the chosen set; float32 distance
 *   fma(dz,dz, fma(dx,dx, dy*dy)) and the reference kernel's tie order (smaller k mod 512, then smaller k).
 *   temp [b,n] floats of scratch.  cluster_size: CTAs per batch item (thread-block cluster), 0 = choose, else 1/2/4/8.
 * dmcf_approx_match  replaces op ApproxMatch (utils/tools/tf_approxmatch.cpp:33-38, kernels tf_approxmatch.cu:27-160,
 *   CPU twin tf_approxmatch.cpp:51-112; called at utils/tools/losses.py:406, 413) for ONE batch item with n points in
 *   xyz1 and m in xyz2 (the reference's per-item counts n[i], m[i]: pass those counts and the row stride of the padded
 *   matrix).  match [m][match_ld] (match[l*match_ld + k], like the reference's [b, m, n] layout) may be NULL;
 *   cost_out (one float, may be NULL) receives sum_kl |x1_k - x2_l| match[l][k] (= op MatchCost) accumulated on the
 *   fly, so the EMD metric never needs the n x m matrix.  first_level: the annealing starts at exp(-4^first_level d^2):
 *   7 = the reference's CUDA kernel (tf_approxmatch.cu:47), 8 = its CPU kernel (tf_approxmatch.cpp:59).
 * dmcf_match_cost  replaces op MatchCost (tf_approxmatch.cpp:39-43,177-196; losses.py:407) for one batch item.
 * dmcf_nn_distance replaces one direction of op NnDistance (utils/tools/nn_distance.cpp:29-35,47-70): for every point of
 *   xyz1 the squared distance to / index of its nearest point in xyz2 ((x*x + y*y) + z*z uncontracted, first minimum).
 * ------------------------------------------------------------------------------------------------- */
int dmcf_farthest_point_sample(const float* points, int32_t b, int32_t n, int32_t m, float* temp, int32_t* idx_out,
                               int32_t cluster_size, void* stream);
size_t dmcf_approx_match_workspace_bytes(int32_t n, int32_t m);
int dmcf_approx_match(const float* xyz1, int32_t n, const float* xyz2, int32_t m, int32_t first_level, float* match,
                      int64_t match_ld, float* cost_out, void* workspace, size_t workspace_bytes, void* stream);
size_t dmcf_match_cost_workspace_bytes(int32_t n, int32_t m);
int dmcf_match_cost(const float* xyz1, int32_t n, const float* xyz2, int32_t m, const float* match, int64_t match_ld,
                    float* cost_out, void* workspace, size_t workspace_bytes, void* stream);
/* op MatchCostGrad (tf_approxmatch.cpp:44-50,198-232): gradients of the match cost w.r.t. both point sets, match held constant;
 * grad1 [n,3], grad2 [m,3]. */
int dmcf_match_cost_grad(const float* xyz1, int32_t n, const float* xyz2, int32_t m, const float* match, int64_t match_ld,
                         float* grad1, float* grad2, void* stream);
int dmcf_nn_distance(const float* xyz1, int32_t n, const float* xyz2, int32_t m, float* dist, int32_t* idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMCF_B200_H */
