/*
 * pct_b200 -- C ABI of the B200-native curvature hot path.
 *
 * The reference (masnottuh/point-cloud-toolbox) has no FFI, plugin or operator
 * interface: its hot path is five pure-Python methods of one class,
 * `PointCloud` in pointCloudToolbox.py, calling scipy/numpy.  The drop-in
 * boundary is therefore that class's surface; this header is what a Python
 * `PointCloud` binds through `ctypes` (see INTEGRATION.md for the stub) and
 * every entry point below names the reference lines it replaces.
 *
 * Conventions
 *   - plain C, no C++ exceptions cross the boundary; return 0 (PCT_OK) or a
 *     negative PCT_ERR_* code, `pct_last_error()` gives a thread-local message
 *   - every pointer except `pct_index*` and the `*_host` entry is a DEVICE
 *     pointer owned by the caller (PyTorch allocates), 16-byte aligned
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream);
 *     calls are asynchronous on it unless stated otherwise
 *   - the library owns only `pct_index` (opaque; immutable after build)
 *   - per-point failures are reported in `status[]` bits, never as an error
 *     code; outputs of a failed point are NaN
 *   - threading: calls on ONE stream must come from one host thread at a time
 *     (the temporaries of a call live in a scratch arena keyed by the stream,
 *     which the next call on that stream reuses); different streams may be
 *     driven from different threads concurrently
 *
 * Query ranges and output layout
 *   The index keeps the cloud Morton-sorted.  Queries are addressed by SORTED
 *   position [q_begin, q_end) so that ranks of a multi-GPU job take contiguous,
 *   spatially coherent slices.  `layout` selects where row r of an output goes:
 *     PCT_LAYOUT_ORIGINAL  row = original index of the query point; outputs are
 *                          sized for all N points (single-GPU drop-in)
 *     PCT_LAYOUT_SLICE     row = sorted position - q_begin; outputs are sized
 *                          (q_end - q_begin) (multi-GPU shard, gathered later)
 *   Neighbour indices written to `idx` are always ORIGINAL point indices.
 */
#ifndef PCT_B200_H
#define PCT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCT_VERSION 200

#define PCT_OK 0
#define PCT_ERR_INVALID_ARGUMENT (-1)
#define PCT_ERR_CUDA (-2)
#define PCT_ERR_NONFINITE (-3)   /* cloud holds NaN/Inf: ref :273-274 ValueError */
#define PCT_ERR_K_TOO_LARGE (-4) /* k + 1 > N: ref :640 IndexError */
#define PCT_ERR_NO_DEVICE (-5)

#define PCT_LAYOUT_ORIGINAL 0
#define PCT_LAYOUT_SLICE 1

/* status[] bits */
#define PCT_STATUS_OK 0u
#define PCT_STATUS_EXACT_PATH 1u     /* informational: resolved by the exact tie/expansion path */
#define PCT_STATUS_FEW_NEIGHBORS 2u  /* < 2 neighbours (no covariance) -> NaN */
#define PCT_STATUS_RANK_DEFICIENT 4u /* rank-deficient design: the coefficients are lstsq's minimum-norm solution (ref :359), finite */
#define PCT_STATUS_NONFINITE 8u      /* non-finite intermediate: ref :318-319 / :356-357 ValueError */
#define PCT_STATUS_BAD_INDEX 32u     /* rows entry points: a caller-provided index lies outside [-N, N) -> NaN */
#define PCT_STATUS_UNRESOLVED 16u    /* slab index only (pct_index_set_slab): the k-th neighbour may lie outside the slab */

typedef struct pct_index pct_index;

typedef struct pct_index_info {
    int64_t num_points;
    float cell_size;      /* level-0 cell edge */
    float origin[3];      /* bounding-box minimum */
    float extent[3];      /* bounding-box size */
    int32_t dims[3];      /* level-0 grid dimensions */
    int32_t bits_per_axis;
    int32_t num_levels;
    int64_t cells_level0; /* occupied cells */
    int64_t device_bytes; /* HBM held by the index */
    float est_dimension;  /* intrinsic dimension seen by the density pilot */
    int32_t build_launches; /* kernels of this library the build launched (CUB's sort kernels not counted) */
} pct_index_info;

typedef struct pct_query_stats {
    int64_t queries;
    int64_t level1_retries; /* queries whose k-th neighbour left the level-0 stencil */
    int64_t exact_path;     /* queries resolved by the exact (tie / expansion) kernel */
    int64_t kernel_launches;
    int64_t unstaged;       /* queries whose chunk did not fit the shared-memory staging buffer */
    int64_t unresolved;     /* slab index only: queries returned with PCT_STATUS_UNRESOLVED */
    int64_t rank_deficient; /* fused kNN fit: neighbourhoods answered with lstsq's minimum-norm solution */
} pct_query_stats;

int pct_version(void);
const char* pct_last_error(void);

/* Spatial index: replaces `sp.spatial.cKDTree(points)` (ref :74).
 * xyz: N rows of `stride` floats (3 = packed xyz, 4 = padded), fp32.
 * cell_hint > 0 fixes the level-0 cell edge; otherwise it is chosen from a
 * density pilot so that a cell holds about 0.46 * k_hint points (k_hint <= 24; fewer per k above) (k_hint <= 0: 20).
 * Synchronises `stream` (the grid shape is needed on the host).
 * PCT_ERR_NONFINITE if any coordinate is NaN/Inf. */
int pct_index_build(const float* xyz, int64_t n, int stride, float cell_hint, int k_hint,
                    void* stream, pct_index** out);
/* The cell edge pct_index_build would choose for this cloud (density pilot only, no index), and the
 * bounding box {min xyz, max xyz} if bbox_min_max != NULL.  Lets the ranks of a spatially partitioned
 * job agree on one cell size before each builds the index of its own slab.  Synchronises `stream`. */
int pct_estimate_cell_size(const float* xyz, int64_t n, int stride, int k_hint, void* stream,
                           float* cell_size, float* bbox_min_max);
/* Declares the index to be one slab of a spatially partitioned cloud (multi-GPU, new; the reference is
 * single process): it was built from ALL cloud points with complete_lo <= coordinate[axis] <= complete_hi,
 * and answers only queries with own_lo <= coordinate[axis] < own_hi (other query rows are left untouched).
 * No search radius reaches beyond the complete range; a query whose k-th neighbour cannot be proven to
 * lie inside it gets PCT_STATUS_UNRESOLVED (NaN outputs; idx -1) and must be answered from a larger index.
 * row_map (device, N int32, may be NULL; the caller keeps it alive while the index is used): with
 * PCT_LAYOUT_ORIGINAL the row of an owned point is row_map[original index] instead of the original index,
 * so the outputs can be sized for the owned points only.
 * axis = -1 removes the restriction.  kNN entry points only. */
int pct_index_set_slab(pct_index* index, int axis, float complete_lo, float complete_hi,
                       float own_lo, float own_hi, const int32_t* row_map);
/* Slab selection on the device (new, multi-GPU): pct_slab_select writes to `sel` (device, capacity N int32)
 * the ASCENDING original indices of the points with complete_lo <= coordinate[axis] <= complete_hi and
 * returns their number; pct_slab_gather copies those points to `local_xyz` (m x 3 packed) and writes
 * `row_map` (m int32): for a point with own_lo <= coordinate[axis] < own_hi its rank among the owned
 * points (the compact output row pct_index_set_slab expects), and returns the number of owned points.
 * Both synchronise `stream`. */
int pct_slab_select(const float* xyz, int64_t n, int stride, int axis, float complete_lo, float complete_hi,
                    int32_t* sel, int64_t* num_selected, void* stream);
int pct_slab_gather(const float* xyz, int stride, int axis, const int32_t* sel, int64_t m, float own_lo,
                    float own_hi, float* local_xyz, int32_t* row_map, int64_t* num_owned, void* stream);
/* Slab exchange (new, multi-GPU): every rank holds a contiguous share of the cloud (original indices id_base ..
 * id_base + n - 1) and the ranks trade points with ONE all-to-all instead of replicating the cloud.
 * bounds (host, world x 4 floats): complete_lo, complete_hi, own_lo, own_hi of every slab, as pct_index_set_slab takes them.
 * pct_estimate_cell_size_sample: the cell edge pct_index_build would choose for a cloud of n_total points of which
 *   `sample` (device) holds n_sample; bbox_min_max (host, 6 floats) is the WHOLE cloud's box.  Synchronises `stream`.
 * pct_slab_bin_count: counts (host, 2 x world int64) = number of the share's points inside every slab's complete range,
 *   then the number every slab owns; block_pos (device, 2 * world * pct_slab_bin_blocks(n) + 1 int32) is working storage
 *   that pct_slab_bin_fill reads.  Synchronises `stream`.
 * pct_slab_bin_fill: records (device, total_complete x 4 floats {x, y, z, original index as int bits}) grouped by
 *   destination slab, ascending index inside a group (distance ties keep the whole cloud's order at the receiver);
 *   owned_local (device, n int32): the share's local indices grouped by OWNER slab, ascending -- the order in which the
 *   owners return their rows.
 * pct_slab_rows: row_map of pct_index_set_slab for a received slab cloud (m rows): rank of every owned point.
 * pct_slab_row_ids: row_ids (device, owned-rows int32) = the original index of every output row of that slab (the id
 *   column of the owned points in row order), which pct_index_set_peers needs. */
int pct_slab_row_ids(const float* xyz4, int64_t m, int axis, float own_lo, float own_hi, const int32_t* row_map,
                     int32_t* row_ids, void* stream);
/* Return fused into the kernel (new, multi-GPU; replaces the gather SURVEY.md 8(e) puts after the queries): rank r of
 * `world` holds the original indices begins[r] .. begins[r + 1] - 1 and a result array of (its rows) x {K, H} floats
 * that the other ranks have mapped (CUDA IPC / NVLink peer memory); peer_rows[r] (host array of `world` device
 * pointers) is that array as THIS process sees it.  After this call pct_curvature_fused_knn_records on the slab index
 * stores K and H of every answered query with one 8-byte store into the array of the rank that holds the query's
 * point (besides its local record), so no all-to-all and no scatter follow the kernel -- the ranks only need a
 * barrier before they read their arrays.  row_ids: pct_slab_row_ids (caller keeps it alive).  begins: host, world + 1.
 * world = 0 removes the routing.  Asynchronous on the stream the index was built on. */
int pct_index_set_peers(pct_index* index, int world, const int64_t* begins, void* const* peer_rows,
                        const int32_t* row_ids);
int pct_estimate_cell_size_sample(const float* sample, int64_t n_sample, int stride, int64_t n_total,
                                  const float* bbox_min_max, int k_hint, void* stream, float* cell_size);
int64_t pct_slab_bin_blocks(int64_t n);
int pct_slab_bin_count(const float* xyz, int64_t n, int stride, int axis, int world, const float* bounds,
                       int32_t* block_pos, int64_t* counts, void* stream);
int pct_slab_bin_fill(const float* xyz, int64_t n, int stride, int axis, int world, const float* bounds,
                      const int32_t* block_pos, int64_t total_complete, int64_t id_base, float* records,
                      int32_t* owned_local, void* stream);
/* The forward exchange fused into the binning kernel (new, multi-GPU): like pct_slab_bin_fill, but the records of
 * destination d are stored straight into rank d's slab buffer -- peer_slabs[d] (host array of `world` device pointers:
 * the buffer as THIS process maps it, CUDA IPC / NVLink peer memory), from row dest_rows[d] on (host; the number of
 * records the lower ranks send to d, so that the received slab is in ascending original index).  complete_counts:
 * the first `world` counts of pct_slab_bin_count.  The ranks need a barrier before they read their buffers. */
int pct_slab_bin_fill_peers(const float* xyz, int64_t n, int stride, int axis, int world, const float* bounds,
                            const int32_t* block_pos, const int64_t* complete_counts, int64_t id_base,
                            void* const* peer_slabs, const int64_t* dest_rows, int32_t* owned_local, void* stream);
int pct_slab_rows(const float* xyz, int64_t m, int stride, int axis, float own_lo, float own_hi, int32_t* row_map,
                  void* stream);
/* frees the index in the order of the stream it was built on (no device synchronisation);
 * queries issued on OTHER streams must have completed */
int pct_index_destroy(pct_index* index);
int pct_index_get_info(const pct_index* index, pct_index_info* info);
/* perm[sorted position] = original index, int32 x N */
int pct_index_permutation(const pct_index* index, int32_t* perm, void* stream);
/* counters of the most recent query call on this index (synchronises `stream`) */
int pct_index_last_stats(const pct_index* index, void* stream, pct_query_stats* stats);

/* kNN lists: replaces the `kdtree.query(point, k+1)` loop of plant_kdtree (ref :81-85).
 * Row = the k nearest OTHER entries of the (k+1)-list ordered by (fp64 squared
 * distance, index) with the first entry dropped, exactly like the reference.
 * idx: rows x k int32 (original indices), dist: rows x k fp32 (either may be NULL). */
int pct_knn(const pct_index* index, int64_t q_begin, int64_t q_end, int k,
            int32_t* idx, float* dist, int layout, void* stream);

/* kNN lists of selected cloud points: replaces `self.kdtree.query(point, n + 1)` of the neighbour study
 * (ref :759), which probes a few hundred sampled points with up to 100 neighbours.  query_ids: nq original
 * indices; xyz: the cloud the index was built from (the points are located through their own cell).
 * Row r = the k nearest OTHER points of point query_ids[r], ordered like pct_knn; idx nq x k, dist nq x k.
 * Synchronises `stream`; PCT_ERR_INVALID_ARGUMENT if an id is not a point of the indexed cloud. */
int pct_knn_points(const pct_index* index, const float* xyz, int stride, const int32_t* query_ids,
                   int64_t nq, int k, int32_t* idx, float* dist, void* stream);

/* kNN of ARBITRARY coordinates: `self.kdtree.query(x, k)` as scipy defines it (the reference only passes cloud points,
 * ref :83, :759).  queries: nq x 3 fp32 on the device.  Row r = the k nearest cloud points of queries[r] ordered by
 * (fp64 squared distance, index) -- a query that is a cloud point finds itself first, at distance 0; nothing is dropped.
 * idx nq x k int32, dist nq x k fp64 (scipy returns float64).  Whole-cloud indexes only.  A query outside the cloud's
 * bounding box is answered by a scan of the whole cloud (correct, slow for large clouds). */
int pct_knn_query(const pct_index* index, const float* queries, int64_t nq, int k, int32_t* idx, double* dist,
                  void* stream);

/* fused search + fit for selected cloud points (same addressing as pct_knn_points), packed records nq x 8 */
int pct_curvature_points_records(const pct_index* index, const float* xyz, int stride,
                                 const int32_t* query_ids, int64_t nq, int k, float* records, void* stream);

/* epsilon-ball (advertised README.md:8, absent in the reference; semantics of
 * scipy `query_ball_point`: d2 <= radius*radius in fp64, self excluded).
 * count: rows x int32.  fill: CSR rows given exclusive `offsets` (rows + 1, int64),
 * each row ordered by (d2, index); nnz = offsets[rows] = length of idx / dist. */
int pct_ball_count(const pct_index* index, int64_t q_begin, int64_t q_end, double radius,
                   int32_t* counts, int layout, void* stream);
int pct_ball_fill(const pct_index* index, int64_t q_begin, int64_t q_end, double radius,
                  const int64_t* offsets, int64_t nnz, int32_t* idx, float* dist, int layout,
                  void* stream);

/* Fit from given neighbour lists: replaces fit_explicit_quadratic_surfaces_to_neighborhoods
 * (ref :635-647) + calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points
 * (ref :657-674) for rows of `idx` (nq x k, original indices, reference order:
 * first = nearest, last = farthest).  query_ids (nq, may be NULL = 0..nq-1) names
 * the cloud point each row belongs to.  xyz: N x 3 packed fp32, original order.
 * normals nq x 3, coeffs nq x 6 [A,B,C,D,E,F], curv nq x 5 [K,H,k1,k2,H^2],
 * status nq (any output may be NULL). */
int pct_fit_from_neighbors(const float* xyz, int64_t n, const int32_t* idx, int64_t nq, int k,
                           const int32_t* query_ids, float* normals, float* coeffs, float* curv,
                           uint8_t* status, void* stream);
/* same for variable-size rows (CSR offsets nq + 1 int64) */
int pct_fit_from_csr(const float* xyz, int64_t n, const int64_t* offsets, const int32_t* idx,
                     int64_t nq, const int32_t* query_ids, float* normals, float* coeffs,
                     float* curv, uint8_t* status, void* stream);

/* Fused search + fit, neighbourhoods never leave the SM: the throughput path of
 * plant_kdtree(k) + compute_pointwise_explicit_quadratic_curvature() (ref :69-89, :505-509). */
int pct_curvature_fused_knn(const pct_index* index, int64_t q_begin, int64_t q_end, int k,
                            float* normals, float* coeffs, float* curv, uint8_t* status,
                            int layout, void* stream);
int pct_curvature_fused_ball(const pct_index* index, int64_t q_begin, int64_t q_end, double radius,
                             int32_t* counts, float* normals, float* coeffs, float* curv,
                             uint8_t* status, int layout, void* stream);

/* Same kernels, packed output: `records` = rows x 8 fp32 {nx, ny, nz, K, H, k1, k2, status bits},
 * 32-byte aligned.  One full DRAM sector per point even when rows are scattered to original
 * order (the separate arrays above cost three partial sectors); H^2 = H*H is left to the caller. */
int pct_curvature_fused_knn_records(const pct_index* index, int64_t q_begin, int64_t q_end, int k,
                                    float* records, int layout, void* stream);
int pct_curvature_fused_ball_records(const pct_index* index, int64_t q_begin, int64_t q_end,
                                     double radius, int32_t* counts, float* records, int layout,
                                     void* stream);

/* Batched forms of the three static methods.
 * pct_plane_rotate: get_best_fit_plane_and_rotate (ref :270-321) on nq neighbourhoods of
 *   k centred points (fp32, nq x k x 3) -> rotated fp64 (nq x k x 3), unit normals fp64 (nq x 3).
 * pct_quadric_fit: fit_quadratic_surface (ref :331-360) on nq x k x 3 fp64 -> coeffs fp32 nq x 6.
 * pct_quadric_curvature: calculate_explicit_quadratic_curvatures (ref :398-431), nq x 6 -> nq x 5. */
int pct_plane_rotate(const float* centered, int64_t nq, int k, double* rotated, double* normals,
                     uint8_t* status, void* stream);
int pct_quadric_fit(const double* rotated, int64_t nq, int k, float* coeffs, uint8_t* status,
                    void* stream);
int pct_quadric_curvature(const float* coeffs, int64_t nq, float* curv, void* stream);

/* Implicit 10-coefficient quadric A x^2 + B y^2 + C z^2 + D xy + E xz + F yz + G x + H y + I z + J = 0 (ref :363-396,
 * :435-480, :617-633, :676-689).
 * pct_implicit_quadric_fit: the minimiser of |A c|^2 on the unit sphere -- the problem the reference hands to SLSQP
 *   from the all-ones start, which stops far from the minimiser: coefficients are NOT comparable
 *   with the reference's ("parity unpinned").  Neighbourhoods either as rows of original indices (nq x k, centred on
 *   point query_ids[r] in fp32 like ref :627) or, when `centered` != NULL, as nq x k x 3 centred fp32 points.
 *   coeffs nq x 10 fp64, unit norm, signed so that (G, H, I) points away from the neighbours' centroid.
 * pct_implicit_quadric_curvature: calculate_implicit_quadric_curvatures (ref :435-480) as written, fp64:
 *   curv nq x 4 = [K_g = det(Hess) / |g|^4, K_h, k1, k2] at the origin. */
int pct_implicit_quadric_fit(const float* xyz, int64_t n, const int32_t* idx, int64_t nq, int k,
                             const int32_t* query_ids, const float* centered, double* coeffs, void* stream);
int pct_implicit_quadric_curvature(const double* coeffs, int64_t nq, double* curv, void* stream);

/* Text loader: replaces `np.loadtxt(file_path)` of read_from_file (ref :51) with a memory-mapped,
 * multi-threaded parser (host code).  One row per line; values separated by blanks or tabs (np.loadtxt's default delimiter: a comma is an error); '#'
 * starts a comment; empty lines are skipped; every row has the same number of columns
 * (PCT_ERR_INVALID_ARGUMENT otherwise, np.loadtxt raises ValueError).  Values are converted like Python's
 * float() (correctly rounded), so the table equals np.loadtxt's bit for bit.
 * pct_text_shape: rows and columns; pct_text_load: the rows x cols float64 table, `threads` <= 0 = all cores;
 * pct_text_load_f32: the same table rounded to float32 (`.astype(np.float32)` of ref :52-53) without the
 * float64 intermediate in memory -- what read_from_file keeps.  `out` may be pinned host memory. */
int pct_text_shape(const char* path, int64_t* rows, int64_t* cols);
int pct_text_load(const char* path, int64_t rows, int64_t cols, double* out, int threads);
int pct_text_load_f32(const char* path, int64_t rows, int64_t cols, float* out, int threads);

/* PCA estimators on given neighbour rows: replaces the per-point body of
 * principal_curvatures_via_principal_component_analysis (ref :901-945): covariance (np.cov, ddof = 1, fp64) of the
 * k listed points of every row (plus the query point itself when include_self != 0, as sklearn's
 * kneighbors(points) lists it, utils.py:812-815), its eigen-decomposition, and what the reference derives:
 * values nq x 6 = [l1 >= l2 >= l3, K = l1 * l2, H = (l1 + l2) / 2 (ref :935-936), l3 / (l1 + l2 + l3 + 1e-10)];
 * directions nq x 3 x 2 = eigenvectors of l1 and l2 as columns (ref :933; sign is arbitrary, as in eigh); may be NULL.
 * Rows come from pct_knn (the reference ranks all N points by fp32 distance per query, O(N^2)). */
int pct_pca_from_neighbors(const float* xyz, int64_t n, const int32_t* idx, int64_t nq, int k, int include_self,
                           const int32_t* query_ids, double* values, double* directions, void* stream);

/* Energy integration over a triangle mesh whose vertices carry the path's K and H: replaces
 * load_mesh_compute_energies (utils.py:702-765).  vertices n x 3 fp32, triangles t x 3 int32 (negative indices
 * count from the end like numpy's), gaussian / mean fp32 per vertex (NULL = zeros, utils.py:744-748), all on the
 * device.  out (device, 4 doubles): bending = nansum(mean(H^2 at the corners) * area), stretching =
 * nansum(mean(K at the corners) * area), total area, number of triangles with an index out of range (the
 * reference raises IndexError; such triangles contribute nothing here). */
int pct_mesh_energies(const float* vertices, int64_t n_vertices, const int32_t* triangles, int64_t n_triangles,
                      const float* gaussian, const float* mean, double* out, void* stream);

/* PLY body reader: replaces parse_ply (utils.py:979-1004) -- skip to the line "end_header", then float() of the
 * first three tokens of EVERY following line (extra columns are ignored, face lines are read as points, exactly
 * like the reference), rounded to float32.  A line with fewer than three numbers is an error (the reference
 * prints it and returns None).  pct_ply_shape: number of body lines and the byte offset of the body. */
int pct_ply_shape(const char* path, int64_t* rows, int64_t* body_offset);
int pct_ply_load_f32(const char* path, int64_t body_offset, int64_t rows, float* out /* rows x 3 */, int threads);

/* Writers (host code, all threads format blocks of rows, pwrite at each block's offset).
 * pct_write_points_ply: save_points_to_ply (utils.py:963-976): ASCII PLY header + np.savetxt '%.6f %.6f %.6f';
 *   `points` is n x 3 float32 (is_f64 = 0) or float64 (is_f64 = 1), host memory.
 * pct_write_curvature_ply: the output block of validate_shape (utils.py:538-551): header with
 *   gaussian_curvature / mean_curvature properties, rows f'{x} {y} {z} {K} {H}' of float32 scalars, i.e.
 *   repr(float(v)) of every value.  Both files are byte-identical to the reference's. */
int pct_write_points_ply(const char* path, const void* points, int is_f64, int64_t n, int threads);
int pct_write_curvature_ply(const char* path, const float* points, const float* gaussian, const float* mean, int64_t n,
                            int threads);

/* Host -> device copy on `stream` that does not depend on the source being page-locked: a pageable source (a plain
 * numpy array given to PointCloud(points=...), ref :26-47) is staged by several host threads through page-locked
 * double buffers, piece by piece, each piece's DMA queued as soon as it is staged.  Like cudaMemcpyAsync from
 * pageable memory, the call returns once the source has been read; the transfer completes in stream order. */
int pct_upload(void* dst_device, const void* src_host, int64_t bytes, void* stream);

/* Diagnostics: measured FP32 and FP64 FMA throughput (TFLOP/s, 2 flops per FMA) of the current device -- the
 * secondary roofline of SURVEY.md section 8(d); MEASURED_PEAKS.json has no CUDA-core figure.  Synchronises `stream`. */
int pct_measure_fma_peaks(double* fp32_tflops, double* fp64_tflops, void* stream);

/* Frees the per-stream scratch arenas the library keeps for the temporaries of its calls (they grow to the
 * largest call seen on a stream: about 37 bytes per point for an index build) and the page-locked staging buffers
 * of pct_upload.  Synchronises those streams. */
int pct_release_scratch(void);

/* Host-buffer convenience for callers without PyTorch: H2D + build + fused kNN
 * curvature + D2H, synchronous.  K, H: N fp32 host arrays. */
int pct_curvature_knn_host(const float* xyz_host, int64_t n, int k, float* K_host, float* H_host);

#ifdef __cplusplus
}
#endif
#endif /* PCT_B200_H */
