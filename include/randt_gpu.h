/* randt_gpu.h — C-ABI of the B200-native NDT scan-matching hot path.
 *
 * Drop-in boundary for IGMR-RWTH/RaNDT-SLAM's ndt_representation voxelisation and ndt_registration
 * cost/Jacobian evaluation.  The reference has no FFI/plugin mechanism; its seam is the
 * C++ surface of the catkin libraries `ndt_representation` / `ndt_registration` plus the Ceres
 * `CostFunction::Evaluate` interface.  Each entry point below names the reference interface it replaces
 * (R/ = ros/ndt_radar_slam/ in the reference tree).  The host-side C++ mirror of the reference classes
 * (randt_slam_b200/host/) and the Python test/bench bindings are thin layers over exactly these symbols.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.  All functions return 0 on success, <0 on error
 *    (RANDT_E_*); they never throw.  randt_last_error(ctx) returns a description of the last failure on ctx.
 *  - a randt_ctx owns one CUDA stream (or borrows the caller's); contexts are independent and may be used
 *    from different host threads concurrently (one thread per context at a time).  No global mutable state.
 *  - "cell" = 12 float32: mean (x, y, intensity) then row-major 3x3 covariance — the values
 *    Cell::new_mean_ / new_cov_ hold (R/include/ndt_representation/ndt_cell.h:167-168).
 *  - "pose" = 4 float64 in Sophus::SE2d::data() order [cos, sin, tx, ty] (R/src/ndt_registration/ndt_matcher.cpp:233,292)
 *    for the SE2 variants, or 3 float64 [x, y, theta] for the vector variants.
 *  - a randt_map / randt_problem carries scratch buffers for the calls made on it: use one object from one thread at a time (different
 *    objects may be used concurrently on different contexts).  Per-call device memory comes from the stream-ordered pool of the device.
 *  - *_dev entry points take device pointers, enqueue on the context stream and do not synchronise (randt_register_batch_dev is the
 *    exception: it polls the device and returns when every segment has finished).
 *  - kernels are launched with programmatic stream serialisation: a randt kernel may begin its prologue while the previous kernel of the
 *    stream drains, but reads caller-visible data only after that kernel has completed (griddepcontrol.wait).
 */
#ifndef RANDT_GPU_H
#define RANDT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RANDT_API __attribute__((visibility("default")))

enum {
  RANDT_OK = 0,
  RANDT_E_INVALID = -1,     /* bad argument */
  RANDT_E_CUDA = -2,        /* CUDA runtime error (see randt_last_error) */
  RANDT_E_CAPACITY = -3,    /* input exceeds a kernel capacity (e.g. label span, scan too large) */
  RANDT_E_NONFINITE = -4,   /* evaluation produced non-finite output (singular covariance sum, ...) */
  RANDT_E_NOMEM = -5
};

/* residual functor variants — R/include/ndt_registration/ceres_residuals.h */
enum {
  RANDT_VAR_SE2_INTENSITY = 0,  /* NDTFrameToMapIntensityFactorResidualSE2  :520-552  (live in every shipped config) */
  RANDT_VAR_SE2_XY = 1,         /* NDTFrameToMapFactorResidualSE2           :454-484 */
  RANDT_VAR_VEC_INTENSITY = 2,  /* NDTFrameToMapIntensityFactorResidual     :486-518  params [x, y | theta] */
  RANDT_VAR_VEC_XY = 3          /* NDTFrameToMapFactorResidual              :421-451  params [x, y | theta] */
};

/* robust loss — R/include/ndt_registration/ceres_loss_functions.h:9-48, wrapped in ceres::ScaledLoss(loss, weight)
 * as R/src/ndt_registration/ndt_matcher.cpp:392,479 does. */
enum { RANDT_LOSS_NONE = 0, RANDT_LOSS_BARRON = 1, RANDT_LOSS_WELSCH = 2 };
typedef struct randt_loss {
  int32_t kind;
  double scale;      /* a      : loss_function_scale */
  double alpha;      /* alpha  : loss_function_convexity (Barron) */
  double mu;         /* GNC control parameter (>= 1 in the reference's schedule) */
  double weight;     /* ScaledLoss factor, e.g. ndt_weight / (n_cells * k) */
} randt_loss;

/* nearest-neighbour metric of Map::getClosestCells — R/src/ndt_representation/ndt_map.cpp:101-151 */
enum { RANDT_LOOKUP_MAHALANOBIS_INTENSITY = 0, RANDT_LOOKUP_EUCLID_XY = 1 };

/* voxel grid + map geometry: RadarPreprocessorParameters / NDTMapParameters after NDTSlam::readParameters
 * (R/include/ndt_slam/ndt_slam_parameters.h:17-50, R/src/ndt_slam/ndt_slam.cpp:653-654,691) */
typedef struct randt_grid_params {
  float max_range;        /* radar_preprocessor/max_range (float, as ClusterGenerator::max_sensor_range_) */
  int32_t n_clusters;     /* int((2*max_range/resolution)^2) */
  int32_t min_points;     /* ndt_map/min_points_per_cell; a cell needs n > min_points */
  int32_t size_x, size_y; /* map size in cells (after the int /= resolution truncation) */
  double resolution;      /* ndt_map/resolution */
  double max_linf;        /* ndt_map/max_neighbor_linf_distance */
} randt_grid_params;

/* one fused evaluation result per segment (= per pose): 24 float64 */
enum {
  RANDT_FUSED_H = 0,       /* [16] row-major 4x4  sum J~^T J~  (ambient parameters; 3x3 in the top-left for VEC variants) */
  RANDT_FUSED_G = 16,      /* [4]  sum J~^T r~ */
  RANDT_FUSED_COST = 20,   /* sum 0.5 * rho(r^2)  (what ceres reports as cost for these blocks) */
  RANDT_FUSED_MAXR = 21,   /* max raw residual (GNC mu seed, ndt_matcher.cpp:387) */
  RANDT_FUSED_SUMSQ = 22,  /* sum raw r^2 */
  RANDT_FUSED_N = 23,      /* residual blocks in the segment */
  RANDT_FUSED_STRIDE = 24
};
/* the same record without the redundant half of the symmetric H: 18 float64 (what randt_eval_fused_async can ship over PCIe) */
enum {
  RANDT_PACKED_H = 0,      /* [10] upper triangle of the 4x4, row by row: 00 01 02 03 11 12 13 22 23 33 */
  RANDT_PACKED_G = 10,     /* [4] */
  RANDT_PACKED_COST = 14, RANDT_PACKED_MAXR = 15, RANDT_PACKED_SUMSQ = 16, RANDT_PACKED_N = 17,
  RANDT_PACKED_STRIDE = 18,
  RANDT_CORE_STRIDE = 15,  /* packed == 2: the first 15 entries of the packed record (H upper triangle, g, cost) — what one LM iteration consumes */
  /* packed == 3: the normal equations in the functor's own three-dimensional derivative basis, before any chain rule — 80 bytes per
   * segment, the least a host-side LM iteration needs.  Basis of RANDT_VAR_SE2_INTENSITY: (theta, tx, ty) with theta = atan2(s, c), i.e.
   * the ambient record is F^T H_b F, F^T g_b with F = [[-s/n2, c/n2, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], n2 = c^2 + s^2, and the
   * Sophus::Manifold<SE2> tangent record (ux, uy, theta) is Q^T H_b Q, Q^T g_b with Q = [[0, 0, 1], [c, -s, 0], [s, c, 0]] (unit c, s);
   * basis of the VEC variants: their parameters (x, y, theta).  Not available for RANDT_VAR_SE2_XY (four-dimensional basis) or want_jac == 0. */
  RANDT_BASIS_H = 0,       /* [6] upper triangle of the 3x3, row by row: 00 01 02 11 12 22 */
  RANDT_BASIS_G = 6,       /* [3] */
  RANDT_BASIS_COST = 9,
  RANDT_BASIS_STRIDE = 10
};

typedef struct randt_ctx randt_ctx;
typedef struct randt_map randt_map;          /* a batch of B NDT maps resident on the device */
typedef struct randt_problem randt_problem;  /* cell tables + frozen pair list (construction-time state of addNDTFactor) */

/* ---- context ------------------------------------------------------------------------------------------ */
RANDT_API int randt_version(void);
/* stream: a cudaStream_t to borrow (e.g. the caller's current stream), or NULL to create an owned non-blocking stream. */
RANDT_API int randt_ctx_create(int device, void* stream, randt_ctx** out);
RANDT_API void randt_ctx_destroy(randt_ctx* ctx);
RANDT_API const char* randt_last_error(const randt_ctx* ctx);
RANDT_API void* randt_ctx_stream(const randt_ctx* ctx);
RANDT_API int randt_ctx_sync(randt_ctx* ctx);
/* number of kernels this context has launched since creation (bench "gpu_launches") */
RANDT_API uint64_t randt_ctx_launch_count(const randt_ctx* ctx);

/* returns (and clears) the number of degenerate pairs K3 has met on this context: pairs whose d^T B^-1 d is negative or
 * non-finite (singular covariance sum).  They are excluded from the fused sums and emitted as NaN by randt_eval_emit. */
RANDT_API int randt_ctx_take_bad_pairs(randt_ctx* ctx, uint64_t* count);
/* pinned host / device buffers and stream-ordered copies for callers that stage their own data (bench, host adapters) */
RANDT_API void* randt_host_alloc(size_t bytes);
RANDT_API void randt_host_free(void* p);
RANDT_API void* randt_dev_alloc(size_t bytes);
RANDT_API void randt_dev_free(void* p);
RANDT_API int randt_memcpy_h2d(randt_ctx* ctx, void* dst, const void* src, size_t bytes);
RANDT_API int randt_memcpy_d2h(randt_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---- K6: raw-scan peak filter ---------------------------------------------------------------------------
 * Replaces RadarPreprocessor::filterScan (R/src/radar_preprocessing/radar_preprocessor.cpp:45-125): per-azimuth strongest return, the falling
 * flanks around it, range / intensity gates, sensor -> base transform.  raw4: float32 [n_azimuths * n_bins][4] = (x, y, z, intensity) in
 * the sensor frame, azimuth-major (the organised cloud cloud_in->height x width carries).  out4: filtered points (x, y, z, intensity) in the
 * base frame, in the reference's order — exactly randt_voxelize()'s input.  *_on_device != 0: the pointer is a device pointer.
 * Returns RANDT_E_INVALID if the reference's angle rule (new azimuth when |atan2(y, x) - first angle| > 1e-4) would not cut the scan at the
 * given rows, RANDT_E_CAPACITY if more than `cap` points survive (n_out then holds the required count). */
typedef struct randt_filter_params {
  float min_range, max_range, min_intensity;   /* radar_preprocessor/{min_range, max_range, min_intensity} (float members, radar_preprocessor.h:59-61) */
  float pad_;
  double beam_distance_increment_threshold;     /* radar_preprocessor/beam_distance_increment_threshold */
  float sensor_to_base[12];                     /* row-major 3x4 of initial_transform_radar_baselink (Eigen::Affine3f) */
} randt_filter_params;
/* pcl::PointXYZI records as PCL lays them out (32 bytes: x, y, z, 1 | intensity, 3 pad floats; what `cloud.points.data()` points at in
 * RadarPreprocessor::filterScan's input, radar_preprocessor.cpp:45-51) -> the path's (x, y, 0, intensity) float4 points in DEVICE memory
 * d_out4 (n * 16 bytes), ready for randt_filter_scan(raw_on_device = 1) or randt_voxelize(pts_on_device = 1): the caller hands the PCL
 * buffer over as it is instead of repacking 1.2 M points per scan on the host.  pcl_points: host (on_device = 0) or device memory. */
RANDT_API int randt_points_from_pcl_xyzi(randt_ctx* ctx, const void* pcl_points, uint32_t n, int on_device, float* d_out4);
RANDT_API int randt_filter_scan(randt_ctx* ctx, const float* raw4, uint32_t n_azimuths, uint32_t n_bins, const randt_filter_params* params,
                                int raw_on_device, float* out4, int out_on_device, uint32_t cap, uint32_t* n_out);
/* The same for n_scans scans of one shape laid back to back in raw4 (several sequences replayed side by side, or a backlog of one):
 * every scan is filtered independently, exactly as randt_filter_scan would; out4 receives the kept points scan after scan and
 * scan_off[n_scans + 1] (host) their offsets — together the input of randt_voxelize().  cap: capacity of out4 in points over all scans
 * (RANDT_E_CAPACITY: scan_off[n_scans] then holds the required count). */
RANDT_API int randt_filter_scans(randt_ctx* ctx, const float* raw4, uint32_t n_scans, uint32_t n_azimuths, uint32_t n_bins,
                                 const randt_filter_params* params, int raw_on_device, float* out4, int out_on_device, uint32_t cap,
                                 uint32_t* scan_off);

/* ---- K1: voxelisation ---------------------------------------------------------------------------------
 * Replaces Grid::cluster (R/src/radar_preprocessing/grid.cpp:7-14), ClusterGenerator::labelClouds
 * (R/src/radar_preprocessing/radar_preprocessor.cpp:151-169), Map::insertCluster (R/src/ndt_representation/ndt_map.cpp:238-245)
 * and Cell::addPointCloud/updateCell (R/src/ndt_representation/ndt_cell.cpp:25-114) — i.e. what
 * HierarchicalMap::addClusters does per scan (R/src/local_fuser/local_fuser.cpp:102-105).
 * pts4: float32 [n_pts_total][4] = (x, y, z-unused, intensity); scan b owns points [scan_off[b], scan_off[b+1]).
 * pts_on_device != 0: pts4 is a device pointer (scan_off is always a host pointer). */
RANDT_API int randt_voxelize(randt_ctx* ctx, const float* pts4, const uint32_t* scan_off, uint32_t n_scans,
                             const randt_grid_params* params, int pts_on_device, randt_map** out);

/* ---- maps ---------------------------------------------------------------------------------------------- */
/* Upload B maps built elsewhere.  cell_off[B+1]; npts may be NULL (treated as 0); slot may be NULL, in which case the
 * dense lookup table grid_indizes_ is rebuilt from the cell means in cell order (later cell wins, as insertCluster does). */
RANDT_API int randt_map_upload(randt_ctx* ctx, const float* cells, const uint32_t* npts, const uint32_t* cell_off, uint32_t n_maps,
                               const int32_t* slot, const randt_grid_params* params, randt_map** out);
RANDT_API int randt_map_info(const randt_map* map, uint32_t* n_maps, uint32_t* n_cells_total, uint32_t* n_slots);
/* any output pointer may be NULL.  cells [n_cells_total][12], npts [n_cells_total], labels [n_cells_total] (voxelised maps only),
 * cell_off [n_maps+1], slot [n_maps][n_slots] */
RANDT_API int randt_map_download(randt_ctx* ctx, const randt_map* map, float* cells, uint32_t* npts, int32_t* labels,
                                 uint32_t* cell_off, int32_t* slot);
/* Map::transformMap / Cell::transformCell (R/src/ndt_representation/ndt_map.cpp:177-182, ndt_cell.cpp:117-123).
 * trans: float32 [n_maps][4] = (cos, sin, tx, ty) of the Eigen::Affine2f the caller holds, on the host.  Means move by the affine as given;
 * covariances by Eigen's Transform::rotation() of it (the SVD polar factor, as transformCell computes it: equal to the linear part only up
 * to float rounding).  The slot table is NOT updated (neither is the reference's). */
RANDT_API int randt_map_transform(randt_ctx* ctx, randt_map* map, const float* trans);
/* The same for poses held as Sophus::SE2d: poses float64 [n_maps][4] = [cos, sin, tx, ty]; the affine is
 * Eigen::Affine2f(pose.cast<float>().matrix()) as the reference builds it (R/src/local_fuser/local_fuser.cpp:175,280,338): Sophus' cast
 * re-normalises the float unit complex. */
RANDT_API int randt_map_transform_se2d(randt_ctx* ctx, randt_map* map, const double* poses);
/* Map::mergeMapCell + Cell::operator+= (R/src/ndt_representation/ndt_map.cpp:191-207, ndt_cell.h:133-142): merge moving map b
 * into fixed map b for every b.  `fixed` is rebuilt in place (cell order: existing cells, then appended cells in moving order). */
RANDT_API int randt_map_merge(randt_ctx* ctx, randt_map* fixed, const randt_map* moving);
RANDT_API void randt_map_destroy(randt_map* map);

/* Map::calculateCSDivergence (R/src/ndt_representation/ndt_map.cpp:42-99): Cauchy-Schwarz divergence between fixed map b and moving map b
 * (already transformed into the fixed frame, as R/src/local_fuser/local_fuser.cpp:338-339 does) for every b; out: host float64 [n_maps].
 * The reference leaves its three accumulators uninitialised; they start at zero here. */
RANDT_API int randt_cs_divergence(randt_ctx* ctx, const randt_map* fixed, const randt_map* moving, double* out);

/* ---- K2: association ----------------------------------------------------------------------------------
 * Replaces the association half of Matcher::addNDTFactor (R/src/ndt_registration/ndt_matcher.cpp:200-217,249-253):
 * transform each moving cell by the initial guess (float32), Map::getClosestCells (ndt_map.cpp:101-151,163-175) with the
 * expanding window, Mahalanobis (ndt_cell.cpp:172-176) or Euclidean metric, first k by (distance, index).
 * Problem b pairs moving map b with fixed map b at pose0[b] (float64 [n_maps][4], host).  The resulting problem has one
 * segment per map; pairs are stored in the reference's residual-block order. */
RANDT_API int randt_associate(randt_ctx* ctx, const randt_map* fixed, const randt_map* moving, const double* pose0, int k, int metric,
                              randt_problem** out);

/* ---- problems -------------------------------------------------------------------------------------------
 * A problem snapshots the cell tables (like the functors copy their cells, ceres_residuals.h:528-535) and a frozen pair list.
 * pair_m / pair_f index rows of cells_m / cells_f; segment s owns pairs [seg_off[s], seg_off[s+1]) and is evaluated at pose s. */
RANDT_API int randt_problem_create(randt_ctx* ctx, const float* cells_m, uint32_t n_m, const float* cells_f, uint32_t n_f,
                                   const uint32_t* pair_m, const uint32_t* pair_f, uint32_t n_pairs, const uint32_t* seg_off,
                                   uint32_t n_segments, randt_problem** out);
/* Joins problems on the device (no host copy of their tables): part i contributes its cell snapshots, pairs and duos with shifted indices;
 * its pairs go to segment seg_of_part[i] of the result (seg_of_part non-decreasing, every part a single-segment problem of this context's
 * device).  This is how the residual blocks of several addNDTFactor calls become one problem: the blocks of every window state of
 * Matcher::estimateTransformCeres (ndt_matcher.cpp:356-360: per state, one call per fixed map) -> one segment per state. */
RANDT_API int randt_problem_concat(randt_ctx* ctx, const randt_problem* const* parts, uint32_t n_parts, const uint32_t* seg_of_part,
                                   uint32_t n_segments, randt_problem** out);
RANDT_API int randt_problem_info(const randt_problem* p, uint32_t* n_segments, uint32_t* n_pairs, uint32_t* n_m, uint32_t* n_f);
/* how K3 holds the problem in HBM: records of record_bytes each (one per duo = two pairs sharing their moving cell), of which
 * n_overflow needed a full-precision side record; any pointer may be NULL */
RANDT_API int randt_problem_layout(const randt_problem* p, uint32_t* n_duos, uint32_t* record_bytes, uint32_t* n_overflow);
/* the work schedule K3 walks (introspection; tests compare it with the host builder in csrc/schedule.hpp).  counts [5]: warps of the
 * schedule, tiles, chunks of plan A (split chunks allowed), chunks of plan B, the warp budget the schedule was built for.  plan_a4 /
 * plan_b4: the chunk descriptors as 4 x uint32 {duo_begin, meta, seg, part} (at most cap_chunks each); woff_a / woff_b: [warps + 1]
 * chunk ranges per warp (at most cap_warps + 1 each); duo_off: [n_segments + 1].  Any pointer but counts may be NULL. */
RANDT_API int randt_problem_schedule(randt_ctx* ctx, const randt_problem* p, uint32_t* counts, uint32_t* plan_a4, uint32_t* plan_b4,
                                     uint32_t cap_chunks, uint32_t* woff_a, uint32_t* woff_b, uint32_t cap_warps, uint32_t* duo_off);
/* any pointer may be NULL */
RANDT_API int randt_problem_download(randt_ctx* ctx, const randt_problem* p, uint32_t* pair_m, uint32_t* pair_f, uint32_t* seg_off);
/* the snapshotted cell tables: cells_m [n_m][12], cells_f [n_f][12]; either may be NULL */
RANDT_API int randt_problem_download_cells(randt_ctx* ctx, const randt_problem* p, float* cells_m, float* cells_f);
RANDT_API void randt_problem_destroy(randt_problem* p);

/* ---- K3: residual / Jacobian evaluation ---------------------------------------------------------------
 * Replaces ceres::AutoDiffCostFunction<NDTFrameToMap*Residual*, 1, ...>::Evaluate over every residual block of the problem
 * (hot loop C), the per-block robust-loss Corrector, and the normal-equation accumulation.
 *
 * EMIT: r[n_pairs] and J[n_pairs][np] (np = 4 SE2 variants, 3 VEC variants), raw (no loss) — exactly what Evaluate returns per
 * block.  J == NULL gives the residual-only pass used for the GNC seed (ndt_matcher.cpp:382-387). */
RANDT_API int randt_eval_emit(randt_ctx* ctx, const randt_problem* p, int variant, const double* poses, double* r, double* J);
RANDT_API int randt_eval_emit_dev(randt_ctx* ctx, const randt_problem* p, int variant, const double* d_poses, double* d_r, double* d_J);
/* FUSED: per segment, loss-corrected J~^T J~, J~^T r~, cost, max raw residual (layout RANDT_FUSED_*).  mu_per_seg (may be NULL)
 * overrides loss->mu per segment.  want_jac == 0 computes cost/max/sumsq only (trust-region candidate evaluation, BnB sweep). */
RANDT_API int randt_eval_fused(randt_ctx* ctx, const randt_problem* p, int variant, const double* poses, const randt_loss* loss,
                               const double* mu_per_seg, int want_jac, double* out);
/* Enqueue-only form for a caller that scores one pose set after another (the BnB levels of ndt_matcher.cpp:560-576, pose grids):
 * poses, mu_per_seg and out must be pinned host memory (randt_host_alloc); the upload of call i+1 and the copy-out of call i-1 overlap
 * the kernel of call i (two copy streams, two device slots each way).  The buffers of a call may be read / reused after
 * randt_ctx_sync(), or once call i+2 has left the context's stream; give the calls in flight their own buffers.
 * packed == 1: `out` receives RANDT_PACKED_STRIDE doubles per segment (layout RANDT_PACKED_*) instead of RANDT_FUSED_STRIDE; packed == 2:
 * RANDT_CORE_STRIDE doubles (the packed record without max r, sum r^2, n: the normal equations and the cost, 120 bytes); packed == 3:
 * RANDT_BASIS_STRIDE doubles (layout RANDT_BASIS_*: the same normal equations before the chain rule to the ambient parameters, 80 bytes) — the
 * copy-out is what bounds a pipelined step, and a quarter of the full record is the mirrored half of H. */
RANDT_API int randt_eval_fused_async(randt_ctx* ctx, const randt_problem* p, int variant, const double* poses, const randt_loss* loss,
                                     const double* mu_per_seg, int want_jac, int packed, double* out);
/* Completion of individual async calls: randt_ctx_async_count() = number of randt_eval_fused_async calls issued on this context so far
 * (the ticket of the latest one); randt_ctx_wait_async(ticket) returns once that call's records are in host memory and its input
 * buffers may be reused (it may wait for a later call of the same parity: calls complete in order). */
RANDT_API uint64_t randt_ctx_async_count(const randt_ctx* ctx);
RANDT_API int randt_ctx_wait_async(randt_ctx* ctx, uint64_t ticket);
RANDT_API int randt_eval_fused_dev(randt_ctx* ctx, const randt_problem* p, int variant, const double* d_poses, const randt_loss* loss,
                                   const double* d_mu_per_seg, int want_jac, double* d_out);
/* All-pairs evaluation (K8): every moving cell of map b against every fixed cell of map b — optionally only pairs whose fixed mean lies
 * within `window` metres (L-infinity) of the transformed moving mean; window <= 0: all N_m x N_f pairs — with the same functor, loss
 * corrector and per-pose reduction as randt_eval_fused (record layout RANDT_FUSED_*, entry n = pairs used).  The reference keeps only the k
 * nearest fixed cells per moving cell (Matcher::addNDTFactor, ndt_matcher.cpp:183-288); this is the "every overlapping submap cell" reading
 * of the cost (BASELINE.json north_star), a cache-resident, fp64-bound workload.  poses [n_maps][np], out [n_maps][24]. */
RANDT_API int randt_eval_allpairs(randt_ctx* ctx, const randt_map* fixed, const randt_map* moving, int variant, const double* poses,
                                  const randt_loss* loss, double window, double* out);
RANDT_API int randt_eval_allpairs_dev(randt_ctx* ctx, const randt_map* fixed, const randt_map* moving, int variant, const double* d_poses,
                                      const randt_loss* loss, double window, double* d_out);

/* Cost of ONE segment's pair list at many candidate poses (the inner loop of Matcher::estimateTransformGlobalBNB,
 * ndt_matcher.cpp:560-576): cost[i] = sum over the pairs of segment `seg` of 0.5 * rho(r^2) at poses[i]. */
RANDT_API int randt_sweep_costs(randt_ctx* ctx, const randt_problem* p, uint32_t seg, int variant, const double* poses, uint32_t n_poses,
                                const randt_loss* loss, double* cost);

/* ---- K4: batched GNC + Levenberg-Marquardt registration ------------------------------------------------
 * Replaces, for every segment (= one registration) of the problem at once, the solve half of Matcher::estimateLoopConstraint
 * (R/src/ndt_registration/ndt_matcher.cpp:457-492; the same loop closes Matcher::estimateTransformCeres, :372-397): raw residual
 * maximum -> mu0 = min(2 max_r^2 / gnc_loss_scale^2, gnc_divisor^(gnc_max_steps-1)); do { mu = max(mu, 1); ceres::Solve (trust
 * region, LEVENBERG_MARQUARDT, dense) with loss ScaledLoss(Barron(loss->scale, loss->alpha, mu), loss->weight); mu /= gnc_divisor }
 * while (mu > 1/sqrt(gnc_divisor)).  Defaults of the solver fields are ceres 2.1.0's Solver::Options defaults plus the
 * reference's max_num_iterations (ndt_radar_slam_base_parameters.yaml: max_iteration). */
typedef struct randt_solver_options {
  int32_t max_num_iterations;                 /* 200 */
  int32_t use_manifold;                       /* 1: Sophus::Manifold<SE2> on the 4-parameter pose (estimateTransformCeres, ndt_matcher.cpp:334);
                                                 0: raw ambient parameters — what estimateLoopConstraint ends up optimising */
  int32_t max_num_consecutive_invalid_steps;  /* 5 */
  int32_t jacobi_scaling;                     /* 1 */
  double function_tolerance;                  /* 1e-6 */
  double gradient_tolerance;                  /* 1e-10 */
  double parameter_tolerance;                 /* 1e-8 */
  double initial_trust_region_radius;         /* 1e4 */
  double max_trust_region_radius;             /* 1e16 */
  double min_trust_region_radius;             /* 1e-32 */
  double min_lm_diagonal, max_lm_diagonal;    /* 1e-6, 1e32 */
  double min_relative_decrease;               /* 1e-3 */
  double gnc_loss_scale;                      /* ndt_matcher/loss_function_scale (the scale in mu0, ndt_matcher.cpp:386,473) */
  double gnc_divisor;                         /* ndt_matcher/gnc_control_parameter_divisor */
  int32_t gnc_max_steps;                      /* gnc_steps / loop_closure gnc steps */
  int32_t poll_interval;                      /* >= 0: registrations of up to 1024 pairs are solved start to finish by the persistent kernel (one
                                                 warp each, one launch for the batch); longer ones step through one evaluation + one
                                                 solver launch per LM iteration, the host polling the active count every poll_interval
                                                 iterations (0: default 4).  < 0: the stepwise path for every registration (diagnostic). */
} randt_solver_options;
RANDT_API void randt_solver_options_default(randt_solver_options* o);

/* per-segment result record of randt_register_batch: 8 float64 */
enum {
  RANDT_REG_SCORE = 0,        /* summary.final_cost / num_residual_blocks of the last solve (return value of estimateLoopConstraint, :492) */
  RANDT_REG_FINAL_COST = 1,
  RANDT_REG_GNC_SOLVES = 2,   /* ceres::Solve calls */
  RANDT_REG_ITERATIONS = 3,   /* minimiser iterations over all solves (iteration 0 of each solve included) */
  RANDT_REG_EVALS = 4,        /* cost + Jacobian evaluations ceres would have made */
  RANDT_REG_MU_FIRST = 5,     /* mu0 before the max(mu, 1) clamp */
  RANDT_REG_STATUS = 6,       /* 0 ok, 1 the segment has no residual blocks ("NO RESIDUALS ADDED", :454-456): pose returned unchanged */
  RANDT_REG_TERMINATION = 7,  /* last solve: 0 convergence, 1 iteration limit, 2 failure */
  RANDT_REG_STRIDE = 8
};
/* poses: [n_segments][np] initial guesses in, refined poses out; result: [n_segments][RANDT_REG_STRIDE].  loss->mu is ignored (the GNC
 * schedule sets it per segment).  The whole batch advances in lock step: one K3 launch + one K4 launch per LM iteration over the
 * segments still active; nothing but an active counter crosses PCIe until every segment has finished.  *_dev: device pointers. */
RANDT_API int randt_register_batch(randt_ctx* ctx, const randt_problem* p, int variant, double* poses, const randt_loss* loss,
                                   const randt_solver_options* opt, double* result);
RANDT_API int randt_register_batch_dev(randt_ctx* ctx, const randt_problem* p, int variant, double* d_poses, const randt_loss* loss,
                                       const randt_solver_options* opt, double* d_result);
/* randt_register_batch with one ScaledLoss weight per registration (host [n_segments]; NULL = loss->weight for all): the odometry weight
 * ndt_weight / (n_cells k) depends on the moving scan's own cell count (ndt_matcher.cpp:367,392), so a batch of scans solved together
 * gives each scan exactly the weight it would get alone.  Needs registrations of <= 1024 pairs (the persistent solver). */
RANDT_API int randt_register_batch_weighted(randt_ctx* ctx, const randt_problem* p, int variant, double* poses, const randt_loss* loss,
                                            const double* weight_per_seg, const randt_solver_options* opt, double* result);

/* One scan of a live stream in one call (the per-scan chain of LocalFuser::processScan, R/src/local_fuser/local_fuser.cpp:95-190, for the
 * parts that are on this path): voxelise the filtered scan (K1) -> associate it with the submap at pose_io (K2, addNDTFactor) ->
 * GNC + LM registration (K7; loss ScaledLoss(Barron(loss->scale, loss->alpha, mu), ndt_weight / (n_cells k)), ndt_matcher.cpp:392) ->
 * if insert_keyframe != 0: transformMap by the estimate + mergeMapCell into the submap (local_fuser.cpp:164-190).  An EMPTY submap
 * (created by randt_map_upload with no cells) takes the scan as its first keyframe at pose_io.  pose_io [4]: prior in, estimate out;
 * result [RANDT_REG_STRIDE] (may be NULL); n_cells_out: cells of the scan (may be NULL).  pts4 is host memory. */
RANDT_API int randt_scan_step(randt_ctx* ctx, randt_map* submap, const float* pts4, uint32_t n_pts, const randt_grid_params* gp, int k, int metric,
                              const randt_loss* loss, double ndt_weight, const randt_solver_options* opt, int insert_keyframe, double* pose_io,
                              double* result, uint32_t* n_cells_out);

#ifdef __cplusplus
}
#endif
#endif /* RANDT_GPU_H */
