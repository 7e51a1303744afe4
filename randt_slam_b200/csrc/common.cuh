// Internal declarations shared by the CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/randt_gpu.h"

#include "schedule.hpp"

namespace randt {


struct LossParams {  // host-evaluated constants of the loss that do not depend on mu
  int kind;
  double a2;      // scale^2           (b = mu * a2)
  double alpha;
  double weight;
  double mu;      // used when no per-segment mu is given
  double fa, tf;  // Barron: |alpha - 2| / alpha and 2 / |alpha - 2|  (pre_factor = b * fa, times_s = tf / b)
};

// K3 evaluates "duos": one or two consecutive pairs that share their moving cell (the k neighbours of a moving cell are
// adjacent in the reference's residual-block order, R/src/ndt_registration/ndt_matcher.cpp:217-246).  One lane owns a duo, so
// the moving cell is loaded, converted and rotated once for both pairs.
struct Duo { uint32_t im, jf0, jf1, p0; };   // moving cell, fixed cell of pair p0, fixed cell of pair p0 + 1 (kNoCell: none), first pair
constexpr uint32_t kNoCell = 0xffffffffu;
// What K3 streams: the duo with its three cells inlined (moving, fixed of pair p0, fixed of pair p0 + 1) in the COMPACT form, 112 B.
// K3 only ever uses the symmetric part of a covariance, i.e. per off-diagonal couple (a, b) the fp64 sum s = (double)a + (double)b.  A
// cell is therefore stored as 9 floats — mean (x, y, i), diagonal (S00, S11, S22) and hs = float(s) rounded toward zero for the couples
// (01,10), (02,20), (12,21) — plus a 2-bit code per couple: s has at most 26 significant bits (a and b are a few float-ulps apart), so
// the bits of s are the bits of (double)hs with the code OR-ed in at bit 27.  The encoding is verified bit for bit when the table is
// built; a duo with a couple that does not fit (exponents far apart: an off-diagonal that is rounding noise around zero) is flagged
// kRecEscape and its word 0 indexes a full-precision record (9 x float4 = the three cells as stored) in the overflow table.
//   words 9 c + 0..8 : cell c = (mx, my, mi, S00, S11, S22, hs01, hs02, hs12), c = 0 moving, 1 fixed of p0, 2 fixed of p0 + 1
//   word 27          : bits 2 e, 2 e + 1 = code of couple e = 3 c + j;  kRecNoSecond: the duo has one pair only;  kRecEscape
struct __align__(16) DuoRec { float4 v[7]; };
struct __align__(16) DuoRecFull { float4 v[9]; };
constexpr uint32_t kRecEscape = 0x80000000u, kRecNoSecond = 0x40000000u;

struct DeviceProblem {
  const float4* cells_m;   // 3 x float4 per cell
  const float4* cells_f;
  const uint2* pairs;      // (im, jf), reference residual-block order
  const Duo* duos;         // the same pairs grouped two by two (shared moving cell)
  const DuoRec* duo_recs;  // [n_duos] record-major table K3 streams with bulk copies (compact form)
  const DuoRecFull* duo_overflow;  // full-precision records of the (rare) duos flagged kRecEscape
  const uint32_t* duo_p0;  // [n_duos] first pair of each duo (EMIT output rows)
  const uint32_t* seg_off; // [S+1] pair offsets per segment
  const ChunkDesc* chunks; // in the balanced order: warp w owns chunks [warp_off[w], warp_off[w+1])
  uint32_t n_chunks;
  const uint32_t* warp_off;  // [n_warps+1]
  uint32_t n_warps;
  uint32_t out_packed;       // FUSED only: 1 = write RANDT_PACKED_STRIDE-double records (upper triangle of H) instead of RANDT_FUSED_STRIDE, 2 = RANDT_CORE_STRIDE, 3 = RANDT_BASIS_STRIDE (slot totals, no chain rule)
  uint32_t plan_static;      // 1: chunks / warp_off are the problem's immutable schedule; 0: produced by a preceding kernel (solver re-plan)
  const uint32_t* seg_first_tile;  // [S+1]
  uint32_t n_segments;
  uint32_t n_pairs;
  double* partials;        // [n_tiles][kMaxAcc]
  uint32_t* seg_counters;  // [S] zero between launches
  const uint32_t* seg_active;  // [S] or NULL: segments with a zero flag are skipped by the fused kernel (batched solver)
};

#ifndef RANDT_K3_THREADS
#define RANDT_K3_THREADS 128
#endif
constexpr int kK3Threads = RANDT_K3_THREADS;   // threads per CTA in the pair-evaluation kernels
#ifndef RANDT_K3_MIN_CTAS
#define RANDT_K3_MIN_CTAS 4       // CTAs per SM the K3 register allocation is bounded for
#endif
constexpr int kK3MaxWarps = kSmCount * RANDT_K3_MIN_CTAS * (kK3Threads / 32);   // resident warps of the persistent grid
constexpr int kMaxAcc = 20;       // accumulators per tile partial (<= 10 H + 4 g + cost + max + sumsq + nonfinite)

// launchers (k3_pair_eval.cu)
cudaError_t launch_eval_fused(const DeviceProblem& p, int variant, const double* d_poses, const LossParams& lp, const double* d_mu,
                              bool want_jac, double* d_out, unsigned long long* d_bad, cudaStream_t s, int* n_launches);
cudaError_t launch_eval_emit(const DeviceProblem& p, int variant, const double* d_poses, double* d_r, double* d_J,
                             unsigned long long* d_bad, cudaStream_t s, int* n_launches);
cudaError_t launch_permute_duos(const Duo* in, const uint32_t* tile_rec_begin, const uint32_t* tile_duo_begin, uint32_t n_tiles, uint32_t n_duos,
                                Duo* out, cudaStream_t s, int* n_launches);
// d_n_overflow: device counter (zeroed by the caller) of the duos that needed a full record; records beyond overflow_cap are dropped
// (the caller re-runs with a larger table when the count exceeds the capacity)
// (tile_rec_begin / tile_duo_begin / n_tiles: the schedule-order permutation of the duos, or NULL: records in duo order)
cudaError_t launch_build_duo_records(const float4* cells_m, const float4* cells_f, const Duo* duos, uint32_t n_duos, const uint32_t* tile_rec_begin,
                                     const uint32_t* tile_duo_begin, uint32_t n_tiles, DuoRec* recs, uint32_t* duo_p0, DuoRecFull* overflow,
                                     uint32_t overflow_cap, uint32_t* d_n_overflow, cudaStream_t s, int* n_launches);
// the chunk lists of both plans from the tile assignment (schedule.hpp walk_warp, one thread per warp of the schedule)
cudaError_t launch_emit_chunks(const uint32_t* tile_seg, const uint32_t* first, const uint32_t* duo_off, uint32_t tile_duos, const uint32_t* mine,
                               const uint32_t* mine_off, const uint32_t* warp_rec_begin, const uint32_t* woffA, const uint32_t* woffB, uint32_t n_warps,
                               ChunkDesc* planA, ChunkDesc* planB, cudaStream_t s, int* n_launches);
cudaError_t launch_sweep_costs(const DeviceProblem& p, uint32_t pair_begin, uint32_t pair_end, int variant, const double* d_poses,
                               uint32_t n_poses, const LossParams& lp, double* d_cost, cudaStream_t s, int* n_launches);

// map geometry as the kernels need it
struct MapGeomDev {
  int size_x, size_y;
  uint32_t n_slots;
  double res, off_x, off_y;
  int r_stop;  // int(max_linf / res)
};
inline MapGeomDev make_geom(const randt_grid_params& g) {
  MapGeomDev m;
  m.size_x = g.size_x; m.size_y = g.size_y; m.n_slots = (uint32_t)g.size_x * (uint32_t)g.size_y;
  m.res = g.resolution;
  m.off_x = -static_cast<double>((unsigned)g.size_x) / 2.0 * g.resolution;
  m.off_y = -static_cast<double>((unsigned)g.size_y) / 2.0 * g.resolution;
  m.r_stop = static_cast<int>(g.max_linf / g.resolution);
  return m;
}

// k2_associate.cu
constexpr int kMaxNeighbours = 8;
cudaError_t launch_associate(const float4* cells_f, const uint32_t* cell_off_f, const int32_t* slot_f, const float4* cells_m,
                             const uint32_t* cell_off_m, uint32_t n_maps, uint32_t n_m_total, uint32_t max_m_per_map,
                             const MapGeomDev& geom, const float4* d_pose_f /*[B] (c,s,tx,ty)*/, int k, int metric,
                             uint32_t* d_nn /*[n_m_total*k]*/, uint32_t* d_cnt /*[n_m_total]*/, cudaStream_t s, int* n_launches);
// one map pair, everything (affine, association, pair / duo lists, K3 records, cell snapshots, totals) in one launch; d_totals[3] zeroed by
// the caller.  d_pairs == NULL: records only (no pair / duo lists, no snapshots).  d_layout != NULL: also the 6-word layout
// {seg_off[2], seg_duo_off[2], seg_first_tile, tile_rec_begin} that lets K7 solve this registration without a host round trip.
// d_n_m != NULL: the moving map's size is read from the device (chained behind K1); d_weight != NULL: receives ndt_weight / (n_m k).
cudaError_t launch_associate_single(const float4* cells_f, uint32_t n_f, const int32_t* slot_f, const float4* cells_m, uint32_t n_m,
                                    const MapGeomDev& geom, const double* d_pose0, int k, int metric, uint2* d_pairs, Duo* d_duos, DuoRec* d_recs,
                                    uint32_t* d_duo_p0, DuoRecFull* d_overflow, uint32_t overflow_cap, float4* d_snap_m, float4* d_snap_f,
                                    uint32_t* d_totals, uint32_t* d_layout, const uint32_t* d_n_m /* or NULL: n_m */, double ndt_weight,
                                    double* d_weight /* or NULL */, cudaStream_t s, int* n_launches);
cudaError_t launch_compact_pairs(const uint32_t* d_nn, const uint32_t* d_cnt, const uint32_t* d_scan /*exclusive scan of cnt*/,
                                 const uint32_t* cell_off_m, const uint32_t* cell_off_f, uint32_t n_maps, uint32_t n_m_total,
                                 uint32_t max_m_per_map, int k, uint2* d_pairs, cudaStream_t s, int* n_launches);
cudaError_t launch_compact_duos(const uint32_t* d_nn, const uint32_t* d_cnt, const uint32_t* d_scan /*pairs*/, const uint32_t* d_scan2 /*duos*/,
                                const uint32_t* cell_off_m, const uint32_t* cell_off_f, uint32_t n_maps, uint32_t n_m_total,
                                uint32_t max_m_per_map, int k, Duo* d_duos, cudaStream_t s, int* n_launches);
cudaError_t launch_duo_counts(const uint32_t* d_cnt, uint32_t n, uint32_t* d_cnt2, cudaStream_t s, int* n_launches);
cudaError_t launch_shift_part(const uint2* pairs_in, uint32_t n_pairs, const Duo* duos_in, uint32_t n_duos, uint32_t m_base, uint32_t f_base,
                              uint32_t p_base, uint2* pairs_out, Duo* duos_out, cudaStream_t s, int* n_launches);
cudaError_t launch_gather_offsets(const uint32_t* d_scan, const uint32_t* d_scan2, const uint32_t* cell_off, uint32_t n_off, uint32_t* d_out,
                                  cudaStream_t s, int* n_launches);
cudaError_t launch_exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out /*[n+1]*/, uint32_t n, uint32_t* d_block_sums, cudaStream_t s,
                                      int* n_launches);

// k1_voxelize.cu
enum { VOX_OK = 0, VOX_SPAN = 1, VOX_CELL_CAP = 2, VOX_OUT_OF_MAP = 3 };
cudaError_t launch_voxelize(const float4* d_pts, const uint32_t* d_scan_off, uint32_t n_scans, uint32_t max_pts_per_scan,
                            const randt_grid_params& gp, const MapGeomDev& geom, uint32_t cell_cap_per_scan, float4* d_cells_p,
                            uint32_t* d_npts_p, int32_t* d_labels_p, uint32_t* d_cell_count, int32_t* d_slot, unsigned short* d_bins_scratch /* n_points, or NULL */,
                            int* d_status, cudaStream_t s, int* n_launches);
bool voxelize_needs_bins_scratch(uint32_t max_pts_per_scan, uint32_t cell_cap_per_scan, const randt_grid_params& gp);
cudaError_t launch_compact_cells(const float4* d_cells_p, const uint32_t* d_npts_p, const int32_t* d_labels_p, const uint32_t* d_cell_off,
                                 uint32_t n_scans, uint32_t cell_cap_per_scan, uint32_t max_cells_per_scan, float4* d_cells,
                                 uint32_t* d_npts, int32_t* d_labels, cudaStream_t s, int* n_launches);
cudaError_t launch_build_slots(const float4* d_cells, const uint32_t* d_cell_off, uint32_t n_maps, uint32_t max_per_map,
                               const MapGeomDev& geom, int32_t* d_slot, cudaStream_t s, int* n_launches);
// AffineRec: 4 x float4 per map = (c, s, tx, ty) | (R00 R01 R02 R10) | (R11 R12 R20 R21) | (R22 0 0 0): the reference's Eigen::Affine2f and the
// Eigen Transform::rotation() of its 3-D lift (what Cell::transformCell rotates covariances with)
cudaError_t launch_prepare_affine(const void* d_in, int from_se2d, uint32_t n_maps, float4* d_aff, cudaStream_t s, int* n_launches);
cudaError_t launch_transform_cells(float4* d_cells, const uint32_t* d_cell_off, uint32_t n_maps, uint32_t max_per_map, const float4* d_aff,
                                   cudaStream_t s, int* n_launches);
cudaError_t launch_merge_maps(const float4* f_cells, const uint32_t* f_npts, const uint32_t* f_off, int32_t* f_slot, const float4* m_cells,
                              const uint32_t* m_npts, const uint32_t* m_off, uint32_t n_maps, const MapGeomDev& geom, const uint32_t* o_off,
                              float4* o_cells, uint32_t* o_npts, uint32_t* o_count, uint32_t max_m_per_map, cudaStream_t s, int* n_launches);

// k4_lm_step.cu — per-segment state of the batched GNC + LM solver
struct LmState {
  double x[4], cand[4];          // accepted point, candidate point
  double cost, g[4], H[16];      // at x: cost, tangent gradient, tangent J^T J (leading dimension 4)
  double scale[4], diag[4];      // Jacobi column scaling (fixed per solve), LM diagonal
  double radius, decrease_factor, x_norm, model_cost_change, gmax, min_cost;
  double mu, mu_first, final_cost;
  int32_t phase, iteration, solve_iterations, consecutive_invalid, last_successful, reuse_diagonal;
  int32_t gnc_solves, total_iterations, n_cost_evals, n_jac_evals, status, termination;
  uint32_t n_blocks, pad_;
};
cudaError_t launch_lm_init(uint32_t S, int np, const double* d_poses0, LmState* state, double* eval_pose, double* mu, uint32_t* active,
                           double* rec, uint32_t* n_active, cudaStream_t s, int* n_launches);
cudaError_t launch_lm_step(uint32_t S, int np, int use_manifold, const randt_solver_options& o, const double* rec, LmState* state,
                           double* eval_pose, double* mu, uint32_t* active, uint32_t* n_active, double* poses_out, double* result,
                           cudaStream_t s, int* n_launches);
cudaError_t launch_replan(const ChunkDesc* chunks, uint32_t n_chunks, const uint32_t* active, uint32_t n_warps, uint32_t* flags, uint32_t* scan,
                          uint32_t* block_sums, ChunkDesc* kept, uint32_t* warp_off, cudaStream_t s, int* n_launches);

// k7_solve.cu — persistent solver: one warp per registration, the whole GNC + LM solve in one launch
constexpr uint32_t kSolveMaxDuos = 512;   // registrations up to this many duos (1024 pairs) are solved by one warp
struct SolveLayout {
  const uint32_t* seg_duo_off;     // [S+1] duo offsets per segment
  const uint32_t* tile_rec_begin;  // [n_tiles] record offset of every tile, tiles in segment order
  uint32_t tile_duos;              // duos per full tile (multiple of 32)
  const uint32_t* items;           // [n_items] segments to solve, or NULL = segments 0 .. n_items-1
  uint32_t n_items;
  uint32_t* next_item;             // device counter, zero at launch
};
cudaError_t launch_solve_persistent(const DeviceProblem& p, const SolveLayout& L, int variant, int use_manifold, const LossParams& lp,
                                    const double* d_weight_per_seg /* [S] ScaledLoss weight per registration, or NULL: lp.weight */, const randt_solver_options& o, const double* d_poses0, double* d_poses_out, double* d_result,
                                    unsigned long long* d_bad, cudaStream_t s, int* n_launches);

// k8_allpairs.cu — every moving cell against every fixed cell (optionally within an L-infinity window), fused normal equations per map pair
void allpairs_tiles(uint32_t max_m, uint32_t max_f, uint32_t* tiles_m, uint32_t* tiles_f);
cudaError_t launch_allpairs(const float4* cells_f, const uint32_t* off_f, uint32_t max_f, const float4* cells_m, const uint32_t* off_m, uint32_t max_m,
                            uint32_t n_maps, int variant, const double* d_poses, const LossParams& lp, double window, double* d_partials,
                            uint32_t* d_tickets, double* d_out, unsigned long long* d_bad, cudaStream_t s, int* n_launches);

// k5_cs_divergence.cu
cudaError_t launch_cs_divergence(const float4* cells_f, const uint32_t* off_f, const float4* cells_m, const uint32_t* off_m, uint32_t n_maps,
                                 double* d_partials, uint32_t* d_tickets, double* d_out, cudaStream_t s, int* n_launches);
int cs_divergence_split();

// k6_filter_scan.cu
cudaError_t launch_filter_scans(const float4* d_raw, uint32_t n_scans, uint32_t n_az, uint32_t n_bins, const randt_filter_params& fp, uint32_t* d_peak,
                                float* d_angle, float4* d_out, uint32_t cap, uint32_t* d_counts, uint32_t* d_scan_off, uint32_t* d_block_sums,
                                int* d_status, cudaStream_t s, int* n_launches);
cudaError_t launch_filter_scan(const float4* d_raw, uint32_t n_az, uint32_t n_bins, const randt_filter_params& fp, uint32_t* d_peak, float* d_angle,
                               float4* d_out, uint32_t cap, uint32_t* d_n_out, int* d_status, cudaStream_t s, int* n_launches);

cudaError_t launch_pcl_xyzi_to_float4(const void* d_in32, uint32_t n, float4* d_out, cudaStream_t s, int* n_launches);

// static_cast<unsigned>(double) as x86-64 gcc defines it for negative inputs: truncate to int64, keep the low 32 bits
__host__ __device__ inline uint32_t to_u32_trunc(double v) {
  if (!(v > -9.2e18 && v < 9.2e18)) return 0u;
  return (uint32_t)(long long)v;
}
__host__ __device__ inline uint32_t coord_to_index(const MapGeomDev& g, float x, float y) {
  const uint32_t mx = to_u32_trunc(((double)x - g.off_x) / g.res);
  const uint32_t my = to_u32_trunc(((double)y - g.off_y) / g.res);
  return my * (uint32_t)g.size_x + mx;
}

#ifdef __CUDACC__
// Cell::transformCell (R/src/ndt_representation/ndt_cell.cpp:117-123) on the 3 x float4 cell layout.  Only meaningful in translation units
// compiled with -fmad=false (k1, k2): every product / sum is a separate IEEE operation in Eigen's order, x0 + (x1 + x2) per coefficient.
struct CellRaw { float4 a, b, c; };
__device__ __forceinline__ void transform_cell_affine(CellRaw& q, const float4* aff) {
  const float4 A = aff[0], R0 = aff[1], R1 = aff[2], R2 = aff[3];   // (plain loads: the fused association keeps the record in shared memory)
  const float c = A.x, s = A.y;
  const float R[3][3] = {{R0.x, R0.y, R0.z}, {R0.w, R1.x, R1.y}, {R1.z, R1.w, R2.x}};
  const float S[3][3] = {{q.a.w, q.b.x, q.b.y}, {q.b.z, q.b.w, q.c.x}, {q.c.y, q.c.z, q.c.w}};
  const float x = q.a.x, y = q.a.y, in = q.a.z;
  const float mx = A.z + (c * x + ((-s) * y + 0.f * in));
  const float my = A.w + (s * x + (c * y + 0.f * in));
  const float mi = 0.f + (0.f * x + (0.f * y + 1.f * in));
  float T[3][3], O[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[i][j] = R[i][0] * S[0][j] + (R[i][1] * S[1][j] + R[i][2] * S[2][j]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) O[i][j] = T[i][0] * R[j][0] + (T[i][1] * R[j][1] + T[i][2] * R[j][2]);
  q.a = make_float4(mx, my, mi, O[0][0]);
  q.b = make_float4(O[0][1], O[0][2], O[1][0], O[1][1]);
  q.c = make_float4(O[1][2], O[2][0], O[2][1], O[2][2]);
}
#endif

}  // namespace randt
