// randt_host.hpp — C++17 host-side mirror of the RaNDT-SLAM surface for the NDT hot path, over the C-ABI in randt_gpu.h.
//
// The reference (R/ = ros/ndt_radar_slam/) is compiled C++, so the host side above the C-ABI is C++ with the reference's class
// and method names, argument meaning and error behaviour for this path:
//
//   randt::Map              <- ndt_representation::Map / HierarchicalMap      R/include/ndt_representation/ndt_map.h:14-199,
//                                                                             ndt_hierarchical_map.h:34-69 (addClusters, transformMap, mergeMapCell)
//   randt::Matcher          <- ndt_registration::Matcher                      R/include/ndt_registration/ndt_matcher.h:46-87
//   randt::NdtCostFunction  <- the N x AutoDiffCostFunction<NDTFrameToMap...Residual..., 1, 4> blocks Matcher::addNDTFactor adds
//                              (R/src/ndt_registration/ndt_matcher.cpp:183-288), as ONE ceres::CostFunction
//   randt::SE2d             <- Sophus::SE2d (storage order [cos, sin, tx, ty], group product, exp / log of the rotation)
//
// No Eigen / Sophus / Ceres / PCL / ROS types appear here (none exist in this image); INTEGRATION.md shows the few lines that
// adapt them in the reference tree.  The classes are exported from librandt_host.so (link with -lrandt_host -lrandt_gpu).  Every method throws randt::Error (std::runtime_error) on a C-ABI failure — there is no CPU
// fallback.
#pragma once
#include <cstddef>
#include <cstdint>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "randt_gpu.h"

#if __has_include(<ceres/cost_function.h>)
#include <ceres/cost_function.h>
#else
// Stand-in with the exact interface of ceres::CostFunction (Ceres 2.1.0 include/ceres/cost_function.h) so that NdtCostFunction
// compiles here and drops into a real ceres::Problem unchanged where Ceres is installed.
namespace ceres {
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  CostFunction(const CostFunction&) = delete;
  void operator=(const CostFunction&) = delete;
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};
}  // namespace ceres
#endif

namespace randt {

struct RANDT_API Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// Sophus::SE2d stand-in: unit complex number + translation, data() in Sophus' storage order.
struct RANDT_API SE2d {
  double v[4] = {1.0, 0.0, 0.0, 0.0};   // cos, sin, tx, ty
  SE2d() {}
  SE2d(double theta, double tx, double ty);
  double* data() { return v; }
  const double* data() const { return v; }
  double angle() const;                  // so2().log()
  SE2d operator*(const SE2d& o) const;   // group product with Sophus' conditional renormalisation of the complex number
  void matrix3f(float out[9]) const;     // column-major 3x3 homogeneous matrix as floats (what the BnB search de-duplicates on)
  static SE2d exp(double ux, double uy, double theta);   // Sophus::SE2d::exp of the screw (ux, uy, theta)
};

// rc::navigation::ndt::State (R/include/ndt_slam/trajectory_representation.h:12-22): a window state in both representations
struct RANDT_API State {
  SE2d pose;
  double pos[2] = {0.0, 0.0};
  double rot = 0.0;
  double lin_vel[2] = {0.0, 0.0};
  double rot_vel = 0.0;
  double lin_acc[2] = {0.0, 0.0};
  double imu_bias = 0.0;
  double stamp = 0.0;
};
// predict / predictSE2 (R/include/ndt_registration/ceres_residuals.h:25-85): constant-acceleration motion model over max(dt, 0.2 s)
RANDT_API void predict(const State& old_state, double raw_dt, State& new_state);      // vector representation (pos, rot)
RANDT_API void predictSE2(const State& old_state, double raw_dt, State& new_state);   // Lie-group representation (pose)

// The parameter fields of the reference this path reads (R/include/ndt_slam/ndt_slam_parameters.h:17-50,56-84), already derived as
// NDTSlam::readParameters leaves them (size in cells after the int /= resolution; n_clusters = int((2 max_range / resolution)^2)).
struct RANDT_API NDTMapParameters {
  double resolution = 1.0;
  int size_x = 50, size_y = 50;
  double max_neighbour_manhattan_distance = 4.0;
  int min_points_per_cell = 3;
  float max_range = 16.0f;      // radar_preprocessor/max_range (cluster grid)
  int n_clusters = 1024;
  randt_grid_params grid() const;
};
struct NDTMatcherParameters {
  int gnc_steps = 3;
  int smoothing_steps = 3;
  double loss_function_convexity = -2.0;
  double loss_function_scale = 1.5;
  double gnc_control_parameter_divisor = 1.3;
  int max_iteration = 200;
  double pose_reject_translation = 2.0, pose_reject_rotation = 2.0;
  int n_results_kd_lookup = 4;
  double ndt_weight = 50000.0;
  bool use_analytic_expressions_for_optimization = false;
  bool use_intensity_as_dimension = true;
  bool optimize_on_manifold = true;
  bool lookup_mahalanobis = true;
  bool use_constant_velocity_model = true;
  double csm_window_linear = 4.5, csm_window_angular = 0.45, csm_linear_step = 0.4, csm_cost_threshold = 0.82, csm_max_px_accurate_range = 4.0;
  int csm_n_iter = 2;
  // the host factors of the joint window problem (ndt_slam_parameters.h:53-58,75): Eigen::Matrix<double,8,8> motion_sqrtI, entry (i, j)
  // at [i * 8 + j] (NDTSlam::readParameters maps the yaml list column-major, ndt_slam.cpp:556 — the shipped matrices are diagonal)
  double motion_sqrtI[64] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0,
                             0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1};
  double covariance_scaling_factor = 25.0;
  double weight_imu = 64.0, weight_imu_bias = 750000.1;
  bool use_imu = false;
};

// what the last estimateTransformCeres call did (the reference keeps only total_optimization_time_; these are for tests and tracing)
struct WindowSummary {
  int status = 0;            // 0 solved, 1 nothing to solve (fewer than two states or no NDT residual block: the reference would read an empty vector)
  int rejected = 0;          // the rejection gate of ndt_matcher.cpp:408-422 fired
  int gnc_solves = 0, total_iterations = 0, n_free_states = 0, n_tangent = 0, evaluations = 0;
  double final_cost = 0, mu_first = 0, max_residual = 0;
  double setup_us = 0, solve_us = 0;   // wall time of building the problem (associations, record table) and of the GNC / LM loop
};

class RANDT_API Context {
 public:
  explicit Context(int device = 0, void* cuda_stream = nullptr);
  ~Context();
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  randt_ctx* get() const { return ctx_; }
  void check(int rc) const;   // throws Error with randt_last_error()
  uint64_t launchCount() const { return randt_ctx_launch_count(ctx_); }

 private:
  randt_ctx* ctx_ = nullptr;
};

// A batch of B independent NDT maps resident on the device (B = 1 is the reference's Map).
class RANDT_API Map {
 public:
  Map(Context& ctx, const NDTMapParameters& p, uint32_t n_maps = 1);   // Map::initialize: empty maps
  ~Map();
  Map(Map&&) noexcept;
  Map(const Map&) = delete;
  // RadarPreprocessor clustering + HierarchicalMap::addClusters (R/src/local_fuser/local_fuser.cpp:102-105): replaces the content
  // with the cells of the given scans.  pts4 = (x, y, unused, intensity) per point; scan b owns [scan_off[b], scan_off[b+1]).
  void addClusters(const float* pts4, const uint32_t* scan_off, uint32_t n_scans);
  void addClusters(const std::vector<float>& pts4) { const uint32_t off[2] = {0, (uint32_t)(pts4.size() / 4)}; addClusters(pts4.data(), off, 1); }
  void transformMap(const SE2d* trans);            // one transform per map (Map::transformMap, float32 like the reference)
  void transformMap(const SE2d& t) { transformMap(&t); }
  void mergeMapCell(const Map& moving);             // Map::mergeMapCell for every map of the batch
  // Map::calculateCSDivergence (ndt_map.cpp:42-99) against the (already transformed) moving map, one value per map of the batch
  std::vector<double> calculateCSDivergence(const Map& m_map) const;
  size_t get_n_cells() const;                       // total over the batch
  uint32_t n_maps() const;
  // Map::getCellMeanAndCovariance (ndt_map.h:112-119): mean[3], row-major cov[9]; cached host copy, refreshed after mutation
  bool getCellMeanAndCovariance(size_t idx, float* mean3, float* cov9) const;
  const std::vector<uint32_t>& cellOffsets() const;  // [B+1]
  // The wire format of an NDT cell (ndt_msgs/Mean + ndt_msgs/Covariance, ros/ndt_msgs/msg/{Mean,Covariance}.msg) as
  // NDTSlam::createVisualizationMsg fills it (R/src/ndt_slam/ndt_slam.cpp:370-393): per cell float64 (x, y, i) and
  // (xx, xy, xi, yy, yi, ii) — the upper triangle of the float32 covariance.  mean_intensity (a per-cluster maximum kept only for
  // rviz) is not produced by this path.
  void exportNormalDistributions(std::vector<double>& mean_xyi, std::vector<double>& cov_xx_xy_xi_yy_yi_ii) const;
  randt_map* handle() const { return map_; }
  Context& context() const { return *ctx_; }
  const NDTMapParameters& parameters() const { return p_; }

 private:
  void sync_host() const;
  Context* ctx_;
  NDTMapParameters p_;
  randt_map* map_ = nullptr;
  mutable bool host_valid_ = false;
  mutable std::vector<float> h_cells_;
  mutable std::vector<uint32_t> h_off_;
};

// One ceres::CostFunction standing in for all residual blocks addNDTFactor would add for one pose: parameter block {4} =
// [cos, sin, tx, ty] (manifold mode) or {2, 1} = pos, rot (vector mode).  Residuals: the P loss-corrected pair residuals, then one
// extra entry sqrt(2 (sum rho/2 - sum rho' r^2/2)) with a zero Jacobian, so that 1/2 |residuals|^2 equals the robustified cost ceres
// would report for the P blocks, and J^T J / J^T r equal what ceres accumulates with its per-block Corrector (exact for rho'' <= 0,
// i.e. every Barron alpha < 2 and Welsch).  With no loss set the residuals and Jacobians are the raw ones and the extra entry is 0.
class RANDT_API NdtCostFunction : public ceres::CostFunction {
 public:
  NdtCostFunction(Context& ctx, randt_problem* problem /*takes ownership*/, int variant);
  ~NdtCostFunction() override;
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
  void setLoss(const randt_loss* loss);            // LossFunctionWrapper::Reset of the GNC loop (ndt_matcher.cpp:392-393)
  uint32_t numPairs() const { return n_pairs_; }
  double maxRawResidual(const double* pose) const;  // Problem::Evaluate(apply_loss_function = false) + max (ndt_matcher.cpp:382-387)
  randt_problem* problem() const { return problem_; }

 private:
  Context* ctx_;
  randt_problem* problem_;
  int variant_;
  uint32_t n_pairs_ = 0;
  bool has_loss_ = false;
  randt_loss loss_{};
  mutable std::vector<double> r_, J_;
};

class RANDT_API Matcher {
 public:
  explicit Matcher(Context& ctx) : ctx_(&ctx) {}
  ~Matcher();
  Matcher(const Matcher&) = delete;
  Matcher& operator=(const Matcher&) = delete;
  void initialize(const NDTMatcherParameters& parameters) { parameters_ = parameters; }
  // Matcher::resetMatcher (ndt_matcher.cpp:18-20): forget the relative IMU constraints collected by predictTransform
  void resetMatcher() { imu_constraints_.clear(); }
  // Matcher::predictTransform (ndt_matcher.cpp:22-59): append the motion-model prediction of the last state at `stamp` to the trajectory
  // (vector model when the analytic functors or the vector parametrisation are configured, SE(2) model otherwise; the predicted state
  // carries zero acceleration) and remember the IMU yaw guess of the step.  Both representations of the new state agree.
  void predictTransform(const double& initial_angle_guess, const double& stamp, std::vector<State>& trajectory);
  const std::vector<double>& imuConstraints() const { return imu_constraints_; }
  // association half of Matcher::addNDTFactor for a batch: problem b pairs moving map b with fixed map b at initial_guess[b]
  randt_problem* associate(const SE2d* initial_guess, const Map& fixed_ndt, const Map& moving_ndt, bool use_intensity_as_dimension,
                           int n_neighbours) const;
  // Matcher::addNDTFactor as one batched cost function (single map pair)
  std::unique_ptr<NdtCostFunction> addNDTFactor(const SE2d& initial_guess, const Map& fixed_ndt, const Map& moving_ndt,
                                                bool use_intensity_as_dimension, int n_neighbours) const;
  // Matcher::estimateTransformCeres (ndt_matcher.cpp:322-424), same arguments: the joint problem over the last smoothing_steps states of
  // `trajectory` (the state before them is constant): per free state the NDT residual blocks of its scan (moving_ndts.end()[-i]) against
  // every fixed map, associated at the state's pose; the motion-model factor between consecutive states (MotionModelFactorSE2 /
  // MotionModelFactor, ceres_residuals.h:554-679); with use_imu the relative yaw factor (RotationalResidualSE2 / RotationalResidual,
  // :307-370) and the IMU bias blocks; the acceleration blocks are constant under use_constant_velocity_model; GNC loop of :382-397 with
  // ScaledLoss(Barron(scale, convexity, mu), ndt_weight / (n_cells k)), n_cells = the window's moving cells; ceres' trust-region
  // Levenberg-Marquardt on the joint tangent space (Sophus::Manifold<SE2> for the poses when optimize_on_manifold), DENSE_QR replaced by
  // the damped normal equations; afterwards the newest state's two representations are synchronised, the rejection gate of :408-422 is
  // applied and `trans` returns the newest pose.  Every LM evaluation is ONE K3 launch over all window states (the normal equations of
  // the NDT blocks per state) plus the 8 (+2) host residuals per factor.  The IMU constraint of the factor ending at trajectory.end()[-i]
  // is imu_constraints_.end()[-i-1] as in the reference (:352), 0 where that reads before the first element.
  void estimateTransformCeres(SE2d& trans, std::vector<State>& trajectory, const double& initial_angle_guess, const double& stamp,
                              const std::deque<Map>& fixed_ndts, const std::deque<Map>& moving_ndts);
  // the same over maps the caller keeps elsewhere (randt::Map is move-only; the reference copies its maps into the deques)
  void estimateTransformCeres(SE2d& trans, std::vector<State>& trajectory, const double& initial_angle_guess, const double& stamp,
                              const std::vector<const Map*>& fixed_ndts, const std::vector<const Map*>& moving_ndts);
  const WindowSummary& lastWindowSummary() const { return window_summary_; }
  // ceres' convergence tolerances for the window solve (<= 0: Solver::Options defaults 1e-6 / 1e-8 / 1e-10)
  void setWindowTolerances(double function_tolerance, double parameter_tolerance, double gradient_tolerance) {
    window_tol_[0] = function_tolerance; window_tol_[1] = parameter_tolerance; window_tol_[2] = gradient_tolerance;
  }
  // the same with the IMU constraints given explicitly (imu[j] belongs to the factor ending at the j-th free state, oldest first)
  void solveWindow(SE2d& trans, std::vector<State>& trajectory, const std::vector<const Map*>& fixed_ndts,
                   const std::vector<const Map*>& moving_window, const std::vector<double>& imu);
  // The NDT part of Matcher::estimateTransformCeres (ndt_matcher.cpp:321-424) for the newest state of the window: residual blocks
  // of the moving scan against EVERY fixed map (the current submap, plus the previous one while they overlap: local_fuser.cpp:129-138)
  // in the reference's block order, loss ScaledLoss(Barron(loss_function_scale, convexity, mu), ndt_weight / (n_cells k)) (:392), the GNC
  // schedule of :382-397 with gnc_steps, the manifold when optimize_on_manifold (and not the analytic functors), and the rejection gate
  // of :408-422.  The motion-model and IMU factors of the reference's joint problem are host-side factors outside this path: a caller
  // that needs them keeps ceres and plugs NdtCostFunction in instead.  trans: prior in, estimate out (untouched when rejected).
  // Batched: problem b = moving map b against map b of every entry of fixed_ndts.  Returns accepted[b].
  std::vector<char> estimateTransformsNDT(std::vector<SE2d>& trans, const std::vector<const Map*>& fixed_ndts, const Map& moving_ndts) const;
  bool estimateTransformNDT(SE2d& trans, const std::vector<const Map*>& fixed_ndts, const Map& moving_ndt) const;
  // Matcher::estimateLoopConstraint (ndt_matcher.cpp:426-493), B = 1
  double estimateLoopConstraint(SE2d& trans, const Map& old_ndt, Map& new_ndt, int max_gnc_steps, bool use_intensity_as_dimension,
                                double scale) const;
  // the same for B independent (submap, scan) pairs in one device-resident solve (loop-closure candidate sweep); returns the scores
  std::vector<double> estimateLoopConstraints(std::vector<SE2d>& trans, const Map& old_ndts, Map& new_ndts, int max_gnc_steps,
                                              bool use_intensity_as_dimension, double scale) const;
  // Matcher::estimateTransformGlobalBNB (ndt_matcher.cpp:495-608): same coarse-to-fine tree, every queued level evaluated in one sweep
  double estimateTransformGlobalBNB(SE2d& trans, const Map& fixed_ndt, Map& moving_ndt, bool use_intensity_as_dimension, double scale,
                                    double search_window_size_linear, double search_window_size_angular) const;
  const NDTMatcherParameters& parameters() const { return parameters_; }
  int variant(bool use_intensity_as_dimension) const;

 private:
  Context* ctx_;
  NDTMatcherParameters parameters_;
  std::vector<double> imu_constraints_;
  State X_next_;
  WindowSummary window_summary_;
  double window_tol_[3] = {0.0, 0.0, 0.0};
  double* window_staging_ = nullptr;   // pinned poses + records of one window evaluation (kept across calls)
  size_t window_staging_cap_ = 0;
};

}  // namespace randt

// C entry points over the classes above, used by the Python tests (ctypes) to drive the C++ layer; not part of the drop-in ABI.
extern "C" {
RANDT_API int randt_hostapi_loop_constraints(int device, const randt_grid_params* gp, const float* fixed_pts4, const uint32_t* fixed_off,
                                             const float* moving_pts4, const uint32_t* moving_off, uint32_t n_problems, int k,
                                             double loss_function_scale, double convexity, double divisor, int max_gnc_steps, double loop_scale,
                                             int optimize_on_manifold, double* poses_io /*[n][4]*/, double* scores /*[n]*/);
RANDT_API int randt_hostapi_cost_function(int device, const randt_grid_params* gp, const float* fixed_pts4, uint32_t n_fixed,
                                          const float* moving_pts4, uint32_t n_moving, int k, const double* guess4, const randt_loss* loss,
                                          const double* pose4, double* residuals /*[cap]*/, double* jacobian /*[cap][4]*/, uint32_t cap,
                                          uint32_t* num_residuals, double* max_raw);
RANDT_API int randt_hostapi_bnb(int device, const randt_grid_params* gp, const float* fixed_pts4, uint32_t n_fixed, const float* moving_pts4,
                                uint32_t n_moving, double convexity, double scale, double window_linear, double window_angular,
                                double linear_step, double max_px_range, double cost_threshold, int n_iter, double* pose_io4, double* min_cost,
                                uint32_t* n_evaluated);
RANDT_API int randt_hostapi_odometry(int device, const randt_grid_params* gp, const float* const* fixed_pts4, const uint32_t* n_fixed_pts,
                                     const double* fixed_pose4 /*[n_fixed][4]: each fixed scan is voxelised, then moved by its pose*/,
                                     uint32_t n_fixed, const float* moving_pts4, uint32_t n_moving, int k, double loss_function_scale,
                                     double convexity, double divisor, int gnc_steps, double ndt_weight, int optimize_on_manifold,
                                     double reject_translation, double reject_rotation, double* pose_io4, int* accepted);
RANDT_API int randt_hostapi_export(int device, const randt_grid_params* gp, const float* pts4, uint32_t n_pts, double* mean3 /*[cap][3]*/,
                                   double* cov6 /*[cap][6]*/, uint32_t cap, uint32_t* n_cells);
/* `steps` pipelined evaluations through randt_eval_fused_async from a C++ caller (what bench.py times as e2e): step i uses the i-th of
 * `depth` caller-owned pinned buffer sets (poses [S][np] in, records out); a set is reused once its previous call has delivered. */
RANDT_API int randt_hostapi_eval_async_loop(randt_ctx* ctx, const randt_problem* problem, int variant, const randt_loss* loss, double* const* poses_ring,
                                            double* const* out_ring, uint32_t depth, uint32_t steps, int packed);
/* The K3 work schedule (randt_slam_b200/csrc/schedule.hpp: tiles, longest-processing-time assignment to `max_warps` warps, chunk plans A and
 * B) for the given per-segment duo offsets, without a device: what tests/test_schedule_cpu.py checks.  counts[5] = tiles, warps, chunks
 * of plan A, chunks of plan B, records; a NULL output array is skipped; arrays must hold cap_tiles (+1) / cap_chunks / max_warps+1 /
 * n_segments+1 entries (RANDT_E_CAPACITY otherwise, counts still filled).  tiles4 / plan_*4: 4 x uint32 per entry (seg, begin, end, part /
 * duo_begin, meta, seg, part). */
RANDT_API int randt_hostapi_build_schedule(const uint32_t* duo_off, uint32_t n_segments, uint32_t max_warps, uint32_t* counts, uint32_t* tiles4,
                                           uint32_t cap_tiles, uint32_t* plan_a4, uint32_t* plan_b4, uint32_t cap_chunks, uint32_t* woff_a,
                                           uint32_t* woff_b, uint32_t* tile_rec_begin, uint32_t* tile_duo_begin, uint32_t* first);
/* predict (se2_model == 0) / predictSE2 (se2_model != 0) on one state: state12 = [cos, sin, tx, ty, pos_x, pos_y, rot, vx, vy, omega, ax, ay] in and out */
RANDT_API void randt_hostapi_predict(int se2_model, const double* state12, double raw_dt, double* out12);
/* The host factors of the window problem without a device (tests): the motion-model (+ IMU) factors over W + 1 states
 * (states [(W + 1)][14]: cos, sin, tx, ty, pos_x, pos_y, rot, vx, vy, omega, ax, ay, imu_bias, stamp; the first one constant), evaluated
 * at the states as given: cost, tangent gradient g [nt] and J^T J H [nt * nt] in the solver's parameter order (first state: lin_vel 2 |
 * rot_vel 1 | lin_acc 2 unless constant velocity — only its pose and bias are constant, ndt_matcher.cpp:304-320; every later state: pose 3 |
 * lin_vel 2 | rot_vel 1 | lin_acc 2 unless constant velocity | imu_bias 1 with use_imu).  params as in randt_hostapi_window_solve.
 * Returns nt (or a negative error). */
RANDT_API int randt_hostapi_window_factors(const double* states, uint32_t W, const double* imu, const double* params80, double* cost, double* g,
                                           double* H);
/* The trust-region loop of the window solve on the host factors alone (no NDT term, no device): minimises the motion-model (+ IMU) cost
 * over the free parameters of the window states with window::minimize (ceres' Levenberg-Marquardt restated), states in / out.
 * tolerances3 / max_iterations <= 0: ceres' defaults / 200.  summary4: initial cost, final cost, iterations, termination (0 convergence,
 * 1 iteration limit, 2 failure).  Returns 0 or a negative error. */
RANDT_API int randt_hostapi_window_minimize_factors(double* states, uint32_t W, const double* imu, const double* params80, const double* tolerances3,
                                                    int max_iterations, double* summary4);
/* Matcher::estimateTransformCeres over scans given as points: fixed scan f is voxelised and moved by fixed_pose4[f] (one fixed map each);
 * window scan w (oldest first, W of them) is voxelised as it is.  states [(W + 1)][14] in / out, trans4 in / out.  params80 [16 + 64]: k,
 * gnc_steps, max_iteration, loss_scale, alpha, divisor, ndt_weight, manifold, constant_velocity, use_imu, weight_imu, weight_imu_bias,
 * reject_translation, reject_rotation, use_intensity, (reserved), then covariance_scaling_factor * motion_sqrtI, entry (i, j) at [i * 8 + j].
 * tolerances3 (may be NULL): function, parameter, gradient tolerance.  out10: status, rejected, gnc_solves, total_iterations, final_cost,
 * mu_first, max_residual, n_tangent, evaluations, n_cells of the window. */
RANDT_API int randt_hostapi_window_solve(int device, const randt_grid_params* gp, const float* const* fixed_pts4, const uint32_t* n_fixed_pts,
                                         const double* fixed_pose4, uint32_t n_fixed, const float* const* window_pts4, const uint32_t* n_window_pts,
                                         uint32_t W, double* states, const double* imu, const double* params80, const double* tolerances3,
                                         double* trans4, double* out10);
/* LocalFuser::processScan in miniature (R/src/local_fuser/local_fuser.cpp:99-300, one submap, no submap roll-over): per scan voxelise ->
 * Matcher::predictTransform -> Matcher::estimateTransformCeres against the submap over the window of the last smoothing_steps scans -> both
 * pose representations of the window states -> keyframes (every insertion_step-th state) enter the submap insertion_delay = smoothing_steps + 1
 * scans later at their smoothed pose (transformMap + mergeMapCell); the first scan initialises the trajectory (zero velocities) and the
 * submap.  pts4 / scan_off: the scans back to back; stamps [n]; yaw [n] (relative IMU yaw per scan, may be NULL); params80 as in
 * randt_hostapi_window_solve.  poses_out [n][4]: the estimate returned for each scan when it arrived; states_out [n][14]: the trajectory
 * at the end (smoothed); stats_out [n][4]: total LM iterations, device evaluations, rejected, wall microseconds of the scan;
 * totals [6]: wall seconds of the whole loop, submap cells at the end, kernel launches, keyframes merged, seconds spent building the window
 * problems, seconds spent in their GNC / LM loops. */
RANDT_API int randt_hostapi_window_replay(int device, const randt_grid_params* gp, const float* pts4, const uint32_t* scan_off, uint32_t n_scans,
                                          const double* stamps, const double* yaw, const double* params80, int smoothing_steps, int insertion_step,
                                          double* poses_out, double* states_out, double* stats_out, double* totals);
RANDT_API const char* randt_hostapi_last_error(void);
}
