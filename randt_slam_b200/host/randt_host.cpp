// Host-side C++ mirror of the reference's Map / Matcher surface for the NDT hot path (see include/randt_host.hpp).
// Pure orchestration over the C-ABI: every per-point / per-cell / per-pair operation runs in the CUDA kernels.
#include "../../include/randt_host.hpp"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>
#include <deque>
#include <set>
#include <array>
#include <tuple>

#include "../csrc/schedule.hpp"   // the K3 schedule builder (plain C++), exposed for the CPU tests
#include "window_solver.hpp"

namespace randt {

// ---------------------------------------------------------------------------------------------------------------------
// SE2d
// ---------------------------------------------------------------------------------------------------------------------
SE2d::SE2d(double theta, double tx, double ty) { v[0] = std::cos(theta); v[1] = std::sin(theta); v[2] = tx; v[3] = ty; }
double SE2d::angle() const { return std::atan2(v[1], v[0]); }
SE2d SE2d::operator*(const SE2d& o) const {
  SE2d r;
  double re = v[0] * o.v[0] - v[1] * o.v[1], im = v[0] * o.v[1] + v[1] * o.v[0];
  const double n2 = re * re + im * im;
  if (n2 != 1.0) { const double sc = 2.0 / (1.0 + n2); re *= sc; im *= sc; }   // Sophus SO2 product: first-order renormalisation
  r.v[0] = re; r.v[1] = im;
  r.v[2] = v[2] + (v[0] * o.v[2] - v[1] * o.v[3]);
  r.v[3] = v[3] + (v[1] * o.v[2] + v[0] * o.v[3]);
  return r;
}
SE2d SE2d::exp(double ux, double uy, double theta) {
  // Sophus::SE2::exp: SO2::exp(theta), translation = V(theta) * (ux, uy) with the small-angle series below Constants<double>::epsilon()
  const double s = std::sin(theta), c = std::cos(theta);
  double sbt, omcbt;   // sin(theta) / theta, (1 - cos(theta)) / theta
  if (std::fabs(theta) < 1e-10) {
    const double t2 = theta * theta;
    sbt = 1.0 - (1.0 / 6.0) * t2;
    omcbt = 0.5 * theta - (1.0 / 24.0) * theta * t2;
  } else { sbt = s / theta; omcbt = (1.0 - c) / theta; }
  SE2d r;
  r.v[0] = c; r.v[1] = s;
  r.v[2] = sbt * ux - omcbt * uy;
  r.v[3] = omcbt * ux + sbt * uy;
  return r;
}

namespace {
// NormalizeAngle (R/include/ndt_registration/state_manifold.h:17-23)
double normalize_angle(double a) { const double two_pi = 2.0 * M_PI; return a - two_pi * std::floor((a + M_PI) / two_pi); }
}  // namespace

void predict(const State& o, double raw_dt, State& n) {
  const double dt = std::max(raw_dt, 0.2);          // identical stamps happen (driver hiccups): ceres_residuals.h:38
  const double rot = normalize_angle(o.rot + 0.5 * dt * o.rot_vel);
  const double sy = std::sin(rot), cy = std::cos(rot);
  const double dx = o.lin_vel[0] * dt + 0.5 * o.lin_acc[0] * dt * dt, dy = o.lin_vel[1] * dt + 0.5 * o.lin_acc[1] * dt * dt;
  State r = o;
  r.rot = normalize_angle(o.rot + dt * o.rot_vel);
  r.pos[0] = o.pos[0] + (cy * dx - sy * dy);
  r.pos[1] = o.pos[1] + (sy * dx + cy * dy);
  r.lin_vel[0] = o.lin_vel[0] + dt * o.lin_acc[0];
  r.lin_vel[1] = o.lin_vel[1] + dt * o.lin_acc[1];
  n.pos[0] = r.pos[0]; n.pos[1] = r.pos[1]; n.rot = r.rot;
  n.lin_vel[0] = r.lin_vel[0]; n.lin_vel[1] = r.lin_vel[1]; n.rot_vel = o.rot_vel; n.lin_acc[0] = o.lin_acc[0]; n.lin_acc[1] = o.lin_acc[1];
}

void predictSE2(const State& o, double raw_dt, State& n) {
  const double dt = std::max(raw_dt, 0.2);
  // screw = (v dt + dt a / 2, omega dt): the reference's SE(2) model integrates the acceleration with dt / 2, not dt^2 / 2 (ceres_residuals.h:77-79)
  const SE2d step = SE2d::exp(o.lin_vel[0] * dt + 0.5 * dt * o.lin_acc[0], o.lin_vel[1] * dt + 0.5 * dt * o.lin_acc[1], o.rot_vel * dt);
  n.pose = o.pose * step;
  n.lin_vel[0] = o.lin_vel[0] + dt * o.lin_acc[0];
  n.lin_vel[1] = o.lin_vel[1] + dt * o.lin_acc[1];
  n.rot_vel = o.rot_vel; n.lin_acc[0] = o.lin_acc[0]; n.lin_acc[1] = o.lin_acc[1];
}

void Matcher::predictTransform(const double& initial_angle_guess, const double& stamp, std::vector<State>& trajectory) {
  if (trajectory.empty()) return;
  State last_state = trajectory.back();
  X_next_.stamp = stamp;
  last_state.lin_acc[0] = 0.0; last_state.lin_acc[1] = 0.0;
  const double dt = stamp - trajectory.back().stamp;
  if (parameters_.use_analytic_expressions_for_optimization || !parameters_.optimize_on_manifold) {
    predict(last_state, dt, X_next_);
    X_next_.pose = SE2d(X_next_.rot, X_next_.pos[0], X_next_.pos[1]);     // both representations hold the same information
  } else {
    predictSE2(last_state, dt, X_next_);
    X_next_.pos[0] = X_next_.pose.v[2]; X_next_.pos[1] = X_next_.pose.v[3];
    X_next_.rot = X_next_.pose.angle();
  }
  trajectory.push_back(X_next_);
  imu_constraints_.push_back(initial_angle_guess);
}

void SE2d::matrix3f(float out[9]) const {
  const float m[9] = {(float)v[0], (float)v[1], 0.f, (float)-v[1], (float)v[0], 0.f, (float)v[2], (float)v[3], 1.f};
  std::memcpy(out, m, sizeof(m));
}

randt_grid_params NDTMapParameters::grid() const {
  randt_grid_params g;
  g.max_range = max_range; g.n_clusters = n_clusters; g.min_points = min_points_per_cell; g.size_x = size_x; g.size_y = size_y;
  g.resolution = resolution; g.max_linf = max_neighbour_manhattan_distance;
  return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------------------
Context::Context(int device, void* cuda_stream) {
  const int rc = randt_ctx_create(device, cuda_stream, &ctx_);
  if (rc != RANDT_OK) throw Error(rc, "randt_ctx_create failed: no usable CUDA device " + std::to_string(device));
}
Context::~Context() { randt_ctx_destroy(ctx_); }
void Context::check(int rc) const {
  if (rc != RANDT_OK) throw Error(rc, std::string("randt: ") + randt_last_error(ctx_));
}

// ---------------------------------------------------------------------------------------------------------------------
// Map
// ---------------------------------------------------------------------------------------------------------------------
Map::Map(Context& ctx, const NDTMapParameters& p, uint32_t n_maps) : ctx_(&ctx), p_(p) {
  const std::vector<uint32_t> off(n_maps + 1, 0u);
  const randt_grid_params g = p_.grid();
  ctx_->check(randt_map_upload(ctx_->get(), nullptr, nullptr, off.data(), n_maps, nullptr, &g, &map_));
}
Map::~Map() { randt_map_destroy(map_); }
Map::Map(Map&& o) noexcept : ctx_(o.ctx_), p_(o.p_), map_(o.map_) { o.map_ = nullptr; }

void Map::addClusters(const float* pts4, const uint32_t* scan_off, uint32_t n_scans) {
  const randt_grid_params g = p_.grid();
  randt_map* fresh = nullptr;
  ctx_->check(randt_voxelize(ctx_->get(), pts4, scan_off, n_scans, &g, 0, &fresh));
  randt_map_destroy(map_);
  map_ = fresh;
  host_valid_ = false;
}
void Map::transformMap(const SE2d* trans) {
  // Eigen::Affine2f(trans.cast<float>().matrix()) — the cast (with Sophus' re-normalisation) happens on the device
  const uint32_t B = n_maps();
  std::vector<double> t(4 * (size_t)B);
  for (uint32_t b = 0; b < B; ++b) for (int i = 0; i < 4; ++i) t[4 * b + i] = trans[b].v[i];
  ctx_->check(randt_map_transform_se2d(ctx_->get(), map_, t.data()));
  host_valid_ = false;
}
void Map::mergeMapCell(const Map& moving) {
  ctx_->check(randt_map_merge(ctx_->get(), map_, moving.map_));
  host_valid_ = false;
}
std::vector<double> Map::calculateCSDivergence(const Map& m_map) const {
  std::vector<double> out(n_maps(), 0.0);
  ctx_->check(randt_cs_divergence(ctx_->get(), map_, m_map.map_, out.data()));
  return out;
}
uint32_t Map::n_maps() const { uint32_t b = 0; randt_map_info(map_, &b, nullptr, nullptr); return b; }
size_t Map::get_n_cells() const { uint32_t n = 0; randt_map_info(map_, nullptr, &n, nullptr); return n; }
void Map::sync_host() const {
  if (host_valid_) return;
  uint32_t b = 0, n = 0;
  randt_map_info(map_, &b, &n, nullptr);
  h_cells_.assign((size_t)n * 12, 0.f); h_off_.assign(b + 1, 0u);
  ctx_->check(randt_map_download(ctx_->get(), map_, h_cells_.data(), nullptr, nullptr, h_off_.data(), nullptr));
  host_valid_ = true;
}
bool Map::getCellMeanAndCovariance(size_t idx, float* mean3, float* cov9) const {
  sync_host();
  if (idx >= h_cells_.size() / 12) return false;
  std::memcpy(mean3, &h_cells_[idx * 12], 3 * sizeof(float));
  std::memcpy(cov9, &h_cells_[idx * 12 + 3], 9 * sizeof(float));
  return true;
}
const std::vector<uint32_t>& Map::cellOffsets() const { sync_host(); return h_off_; }
void Map::exportNormalDistributions(std::vector<double>& mean, std::vector<double>& cov) const {
  sync_host();
  const size_t n = h_cells_.size() / 12;
  mean.resize(3 * n); cov.resize(6 * n);
  for (size_t i = 0; i < n; ++i) {
    const float* c = &h_cells_[12 * i];
    mean[3 * i] = c[0]; mean[3 * i + 1] = c[1]; mean[3 * i + 2] = c[2];
    cov[6 * i] = c[3]; cov[6 * i + 1] = c[4]; cov[6 * i + 2] = c[5];            // cov(0,0), cov(0,1), cov(0,2)
    cov[6 * i + 3] = c[7]; cov[6 * i + 4] = c[8]; cov[6 * i + 5] = c[11];       // cov(1,1), cov(1,2), cov(2,2)
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// NdtCostFunction
// ---------------------------------------------------------------------------------------------------------------------
namespace {
// BarronLoss / WelschLoss::Evaluate wrapped in ScaledLoss (R/src/ndt_registration/ceres_loss_functions.cpp:10-39) — host copy used
// only for the per-residual correction of the ceres adapter; the solver path evaluates the loss inside K3.
void loss_rho(const randt_loss& l, double s, double rho[3]) {
  if (l.kind == RANDT_LOSS_WELSCH) {
    const double b = l.mu * l.scale * l.scale, c = -1.0 / b, ex = std::exp(s * c);
    rho[0] = b * (1 - ex); rho[1] = ex; rho[2] = c * ex;
  } else if (l.kind == RANDT_LOSS_BARRON) {
    const double b = l.mu * l.scale * l.scale, c = 1 / b, factor = std::abs(l.alpha - 2.0), e = 0.5 * l.alpha;
    if (l.alpha >= 2.0) { rho[0] = s; rho[1] = 1; rho[2] = 0; }
    else if (std::abs(l.alpha) <= 0.05) {
      const double sum = 1.0 + s * c, inv = 1.0 / sum;
      rho[0] = b * std::log(sum); rho[1] = std::max(2.2250738585072014e-308, inv); rho[2] = -c * (inv * inv);
    } else {
      const double pre = b * factor / l.alpha, ts = 2 * c / factor, u = s * ts + 1.0;
      rho[0] = pre * (std::pow(u, e) - 1.); rho[1] = pre * e * std::pow(u, e - 1.) * ts;
      rho[2] = pre * e * (e - 1) * std::pow(u, e - 2.) * ts * ts;
    }
  } else { rho[0] = s; rho[1] = 1; rho[2] = 0; }
  rho[0] *= l.weight; rho[1] *= l.weight; rho[2] *= l.weight;
}
}  // namespace

NdtCostFunction::NdtCostFunction(Context& ctx, randt_problem* problem, int variant) : ctx_(&ctx), problem_(problem), variant_(variant) {
  uint32_t S = 0;
  randt_problem_info(problem_, &S, &n_pairs_, nullptr, nullptr);
  if (S != 1) { randt_problem_destroy(problem_); throw Error(RANDT_E_INVALID, "NdtCostFunction needs a single-segment problem"); }
  if (variant_ <= RANDT_VAR_SE2_XY) mutable_parameter_block_sizes()->push_back(4);
  else { mutable_parameter_block_sizes()->push_back(2); mutable_parameter_block_sizes()->push_back(1); }   // pos[2], rot[1] (ndt_matcher.cpp:245)
  set_num_residuals((int)n_pairs_ + 1);
  r_.resize(n_pairs_); J_.resize((size_t)n_pairs_ * 4);
}
NdtCostFunction::~NdtCostFunction() { randt_problem_destroy(problem_); }
void NdtCostFunction::setLoss(const randt_loss* loss) { has_loss_ = loss != nullptr; if (loss) loss_ = *loss; }

double NdtCostFunction::maxRawResidual(const double* pose) const {
  ctx_->check(randt_eval_emit(ctx_->get(), problem_, variant_, pose, r_.data(), nullptr));
  double m = 0;
  for (double r : r_) if (r > m) m = r;
  return m;
}

bool NdtCostFunction::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  const bool se2 = variant_ <= RANDT_VAR_SE2_XY;
  double params[4];
  if (se2) std::memcpy(params, parameters[0], 4 * sizeof(double));
  else { params[0] = parameters[0][0]; params[1] = parameters[0][1]; params[2] = parameters[1][0]; }
  const int np = se2 ? 4 : 3;
  const bool want_j = jacobians && (jacobians[0] || (!se2 && jacobians[1]));
  if (randt_eval_emit(ctx_->get(), problem_, variant_, params, r_.data(), want_j ? J_.data() : nullptr) != RANDT_OK) return false;
  double deficit = 0.0;   // sum rho - sum rho' r^2  (>= 0 for a concave rho with rho(0) = 0)
  for (uint32_t p = 0; p < n_pairs_; ++p) {
    double r = r_[p];
    if (!std::isfinite(r)) return false;   // degenerate pair: ceres rejects the step (Evaluate returns false)
    double sc = 1.0;
    if (has_loss_) {
      double rho[3];
      const double sq = r * r;
      loss_rho(loss_, sq, rho);
      sc = std::sqrt(rho[1]);               // Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho')
      deficit += rho[0] - rho[1] * sq;
    }
    residuals[p] = sc * r;
    if (want_j) {
      const double* Jp = &J_[(size_t)p * np];
      if (se2) { if (jacobians[0]) for (int i = 0; i < 4; ++i) jacobians[0][(size_t)p * 4 + i] = sc * Jp[i]; }
      else {
        if (jacobians[0]) { jacobians[0][(size_t)p * 2] = sc * Jp[0]; jacobians[0][(size_t)p * 2 + 1] = sc * Jp[1]; }
        if (jacobians[1]) jacobians[1][p] = sc * Jp[2];
      }
    }
  }
  residuals[n_pairs_] = std::sqrt(std::max(0.0, deficit));
  if (want_j) {
    if (se2) { if (jacobians[0]) for (int i = 0; i < 4; ++i) jacobians[0][(size_t)n_pairs_ * 4 + i] = 0.0; }
    else {
      if (jacobians[0]) { jacobians[0][(size_t)n_pairs_ * 2] = 0.0; jacobians[0][(size_t)n_pairs_ * 2 + 1] = 0.0; }
      if (jacobians[1]) jacobians[1][n_pairs_] = 0.0;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Matcher
// ---------------------------------------------------------------------------------------------------------------------
int Matcher::variant(bool use_intensity_as_dimension) const {
  const bool manifold = parameters_.optimize_on_manifold && !parameters_.use_analytic_expressions_for_optimization;
  if (manifold) return use_intensity_as_dimension ? RANDT_VAR_SE2_INTENSITY : RANDT_VAR_SE2_XY;
  return use_intensity_as_dimension ? RANDT_VAR_VEC_INTENSITY : RANDT_VAR_VEC_XY;
}

randt_problem* Matcher::associate(const SE2d* initial_guess, const Map& fixed_ndt, const Map& moving_ndt, bool use_intensity_as_dimension,
                                  int n_neighbours) const {
  const uint32_t B = fixed_ndt.n_maps();
  std::vector<double> pose0(4 * (size_t)B);
  for (uint32_t b = 0; b < B; ++b) std::memcpy(&pose0[4 * b], initial_guess[b].v, 4 * sizeof(double));
  // ndt_matcher.cpp:201: Mahalanobis lookup needs both use_intensity_as_dimension and lookup_mahalanobis, else Euclidean xy
  const int metric = (use_intensity_as_dimension && parameters_.lookup_mahalanobis) ? RANDT_LOOKUP_MAHALANOBIS_INTENSITY : RANDT_LOOKUP_EUCLID_XY;
  randt_problem* prob = nullptr;
  ctx_->check(randt_associate(ctx_->get(), fixed_ndt.handle(), moving_ndt.handle(), pose0.data(), n_neighbours, metric, &prob));
  return prob;
}

std::unique_ptr<NdtCostFunction> Matcher::addNDTFactor(const SE2d& initial_guess, const Map& fixed_ndt, const Map& moving_ndt,
                                                       bool use_intensity_as_dimension, int n_neighbours) const {
  if (fixed_ndt.n_maps() != 1 || moving_ndt.n_maps() != 1) throw Error(RANDT_E_INVALID, "addNDTFactor takes single maps");
  randt_problem* prob = associate(&initial_guess, fixed_ndt, moving_ndt, use_intensity_as_dimension, n_neighbours);
  return std::unique_ptr<NdtCostFunction>(new NdtCostFunction(*ctx_, prob, variant(use_intensity_as_dimension)));
}

std::vector<char> Matcher::estimateTransformsNDT(std::vector<SE2d>& trans, const std::vector<const Map*>& fixed_ndts, const Map& moving_ndts) const {
  const uint32_t B = moving_ndts.n_maps();
  if (fixed_ndts.empty()) throw Error(RANDT_E_INVALID, "estimateTransformsNDT: no fixed map");
  if (trans.size() != B) throw Error(RANDT_E_INVALID, "estimateTransformsNDT: one prior per moving map");
  for (const Map* f : fixed_ndts) if (!f || f->n_maps() != B) throw Error(RANDT_E_INVALID, "estimateTransformsNDT: batch sizes differ");
  const bool use_intensity = parameters_.use_intensity_as_dimension;
  const int k = parameters_.n_results_kd_lookup;
  struct Problems { std::vector<randt_problem*> v; ~Problems() { for (randt_problem* p : v) randt_problem_destroy(p); } } probs;
  for (const Map* f : fixed_ndts) probs.v.push_back(associate(trans.data(), *f, moving_ndts, use_intensity, k));
  randt_problem* prob = probs.v[0];
  if (probs.v.size() > 1) {
    // one residual-block list per state: the blocks of fixed map 0, then those of fixed map 1, ... (estimateTransformCeres appends
    // the ids addNDTFactor returns map after map).  The moving cells are shared, the fixed tables are laid one after the other.
    std::vector<float> cells_m, cells_f;
    std::vector<std::vector<uint32_t>> pm(probs.v.size()), pf(probs.v.size()), so(probs.v.size());
    uint32_t n_m = 0, f_base = 0;
    for (size_t i = 0; i < probs.v.size(); ++i) {
      uint32_t S = 0, P = 0, nm = 0, nf = 0;
      ctx_->check(randt_problem_info(probs.v[i], &S, &P, &nm, &nf));
      pm[i].resize(P); pf[i].resize(P); so[i].resize((size_t)S + 1);
      ctx_->check(randt_problem_download(ctx_->get(), probs.v[i], pm[i].data(), pf[i].data(), so[i].data()));
      std::vector<float> cm((size_t)nm * 12), cf((size_t)nf * 12);
      ctx_->check(randt_problem_download_cells(ctx_->get(), probs.v[i], cm.data(), cf.data()));
      if (i == 0) { cells_m = cm; n_m = nm; }
      cells_f.insert(cells_f.end(), cf.begin(), cf.end());
      for (uint32_t& j : pf[i]) j += f_base;
      f_base += nf;
    }
    std::vector<uint32_t> pair_m, pair_f, seg_off((size_t)B + 1, 0);
    for (uint32_t b = 0; b < B; ++b) {
      for (size_t i = 0; i < probs.v.size(); ++i) {
        pair_m.insert(pair_m.end(), pm[i].begin() + so[i][b], pm[i].begin() + so[i][b + 1]);
        pair_f.insert(pair_f.end(), pf[i].begin() + so[i][b], pf[i].begin() + so[i][b + 1]);
      }
      seg_off[b + 1] = (uint32_t)pair_m.size();
    }
    randt_problem* merged = nullptr;
    ctx_->check(randt_problem_create(ctx_->get(), cells_m.data(), n_m, cells_f.data(), f_base, pair_m.data(), pair_f.data(), (uint32_t)pair_m.size(),
                                     seg_off.data(), B, &merged));
    probs.v.push_back(merged);
    prob = merged;
  }
  const int var = variant(use_intensity);
  const int np = var <= RANDT_VAR_SE2_XY ? 4 : 3;
  std::vector<double> poses((size_t)B * np), result((size_t)B * RANDT_REG_STRIDE);
  for (uint32_t b = 0; b < B; ++b) {
    if (np == 4) std::memcpy(&poses[4 * b], trans[b].v, 4 * sizeof(double));
    else { poses[3 * b] = trans[b].v[2]; poses[3 * b + 1] = trans[b].v[3]; poses[3 * b + 2] = trans[b].angle(); }
  }
  // n_cells counts the moving map's cells (ndt_matcher.cpp:372), once, whatever the number of fixed maps: every registration of the
  // batch gets the weight its own scan would get alone
  const std::vector<uint32_t>& m_off = moving_ndts.cellOffsets();
  std::vector<double> weights(B);
  for (uint32_t b = 0; b < B; ++b) weights[b] = parameters_.ndt_weight / ((double)(m_off[b + 1] - m_off[b]) * (double)k);
  randt_loss loss;
  loss.kind = RANDT_LOSS_BARRON; loss.scale = parameters_.loss_function_scale; loss.alpha = parameters_.loss_function_convexity; loss.mu = 1.0;
  loss.weight = 1.0;
  randt_solver_options opt;
  randt_solver_options_default(&opt);
  opt.max_num_iterations = parameters_.max_iteration;
  opt.use_manifold = (parameters_.optimize_on_manifold && !parameters_.use_analytic_expressions_for_optimization) ? 1 : 0;
  opt.gnc_loss_scale = parameters_.loss_function_scale;
  opt.gnc_divisor = parameters_.gnc_control_parameter_divisor;
  opt.gnc_max_steps = parameters_.gnc_steps;
  ctx_->check(randt_register_batch_weighted(ctx_->get(), prob, var, poses.data(), &loss, weights.data(), &opt, result.data()));
  std::vector<char> accepted(B, 1);
  for (uint32_t b = 0; b < B; ++b) {
    SE2d est;
    if (np == 4) std::memcpy(est.v, &poses[4 * b], 4 * sizeof(double));
    else est = SE2d(poses[3 * b + 2], poses[3 * b], poses[3 * b + 1]);
    // reject an estimate that strays too far from the prior (ndt_matcher.cpp:408-411); no residual block at all is a failure too
    double da = est.angle() - trans[b].angle();
    da = std::atan2(std::sin(da), std::cos(da));
    const bool stray = std::fabs(est.v[2] - trans[b].v[2]) > parameters_.pose_reject_translation ||
                       std::fabs(est.v[3] - trans[b].v[3]) > parameters_.pose_reject_translation || std::fabs(da) > parameters_.pose_reject_rotation;
    if (stray || result[(size_t)b * RANDT_REG_STRIDE + RANDT_REG_STATUS] != 0.0) accepted[b] = 0;
    else trans[b] = est;
  }
  return accepted;
}

bool Matcher::estimateTransformNDT(SE2d& trans, const std::vector<const Map*>& fixed_ndts, const Map& moving_ndt) const {
  std::vector<SE2d> t(1, trans);
  const std::vector<char> ok = estimateTransformsNDT(t, fixed_ndts, moving_ndt);
  trans = t[0];
  return ok[0] != 0;
}

std::vector<double> Matcher::estimateLoopConstraints(std::vector<SE2d>& trans, const Map& old_ndts, Map& new_ndts, int max_gnc_steps,
                                                     bool use_intensity_as_dimension, double scale) const {
  const uint32_t B = old_ndts.n_maps();
  if (trans.size() != B || new_ndts.n_maps() != B) throw Error(RANDT_E_INVALID, "estimateLoopConstraints: batch sizes differ");
  randt_problem* prob = associate(trans.data(), old_ndts, new_ndts, use_intensity_as_dimension, parameters_.n_results_kd_lookup);
  struct Guard { randt_problem* p; ~Guard() { randt_problem_destroy(p); } } guard{prob};
  const int var = variant(use_intensity_as_dimension);
  const int np = var <= RANDT_VAR_SE2_XY ? 4 : 3;
  std::vector<double> poses((size_t)B * np), result((size_t)B * RANDT_REG_STRIDE);
  for (uint32_t b = 0; b < B; ++b) {
    if (np == 4) std::memcpy(&poses[4 * b], trans[b].v, 4 * sizeof(double));
    else { poses[3 * b] = trans[b].v[2]; poses[3 * b + 1] = trans[b].v[3]; poses[3 * b + 2] = trans[b].angle(); }   // pos, rot = log()(2)
  }
  randt_loss loss;   // ScaledLoss(BarronLoss(scale, convexity, mu), 1)  (ndt_matcher.cpp:479)
  loss.kind = RANDT_LOSS_BARRON; loss.scale = scale; loss.alpha = parameters_.loss_function_convexity; loss.mu = 1.0; loss.weight = 1.0;
  randt_solver_options opt;
  randt_solver_options_default(&opt);
  opt.max_num_iterations = parameters_.max_iteration;
  // the residual blocks hang off a copy of the pose that carries no manifold (ndt_matcher.cpp:445 vs 452 -> 233): raw ambient parameters
  opt.use_manifold = 0;
  opt.gnc_loss_scale = parameters_.loss_function_scale;
  opt.gnc_divisor = parameters_.gnc_control_parameter_divisor;
  opt.gnc_max_steps = max_gnc_steps;
  ctx_->check(randt_register_batch(ctx_->get(), prob, var, poses.data(), &loss, &opt, result.data()));
  std::vector<double> scores(B);
  for (uint32_t b = 0; b < B; ++b) {
    if (np == 4) std::memcpy(trans[b].v, &poses[4 * b], 4 * sizeof(double));
    else trans[b] = SE2d(poses[3 * b + 2], poses[3 * b], poses[3 * b + 1]);
    scores[b] = result[(size_t)b * RANDT_REG_STRIDE + RANDT_REG_SCORE];
  }
  return scores;
}

double Matcher::estimateLoopConstraint(SE2d& trans, const Map& old_ndt, Map& new_ndt, int max_gnc_steps, bool use_intensity_as_dimension,
                                       double scale) const {
  std::vector<SE2d> t(1, trans);
  const std::vector<double> s = estimateLoopConstraints(t, old_ndt, new_ndt, max_gnc_steps, use_intensity_as_dimension, scale);
  trans = t[0];
  return s[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// Matcher::estimateTransformCeres: the joint window problem
// ---------------------------------------------------------------------------------------------------------------------
namespace {
// addMotionModelFactor (ndt_matcher.cpp:60-110) / addImuFactor (:144-181) over the blocks addMotionParameterBlock (:290-314) and
// addImuParameterBlock (:316-320) registered; the acceleration blocks are constant under the constant-velocity model, the oldest
// state's blocks are all constant
struct WindowBlocks { int pose = -1, pos = -1, rot = -1, lin_vel = -1, rot_vel = -1, lin_acc = -1, imu_bias = -1; };

WindowBlocks addMotionParameterBlock(window::JointProblem& problem, bool manifold, State& X, bool set_constant, bool constant_velocity) {
  WindowBlocks b;
  if (manifold) b.pose = problem.addParameterBlock(X.pose.data(), 4, true);
  else { b.pos = problem.addParameterBlock(X.pos, 2, false); b.rot = problem.addParameterBlock(&X.rot, 1, false); }
  b.lin_vel = problem.addParameterBlock(X.lin_vel, 2, false);
  b.rot_vel = problem.addParameterBlock(&X.rot_vel, 1, false);
  b.lin_acc = problem.addParameterBlock(X.lin_acc, 2, false);
  if (constant_velocity) problem.setParameterBlockConstant(b.lin_acc);
  if (set_constant) {
    // the reference pins only the pose of the oldest state (:304-312) ...
    if (manifold) problem.setParameterBlockConstant(b.pose);
    else { problem.setParameterBlockConstant(b.pos); problem.setParameterBlockConstant(b.rot); }
  }
  return b;
}

void addMotionModelFactor(window::JointProblem& problem, bool manifold, const WindowBlocks& a, const WindowBlocks& b, double dt, const double* sqrtI) {
  window::HostFactor f;
  f.nres = 8;
  if (manifold) {
    f.blocks = {a.pose, a.lin_vel, a.rot_vel, a.lin_acc, b.pose, b.lin_vel, b.rot_vel, b.lin_acc};
    f.eval = [dt, sqrtI](const window::D20* const* p, window::D20* r) { window::MotionModelFactorSE2(dt, sqrtI, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], r); };
  } else {
    f.blocks = {a.pos, a.rot, a.lin_vel, a.rot_vel, a.lin_acc, b.pos, b.rot, b.lin_vel, b.rot_vel, b.lin_acc};
    f.eval = [dt, sqrtI](const window::D20* const* p, window::D20* r) { window::MotionModelFactor(dt, sqrtI, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], r); };
  }
  problem.addFactor(std::move(f));
}

void addImuFactor(window::JointProblem& problem, bool manifold, const WindowBlocks& a, const WindowBlocks& b, double imu_constraint, double weight_imu,
                  double dt, double weight_imu_bias) {
  window::HostFactor f;
  f.nres = 2;
  if (manifold) {
    f.blocks = {a.pose, b.pose, a.imu_bias, b.imu_bias};
    f.eval = [=](const window::D20* const* p, window::D20* r) { window::RotationalResidualSE2(imu_constraint, weight_imu, dt, weight_imu_bias, p[0], p[1], p[2], p[3], r); };
  } else {
    f.blocks = {a.rot, b.rot, a.imu_bias, b.imu_bias};
    f.eval = [=](const window::D20* const* p, window::D20* r) { window::RotationalResidual(imu_constraint, weight_imu, dt, weight_imu_bias, p[0], p[1], p[2], p[3], r); };
  }
  problem.addFactor(std::move(f));
}

// registers the blocks and host factors of the window states[0 .. W] (states[0] constant); returns the blocks per state
std::vector<WindowBlocks> buildWindowFactors(window::JointProblem& problem, const NDTMatcherParameters& p, bool manifold, State* const* states, int W,
                                             const double* imu, const double* sqrtI_scaled) {
  std::vector<WindowBlocks> blocks((size_t)W + 1);
  blocks[0] = addMotionParameterBlock(problem, manifold, *states[0], true, p.use_constant_velocity_model);
  if (p.use_imu) { blocks[0].imu_bias = problem.addParameterBlock(&states[0]->imu_bias, 1, false); problem.setParameterBlockConstant(blocks[0].imu_bias); }
  for (int j = 1; j <= W; ++j) {
    blocks[j] = addMotionParameterBlock(problem, manifold, *states[j], false, p.use_constant_velocity_model);
    if (p.use_imu) blocks[j].imu_bias = problem.addParameterBlock(&states[j]->imu_bias, 1, false);
    const double dt = states[j]->stamp - states[j - 1]->stamp;
    addMotionModelFactor(problem, manifold, blocks[j - 1], blocks[j], dt, sqrtI_scaled);
    if (p.use_imu) addImuFactor(problem, manifold, blocks[j - 1], blocks[j], imu[j - 1], p.weight_imu, dt, p.weight_imu_bias);
  }
  return blocks;
}
}  // namespace

Matcher::~Matcher() { randt_host_free(window_staging_); }

void Matcher::solveWindow(SE2d& trans, std::vector<State>& trajectory, const std::vector<const Map*>& fixed_ndts,
                          const std::vector<const Map*>& moving_window, const std::vector<double>& imu) {
  window_summary_ = WindowSummary();
  const std::chrono::steady_clock::time_point t_begin = std::chrono::steady_clock::now();
  const bool manifold = parameters_.optimize_on_manifold && !parameters_.use_analytic_expressions_for_optimization;
  const bool use_intensity = parameters_.use_intensity_as_dimension;
  const int k = parameters_.n_results_kd_lookup;
  if (trajectory.size() < 2 || fixed_ndts.empty()) { window_summary_.status = 1; return; }
  const size_t W = std::min(trajectory.size() - 1, (size_t)parameters_.smoothing_steps);   // ndt_matcher.cpp:343
  if (moving_window.size() < W) throw Error(RANDT_E_INVALID, "estimateTransformCeres: fewer moving maps than window states");
  if (parameters_.use_imu && imu.size() < W) throw Error(RANDT_E_INVALID, "estimateTransformCeres: one IMU constraint per window factor");
  const double prior_translation[2] = {trans.v[2], trans.v[3]};
  const double prior_rotation = trans.angle();

  // ---- NDT residual blocks: per free state (oldest first), per fixed map, addNDTFactor at the state's own pose (:356-359)
  struct Problems { std::vector<randt_problem*> v; ~Problems() { for (randt_problem* q : v) randt_problem_destroy(q); } } probs;
  std::vector<uint32_t> seg_of_part;
  size_t n_cells = 0, n_blocks = 0;
  for (size_t j = 1; j <= W; ++j) {
    State& X = trajectory[trajectory.size() - 1 - W + j];
    const Map& moving = *moving_window[moving_window.size() - W + (j - 1)];
    if (moving.n_maps() != 1) throw Error(RANDT_E_INVALID, "estimateTransformCeres takes single maps");
    const SE2d guess = X.pose;   // the association always starts from the Lie-group representation (:357), whatever is optimised
    for (const Map* fixed : fixed_ndts) {
      randt_problem* q = associate(&guess, *fixed, moving, use_intensity, k);
      probs.v.push_back(q);
      seg_of_part.push_back((uint32_t)(j - 1));
      uint32_t P = 0;
      ctx_->check(randt_problem_info(q, nullptr, &P, nullptr, nullptr));
      n_blocks += P;
    }
    n_cells += moving.get_n_cells();   // :361
  }
  if (n_blocks == 0 || n_cells == 0) { window_summary_.status = 1; return; }
  // one problem, one segment per window state, joined on the device (cell snapshots, pairs and duos never visit the host)
  randt_problem* merged = nullptr;
  ctx_->check(randt_problem_concat(ctx_->get(), probs.v.data(), (uint32_t)probs.v.size(), seg_of_part.data(), (uint32_t)W, &merged));
  probs.v.push_back(merged);

  // ---- parameter blocks and host factors
  double sqrtI[64];
  for (int i = 0; i < 64; ++i) sqrtI[i] = parameters_.covariance_scaling_factor * parameters_.motion_sqrtI[i];   // :67
  window::JointProblem problem;
  std::vector<State*> states(W + 1);
  for (size_t j = 0; j <= W; ++j) states[j] = &trajectory[trajectory.size() - 1 - W + j];
  const std::vector<WindowBlocks> blocks = buildWindowFactors(problem, parameters_, manifold, states.data(), (int)W, imu.data(), sqrtI);
  window::NdtTerm term;
  const size_t staging = W * (4 + RANDT_FUSED_STRIDE);
  if (window_staging_cap_ < staging) {
    randt_host_free(window_staging_);
    window_staging_cap_ = 0;
    window_staging_ = static_cast<double*>(randt_host_alloc(sizeof(double) * staging));
    if (!window_staging_) throw Error(RANDT_E_NOMEM, "estimateTransformCeres: no pinned host memory");
    window_staging_cap_ = staging;
  }
  term.poses = window_staging_; term.records = window_staging_ + W * 4;
  term.ctx = ctx_->get(); term.problem = merged; term.variant = variant(use_intensity); term.np = manifold ? 4 : 3;
  for (size_t j = 1; j <= W; ++j) {
    if (manifold) term.seg_blocks.push_back({blocks[j].pose});
    else term.seg_blocks.push_back({blocks[j].pos, blocks[j].rot});
  }
  problem.setNdtTerm(term);
  problem.finalize();

  randt_solver_options opt;
  randt_solver_options_default(&opt);
  opt.max_num_iterations = parameters_.max_iteration;
  if (window_tol_[0] > 0.0) opt.function_tolerance = window_tol_[0];
  if (window_tol_[1] > 0.0) opt.parameter_tolerance = window_tol_[1];
  if (window_tol_[2] > 0.0) opt.gradient_tolerance = window_tol_[2];

  const std::chrono::steady_clock::time_point t_built = std::chrono::steady_clock::now();
  window_summary_.setup_us = std::chrono::duration<double, std::micro>(t_built - t_begin).count();
  // ---- GNC loop (:382-397)
  std::vector<double> x((size_t)problem.numAmbient());
  problem.gather(x.data());
  double raw_cost = 0.0, max_residual = 0.0;
  if (!problem.evaluate(x.data(), nullptr, false, &raw_cost, nullptr, nullptr, &max_residual))
    throw Error(RANDT_E_INVALID, std::string("estimateTransformCeres: evaluation failed: ") + randt_last_error(ctx_->get()));
  double gnc_mu = 2.0 * std::pow(max_residual, 2) / std::pow(parameters_.loss_function_scale, 2);
  gnc_mu = std::min(gnc_mu, std::pow(parameters_.gnc_control_parameter_divisor, parameters_.gnc_steps - 1));
  window_summary_.mu_first = gnc_mu; window_summary_.max_residual = max_residual;
  window_summary_.n_free_states = (int)W; window_summary_.n_tangent = problem.numTangent();
  window::MinimizerSummary last;
  do {
    gnc_mu = std::max(gnc_mu, 1.0);
    randt_loss loss;
    loss.kind = RANDT_LOSS_BARRON; loss.scale = parameters_.loss_function_scale; loss.alpha = parameters_.loss_function_convexity; loss.mu = gnc_mu;
    loss.weight = parameters_.ndt_weight / (double)(n_cells * (size_t)k);
    last = window::minimize(problem, loss, opt, x.data());
    if (problem.deviceFailed()) throw Error(RANDT_E_CUDA, std::string("estimateTransformCeres: evaluation failed: ") + randt_last_error(ctx_->get()));
    window_summary_.gnc_solves++;
    window_summary_.total_iterations += last.num_iterations;
    gnc_mu /= parameters_.gnc_control_parameter_divisor;
  } while (gnc_mu > 1.0 / std::sqrt(parameters_.gnc_control_parameter_divisor));
  window_summary_.final_cost = last.final_cost;
  window_summary_.evaluations = problem.evaluations();
  window_summary_.solve_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_built).count();
  problem.scatter(x.data());

  // ---- both representations of the newest state (:399-406), rejection gate (:408-422)
  State& newest = trajectory.back();
  if (!manifold) newest.pose = SE2d(newest.rot, newest.pos[0], newest.pos[1]);
  else { newest.pos[0] = newest.pose.v[2]; newest.pos[1] = newest.pose.v[3]; newest.rot = newest.pose.angle(); }
  SE2d inv_rot; inv_rot.v[0] = newest.pose.v[0]; inv_rot.v[1] = -newest.pose.v[1];
  const double rot_diff = (inv_rot * SE2d(prior_rotation, 0.0, 0.0)).angle();
  if (std::fabs(newest.pose.v[2] - prior_translation[0]) > parameters_.pose_reject_translation ||
      std::fabs(newest.pose.v[3] - prior_translation[1]) > parameters_.pose_reject_translation || std::fabs(rot_diff) > parameters_.pose_reject_rotation) {
    const State& before = trajectory[trajectory.size() - 2];
    window_summary_.rejected = 1;
    newest.pos[0] = before.pos[0]; newest.pos[1] = before.pos[1]; newest.pose = before.pose; newest.rot = before.rot;
    newest.lin_vel[0] = newest.lin_vel[1] = 0.0; newest.rot_vel = 0.0; newest.lin_acc[0] = newest.lin_acc[1] = 0.0;
    newest.imu_bias = before.imu_bias;
  }
  trans = newest.pose;
}

void Matcher::estimateTransformCeres(SE2d& trans, std::vector<State>& trajectory, const double& initial_angle_guess, const double& stamp,
                                     const std::deque<Map>& fixed_ndts, const std::deque<Map>& moving_ndts) {
  std::vector<const Map*> fixed, moving;
  for (const Map& m : fixed_ndts) fixed.push_back(&m);
  for (const Map& m : moving_ndts) moving.push_back(&m);
  estimateTransformCeres(trans, trajectory, initial_angle_guess, stamp, fixed, moving);
}

void Matcher::estimateTransformCeres(SE2d& trans, std::vector<State>& trajectory, const double& /*initial_angle_guess*/, const double& /*stamp*/,
                                     const std::vector<const Map*>& fixed, const std::vector<const Map*>& moving) {
  const size_t W = trajectory.size() < 2 ? 0 : std::min(trajectory.size() - 1, (size_t)parameters_.smoothing_steps);
  // the factor ending at trajectory.end()[-i] reads imu_constraints_.end()[-i-1] (:352); before the first element: 0
  std::vector<double> imu(W, 0.0);
  for (size_t i = 1; i <= W; ++i) {
    const size_t back = i + 1;
    if (back <= imu_constraints_.size()) imu[W - i] = imu_constraints_[imu_constraints_.size() - back];
  }
  solveWindow(trans, trajectory, fixed, moving, imu);
}

double Matcher::estimateTransformGlobalBNB(SE2d& trans, const Map& fixed_ndt, Map& moving_ndt, bool use_intensity_as_dimension, double scale,
                                           double search_window_size_linear, double search_window_size_angular) const {
  search_window_size_linear = std::min(search_window_size_linear, parameters_.csm_window_linear);
  search_window_size_angular = std::min(search_window_size_angular, parameters_.csm_window_angular);
  randt_problem* prob = associate(&trans, fixed_ndt, moving_ndt, use_intensity_as_dimension, 4);   // n_neighbours = 4 (ndt_matcher.cpp:521)
  struct Guard { randt_problem* p; ~Guard() { randt_problem_destroy(p); } } guard{prob};
  uint32_t n_blocks = 0;
  randt_problem_info(prob, nullptr, &n_blocks, nullptr, nullptr);
  const int var = variant(use_intensity_as_dimension);
  randt_loss loss;   // bare BarronLoss(scale, convexity): mu = 1, no ScaledLoss (ndt_matcher.cpp:517)
  loss.kind = RANDT_LOSS_BARRON; loss.scale = scale; loss.alpha = parameters_.loss_function_convexity; loss.mu = 1.0; loss.weight = 1.0;

  const double linear_step = parameters_.csm_linear_step;
  const double max_range = parameters_.csm_max_px_accurate_range;
  const double angular_step = std::acos(1 - ((linear_step * linear_step) / (2 * max_range * max_range)));
  const double cost_threshold = parameters_.csm_cost_threshold;
  const size_t n_iter = (size_t)parameters_.csm_n_iter;
  const double initial_linear_step = std::pow(2, (double)n_iter - 1) * linear_step;

  typedef std::array<float, 9> Key;
  std::set<Key> calculated_points;                       // the reference keeps a vector + std::find (O(n^2)); same membership test
  auto key = [](const SE2d& t) { Key k; t.matrix3f(k.data()); return k; };
  std::vector<std::pair<SE2d, size_t>> level;            // (transform, tree level) in the reference's queue order
  for (double tx = -search_window_size_linear / 2.0; tx <= search_window_size_linear / 2.0; tx += initial_linear_step)
    for (double ty = -search_window_size_linear / 2.0; ty <= search_window_size_linear / 2.0; ty += initial_linear_step)
      for (double a = -search_window_size_angular / 2.0; a < search_window_size_angular / 2.0; a += angular_step) {
        const SE2d cur = trans * SE2d(a, tx, ty);
        level.emplace_back(cur, (size_t)1);
        calculated_points.insert(key(cur));               // the coarsest level is not de-duplicated against itself (push_back only)
      }
  double min_cost = 100000.0;
  SE2d best_trans;
  const int np = var <= RANDT_VAR_SE2_XY ? 4 : 3;
  // The reference pops one transform at a time; children always join the back of the queue, so evaluating everything that is
  // queued in one sweep and then walking the results in queue order visits, prunes and expands exactly the same nodes.
  while (!level.empty()) {
    std::vector<double> poses(level.size() * (size_t)np), cost(level.size());
    for (size_t i = 0; i < level.size(); ++i) {
      const SE2d& t = level[i].first;
      if (np == 4) std::memcpy(&poses[4 * i], t.v, 4 * sizeof(double));
      else { poses[3 * i] = t.v[2]; poses[3 * i + 1] = t.v[3]; poses[3 * i + 2] = t.angle(); }
    }
    ctx_->check(randt_sweep_costs(ctx_->get(), prob, 0, var, poses.data(), (uint32_t)level.size(), &loss, cost.data()));
    std::vector<std::pair<SE2d, size_t>> next;
    for (size_t i = 0; i < level.size(); ++i) {
      const double current_cost = cost[i] / (double)n_blocks;
      const size_t current_level = level[i].second;
      if (current_cost < cost_threshold) {
        if (current_cost < min_cost) { best_trans = level[i].first; min_cost = current_cost; }
        if (current_level < n_iter) {
          const double step = std::pow(2.0, (double)current_level) * linear_step;
          for (double tx = -step; tx <= step; tx += step)
            for (double ty = -step; ty <= step; ty += step)
              for (double a = -angular_step; a <= angular_step; a += angular_step) {
                const SE2d sampled = level[i].first * SE2d(a, tx, ty);
                if (calculated_points.insert(key(sampled)).second) next.emplace_back(sampled, current_level + 1);
              }
        }
      }
    }
    level.swap(next);
  }
  trans = best_trans;
  return min_cost;
}

}  // namespace randt

// ---------------------------------------------------------------------------------------------------------------------
// C hooks for the Python tests
// ---------------------------------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_last_error;
randt::NDTMapParameters map_params(const randt_grid_params& g) {
  randt::NDTMapParameters p;
  p.resolution = g.resolution; p.size_x = g.size_x; p.size_y = g.size_y; p.max_neighbour_manhattan_distance = g.max_linf;
  p.min_points_per_cell = g.min_points; p.max_range = g.max_range; p.n_clusters = g.n_clusters;
  return p;
}
template <typename F>
int guarded(F&& f) {
  try { f(); return RANDT_OK; }
  catch (const randt::Error& e) { g_last_error = e.what(); return e.code; }
  catch (const std::exception& e) { g_last_error = e.what(); return RANDT_E_INVALID; }
}
}  // namespace

extern "C" {

const char* randt_hostapi_last_error(void) { return g_last_error.c_str(); }

int randt_hostapi_loop_constraints(int device, const randt_grid_params* gp, const float* fixed_pts4, const uint32_t* fixed_off,
                                   const float* moving_pts4, const uint32_t* moving_off, uint32_t n, int k, double loss_function_scale,
                                   double convexity, double divisor, int max_gnc_steps, double loop_scale, int optimize_on_manifold,
                                   double* poses_io, double* scores) {
  return guarded([&] {
    randt::Context ctx(device);
    const randt::NDTMapParameters mp = map_params(*gp);
    randt::Map fixed(ctx, mp, n), moving(ctx, mp, n);
    fixed.addClusters(fixed_pts4, fixed_off, n);
    moving.addClusters(moving_pts4, moving_off, n);
    randt::NDTMatcherParameters p;
    p.n_results_kd_lookup = k; p.loss_function_scale = loss_function_scale; p.loss_function_convexity = convexity;
    p.gnc_control_parameter_divisor = divisor; p.optimize_on_manifold = optimize_on_manifold != 0;
    randt::Matcher m(ctx);
    m.initialize(p);
    std::vector<randt::SE2d> t(n);
    for (uint32_t b = 0; b < n; ++b) std::memcpy(t[b].v, poses_io + 4 * b, 4 * sizeof(double));
    std::vector<double> s;
    if (n == 1) { s.push_back(m.estimateLoopConstraint(t[0], fixed, moving, max_gnc_steps, true, loop_scale)); }
    else s = m.estimateLoopConstraints(t, fixed, moving, max_gnc_steps, true, loop_scale);
    for (uint32_t b = 0; b < n; ++b) { std::memcpy(poses_io + 4 * b, t[b].v, 4 * sizeof(double)); scores[b] = s[b]; }
  });
}

int randt_hostapi_cost_function(int device, const randt_grid_params* gp, const float* fixed_pts4, uint32_t n_fixed, const float* moving_pts4,
                                uint32_t n_moving, int k, const double* guess4, const randt_loss* loss, const double* pose4, double* residuals,
                                double* jacobian, uint32_t cap, uint32_t* num_residuals, double* max_raw) {
  return guarded([&] {
    randt::Context ctx(device);
    const randt::NDTMapParameters mp = map_params(*gp);
    randt::Map fixed(ctx, mp), moving(ctx, mp);
    const uint32_t fo[2] = {0, n_fixed}, mo[2] = {0, n_moving};
    fixed.addClusters(fixed_pts4, fo, 1);
    moving.addClusters(moving_pts4, mo, 1);
    randt::NDTMatcherParameters p;
    p.n_results_kd_lookup = k;
    randt::Matcher m(ctx);
    m.initialize(p);
    randt::SE2d g;
    std::memcpy(g.v, guess4, 4 * sizeof(double));
    std::unique_ptr<randt::NdtCostFunction> cf = m.addNDTFactor(g, fixed, moving, true, k);
    const ceres::CostFunction* base = cf.get();   // driven through the ceres interface only
    *num_residuals = (uint32_t)base->num_residuals();
    if ((uint32_t)base->num_residuals() > cap) throw randt::Error(RANDT_E_CAPACITY, "residual buffer too small");
    if (base->parameter_block_sizes().size() != 1 || base->parameter_block_sizes()[0] != 4) throw randt::Error(RANDT_E_INVALID, "unexpected block sizes");
    cf->setLoss(loss);
    const double* params[1] = {pose4};
    double* jac[1] = {jacobian};
    if (!base->Evaluate(params, residuals, jacobian ? jac : nullptr)) throw randt::Error(RANDT_E_NONFINITE, "Evaluate returned false");
    if (max_raw) *max_raw = cf->maxRawResidual(pose4);
  });
}

int randt_hostapi_odometry(int device, const randt_grid_params* gp, const float* const* fixed_pts4, const uint32_t* n_fixed_pts, const double* fixed_pose4,
                           uint32_t n_fixed, const float* moving_pts4, uint32_t n_moving, int k, double loss_function_scale, double convexity,
                           double divisor, int gnc_steps, double ndt_weight, int optimize_on_manifold, double reject_translation,
                           double reject_rotation, double* pose_io4, int* accepted) {
  return guarded([&] {
    randt::Context ctx(device);
    const randt::NDTMapParameters mp = map_params(*gp);
    std::vector<std::unique_ptr<randt::Map>> fixed;
    std::vector<const randt::Map*> fixed_ptr;
    for (uint32_t i = 0; i < n_fixed; ++i) {
      fixed.emplace_back(new randt::Map(ctx, mp));
      const uint32_t off[2] = {0, n_fixed_pts[i]};
      fixed.back()->addClusters(fixed_pts4[i], off, 1);
      randt::SE2d T;
      std::memcpy(T.v, fixed_pose4 + 4 * i, 4 * sizeof(double));
      fixed.back()->transformMap(&T);
      fixed_ptr.push_back(fixed.back().get());
    }
    randt::Map moving(ctx, mp);
    const uint32_t mo[2] = {0, n_moving};
    moving.addClusters(moving_pts4, mo, 1);
    randt::NDTMatcherParameters p;
    p.n_results_kd_lookup = k; p.loss_function_scale = loss_function_scale; p.loss_function_convexity = convexity;
    p.gnc_control_parameter_divisor = divisor; p.gnc_steps = gnc_steps; p.ndt_weight = ndt_weight; p.optimize_on_manifold = optimize_on_manifold != 0;
    p.pose_reject_translation = reject_translation; p.pose_reject_rotation = reject_rotation;
    randt::Matcher m(ctx);
    m.initialize(p);
    randt::SE2d t;
    std::memcpy(t.v, pose_io4, 4 * sizeof(double));
    *accepted = m.estimateTransformNDT(t, fixed_ptr, moving) ? 1 : 0;
    std::memcpy(pose_io4, t.v, 4 * sizeof(double));
  });
}

int randt_hostapi_eval_async_loop(randt_ctx* ctx, const randt_problem* problem, int variant, const randt_loss* loss, double* const* poses_ring,
                                  double* const* out_ring, uint32_t depth, uint32_t steps, int packed) {
  return guarded([&] {
    if (!ctx || !problem || !poses_ring || !out_ring || depth < 1) throw randt::Error(RANDT_E_INVALID, "eval_async_loop: needs buffer sets");
    auto check = [&](int rc) { if (rc != RANDT_OK) throw randt::Error(rc, randt_last_error(ctx)); };
    std::vector<uint64_t> ticket(depth, 0);
    for (uint32_t i = 0; i < steps; ++i) {
      const uint32_t j = i % depth;
      // buffer set j is reused: the call that had it (step i - depth) must have delivered.  The device slots are ordered by events inside
      // the library, so the host may run up to `depth` steps ahead and the next upload is never waiting for the host.
      if (ticket[j]) check(randt_ctx_wait_async(ctx, ticket[j]));
      check(randt_eval_fused_async(ctx, problem, variant, poses_ring[j], loss, nullptr, 1, packed, out_ring[j]));
      ticket[j] = randt_ctx_async_count(ctx);
    }
    check(randt_ctx_sync(ctx));
  });
}

int randt_hostapi_build_schedule(const uint32_t* duo_off, uint32_t n_segments, uint32_t max_warps, uint32_t* counts, uint32_t* tiles4, uint32_t cap_tiles,
                                 uint32_t* plan_a4, uint32_t* plan_b4, uint32_t cap_chunks, uint32_t* woff_a, uint32_t* woff_b,
                                 uint32_t* tile_rec_begin, uint32_t* tile_duo_begin, uint32_t* first) {
  return guarded([&] {
    if (!duo_off || !counts || max_warps == 0) throw randt::Error(RANDT_E_INVALID, "build_schedule: null argument");
    randt::Schedule sch;
    randt::build_schedule(duo_off, n_segments, max_warps, sch);
    counts[0] = (uint32_t)sch.tiles.size(); counts[1] = sch.n_warps; counts[2] = (uint32_t)sch.planA.size(); counts[3] = (uint32_t)sch.planB.size();
    counts[4] = sch.tile_rec_begin.back();
    if (sch.tiles.size() > cap_tiles || sch.planA.size() > cap_chunks || sch.planB.size() > cap_chunks)
      throw randt::Error(RANDT_E_CAPACITY, "build_schedule: output arrays too small");
    static_assert(sizeof(randt::Tile) == 16 && sizeof(randt::ChunkDesc) == 16, "4 x uint32 records");
    if (tiles4 && !sch.tiles.empty()) std::memcpy(tiles4, sch.tiles.data(), sch.tiles.size() * 16);
    if (plan_a4 && !sch.planA.empty()) std::memcpy(plan_a4, sch.planA.data(), sch.planA.size() * 16);
    if (plan_b4 && !sch.planB.empty()) std::memcpy(plan_b4, sch.planB.data(), sch.planB.size() * 16);
    if (woff_a) std::memcpy(woff_a, sch.woffA.data(), sch.woffA.size() * 4);
    if (woff_b) std::memcpy(woff_b, sch.woffB.data(), sch.woffB.size() * 4);
    if (tile_rec_begin) std::memcpy(tile_rec_begin, sch.tile_rec_begin.data(), sch.tile_rec_begin.size() * 4);
    if (tile_duo_begin && !sch.tile_duo_begin.empty()) std::memcpy(tile_duo_begin, sch.tile_duo_begin.data(), sch.tile_duo_begin.size() * 4);
    if (first) std::memcpy(first, sch.first.data(), sch.first.size() * 4);
  });
}

void randt_hostapi_predict(int se2_model, const double* s, double raw_dt, double* out) {
  randt::State a, b;
  std::memcpy(a.pose.v, s, 4 * sizeof(double));
  a.pos[0] = s[4]; a.pos[1] = s[5]; a.rot = s[6]; a.lin_vel[0] = s[7]; a.lin_vel[1] = s[8]; a.rot_vel = s[9]; a.lin_acc[0] = s[10]; a.lin_acc[1] = s[11];
  b = a;
  if (se2_model) randt::predictSE2(a, raw_dt, b); else randt::predict(a, raw_dt, b);
  std::memcpy(out, b.pose.v, 4 * sizeof(double));
  out[4] = b.pos[0]; out[5] = b.pos[1]; out[6] = b.rot; out[7] = b.lin_vel[0]; out[8] = b.lin_vel[1]; out[9] = b.rot_vel; out[10] = b.lin_acc[0]; out[11] = b.lin_acc[1];
}

namespace {
void state_from14(const double* s, randt::State& X) {
  std::memcpy(X.pose.v, s, 4 * sizeof(double));
  X.pos[0] = s[4]; X.pos[1] = s[5]; X.rot = s[6]; X.lin_vel[0] = s[7]; X.lin_vel[1] = s[8]; X.rot_vel = s[9];
  X.lin_acc[0] = s[10]; X.lin_acc[1] = s[11]; X.imu_bias = s[12]; X.stamp = s[13];
}
void state_to14(const randt::State& X, double* s) {
  std::memcpy(s, X.pose.v, 4 * sizeof(double));
  s[4] = X.pos[0]; s[5] = X.pos[1]; s[6] = X.rot; s[7] = X.lin_vel[0]; s[8] = X.lin_vel[1]; s[9] = X.rot_vel;
  s[10] = X.lin_acc[0]; s[11] = X.lin_acc[1]; s[12] = X.imu_bias; s[13] = X.stamp;
}
randt::NDTMatcherParameters matcher_params80(const double* q) {
  randt::NDTMatcherParameters p;
  p.n_results_kd_lookup = (int)q[0]; p.gnc_steps = (int)q[1]; p.max_iteration = (int)q[2]; p.loss_function_scale = q[3];
  p.loss_function_convexity = q[4]; p.gnc_control_parameter_divisor = q[5]; p.ndt_weight = q[6]; p.optimize_on_manifold = q[7] != 0.0;
  p.use_constant_velocity_model = q[8] != 0.0; p.use_imu = q[9] != 0.0; p.weight_imu = q[10]; p.weight_imu_bias = q[11];
  p.pose_reject_translation = q[12]; p.pose_reject_rotation = q[13]; p.use_intensity_as_dimension = q[14] != 0.0;
  p.covariance_scaling_factor = 1.0;   // the caller passes the scaled matrix
  for (int i = 0; i < 64; ++i) p.motion_sqrtI[i] = q[16 + i];
  return p;
}
}  // namespace

int randt_hostapi_window_factors(const double* states, uint32_t W, const double* imu, const double* params80, double* cost, double* g, double* H) {
  int nt = 0;
  const int rc = guarded([&] {
    const randt::NDTMatcherParameters p = matcher_params80(params80);
    const bool manifold = p.optimize_on_manifold && !p.use_analytic_expressions_for_optimization;
    std::vector<randt::State> st((size_t)W + 1);
    std::vector<randt::State*> ptr((size_t)W + 1);
    for (uint32_t j = 0; j <= W; ++j) { state_from14(states + 14 * (size_t)j, st[j]); ptr[j] = &st[j]; }
    std::vector<double> zero(W, 0.0);
    randt::window::JointProblem problem;
    randt::buildWindowFactors(problem, p, manifold, ptr.data(), (int)W, imu ? imu : zero.data(), p.motion_sqrtI);
    problem.finalize();
    nt = problem.numTangent();
    std::vector<double> x((size_t)problem.numAmbient());
    problem.gather(x.data());
    *cost = 0.0;
    std::fill(g, g + nt, 0.0); std::fill(H, H + (size_t)nt * nt, 0.0);
    problem.evaluateHostFactors(x.data(), true, cost, g, H);
  });
  return rc == RANDT_OK ? nt : rc;
}

int randt_hostapi_window_minimize_factors(double* states, uint32_t W, const double* imu, const double* params80, const double* tolerances3,
                                          int max_iterations, double* summary4) {
  return guarded([&] {
    const randt::NDTMatcherParameters p = matcher_params80(params80);
    const bool manifold = p.optimize_on_manifold && !p.use_analytic_expressions_for_optimization;
    std::vector<randt::State> st((size_t)W + 1);
    std::vector<randt::State*> ptr((size_t)W + 1);
    for (uint32_t j = 0; j <= W; ++j) { state_from14(states + 14 * (size_t)j, st[j]); ptr[j] = &st[j]; }
    std::vector<double> zero(W, 0.0);
    randt::window::JointProblem problem;
    randt::buildWindowFactors(problem, p, manifold, ptr.data(), (int)W, imu ? imu : zero.data(), p.motion_sqrtI);
    problem.finalize();
    randt_solver_options opt;
    randt_solver_options_default(&opt);
    if (max_iterations > 0) opt.max_num_iterations = max_iterations;
    if (tolerances3) {
      if (tolerances3[0] > 0.0) opt.function_tolerance = tolerances3[0];
      if (tolerances3[1] > 0.0) opt.parameter_tolerance = tolerances3[1];
      if (tolerances3[2] > 0.0) opt.gradient_tolerance = tolerances3[2];
    }
    std::vector<double> x((size_t)problem.numAmbient());
    problem.gather(x.data());
    randt_loss none;
    none.kind = RANDT_LOSS_NONE; none.scale = 1.0; none.alpha = 2.0; none.mu = 1.0; none.weight = 1.0;
    const randt::window::MinimizerSummary ms = randt::window::minimize(problem, none, opt, x.data());
    problem.scatter(x.data());
    for (uint32_t j = 0; j <= W; ++j) state_to14(st[j], states + 14 * (size_t)j);
    summary4[0] = ms.initial_cost; summary4[1] = ms.final_cost; summary4[2] = ms.num_iterations; summary4[3] = ms.termination;
  });
}

int randt_hostapi_window_solve(int device, const randt_grid_params* gp, const float* const* fixed_pts4, const uint32_t* n_fixed_pts,
                               const double* fixed_pose4, uint32_t n_fixed, const float* const* window_pts4, const uint32_t* n_window_pts, uint32_t W,
                               double* states, const double* imu, const double* params80, const double* tolerances3, double* trans4, double* out10) {
  return guarded([&] {
    randt::Context ctx(device);
    const randt::NDTMapParameters mp = map_params(*gp);
    std::vector<std::unique_ptr<randt::Map>> maps;
    std::vector<const randt::Map*> fixed, window;
    for (uint32_t i = 0; i < n_fixed; ++i) {
      maps.emplace_back(new randt::Map(ctx, mp));
      const uint32_t off[2] = {0, n_fixed_pts[i]};
      maps.back()->addClusters(fixed_pts4[i], off, 1);
      randt::SE2d T;
      std::memcpy(T.v, fixed_pose4 + 4 * i, 4 * sizeof(double));
      maps.back()->transformMap(&T);
      fixed.push_back(maps.back().get());
    }
    size_t n_cells = 0;
    for (uint32_t w = 0; w < W; ++w) {
      maps.emplace_back(new randt::Map(ctx, mp));
      const uint32_t off[2] = {0, n_window_pts[w]};
      maps.back()->addClusters(window_pts4[w], off, 1);
      window.push_back(maps.back().get());
      n_cells += maps.back()->get_n_cells();
    }
    randt::NDTMatcherParameters p = matcher_params80(params80);
    p.smoothing_steps = (int)W;
    randt::Matcher m(ctx);
    m.initialize(p);
    if (tolerances3) m.setWindowTolerances(tolerances3[0], tolerances3[1], tolerances3[2]);
    std::vector<randt::State> trajectory((size_t)W + 1);
    for (uint32_t j = 0; j <= W; ++j) state_from14(states + 14 * (size_t)j, trajectory[j]);
    randt::SE2d t;
    std::memcpy(t.v, trans4, 4 * sizeof(double));
    std::vector<double> imu_v(W, 0.0);
    if (imu) imu_v.assign(imu, imu + W);
    m.solveWindow(t, trajectory, fixed, window, imu_v);
    std::memcpy(trans4, t.v, 4 * sizeof(double));
    for (uint32_t j = 0; j <= W; ++j) state_to14(trajectory[j], states + 14 * (size_t)j);
    const randt::WindowSummary& ws = m.lastWindowSummary();
    out10[0] = ws.status; out10[1] = ws.rejected; out10[2] = ws.gnc_solves; out10[3] = ws.total_iterations; out10[4] = ws.final_cost;
    out10[5] = ws.mu_first; out10[6] = ws.max_residual; out10[7] = ws.n_tangent; out10[8] = ws.evaluations; out10[9] = (double)n_cells;
  });
}

int randt_hostapi_window_replay(int device, const randt_grid_params* gp, const float* pts4, const uint32_t* scan_off, uint32_t n_scans,
                                const double* stamps, const double* yaw, const double* params80, int smoothing_steps, int insertion_step,
                                double* poses_out, double* states_out, double* stats_out, double* totals) {
  return guarded([&] {
    using clock = std::chrono::steady_clock;
    randt::Context ctx(device);
    const randt::NDTMapParameters mp = map_params(*gp);
    randt::NDTMatcherParameters p = matcher_params80(params80);
    p.smoothing_steps = smoothing_steps;
    const bool manifold = p.optimize_on_manifold && !p.use_analytic_expressions_for_optimization;
    const size_t insertion_delay = (size_t)smoothing_steps + 1;                  // ndt_slam.cpp:580
    randt::Matcher matcher(ctx);
    matcher.initialize(p);
    std::vector<randt::State> trajectory;
    std::deque<std::shared_ptr<randt::Map>> map_window, next_maps_to_insert;
    std::shared_ptr<randt::Map> submap;      // _current_submap: becomes the first scan, then grows by mergeMapCell
    randt::SE2d current_transform;
    uint32_t keyframes = 0;
    totals[4] = totals[5] = 0.0;
    const uint64_t launches0 = ctx.launchCount();
    const clock::time_point t_begin = clock::now();
    for (uint32_t i = 0; i < n_scans; ++i) {
      const clock::time_point t0 = clock::now();
      std::shared_ptr<randt::Map> current_scan(new randt::Map(ctx, mp));
      const uint32_t off[2] = {0, scan_off[i + 1] - scan_off[i]};
      current_scan->addClusters(pts4 + 4 * (size_t)scan_off[i], off, 1);
      double* st = stats_out + 4 * (size_t)i;
      st[0] = st[1] = st[2] = 0.0;
      if (submap && submap->get_n_cells() > 0) {
        matcher.predictTransform(yaw ? yaw[i] : 0.0, stamps[i], trajectory);
        map_window.push_back(current_scan);
        std::vector<const randt::Map*> f_maps{submap.get()}, window;
        for (const auto& m : map_window) window.push_back(m.get());
        matcher.estimateTransformCeres(current_transform, trajectory, yaw ? yaw[i] : 0.0, stamps[i], f_maps, window);
        const randt::WindowSummary& ws = matcher.lastWindowSummary();
        st[0] = ws.total_iterations; st[1] = ws.evaluations; st[2] = ws.rejected;
        totals[4] += ws.setup_us * 1e-6; totals[5] += ws.solve_us * 1e-6;
        // write both pose representations (local_fuser.cpp:141-150)
        for (size_t b = 1; b <= std::min((size_t)smoothing_steps, trajectory.size()); ++b) {
          randt::State& X = trajectory[trajectory.size() - b];
          if (!manifold) X.pose = randt::SE2d(X.rot, X.pos[0], X.pos[1]);
          else { X.pos[0] = X.pose.v[2]; X.pos[1] = X.pose.v[3]; X.rot = X.pose.angle(); }
        }
        const size_t trajectory_size = trajectory.size();
        if (map_window.size() >= (size_t)smoothing_steps) map_window.pop_front();
        if (trajectory_size % (size_t)insertion_step == 0) next_maps_to_insert.push_back(current_scan);   // keyframe, pushed on the buffer
        if (trajectory_size >= insertion_delay + (size_t)insertion_step && (trajectory_size - insertion_delay) % (size_t)insertion_step == 0 &&
            !next_maps_to_insert.empty()) {
          // the keyframe leaves the estimator: insert it at its smoothed pose (:164-190)
          const randt::State& X_smoothed = trajectory[trajectory_size - insertion_delay - 1];
          const randt::SE2d smoothed = manifold ? X_smoothed.pose : randt::SE2d(X_smoothed.rot, X_smoothed.pos[0], X_smoothed.pos[1]);
          next_maps_to_insert.front()->transformMap(smoothed);
          submap->mergeMapCell(*next_maps_to_insert.front());
          next_maps_to_insert.pop_front();
          ++keyframes;
        }
      } else {
        // the first scan of the submap (:221-296)
        randt::State initial_state;
        initial_state.pose = current_transform;
        initial_state.pos[0] = current_transform.v[2]; initial_state.pos[1] = current_transform.v[3]; initial_state.rot = current_transform.angle();
        initial_state.stamp = stamps[i];
        trajectory.push_back(initial_state);
        current_scan->transformMap(current_transform);
        submap = current_scan;               // mergeMapCell into the empty submap: every cell is appended as it is
      }
      std::memcpy(poses_out + 4 * (size_t)i, current_transform.v, 4 * sizeof(double));
      st[3] = std::chrono::duration<double, std::micro>(clock::now() - t0).count();
    }
    totals[0] = std::chrono::duration<double>(clock::now() - t_begin).count();
    totals[1] = submap ? (double)submap->get_n_cells() : 0.0;
    totals[2] = (double)(ctx.launchCount() - launches0);
    totals[3] = (double)keyframes;
    for (uint32_t i = 0; i < n_scans && i < trajectory.size(); ++i) state_to14(trajectory[i], states_out + 14 * (size_t)i);
  });
}

int randt_hostapi_export(int device, const randt_grid_params* gp, const float* pts4, uint32_t n_pts, double* mean3, double* cov6, uint32_t cap,
                         uint32_t* n_cells) {
  return guarded([&] {
    randt::Context ctx(device);
    randt::Map map(ctx, map_params(*gp));
    const uint32_t off[2] = {0, n_pts};
    map.addClusters(pts4, off, 1);
    std::vector<double> m, c;
    map.exportNormalDistributions(m, c);
    *n_cells = (uint32_t)(m.size() / 3);
    if (*n_cells > cap) throw randt::Error(RANDT_E_CAPACITY, "export buffer too small");
    std::memcpy(mean3, m.data(), m.size() * sizeof(double));
    std::memcpy(cov6, c.data(), c.size() * sizeof(double));
  });
}

int randt_hostapi_bnb(int device, const randt_grid_params* gp, const float* fixed_pts4, uint32_t n_fixed, const float* moving_pts4, uint32_t n_moving,
                      double convexity, double scale, double window_linear, double window_angular, double linear_step, double max_px_range,
                      double cost_threshold, int n_iter, double* pose_io4, double* min_cost, uint32_t* n_evaluated) {
  return guarded([&] {
    randt::Context ctx(device);
    const randt::NDTMapParameters mp = map_params(*gp);
    randt::Map fixed(ctx, mp), moving(ctx, mp);
    const uint32_t fo[2] = {0, n_fixed}, mo[2] = {0, n_moving};
    fixed.addClusters(fixed_pts4, fo, 1);
    moving.addClusters(moving_pts4, mo, 1);
    randt::NDTMatcherParameters p;
    p.loss_function_convexity = convexity; p.csm_linear_step = linear_step; p.csm_max_px_accurate_range = max_px_range;
    p.csm_cost_threshold = cost_threshold; p.csm_n_iter = n_iter; p.csm_window_linear = window_linear; p.csm_window_angular = window_angular;
    randt::Matcher m(ctx);
    m.initialize(p);
    randt::SE2d t;
    std::memcpy(t.v, pose_io4, 4 * sizeof(double));
    const uint64_t l0 = ctx.launchCount();
    *min_cost = m.estimateTransformGlobalBNB(t, fixed, moving, true, scale, window_linear, window_angular);
    if (n_evaluated) *n_evaluated = (uint32_t)(ctx.launchCount() - l0);
    std::memcpy(pose_io4, t.v, 4 * sizeof(double));
  });
}

}  // extern "C"
