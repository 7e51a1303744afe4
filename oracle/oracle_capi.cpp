// ORACLE — TEST INFRASTRUCTURE ONLY (see ndt_oracle.h header).  PARITY UNPINNED.
// Flat-array C exports of the CPU restatement so the Python tests / bench baseline can drive it via ctypes.
#include <chrono>
#include <cstdio>
#include <thread>

#include <atomic>

#include "lm_oracle.h"
#include "window_oracle.h"

using namespace orc;

namespace {
NdtMap view_map(const float* cells, const uint32_t* npts, int n, const int32_t* slot, int size_x, int size_y, double res, double max_linf) {
  NdtMap m;
  MapGeom g; g.set(size_x, size_y, res, max_linf);
  m.geom = g;
  m.cells.resize(n);
  if (n) std::memcpy(m.cells.data(), cells, sizeof(Cell12) * n);
  m.npts.assign(n, 0);
  if (npts) for (int i = 0; i < n; ++i) m.npts[i] = npts[i];
  if (slot) m.slot.assign(slot, slot + g.n_slots()); else m.slot.assign(g.n_slots(), -1);
  return m;
}
Loss make_loss(int kind, double a, double alpha, double mu, double weight) {
  Loss l; l.kind = kind; l.a = a; l.alpha = alpha; l.mu = mu; l.weight = weight; return l;
}
}  // namespace

extern "C" {

int orc_n_clusters(double max_range, double resolution) { return n_clusters_from_params(max_range, resolution); }
int orc_grid_row_size(size_t n_clusters) { return grid_row_size(n_clusters); }

void orc_grid_labels(const float* pts4, size_t n, size_t n_clusters, float max_range, int32_t* labels) {
  grid_labels(reinterpret_cast<const Pt4*>(pts4), n, n_clusters, max_range, labels);
}

uint32_t orc_coord_to_index(int size_x, int size_y, double res, float x, float y) {
  MapGeom g; g.set(size_x, size_y, res, 0.0);
  return coord_to_index(g, x, y);
}

void orc_sym2_eigen(float a00, float a10, float a11, float* ev2, float* V4) {
  float V[2][2];
  sym2_eigen_f(a00, a10, a11, ev2, V);
  V4[0] = V[0][0]; V4[1] = V[0][1]; V4[2] = V[1][0]; V4[3] = V[1][1];
}

int orc_cell_from_points(const float* pts4, size_t n, int min_points, float* cell12) {
  Cell12 c;
  if (!cell_from_points(reinterpret_cast<const Pt4*>(pts4), nullptr, n, min_points, c)) return 0;
  std::memcpy(cell12, &c, sizeof(c));
  return 1;
}

// returns number of cells (<= cap); slot_out has size_x*size_y entries
int orc_voxelize(const float* pts4, size_t n, size_t n_clusters, float max_range, int min_points, int size_x, int size_y,
                 double res, double max_linf, float* cells_out, uint32_t* npts_out, int32_t* labels_out, int32_t* slot_out,
                 int cap, int* dropped) {
  MapGeom g; g.set(size_x, size_y, res, max_linf);
  NdtMap m;
  std::vector<int32_t> labels;
  voxelize(reinterpret_cast<const Pt4*>(pts4), n, n_clusters, max_range, min_points, g, m, &labels);
  const int nc = (int)m.cells.size();
  if (nc > cap) return -nc;
  if (nc) std::memcpy(cells_out, m.cells.data(), sizeof(Cell12) * nc);
  for (int i = 0; i < nc; ++i) { if (npts_out) npts_out[i] = m.npts[i]; if (labels_out) labels_out[i] = labels[i]; }
  if (slot_out) std::memcpy(slot_out, m.slot.data(), sizeof(int32_t) * m.slot.size());
  if (dropped) *dropped = m.dropped_out_of_map;
  return nc;
}

void orc_se2d_cast_float(const double* pose, float* out4) { se2d_cast_float(pose, out4); }
void orc_affine_rotation(float c, float s, float* R9) { float R[3][3]; affine_rotation_f(c, s, R); for (int i = 0; i < 9; ++i) R9[i] = R[i / 3][i % 3]; }
void orc_transform_cells(float* cells, size_t n, float c, float s, float tx, float ty) {
  Cell12* cc = reinterpret_cast<Cell12*>(cells);
  float R[3][3];
  affine_rotation_f(c, s, R);
  for (size_t i = 0; i < n; ++i) transform_cell(cc[i], c, s, tx, ty, R);
}

// in-place merge of a moving map into a fixed map (arrays have capacity cap_f); returns new fixed cell count or -needed
int orc_merge_map_cell(float* f_cells, uint32_t* f_npts, int n_f, int cap_f, int32_t* f_slot, int size_x, int size_y, double res,
                       const float* m_cells, const uint32_t* m_npts, int n_m) {
  NdtMap F = view_map(f_cells, f_npts, n_f, f_slot, size_x, size_y, res, 0.0);
  NdtMap M = view_map(m_cells, m_npts, n_m, nullptr, size_x, size_y, res, 0.0);
  merge_map_cell(F, M);
  const int nf = (int)F.cells.size();
  if (nf > cap_f) return -nf;
  std::memcpy(f_cells, F.cells.data(), sizeof(Cell12) * nf);
  for (int i = 0; i < nf; ++i) f_npts[i] = F.npts[i];
  std::memcpy(f_slot, F.slot.data(), sizeof(int32_t) * F.slot.size());
  return nf;
}

// returns P (number of pairs) or -needed when cap is too small
int orc_associate(const float* f_cells, int n_f, const int32_t* f_slot, int size_x, int size_y, double res, double max_linf,
                  const float* m_cells, int n_m, const double* pose4, int k, int metric, uint32_t* im_out, uint32_t* jf_out, int cap) {
  NdtMap F = view_map(f_cells, nullptr, n_f, f_slot, size_x, size_y, res, max_linf);
  NdtMap M = view_map(m_cells, nullptr, n_m, nullptr, size_x, size_y, res, max_linf);
  PairList pl;
  associate(F, M, pose4, k, metric, pl);
  const int P = (int)pl.im.size();
  if (P > cap) return -P;
  for (int i = 0; i < P; ++i) { im_out[i] = pl.im[i]; jf_out[i] = pl.jf[i]; }
  return P;
}

// mode: 0 = autodiff (Jet), 1 = closed form (VAR_SE2_INTENSITY only), 2 = value only
void orc_eval_pairs(int variant, int mode, const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf,
                    size_t P, const double* params, double* r_out, double* J_out) {
  const Cell12* cm = reinterpret_cast<const Cell12*>(cells_m);
  const Cell12* cf = reinterpret_cast<const Cell12*>(cells_f);
  const int np = variant_num_params(variant);
  for (size_t p = 0; p < P; ++p) {
    double r, J[4];
    if (mode == 0) eval_pair_autodiff(variant, params, cm[im[p]], cf[jf[p]], &r, J);
    else if (mode == 1) eval_pair_closed_form(params, cm[im[p]], cf[jf[p]], &r, J);
    else { r = eval_pair_value(variant, params, cm[im[p]], cf[jf[p]]); }
    r_out[p] = r;
    if (J_out && mode != 2) for (int a = 0; a < np; ++a) J_out[p * np + a] = J[a];
  }
}

void orc_loss_eval(int kind, double a, double alpha, double mu, double weight, double s, double* rho3) {
  make_loss(kind, a, alpha, mu, weight).evaluate(s, rho3);
}

void orc_corrector(double sq_norm, const double* rho3, double* out3) {
  Corrector c(sq_norm, rho3);
  out3[0] = c.sqrt_rho1; out3[1] = c.residual_scaling; out3[2] = c.alpha_sq_norm;
}

double orc_gnc_initial_mu(double max_residual, double loss_scale, double divisor, int steps) {
  return gnc_initial_mu(max_residual, loss_scale, divisor, steps);
}

// out24 = H[16], g[4], cost, max_r, sum_sq, n
static void pack_fused(const FusedOut& f, double* out24) {
  for (int i = 0; i < 16; ++i) out24[i] = f.H[i];
  for (int i = 0; i < 4; ++i) out24[16 + i] = f.g[i];
  out24[20] = f.cost; out24[21] = f.max_r; out24[22] = f.sum_sq; out24[23] = (double)f.n;
}

void orc_fused(int variant, const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf, size_t P,
               const double* params, int loss_kind, double a, double alpha, double mu, double weight, int want_jac, double* out24) {
  FusedOut f;
  accumulate_pairs(variant, params, reinterpret_cast<const Cell12*>(cells_m), reinterpret_cast<const Cell12*>(cells_f), im, jf, P,
                   make_loss(loss_kind, a, alpha, mu, weight), want_jac != 0, f);
  pack_fused(f, out24);
}

int orc_hw_threads() { return (int)std::thread::hardware_concurrency(); }

// Batched: S segments; pairs of segment s are [seg_off[s], seg_off[s+1]); im/jf index the concatenated tables.
// pose stride = 4 (SE2 variants) or 3.  mu may be per segment (mu_per_seg != NULL) else scalar `mu`.
// Returns wall seconds for `repeats` passes.  closed_form = 1 uses the closed form instead of the Jet path (informational).
double orc_fused_batch(int variant, const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf,
                       const uint32_t* seg_off, int S, const double* poses, int loss_kind, double a, double alpha, double mu,
                       double weight, const double* mu_per_seg, int want_jac, double* out24, int n_threads, int repeats) {
  const Cell12* cm = reinterpret_cast<const Cell12*>(cells_m);
  const Cell12* cf = reinterpret_cast<const Cell12*>(cells_f);
  const int np = variant_num_params(variant);
  const auto t0 = std::chrono::steady_clock::now();
  const int T = n_threads > 0 ? n_threads : 1;
  for (int rep = 0; rep < repeats; ++rep) {
    std::atomic<int> next(0);
    auto worker = [&]() {
      for (;;) {
        const int s0 = next.fetch_add(4);
        if (s0 >= S) break;
        for (int s = s0; s < std::min(S, s0 + 4); ++s) {
          FusedOut f;
          const Loss loss = make_loss(loss_kind, a, alpha, mu_per_seg ? mu_per_seg[s] : mu, weight);
          accumulate_pairs(variant, poses + (size_t)s * np, cm, cf, im + seg_off[s], jf + seg_off[s], seg_off[s + 1] - seg_off[s],
                           loss, want_jac != 0, f);
          pack_fused(f, out24 + (size_t)s * 24);
        }
      }
    };
    if (T == 1) { worker(); }
    else {
      std::vector<std::thread> pool;
      for (int t = 0; t < T; ++t) pool.emplace_back(worker);
      for (auto& th : pool) th.join();
    }
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Matcher::estimateLoopConstraint restated.  out = pose[4], score, gnc_solves, total_iterations, total_evals, mu_first ; returns status
// ceres tolerances for the next orc_loop_constraint calls of this thread (<= 0: ceres' defaults): lets a test remove the slack the
// default function_tolerance leaves when two implementations are compared at their stopping points
static thread_local double g_function_tol = 0.0, g_parameter_tol = 0.0, g_gradient_tol = 0.0;
void orc_set_tolerances(double function_tol, double parameter_tol, double gradient_tol) {
  g_function_tol = function_tol; g_parameter_tol = parameter_tol; g_gradient_tol = gradient_tol;
}

int orc_loop_constraint(const float* f_cells, int n_f, const int32_t* f_slot, int size_x, int size_y, double res, double max_linf,
                        const float* m_cells, int n_m, const double* pose4, int k, int metric, int variant,
                        double matcher_loss_scale, double loop_scale, double alpha, double divisor, int max_gnc_steps,
                        int max_iterations, int on_manifold, double loss_weight, const uint32_t* im_in, const uint32_t* jf_in, int P_in,
                        double* out9) {
  NdtMap F = view_map(f_cells, nullptr, n_f, f_slot, size_x, size_y, res, max_linf);
  NdtMap M = view_map(m_cells, nullptr, n_m, nullptr, size_x, size_y, res, max_linf);
  LmOptions opt; opt.max_num_iterations = max_iterations;
  if (g_function_tol > 0.0) opt.function_tolerance = g_function_tol;
  if (g_parameter_tol > 0.0) opt.parameter_tolerance = g_parameter_tol;
  if (g_gradient_tol > 0.0) opt.gradient_tolerance = g_gradient_tol;
  PairList pl;
  if (im_in && P_in > 0) { pl.im.assign(im_in, im_in + P_in); pl.jf.assign(jf_in, jf_in + P_in); }
  LoopConstraintResult R = loop_constraint(F, M, pose4, k, metric, variant, matcher_loss_scale, loop_scale, alpha, divisor,
                                           max_gnc_steps, opt, on_manifold != 0, loss_weight, (im_in && P_in > 0) ? &pl : nullptr);
  for (int i = 0; i < 4; ++i) out9[i] = R.pose[i];
  out9[4] = R.score; out9[5] = R.gnc_solves; out9[6] = R.total_iterations; out9[7] = R.total_evals; out9[8] = R.gnc_mu_first;
  return R.status;
}

void orc_sweep_costs(int variant, const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf, size_t P,
                     int loss_kind, double a, double alpha, double mu, double weight, const double* poses, size_t S, double* cost_out) {
  sweep_costs(variant, reinterpret_cast<const Cell12*>(cells_m), reinterpret_cast<const Cell12*>(cells_f), im, jf, P,
              make_loss(loss_kind, a, alpha, mu, weight), poses, S, cost_out);
}

void orc_se2_plus(const double* T4, const double* d3, double* out4) { se2_plus(T4, d3, out4); }
void orc_se2_plus_jacobian(const double* T4, double* J12) { se2_plus_jacobian(T4, J12); }

int orc_bnb(const float* f_cells, int n_f, const int32_t* f_slot, int size_x, int size_y, double res, double max_linf, const float* m_cells, int n_m,
            const double* pose4, int variant, double alpha, double scale, double window_linear, double window_angular, double linear_step,
            double max_px_range, double cost_threshold, int n_iter, double* out6) {
  NdtMap F = view_map(f_cells, nullptr, n_f, f_slot, size_x, size_y, res, max_linf);
  NdtMap M = view_map(m_cells, nullptr, n_m, nullptr, size_x, size_y, res, max_linf);
  BnbResult R = bnb_search(F, M, pose4, variant, alpha, scale, window_linear, window_angular, linear_step, max_px_range, cost_threshold, n_iter);
  for (int i = 0; i < 4; ++i) out6[i] = R.pose[i];
  out6[4] = R.min_cost; out6[5] = R.n_evaluated;
  return 0;
}

double orc_cs_divergence(const float* f_cells, size_t n_f, const float* m_cells, size_t n_m, double* terms3) {
  return cs_divergence(reinterpret_cast<const Cell12*>(f_cells), n_f, reinterpret_cast<const Cell12*>(m_cells), n_m, terms3);
}

int orc_filter_scan(const float* raw4, size_t n, float min_d, float max_d, float min_i, double beam_thr, const float* tf12, float* out4, size_t cap,
                    size_t* n_peaks) {
  FilterParams fp; fp.min_distance = min_d; fp.max_distance = max_d; fp.min_intensity = min_i; fp.beam_thr = beam_thr;
  for (int i = 0; i < 12; ++i) fp.tf[i] = tf12[i];
  std::vector<Pt4> out; std::vector<size_t> peaks;
  filter_scan(reinterpret_cast<const Pt4*>(raw4), n, fp, out, &peaks);
  if (n_peaks) *n_peaks = peaks.size();
  if (out.size() > cap) return -1;
  std::memcpy(out4, out.data(), out.size() * sizeof(Pt4));
  return (int)out.size();
}

// ---- estimateTransformCeres window problem (window_oracle.h) ----------------------------------------------------------
// params [16 + 64]: k, gnc_steps, max_iteration, loss_scale, alpha, divisor, ndt_weight, manifold, constant_velocity, use_imu, weight_imu,
// weight_imu_bias, reject_translation, reject_rotation, variant, (reserved), then covariance_scaling_factor * motion_sqrtI row-major.
// states [(W + 1)][14]: cos, sin, tx, ty, pos_x, pos_y, rot, vx, vy, omega, ax, ay, imu_bias, stamp — oldest (constant) state first.
static void window_fill(WindowProblem& wp, const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf, const uint32_t* seg_off,
                        int W, const double* states, const double* imu, const double* params) {
  WindowParams& P = wp.P;
  P.k = (int)params[0]; P.gnc_steps = (int)params[1]; P.max_iteration = (int)params[2]; P.loss_scale = params[3]; P.alpha = params[4];
  P.divisor = params[5]; P.ndt_weight = params[6]; P.manifold = params[7] != 0.0; P.constant_velocity = params[8] != 0.0;
  P.use_imu = params[9] != 0.0; P.weight_imu = params[10]; P.weight_imu_bias = params[11]; P.reject_translation = params[12];
  P.reject_rotation = params[13]; P.variant = (int)params[14];
  for (int i = 0; i < 64; ++i) P.sqrtI[i] = params[16 + i];
  wp.st.resize(W + 1);
  for (int j = 0; j <= W; ++j) {
    const double* s = states + (size_t)j * WSTATE_DOUBLES; WState& o = wp.st[j];
    for (int i = 0; i < 4; ++i) o.pose[i] = s[i];
    o.pos[0] = s[4]; o.pos[1] = s[5]; o.rot = s[6]; o.lin_vel[0] = s[7]; o.lin_vel[1] = s[8]; o.rot_vel = s[9];
    o.lin_acc[0] = s[10]; o.lin_acc[1] = s[11]; o.imu_bias = s[12]; o.stamp = s[13];
  }
  wp.imu.assign(W, 0.0);
  if (imu) for (int j = 0; j < W; ++j) wp.imu[j] = imu[j];
  wp.cm = reinterpret_cast<const Cell12*>(cells_m); wp.cf = reinterpret_cast<const Cell12*>(cells_f);
  wp.im = im; wp.jf = jf;
  wp.seg_off.assign(seg_off, seg_off + W + 1);
  wp.layout();
}
static void window_store(const WindowProblem& wp, double* states) {
  for (size_t j = 0; j < wp.st.size(); ++j) {
    double* s = states + j * WSTATE_DOUBLES; const WState& o = wp.st[j];
    for (int i = 0; i < 4; ++i) s[i] = o.pose[i];
    s[4] = o.pos[0]; s[5] = o.pos[1]; s[6] = o.rot; s[7] = o.lin_vel[0]; s[8] = o.lin_vel[1]; s[9] = o.rot_vel;
    s[10] = o.lin_acc[0]; s[11] = o.lin_acc[1]; s[12] = o.imu_bias; s[13] = o.stamp;
  }
}
// one evaluation of the whole problem at the given states: cost, tangent gradient [nt], tangent J^T J [nt * nt]; returns nt
int orc_window_evaluate(const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf, const uint32_t* seg_off, int W,
                        const double* states, const double* imu, const double* params, int loss_kind, double mu, double weight, double* cost,
                        double* g, double* H, double* max_raw) {
  WindowProblem wp;
  window_fill(wp, cells_m, cells_f, im, jf, seg_off, W, states, imu, params);
  std::vector<double> x(wp.n_amb());
  wp.pack(x.data());
  Loss loss = make_loss(loss_kind, wp.P.loss_scale, wp.P.alpha, mu, weight);
  wp.evaluate(x.data(), loss, cost, g, H, max_raw);
  return wp.n_tan();
}
// out8: status, rejected, gnc_solves, total_iterations, final_cost, mu_first, max_residual, n_tangent
int orc_window_solve(const float* cells_m, const float* cells_f, const uint32_t* im, const uint32_t* jf, const uint32_t* seg_off, int W,
                     double* states, const double* imu, const double* params, size_t n_cells_total, double* trans4, double* out8) {
  WindowProblem wp;
  window_fill(wp, cells_m, cells_f, im, jf, seg_off, W, states, imu, params);
  LmOptions opt; opt.max_num_iterations = wp.P.max_iteration;
  if (g_function_tol > 0.0) opt.function_tolerance = g_function_tol;
  if (g_parameter_tol > 0.0) opt.parameter_tolerance = g_parameter_tol;
  if (g_gradient_tol > 0.0) opt.gradient_tolerance = g_gradient_tol;
  const WindowResult R = window_solve(wp, trans4, opt, n_cells_total);
  window_store(wp, states);
  out8[0] = R.status; out8[1] = R.rejected; out8[2] = R.gnc_solves; out8[3] = R.total_iterations; out8[4] = R.final_cost; out8[5] = R.mu_first;
  out8[6] = R.max_residual; out8[7] = R.n_tangent;
  return R.status;
}

// One host factor of the window problem as ceres sees it: residuals and the AMBIENT Jacobian over the 2 x 10 parameter slots of its two
// states (slot of state side s: 10 s + pose 0-3 | pos 0-1, rot 2 | lin_vel 4-5 | rot_vel 6 | lin_acc 7-8 | imu_bias 9), every parameter
// free.  kind 0: MotionModelFactorSE2 / MotionModelFactor (8 residuals), kind 1: RotationalResidualSE2 / RotationalResidual (2).
// states: a14, b14 as in orc_window_solve.  Returns the number of residuals; jac is [nres][20] row-major.
int orc_factor_block(int kind, int manifold, const double* a14, const double* b14, const double* sqrtI64, double imu_rot, double weight_imu,
                     double weight_bias, double* residuals, double* jac) {
  typedef Jet<20> J;
  auto load = [&](const double* s, int base, StateT<J>& S, J& bias) {
    for (int i = 0; i < 4; ++i) S.pose[i] = manifold ? J(s[i], base + i) : J(s[i]);
    S.pos[0] = manifold ? J(s[4]) : J(s[4], base + 0); S.pos[1] = manifold ? J(s[5]) : J(s[5], base + 1); S.rot = manifold ? J(s[6]) : J(s[6], base + 2);
    S.vel[0] = J(s[7], base + 4); S.vel[1] = J(s[8], base + 5); S.omega = J(s[9], base + 6);
    S.acc[0] = J(s[10], base + 7); S.acc[1] = J(s[11], base + 8);
    bias = J(s[12], base + 9);
  };
  StateT<J> A, B; J ba, bb;
  load(a14, 0, A, ba); load(b14, 10, B, bb);
  const double dt = b14[13] - a14[13];
  J res[8];
  int n = 8;
  if (kind == 0) { if (manifold) motion_factor_se2(A, B, dt, sqrtI64, res); else motion_factor_vec(A, B, dt, sqrtI64, res); }
  else { n = 2; imu_factor(manifold != 0, A, B, ba, bb, imu_rot, weight_imu, dt, weight_bias, res); }
  for (int r = 0; r < n; ++r) { residuals[r] = res[r].a; for (int c = 0; c < 20; ++c) jac[r * 20 + c] = res[r].v[c]; }
  return n;
}

// lm_minimize on the host factors of the window problem alone (no NDT blocks): states in / out; summary4 = initial cost, final cost,
// iterations, termination
int orc_window_minimize_factors(double* states, int W, const double* imu, const double* params, int max_iterations, double* summary4) {
  WindowProblem wp;
  std::vector<uint32_t> seg_off((size_t)W + 1, 0u);
  const float no_cell[12] = {0};
  const uint32_t no_pair = 0;
  window_fill(wp, no_cell, no_cell, &no_pair, &no_pair, seg_off.data(), W, states, imu, params);
  LmOptions opt; opt.max_num_iterations = max_iterations > 0 ? max_iterations : 200;
  if (g_function_tol > 0.0) opt.function_tolerance = g_function_tol;
  if (g_parameter_tol > 0.0) opt.parameter_tolerance = g_parameter_tol;
  if (g_gradient_tol > 0.0) opt.gradient_tolerance = g_gradient_tol;
  std::vector<double> x(wp.n_amb());
  wp.pack(x.data());
  Loss none; none.kind = LOSS_NONE;
  NormalEqProblem prob;
  prob.n_amb = wp.n_amb(); prob.n_tan = wp.n_tan();
  prob.eval_full = [&](const double* xx, double* cost, double* g, double* H) { return wp.evaluate(xx, none, cost, g, H); };
  prob.eval_cost = [&](const double* xx, double* cost) { return wp.evaluate(xx, none, cost, nullptr, nullptr); };
  prob.plus = [&](const double* xx, const double* d, double* xp) { wp.plus(xx, d, xp); };
  const LmSummary S = lm_minimize(prob, opt, x.data());
  wp.unpack(x.data());
  window_store(wp, states);
  summary4[0] = S.initial_cost; summary4[1] = S.final_cost; summary4[2] = S.num_iterations; summary4[3] = S.termination;
  return 0;
}

}  // extern "C"
