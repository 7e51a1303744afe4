// Joint window problem of Matcher::estimateTransformCeres: see window_solver.hpp.
#include "window_solver.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace randt {
namespace window {

// ---------------------------------------------------------------------------------------------------------------------
// dual-number arithmetic (the operator set the factors below need)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; const double bi = 1.0 / b.v, q = a.v * bi;
  r.v = q; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - q * b.d[i]) * bi; return r;
}
template <int N> inline Dual<N> operator+(const Dual<N>& a, double s) { Dual<N> r = a; r.v += s; return r; }
template <int N> inline Dual<N> operator+(double s, const Dual<N>& a) { Dual<N> r = a; r.v += s; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double s) { Dual<N> r = a; r.v -= s; return r; }
template <int N> inline Dual<N> operator-(double s, const Dual<N>& a) { Dual<N> r; r.v = s - a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double s) { Dual<N> r; r.v = a.v * s; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> inline Dual<N> operator*(double s, const Dual<N>& a) { return a * s; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double s) { return a * (1.0 / s); }
template <int N> inline Dual<N> dsin(const Dual<N>& a) { Dual<N> r; const double c = std::cos(a.v); r.v = std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> dcos(const Dual<N>& a) { Dual<N> r; const double s = -std::sin(a.v); r.v = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <int N> inline Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r; const double t = 1.0 / (x.v * x.v + y.v * y.v);
  r.v = std::atan2(y.v, x.v); for (int i = 0; i < N; ++i) r.d[i] = t * (x.v * y.d[i] - y.v * x.d[i]); return r;
}
template <int N> inline Dual<N> dfloor(const Dual<N>& a) { return Dual<N>(std::floor(a.v)); }   // piecewise constant: no partials (ceres::floor)
inline double dsin(double a) { return std::sin(a); }
inline double dcos(double a) { return std::cos(a); }
inline double datan2(double y, double x) { return std::atan2(y, x); }
inline double dfloor(double a) { return std::floor(a); }
template <int N> inline double val(const Dual<N>& a) { return a.v; }
inline double val(double a) { return a; }

// NormalizeAngle (R/include/ndt_registration/state_manifold.h:17-23)
template <typename T> inline T NormalizeAngle(const T& a) { const T two_pi(2.0 * M_PI); return a - two_pi * dfloor((a + T(M_PI)) / two_pi); }

// Sophus::SE2<T> (1.22.10): unit complex number + translation; the branches look at values only, as Jet comparisons do
template <typename T> struct Group { T re, im, tx, ty; };
template <typename T> inline Group<T> mapPose(const T* p) { return Group<T>{p[0], p[1], p[2], p[3]}; }
template <typename T> inline Group<T> product(const Group<T>& a, const Group<T>& b) {
  Group<T> r;
  r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re;
  const T sq = r.re * r.re + r.im * r.im;
  if (val(sq) != 1.0) { const T scale = T(2.0) / (T(1.0) + sq); r.re = r.re * scale; r.im = r.im * scale; }   // SO2 product's first-order renormalisation
  r.tx = a.tx + (a.re * b.tx - a.im * b.ty);
  r.ty = a.ty + (a.im * b.tx + a.re * b.ty);
  return r;
}
template <typename T> inline Group<T> inverse(const Group<T>& a) {
  Group<T> r;
  r.re = a.re; r.im = -a.im;
  const T mx = a.tx * T(-1.0), my = a.ty * T(-1.0);
  r.tx = r.re * mx - r.im * my;
  r.ty = r.im * mx + r.re * my;
  return r;
}
template <typename T> inline Group<T> expMap(const T& ux, const T& uy, const T& theta) {
  Group<T> r;
  r.re = dcos(theta); r.im = dsin(theta);
  T sin_by_theta, one_minus_cos_by_theta;
  if (std::fabs(val(theta)) < 1e-10) {   // Sophus::Constants<double>::epsilon()
    const T theta_sq = theta * theta;
    sin_by_theta = T(1.0) - T(1.0 / 6.0) * theta_sq;
    one_minus_cos_by_theta = T(0.5) * theta - T(1.0 / 24.0) * theta * theta_sq;
  } else {
    sin_by_theta = r.im / theta;
    one_minus_cos_by_theta = (T(1.0) - r.re) / theta;
  }
  r.tx = sin_by_theta * ux - one_minus_cos_by_theta * uy;
  r.ty = one_minus_cos_by_theta * ux + sin_by_theta * uy;
  return r;
}
template <typename T> inline void logMap(const Group<T>& a, T out[3]) {
  const T theta = datan2(a.im, a.re);
  const T halftheta = T(0.5) * theta;
  const T real_minus_one = a.re - T(1.0);
  T halftheta_by_tan_of_halftheta;
  if (std::fabs(val(real_minus_one)) < 1e-10) halftheta_by_tan_of_halftheta = T(1.0) - T(1.0 / 12.0) * theta * theta;
  else halftheta_by_tan_of_halftheta = -(halftheta * a.im) / real_minus_one;
  out[0] = halftheta_by_tan_of_halftheta * a.tx + halftheta * a.ty;
  out[1] = -halftheta * a.tx + halftheta_by_tan_of_halftheta * a.ty;
  out[2] = theta;
}
// residuals_map.applyOnTheLeft(sqrtI_)
template <typename T> inline void applySqrtInformation(const double* sqrtI, const T e[8], T* residuals) {
  for (int i = 0; i < 8; ++i) {
    T acc = T(sqrtI[i * 8]) * e[0];
    for (int j = 1; j < 8; ++j) acc = acc + T(sqrtI[i * 8 + j]) * e[j];
    residuals[i] = acc;
  }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// the reference's factors
// ---------------------------------------------------------------------------------------------------------------------
// MotionModelFactorSE2::operator() (ceres_residuals.h:629-675) around predictSE2 (:67-85)
template <typename T>
void MotionModelFactorSE2(double raw_dt, const double* sqrtI, const T* old_pose, const T* old_lin_vel, const T* old_rot_vel, const T* old_lin_acc,
                          const T* new_pose, const T* new_lin_vel, const T* new_rot_vel, const T* new_lin_acc, T* residuals) {
  const double dt = std::max(raw_dt, 0.2);
  const Group<T> step = expMap(old_lin_vel[0] * dt + 0.5 * dt * old_lin_acc[0], old_lin_vel[1] * dt + 0.5 * dt * old_lin_acc[1], old_rot_vel[0] * dt);
  const Group<T> pose_pred = product(mapPose(old_pose), step);
  T error[3];
  logMap(product(inverse(pose_pred), mapPose(new_pose)), error);
  T e[8];
  e[0] = error[0]; e[1] = error[1]; e[2] = error[2];
  e[3] = new_lin_vel[0] - (old_lin_vel[0] + dt * old_lin_acc[0]);
  e[4] = new_lin_vel[1] - (old_lin_vel[1] + dt * old_lin_acc[1]);
  e[5] = new_rot_vel[0] - old_rot_vel[0];
  e[6] = new_lin_acc[0] - old_lin_acc[0];
  e[7] = new_lin_acc[1] - old_lin_acc[1];
  applySqrtInformation(sqrtI, e, residuals);
}
// MotionModelFactor::operator() (ceres_residuals.h:562-615) around predict (:25-57)
template <typename T>
void MotionModelFactor(double raw_dt, const double* sqrtI, const T* old_pos, const T* old_rot, const T* old_lin_vel, const T* old_rot_vel,
                       const T* old_lin_acc, const T* new_pos, const T* new_rot, const T* new_lin_vel, const T* new_rot_vel, const T* new_lin_acc,
                       T* residuals) {
  const double dt = std::max(raw_dt, 0.2);
  const T mid_rot = NormalizeAngle(old_rot[0] + 0.5 * dt * old_rot_vel[0]);
  const T rot_pred = NormalizeAngle(old_rot[0] + dt * old_rot_vel[0]);
  const T sy = dsin(mid_rot), cy = dcos(mid_rot);
  const T delta_x = old_lin_vel[0] * dt + 0.5 * old_lin_acc[0] * dt * dt;
  const T delta_y = old_lin_vel[1] * dt + 0.5 * old_lin_acc[1] * dt * dt;
  T e[8];
  e[0] = new_pos[0] - (old_pos[0] + (cy * delta_x - sy * delta_y));
  e[1] = new_pos[1] - (old_pos[1] + (sy * delta_x + cy * delta_y));
  e[2] = NormalizeAngle(new_rot[0] - rot_pred);
  e[3] = new_lin_vel[0] - (old_lin_vel[0] + dt * old_lin_acc[0]);
  e[4] = new_lin_vel[1] - (old_lin_vel[1] + dt * old_lin_acc[1]);
  e[5] = new_rot_vel[0] - old_rot_vel[0];
  e[6] = new_lin_acc[0] - old_lin_acc[0];
  e[7] = new_lin_acc[1] - old_lin_acc[1];
  applySqrtInformation(sqrtI, e, residuals);
}
// RotationalResidualSE2::operator() (ceres_residuals.h:355-363): the time step is NOT clamped here
template <typename T>
void RotationalResidualSE2(double rot, double weight, double dt, double bias_weight, const T* pose_old, const T* pose_new, const T* bias_old,
                           const T* bias_new, T* residuals) {
  const Group<T> M_0 = mapPose(pose_old);
  const Group<T> M_1 = product(mapPose(pose_new), expMap(T(0.0), T(0.0), bias_new[0] * dt));
  T lg[3];
  logMap(product(inverse(M_0), M_1), lg);
  residuals[0] = weight * (rot - lg[2]);
  residuals[1] = bias_weight * (bias_new[0] - bias_old[0]);
}
// RotationalResidual::operator() (ceres_residuals.h:325-329)
template <typename T>
void RotationalResidual(double rot, double weight, double dt, double bias_weight, const T* rot_old, const T* rot_new, const T* bias_old,
                        const T* bias_new, T* residuals) {
  residuals[0] = weight * (rot - NormalizeAngle(rot_new[0] - rot_old[0] + bias_new[0] * dt));
  residuals[1] = bias_weight * (bias_new[0] - bias_old[0]);
}
template void MotionModelFactorSE2<D20>(double, const double*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, D20*);
template void MotionModelFactor<D20>(double, const double*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, const D20*, D20*);
template void RotationalResidualSE2<D20>(double, double, double, double, const D20*, const D20*, const D20*, const D20*, D20*);
template void RotationalResidual<D20>(double, double, double, double, const D20*, const D20*, const D20*, const D20*, D20*);

// ---------------------------------------------------------------------------------------------------------------------
// JointProblem
// ---------------------------------------------------------------------------------------------------------------------
namespace {
// Sophus::Manifold<SE2>::PlusJacobian = Dx_this_mul_exp_x_at_0: rows (0, 0, -s), (0, 0, c), (c, -s, 0), (s, c, 0)
inline void plusJacobianSE2(const double* T, double P[4][3]) {
  const double c = T[0], s = T[1];
  const double v[4][3] = {{0, 0, -s}, {0, 0, c}, {c, -s, 0}, {s, c, 0}};
  std::memcpy(P, v, sizeof(v));
}
// Sophus::Manifold<SE2>::Plus: T * exp(delta)
inline void plusSE2(const double* T, const double* delta, double* out) {
  const Group<double> r = product(mapPose(T), expMap(delta[0], delta[1], delta[2]));
  out[0] = r.re; out[1] = r.im; out[2] = r.tx; out[3] = r.ty;
}
}  // namespace

int JointProblem::addParameterBlock(double* values, int size, bool se2_manifold) {
  for (size_t i = 0; i < blocks_.size(); ++i) if (blocks_[i].user == values) return (int)i;
  ParameterBlock b;
  b.user = values; b.size = size; b.se2 = se2_manifold; b.tangent = se2_manifold ? 3 : size;
  blocks_.push_back(b);
  return (int)blocks_.size() - 1;
}

void JointProblem::finalize() {
  n_amb_ = n_tan_ = 0;
  for (ParameterBlock& b : blocks_) {
    if (b.constant) { b.x_off = b.t_off = -1; continue; }
    b.x_off = n_amb_; b.t_off = n_tan_;
    n_amb_ += b.size; n_tan_ += b.tangent;
  }
  if (ndt_.problem) {
    uint32_t S = 0, P = 0;
    randt_problem_info(ndt_.problem, &S, &P, nullptr, nullptr);
    ndt_blocks_ = P;
    if (!ndt_.poses || !ndt_.records) throw Error(RANDT_E_INVALID, "window problem: the NDT term needs staging buffers");
    (void)S;
  }
}

void JointProblem::gather(double* x) const {
  for (const ParameterBlock& b : blocks_) if (!b.constant) std::memcpy(x + b.x_off, b.user, sizeof(double) * b.size);
}
void JointProblem::scatter(const double* x) const {
  for (const ParameterBlock& b : blocks_) if (!b.constant) std::memcpy(b.user, x + b.x_off, sizeof(double) * b.size);
}
void JointProblem::plus(const double* x, const double* delta, double* x_plus) const {
  for (const ParameterBlock& b : blocks_) {
    if (b.constant) continue;
    if (b.se2) plusSE2(x + b.x_off, delta + b.t_off, x_plus + b.x_off);
    else for (int i = 0; i < b.size; ++i) x_plus[b.x_off + i] = x[b.x_off + i] + delta[b.t_off + i];
  }
}

void JointProblem::evaluateHostFactors(const double* x, bool want_jac, double* cost, double* g, double* H) const {
  const int nt = n_tan_;
  D20 storage[kSeeds];
  const D20* params[10];
  D20 residuals[8];
  int cols[kSeeds];
  double row[kSeeds];
  for (const HostFactor& f : factors_) {
    int slot = 0;
    for (size_t bi = 0; bi < f.blocks.size(); ++bi) {
      const ParameterBlock& b = blocks_[f.blocks[bi]];
      const double* v = blockValues(f.blocks[bi], x);
      params[bi] = &storage[slot];
      for (int c = 0; c < b.size; ++c) {
        storage[slot + c] = D20(v[c]);
        if (!b.constant) storage[slot + c].d[slot + c] = 1.0;
      }
      slot += b.size;
    }
    f.eval(params, residuals);
    for (int r = 0; r < f.nres; ++r) {
      const double rv = residuals[r].v;
      *cost += 0.5 * rv * rv;
      if (!want_jac) continue;
      // local Jacobian row: ambient partials through the manifold's PlusJacobian, gathered as (column, value)
      int n = 0, s = 0;
      for (size_t bi = 0; bi < f.blocks.size(); ++bi) {
        const ParameterBlock& b = blocks_[f.blocks[bi]];
        if (!b.constant) {
          const double* dv = residuals[r].d + s;
          if (b.se2) {
            double P[4][3];
            plusJacobianSE2(blockValues(f.blocks[bi], x), P);
            for (int a = 0; a < 3; ++a) { cols[n] = b.t_off + a; row[n] = dv[0] * P[0][a] + dv[1] * P[1][a] + dv[2] * P[2][a] + dv[3] * P[3][a]; ++n; }
          } else {
            for (int a = 0; a < b.size; ++a) { cols[n] = b.t_off + a; row[n] = dv[a]; ++n; }
          }
        }
        s += b.size;
      }
      for (int a = 0; a < n; ++a) {
        g[cols[a]] += row[a] * rv;
        for (int c = 0; c < n; ++c) H[(size_t)cols[a] * nt + cols[c]] += row[a] * row[c];
      }
    }
  }
}

bool JointProblem::evaluate(const double* x, const randt_loss* loss, bool want_jac, double* cost, double* g, double* H, double* max_raw) {
  const int nt = n_tan_;
  *cost = 0.0;
  if (want_jac) { std::fill(g, g + nt, 0.0); std::fill(H, H + (size_t)nt * nt, 0.0); }
  ++n_evals_;
  if (ndt_.problem) {
    const size_t S = ndt_.seg_blocks.size();
    const int np = ndt_.np;
    for (size_t s = 0; s < S; ++s) {
      int o = 0;
      for (int id : ndt_.seg_blocks[s]) { const double* v = blockValues(id, x); for (int c = 0; c < blocks_[id].size; ++c) ndt_.poses[s * np + o++] = v[c]; }
    }
    randt_loss none;
    none.kind = RANDT_LOSS_NONE; none.scale = 1.0; none.alpha = 2.0; none.mu = 1.0; none.weight = 1.0;
    // ONE K3 launch: per window state the loss-corrected normal equations of all its residual blocks
    if (randt_eval_fused(ndt_.ctx, ndt_.problem, ndt_.variant, ndt_.poses, loss ? loss : &none, nullptr, want_jac ? 1 : 0, ndt_.records) != RANDT_OK) {
      device_failed_ = true;
      return false;
    }
    double mr = 0.0;
    for (size_t s = 0; s < S; ++s) {
      const double* rec = &ndt_.records[s * RANDT_FUSED_STRIDE];
      *cost += rec[RANDT_FUSED_COST];
      if (rec[RANDT_FUSED_N] > 0.0) mr = std::max(mr, rec[RANDT_FUSED_MAXR]);
      if (!want_jac) continue;
      // ambient index of the variant's parameter vector -> tangent columns
      double M[4][4] = {{0}};
      int cols[4], n = 0, a0 = 0;
      for (int id : ndt_.seg_blocks[s]) {
        const ParameterBlock& b = blocks_[id];
        if (!b.constant) {
          if (b.se2) {
            double P[4][3];
            plusJacobianSE2(blockValues(id, x), P);
            for (int a = 0; a < 3; ++a) { cols[n + a] = b.t_off + a; for (int i = 0; i < 4; ++i) M[a0 + i][n + a] = P[i][a]; }
            n += 3;
          } else {
            for (int a = 0; a < b.size; ++a) { cols[n + a] = b.t_off + a; M[a0 + a][n + a] = 1.0; }
            n += b.size;
          }
        }
        a0 += b.size;
      }
      for (int a = 0; a < n; ++a) {
        double ga = 0.0;
        for (int i = 0; i < np; ++i) ga += M[i][a] * rec[RANDT_FUSED_G + i];
        g[cols[a]] += ga;
        for (int c = 0; c < n; ++c) {
          double h = 0.0;
          for (int i = 0; i < np; ++i) for (int j = 0; j < np; ++j) h += M[i][a] * rec[RANDT_FUSED_H + i * 4 + j] * M[j][c];
          H[(size_t)cols[a] * nt + cols[c]] += h;
        }
      }
    }
    if (max_raw) *max_raw = mr;
  }
  evaluateHostFactors(x, want_jac, cost, g, H);
  return std::isfinite(*cost);
}

// ---------------------------------------------------------------------------------------------------------------------
// ceres 2.1.0 TrustRegionMinimizer + LevenbergMarquardtStrategy on the normal equations (the reference selects DENSE_QR on [J; sqrt(D)];
// the NDT blocks arrive as J^T J from the device, so the mathematically identical damped normal equations are solved by Cholesky)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
bool choleskySolve(std::vector<double>& A, const std::vector<double>& b, int n, std::vector<double>& y) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double l = std::sqrt(d);
    A[(size_t)j * n + j] = l;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / l;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * y[k]; y[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * y[k]; y[i] = s / A[(size_t)i * n + i]; }
  for (int i = 0; i < n; ++i) if (!std::isfinite(y[i])) return false;
  return true;
}
double norm2(const std::vector<double>& v) { double s = 0; for (double e : v) s += e * e; return std::sqrt(s); }
}  // namespace

MinimizerSummary minimize(JointProblem& problem, const randt_loss& loss, const randt_solver_options& o, double* x_io) {
  MinimizerSummary summary;
  const int na = problem.numAmbient(), nt = problem.numTangent();
  std::vector<double> x(x_io, x_io + na), candidate(na), probe(na);
  std::vector<double> g(nt), H((size_t)nt * nt), g_cand(nt), H_cand((size_t)nt * nt), gs(nt), Hs((size_t)nt * nt), A, scale(nt, 1.0), lm_diag(nt), y(nt), delta(nt), neg_g(nt);
  double cost = 0.0;
  if (!problem.evaluate(x.data(), &loss, true, &cost, g.data(), H.data())) { summary.termination = 2; return summary; }
  summary.initial_cost = cost;
  double min_cost = cost;
  summary.num_iterations = 1;
  if (o.jacobi_scaling) for (int i = 0; i < nt; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(H[(size_t)i * nt + i]));
  auto scaled_system = [&]() {
    for (int i = 0; i < nt; ++i) { gs[i] = g[i] * scale[i]; for (int j = 0; j < nt; ++j) Hs[(size_t)i * nt + j] = H[(size_t)i * nt + j] * scale[i] * scale[j]; }
  };
  // the gradient test: max |x - Plus(x, -g)| over the ambient parameters
  auto gradient_max_norm = [&]() {
    for (int i = 0; i < nt; ++i) neg_g[i] = -g[i];
    problem.plus(x.data(), neg_g.data(), probe.data());
    double m = 0.0;
    for (int i = 0; i < na; ++i) m = std::max(m, std::fabs(x[i] - probe[i]));
    return m;
  };
  scaled_system();
  double x_norm = norm2(x), radius = o.initial_trust_region_radius, decrease_factor = 2.0, gmax = gradient_max_norm();
  bool reuse_diagonal = false, step_was_successful = true;
  int invalid_steps = 0, iteration = 0;
  for (;;) {
    if (iteration >= o.max_num_iterations) { summary.termination = 1; break; }
    if (step_was_successful && gmax <= o.gradient_tolerance) { summary.termination = 0; break; }
    if (radius <= o.min_trust_region_radius) { summary.termination = 0; break; }
    ++iteration;
    if (!reuse_diagonal) for (int i = 0; i < nt; ++i) lm_diag[i] = std::min(std::max(Hs[(size_t)i * nt + i], o.min_lm_diagonal), o.max_lm_diagonal);
    A = Hs;
    for (int i = 0; i < nt; ++i) A[(size_t)i * nt + i] += lm_diag[i] / radius;
    bool usable = choleskySolve(A, gs, nt, y);
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    if (usable) {
      // step = -y; model_cost_change = -step^T (gs + Hs step / 2)
      double sg = 0.0, sHs = 0.0;
      for (int i = 0; i < nt; ++i) {
        sg -= y[i] * gs[i];
        double r = 0.0;
        for (int j = 0; j < nt; ++j) r -= Hs[(size_t)i * nt + j] * y[j];
        sHs -= y[i] * r;
      }
      model_cost_change = -(sg + 0.5 * sHs);
      usable = model_cost_change > 0.0;
    }
    if (!usable) {
      summary.num_iterations++;
      step_was_successful = false;
      if (++invalid_steps >= o.max_num_consecutive_invalid_steps) { summary.termination = 2; break; }
      radius *= 0.5;
      continue;
    }
    invalid_steps = 0;
    for (int i = 0; i < nt; ++i) delta[i] = -y[i] * scale[i];
    problem.plus(x.data(), delta.data(), candidate.data());
    // the candidate is evaluated with its Jacobian in the same device launch: an accepted step needs no second evaluation
    double cand_cost = 0.0;
    if (!problem.evaluate(candidate.data(), &loss, true, &cand_cost, g_cand.data(), H_cand.data()) || !std::isfinite(cand_cost)) cand_cost = DBL_MAX;
    if (problem.deviceFailed()) { summary.termination = 2; break; }     // not a bad step: the device call itself failed
    double step_norm = 0.0;
    for (int i = 0; i < na; ++i) step_norm += (x[i] - candidate[i]) * (x[i] - candidate[i]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { summary.termination = 0; break; }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= o.function_tolerance * cost) { summary.termination = 0; break; }
    const double relative_decrease = cost_change / model_cost_change;
    summary.num_iterations++;
    if (relative_decrease > o.min_relative_decrease) {
      x = candidate; x_norm = norm2(x);
      cost = cand_cost; g.swap(g_cand); H.swap(H_cand);
      scaled_system();
      gmax = gradient_max_norm();
      const double t = 2.0 * relative_decrease - 1.0;
      radius = std::min(o.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - t * t * t));
      decrease_factor = 2.0; reuse_diagonal = false; step_was_successful = true;
      min_cost = std::min(min_cost, cost);
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true; step_was_successful = false;
      min_cost = std::min(min_cost, cand_cost);
    }
  }
  summary.final_cost = min_cost;
  std::memcpy(x_io, x.data(), sizeof(double) * na);
  return summary;
}

}  // namespace window
}  // namespace randt
