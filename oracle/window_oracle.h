// ORACLE — TEST INFRASTRUCTURE ONLY (see ndt_oracle.h header).  PARITY UNPINNED.
//
// CPU restatement of the joint window problem Matcher::estimateTransformCeres builds and solves
// (R/src/ndt_registration/ndt_matcher.cpp:322-424): the NDT residual blocks of every free window state (addNDTFactor, :183-288), the
// motion-model factor between consecutive states (addMotionModelFactor :60-110 -> MotionModelFactor / MotionModelFactorSE2,
// R/include/ndt_registration/ceres_residuals.h:554-679, predict / predictSE2 :25-85), the relative IMU yaw factor (addImuFactor
// :144-181 -> RotationalResidual / RotationalResidualSE2, ceres_residuals.h:307-370), the parameter blocks of
// addMotionParameterBlock / addImuParameterBlock (:290-320: the POSE and the IMU bias of the oldest window state are constant — its
// velocities are not —, the acceleration blocks are constant under the constant-velocity model), the GNC loop (:382-397) and the rejection gate (:408-422).  Every factor is differentiated with
// dual numbers (what ceres::AutoDiffCostFunction does), pose blocks go through Sophus::Manifold<SE2>::PlusJacobian, and the
// trust-region loop is lm_oracle.h's restatement of ceres 2.1.0.  Sophus 1.22.10's SE2 exp / log / inverse / product are restated
// on the generic scalar (branches compare the value part of a dual number, as ceres' Jet comparisons do).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "lm_oracle.h"

namespace orc {

// rc::navigation::ndt::State, flat (R/include/ndt_slam/trajectory_representation.h:12-22)
struct WState {
  double pose[4];      // Sophus::SE2d::data(): cos, sin, tx, ty
  double pos[2], rot;
  double lin_vel[2], rot_vel, lin_acc[2], imu_bias, stamp;
};
constexpr int WSTATE_DOUBLES = 14;

struct WindowParams {
  int k = 2, gnc_steps = 2, max_iteration = 200;
  double loss_scale = 1.0, alpha = -2.0, divisor = 1.1, ndt_weight = 5000.0;
  bool manifold = true, constant_velocity = true, use_imu = false;
  double weight_imu = 64.0, weight_imu_bias = 750000.1;
  double sqrtI[64];    // covariance_scaling_factor * motion_sqrtI, entry (i, j) at [i * 8 + j]
  double reject_translation = 5.0, reject_rotation = 2.0;
  int variant = VAR_SE2_INTENSITY;
};

struct WindowResult {
  int status = 0;          // 0 solved, 1 nothing to solve (fewer than two states, or no NDT residual block)
  int rejected = 0;
  int gnc_solves = 0, total_iterations = 0, n_free = 0, n_tangent = 0;
  double final_cost = 0, mu_first = 0, max_residual = 0;
};

// ---- Sophus SE2<T> on a generic scalar -------------------------------------------------------------------------------
template <typename T> struct Se2 { T c, s, x, y; };

template <typename T>
inline Se2<T> se2t_mul(const Se2<T>& A, const Se2<T>& B) {
  T re = A.c * B.c - A.s * B.s, im = A.c * B.s + A.s * B.c;
  const T n2 = re * re + im * im;
  if (value_of(n2) != 1.0) { const T sc = T(2.0) / (T(1.0) + n2); re = re * sc; im = im * sc; }
  Se2<T> R;
  R.c = re; R.s = im;
  R.x = A.x + (A.c * B.x - A.s * B.y);
  R.y = A.y + (A.s * B.x + A.c * B.y);
  return R;
}
template <typename T>
inline Se2<T> se2t_inverse(const Se2<T>& A) {
  Se2<T> R;
  R.c = A.c; R.s = -A.s;
  const T nx = A.x * T(-1.0), ny = A.y * T(-1.0);
  R.x = R.c * nx - R.s * ny;
  R.y = R.s * nx + R.c * ny;
  return R;
}
template <typename T>
inline Se2<T> se2t_exp(const T& ux, const T& uy, const T& theta) {
  Se2<T> R;
  R.c = cos(theta); R.s = sin(theta);
  T sbt, omcbt;
  if (std::fabs(value_of(theta)) < 1e-10) {
    const T t2 = theta * theta;
    sbt = T(1.0) - T(1.0 / 6.0) * t2;
    omcbt = T(0.5) * theta - T(1.0 / 24.0) * theta * t2;
  } else { sbt = R.s / theta; omcbt = (T(1.0) - R.c) / theta; }
  R.x = sbt * ux - omcbt * uy;
  R.y = omcbt * ux + sbt * uy;
  return R;
}
template <typename T>
inline void se2t_log(const Se2<T>& A, T out[3]) {
  const T theta = atan2(A.s, A.c);
  const T half = T(0.5) * theta;
  const T real_minus_one = A.c - T(1.0);
  T h;
  if (std::fabs(value_of(real_minus_one)) < 1e-10) h = T(1.0) - T(1.0 / 12.0) * theta * theta;
  else h = -(half * A.s) / real_minus_one;
  out[0] = h * A.x + half * A.y;
  out[1] = -half * A.x + h * A.y;
  out[2] = theta;
}

// ---- the factors, templated like the reference functors ----------------------------------------------------------------
// one state's parameters as the factor sees them: pose (manifold) or pos, rot (vector), lin_vel, rot_vel, lin_acc
template <typename T> struct StateT { T pose[4]; T pos[2]; T rot; T vel[2]; T omega; T acc[2]; };

// MotionModelFactorSE2::operator()  ceres_residuals.h:629-675 with predictSE2 :67-85
template <typename T>
inline void motion_factor_se2(const StateT<T>& a, const StateT<T>& b, double raw_dt, const double* sqrtI, T res[8]) {
  const double dt = std::max(raw_dt, 0.2);
  const Se2<T> p0{a.pose[0], a.pose[1], a.pose[2], a.pose[3]}, p1{b.pose[0], b.pose[1], b.pose[2], b.pose[3]};
  const Se2<T> pred = se2t_mul(p0, se2t_exp(a.vel[0] * dt + 0.5 * dt * a.acc[0], a.vel[1] * dt + 0.5 * dt * a.acc[1], a.omega * dt));
  T lg[3];
  se2t_log(se2t_mul(se2t_inverse(pred), p1), lg);
  T e[8];
  e[0] = lg[0]; e[1] = lg[1]; e[2] = lg[2];
  e[3] = b.vel[0] - (a.vel[0] + dt * a.acc[0]);
  e[4] = b.vel[1] - (a.vel[1] + dt * a.acc[1]);
  e[5] = b.omega - a.omega;
  e[6] = b.acc[0] - a.acc[0];
  e[7] = b.acc[1] - a.acc[1];
  for (int i = 0; i < 8; ++i) { T s = T(sqrtI[i * 8]) * e[0]; for (int j = 1; j < 8; ++j) s = s + T(sqrtI[i * 8 + j]) * e[j]; res[i] = s; }
}
// MotionModelFactor::operator()  ceres_residuals.h:562-615 with predict :25-57
template <typename T>
inline void motion_factor_vec(const StateT<T>& a, const StateT<T>& b, double raw_dt, const double* sqrtI, T res[8]) {
  const double dt = std::max(raw_dt, 0.2);
  const T rot = normalize_angle(a.rot + 0.5 * dt * a.omega);
  const T new_rot = normalize_angle(a.rot + dt * a.omega);
  const T sy = sin(rot), cy = cos(rot);
  const T dx = a.vel[0] * dt + 0.5 * a.acc[0] * dt * dt, dy = a.vel[1] * dt + 0.5 * a.acc[1] * dt * dt;
  const T px = a.pos[0] + (cy * dx - sy * dy), py = a.pos[1] + (sy * dx + cy * dy);
  T e[8];
  e[0] = b.pos[0] - px; e[1] = b.pos[1] - py;
  e[2] = normalize_angle(b.rot - new_rot);
  e[3] = b.vel[0] - (a.vel[0] + dt * a.acc[0]);
  e[4] = b.vel[1] - (a.vel[1] + dt * a.acc[1]);
  e[5] = b.omega - a.omega;
  e[6] = b.acc[0] - a.acc[0];
  e[7] = b.acc[1] - a.acc[1];
  for (int i = 0; i < 8; ++i) { T s = T(sqrtI[i * 8]) * e[0]; for (int j = 1; j < 8; ++j) s = s + T(sqrtI[i * 8 + j]) * e[j]; res[i] = s; }
}
// RotationalResidualSE2 (ceres_residuals.h:355-363) / RotationalResidual (:325-329)
template <typename T>
inline void imu_factor(bool manifold, const StateT<T>& a, const StateT<T>& b, const T& bias_old, const T& bias_new, double imu_rot, double weight,
                       double raw_dt, double bias_weight, T res[2]) {
  if (manifold) {
    const Se2<T> M0{a.pose[0], a.pose[1], a.pose[2], a.pose[3]};
    const Se2<T> M1 = se2t_mul(Se2<T>{b.pose[0], b.pose[1], b.pose[2], b.pose[3]}, se2t_exp(T(0.0), T(0.0), bias_new * raw_dt));
    T lg[3];
    se2t_log(se2t_mul(se2t_inverse(M0), M1), lg);
    res[0] = weight * (imu_rot - lg[2]);
  } else {
    res[0] = weight * (imu_rot - normalize_angle(b.rot - a.rot + bias_new * raw_dt));
  }
  res[1] = bias_weight * (bias_new - bias_old);
}

// ---- the window problem -------------------------------------------------------------------------------------------------
// states[0] is the constant state trajectory.end()[-W-1]; states[1..W] are free.  Segment w (0-based) of the pair list belongs to
// states[w + 1]: its residual blocks over all fixed maps in the order estimateTransformCeres adds them.
struct WindowProblem {
  WindowParams P;
  std::vector<WState> st;                 // W + 1 states (current values are written back by unpack)
  std::vector<double> imu;                // the IMU constraint of the factor between states[j-1] and states[j]: imu[j-1]
  const Cell12* cm = nullptr; const Cell12* cf = nullptr;
  const uint32_t* im = nullptr; const uint32_t* jf = nullptr;
  std::vector<uint32_t> seg_off;          // [W + 1]
  int W = 0;
  // states[0] = trajectory.end()[-W-1]: addMotionParameterBlock(..., set_constant = true) pins its POSE only (ndt_matcher.cpp:304-312) and
  // addImuParameterBlock(..., true) its bias; its velocities (and, without the constant-velocity model, its acceleration) stay free.
  std::vector<int> amb_off, tan_off;      // per state: start of its free parameters in the ambient / tangent vectors
  int n_amb_ = 0, n_tan_ = 0;

  bool pose_free(int j) const { return j >= 1; }
  bool bias_free(int j) const { return j >= 1 && P.use_imu; }
  bool acc_free() const { return !P.constant_velocity; }
  void layout() {
    W = (int)st.size() - 1;
    amb_off.assign(W + 1, 0); tan_off.assign(W + 1, 0);
    n_amb_ = n_tan_ = 0;
    for (int j = 0; j <= W; ++j) {
      amb_off[j] = n_amb_; tan_off[j] = n_tan_;
      const int rest = 3 + (acc_free() ? 2 : 0) + (bias_free(j) ? 1 : 0);
      n_amb_ += rest + (pose_free(j) ? (P.manifold ? 4 : 3) : 0);
      n_tan_ += rest + (pose_free(j) ? 3 : 0);
    }
  }
  int n_amb() const { return n_amb_; }
  int n_tan() const { return n_tan_; }
  // free parameters of state j in the ambient vector: [pose] | vel | omega | [acc] | [bias]
  void put_state(const WState& s, int j, double* p) const {
    int o = 0;
    if (pose_free(j)) {
      if (P.manifold) { for (int i = 0; i < 4; ++i) p[o++] = s.pose[i]; }
      else { p[o++] = s.pos[0]; p[o++] = s.pos[1]; p[o++] = s.rot; }
    }
    p[o++] = s.lin_vel[0]; p[o++] = s.lin_vel[1]; p[o++] = s.rot_vel;
    if (acc_free()) { p[o++] = s.lin_acc[0]; p[o++] = s.lin_acc[1]; }
    if (bias_free(j)) p[o++] = s.imu_bias;
  }
  void get_state(WState& s, int j, const double* p) const {
    int o = 0;
    if (pose_free(j)) {
      if (P.manifold) { for (int i = 0; i < 4; ++i) s.pose[i] = p[o++]; }
      else { s.pos[0] = p[o++]; s.pos[1] = p[o++]; s.rot = p[o++]; }
    }
    s.lin_vel[0] = p[o++]; s.lin_vel[1] = p[o++]; s.rot_vel = p[o++];
    if (acc_free()) { s.lin_acc[0] = p[o++]; s.lin_acc[1] = p[o++]; }
    if (bias_free(j)) s.imu_bias = p[o++];
  }
  void pack(double* x) const { for (int j = 0; j <= W; ++j) put_state(st[j], j, x + amb_off[j]); }
  void unpack(const double* x) { for (int j = 0; j <= W; ++j) get_state(st[j], j, x + amb_off[j]); }
  void plus(const double* x, const double* d, double* xp) const {
    for (int j = 0; j <= W; ++j) {
      const double* p = x + amb_off[j]; const double* dd = d + tan_off[j]; double* q = xp + amb_off[j];
      const int na = (j < W ? amb_off[j + 1] : n_amb_) - amb_off[j];
      if (pose_free(j) && P.manifold) { se2_plus(p, dd, q); for (int i = 0; i < na - 4; ++i) q[4 + i] = p[4 + i] + dd[3 + i]; }
      else for (int i = 0; i < na; ++i) q[i] = p[i] + dd[i];
    }
  }

  // the parameters of state j under the ambient vector x as dual numbers seeded at slots base + (pose 0-3 | pos 0-1, rot 2 | vel 4-5 |
  // omega 6 | acc 7-8 | bias 9); constant parameters get plain values
  template <int N>
  void load_state(int j, const double* x, int base, StateT<Jet<N>>& S, Jet<N>& bias) const {
    WState s = st[j];
    get_state(s, j, x + amb_off[j]);
    auto mk = [&](double v, int slot, bool free_) { return free_ ? Jet<N>(v, base + slot) : Jet<N>(v); };
    for (int i = 0; i < 4; ++i) S.pose[i] = mk(s.pose[i], i, pose_free(j) && P.manifold);
    S.pos[0] = mk(s.pos[0], 0, pose_free(j) && !P.manifold); S.pos[1] = mk(s.pos[1], 1, pose_free(j) && !P.manifold);
    S.rot = mk(s.rot, 2, pose_free(j) && !P.manifold);
    S.vel[0] = mk(s.lin_vel[0], 4, true); S.vel[1] = mk(s.lin_vel[1], 5, true); S.omega = mk(s.rot_vel, 6, true);
    S.acc[0] = mk(s.lin_acc[0], 7, acc_free()); S.acc[1] = mk(s.lin_acc[1], 8, acc_free());
    bias = mk(s.imu_bias, 9, bias_free(j));
  }
  // adds a residual block (dual-number residuals over the 2 x 10 seed slots of states js[0], js[1]) to cost / g / H
  void add_block(const Jet<20>* res, int nres, const int js[2], const double* x, double* cost, double* g, double* H) const {
    const int nt = n_tan();
    std::vector<double> Jrow(nt);
    for (int r = 0; r < nres; ++r) {
      std::fill(Jrow.begin(), Jrow.end(), 0.0);
      for (int side = 0; side < 2; ++side) {
        const int j = js[side];
        const double* v = res[r].v + side * 10;
        int o = tan_off[j];
        if (pose_free(j)) {
          if (P.manifold) {
            double Pj[12]; se2_plus_jacobian(x + amb_off[j], Pj);
            for (int a = 0; a < 3; ++a) { double s = 0; for (int i = 0; i < 4; ++i) s += v[i] * Pj[i * 3 + a]; Jrow[o + a] += s; }
          } else for (int a = 0; a < 3; ++a) Jrow[o + a] += v[a];
          o += 3;
        }
        Jrow[o++] += v[4]; Jrow[o++] += v[5]; Jrow[o++] += v[6];
        if (acc_free()) { Jrow[o++] += v[7]; Jrow[o++] += v[8]; }
        if (bias_free(j)) Jrow[o++] += v[9];
      }
      *cost += 0.5 * res[r].a * res[r].a;
      if (g) for (int a = 0; a < nt; ++a) {
        g[a] += Jrow[a] * res[r].a;
        for (int b = 0; b < nt; ++b) H[a * nt + b] += Jrow[a] * Jrow[b];
      }
    }
  }

  // cost (+ tangent gradient and J^T J when g != nullptr) of the whole problem at x under `loss` on the NDT blocks
  bool evaluate(const double* x, const Loss& loss, double* cost, double* g, double* H, double* max_raw = nullptr) const {
    const int nt = n_tan();
    *cost = 0;
    if (g) { std::fill(g, g + nt, 0.0); std::fill(H, H + nt * nt, 0.0); }
    const int np = variant_num_params(P.variant);
    double mr = 0; bool first = true;
    for (int j = 1; j <= W; ++j) {
      const double* p = x + amb_off[j];
      const uint32_t b = seg_off[j - 1], e = seg_off[j];
      FusedOut f;
      accumulate_pairs(P.variant, p, cm, cf, im + b, jf + b, e - b, loss, g != nullptr, f);
      if (e > b) { mr = first ? f.max_r : std::max(mr, f.max_r); first = false; }
      *cost += f.cost;
      if (g) {
        const int t0 = tan_off[j];
        if (P.manifold) {
          double Pj[12]; se2_plus_jacobian(p, Pj);
          for (int a = 0; a < 3; ++a) {
            double s = 0; for (int i = 0; i < 4; ++i) s += Pj[i * 3 + a] * f.g[i];
            g[t0 + a] += s;
            for (int c2 = 0; c2 < 3; ++c2) { double h = 0; for (int i = 0; i < 4; ++i) for (int l = 0; l < 4; ++l) h += Pj[i * 3 + a] * f.H[i * 4 + l] * Pj[l * 3 + c2]; H[(t0 + a) * nt + t0 + c2] += h; }
          }
        } else {
          for (int a = 0; a < np; ++a) { g[t0 + a] += f.g[a]; for (int c2 = 0; c2 < np; ++c2) H[(t0 + a) * nt + t0 + c2] += f.H[a * 4 + c2]; }
        }
      }
    }
    if (max_raw) *max_raw = mr;
    for (int j = 1; j <= W; ++j) {
      StateT<Jet<20>> A, B; Jet<20> ba, bb;
      load_state<20>(j - 1, x, 0, A, ba);
      load_state<20>(j, x, 10, B, bb);
      const double dt = st[j].stamp - st[j - 1].stamp;
      const int js[2] = {j - 1, j};
      Jet<20> res[8];
      if (P.manifold) motion_factor_se2(A, B, dt, P.sqrtI, res); else motion_factor_vec(A, B, dt, P.sqrtI, res);
      add_block(res, 8, js, x, cost, g, H);
      if (P.use_imu) {
        Jet<20> ri[2];
        imu_factor(P.manifold, A, B, ba, bb, imu[j - 1], P.weight_imu, dt, P.weight_imu_bias, ri);
        add_block(ri, 2, js, x, cost, g, H);
      }
    }
    return std::isfinite(*cost);
  }
};

// Matcher::estimateTransformCeres.  trans4: prior in (rejection gate), estimate out.  states: W + 1 window states, oldest (constant) first;
// the free ones are updated in place (pose / pos, rot as ceres leaves them: only the newest state has both representations synchronised).
inline WindowResult window_solve(WindowProblem& wp, double trans4[4], const LmOptions& opt, size_t n_cells_total) {
  WindowResult R;
  wp.layout();
  const int W = wp.W;
  if (W < 1 || wp.seg_off[W] == 0) { R.status = 1; return R; }
  R.n_free = W; R.n_tangent = wp.n_tan();
  const double prior_t[2] = {trans4[2], trans4[3]};
  const double prior_rot = std::atan2(trans4[1], trans4[0]);
  std::vector<double> x(wp.n_amb());
  wp.pack(x.data());
  Loss none; none.kind = LOSS_NONE;
  double c0, max_r = 0;
  wp.evaluate(x.data(), none, &c0, nullptr, nullptr, &max_r);
  R.max_residual = max_r;
  double mu = gnc_initial_mu(max_r, wp.P.loss_scale, wp.P.divisor, wp.P.gnc_steps);
  R.mu_first = mu;
  const double weight = wp.P.ndt_weight / (double)(n_cells_total * (size_t)wp.P.k);
  LmSummary last;
  do {
    mu = std::max(mu, 1.0);
    Loss loss; loss.kind = LOSS_BARRON; loss.a = wp.P.loss_scale; loss.alpha = wp.P.alpha; loss.mu = mu; loss.weight = weight;
    NormalEqProblem prob;
    prob.n_amb = wp.n_amb(); prob.n_tan = wp.n_tan();
    prob.eval_full = [&](const double* xx, double* cost, double* g, double* H) { return wp.evaluate(xx, loss, cost, g, H); };
    prob.eval_cost = [&](const double* xx, double* cost) { return wp.evaluate(xx, loss, cost, nullptr, nullptr); };
    prob.plus = [&](const double* xx, const double* d, double* xp) { wp.plus(xx, d, xp); };
    last = lm_minimize(prob, opt, x.data());
    R.gnc_solves++;
    R.total_iterations += last.num_iterations;
    mu /= wp.P.divisor;
  } while (mu > 1.0 / std::sqrt(wp.P.divisor));
  R.final_cost = last.final_cost;
  wp.unpack(x.data());
  WState& n = wp.st[W];
  if (!wp.P.manifold) { n.pose[0] = std::cos(n.rot); n.pose[1] = std::sin(n.rot); n.pose[2] = n.pos[0]; n.pose[3] = n.pos[1]; }
  else { n.pos[0] = n.pose[2]; n.pos[1] = n.pose[3]; n.rot = std::atan2(n.pose[1], n.pose[0]); }
  // (pose.so2().inverse() * SO2d(prior_rotation)).log()
  double re = n.pose[0] * std::cos(prior_rot) + n.pose[1] * std::sin(prior_rot), imv = n.pose[0] * std::sin(prior_rot) - n.pose[1] * std::cos(prior_rot);
  const double n2 = re * re + imv * imv;
  if (n2 != 1.0) { const double sc = 2.0 / (1.0 + n2); re *= sc; imv *= sc; }
  if (std::fabs(n.pose[2] - prior_t[0]) > wp.P.reject_translation || std::fabs(n.pose[3] - prior_t[1]) > wp.P.reject_translation ||
      std::fabs(std::atan2(imv, re)) > wp.P.reject_rotation) {
    const WState& p = wp.st[W - 1];
    R.rejected = 1;
    n.pos[0] = p.pos[0]; n.pos[1] = p.pos[1]; n.rot = p.rot;
    for (int i = 0; i < 4; ++i) n.pose[i] = p.pose[i];
    n.lin_vel[0] = n.lin_vel[1] = 0.0; n.rot_vel = 0.0; n.lin_acc[0] = n.lin_acc[1] = 0.0;
    n.imu_bias = p.imu_bias;
  }
  for (int i = 0; i < 4; ++i) trans4[i] = n.pose[i];
  return R;
}

}  // namespace orc
