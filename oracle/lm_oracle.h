// ORACLE — TEST INFRASTRUCTURE ONLY (see ndt_oracle.h header).  PARITY UNPINNED.
//
// Restates, from the published algorithm of Ceres Solver 2.1.0 (pinned by /root/reference/Dockerfile:11-17;
// not vendored under /root/reference), the pieces of `ceres::Solve` the reference's registration calls rely on
// (R/src/ndt_registration/ndt_matcher.cpp:372-397, 457-483): trust-region minimiser with the
// Levenberg-Marquardt strategy (initial radius 1e4, min/max LM diagonal 1e-6/1e32, radius update
// r /= max(1/3, 1-(2q-1)^3) on accept, r /= 2,4,8.. on reject), Jacobi column scaling fixed at iteration 0,
// convergence tests (function 1e-6, gradient 1e-10, parameter 1e-8), per-block robust-loss correction.
// The reference selects DENSE_QR on [J; D]; the oracle solves the mathematically identical damped normal
// equations (J^T J + D^T D) y = J^T r by Cholesky (n <= 18 unknowns).
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <vector>

#include "ndt_oracle.h"

namespace orc {

struct LmOptions {
  int max_num_iterations = 200;          // ndt_radar_slam_base_parameters.yaml: max_iteration
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  double min_relative_decrease = 1e-3;
  int max_num_consecutive_invalid_steps = 5;
  bool jacobi_scaling = true;
};

enum LmTermination { LM_CONVERGENCE = 0, LM_NO_CONVERGENCE = 1, LM_FAILURE = 2 };

struct LmSummary {
  double initial_cost = 0, final_cost = 0;
  int num_iterations = 0;      // iterations.size() in ceres terms (includes iteration 0)
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  int num_cost_evals = 0, num_jac_evals = 0;
  int termination = LM_NO_CONVERGENCE;
  int num_residual_blocks = 0;
};

// A problem in "normal equation" form.  n_amb ambient parameters, n_tan tangent dimensions.
struct NormalEqProblem {
  int n_amb = 0, n_tan = 0;
  // full evaluation: cost, gradient g = J^T r (n_tan), H = J^T J (n_tan x n_tan row-major); J = local (tangent) Jacobian
  std::function<bool(const double* x, double* cost, double* g, double* H)> eval_full;
  std::function<bool(const double* x, double* cost)> eval_cost;
  std::function<void(const double* x, const double* delta, double* x_plus)> plus;
};

inline bool cholesky_solve(std::vector<double> A, std::vector<double> b, int n, double* x) {
  // in-place LL^T on copies
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i * n + k] * x[k]; x[i] = s / A[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * x[k]; x[i] = s / A[i * n + i]; }
  for (int i = 0; i < n; ++i) if (!std::isfinite(x[i])) return false;
  return true;
}

inline double vec_norm(const double* v, int n) { double s = 0; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return std::sqrt(s); }

inline LmSummary lm_minimize(const NormalEqProblem& P, const LmOptions& o, double* x_io) {
  LmSummary S;
  const int na = P.n_amb, nt = P.n_tan;
  std::vector<double> x(x_io, x_io + na), cand(na), g(nt), H(nt * nt), gs(nt), Hs(nt * nt), scale(nt, 1.0),
      diag(nt), step(nt), delta(nt), y(nt), tmp(na), negg(nt);
  double cost = 0;
  if (!P.eval_full(x.data(), &cost, g.data(), H.data())) { S.termination = LM_FAILURE; return S; }
  S.num_jac_evals++;
  S.initial_cost = cost;
  double min_recorded_cost = cost;
  S.num_iterations = 1;
  if (o.jacobi_scaling) for (int i = 0; i < nt; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(H[i * nt + i]));
  auto rescale = [&]() {
    for (int i = 0; i < nt; ++i) { gs[i] = g[i] * scale[i]; for (int j = 0; j < nt; ++j) Hs[i * nt + j] = H[i * nt + j] * scale[i] * scale[j]; }
  };
  auto grad_max_norm = [&]() {
    for (int i = 0; i < nt; ++i) negg[i] = -g[i];
    P.plus(x.data(), negg.data(), tmp.data());
    double m = 0; for (int i = 0; i < na; ++i) m = std::max(m, std::fabs(x[i] - tmp[i])); return m;
  };
  rescale();
  double x_norm = vec_norm(x.data(), na);
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  bool last_successful = true;  // iteration 0 counts as successful for the gradient test
  int consecutive_invalid = 0;
  double gmax = grad_max_norm();
  int iteration = 0;
  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= o.max_num_iterations) { S.termination = LM_NO_CONVERGENCE; break; }
    if (last_successful && gmax <= o.gradient_tolerance) { S.termination = LM_CONVERGENCE; break; }
    if (radius <= o.min_trust_region_radius) { S.termination = LM_CONVERGENCE; break; }
    ++iteration;
    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal)
      for (int i = 0; i < nt; ++i) diag[i] = std::min(std::max(Hs[i * nt + i], o.min_lm_diagonal), o.max_lm_diagonal);
    std::vector<double> A(Hs);
    for (int i = 0; i < nt; ++i) A[i * nt + i] += diag[i] / radius;
    bool valid = cholesky_solve(A, gs, nt, y.data());
    reuse_diagonal = true;
    double model_cost_change = 0;
    if (valid) {
      for (int i = 0; i < nt; ++i) step[i] = -y[i];
      double sg = 0, sHs = 0;
      for (int i = 0; i < nt; ++i) { sg += step[i] * gs[i]; double r = 0; for (int j = 0; j < nt; ++j) r += Hs[i * nt + j] * step[j]; sHs += step[i] * r; }
      model_cost_change = -(sg + 0.5 * sHs);
      if (!(model_cost_change > 0.0)) valid = false;
    }
    if (!valid) {
      S.num_iterations++; S.num_unsuccessful_steps++;
      last_successful = false;
      if (++consecutive_invalid >= o.max_num_consecutive_invalid_steps) { S.termination = LM_FAILURE; break; }
      radius *= 0.5; reuse_diagonal = true;
      continue;
    }
    consecutive_invalid = 0;
    for (int i = 0; i < nt; ++i) delta[i] = step[i] * scale[i];
    P.plus(x.data(), delta.data(), cand.data());
    double cand_cost = 0;
    S.num_cost_evals++;
    if (!P.eval_cost(cand.data(), &cand_cost) || !std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached
    double sn = 0; for (int i = 0; i < na; ++i) sn += (x[i] - cand[i]) * (x[i] - cand[i]); sn = std::sqrt(sn);
    if (sn <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { S.termination = LM_CONVERGENCE; break; }
    // FunctionToleranceReached
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= o.function_tolerance * cost) { S.termination = LM_CONVERGENCE; break; }
    const double relative_decrease = cost_change / model_cost_change;
    S.num_iterations++;
    if (relative_decrease > o.min_relative_decrease) {
      x = cand; x_norm = vec_norm(x.data(), na);
      if (!P.eval_full(x.data(), &cost, g.data(), H.data())) { S.termination = LM_FAILURE; break; }
      S.num_jac_evals++;
      rescale();
      gmax = grad_max_norm();
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      last_successful = true; S.num_successful_steps++;
      min_recorded_cost = std::min(min_recorded_cost, cost);
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      last_successful = false; S.num_unsuccessful_steps++;
      min_recorded_cost = std::min(min_recorded_cost, cand_cost);
    }
  }
  S.final_cost = min_recorded_cost;
  for (int i = 0; i < na; ++i) x_io[i] = x[i];
  return S;
}

// ------------------------------------------------------------------------------------------
// Sophus 1.22.10 SE2 pieces (restated): exp, group product (with the conditional renormalisation of
// the unit complex), Manifold<SE2>::Plus = T * exp(delta), PlusJacobian = Dx_this_mul_exp_x_at_0.
// ------------------------------------------------------------------------------------------
inline void se2_exp(const double a[3], double out[4]) {
  const double theta = a[2];
  const double c = std::cos(theta), s = std::sin(theta);
  double sbt, omcbt;
  if (std::fabs(theta) < 1e-10) {
    const double t2 = theta * theta;
    sbt = 1.0 - (1.0 / 6.0) * t2;
    omcbt = 0.5 * theta - (1.0 / 24.0) * theta * t2;
  } else { sbt = s / theta; omcbt = (1.0 - c) / theta; }
  out[0] = c; out[1] = s;
  out[2] = sbt * a[0] - omcbt * a[1];
  out[3] = omcbt * a[0] + sbt * a[1];
}
inline void se2_mul(const double A[4], const double B[4], double out[4]) {
  double re = A[0] * B[0] - A[1] * B[1], im = A[0] * B[1] + A[1] * B[0];
  const double n2 = re * re + im * im;
  if (n2 != 1.0) { const double sc = 2.0 / (1.0 + n2); re *= sc; im *= sc; }
  const double tx = A[2] + (A[0] * B[2] - A[1] * B[3]);
  const double ty = A[3] + (A[1] * B[2] + A[0] * B[3]);
  out[0] = re; out[1] = im; out[2] = tx; out[3] = ty;
}
inline void se2_plus(const double T[4], const double d[3], double out[4]) { double e[4]; se2_exp(d, e); se2_mul(T, e, out); }
inline void se2_plus_jacobian(const double T[4], double J[12]) {  // 4 x 3 row-major
  const double c = T[0], s = T[1];
  const double v[12] = {0, 0, -s, 0, 0, c, c, -s, 0, s, c, 0};
  for (int i = 0; i < 12; ++i) J[i] = v[i];
}

// ------------------------------------------------------------------------------------------
// Single-pose NDT problem (all residual blocks hang off one 4-parameter pose block).
// accumulate: sum over pairs, in pair order, of the loss-corrected blocks.
// ------------------------------------------------------------------------------------------
struct FusedOut {           // what one evaluation pass produces for one pose
  double H[16];             // ambient 4x4 J~^T J~ (row-major, symmetric)
  double g[4];              // J~^T r~
  double cost;              // sum 0.5 * rho(r^2)
  double max_r;             // max raw residual (GNC)
  double sum_sq;            // sum raw r^2
  uint32_t n;               // residual blocks
};
inline void accumulate_pairs(int variant, const double* params, const Cell12* cm, const Cell12* cf,
                             const uint32_t* im, const uint32_t* jf, size_t P, const Loss& loss, bool want_jac, FusedOut& out) {
  const int np = variant_num_params(variant);
  std::memset(&out, 0, sizeof(out));
  for (size_t p = 0; p < P; ++p) {
    double r, J[4] = {0, 0, 0, 0}, cost;
    eval_pair_autodiff(variant, params, cm[im[p]], cf[jf[p]], &r, want_jac ? J : nullptr);
    out.max_r = (p == 0) ? r : std::max(out.max_r, r);
    const double sq = r * r;
    out.sum_sq += sq;
    double rho[3];
    loss.evaluate(sq, rho);
    cost = 0.5 * rho[0];
    if (want_jac) {
      Corrector corr(sq, rho);
      corr.correct(r, J, np);
      for (int a = 0; a < np; ++a) { out.g[a] += J[a] * r; for (int b = 0; b < np; ++b) out.H[a * 4 + b] += J[a] * J[b]; }
    }
    out.cost += cost;
  }
  out.n = (uint32_t)P;
}

struct LoopConstraintResult {
  double pose[4];
  double score;            // summary.final_cost / summary.num_residual_blocks  (ndt_matcher.cpp:492)
  int gnc_solves;
  int total_iterations;
  int total_evals;
  double gnc_mu_first;
  int status;              // 0 ok, 1 no residuals
};

// Matcher::estimateLoopConstraint   R/src/ndt_registration/ndt_matcher.cpp:426-493, shipped flags
// (optimize_on_manifold = true, autodiff, intensity): the residuals hang off a 4-double block that has NO
// manifold (quirk B.13), so ceres optimises the raw ambient [c, s, tx, ty].
// `on_manifold` = true instead attaches Sophus::Manifold<SE2> (what estimateTransformCeres does for its poses).
inline LoopConstraintResult loop_constraint(const NdtMap& fixed, const NdtMap& moving, const double pose0[4], int k, int metric,
                                            int variant, double matcher_loss_scale, double loop_scale, double alpha,
                                            double divisor, int max_gnc_steps, const LmOptions& opt, bool on_manifold,
                                            double loss_weight = 1.0, const PairList* pairs_in = nullptr) {
  LoopConstraintResult R{};
  for (int i = 0; i < 4; ++i) R.pose[i] = pose0[i];
  PairList pl;
  if (pairs_in) pl = *pairs_in; else associate(fixed, moving, pose0, k, metric, pl);
  const size_t P = pl.im.size();
  if (P == 0) { R.status = 1; return R; }
  const int np = variant_num_params(variant);
  Loss none; none.kind = LOSS_NONE;
  FusedOut f0;
  accumulate_pairs(variant, R.pose, moving.cells.data(), fixed.cells.data(), pl.im.data(), pl.jf.data(), P, none, false, f0);
  double mu = gnc_initial_mu(f0.max_r, matcher_loss_scale, divisor, max_gnc_steps);
  R.gnc_mu_first = mu;
  LmSummary last;
  do {
    mu = std::max(mu, 1.0);
    Loss loss; loss.kind = LOSS_BARRON; loss.a = loop_scale; loss.alpha = alpha; loss.mu = mu; loss.weight = loss_weight;
    NormalEqProblem prob;
    prob.n_amb = np; prob.n_tan = (on_manifold && np == 4) ? 3 : np;
    prob.eval_full = [&](const double* x, double* cost, double* g, double* H) {
      FusedOut f;
      accumulate_pairs(variant, x, moving.cells.data(), fixed.cells.data(), pl.im.data(), pl.jf.data(), P, loss, true, f);
      *cost = f.cost;
      if (prob.n_tan == np) {
        for (int a = 0; a < np; ++a) { g[a] = f.g[a]; for (int b = 0; b < np; ++b) H[a * np + b] = f.H[a * 4 + b]; }
      } else {
        double Pj[12]; se2_plus_jacobian(x, Pj);
        for (int a = 0; a < 3; ++a) {
          g[a] = 0; for (int i = 0; i < 4; ++i) g[a] += Pj[i * 3 + a] * f.g[i];
          for (int b = 0; b < 3; ++b) { double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += Pj[i * 3 + a] * f.H[i * 4 + j] * Pj[j * 3 + b]; H[a * 3 + b] = s; }
        }
      }
      return std::isfinite(f.cost);
    };
    prob.eval_cost = [&](const double* x, double* cost) {
      FusedOut f;
      accumulate_pairs(variant, x, moving.cells.data(), fixed.cells.data(), pl.im.data(), pl.jf.data(), P, loss, false, f);
      *cost = f.cost; return std::isfinite(f.cost);
    };
    if (prob.n_tan == np) prob.plus = [np](const double* x, const double* d, double* xp) { for (int i = 0; i < np; ++i) xp[i] = x[i] + d[i]; };
    else prob.plus = [](const double* x, const double* d, double* xp) { se2_plus(x, d, xp); };
    last = lm_minimize(prob, opt, R.pose);
    R.gnc_solves++;
    R.total_iterations += last.num_iterations;
    R.total_evals += last.num_cost_evals + last.num_jac_evals;
    mu /= divisor;
  } while (mu > 1.0 / std::sqrt(divisor));
  R.score = last.final_cost / (double)P;
  return R;
}

// a13  cost-only sweep over candidate poses (the inner evaluation of estimateTransformGlobalBNB,
// R/src/ndt_registration/ndt_matcher.cpp:560-576): cost_s = sum 0.5 rho(r^2) / num_blocks.
inline void sweep_costs(int variant, const Cell12* cm, const Cell12* cf, const uint32_t* im, const uint32_t* jf, size_t P,
                        const Loss& loss, const double* poses, size_t S, double* cost_out) {
  const int np = variant_num_params(variant);
  for (size_t s = 0; s < S; ++s) {
    FusedOut f;
    accumulate_pairs(variant, poses + s * np, cm, cf, im, jf, P, loss, false, f);
    cost_out[s] = f.cost;
  }
}

// a13  Matcher::estimateTransformGlobalBNB   R/src/ndt_registration/ndt_matcher.cpp:495-608, restated sequentially exactly as written:
// coarse grid over the window, FIFO queue, cost = sum 0.5 rho(r^2) / num_blocks under a bare BarronLoss(scale, alpha) (mu = 1),
// prune at cost_threshold, expand 3 x 3 x 3 children at half the linear step, de-duplicate on the float-cast 3x3 matrix.
struct BnbResult { double pose[4]; double min_cost; int n_evaluated; };
inline void se2_from(double a, double tx, double ty, double out[4]) { out[0] = std::cos(a); out[1] = std::sin(a); out[2] = tx; out[3] = ty; }
inline BnbResult bnb_search(const NdtMap& fixed, const NdtMap& moving, const double trans0[4], int variant, double alpha, double scale,
                            double window_linear, double window_angular, double linear_step, double max_px_range, double cost_threshold,
                            int n_iter_in, int metric = LOOKUP_MAHALANOBIS_INTENSITY) {
  BnbResult R{};
  PairList pl;
  associate(fixed, moving, trans0, 4, metric, pl);   // n_neighbours = 4 (ndt_matcher.cpp:521)
  const size_t P = pl.im.size();
  Loss loss; loss.kind = LOSS_BARRON; loss.a = scale; loss.alpha = alpha; loss.mu = 1.0; loss.weight = 1.0;
  const double angular_step = std::acos(1 - ((linear_step * linear_step) / (2 * max_px_range * max_px_range)));
  const size_t n_iter = (size_t)n_iter_in;
  const double initial_linear_step = std::pow(2, (double)n_iter - 1) * linear_step;
  struct Node { double t[4]; size_t level; };
  std::vector<std::vector<float>> calculated;
  auto key = [](const double t[4]) { return std::vector<float>{(float)t[0], (float)t[1], 0.f, (float)-t[1], (float)t[0], 0.f, (float)t[2], (float)t[3], 1.f}; };
  std::vector<Node> queue;   // FIFO: index `head` is the front
  for (double tx = -window_linear / 2.0; tx <= window_linear / 2.0; tx += initial_linear_step)
    for (double ty = -window_linear / 2.0; ty <= window_linear / 2.0; ty += initial_linear_step)
      for (double a = -window_angular / 2.0; a < window_angular / 2.0; a += angular_step) {
        double d[4]; se2_from(a, tx, ty, d);
        Node n; se2_mul(trans0, d, n.t); n.level = 1;
        queue.push_back(n);
        calculated.push_back(key(n.t));
      }
  double min_cost = 100000.0;
  double best[4] = {1, 0, 0, 0};
  const int np = variant_num_params(variant);
  for (size_t head = 0; head < queue.size(); ++head) {
    const Node cur = queue[head];
    double params[4] = {cur.t[0], cur.t[1], cur.t[2], cur.t[3]};
    if (np == 3) { params[0] = cur.t[2]; params[1] = cur.t[3]; params[2] = std::atan2(cur.t[1], cur.t[0]); }
    FusedOut f;
    accumulate_pairs(variant, params, moving.cells.data(), fixed.cells.data(), pl.im.data(), pl.jf.data(), P, loss, false, f);
    R.n_evaluated++;
    const double current_cost = f.cost / (double)P;
    if (current_cost < cost_threshold) {
      if (current_cost < min_cost) { for (int i = 0; i < 4; ++i) best[i] = cur.t[i]; min_cost = current_cost; }
      if (cur.level < n_iter) {
        const double step = std::pow(2.0, (double)cur.level) * linear_step;
        for (double tx = -step; tx <= step; tx += step)
          for (double ty = -step; ty <= step; ty += step)
            for (double a = -angular_step; a <= angular_step; a += angular_step) {
              double d[4]; se2_from(a, tx, ty, d);
              Node n; se2_mul(cur.t, d, n.t); n.level = cur.level + 1;
              const std::vector<float> k = key(n.t);
              if (std::find(calculated.begin(), calculated.end(), k) == calculated.end()) { calculated.push_back(k); queue.push_back(n); }
            }
      }
    }
  }
  for (int i = 0; i < 4; ++i) R.pose[i] = best[i];
  R.min_cost = min_cost;
  return R;
}

}  // namespace orc
