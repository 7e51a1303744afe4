// The per-registration state machine of the batched GNC + Levenberg-Marquardt solver (ceres 2.1.0's trust-region minimiser restated, see
// k4_lm_step.cu), shared by the step kernel K4 (one thread per registration, one launch per iteration) and the persistent solver K7
// (k7_solve.cu: one warp per registration, the whole solve in one launch).
#pragma once
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace randt {

namespace {

enum { PH_INIT = 0, PH_SOLVE_START = 1, PH_CANDIDATE = 2, PH_DONE = 3 };
enum { TERM_CONVERGENCE = 0, TERM_NO_CONVERGENCE = 1, TERM_FAILURE = 2 };

__device__ __forceinline__ double lm_sqrt(double x);
__device__ __forceinline__ double norm_n(const double* v, int n) { double s = 0; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return lm_sqrt(s); }

// Sophus SE2::exp and group product (with the conditional renormalisation of the unit complex number)
__device__ void se2_plus(const double* T, const double* d, double* out) {
  const double theta = d[2];
  double s, c;
  sincos(theta, &s, &c);
  double sbt, omcbt;
  if (fabs(theta) < 1e-10) {
    const double t2 = theta * theta;
    sbt = 1.0 - (1.0 / 6.0) * t2;
    omcbt = 0.5 * theta - (1.0 / 24.0) * theta * t2;
  } else { sbt = s / theta; omcbt = (1.0 - c) / theta; }
  const double ex = sbt * d[0] - omcbt * d[1], ey = omcbt * d[0] + sbt * d[1];
  double re = T[0] * c - T[1] * s, im = T[0] * s + T[1] * c;
  const double n2 = re * re + im * im;
  if (n2 != 1.0) { const double sc = 2.0 / (1.0 + n2); re *= sc; im *= sc; }
  out[0] = re; out[1] = im;
  out[2] = T[2] + (T[0] * ex - T[1] * ey);
  out[3] = T[3] + (T[1] * ex + T[0] * ey);
}

// compile-time problem shape: NP ambient parameters, NT tangent dimensions (all loops unroll, every array stays in registers)
template <int NP_, bool MANIFOLD_>
struct Dims {
  static constexpr int np = NP_;
  static constexpr bool manifold = MANIFOLD_;
  static constexpr int nt = MANIFOLD_ ? 3 : NP_;
};

template <typename D>
__device__ __forceinline__ void plus(const double* x, const double* delta, double* out) {
  if (D::manifold) se2_plus(x, delta, out);
  else {
#pragma unroll
    for (int i = 0; i < D::np; ++i) out[i] = x[i] + delta[i];
  }
}

// cost, tangent gradient and tangent J^T J of the evaluated point `x` from a K3 fused record (ambient 4x4 layout)
template <typename D>
__device__ __forceinline__ void load_normal_eq(const double* __restrict__ rec, const double* x, LmState& st) {
  st.cost = rec[RANDT_FUSED_COST];
  if (!D::manifold) {
    _Pragma("unroll") for (int a = 0; a < D::nt; ++a) {
      st.g[a] = rec[RANDT_FUSED_G + a];
      _Pragma("unroll") for (int b = 0; b < D::nt; ++b) st.H[a * 4 + b] = rec[RANDT_FUSED_H + a * 4 + b];
    }
  } else {
    // Sophus::Manifold<SE2>::PlusJacobian = Dx_this_mul_exp_x_at_0 (4 x 3): rows [0,0,-s], [0,0,c], [c,-s,0], [s,c,0]
    const double c = x[0], s = x[1];
    const double Pj[4][3] = {{0, 0, -s}, {0, 0, c}, {c, -s, 0}, {s, c, 0}};
    _Pragma("unroll") for (int a = 0; a < 3; ++a) {
      double ga = 0;
      _Pragma("unroll") for (int i = 0; i < 4; ++i) ga += Pj[i][a] * rec[RANDT_FUSED_G + i];
      st.g[a] = ga;
      _Pragma("unroll") for (int b = 0; b < 3; ++b) {
        double h = 0;
        _Pragma("unroll") for (int i = 0; i < 4; ++i) _Pragma("unroll") for (int j = 0; j < 4; ++j) h += Pj[i][a] * rec[RANDT_FUSED_H + i * 4 + j] * Pj[j][b];
        st.H[a * 4 + b] = h;
      }
    }
  }
}

template <typename D>
__device__ __forceinline__ double grad_max_norm(const LmState& st) {
  double negg[4], tmp[4];
  _Pragma("unroll") for (int i = 0; i < D::nt; ++i) negg[i] = -st.g[i];
  plus<D>(st.x, negg, tmp);
  double m = 0;
  _Pragma("unroll") for (int i = 0; i < D::np; ++i) m = fmax(m, fabs(st.x[i] - tmp[i]));
  return m;
}

// 1/x and 1/sqrt(x) for a normal, finite, positive x: special-function seed (2^-23) + two Newton steps, no slow path.  The solver's
// serial chain is made of these (a correctly rounded fp64 divide or square root costs ~10x the latency); results are within 1-2 ulp.
__device__ __forceinline__ double lm_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double lm_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
}
__device__ __forceinline__ double lm_sqrt(double x) { return x > 0.0 ? x * lm_rsqrt(x) : 0.0; }

// in-place Cholesky solve of the n x n system A y = b (row-major, leading dimension 4); false if A is not positive definite.
// One reciprocal square root per column: L_ij and the two substitutions multiply by 1 / L_jj instead of dividing.
template <int n>
__device__ __forceinline__ bool chol_solve(double* A, const double* b, double* y) {
  double inv[n];
  _Pragma("unroll") for (int j = 0; j < n; ++j) {
    double d = A[j * 4 + j];
    _Pragma("unroll") for (int k = 0; k < j; ++k) d -= A[j * 4 + k] * A[j * 4 + k];
    if (!(d > 0.0) || !isfinite(d)) return false;
    inv[j] = lm_rsqrt(d);
    _Pragma("unroll") for (int i = j + 1; i < n; ++i) {
      double s = A[i * 4 + j];
      _Pragma("unroll") for (int k = 0; k < j; ++k) s -= A[i * 4 + k] * A[j * 4 + k];
      A[i * 4 + j] = s * inv[j];
    }
  }
  _Pragma("unroll") for (int i = 0; i < n; ++i) { double s = b[i]; _Pragma("unroll") for (int k = 0; k < i; ++k) s -= A[i * 4 + k] * y[k]; y[i] = s * inv[i]; }
  _Pragma("unroll") for (int i = n - 1; i >= 0; --i) { double s = y[i]; _Pragma("unroll") for (int k = i + 1; k < n; ++k) s -= A[k * 4 + i] * y[k]; y[i] = s * inv[i]; }
  _Pragma("unroll") for (int i = 0; i < n; ++i) if (!isfinite(y[i])) return false;
  return true;
}

template <typename D>
__device__ __forceinline__ void begin_solve(const randt_solver_options& o, const double* __restrict__ rec, LmState& st) {
  load_normal_eq<D>(rec, st.x, st);
  st.n_jac_evals += 1;
  st.min_cost = st.cost;
  st.solve_iterations = 1;
  _Pragma("unroll") for (int i = 0; i < D::nt; ++i) st.scale[i] = o.jacobi_scaling ? 1.0 / (1.0 + sqrt(st.H[i * 4 + i])) : 1.0;
  st.x_norm = norm_n(st.x, D::np);
  st.radius = o.initial_trust_region_radius;
  st.decrease_factor = 2.0;
  st.reuse_diagonal = 0;
  st.last_successful = 1;     // iteration 0 counts as successful for the gradient test
  st.consecutive_invalid = 0;
  st.gmax = grad_max_norm<D>(st);
  st.iteration = 0;
}

// Runs the minimiser until it needs the cost at a candidate point (true, st.cand set) or terminates (false, st.termination set).
template <typename D>
__device__ __forceinline__ bool next_candidate(const randt_solver_options& o, LmState& st) {
  constexpr int nt = D::nt;
  while (true) {
    if (st.iteration >= o.max_num_iterations) { st.termination = TERM_NO_CONVERGENCE; return false; }
    if (st.last_successful && st.gmax <= o.gradient_tolerance) { st.termination = TERM_CONVERGENCE; return false; }
    if (st.radius <= o.min_trust_region_radius) { st.termination = TERM_CONVERGENCE; return false; }
    st.iteration += 1;
    double gs[4], Hs[16], A[16], y[4], step[4];
    _Pragma("unroll") for (int i = 0; i < nt; ++i) {
      gs[i] = st.g[i] * st.scale[i];
      _Pragma("unroll") for (int j = 0; j < nt; ++j) Hs[i * 4 + j] = st.H[i * 4 + j] * st.scale[i] * st.scale[j];
    }
    if (!st.reuse_diagonal)
      _Pragma("unroll") for (int i = 0; i < nt; ++i) st.diag[i] = fmin(fmax(Hs[i * 4 + i], o.min_lm_diagonal), o.max_lm_diagonal);
    const double inv_radius = lm_rcp(st.radius);
    _Pragma("unroll") for (int i = 0; i < nt; ++i) _Pragma("unroll") for (int j = 0; j < nt; ++j) A[i * 4 + j] = Hs[i * 4 + j] + (i == j ? st.diag[i] * inv_radius : 0.0);
    bool valid = chol_solve<nt>(A, gs, y);
    st.reuse_diagonal = 1;
    if (valid) {
      double sg = 0, sHs = 0;
      _Pragma("unroll") for (int i = 0; i < nt; ++i) step[i] = -y[i];
      _Pragma("unroll") for (int i = 0; i < nt; ++i) {
        sg += step[i] * gs[i];
        double r = 0;
        _Pragma("unroll") for (int j = 0; j < nt; ++j) r += Hs[i * 4 + j] * step[j];
        sHs += step[i] * r;
      }
      st.model_cost_change = -(sg + 0.5 * sHs);
      if (!(st.model_cost_change > 0.0)) valid = false;
    }
    if (!valid) {
      st.solve_iterations += 1;
      st.last_successful = 0;
      if (++st.consecutive_invalid >= o.max_num_consecutive_invalid_steps) { st.termination = TERM_FAILURE; return false; }
      st.radius *= 0.5;
      continue;
    }
    st.consecutive_invalid = 0;
    double delta[4];
    _Pragma("unroll") for (int i = 0; i < nt; ++i) delta[i] = step[i] * st.scale[i];
    plus<D>(st.x, delta, st.cand);
    return true;
  }
}

// Consumes the evaluation of st.cand.  Returns false when the solve terminated on a tolerance.
template <typename D>
__device__ __forceinline__ bool on_candidate(const randt_solver_options& o, const double* __restrict__ rec, LmState& st) {
  double cand_cost = rec[RANDT_FUSED_COST];
  if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
  st.n_cost_evals += 1;
  double sn = 0;
  _Pragma("unroll") for (int i = 0; i < D::np; ++i) sn += (st.x[i] - st.cand[i]) * (st.x[i] - st.cand[i]);
  sn = lm_sqrt(sn);
  if (sn <= o.parameter_tolerance * (st.x_norm + o.parameter_tolerance)) { st.termination = TERM_CONVERGENCE; return false; }
  const double cost_change = st.cost - cand_cost;
  if (fabs(cost_change) <= o.function_tolerance * st.cost) { st.termination = TERM_CONVERGENCE; return false; }
  const double relative_decrease = cost_change * lm_rcp(st.model_cost_change);
  st.solve_iterations += 1;
  if (relative_decrease > o.min_relative_decrease) {
    _Pragma("unroll") for (int i = 0; i < D::np; ++i) st.x[i] = st.cand[i];
    st.x_norm = norm_n(st.x, D::np);
    load_normal_eq<D>(rec, st.x, st);       // the candidate was evaluated with its Jacobian: nothing to re-evaluate
    st.n_jac_evals += 1;
    st.gmax = grad_max_norm<D>(st);
    const double t = 2.0 * relative_decrease - 1.0;
    st.radius = st.radius * lm_rcp(fmax(1.0 / 3.0, 1.0 - t * t * t));
    st.radius = fmin(o.max_trust_region_radius, st.radius);
    st.decrease_factor = 2.0; st.reuse_diagonal = 0;
    st.last_successful = 1;
    st.min_cost = fmin(st.min_cost, st.cost);
  } else {
    st.radius = st.radius * lm_rcp(st.decrease_factor); st.decrease_factor *= 2.0; st.reuse_diagonal = 1;
    st.last_successful = 0;
    st.min_cost = fmin(st.min_cost, cand_cost);
  }
  return true;
}


// One transition of a registration's state machine on the record `rec` of the evaluation it asked for: seeds the GNC from the first
// (loss-free) evaluation, starts a solve, consumes a candidate, moves to the next GNC solve or finishes.  Afterwards, unless
// st.phase == PH_DONE, `eval_pose` / `mu_out` hold the next evaluation the registration needs.
template <typename D>
__device__ __forceinline__ void lm_advance(const randt_solver_options& o, const double* __restrict__ rec, LmState& st, double* eval_pose,
                                           double* mu_out) {
  constexpr int np = D::np;
  bool solve_ended = false, need_start = false;
  if (st.phase == PH_INIT) {
    st.n_blocks = (uint32_t)rec[RANDT_FUSED_N];
    if (st.n_blocks == 0u) { st.status = 1; st.phase = PH_DONE; }     // "WARNING: NO RESIDUALS ADDED!" (ndt_matcher.cpp:454-456)
    else {
      const double max_r = rec[RANDT_FUSED_MAXR];
      double m = 2.0 * (max_r * max_r) / (o.gnc_loss_scale * o.gnc_loss_scale);
      m = fmin(m, pow(o.gnc_divisor, (double)(o.gnc_max_steps - 1)));
      st.mu_first = m;
      st.mu = fmax(m, 1.0);
      need_start = true;
    }
  } else if (st.phase == PH_SOLVE_START) {
    begin_solve<D>(o, rec, st);
    if (!next_candidate<D>(o, st)) solve_ended = true;
    else st.phase = PH_CANDIDATE;
  } else if (st.phase == PH_CANDIDATE) {
    if (!on_candidate<D>(o, rec, st) || !next_candidate<D>(o, st)) solve_ended = true;
  }
  if (solve_ended) {
    st.gnc_solves += 1;
    st.total_iterations += st.solve_iterations;
    st.final_cost = st.min_cost;
    st.mu /= o.gnc_divisor;
    if (st.mu > 1.0 / sqrt(o.gnc_divisor)) { st.mu = fmax(st.mu, 1.0); need_start = true; }
    else st.phase = PH_DONE;
  }
  if (need_start) {
    st.phase = PH_SOLVE_START;
    *mu_out = st.mu;
    for (int i = 0; i < np; ++i) eval_pose[i] = st.x[i];
  } else if (st.phase == PH_CANDIDATE) {
    for (int i = 0; i < np; ++i) eval_pose[i] = st.cand[i];
  }
}
// what a finished registration reports (RANDT_REG_*)
__device__ __forceinline__ void lm_write_result(const LmState& st, double* __restrict__ r) {
  r[RANDT_REG_SCORE] = st.n_blocks ? st.final_cost / (double)st.n_blocks : 0.0;    // summary.final_cost / num_residual_blocks (:492)
  r[RANDT_REG_FINAL_COST] = st.final_cost;
  r[RANDT_REG_GNC_SOLVES] = (double)st.gnc_solves;
  r[RANDT_REG_ITERATIONS] = (double)st.total_iterations;
  r[RANDT_REG_EVALS] = (double)(st.n_cost_evals + st.n_jac_evals);
  r[RANDT_REG_MU_FIRST] = st.mu_first;
  r[RANDT_REG_STATUS] = (double)st.status;
  r[RANDT_REG_TERMINATION] = (double)st.termination;
}

}  // namespace

}  // namespace randt
