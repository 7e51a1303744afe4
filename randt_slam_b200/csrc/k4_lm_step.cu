// K4 — batched graduated-non-convexity + Levenberg-Marquardt bookkeeping for many independent registrations (sm_100a).
//
// Replaces, for every segment (= registration) of a problem at once, the solve half of
//   Matcher::estimateLoopConstraint   R/src/ndt_registration/ndt_matcher.cpp:457-492
//   (and the GNC loop of Matcher::estimateTransformCeres, :372-397, for NDT-only problems):
//     raw residual maximum -> mu0 = min(2 max_r^2 / a^2, div^(steps-1));  do { mu = max(mu,1); ceres::Solve; mu /= div } while (mu > 1/sqrt(div))
// where ceres::Solve is the trust-region minimiser of Ceres 2.1.0 with the LEVENBERG_MARQUARDT strategy on a dense problem of
// 3 or 4 unknowns (the reference asks for DENSE_QR on [J; sqrt(D)]; the damped normal equations (J^T J + D/radius) y = J^T r
// solved here by Cholesky are the same system).  Restated from the published algorithm of Ceres 2.1.0
// (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc; not vendored by the reference): Jacobi column scaling fixed at
// iteration 0, LM diagonal clamped to [1e-6, 1e32] and reused after a rejected step, radius /= max(1/3, 1 - (2 rho - 1)^3) on
// success, radius /= 2, 4, 8.. on failure, parameter / function / gradient tolerance tests in ceres' order, and
// Sophus::Manifold<SE2>::Plus / PlusJacobian (Sophus 1.22.10) when the pose block carries the manifold.
//
// The evaluation itself is K3 (k3_pair_eval.cu): every LM iteration is ONE fused K3 launch over all still-active segments at
// their requested poses (with their own mu), followed by ONE launch of this kernel — one thread per segment — that consumes
// the 24-double records, advances each segment's state machine and posts the next pose to evaluate.  No host round trip per
// iteration; the host only polls the count of active segments every few iterations.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace randt {

namespace {

enum { PH_INIT = 0, PH_SOLVE_START = 1, PH_CANDIDATE = 2, PH_DONE = 3 };
enum { TERM_CONVERGENCE = 0, TERM_NO_CONVERGENCE = 1, TERM_FAILURE = 2 };

__device__ __forceinline__ double norm_n(const double* v, int n) { double s = 0; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return sqrt(s); }

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

// in-place Cholesky solve of the n x n system A y = b (row-major, leading dimension 4); false if A is not positive definite
template <int n>
__device__ __forceinline__ bool chol_solve(double* A, const double* b, double* y) {
  _Pragma("unroll") for (int j = 0; j < n; ++j) {
    double d = A[j * 4 + j];
    _Pragma("unroll") for (int k = 0; k < j; ++k) d -= A[j * 4 + k] * A[j * 4 + k];
    if (!(d > 0.0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j * 4 + j] = d;
    _Pragma("unroll") for (int i = j + 1; i < n; ++i) {
      double s = A[i * 4 + j];
      _Pragma("unroll") for (int k = 0; k < j; ++k) s -= A[i * 4 + k] * A[j * 4 + k];
      A[i * 4 + j] = s / d;
    }
  }
  _Pragma("unroll") for (int i = 0; i < n; ++i) { double s = b[i]; _Pragma("unroll") for (int k = 0; k < i; ++k) s -= A[i * 4 + k] * y[k]; y[i] = s / A[i * 4 + i]; }
  _Pragma("unroll") for (int i = n - 1; i >= 0; --i) { double s = y[i]; _Pragma("unroll") for (int k = i + 1; k < n; ++k) s -= A[k * 4 + i] * y[k]; y[i] = s / A[i * 4 + i]; }
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
    _Pragma("unroll") for (int i = 0; i < nt; ++i) _Pragma("unroll") for (int j = 0; j < nt; ++j) A[i * 4 + j] = Hs[i * 4 + j] + (i == j ? st.diag[i] / st.radius : 0.0);
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
  sn = sqrt(sn);
  if (sn <= o.parameter_tolerance * (st.x_norm + o.parameter_tolerance)) { st.termination = TERM_CONVERGENCE; return false; }
  const double cost_change = st.cost - cand_cost;
  if (fabs(cost_change) <= o.function_tolerance * st.cost) { st.termination = TERM_CONVERGENCE; return false; }
  const double relative_decrease = cost_change / st.model_cost_change;
  st.solve_iterations += 1;
  if (relative_decrease > o.min_relative_decrease) {
    _Pragma("unroll") for (int i = 0; i < D::np; ++i) st.x[i] = st.cand[i];
    st.x_norm = norm_n(st.x, D::np);
    load_normal_eq<D>(rec, st.x, st);       // the candidate was evaluated with its Jacobian: nothing to re-evaluate
    st.n_jac_evals += 1;
    st.gmax = grad_max_norm<D>(st);
    const double t = 2.0 * relative_decrease - 1.0;
    st.radius = st.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
    st.radius = fmin(o.max_trust_region_radius, st.radius);
    st.decrease_factor = 2.0; st.reuse_diagonal = 0;
    st.last_successful = 1;
    st.min_cost = fmin(st.min_cost, st.cost);
  } else {
    st.radius = st.radius / st.decrease_factor; st.decrease_factor *= 2.0; st.reuse_diagonal = 1;
    st.last_successful = 0;
    st.min_cost = fmin(st.min_cost, cand_cost);
  }
  return true;
}

// The per-segment state lives in memory word-major ([word][segment], 53 x 8 bytes per segment): thread s of a warp reads word i of 32
// neighbouring segments from one 256-byte run instead of 32 sectors 424 bytes apart.
constexpr int kStateWords = (int)(sizeof(LmState) / 8);
static_assert(sizeof(LmState) % 8 == 0, "LmState is moved as 8-byte words");
__device__ __forceinline__ void load_state(const LmState* __restrict__ base, uint32_t S, uint32_t s, LmState& st) {
  const unsigned long long* mem = reinterpret_cast<const unsigned long long*>(base);
  unsigned long long w[kStateWords];
#pragma unroll
  for (int i = 0; i < kStateWords; ++i) w[i] = mem[(size_t)i * S + s];
  memcpy(&st, w, sizeof(LmState));
}
__device__ __forceinline__ void store_state(LmState* __restrict__ base, uint32_t S, uint32_t s, const LmState& st) {
  unsigned long long* mem = reinterpret_cast<unsigned long long*>(base);
  unsigned long long w[kStateWords];
  memcpy(w, &st, sizeof(LmState));
#pragma unroll
  for (int i = 0; i < kStateWords; ++i) mem[(size_t)i * S + s] = w[i];
}

__global__ void __launch_bounds__(128) k4_init_kernel(uint32_t S, int np, const double* __restrict__ poses0, LmState* __restrict__ state,
                                                      double* __restrict__ eval_pose, double* __restrict__ mu, uint32_t* __restrict__ active,
                                                      double* __restrict__ rec, uint32_t* __restrict__ n_active) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) *n_active = S;
  if (s >= S) return;
  LmState st;
  memset(&st, 0, sizeof(st));
  for (int i = 0; i < np; ++i) { st.x[i] = poses0[(size_t)s * np + i]; eval_pose[(size_t)s * np + i] = st.x[i]; }
  st.phase = PH_INIT;
  st.mu = 1.0;
  store_state(state, S, s, st);
  mu[s] = 1.0;
  active[s] = 1u;
  for (int i = 0; i < RANDT_FUSED_STRIDE; ++i) rec[(size_t)s * RANDT_FUSED_STRIDE + i] = 0.0;   // segments without pairs get no K3 record
}

template <int NP, bool MANIFOLD>
__global__ void __launch_bounds__(64) k4_lm_step_kernel(uint32_t S, randt_solver_options o,
                                                         const double* __restrict__ rec_all, LmState* __restrict__ state,
                                                         double* __restrict__ eval_pose, double* __restrict__ mu_arr,
                                                         uint32_t* __restrict__ active, uint32_t* __restrict__ n_active,
                                                         double* __restrict__ poses_out, double* __restrict__ result) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the records of the K3 launch just before
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S || active[s] == 0u) return;
  typedef Dims<NP, MANIFOLD> D;
  constexpr int np = NP;
  const double* rec = rec_all + (size_t)s * RANDT_FUSED_STRIDE;
  LmState st;
  load_state(state, S, s, st);
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
    mu_arr[s] = st.mu;
    for (int i = 0; i < np; ++i) eval_pose[(size_t)s * np + i] = st.x[i];
  } else if (st.phase == PH_CANDIDATE) {
    for (int i = 0; i < np; ++i) eval_pose[(size_t)s * np + i] = st.cand[i];
  }
  if (st.phase == PH_DONE) {
    active[s] = 0u;
    atomicSub(n_active, 1u);
    for (int i = 0; i < np; ++i) poses_out[(size_t)s * np + i] = st.x[i];
    double* r = result + (size_t)s * RANDT_REG_STRIDE;
    r[RANDT_REG_SCORE] = st.n_blocks ? st.final_cost / (double)st.n_blocks : 0.0;    // summary.final_cost / num_residual_blocks (:492)
    r[RANDT_REG_FINAL_COST] = st.final_cost;
    r[RANDT_REG_GNC_SOLVES] = (double)st.gnc_solves;
    r[RANDT_REG_ITERATIONS] = (double)st.total_iterations;
    r[RANDT_REG_EVALS] = (double)(st.n_cost_evals + st.n_jac_evals);
    r[RANDT_REG_MU_FIRST] = st.mu_first;
    r[RANDT_REG_STATUS] = (double)st.status;
    r[RANDT_REG_TERMINATION] = (double)st.termination;
  }
  store_state(state, S, s, st);
}

// ---- re-planning the K3 schedule for the segments that are still active ------------------------------------------------
// As segments finish, the static chunk list fills with chunks K3 must step over; when the active count has dropped enough, the
// solver compacts the list (order preserved, so tiles stay contiguous) and hands every warp an equal share of what is left.
__global__ void __launch_bounds__(256) k4_chunk_flags_kernel(const ChunkDesc* __restrict__ chunks, uint32_t n, const uint32_t* __restrict__ active,
                                                             uint32_t* __restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = active[chunks[i].seg] != 0u ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k4_compact_chunks_kernel(const ChunkDesc* __restrict__ chunks, uint32_t n, const uint32_t* __restrict__ flags,
                                                                const uint32_t* __restrict__ scan, ChunkDesc* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) out[scan[i]] = chunks[i];
}
// warp w starts at the first tile boundary at or after its equal share w * n_kept / n_warps
__global__ void __launch_bounds__(256) k4_warp_ranges_kernel(const ChunkDesc* __restrict__ kept, const uint32_t* __restrict__ n_kept_ptr, uint32_t n_warps,
                                                             uint32_t* __restrict__ warp_off) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > n_warps) return;
  const uint32_t n = *n_kept_ptr;
  uint32_t i = (uint32_t)(((unsigned long long)w * n + n_warps - 1u) / n_warps);
  if (w == n_warps) i = n;
  while (i < n && !(kept[i].meta & kChunkFirst)) ++i;
  warp_off[w] = i;
}

}  // namespace

cudaError_t launch_lm_init(uint32_t S, int np, const double* d_poses0, LmState* state, double* eval_pose, double* mu, uint32_t* active,
                           double* rec, uint32_t* n_active, cudaStream_t s, int* n_launches) {
  const int grid = (int)((S + 127u) / 128u);
  k4_init_kernel<<<grid > 0 ? grid : 1, 128, 0, s>>>(S, np, d_poses0, state, eval_pose, mu, active, rec, n_active);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_lm_step(uint32_t S, int np, int use_manifold, const randt_solver_options& o, const double* rec, LmState* state,
                           double* eval_pose, double* mu, uint32_t* active, uint32_t* n_active, double* poses_out, double* result,
                           cudaStream_t s, int* n_launches) {
  if (S == 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((S + 63u) / 64u); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (n_launches) *n_launches += 1;
  if (np == 4 && use_manifold) return cudaLaunchKernelEx(&cfg, k4_lm_step_kernel<4, true>, S, o, rec, state, eval_pose, mu, active, n_active, poses_out, result);
  if (np == 4) return cudaLaunchKernelEx(&cfg, k4_lm_step_kernel<4, false>, S, o, rec, state, eval_pose, mu, active, n_active, poses_out, result);
  return cudaLaunchKernelEx(&cfg, k4_lm_step_kernel<3, false>, S, o, rec, state, eval_pose, mu, active, n_active, poses_out, result);
}

cudaError_t launch_replan(const ChunkDesc* chunks, uint32_t n_chunks, const uint32_t* active, uint32_t n_warps, uint32_t* flags, uint32_t* scan,
                          uint32_t* block_sums, ChunkDesc* kept, uint32_t* warp_off, cudaStream_t s, int* n_launches) {
  if (n_chunks == 0) return cudaSuccess;
  const unsigned grid = (n_chunks + 255u) / 256u;
  k4_chunk_flags_kernel<<<grid, 256, 0, s>>>(chunks, n_chunks, active, flags);
  cudaError_t e = launch_exclusive_scan_u32(flags, scan, n_chunks, block_sums, s, n_launches);
  if (e != cudaSuccess) return e;
  k4_compact_chunks_kernel<<<grid, 256, 0, s>>>(chunks, n_chunks, flags, scan, kept);
  k4_warp_ranges_kernel<<<(n_warps + 256u) / 256u, 256, 0, s>>>(kept, scan + n_chunks, n_warps, warp_off);
  if (n_launches) *n_launches += 3;
  return cudaGetLastError();
}

}  // namespace randt
