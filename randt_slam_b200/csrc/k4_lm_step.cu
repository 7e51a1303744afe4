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

#include "k4_device.cuh"

namespace randt {

namespace {

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
  lm_advance<D>(o, rec, st, eval_pose + (size_t)s * np, mu_arr + s);
  if (st.phase == PH_DONE) {
    active[s] = 0u;
    atomicSub(n_active, 1u);
    for (int i = 0; i < np; ++i) poses_out[(size_t)s * np + i] = st.x[i];
    lm_write_result(st, result + (size_t)s * RANDT_REG_STRIDE);
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
