// K7 — persistent registration solver: ONE WARP solves one registration from its initial guess to its final pose inside a single
// launch (sm_100a).
//
// Replaces Matcher::estimateLoopConstraint's whole solve (R/src/ndt_registration/ndt_matcher.cpp:457-492: the GNC loop around
// ceres::Solve, LM, <= 200 iterations) — and the NDT-only GNC loop of estimateTransformCeres (:372-397) — for every registration of
// a batch whose pair list is short enough for one warp (<= kSolveMaxDuos duos = 1024 pairs; a shipped-config registration has ~180).
// The stepwise path (K3 fused launch + K4 launch per LM iteration, randt_register_batch's first implementation) pays two launches
// and a trip through HBM per iteration and leaves the GPU idle while a batch's last stragglers finish; here a warp
//   * claims a registration from a global counter (dynamic: iteration counts differ by 5x between registrations),
//   * stages its duo records once into shared memory with bulk (TMA) copies — up to 3 chunks = 96 duos stay resident for the whole
//     solve; longer registrations re-stream their chunks through the same 3-buffer ring on every evaluation (L2 hits),
//   * evaluates residuals / Jacobians / Barron corrector / J^T J, J^T r with K3's per-duo code (k3_device.cuh), reduces across the
//     warp in K3's fixed order, and
//   * lets lane 0 run ceres' trust-region bookkeeping (k4_device.cuh: the same state machine K4 runs) on the 24-double record in
//     shared memory, which posts the next pose / mu to evaluate.
// The summation order of a registration depends on nothing but its own pair list, so its result does not depend on which other
// registrations share the batch or on how a batch is sharded over GPUs (tests/test_literal_batch_gpu.py).
// Tried and dropped: four registrations per warp with their solver steps side by side on lanes 0..3 (the serial fp64 chain paid once
// per round) — the records of four registrations no longer stay resident (43 KB per warp), and re-streaming them from L2 on every
// evaluation cost more than the shared solver phase saved (16 384 registrations: 5.3 ms against 4.8 ms; a lone registration 36 % slower).
#include <atomic>

#include "k3_device.cuh"
#include "k4_device.cuh"

namespace randt {

namespace {

#ifndef RANDT_K7_WARPS
#define RANDT_K7_WARPS 4
#endif
constexpr int kK7Warps = RANDT_K7_WARPS;      // warps per CTA
constexpr int kK7Bufs = 3;       // chunk buffers per warp
#ifndef RANDT_K7_MIN_CTAS
#define RANDT_K7_MIN_CTAS 4
#endif
constexpr int kK7MinCtas = RANDT_K7_MIN_CTAS;    // CTAs per SM (shared memory: ~11.6 KB per warp)

template <int NS>
struct __align__(128) SolveWarp {
  float4 rec[kK7Bufs][32 * kRecF4];     // the registration's duo records, chunk c in buffer c % kK7Bufs
  double out[RANDT_FUSED_STRIDE];       // record of the evaluation just made
  LmState st;                           // ceres state of the registration (lane 0)
  PoseConst kc; LossConst lc;           // constants of the evaluation in flight
  double eval_pose[4]; double mu;       // what the state machine asked for
  unsigned long long bar[kK7Bufs];      // one mbarrier per chunk buffer
};

template <int VARIANT, int LOSS, bool MANIFOLD>
__global__ void __launch_bounds__(kK7Warps * 32, kK7MinCtas)
k7_solve_kernel(DeviceProblem P, SolveLayout L, LossParams lp, const double* __restrict__ weight_per_seg, randt_solver_options o,
                const double* __restrict__ poses0, double* __restrict__ poses_out, double* __restrict__ result,
                unsigned long long* __restrict__ bad_counter) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NP = VarTraits<VARIANT>::NP;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NS = NH + NB + 2;
  typedef Dims<NP, MANIFOLD> D;
  typedef SolveWarp<NS> W;
  extern __shared__ __align__(128) unsigned char k7_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  W& w = reinterpret_cast<W*>(k7_smem)[warp];
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kK7Bufs; ++b) mbar_init(&w.bar[b], 1u);
    fence_mbar_init();
  }
  __syncwarp();
  const uint32_t omap = out_map<VARIANT, true>(lane);
  const int src_lane = bfly_owner<NS>((int)(omap & 31u));      // where the butterfly leaves the slot this lane's record entry needs
  uint32_t phase_bits = 0u;     // bit b: parity the next completion of buffer b's barrier will have
  const uint32_t cpt = L.tile_duos >> 5;     // chunks per full tile (tiles are multiples of 32 duos except a segment's last)
  while (true) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(L.next_item, 1u);
    item = __shfl_sync(kFull, item, 0);
    if (item >= L.n_items) break;
    const uint32_t seg = L.items ? L.items[item] : item;
    const uint32_t nd = L.seg_duo_off[seg + 1] - L.seg_duo_off[seg];
    const uint32_t n_chunks = (nd + 31u) >> 5;
    const uint32_t first_tile = P.seg_first_tile[seg];
    const uint32_t n_pairs_seg = P.seg_off[seg + 1] - P.seg_off[seg];
    const bool resident = n_chunks <= (uint32_t)kK7Bufs;
    LossParams lps = lp;                       // ScaledLoss weight of this registration (ndt_weight / (n_cells k) differs per scan)
    if (weight_per_seg) lps.weight = weight_per_seg[seg];
    const uint32_t ahead = resident ? (uint32_t)kK7Bufs : (uint32_t)(kK7Bufs - 1);
    auto issue = [&](uint32_t c) {           // lane 0: chunk c of this registration -> buffer c % kK7Bufs
      const uint32_t n_here = min(32u, nd - (c << 5));
      const uint32_t t = c / cpt;
      const DuoRec* src = P.duo_recs + L.tile_rec_begin[first_tile + t] + ((c - t * cpt) << 5);
      const uint32_t b = c % (uint32_t)kK7Bufs;
      const uint32_t bytes = n_here * (uint32_t)sizeof(DuoRec);
      mbar_expect_tx(&w.bar[b], bytes);
      bulk_g2s(&w.rec[b][0], src, bytes, &w.bar[b]);
    };
    if (lane == 0) {
      LmState& st = w.st;
      memset(&st, 0, sizeof(LmState));
#pragma unroll
      for (int i = 0; i < NP; ++i) { st.x[i] = poses0[(size_t)seg * NP + i]; w.eval_pose[i] = st.x[i]; }
      st.phase = PH_INIT;
      st.mu = 1.0;
      w.mu = 1.0;
    }
    __syncwarp();
    bool first_eval = true;
    while (true) {
      // ---- one evaluation at (w.eval_pose, w.mu) ----
      const bool load = !resident || first_eval;
      if (lane == 0) {
        if (load) for (uint32_t c = 0; c < min(n_chunks, ahead); ++c) issue(c);
        PoseConst k0; LossConst l0;
        make_pose_const<VARIANT>(w.eval_pose, k0);
        make_loss_const(lps, w.mu, l0);
        w.kc = k0; w.lc = l0;
      }
      __syncwarp();
      double acc[NS]; double max_dd = 0.0; uint32_t n_bad = 0;
#pragma unroll
      for (int e = 0; e < NS; ++e) acc[e] = 0.0;
      for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t b = c % (uint32_t)kK7Bufs;
        if (load) {
          // buffer (c + ahead) % kK7Bufs held chunk c - 1, which every lane has consumed (warp barrier at the end of the last round)
          if (!resident && lane == 0 && c + ahead < n_chunks) issue(c + ahead);
          mbar_wait(&w.bar[b], (phase_bits >> b) & 1u);
          phase_bits ^= 1u << b;
        }
        const uint32_t n_here = min(32u, nd - (c << 5));
        if ((uint32_t)lane < n_here) accumulate_duo<VARIANT, LOSS, true, NS>(w.kc, w.lc, &w.rec[b][0], P.duo_overflow, lane, acc, max_dd, n_bad);
        __syncwarp();
      }
      bfly_reduce<NS, 16>(acc, lane);             // fixed exchange pattern, no scratch: acc[0] = the 32-lane total of this lane's slot
      const double mx = warp_max_nonneg(max_dd);
      const uint32_t bad = __reduce_add_sync(kFull, n_bad);
      write_segment_out_from(acc[0], mx, w.kc.ja, w.kc.jb, n_pairs_seg, w.out, 0u, 0u, lane, omap, src_lane);
      if (lane == 0 && bad) atomicAdd(bad_counter, (unsigned long long)bad);
      __syncwarp();
      first_eval = false;
      // ---- ceres' bookkeeping on the record: next candidate / next GNC solve / done ----
      if (lane == 0) lm_advance<D>(o, w.out, w.st, w.eval_pose, &w.mu);
      __syncwarp();
      if (w.st.phase == PH_DONE) break;
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) poses_out[(size_t)seg * NP + i] = w.st.x[i];
      lm_write_result(w.st, result + (size_t)seg * RANDT_REG_STRIDE);
    }
    __syncwarp();
  }
}

template <typename K>
cudaError_t k7_allow_smem(K kernel, size_t bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

int k7_loss_code(const LossParams& lp) {
  if (lp.kind == RANDT_LOSS_NONE) return L_NONE;
  if (lp.kind == RANDT_LOSS_WELSCH) return L_WELSCH;
  if (lp.alpha == -2.0) return L_BARRON_M2;
  if (lp.alpha == -1.0) return L_BARRON_M1;
  return L_BARRON;
}

template <int VARIANT, int LOSS, bool MANIFOLD>
cudaError_t launch_solve_vlm(const DeviceProblem& p, const SolveLayout& L, const LossParams& lp, const double* wps, const randt_solver_options& o,
                             const double* poses0, double* poses_out, double* result, unsigned long long* bad, cudaStream_t s) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NS = NB * (NB + 1) / 2 + NB + 2;
  constexpr size_t smem = (size_t)kK7Warps * sizeof(SolveWarp<NS>);
  static std::atomic<unsigned long long> done{0ull};
  if (cudaError_t rc = k7_allow_smem(k7_solve_kernel<VARIANT, LOSS, MANIFOLD>, smem, done)) return rc;
  const uint32_t max_ctas = (uint32_t)(kSmCount * kK7MinCtas);
  const uint32_t grid = std::max(1u, std::min(max_ctas, (L.n_items + kK7Warps - 1u) / kK7Warps));
  k7_solve_kernel<VARIANT, LOSS, MANIFOLD><<<grid, kK7Warps * 32, smem, s>>>(p, L, lp, wps, o, poses0, poses_out, result, bad);
  return cudaGetLastError();
}
template <int VARIANT, bool MANIFOLD>
cudaError_t launch_solve_vm(const DeviceProblem& p, const SolveLayout& L, const LossParams& lp, const double* wps, const randt_solver_options& o,
                            const double* poses0, double* poses_out, double* result, unsigned long long* bad, cudaStream_t s) {
  switch (k7_loss_code(lp)) {
    case L_NONE: return launch_solve_vlm<VARIANT, L_NONE, MANIFOLD>(p, L, lp, wps, o, poses0, poses_out, result, bad, s);
    case L_WELSCH: return launch_solve_vlm<VARIANT, L_WELSCH, MANIFOLD>(p, L, lp, wps, o, poses0, poses_out, result, bad, s);
    case L_BARRON_M2: return launch_solve_vlm<VARIANT, L_BARRON_M2, MANIFOLD>(p, L, lp, wps, o, poses0, poses_out, result, bad, s);
    case L_BARRON_M1: return launch_solve_vlm<VARIANT, L_BARRON_M1, MANIFOLD>(p, L, lp, wps, o, poses0, poses_out, result, bad, s);
    default: return launch_solve_vlm<VARIANT, L_BARRON, MANIFOLD>(p, L, lp, wps, o, poses0, poses_out, result, bad, s);
  }
}

}  // namespace

cudaError_t launch_solve_persistent(const DeviceProblem& p, const SolveLayout& L, int variant, int use_manifold, const LossParams& lp,
                                    const double* d_weight_per_seg, const randt_solver_options& o, const double* d_poses0, double* d_poses_out, double* d_result,
                                    unsigned long long* d_bad, cudaStream_t s, int* n_launches) {
  if (L.n_items == 0) return cudaSuccess;
  cudaError_t e;
  // the manifold only exists for the 4-parameter SE2 pose blocks (variants 0 and 1), as in K4
  switch (variant) {
    case 0: e = use_manifold ? launch_solve_vm<0, true>(p, L, lp, d_weight_per_seg, o, d_poses0, d_poses_out, d_result, d_bad, s)
                             : launch_solve_vm<0, false>(p, L, lp, d_weight_per_seg, o, d_poses0, d_poses_out, d_result, d_bad, s); break;
    case 1: e = use_manifold ? launch_solve_vm<1, true>(p, L, lp, d_weight_per_seg, o, d_poses0, d_poses_out, d_result, d_bad, s)
                             : launch_solve_vm<1, false>(p, L, lp, d_weight_per_seg, o, d_poses0, d_poses_out, d_result, d_bad, s); break;
    case 2: e = launch_solve_vm<2, false>(p, L, lp, d_weight_per_seg, o, d_poses0, d_poses_out, d_result, d_bad, s); break;
    case 3: e = launch_solve_vm<3, false>(p, L, lp, d_weight_per_seg, o, d_poses0, d_poses_out, d_result, d_bad, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (n_launches) *n_launches += 1;
  return e;
}

}  // namespace randt
