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
#include <stdlib.h>

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

// ---- team mode: the four warps of a CTA share ONE registration ---------------------------------------------------------------
// A lone registration (the per-scan call of a live stream) or a small batch leaves most of the GPU idle, and one warp walking a
// ~180-pair registration spends two thirds of an evaluation in the per-duo arithmetic of its five or six chunks.  Here warp w prepares
// the chunks c = w, w + 4, ... (residual, Jacobian row, loss terms per pair: everything up to the sums), and warp 0 then adds the
// prepared terms into the accumulators chunk by chunk, lane by lane, with the same fused multiply-adds in the same order as the
// one-warp kernel: the sums — and with them every bit of the result — do not depend on the mode.
constexpr int kTeam = 4;
constexpr int kTermD = 16;       // doubles a lane hands over per duo: N[2][NB] (<= 8), wd[2], wgt[2], hrho[2], dd[2]

template <int VARIANT, int LOSS>
__device__ __forceinline__ uint32_t prepare_duo(const PoseConst& kc, const LossConst& lc, const float4* __restrict__ rec, const DuoRecFull* __restrict__ ovf,
                                                int lane, double* __restrict__ ex /* [kTermD][32] */) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  CellC m, f[2];
  bool two;
  load_duo(rec, lane, ovf, m, f[0], f[1], two);
  Moving mv;
  moving_part<VARIANT>(kc, m, mv);
  double dd[2], N[2][4], wgt[2], hrho[2], wd[2];
  bool ok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) dd[j] = fixed_part<VARIANT, true>(kc, mv, f[j], N[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) ok[j] = dd_valid(dd[j]);
  const bool use1 = two && ok[1];
  uint32_t flags;
  // (the same warp-uniform choice between the guarded and the unguarded loss as accumulate_duo makes)
  if (__all_sync(__activemask(), ok[0] && use1 && dd[0] > 0.0 && dd[1] > 0.0)) {
#pragma unroll
    for (int j = 0; j < 2; ++j) loss_eval<LOSS, true>(dd[j], lc, wgt[j], hrho[j], wd[j]);
    flags = 3u;
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) loss_eval<LOSS>(dd[j], lc, wgt[j], hrho[j], wd[j]);
    flags = (ok[0] ? 1u : 0u) | (use1 ? 2u : 0u) | (ok[0] ? 0u : 4u) | ((two && !ok[1]) ? 8u : 0u);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int a = 0; a < NB; ++a) ex[(j * NB + a) * 32 + lane] = N[j][a];
    ex[(8 + j) * 32 + lane] = wd[j]; ex[(10 + j) * 32 + lane] = wgt[j]; ex[(12 + j) * 32 + lane] = hrho[j]; ex[(14 + j) * 32 + lane] = dd[j];
  }
  return flags;
}

template <int VARIANT, int NS>
__device__ __forceinline__ void apply_duo(const double* __restrict__ ex, uint32_t flags, int lane, double* acc, double& max_dd, uint32_t& n_bad) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NJ = NH + NB;
  n_bad += ((flags >> 2) & 1u) + ((flags >> 3) & 1u);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if ((flags >> j) & 1u) {
      double N[4];
#pragma unroll
      for (int a = 0; a < NB; ++a) N[a] = ex[(j * NB + a) * 32 + lane];
      const double wd = ex[(8 + j) * 32 + lane], wgt = ex[(10 + j) * 32 + lane], hrho = ex[(12 + j) * 32 + lane], dd = ex[(14 + j) * 32 + lane];
      int q = 0;
#pragma unroll
      for (int a = 0; a < NB; ++a) {
        const double wa = wd * N[a];
#pragma unroll
        for (int b2 = a; b2 < NB; ++b2) { acc[q] = fma(wa, N[b2], acc[q]); ++q; }
        acc[NH + a] = fma(wgt, N[a], acc[NH + a]);
      }
      acc[NJ] += hrho;
      acc[NJ + 1] += dd;
      max_dd = fmax(max_dd, dd);
    }
  }
}

template <int NS>
struct __align__(128) SolveTeam {
  float4 rec[kTeam][kK7Bufs][32 * kRecF4];   // warp w's buffers: its k-th chunk (chunk w + 4 k of the registration) in buffer k % kK7Bufs
  double ex[kTeam][kTermD * 32];             // what warp w prepared in this round, [term][lane]
  uint32_t exf[kTeam][32];                   // ... and which of a lane's two pairs count
  double out[RANDT_FUSED_STRIDE];
  LmState st;
  PoseConst kc; LossConst lc;
  double eval_pose[4]; double mu;
  unsigned long long bar[kTeam][kK7Bufs];
  uint32_t item;
};

template <int VARIANT, int LOSS, bool MANIFOLD>
__global__ void __launch_bounds__(kTeam * 32, 3)
k7_team_kernel(DeviceProblem P, SolveLayout L, LossParams lp, const double* __restrict__ weight_per_seg, randt_solver_options o,
               const double* __restrict__ poses0, double* __restrict__ poses_out, double* __restrict__ result,
               unsigned long long* __restrict__ bad_counter) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NP = VarTraits<VARIANT>::NP;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NS = NH + NB + 2;
  typedef Dims<NP, MANIFOLD> D;
  extern __shared__ __align__(128) unsigned char k7_smem[];
  SolveTeam<NS>& T = *reinterpret_cast<SolveTeam<NS>*>(k7_smem);
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kK7Bufs; ++b) mbar_init(&T.bar[warp][b], 1u);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t omap = out_map<VARIANT, true>(lane);
  const int src_lane = bfly_owner<NS>((int)(omap & 31u));
  uint32_t phase_bits = 0u;     // bit b: parity the next completion of this warp's buffer b will have
  const uint32_t cpt = L.tile_duos >> 5;
  while (true) {
    if (threadIdx.x == 0) T.item = atomicAdd(L.next_item, 1u);
    __syncthreads();
    const uint32_t item = T.item;
    if (item >= L.n_items) break;
    const uint32_t seg = L.items ? L.items[item] : item;
    const uint32_t nd = L.seg_duo_off[seg + 1] - L.seg_duo_off[seg];
    const uint32_t n_chunks = (nd + 31u) >> 5;
    const uint32_t first_tile = P.seg_first_tile[seg];
    const uint32_t n_pairs_seg = P.seg_off[seg + 1] - P.seg_off[seg];
    const uint32_t nw = n_chunks > (uint32_t)warp ? (n_chunks - (uint32_t)warp + (uint32_t)kTeam - 1u) / (uint32_t)kTeam : 0u;   // this warp's chunks
    const bool resident = nw <= (uint32_t)kK7Bufs;
    const uint32_t ahead = resident ? (uint32_t)kK7Bufs : (uint32_t)(kK7Bufs - 1);
    LossParams lps = lp;
    if (weight_per_seg) lps.weight = weight_per_seg[seg];
    auto issue = [&](uint32_t k) {           // lane 0: this warp's k-th chunk -> buffer k % kK7Bufs
      const uint32_t c = (uint32_t)warp + (uint32_t)kTeam * k;
      const uint32_t n_here = min(32u, nd - (c << 5));
      const uint32_t t = c / cpt;
      const DuoRec* src = P.duo_recs + L.tile_rec_begin[first_tile + t] + ((c - t * cpt) << 5);
      const uint32_t b = k % (uint32_t)kK7Bufs;
      const uint32_t bytes = n_here * (uint32_t)sizeof(DuoRec);
      mbar_expect_tx(&T.bar[warp][b], bytes);
      bulk_g2s(&T.rec[warp][b][0], src, bytes, &T.bar[warp][b]);
    };
    if (threadIdx.x == 0) {
      LmState& st = T.st;
      memset(&st, 0, sizeof(LmState));
#pragma unroll
      for (int i = 0; i < NP; ++i) { st.x[i] = poses0[(size_t)seg * NP + i]; T.eval_pose[i] = st.x[i]; }
      st.phase = PH_INIT;
      st.mu = 1.0;
      T.mu = 1.0;
    }
    __syncthreads();
    bool first_eval = true;
    const uint32_t n_rounds = (n_chunks + (uint32_t)kTeam - 1u) / (uint32_t)kTeam;
    while (true) {
      // ---- one evaluation at (T.eval_pose, T.mu) ----
      const bool load = !resident || first_eval;
      if (lane == 0 && load) for (uint32_t k = 0; k < min(nw, ahead); ++k) issue(k);
      if (threadIdx.x == 0) {
        PoseConst k0; LossConst l0;
        make_pose_const<VARIANT>(T.eval_pose, k0);
        make_loss_const(lps, T.mu, l0);
        T.kc = k0; T.lc = l0;
      }
      __syncthreads();
      double acc[NS]; double max_dd = 0.0; uint32_t n_bad = 0;
#pragma unroll
      for (int e = 0; e < NS; ++e) acc[e] = 0.0;
      for (uint32_t r = 0; r < n_rounds; ++r) {
        const uint32_t c = r * (uint32_t)kTeam + (uint32_t)warp;
        if (c < n_chunks) {
          const uint32_t b = r % (uint32_t)kK7Bufs;
          if (load) {
            // buffer (r + ahead) % kK7Bufs held this warp's chunk r - 1, which every lane has consumed (the round's barriers)
            if (!resident && lane == 0 && r + ahead < nw) issue(r + ahead);
            mbar_wait(&T.bar[warp][b], (phase_bits >> b) & 1u);
            phase_bits ^= 1u << b;
          }
          const uint32_t n_here = min(32u, nd - (c << 5));
          uint32_t flags = 0u;
          if ((uint32_t)lane < n_here) flags = prepare_duo<VARIANT, LOSS>(T.kc, T.lc, &T.rec[warp][b][0], P.duo_overflow, lane, &T.ex[warp][0]);
          T.exf[warp][lane] = flags;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll 1
          for (uint32_t ww = 0; ww < (uint32_t)kTeam; ++ww) {
            if (r * (uint32_t)kTeam + ww < n_chunks) {
              const uint32_t flags = T.exf[ww][lane];
              if (flags) apply_duo<VARIANT, NS>(&T.ex[ww][0], flags, lane, acc, max_dd, n_bad);
            }
          }
        }
        __syncthreads();
      }
      if (warp == 0) {
        bfly_reduce<NS, 16>(acc, lane);
        const double mx = warp_max_nonneg(max_dd);
        const uint32_t bad = __reduce_add_sync(kFull, n_bad);
        write_segment_out_from(acc[0], mx, T.kc.ja, T.kc.jb, n_pairs_seg, T.out, 0u, 0u, lane, omap, src_lane);
        if (lane == 0 && bad) atomicAdd(bad_counter, (unsigned long long)bad);
        __syncwarp();
        if (lane == 0) lm_advance<D>(o, T.out, T.st, T.eval_pose, &T.mu);
      }
      __syncthreads();
      first_eval = false;
      if (T.st.phase == PH_DONE) break;
    }
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) poses_out[(size_t)seg * NP + i] = T.st.x[i];
      lm_write_result(T.st, result + (size_t)seg * RANDT_REG_STRIDE);
    }
    __syncthreads();
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

// batches up to this many registrations run in team mode (a CTA of four warps per registration, three CTAs to an SM)
constexpr uint32_t kTeamMaxItems = (uint32_t)kSmCount * 3u;
inline bool use_team(uint32_t n_items) {
  static const char* force = getenv("RANDT_K7_TEAM");      // tests: "0" / "1" pin the mode
  if (force && (force[0] == '0' || force[0] == '1')) return force[0] == '1';
  return n_items <= kTeamMaxItems;
}

template <int VARIANT, int LOSS, bool MANIFOLD>
cudaError_t launch_solve_vlm(const DeviceProblem& p, const SolveLayout& L, const LossParams& lp, const double* wps, const randt_solver_options& o,
                             const double* poses0, double* poses_out, double* result, unsigned long long* bad, cudaStream_t s) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NS = NB * (NB + 1) / 2 + NB + 2;
  if (use_team(L.n_items)) {
    constexpr size_t smem_t = sizeof(SolveTeam<NS>);
    static std::atomic<unsigned long long> done_t{0ull};
    if (cudaError_t rc = k7_allow_smem(k7_team_kernel<VARIANT, LOSS, MANIFOLD>, smem_t, done_t)) return rc;
    const uint32_t grid_t = std::max(1u, std::min((uint32_t)kSmCount * 3u, L.n_items));
    k7_team_kernel<VARIANT, LOSS, MANIFOLD><<<grid_t, kTeam * 32, smem_t, s>>>(p, L, lp, wps, o, poses0, poses_out, result, bad);
    return cudaGetLastError();
  }
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
