// K3 — NDT cell-pair residual / Jacobian evaluation and fused normal-equation reduction (sm_100a).
//
// Replaces hot loop C of the reference: for every residual block, ceres::AutoDiffCostFunction<F,1,np>::Evaluate of
//   NDTFrameToMapIntensityFactorResidualSE2  R/include/ndt_registration/ceres_residuals.h:520-552   (variant 0, live)
//   NDTFrameToMapFactorResidualSE2           :454-484 (1)   NDTFrameToMapIntensityFactorResidual :486-518 (2)
//   NDTFrameToMapFactorResidual              :421-451 (3)
// followed by the robust-loss Corrector (BarronLoss, R/src/ndt_registration/ceres_loss_functions.cpp:19-39, inside
// ceres::ScaledLoss, R/src/ndt_registration/ndt_matcher.cpp:392) and ceres' accumulation of J^T J / J^T r.
//
// Math.  The reference differentiates r = sqrt(d^T B^-1 d), d = R mu_m + t - mu_f, B = R S_m R^T + S_f with 4-wide dual
// numbers.  This kernel evaluates the closed form in fp64 registers (inputs are float32-born but cond(B) reaches ~1e5, so
// fp32 cannot hold the 1e-5 parity bound).  Only the symmetric part of B enters d^T B^-1 d and its derivative up to
// O((eps*cond)^2) (the reference's regularised covariances are asymmetric at float-ulp level, ndt_cell.cpp:110; the first-
// order term of the antisymmetric part cancels in the quadratic form), so B is symmetrised on load and the kernel works with
//   q = B^-1 d (symmetric cofactor inverse),  dd = d.q,  N = r dr/dp = (1/2) d(dd)/dp :
//   N_x = q0, N_y = q1, N_theta = q1 (xr - (Mq)0) - q0 (yr - (Mq)1),  M = R S_m R^T, (xr, yr) = R mu_m.
// FUSED mode never forms r or J: with w = weight*rho'(dd), H += (w/dd) N N^T, g += w N, cost += weight*rho(dd)/2, and for
// the shipped alpha = -2 loss w/dd and rho come from ONE reciprocal.  fp64 pipe: ~110 instructions per pair (was 193).
//
// Work decomposition (B200: 148 SMs, fp64 64 lanes/clk/SM — the co-limiter next to HBM).  Pairs are grouped into duos (two pairs
// that share their moving cell, one lane each) and the duos of a segment (= pose) into tiles; ONE WARP owns a tile.  The host assigns
// tiles to the 148 x 16 resident warps of a persistent grid (longest-processing-time first) and lays the duo records — moving cell
// and both fixed cells inlined in the compact 112-byte form of common.cuh — out in that order, so a warp streams ONE contiguous range: every chunk of 32 records is a
// single bulk copy (TMA, cp.async.bulk + mbarrier) into a two-stage shared-memory ring, described by a 16-byte chunk descriptor.
// Partial sums stay in registers; at a tile's end they are transposed through the just-consumed stage buffer and added in a fixed
// order (deterministic), and lanes e < 24 write the record.  Segments longer than one tile are folded by the last warp to finish
// (ticket counter), in tile order.  Launches chain with programmatic dependent launch (the schedule is fetched before the wait).
#include <math.h>

#include <atomic>

#include "k3_device.cuh"

namespace randt {

namespace {

template <int VARIANT, int LOSS, bool WANT_JAC>
__global__ void __launch_bounds__(kK3Threads, kMinCtas) k3_fused_kernel(DeviceProblem P, const double* __restrict__ poses, LossParams lp,
                                                                       const double* __restrict__ mu_per_seg, double* __restrict__ out,
                                                                       unsigned long long* __restrict__ bad_counter) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NP = VarTraits<VARIANT>::NP;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NJ = WANT_JAC ? NH + NB : 0;     // additive slots: H, g, then cost, sum dd
  constexpr int NS = NJ + 2;
  using StageBuf = StageBufT<NS * 34 * 8>;
  extern __shared__ __align__(128) unsigned char k3_dyn_smem[];       // [kWarpsPerCta][kStages] stage buffers
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  const uint32_t w = blockIdx.x * kWarpsPerCta + warp;
  if (w >= P.n_warps) return;
  StageBuf* stage = reinterpret_cast<StageBuf*>(k3_dyn_smem) + warp * kStages;
  const uint32_t omap = out_map<VARIANT, WANT_JAC>(lane);
  __shared__ ChunkDesc queue_all[kWarpsPerCta][64];
  WarpQueue wq;
  wq.q = queue_all[warp];
  wq.act[0] = 0u; wq.act[1] = 0u;
  griddep_launch_dependents();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&stage[s].bar, 1u);
    fence_mbar_init();
  }
  // Everything the previous kernel of the stream may have written (poses, mu, active flags, `out`, a re-planned schedule) is read
  // after griddep_wait(); the static schedule of a problem is immutable, so its first descriptors are fetched while the
  // previous grid is still draining.
  if (!P.plan_static) griddep_wait();
  wq.c_begin = P.warp_off[w]; wq.c_end = P.warp_off[w + 1];
  const uint4 v_first = wq.load_desc(P, 0u, lane);
  if (P.plan_static) griddep_wait();
  if (wq.c_begin >= wq.c_end) return;
  wq.publish(P, v_first, 0u, lane);    // includes the warp barrier that publishes the mbarrier init
  // prologue: chunks 0 .. kStages-2 in flight (a warp owns at least one chunk; kStages - 1 <= 31 descriptors are in the queue)
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) stage_issue<NP, true>(P, wq, (uint32_t)s, lane, &stage[s], poses, mu_per_seg);
  uint32_t phase_bits = 0u;     // bit s: parity the next completion of stage s's barrier will have

  // Pose and loss constants of the tile(s) in flight live in shared memory (two slots per warp: a split chunk carries the tail of
  // one tile and the head of the next; slot `ts` belongs to the tile at lane 0), read by (mostly broadcast) LDS where they are used:
  // ~26 fewer live registers per thread.
  __shared__ PoseConst kc_all[kWarpsPerCta][2];
  __shared__ LossConst lc_all[kWarpsPerCta][2];
  __shared__ uint32_t np_all[kWarpsPerCta][2];     // pairs of the segment(s) in flight
  int ts = 0;
  double acc[NS]; double max_dd = 0.0; uint32_t n_bad = 0;
#pragma unroll
  for (int e = 0; e < NS; ++e) acc[e] = 0.0;
  int slot = 0;
  uint32_t jj = 0;              // index of the chunk being consumed in the warp's list
  while (true) {
    const ChunkDesc cm = wq.get(jj);
    const uint32_t n_here = cm.meta & kChunkCountMask;
    if (n_here == 0u) break;                                  // past the warp's last chunk
    const bool live = wq.live(jj);                            // false: chunk of an inactive segment (nothing was copied)
    // ---- stage chunk j + kStages - 1 ----
    const uint32_t jn = jj + (uint32_t)(kStages - 1);
    if ((jn & 31u) == 0u) wq.refill(P, jn >> 5, lane);        // first chunk of a new block of 32 descriptors
    int islot = slot + kStages - 1; if (islot >= kStages) islot -= kStages;
    stage_issue<NP, true>(P, wq, jn, lane, &stage[islot], poses, mu_per_seg);
    // ---- chunk j has landed ----
    cp_async_wait<kStages - 1>();
    __syncwarp();
    StageBuf* sb = &stage[slot];
    if (live) { mbar_wait(&sb->bar, (phase_bits >> slot) & 1u); phase_bits ^= 1u << slot; }
    const bool split = (cm.meta & kChunkSplit) != 0u;         // lanes >= sp belong to the next tile (segment cm.part)
    const uint32_t sp = split ? ((cm.meta >> kChunkSplitShift) & 63u) : n_here;
    if (live && (cm.meta & (kChunkFirst | kChunkSplit))) {
      // lane 0: constants of a tile starting at lane 0 -> slot ts; lane 1: constants of the tile starting at lane sp -> slot ts ^ 1
      if (lane < 2 && (cm.meta & (lane == 0 ? kChunkFirst : kChunkSplit))) {
        PoseConst k0; LossConst l0;
        make_pose_const<VARIANT>(sb->pose[lane], k0);
        make_loss_const(lp, mu_per_seg ? sb->mu[lane] : lp.mu, l0);
        kc_all[warp][ts ^ lane] = k0; lc_all[warp][ts ^ lane] = l0;
        np_all[warp][ts ^ lane] = sb->segoff[lane][1] - sb->segoff[lane][0];
      }
      __syncwarp();
    }
    double* scratch = reinterpret_cast<double*>(&sb->rec[0]);
    if (!split) {
      const PoseConst& kc = kc_all[warp][ts]; const LossConst& lc = lc_all[warp][ts];
      if (live && (uint32_t)lane < n_here) accumulate_duo<VARIANT, LOSS, WANT_JAC, NS>(kc, lc, sb->rec, P.duo_overflow, lane, acc, max_dd, n_bad);
      if (live && (cm.meta & kChunkLast)) {
        finish_tile<VARIANT, WANT_JAC, NS>(P, acc, max_dd, n_bad, cm.seg, (cm.meta & kChunkSolo) != 0u, cm.part, kc, np_all[warp][ts], omap, scratch, out, bad_counter, lane);
#pragma unroll
        for (int e = 0; e < NS; ++e) acc[e] = 0.0;
        max_dd = 0.0; n_bad = 0;
      }
    } else {
      // the tail of one (solo) tile in lanes [0, sp) and the head of the next in lanes [sp, n_here): every lane evaluates its duo
      // with its own tile's constants into a private set of sums, which then joins the right tile's totals
      const bool mine_new = (uint32_t)lane >= sp;
      const int my = mine_new ? (ts ^ 1) : ts;
      double c[NS]; double mc = 0.0; uint32_t cb = 0;
#pragma unroll
      for (int e = 0; e < NS; ++e) c[e] = 0.0;
      if ((uint32_t)lane < n_here) accumulate_duo<VARIANT, LOSS, WANT_JAC, NS>(kc_all[warp][my], lc_all[warp][my], sb->rec, P.duo_overflow, lane, c, mc, cb);
      // lanes of the finishing tile add their duo to its sums, lanes of the starting tile keep theirs for the next: 0/1 masks instead of
      // per-accumulator selects (the private sums are finite: degenerate pairs never enter them)
      const double m_old = mine_new ? 0.0 : 1.0, m_new = mine_new ? 1.0 : 0.0;
      double vals[NS];
#pragma unroll
      for (int e = 0; e < NS; ++e) vals[e] = fma(c[e], m_old, acc[e]);
      finish_tile<VARIANT, WANT_JAC, NS>(P, vals, fmax(max_dd, mine_new ? 0.0 : mc), n_bad + (mine_new ? 0u : cb), cm.seg, true, 0u, kc_all[warp][ts],
                                         np_all[warp][ts], omap, scratch, out, bad_counter, lane);
#pragma unroll
      for (int e = 0; e < NS; ++e) acc[e] = fma(c[e], m_new, 0.0);      // (+0 addend: a negative private sum times 0 must not leave -0)
      max_dd = mine_new ? mc : 0.0; n_bad = mine_new ? cb : 0u;
      if (cm.meta & kChunkNewLast) {      // the second tile is shorter than the rest of the chunk: it ends here too
        finish_tile<VARIANT, WANT_JAC, NS>(P, acc, max_dd, n_bad, cm.part, true, 0u, kc_all[warp][ts ^ 1], np_all[warp][ts ^ 1], omap, scratch, out, bad_counter, lane);
#pragma unroll
        for (int e = 0; e < NS; ++e) acc[e] = 0.0;
        max_dd = 0.0; n_bad = 0;
      }
      ts ^= 1;                            // the tile that started here is the one at lane 0 of the next chunk
    }
    // Every lane's LDS of this slot has completed (the values were consumed above), so after the warp barrier lane 0 may let
    // the next bulk copy overwrite it.
    __syncwarp();
    if (++slot == kStages) slot = 0;
    ++jj;
  }
  cp_async_wait<0>();
}

// EMIT: raw residual and ambient Jacobian row per pair (what Evaluate returns for each block); same tile stream, no reduction
template <int VARIANT, bool WANT_JAC>
__device__ __forceinline__ void emit_one(const PoseConst& kc, uint32_t i, double dd, const double* N, double* __restrict__ r_out,
                                         double* __restrict__ J_out, uint32_t& n_bad) {
  const bool ok = dd_valid(dd);
  double r = 0.0, rs = 0.0;   // r = 0: the reference's dual-number sqrt yields 0/0 here; defined as J = 0
  if (ok && dd > 0.0) { rs = rsqrt_fast(dd); r = dd * rs; }
  if (!ok) { ++n_bad; r = __longlong_as_double(0x7ff8000000000000ll); }
  r_out[i] = r;
  if (WANT_JAC) {
    if (VARIANT == 0) {
      double2* dst = reinterpret_cast<double2*>(J_out + (size_t)i * 4);
      const double jt = N[0] * rs;
      dst[0] = make_double2(jt * kc.ja, jt * kc.jb);
      dst[1] = make_double2(N[1] * rs, N[2] * rs);
    } else if (VARIANT == 1) {
      double2* dst = reinterpret_cast<double2*>(J_out + (size_t)i * 4);
      dst[0] = make_double2(N[0] * rs, N[1] * rs);
      dst[1] = make_double2(N[2] * rs, N[3] * rs);
    } else {
      J_out[(size_t)i * 3 + 0] = N[0] * rs; J_out[(size_t)i * 3 + 1] = N[1] * rs; J_out[(size_t)i * 3 + 2] = N[2] * rs;
    }
  }
}

template <int VARIANT, bool WANT_JAC>
__global__ void __launch_bounds__(kK3Threads, kMinCtas) k3_emit_kernel(DeviceProblem P, const double* __restrict__ poses, double* __restrict__ r_out,
                                                                      double* __restrict__ J_out, unsigned long long* __restrict__ bad_counter) {
  constexpr int NP = VarTraits<VARIANT>::NP;
  using StageBuf = StageBufT<0>;
  extern __shared__ __align__(128) unsigned char k3_dyn_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  const uint32_t w = blockIdx.x * kWarpsPerCta + warp;
  if (w >= P.n_warps) return;
  StageBuf* stage = reinterpret_cast<StageBuf*>(k3_dyn_smem) + warp * kStages;
  __shared__ ChunkDesc queue_all[kWarpsPerCta][64];
  WarpQueue wq;
  wq.q = queue_all[warp];
  wq.act[0] = 0u; wq.act[1] = 0u;
  wq.c_begin = P.warp_off[w]; wq.c_end = P.warp_off[w + 1];
  if (wq.c_begin >= wq.c_end) return;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&stage[s].bar, 1u);
    fence_mbar_init();
  }
  wq.refill(P, 0u, lane);
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) stage_issue<NP, false>(P, wq, (uint32_t)s, lane, &stage[s], poses, nullptr);
  PoseConst kc;
  uint32_t n_bad = 0, phase_bits = 0u;
  int slot = 0;
  uint32_t jj = 0;
  while (true) {
    const ChunkDesc cm = wq.get(jj);
    const uint32_t n_here = cm.meta & kChunkCountMask;
    if (n_here == 0u) break;
    const bool live = wq.live(jj);
    const uint32_t jn = jj + (uint32_t)(kStages - 1);
    if ((jn & 31u) == 0u) wq.refill(P, jn >> 5, lane);
    int islot = slot + kStages - 1; if (islot >= kStages) islot -= kStages;
    stage_issue<NP, false>(P, wq, jn, lane, &stage[islot], poses, nullptr);
    cp_async_wait<kStages - 1>();
    __syncwarp();
    StageBuf* sb = &stage[slot];
    if (live) { mbar_wait(&sb->bar, (phase_bits >> slot) & 1u); phase_bits ^= 1u << slot; }
    if ((cm.meta & kChunkFirst) && live) make_pose_const<VARIANT>(sb->pose[0], kc);
    if (live && (uint32_t)lane < n_here) {
      CellC m, f0, f1;
      bool two;
      load_duo(sb->rec, lane, P.duo_overflow, m, f0, f1, two);
      const uint32_t p0 = __ldg(P.duo_p0 + cm.duo_begin + lane);     // first pair of the duo: where its rows go in r / J
      Moving mv;
      moving_part<VARIANT>(kc, m, mv);
      double N[4] = {0.0, 0.0, 0.0, 0.0};
      const double dd0 = fixed_part<VARIANT, WANT_JAC>(kc, mv, f0, N);
      emit_one<VARIANT, WANT_JAC>(kc, p0, dd0, N, r_out, J_out, n_bad);
      if (two) {
        const double dd1 = fixed_part<VARIANT, WANT_JAC>(kc, mv, f1, N);
        emit_one<VARIANT, WANT_JAC>(kc, p0 + 1u, dd1, N, r_out, J_out, n_bad);
      }
    }
    __syncwarp();
    if (++slot == kStages) slot = 0;
    ++jj;
  }
  cp_async_wait<0>();
  const uint32_t bad = __reduce_add_sync(kFull, n_bad);
  if (lane == 0 && bad) atomicAdd(bad_counter, (unsigned long long)bad);
}

// SWEEP: one thread per candidate pose, pairs of one segment staged through shared memory in chunks (broadcast reads).
constexpr int kSweepThreads = 128;
constexpr int kSweepChunk = 128;   // pairs staged per iteration: 128 * 2 cells * 48 B = 12 KB
template <int VARIANT, int LOSS>
__global__ void __launch_bounds__(kSweepThreads) k3_sweep_kernel(DeviceProblem P, uint32_t pair_begin, uint32_t pair_end,
                                                                const double* __restrict__ poses, uint32_t n_poses, LossParams lp,
                                                                double* __restrict__ cost_out) {
  constexpr int NP = VarTraits<VARIANT>::NP;
  __shared__ float4 sm_m[kSweepChunk][3];
  __shared__ float4 sm_f[kSweepChunk][3];
  const uint32_t pi = blockIdx.x * kSweepThreads + threadIdx.x;
  PoseConst k; LossConst lc;
  const bool active = pi < n_poses;
  if (active) make_pose_const<VARIANT>(poses + (size_t)pi * NP, k);
  else { const double idp[4] = {1, 0, 0, 0}; make_pose_const<VARIANT>(idp, k); }
  make_loss_const(lp, lp.mu, lc);
  double cost = 0.0;
  for (uint32_t base = pair_begin; base < pair_end; base += kSweepChunk) {
    const uint32_t n = min((uint32_t)kSweepChunk, pair_end - base);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n * 6; e += kSweepThreads) {
      const uint32_t pp = e / 6, w = e % 6;
      const uint2 pr = P.pairs[base + pp];
      if (w < 3) sm_m[pp][w] = __ldg(P.cells_m + 3 * (size_t)pr.x + w);
      else       sm_f[pp][w - 3] = __ldg(P.cells_f + 3 * (size_t)pr.y + (w - 3));
    }
    __syncthreads();
    if (active) {
      for (uint32_t pp = 0; pp < n; ++pp) {
        RawCell m, f;
        m.a = sm_m[pp][0]; m.b = sm_m[pp][1]; m.c = sm_m[pp][2];
        f.a = sm_f[pp][0]; f.b = sm_f[pp][1]; f.c = sm_f[pp][2];
        double N[4];
        const double dd = pair_core<VARIANT, false>(k, m, f, N);
        if (!dd_valid(dd)) continue;
        double wgt, hrho, wd;
        loss_eval<LOSS>(dd, lc, wgt, hrho, wd);
        cost += hrho;
      }
    }
  }
  if (active) cost_out[pi] = cost;
}

// Record-major duo table in the compact form (common.cuh: DuoRec): record d = [moving cell | fixed cell of pair p0 | fixed cell of pair
// p0 + 1] as 3 x 9 floats + the code word, 112 B, so that the 32 duos of a chunk are one contiguous 3584-byte block that a single bulk
// copy lands.  One thread per duo; a duo whose couples do not fit the encoding goes to the overflow table as stored.
__global__ void __launch_bounds__(128) build_duo_records_kernel(const float4* __restrict__ cells_m, const float4* __restrict__ cells_f,
                                                                const Duo* __restrict__ duos, uint32_t n_duos, const uint32_t* __restrict__ tile_rec_begin,
                                                                const uint32_t* __restrict__ tile_duo_begin, uint32_t n_tiles, DuoRec* __restrict__ recs,
                                                                uint32_t* __restrict__ duo_p0, DuoRecFull* __restrict__ overflow, uint32_t overflow_cap,
                                                                uint32_t* __restrict__ n_overflow) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_duos) return;
  // record d belongs to the tile t with tile_rec_begin[t] <= d < tile_rec_begin[t + 1] (tiles in the order the warps walk them) and
  // is that tile's (d - tile_rec_begin[t])-th duo; without a tile table the records are in duo order
  uint32_t src = d;
  if (tile_rec_begin) {
    uint32_t lo = 0, hi = n_tiles;          // largest t with tile_rec_begin[t] <= d
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (tile_rec_begin[mid] <= d) lo = mid; else hi = mid; }
    src = tile_duo_begin[lo] + (d - tile_rec_begin[lo]);
  }
  const Duo du = duos[src];
  const bool two = du.jf1 != kNoCell;
  RawCell c[3];
  c[0].a = __ldg(cells_m + 3 * (size_t)du.im); c[0].b = __ldg(cells_m + 3 * (size_t)du.im + 1); c[0].c = __ldg(cells_m + 3 * (size_t)du.im + 2);
  c[1].a = __ldg(cells_f + 3 * (size_t)du.jf0); c[1].b = __ldg(cells_f + 3 * (size_t)du.jf0 + 1); c[1].c = __ldg(cells_f + 3 * (size_t)du.jf0 + 2);
  c[2] = c[1];                                  // a missing second pair repeats the first (finite values; kRecNoSecond keeps it out of the sums)
  if (two) { c[2].a = __ldg(cells_f + 3 * (size_t)du.jf1); c[2].b = __ldg(cells_f + 3 * (size_t)du.jf1 + 1); c[2].c = __ldg(cells_f + 3 * (size_t)du.jf1 + 2); }
  encode_duo_record(c, two, recs + d, overflow, overflow_cap, n_overflow);
  duo_p0[d] = du.p0;
}

// the chunk lists of the two plans, one thread per warp of the schedule (schedule.hpp walk_warp: the host only computed the assignment)
__global__ void __launch_bounds__(128) emit_chunks_kernel(const uint32_t* __restrict__ tile_seg, const uint32_t* __restrict__ first,
                                                          const uint32_t* __restrict__ duo_off, uint32_t tile_duos, const uint32_t* __restrict__ mine,
                                                          const uint32_t* __restrict__ mine_off, const uint32_t* __restrict__ warp_rec_begin,
                                                          const uint32_t* __restrict__ woffA, const uint32_t* __restrict__ woffB, uint32_t n_warps,
                                                          ChunkDesc* __restrict__ planA, ChunkDesc* __restrict__ planB) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_warps) return;
  ChunkDesc* a = planA + woffA[w];
  ChunkDesc* b = planB + woffB[w];
  uint32_t nA = 0, nB = 0;
  walk_warp([=](uint32_t t) { return tile_from_seg(t, tile_seg[t], first, duo_off, tile_duos); }, first, mine, mine_off[w], mine_off[w + 1], warp_rec_begin[w],
            [a](uint32_t i, const ChunkDesc& c) { a[i] = c; }, [b](uint32_t i, const ChunkDesc& c) { b[i] = c; }, [](uint32_t, uint32_t, uint32_t) {}, &nA, &nB);
}

int loss_code(const LossParams& lp) {
  if (lp.kind == RANDT_LOSS_NONE) return L_NONE;
  if (lp.kind == RANDT_LOSS_WELSCH) return L_WELSCH;
  if (lp.alpha == -2.0) return L_BARRON_M2;
  if (lp.alpha == -1.0) return L_BARRON_M1;
  return L_BARRON;
}

inline int stream_grid(uint32_t n_warps) { return (int)((n_warps + kWarpsPerCta - 1) / kWarpsPerCta); }

// the stage ring lives in dynamic shared memory (three stages of four warps exceed the 48 KB static limit)
// (the attribute is per device: `done` remembers, per kernel instantiation, on which devices it has been set)
template <typename K>
cudaError_t allow_smem(K kernel, size_t bytes, std::atomic<unsigned long long>& done) {
  if (bytes <= 32u * 1024u) return cudaSuccess;      // static + dynamic stay below the 48 KB that need no opt-in
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}
template <int VARIANT, int LOSS, bool WANT_JAC>
cudaError_t launch_fused_vlj(const DeviceProblem& p, const double* d_poses, const LossParams& lp, const double* d_mu, double* d_out,
                             unsigned long long* bad, cudaStream_t s) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NS = (WANT_JAC ? NB * (NB + 1) / 2 + NB : 0) + 2;
  constexpr size_t smem = (size_t)kWarpsPerCta * kStages * sizeof(StageBufT<NS * 34 * 8>);
  static std::atomic<unsigned long long> smem_done{0ull};
  if (cudaError_t rc = allow_smem(k3_fused_kernel<VARIANT, LOSS, WANT_JAC>, smem, smem_done)) return rc;
  const int grid = stream_grid(p.n_warps);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kK3Threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k3_fused_kernel<VARIANT, LOSS, WANT_JAC>, p, d_poses, lp, d_mu, d_out, bad);
}
template <int VARIANT, int LOSS>
cudaError_t launch_fused_vl(const DeviceProblem& p, const double* d_poses, const LossParams& lp, const double* d_mu, bool want_jac,
                            double* d_out, unsigned long long* bad, cudaStream_t s) {
  if (want_jac) return launch_fused_vlj<VARIANT, LOSS, true>(p, d_poses, lp, d_mu, d_out, bad, s);
  return launch_fused_vlj<VARIANT, LOSS, false>(p, d_poses, lp, d_mu, d_out, bad, s);
}
template <int VARIANT>
cudaError_t launch_fused_v(const DeviceProblem& p, const double* d_poses, const LossParams& lp, const double* d_mu, bool want_jac,
                           double* d_out, unsigned long long* bad, cudaStream_t s) {
  switch (loss_code(lp)) {
    case L_NONE: return launch_fused_vl<VARIANT, L_NONE>(p, d_poses, lp, d_mu, want_jac, d_out, bad, s);
    case L_WELSCH: return launch_fused_vl<VARIANT, L_WELSCH>(p, d_poses, lp, d_mu, want_jac, d_out, bad, s);
    case L_BARRON_M2: return launch_fused_vl<VARIANT, L_BARRON_M2>(p, d_poses, lp, d_mu, want_jac, d_out, bad, s);
    case L_BARRON_M1: return launch_fused_vl<VARIANT, L_BARRON_M1>(p, d_poses, lp, d_mu, want_jac, d_out, bad, s);
    default: return launch_fused_vl<VARIANT, L_BARRON>(p, d_poses, lp, d_mu, want_jac, d_out, bad, s);
  }
}

template <int VARIANT>
cudaError_t launch_sweep_v(const DeviceProblem& p, uint32_t pb, uint32_t pe, const double* d_poses, uint32_t n_poses, const LossParams& lp,
                           double* d_cost, cudaStream_t s) {
  const int grid = (int)((n_poses + kSweepThreads - 1) / kSweepThreads);
  switch (loss_code(lp)) {
    case L_NONE: k3_sweep_kernel<VARIANT, L_NONE><<<grid, kSweepThreads, 0, s>>>(p, pb, pe, d_poses, n_poses, lp, d_cost); break;
    case L_WELSCH: k3_sweep_kernel<VARIANT, L_WELSCH><<<grid, kSweepThreads, 0, s>>>(p, pb, pe, d_poses, n_poses, lp, d_cost); break;
    case L_BARRON_M2: k3_sweep_kernel<VARIANT, L_BARRON_M2><<<grid, kSweepThreads, 0, s>>>(p, pb, pe, d_poses, n_poses, lp, d_cost); break;
    case L_BARRON_M1: k3_sweep_kernel<VARIANT, L_BARRON_M1><<<grid, kSweepThreads, 0, s>>>(p, pb, pe, d_poses, n_poses, lp, d_cost); break;
    default: k3_sweep_kernel<VARIANT, L_BARRON><<<grid, kSweepThreads, 0, s>>>(p, pb, pe, d_poses, n_poses, lp, d_cost); break;
  }
  return cudaGetLastError();
}

}  // namespace

// duos in schedule order: record r belongs to the tile t with tile_rec_begin[t] <= r < tile_rec_begin[t+1] (tiles in the order the
// warps walk them) and is that tile's (r - tile_rec_begin[t])-th duo
__global__ void __launch_bounds__(256) permute_duos_kernel(const Duo* __restrict__ in, const uint32_t* __restrict__ tile_rec_begin,
                                                           const uint32_t* __restrict__ tile_duo_begin, uint32_t n_tiles, uint32_t n_duos,
                                                           Duo* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_duos) return;
  uint32_t lo = 0, hi = n_tiles;          // largest t with tile_rec_begin[t] <= r
  while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (tile_rec_begin[mid] <= r) lo = mid; else hi = mid; }
  out[r] = in[tile_duo_begin[lo] + (r - tile_rec_begin[lo])];
}
cudaError_t launch_permute_duos(const Duo* in, const uint32_t* tile_rec_begin, const uint32_t* tile_duo_begin, uint32_t n_tiles, uint32_t n_duos,
                                Duo* out, cudaStream_t s, int* n_launches) {
  if (n_duos == 0) return cudaSuccess;
  permute_duos_kernel<<<(n_duos + 255u) / 256u, 256, 0, s>>>(in, tile_rec_begin, tile_duo_begin, n_tiles, n_duos, out);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_build_duo_records(const float4* cells_m, const float4* cells_f, const Duo* duos, uint32_t n_duos, const uint32_t* tile_rec_begin,
                                     const uint32_t* tile_duo_begin, uint32_t n_tiles, DuoRec* recs, uint32_t* duo_p0, DuoRecFull* overflow,
                                     uint32_t overflow_cap, uint32_t* d_n_overflow, cudaStream_t s, int* n_launches) {
  if (n_duos == 0) return cudaSuccess;
  build_duo_records_kernel<<<(n_duos + 127u) / 128u, 128, 0, s>>>(cells_m, cells_f, duos, n_duos, tile_rec_begin, tile_duo_begin, n_tiles, recs, duo_p0, overflow,
                                                                  overflow_cap, d_n_overflow);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_emit_chunks(const uint32_t* tile_seg, const uint32_t* first, const uint32_t* duo_off, uint32_t tile_duos, const uint32_t* mine,
                               const uint32_t* mine_off, const uint32_t* warp_rec_begin, const uint32_t* woffA, const uint32_t* woffB, uint32_t n_warps,
                               ChunkDesc* planA, ChunkDesc* planB, cudaStream_t s, int* n_launches) {
  if (n_warps == 0) return cudaSuccess;
  emit_chunks_kernel<<<(n_warps + 127u) / 128u, 128, 0, s>>>(tile_seg, first, duo_off, tile_duos, mine, mine_off, warp_rec_begin, woffA, woffB, n_warps, planA, planB);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_eval_fused(const DeviceProblem& p, int variant, const double* d_poses, const LossParams& lp, const double* d_mu,
                              bool want_jac, double* d_out, unsigned long long* d_bad, cudaStream_t s, int* n_launches) {
  if (p.n_chunks == 0) return cudaSuccess;
  cudaError_t e;
  switch (variant) {
    case 0: e = launch_fused_v<0>(p, d_poses, lp, d_mu, want_jac, d_out, d_bad, s); break;
    case 1: e = launch_fused_v<1>(p, d_poses, lp, d_mu, want_jac, d_out, d_bad, s); break;
    case 2: e = launch_fused_v<2>(p, d_poses, lp, d_mu, want_jac, d_out, d_bad, s); break;
    case 3: e = launch_fused_v<3>(p, d_poses, lp, d_mu, want_jac, d_out, d_bad, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (n_launches) *n_launches += 1;
  return e;
}

cudaError_t launch_eval_emit(const DeviceProblem& p, int variant, const double* d_poses, double* d_r, double* d_J,
                             unsigned long long* d_bad, cudaStream_t s, int* n_launches) {
  if (p.n_chunks == 0) return cudaSuccess;
  const int grid = stream_grid(p.n_warps);
  constexpr size_t smem = (size_t)kWarpsPerCta * kStages * sizeof(StageBufT<0>);
#define RANDT_EMIT(V)                                                                                              \
  if (d_J) { static std::atomic<unsigned long long> dn_{0ull}; if (cudaError_t rc_ = allow_smem(k3_emit_kernel<V, true>, smem, dn_)) return rc_;   \
             k3_emit_kernel<V, true><<<grid, kK3Threads, smem, s>>>(p, d_poses, d_r, d_J, d_bad); }             \
  else     { static std::atomic<unsigned long long> dn_{0ull}; if (cudaError_t rc_ = allow_smem(k3_emit_kernel<V, false>, smem, dn_)) return rc_;  \
             k3_emit_kernel<V, false><<<grid, kK3Threads, smem, s>>>(p, d_poses, d_r, d_J, d_bad); }
  switch (variant) {
    case 0: RANDT_EMIT(0) break;
    case 1: RANDT_EMIT(1) break;
    case 2: RANDT_EMIT(2) break;
    case 3: RANDT_EMIT(3) break;
    default: return cudaErrorInvalidValue;
  }
#undef RANDT_EMIT
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_sweep_costs(const DeviceProblem& p, uint32_t pair_begin, uint32_t pair_end, int variant, const double* d_poses,
                               uint32_t n_poses, const LossParams& lp, double* d_cost, cudaStream_t s, int* n_launches) {
  if (n_poses == 0) return cudaSuccess;
  cudaError_t e;
  switch (variant) {
    case 0: e = launch_sweep_v<0>(p, pair_begin, pair_end, d_poses, n_poses, lp, d_cost, s); break;
    case 1: e = launch_sweep_v<1>(p, pair_begin, pair_end, d_poses, n_poses, lp, d_cost, s); break;
    case 2: e = launch_sweep_v<2>(p, pair_begin, pair_end, d_poses, n_poses, lp, d_cost, s); break;
    case 3: e = launch_sweep_v<3>(p, pair_begin, pair_end, d_poses, n_poses, lp, d_cost, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (n_launches) *n_launches += 1;
  return e;
}

}  // namespace randt
