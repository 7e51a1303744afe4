// K8 — all-pairs evaluation: every moving cell of a scan against every fixed cell of a submap (optionally only those within an
// L-infinity window of the transformed moving mean), reduced to the per-pose normal equations (sm_100a).
//
// The reference never forms this product: Matcher::addNDTFactor (R/src/ndt_registration/ndt_matcher.cpp:183-288) keeps the k nearest
// fixed cells per moving cell.  BASELINE.json's north_star words the cost as "every moving-scan NDT cell against every overlapping
// submap cell" and SURVEY §8a/§8d ask for that reading to be run and labelled next to the kNN one: it is the same functor
// (NDTFrameToMapIntensityFactorResidualSE2 and siblings, ceres_residuals.h:421-552), the same Barron corrector and the same J^T J / J^T r
// accumulation, over N_m x N_f pairs.  Its inputs are cache resident (2 k + 8 k cells = 0.5 MB), so this variant is bound by the
// fp64 pipe, not by HBM: bench.py reports it in pairs/s and fp64 flop/s.
//
// One thread owns one moving cell: its rotated mean / covariance (the `Moving` part of K3's closed form) stays in registers while the
// CTA walks a slab of fixed cells staged in shared memory as ready-made doubles (broadcast reads, no conversions in the loop), two
// fixed cells per iteration as independent instruction streams.  Per-CTA sums go to a partial record; the last CTA of a map pair
// (ticket) folds the partials in tile order, so the result is bitwise reproducible.
#include "k3_device.cuh"

namespace randt {

namespace {

constexpr int kApThreads = 128;   // moving cells per CTA
constexpr int kApSlab = 128;      // fixed cells per CTA

struct FixedD { double mx, my, mi, s00, s11, s22, b2, e2, f2; };

template <int VARIANT, int LOSS>
__global__ void __launch_bounds__(kApThreads) k8_allpairs_kernel(const float4* __restrict__ cells_f, const uint32_t* __restrict__ off_f,
                                                                 const float4* __restrict__ cells_m, const uint32_t* __restrict__ off_m,
                                                                 const double* __restrict__ poses, LossParams lp, double window,
                                                                 double* __restrict__ partials /*[B][tiles_m * tiles_f][kMaxAcc]*/,
                                                                 uint32_t* __restrict__ tickets /*[B], zero*/, uint32_t tiles_m, uint32_t tiles_f,
                                                                 double* __restrict__ out /*[B][24]*/, unsigned long long* __restrict__ bad_counter) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NP = VarTraits<VARIANT>::NP;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NJ = NH + NB;
  constexpr int NS = NJ + 2;
  __shared__ FixedD slab[kApSlab];
  __shared__ PoseConst kc_s;
  __shared__ LossConst lc_s;
  __shared__ double red[kApThreads / 32][NS + 2];
  __shared__ uint32_t s_ticket;
  const uint32_t b = blockIdx.y;
  const uint32_t tm = blockIdx.x / tiles_f, tf = blockIdx.x % tiles_f;
  const uint32_t m0 = off_m[b], nm = off_m[b + 1] - m0, f0 = off_f[b], nf = off_f[b + 1] - f0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    PoseConst k0; LossConst l0;
    make_pose_const<VARIANT>(poses + (size_t)b * NP, k0);
    make_loss_const(lp, lp.mu, l0);
    kc_s = k0; lc_s = l0;
  }
  // stage this CTA's slab of fixed cells as doubles (the couples already summed: K3's symmetric part)
  const uint32_t fb = tf * kApSlab, n_slab = fb < nf ? min((uint32_t)kApSlab, nf - fb) : 0u;
  for (uint32_t e = tid; e < n_slab; e += kApThreads) {
    RawCell c;
    const float4* q = cells_f + 3 * (size_t)(f0 + fb + e);
    c.a = __ldg(q); c.b = __ldg(q + 1); c.c = __ldg(q + 2);
    const CellC cc = cellc_from_raw(c);
    FixedD d; d.mx = cc.mx; d.my = cc.my; d.mi = cc.mi; d.s00 = cc.s00; d.s11 = cc.s11; d.s22 = cc.s22; d.b2 = cc.b2; d.e2 = cc.e2; d.f2 = cc.f2;
    slab[e] = d;
  }
  __syncthreads();
  const PoseConst& kc = kc_s; const LossConst& lc = lc_s;
  double acc[NS]; double max_dd = 0.0; uint32_t n_bad = 0, n_used = 0;
#pragma unroll
  for (int e = 0; e < NS; ++e) acc[e] = 0.0;
  const uint32_t im = tm * kApThreads + tid;
  if (im < nm && n_slab) {
    RawCell c;
    const float4* q = cells_m + 3 * (size_t)(m0 + im);
    c.a = __ldg(q); c.b = __ldg(q + 1); c.c = __ldg(q + 2);
    Moving mv;
    moving_part<VARIANT>(kc, cellc_from_raw(c), mv);
    const double px = mv.xr + kc.tx, py = mv.yr + kc.ty;      // transformed moving mean (window test)
    auto one = [&](const FixedD& fd) {
      if (window > 0.0 && (fabs(px - fd.mx) > window || fabs(py - fd.my) > window)) return;
      // fixed_part with the fixed cell already in double
      const double d0 = px - fd.mx, d1 = py - fd.my;
      const double B00 = mv.M00 + fd.s00, B11 = mv.M11 + fd.s11, B01 = fma(0.5, fd.b2, mv.M01);
      double q0, q1, q2 = 0.0, dd;
      if (VARIANT == 0 || VARIANT == 2) {
        const double d2 = mv.mi - fd.mi;
        const double B22 = mv.S22 + fd.s22, B02 = fma(0.5, fd.e2, mv.M02), B12 = fma(0.5, fd.f2, mv.M12);
        const double C00 = fma(B11, B22, -B12 * B12), C01 = fma(B02, B12, -B01 * B22), C02 = fma(B01, B12, -B02 * B11);
        const double C11 = fma(B00, B22, -B02 * B02), C12 = fma(B01, B02, -B00 * B12), C22 = fma(B00, B11, -B01 * B01);
        const double det = fma(B00, C00, fma(B01, C01, B02 * C02));
        const double idet = rcp_fast(det);
        q0 = fma(C00, d0, fma(C01, d1, C02 * d2)) * idet;
        q1 = fma(C01, d0, fma(C11, d1, C12 * d2)) * idet;
        q2 = fma(C02, d0, fma(C12, d1, C22 * d2)) * idet;
        dd = fma(d0, q0, fma(d1, q1, d2 * q2));
      } else {
        const double det = fma(B00, B11, -B01 * B01);
        const double idet = rcp_fast(det);
        q0 = fma(B11, d0, -B01 * d1) * idet;
        q1 = fma(B00, d1, -B01 * d0) * idet;
        dd = fma(d0, q0, d1 * q1);
      }
      double N[4];
      if (VARIANT == 1) {
        const double g0 = fma(kc.c, q0, kc.s * q1), g1 = fma(kc.c, q1, -kc.s * q0);
        const double Sq0 = fma(mv.S00, q0, mv.bh * q1), Sq1 = fma(mv.bh, q0, mv.S11 * q1);
        const double Sg0 = fma(mv.S00, g0, mv.bh * g1), Sg1 = fma(mv.bh, g0, mv.S11 * g1);
        N[0] = fma(q0, mv.mx, q1 * mv.my) - fma(g0, Sq0, g1 * Sq1);
        N[1] = fma(q1, mv.mx, -q0 * mv.my) - fma(q1, Sg0, -q0 * Sg1);
        N[2] = q0; N[3] = q1;
      } else {
        const double a0 = mv.xr - fma(mv.M00, q0, fma(mv.M01, q1, mv.M02 * q2));
        const double a1 = mv.yr - fma(mv.M01, q0, fma(mv.M11, q1, mv.M12 * q2));
        const double nt = fma(q1, a0, -q0 * a1);
        if (VARIANT == 0) { N[0] = nt; N[1] = q0; N[2] = q1; } else { N[0] = q0; N[1] = q1; N[2] = nt; }
      }
      if (!dd_valid(dd)) { ++n_bad; return; }
      double wgt, hrho, wd;
      loss_eval<LOSS>(dd, lc, wgt, hrho, wd);
      int qi = 0;
#pragma unroll
      for (int a = 0; a < NB; ++a) {
        const double wa = wd * N[a];
#pragma unroll
        for (int b2 = a; b2 < NB; ++b2) { acc[qi] = fma(wa, N[b2], acc[qi]); ++qi; }
        acc[NH + a] = fma(wgt, N[a], acc[NH + a]);
      }
      acc[NJ] += hrho; acc[NJ + 1] += dd;
      max_dd = fmax(max_dd, dd);
      ++n_used;
    };
    uint32_t j = 0;
    for (; j + 2 <= n_slab; j += 2) { one(slab[j]); one(slab[j + 1]); }
    if (j < n_slab) one(slab[j]);
  }
  // CTA reduction (fixed order: lanes by xor-shuffle, then warps in order)
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    double v = acc[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) red[warp][e] = v;
  }
  {
    const double mx = warp_max_nonneg(max_dd);
    const uint32_t bad = __reduce_add_sync(kFull, n_bad), used = __reduce_add_sync(kFull, n_used);
    if (lane == 0) { red[warp][NS] = mx; red[warp][NS + 1] = (double)used; if (bad) atomicAdd(bad_counter, (unsigned long long)bad); }
  }
  __syncthreads();
  const uint32_t n_tiles = tiles_m * tiles_f;
  double* part = partials + ((size_t)b * n_tiles + blockIdx.x) * kMaxAcc;
  if (tid < NS + 2) {
    double v = red[0][tid];
    for (int w2 = 1; w2 < kApThreads / 32; ++w2) v = (tid == NS) ? fmax(v, red[w2][tid]) : v + red[w2][tid];
    part[tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(&tickets[b], 1u);
  __syncthreads();
  if (s_ticket != n_tiles - 1u) return;
  // last CTA of this map pair: fold the partials in tile order and write the 24-double record
  __threadfence();
  if (warp == 0) {
    double v = 0.0;
    if (lane < NS + 2) {
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const double x = __ldcg(partials + ((size_t)b * n_tiles + t) * kMaxAcc + lane);
        v = (lane == NS) ? fmax(v, x) : v + x;
      }
    }
    const double mx_all = __shfl_sync(kFull, v, NS);
    const double used_all = __shfl_sync(kFull, v, NS + 1);
    const uint32_t omap = out_map<VARIANT, true>(lane);
    write_segment_out(v, mx_all, kc.ja, kc.jb, (uint32_t)used_all, out, b, 0u, lane, omap, 1);
    if (lane == 0) tickets[b] = 0u;
  }
}

int k8_loss_code(const LossParams& lp) {
  if (lp.kind == RANDT_LOSS_NONE) return L_NONE;
  if (lp.kind == RANDT_LOSS_WELSCH) return L_WELSCH;
  if (lp.alpha == -2.0) return L_BARRON_M2;
  if (lp.alpha == -1.0) return L_BARRON_M1;
  return L_BARRON;
}

template <int VARIANT>
cudaError_t launch_ap_v(dim3 grid, const float4* cf, const uint32_t* of, const float4* cm, const uint32_t* om, const double* poses, const LossParams& lp,
                        double window, double* partials, uint32_t* tickets, uint32_t tiles_m, uint32_t tiles_f, double* out, unsigned long long* bad,
                        cudaStream_t s) {
#define RANDT_AP(L) k8_allpairs_kernel<VARIANT, L><<<grid, kApThreads, 0, s>>>(cf, of, cm, om, poses, lp, window, partials, tickets, tiles_m, tiles_f, out, bad)
  switch (k8_loss_code(lp)) {
    case L_NONE: RANDT_AP(L_NONE); break;
    case L_WELSCH: RANDT_AP(L_WELSCH); break;
    case L_BARRON_M2: RANDT_AP(L_BARRON_M2); break;
    case L_BARRON_M1: RANDT_AP(L_BARRON_M1); break;
    default: RANDT_AP(L_BARRON); break;
  }
#undef RANDT_AP
  return cudaGetLastError();
}

}  // namespace

void allpairs_tiles(uint32_t max_m, uint32_t max_f, uint32_t* tiles_m, uint32_t* tiles_f) {
  *tiles_m = std::max(1u, (max_m + kApThreads - 1u) / kApThreads);
  *tiles_f = std::max(1u, (max_f + kApSlab - 1u) / kApSlab);
}

cudaError_t launch_allpairs(const float4* cells_f, const uint32_t* off_f, uint32_t max_f, const float4* cells_m, const uint32_t* off_m, uint32_t max_m,
                            uint32_t n_maps, int variant, const double* d_poses, const LossParams& lp, double window, double* d_partials,
                            uint32_t* d_tickets, double* d_out, unsigned long long* d_bad, cudaStream_t s, int* n_launches) {
  if (n_maps == 0) return cudaSuccess;
  uint32_t tm, tf;
  allpairs_tiles(max_m, max_f, &tm, &tf);
  dim3 grid(tm * tf, n_maps);
  cudaError_t e;
  switch (variant) {
    case 0: e = launch_ap_v<0>(grid, cells_f, off_f, cells_m, off_m, d_poses, lp, window, d_partials, d_tickets, tm, tf, d_out, d_bad, s); break;
    case 1: e = launch_ap_v<1>(grid, cells_f, off_f, cells_m, off_m, d_poses, lp, window, d_partials, d_tickets, tm, tf, d_out, d_bad, s); break;
    case 2: e = launch_ap_v<2>(grid, cells_f, off_f, cells_m, off_m, d_poses, lp, window, d_partials, d_tickets, tm, tf, d_out, d_bad, s); break;
    case 3: e = launch_ap_v<3>(grid, cells_f, off_f, cells_m, off_m, d_poses, lp, window, d_partials, d_tickets, tm, tf, d_out, d_bad, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (n_launches) *n_launches += 1;
  return e;
}

}  // namespace randt
