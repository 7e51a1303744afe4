// K3 — NDT cell-pair residual / Jacobian evaluation and fused normal-equation reduction (sm_100a).
//
// Replaces hot loop C of the reference: for every residual block, ceres::AutoDiffCostFunction<F,1,np>::Evaluate of
//   NDTFrameToMapIntensityFactorResidualSE2  R/include/ndt_registration/ceres_residuals.h:520-552   (variant 0, live)
//   NDTFrameToMapFactorResidualSE2           :454-484 (1)   NDTFrameToMapIntensityFactorResidual :486-518 (2)
//   NDTFrameToMapFactorResidual              :421-451 (3)
// followed by the robust-loss Corrector (BarronLoss, R/src/ndt_registration/ceres_loss_functions.cpp:19-39, inside
// ceres::ScaledLoss, R/src/ndt_registration/ndt_matcher.cpp:392) and ceres' accumulation of J^T J / J^T r.
//
// The reference differentiates with 4-wide dual numbers; this kernel evaluates the closed form
//   r = sqrt(d^T B^-1 d),  d = R mu_m + t - mu_f,  B = R S_m R^T + S_f,
//   dr = [ (q+p)^T dd - p^T dB q ] / (2 r),   q = B^-1 d,  p = B^-T d     (B is not assumed symmetric: the reference's
// regularised covariances are asymmetric at float-ulp level, R/src/ndt_representation/ndt_cell.cpp:110)
// in fp64 registers: inputs are float32-born but cond(B) reaches ~1e5, so fp32 cannot hold the 1e-5 parity bound.
//
// Work decomposition: the frozen pair list is cut into tiles of <= 512 pairs that never straddle a segment (= pose).
// One 128-thread CTA per tile; threads stride the tile (coalesced uint2 pair loads, 3 x float4 gathers per cell, the
// k pairs of one moving cell sit in adjacent lanes).  FUSED mode keeps 13/18 fp64 accumulators per thread, reduces
// them through shared memory in a fixed order (deterministic), and the last CTA of a segment folds the tile partials.
#include <math.h>

#include "common.cuh"

namespace randt {

namespace {

enum { L_NONE = 0, L_BARRON = 1, L_WELSCH = 2, L_BARRON_M2 = 3, L_BARRON_M1 = 4 };

struct SegConst {   // per-tile constants, computed by one thread, broadcast through shared memory
  double c, s, tx, ty;      // rotation entries actually used by the variant (normalised for 0/2/3, raw for 1)
  double ja, jb;            // variant 0: dtheta/dc, dtheta/ds
  // loss
  double lb, lc, pre, ts, e, weight;
};

__device__ __forceinline__ void load_cell(const float4* __restrict__ tab, uint32_t idx, double mu[3], double cov[9]) {
  const float4 a = __ldg(tab + 3 * (size_t)idx), b = __ldg(tab + 3 * (size_t)idx + 1), c = __ldg(tab + 3 * (size_t)idx + 2);
  mu[0] = a.x; mu[1] = a.y; mu[2] = a.z;
  cov[0] = a.w; cov[1] = b.x; cov[2] = b.y; cov[3] = b.z; cov[4] = b.w; cov[5] = c.x; cov[6] = c.y; cov[7] = c.z; cov[8] = c.w;
}

// ---- 3-D core (x, y, intensity); rotation about the intensity axis --------------------------------------------
// returns dd = d^T B^-1 d; if WANT_JAC: un-normalised derivative numerators (multiply by 1/(2r)): nt (theta), nx, ny
template <bool WANT_JAC>
__device__ __forceinline__ double core3(double ct, double st, double tx, double ty, const double mm[3], const double S[9],
                                        const double fm[3], const double F[9], double& nt, double& nx, double& ny) {
  const double xr = ct * mm[0] - st * mm[1];
  const double yr = st * mm[0] + ct * mm[1];
  const double d0 = xr + tx - fm[0], d1 = yr + ty - fm[1], d2 = mm[2] - fm[2];
  const double T00 = ct * S[0] - st * S[3], T01 = ct * S[1] - st * S[4];
  const double T10 = st * S[0] + ct * S[3], T11 = st * S[1] + ct * S[4];
  const double M00 = T00 * ct - T01 * st, M01 = T00 * st + T01 * ct;
  const double M10 = T10 * ct - T11 * st, M11 = T10 * st + T11 * ct;
  const double M02 = ct * S[2] - st * S[5], M12 = st * S[2] + ct * S[5];
  const double M20 = S[6] * ct - S[7] * st, M21 = S[6] * st + S[7] * ct, M22 = S[8];
  const double B00 = M00 + F[0], B01 = M01 + F[1], B02 = M02 + F[2];
  const double B10 = M10 + F[3], B11 = M11 + F[4], B12 = M12 + F[5];
  const double B20 = M20 + F[6], B21 = M21 + F[7], B22 = M22 + F[8];
  const double C00 = B11 * B22 - B12 * B21, C01 = B12 * B20 - B10 * B22, C02 = B10 * B21 - B11 * B20;
  const double C10 = B21 * B02 - B22 * B01, C11 = B22 * B00 - B20 * B02, C12 = B20 * B01 - B21 * B00;
  const double C20 = B01 * B12 - B02 * B11, C21 = B02 * B10 - B00 * B12, C22 = B00 * B11 - B01 * B10;
  const double det = C00 * B00 + C10 * B10 + C20 * B20;
  const double idet = 1.0 / det;
  // q = B^-1 d : inv[i][j] = C[j][i] / det ;  p = B^-T d
  const double q0 = (C00 * d0 + C10 * d1 + C20 * d2) * idet;
  const double q1 = (C01 * d0 + C11 * d1 + C21 * d2) * idet;
  const double q2 = (C02 * d0 + C12 * d1 + C22 * d2) * idet;
  const double dd = d0 * q0 + d1 * q1 + d2 * q2;
  if (WANT_JAC) {
    const double p0 = (C00 * d0 + C01 * d1 + C02 * d2) * idet;
    const double p1 = (C10 * d0 + C11 * d1 + C12 * d2) * idet;
    const double p2 = (C20 * d0 + C21 * d1 + C22 * d2) * idet;
    const double u0 = q0 + p0, u1 = q1 + p1;
    const double Mq0 = M00 * q0 + M01 * q1 + M02 * q2, Mq1 = M10 * q0 + M11 * q1 + M12 * q2;
    const double Mtp0 = M00 * p0 + M10 * p1 + M20 * p2, Mtp1 = M01 * p0 + M11 * p1 + M21 * p2;
    const double pSMq = (p1 * Mq0 - p0 * Mq1) + (Mtp0 * q1 - Mtp1 * q0);
    nt = u1 * xr - u0 * yr - pSMq;
    nx = u0; ny = u1;
  }
  return dd;
}

// ---- 2-D core (x, y) ---------------------------------------------------------------------------------------------
// RAW = true : R = [c -s; s c] with the stored, un-normalised complex (Sophus SE2 * point / rotationMatrix()); derivative
//              numerators w.r.t. c and s are independent (nc, ns).
// RAW = false: proper rotation by theta; nc carries the theta numerator.
template <bool WANT_JAC, bool RAW>
__device__ __forceinline__ double core2(double c, double s, double tx, double ty, const double mm[3], const double S[9],
                                        const double fm[3], const double F[9], double& nc, double& ns, double& nx, double& ny) {
  const double xr = c * mm[0] - s * mm[1];
  const double yr = s * mm[0] + c * mm[1];
  const double d0 = xr + tx - fm[0], d1 = yr + ty - fm[1];
  const double T00 = c * S[0] - s * S[3], T01 = c * S[1] - s * S[4];
  const double T10 = s * S[0] + c * S[3], T11 = s * S[1] + c * S[4];
  const double M00 = T00 * c - T01 * s, M01 = T00 * s + T01 * c;
  const double M10 = T10 * c - T11 * s, M11 = T10 * s + T11 * c;
  const double B00 = M00 + F[0], B01 = M01 + F[1], B10 = M10 + F[3], B11 = M11 + F[4];
  const double det = B00 * B11 - B10 * B01;
  const double idet = 1.0 / det;
  const double q0 = (B11 * d0 - B01 * d1) * idet, q1 = (-B10 * d0 + B00 * d1) * idet;
  const double dd = d0 * q0 + d1 * q1;
  if (WANT_JAC) {
    const double p0 = (B11 * d0 - B10 * d1) * idet, p1 = (-B01 * d0 + B00 * d1) * idet;
    const double u0 = q0 + p0, u1 = q1 + p1;
    nx = u0; ny = u1;
    if (RAW) {
      // dB/dc = K + T,  K = A R^T ;  dB/ds = S2 K + T S2^T
      const double K00 = S[0] * c - S[1] * s, K01 = S[0] * s + S[1] * c;
      const double K10 = S[3] * c - S[4] * s, K11 = S[3] * s + S[4] * c;
      const double Dc00 = K00 + T00, Dc01 = K01 + T01, Dc10 = K10 + T10, Dc11 = K11 + T11;
      const double Ds00 = -K10 - T01, Ds01 = -K11 + T00, Ds10 = K00 - T11, Ds11 = K01 + T10;
      const double pDcq = p0 * (Dc00 * q0 + Dc01 * q1) + p1 * (Dc10 * q0 + Dc11 * q1);
      const double pDsq = p0 * (Ds00 * q0 + Ds01 * q1) + p1 * (Ds10 * q0 + Ds11 * q1);
      nc = u0 * mm[0] + u1 * mm[1] - pDcq;
      ns = -u0 * mm[1] + u1 * mm[0] - pDsq;
    } else {
      const double Mq0 = M00 * q0 + M01 * q1, Mq1 = M10 * q0 + M11 * q1;
      const double Mtp0 = M00 * p0 + M10 * p1, Mtp1 = M01 * p0 + M11 * p1;
      const double pSMq = (p1 * Mq0 - p0 * Mq1) + (Mtp0 * q1 - Mtp1 * q0);
      nc = u1 * xr - u0 * yr - pSMq;
      ns = 0.0;
    }
  }
  return dd;
}

template <int VARIANT> struct VarTraits;
template <> struct VarTraits<0> { static constexpr int NB = 3, NP = 4; };  // basis (theta, x, y) -> ambient (c, s, tx, ty)
template <> struct VarTraits<1> { static constexpr int NB = 4, NP = 4; };  // basis (c, s, x, y)
template <> struct VarTraits<2> { static constexpr int NB = 3, NP = 3; };  // basis (x, y, theta)
template <> struct VarTraits<3> { static constexpr int NB = 3, NP = 3; };

// residual + basis Jacobian for one pair.  Returns false when the pair is degenerate (non-finite or negative dd).
template <int VARIANT, bool WANT_JAC>
__device__ __forceinline__ bool eval_pair(const SegConst& k, const float4* __restrict__ cm, const float4* __restrict__ cf, uint2 pr,
                                          double& r, double& dd, double* jb) {
  double mm[3], S[9], fm[3], F[9];
  load_cell(cm, pr.x, mm, S);
  load_cell(cf, pr.y, fm, F);
  double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
  if (VARIANT == 0 || VARIANT == 2) dd = core3<WANT_JAC>(k.c, k.s, k.tx, k.ty, mm, S, fm, F, n0, n1, n2);
  else if (VARIANT == 1) dd = core2<WANT_JAC, true>(k.c, k.s, k.tx, k.ty, mm, S, fm, F, n0, n1, n2, n3);
  else dd = core2<WANT_JAC, false>(k.c, k.s, k.tx, k.ty, mm, S, fm, F, n0, n1, n2, n3);
  const bool ok = (dd >= 0.0) && (dd < 1.0e300);   // false for NaN, inf, negative
  if (!ok) { r = 0.0; dd = 0.0; if (WANT_JAC) { for (int i = 0; i < VarTraits<VARIANT>::NB; ++i) jb[i] = 0.0; } return false; }
  if (dd == 0.0) {  // r = 0: the reference's dual-number sqrt yields 0/0 here; defined as J = 0
    r = 0.0;
    if (WANT_JAC) for (int i = 0; i < VarTraits<VARIANT>::NB; ++i) jb[i] = 0.0;
    return true;
  }
  const double rs = rsqrt(dd);
  r = dd * rs;
  if (WANT_JAC) {
    const double h = 0.5 * rs;
    if (VARIANT == 0) { jb[0] = n0 * h; jb[1] = n1 * h; jb[2] = n2 * h; }                 // (theta, x, y)
    else if (VARIANT == 2) { jb[0] = n1 * h; jb[1] = n2 * h; jb[2] = n0 * h; }            // (x, y, theta)
    else if (VARIANT == 1) { jb[0] = n0 * h; jb[1] = n1 * h; jb[2] = n2 * h; jb[3] = n3 * h; }  // (c, s, x, y)
    else { jb[0] = n2 * h; jb[1] = n3 * h; jb[2] = n0 * h; }                              // (x, y, theta)
  }
  return true;
}

// ---- loss: rho(s) and rho'(s), both already multiplied by the ScaledLoss weight -------------------------------
template <int LOSS>
__device__ __forceinline__ void loss_eval(double s, const SegConst& k, double& rho, double& rho1) {
  if (LOSS == L_NONE) { rho = s * k.weight; rho1 = k.weight; }
  else if (LOSS == L_WELSCH) {
    const double ex = exp(s * k.lc);           // lc = -1/b
    rho = k.lb * (1.0 - ex) * k.weight; rho1 = ex * k.weight;
  } else if (LOSS == L_BARRON_M2) {            // alpha = -2  -> exponent -1
    const double inv = 1.0 / (s * k.ts + 1.0);
    rho = k.pre * (inv - 1.0) * k.weight;
    rho1 = k.pre * k.e * (inv * inv) * k.ts * k.weight;
  } else if (LOSS == L_BARRON_M1) {            // alpha = -1  -> exponent -1/2
    const double rs = rsqrt(s * k.ts + 1.0);
    rho = k.pre * (rs - 1.0) * k.weight;
    rho1 = k.pre * k.e * (rs * rs * rs) * k.ts * k.weight;
  } else {                                     // generic Barron, same branches as BarronLoss::Evaluate
    const double alpha = 2.0 * k.e;
    if (alpha >= 2.0) { rho = s * k.weight; rho1 = k.weight; }
    else if (fabs(alpha) <= 0.05) {
      const double sum = 1.0 + s * k.lc, inv = 1.0 / sum;
      rho = k.lb * log(sum) * k.weight;
      rho1 = fmax(2.2250738585072014e-308, inv) * k.weight;
    } else {
      const double to_exp = s * k.ts + 1.0;
      const double p1 = pow(to_exp, k.e - 1.0);
      rho = k.pre * (p1 * to_exp - 1.0) * k.weight;
      rho1 = k.pre * k.e * p1 * k.ts * k.weight;
    }
  }
}

template <int VARIANT>
__device__ __forceinline__ void make_pose_const(const double* __restrict__ pose, SegConst& k) {
  if (VARIANT == 0) {
    const double c = pose[0], s = pose[1];
    const double n2 = c * c + s * s, n = sqrt(n2);
    k.c = c / n; k.s = s / n; k.tx = pose[2]; k.ty = pose[3];
    k.ja = -s / n2; k.jb = c / n2;
  } else if (VARIANT == 1) {
    k.c = pose[0]; k.s = pose[1]; k.tx = pose[2]; k.ty = pose[3]; k.ja = 0; k.jb = 0;
  } else {
    // NormalizeAngle (R/include/ndt_registration/state_manifold.h:17-23) then cos/sin
    const double two_pi = 2.0 * 3.14159265358979323846;
    const double th = pose[2] - two_pi * floor((pose[2] + 3.14159265358979323846) / two_pi);
    double sn, cs;
    sincos(th, &sn, &cs);
    k.c = cs; k.s = sn; k.tx = pose[0]; k.ty = pose[1]; k.ja = 0; k.jb = 0;
  }
}

__device__ __forceinline__ void make_loss_const(const LossParams& lp, double mu, SegConst& k) {
  k.weight = lp.weight;
  const double b = mu * lp.a2;
  k.lb = b;
  if (lp.kind == RANDT_LOSS_WELSCH) { k.lc = -1.0 / b; k.pre = 0; k.ts = 0; k.e = 0; return; }
  const double c = 1.0 / b, factor = fabs(lp.alpha - 2.0);
  k.lc = c; k.e = 0.5 * lp.alpha; k.pre = b * factor / lp.alpha; k.ts = 2.0 * c / factor;
}

// expand the basis normal equations of one segment into the 24-double output record
template <int VARIANT>
__device__ void write_segment_out(const double* __restrict__ tot, const SegConst& k, uint32_t n_pairs, double* __restrict__ out) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NH = NB * (NB + 1) / 2;
  double H[16], g[4];
  for (int i = 0; i < 16; ++i) H[i] = 0.0;
  for (int i = 0; i < 4; ++i) g[i] = 0.0;
  // unpack upper triangle in (i <= j) row-major order
  double Hb[4][4]; double gb[4];
  int t = 0;
  for (int i = 0; i < NB; ++i) for (int j = i; j < NB; ++j) { Hb[i][j] = tot[t]; Hb[j][i] = tot[t]; ++t; }
  for (int i = 0; i < NB; ++i) gb[i] = tot[NH + i];
  if (VARIANT == 0) {
    // ambient (c, s, tx, ty) = E^T (theta, x, y),  E rows: theta -> (ja, jb, 0, 0), x -> (0,0,1,0), y -> (0,0,0,1)
    const double E[3][4] = {{k.ja, k.jb, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int a = 0; a < 4; ++a) {
      for (int i = 0; i < 3; ++i) g[a] += E[i][a] * gb[i];
      for (int b2 = 0; b2 < 4; ++b2) {
        double s = 0; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s += E[i][a] * Hb[i][j] * E[j][b2];
        H[a * 4 + b2] = s;
      }
    }
  } else {
    for (int a = 0; a < NB; ++a) { g[a] = gb[a]; for (int b2 = 0; b2 < NB; ++b2) H[a * 4 + b2] = Hb[a][b2]; }
  }
  for (int i = 0; i < 16; ++i) out[RANDT_FUSED_H + i] = H[i];
  for (int i = 0; i < 4; ++i) out[RANDT_FUSED_G + i] = g[i];
  out[RANDT_FUSED_COST] = tot[NH + NB + 0];
  out[RANDT_FUSED_MAXR] = tot[NH + NB + 1];
  out[RANDT_FUSED_SUMSQ] = tot[NH + NB + 2];
  out[RANDT_FUSED_N] = (double)n_pairs;
}

template <int VARIANT, int LOSS, bool WANT_JAC>
__global__ void __launch_bounds__(kK3Threads) k3_fused_kernel(DeviceProblem P, const double* __restrict__ poses, LossParams lp,
                                                             const double* __restrict__ mu_per_seg, double* __restrict__ out,
                                                             unsigned long long* __restrict__ bad_counter) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NP = VarTraits<VARIANT>::NP;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NACC = NH + NB + 4;              // H, g, cost, max_r, sum_sq, bad
  constexpr int IDX_MAX = NH + NB + 1;
  __shared__ SegConst kc;
  __shared__ double red[NACC][kK3Threads];
  __shared__ double tot[NACC];
  __shared__ uint32_t ticket;
  const int tid = threadIdx.x;

  for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
    const Tile tile = P.tiles[t];
    if (tid == 0) {
      make_pose_const<VARIANT>(poses + (size_t)tile.seg * NP, kc);
      make_loss_const(lp, mu_per_seg ? mu_per_seg[tile.seg] : lp.mu, kc);
    }
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    for (uint32_t i = tile.begin + tid; i < tile.end; i += kK3Threads) {
      const uint2 pr = P.pairs[i];
      double r, dd, jb[4];
      const bool ok = eval_pair<VARIANT, WANT_JAC>(kc, P.cells_m, P.cells_f, pr, r, dd, jb);
      if (!ok) { acc[NACC - 1] += 1.0; continue; }
      double rho, rho1;
      loss_eval<LOSS>(dd, kc, rho, rho1);
      if (WANT_JAC) {
        int q = 0;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          const double wa = rho1 * jb[a];
#pragma unroll
          for (int b2 = a; b2 < NB; ++b2) { acc[q] += wa * jb[b2]; ++q; }
          acc[NH + a] += wa * r;
        }
      }
      acc[NH + NB + 0] += 0.5 * rho;
      acc[IDX_MAX] = fmax(acc[IDX_MAX], r);
      acc[NH + NB + 2] += dd;
    }
    // ---- block reduction through shared memory, fixed order ----
#pragma unroll
    for (int i = 0; i < NACC; ++i) red[i][tid] = acc[i];
    __syncthreads();
    if (tid < NACC * 4) {
      // 4 lanes per quantity (NACC*4 <= 72 <= 128): each sums 32 strided entries, then 2 shuffle steps
      const int qn = tid >> 2, j = tid & 3;
      double v = red[qn][j];
      if (qn == IDX_MAX) { for (int e = j + 4; e < kK3Threads; e += 4) v = fmax(v, red[qn][e]); }
      else               { for (int e = j + 4; e < kK3Threads; e += 4) v += red[qn][e]; }
      const unsigned mask = __activemask();
      double o = __shfl_xor_sync(mask, v, 2); v = (qn == IDX_MAX) ? fmax(v, o) : v + o;
      o = __shfl_xor_sync(mask, v, 1);        v = (qn == IDX_MAX) ? fmax(v, o) : v + o;
      if (j == 0) tot[qn] = v;
    }
    __syncthreads();
    const uint32_t seg_tiles = P.seg_first_tile[tile.seg + 1] - P.seg_first_tile[tile.seg];
    if (seg_tiles == 1) {
      if (tid == 0) {
        write_segment_out<VARIANT>(tot, kc, tile.end - tile.begin, out + (size_t)tile.seg * RANDT_FUSED_STRIDE);
        if (tot[NACC - 1] != 0.0) atomicAdd(bad_counter, (unsigned long long)tot[NACC - 1]);
      }
    } else {
      if (tid < NACC) P.partials[(size_t)t * kMaxAcc + tid] = tot[tid];
      __threadfence();
      __syncthreads();
      if (tid == 0) ticket = atomicAdd(&P.seg_counters[tile.seg], 1u);
      __syncthreads();
      if (ticket == seg_tiles - 1) {   // last tile of this segment to finish: fold partials in tile order
        __threadfence();
        const uint32_t t0 = P.seg_first_tile[tile.seg];
        if (tid < NACC) {
          double v = 0.0;
          for (uint32_t u = 0; u < seg_tiles; ++u) {
            const double x = __ldcg(&P.partials[(size_t)(t0 + u) * kMaxAcc + tid]);
            v = (tid == IDX_MAX) ? fmax(v, x) : v + x;
          }
          tot[tid] = v;
        }
        __syncthreads();
        if (tid == 0) {
          const uint32_t pb = P.tiles[t0].begin, pe = P.tiles[t0 + seg_tiles - 1].end;
          write_segment_out<VARIANT>(tot, kc, pe - pb, out + (size_t)tile.seg * RANDT_FUSED_STRIDE);
          if (tot[NACC - 1] != 0.0) atomicAdd(bad_counter, (unsigned long long)tot[NACC - 1]);
          P.seg_counters[tile.seg] = 0u;   // re-arm for the next launch
        }
      }
    }
    __syncthreads();
  }
}

// EMIT: raw residual and ambient Jacobian row per pair (what Evaluate returns for each block)
template <int VARIANT, bool WANT_JAC>
__global__ void __launch_bounds__(kK3Threads) k3_emit_kernel(DeviceProblem P, const double* __restrict__ poses, double* __restrict__ r_out,
                                                            double* __restrict__ J_out, unsigned long long* __restrict__ bad_counter) {
  constexpr int NP = VarTraits<VARIANT>::NP;
  __shared__ SegConst kc;
  const int tid = threadIdx.x;
  for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
    const Tile tile = P.tiles[t];
    if (tid == 0) make_pose_const<VARIANT>(poses + (size_t)tile.seg * NP, kc);
    __syncthreads();
    for (uint32_t i = tile.begin + tid; i < tile.end; i += kK3Threads) {
      const uint2 pr = P.pairs[i];
      double r, dd, jb[4];
      const bool ok = eval_pair<VARIANT, WANT_JAC>(kc, P.cells_m, P.cells_f, pr, r, dd, jb);
      if (!ok) atomicAdd(bad_counter, 1ull);
      r_out[i] = ok ? r : __longlong_as_double(0x7ff8000000000000ll);
      if (WANT_JAC) {
        if (VARIANT == 0) {
          double2* dst = reinterpret_cast<double2*>(J_out + (size_t)i * 4);
          dst[0] = make_double2(jb[0] * kc.ja, jb[0] * kc.jb);
          dst[1] = make_double2(jb[1], jb[2]);
        } else if (VARIANT == 1) {
          double2* dst = reinterpret_cast<double2*>(J_out + (size_t)i * 4);
          dst[0] = make_double2(jb[0], jb[1]);
          dst[1] = make_double2(jb[2], jb[3]);
        } else {
          J_out[(size_t)i * 3 + 0] = jb[0]; J_out[(size_t)i * 3 + 1] = jb[1]; J_out[(size_t)i * 3 + 2] = jb[2];
        }
      }
    }
    __syncthreads();
  }
}

// SWEEP: one thread per candidate pose, pairs of one segment staged through shared memory in chunks (broadcast reads).
constexpr int kSweepThreads = 128;
constexpr int kSweepChunk = 64;   // pairs staged per iteration: 64 * 24 doubles = 12 KB
template <int VARIANT, int LOSS>
__global__ void __launch_bounds__(kSweepThreads) k3_sweep_kernel(DeviceProblem P, uint32_t pair_begin, uint32_t pair_end,
                                                                const double* __restrict__ poses, uint32_t n_poses, LossParams lp,
                                                                double* __restrict__ cost_out) {
  constexpr int NP = VarTraits<VARIANT>::NP;
  __shared__ double sm_m[kSweepChunk][12];
  __shared__ double sm_f[kSweepChunk][12];
  const uint32_t pi = blockIdx.x * kSweepThreads + threadIdx.x;
  SegConst k;
  const bool active = pi < n_poses;
  if (active) make_pose_const<VARIANT>(poses + (size_t)pi * NP, k);
  else { const double idp[4] = {1, 0, 0, 0}; make_pose_const<VARIANT>(idp, k); }
  make_loss_const(lp, lp.mu, k);
  double cost = 0.0;
  for (uint32_t base = pair_begin; base < pair_end; base += kSweepChunk) {
    const uint32_t n = min((uint32_t)kSweepChunk, pair_end - base);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n * 24; e += kSweepThreads) {
      const uint32_t pp = e / 24, w = e % 24;
      const uint2 pr = P.pairs[base + pp];
      const float* src = (w < 12) ? reinterpret_cast<const float*>(P.cells_m) + (size_t)pr.x * 12 + w
                                  : reinterpret_cast<const float*>(P.cells_f) + (size_t)pr.y * 12 + (w - 12);
      if (w < 12) sm_m[pp][w] = (double)__ldg(src); else sm_f[pp][w - 12] = (double)__ldg(src);
    }
    __syncthreads();
    if (active) {
      for (uint32_t pp = 0; pp < n; ++pp) {
        double n0, n1, n2, n3, dd;
        if (VARIANT == 0 || VARIANT == 2) dd = core3<false>(k.c, k.s, k.tx, k.ty, &sm_m[pp][0], &sm_m[pp][3], &sm_f[pp][0], &sm_f[pp][3], n0, n1, n2);
        else if (VARIANT == 1) dd = core2<false, true>(k.c, k.s, k.tx, k.ty, &sm_m[pp][0], &sm_m[pp][3], &sm_f[pp][0], &sm_f[pp][3], n0, n1, n2, n3);
        else dd = core2<false, false>(k.c, k.s, k.tx, k.ty, &sm_m[pp][0], &sm_m[pp][3], &sm_f[pp][0], &sm_f[pp][3], n0, n1, n2, n3);
        if (!((dd >= 0.0) && (dd < 1.0e300))) continue;
        double rho, rho1;
        loss_eval<LOSS>(dd, k, rho, rho1);
        cost += 0.5 * rho;
      }
    }
  }
  if (active) cost_out[pi] = cost;
}

int loss_code(const LossParams& lp) {
  if (lp.kind == RANDT_LOSS_NONE) return L_NONE;
  if (lp.kind == RANDT_LOSS_WELSCH) return L_WELSCH;
  if (lp.alpha == -2.0) return L_BARRON_M2;
  if (lp.alpha == -1.0) return L_BARRON_M1;
  return L_BARRON;
}

template <int VARIANT, int LOSS>
cudaError_t launch_fused_vl(const DeviceProblem& p, const double* d_poses, const LossParams& lp, const double* d_mu, bool want_jac,
                            double* d_out, unsigned long long* bad, cudaStream_t s) {
  const int grid = (int)min((uint32_t)(kSmCount * 16), p.n_tiles);
  if (want_jac) k3_fused_kernel<VARIANT, LOSS, true><<<grid, kK3Threads, 0, s>>>(p, d_poses, lp, d_mu, d_out, bad);
  else          k3_fused_kernel<VARIANT, LOSS, false><<<grid, kK3Threads, 0, s>>>(p, d_poses, lp, d_mu, d_out, bad);
  return cudaGetLastError();
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

cudaError_t launch_eval_fused(const DeviceProblem& p, int variant, const double* d_poses, const LossParams& lp, const double* d_mu,
                              bool want_jac, double* d_out, unsigned long long* d_bad, cudaStream_t s, int* n_launches) {
  if (p.n_tiles == 0) return cudaSuccess;
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
  if (p.n_tiles == 0) return cudaSuccess;
  const int grid = (int)min((uint32_t)(kSmCount * 16), p.n_tiles);
#define RANDT_EMIT(V)                                                                                              \
  if (d_J) k3_emit_kernel<V, true><<<grid, kK3Threads, 0, s>>>(p, d_poses, d_r, d_J, d_bad);             \
  else     k3_emit_kernel<V, false><<<grid, kK3Threads, 0, s>>>(p, d_poses, d_r, d_J, d_bad);
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
