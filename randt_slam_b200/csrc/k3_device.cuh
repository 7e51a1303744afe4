// K3 device code shared by the pair-evaluation kernels (k3_pair_eval.cu) and the persistent solver (k7_solve.cu): the per-pair closed
// form, the robust loss, the compact-record loader, the warp reductions, the record writer and the TMA / mbarrier primitives.
// Everything sits in an unnamed namespace: each translation unit that includes this header gets its own copy.
#pragma once
#include <math.h>

#include "common.cuh"

namespace randt {

namespace {

enum { L_NONE = 0, L_BARRON = 1, L_WELSCH = 2, L_BARRON_M2 = 3, L_BARRON_M1 = 4 };

constexpr unsigned kFull = 0xffffffffu;

struct PoseConst {
  double c, s, tx, ty;     // rotation entries used by the variant (normalised for 0/2/3, raw for 1)
  double cc, ss, cs, hd;   // c^2, s^2, c*s, (c^2 - s^2)/2
  double ch, sh;           // c/2, s/2
  double n2;               // c^2 + s^2 (1 unless variant 1)
  double ja, jb;           // variant 0: dtheta/dc, dtheta/ds
};

struct LossConst {
  double lb, lc, pre, ts, e, weight;
  double preW;             // pre * weight / 2      (Barron)
  double K;                // pre * e * ts * weight (Barron: weight * rho' = K * u^(e-1))
};

struct RawCell { float4 a, b, c; };   // mean (x, y, i) + row-major 3x3 covariance, as stored (12 floats)
// What the arithmetic consumes of a cell: mean, diagonal, and the fp64 sums of the off-diagonal couples (twice the symmetric part)
struct CellC { float mx, my, mi, s00, s11, s22; double b2, e2, f2; };   // b2 = S01 + S10, e2 = S02 + S20, f2 = S12 + S21
__device__ __forceinline__ CellC cellc_from_raw(const RawCell& m) {
  CellC c;
  c.mx = m.a.x; c.my = m.a.y; c.mi = m.a.z; c.s00 = m.a.w; c.s11 = m.b.w; c.s22 = m.c.w;
  c.b2 = (double)m.b.x + (double)m.b.z; c.e2 = (double)m.b.y + (double)m.c.y; c.f2 = (double)m.c.x + (double)m.c.z;
  return c;
}
// compact record couple -> fp64 sum: the bits of (double)hs with the 2-bit code at bits 27..28 (common.cuh: DuoRec)
constexpr uint32_t kCodeMask = 0x18000000u;
__device__ __forceinline__ double sym_decode(float hs, uint32_t code_at_27) {
  const double d = (double)hs;
  return __hiloint2double(__double2hiint(d), __double2loint(d) | (int)(code_at_27 & kCodeMask));
}
// the inverse, verified bit for bit: false when s = (double)a + (double)b is not (double)float_rz(s) | code << 27 with code < 4
__device__ __forceinline__ bool sym_encode(float a, float b, float& hs, uint32_t& code) {
  const double s = (double)a + (double)b;
  hs = __double2float_rz(s);
  const unsigned long long sb = (unsigned long long)__double_as_longlong(s), db = (unsigned long long)__double_as_longlong((double)hs);
  const unsigned long long c = (sb - db) >> 27;
  code = (uint32_t)(c & 3ull);
  return c < 4ull && (db | (c << 27)) == sb;
}

// One duo's three cells (moving, fixed of p0, fixed of p0 + 1 or a repeat of the first when the duo has one pair) -> its compact record
// (common.cuh: DuoRec); a duo whose couples do not fit the encoding goes to the overflow table as stored.
__device__ __forceinline__ void encode_duo_record(const RawCell* c, bool two, DuoRec* __restrict__ rec, DuoRecFull* __restrict__ overflow,
                                                  uint32_t overflow_cap, uint32_t* __restrict__ n_overflow) {
  float o[28];
  uint32_t w = two ? 0u : kRecNoSecond;
  bool fits = true;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    o[9 * q + 0] = c[q].a.x; o[9 * q + 1] = c[q].a.y; o[9 * q + 2] = c[q].a.z;
    o[9 * q + 3] = c[q].a.w; o[9 * q + 4] = c[q].b.w; o[9 * q + 5] = c[q].c.w;
    uint32_t code;
    fits = sym_encode(c[q].b.x, c[q].b.z, o[9 * q + 6], code) && fits; w |= code << (2 * (3 * q + 0));
    fits = sym_encode(c[q].b.y, c[q].c.y, o[9 * q + 7], code) && fits; w |= code << (2 * (3 * q + 1));
    fits = sym_encode(c[q].c.x, c[q].c.z, o[9 * q + 8], code) && fits; w |= code << (2 * (3 * q + 2));
  }
  if (!fits) {
    const uint32_t idx = atomicAdd(n_overflow, 1u);
    if (idx < overflow_cap) {
      float4* q = overflow[idx].v;
#pragma unroll
      for (int i = 0; i < 3; ++i) { q[3 * i] = c[i].a; q[3 * i + 1] = c[i].b; q[3 * i + 2] = c[i].c; }
    }
    o[0] = __uint_as_float(idx);
    w = (w & kRecNoSecond) | kRecEscape;
  }
  o[27] = __uint_as_float(w);
  float4* dst = rec->v;
#pragma unroll
  for (int i = 0; i < 7; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}

// 1/x for a normal, finite, non-zero x: MUFU.RCP64H seed (2^-23) + two Newton steps.  No slow path: callers guarantee or
// tolerate garbage-in-garbage-out (degenerate pairs are caught by the validity test on dd).
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}
// 1/sqrt(x), x normal and positive: MUFU.RSQ64H seed + two Newton steps
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

template <int VARIANT> struct VarTraits;
template <> struct VarTraits<0> { static constexpr int NB = 3, NP = 4; };  // basis (theta, x, y) -> ambient (c, s, tx, ty)
template <> struct VarTraits<1> { static constexpr int NB = 4, NP = 4; };  // basis (c, s, x, y)
template <> struct VarTraits<2> { static constexpr int NB = 3, NP = 3; };  // basis (x, y, theta)
template <> struct VarTraits<3> { static constexpr int NB = 3, NP = 3; };

// ---- per-pair core ------------------------------------------------------------------------------------------------
// Everything that depends on the moving cell and the pose only (shared by the pairs of a duo).
struct Moving {
  double mx, my, mi, S00, S11, bh, S22;   // bh = sym(S01)
  double M00, M01, M11, M02, M12;         // R S R^T (symmetric; M22 = S22)
  double xr, yr;                          // R mu_m (no translation)
};
template <int VARIANT>
__device__ __forceinline__ void moving_part(const PoseConst& k, const CellC& m, Moving& o) {
  o.mx = m.mx; o.my = m.my; o.S00 = m.s00; o.S11 = m.s11;
  const double b2 = m.b2;                                       // 2 * sym(S01)
  o.bh = 0.5 * b2;
  o.M00 = fma(k.cc, o.S00, fma(k.ss, o.S11, -k.cs * b2));
  o.M11 = fma(k.n2, o.S00 + o.S11, -o.M00);
  o.M01 = fma(k.cs, o.S00 - o.S11, k.hd * b2);
  o.xr = fma(k.c, o.mx, -k.s * o.my);
  o.yr = fma(k.s, o.mx, k.c * o.my);
  if (VARIANT == 0 || VARIANT == 2) {
    const double e2 = m.e2;                                     // 2 * sym(S02)
    const double f2 = m.f2;                                     // 2 * sym(S12)
    o.M02 = fma(k.ch, e2, -k.sh * f2);
    o.M12 = fma(k.sh, e2, k.ch * f2);
    o.mi = m.mi; o.S22 = m.s22;
  } else {
    o.M02 = 0.0; o.M12 = 0.0; o.mi = 0.0; o.S22 = 0.0;
  }
}
// dd = d^T B^-1 d and (WANT_JAC) the basis numerators N[] = r * dr/d(basis), in the basis order of VarTraits.
template <int VARIANT, bool WANT_JAC>
__device__ __forceinline__ double fixed_part(const PoseConst& k, const Moving& mv, const CellC& f, double* __restrict__ N) {
  const double d0 = (mv.xr + k.tx) - (double)f.mx;
  const double d1 = (mv.yr + k.ty) - (double)f.my;
  const double B00 = mv.M00 + (double)f.s00;
  const double B11 = mv.M11 + (double)f.s11;
  const double B01 = fma(0.5, f.b2, mv.M01);
  double q0, q1, q2 = 0.0, dd;
  if (VARIANT == 0 || VARIANT == 2) {
    const double d2 = mv.mi - (double)f.mi;
    const double B22 = mv.S22 + (double)f.s22;
    const double B02 = fma(0.5, f.e2, mv.M02);
    const double B12 = fma(0.5, f.f2, mv.M12);
    const double C00 = fma(B11, B22, -B12 * B12);
    const double C01 = fma(B02, B12, -B01 * B22);
    const double C02 = fma(B01, B12, -B02 * B11);
    const double C11 = fma(B00, B22, -B02 * B02);
    const double C12 = fma(B01, B02, -B00 * B12);
    const double C22 = fma(B00, B11, -B01 * B01);
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
  if (WANT_JAC) {
    if (VARIANT == 1) {
      // R = [c -s; s c] un-normalised, c and s independent parameters:
      //   N_c = q.(mx, my) - g^T S q,   N_s = q.(-my, mx) - (q1 (S g)0 - q0 (S g)1),   g = R^T q
      const double g0 = fma(k.c, q0, k.s * q1), g1 = fma(k.c, q1, -k.s * q0);
      const double Sq0 = fma(mv.S00, q0, mv.bh * q1), Sq1 = fma(mv.bh, q0, mv.S11 * q1);
      const double Sg0 = fma(mv.S00, g0, mv.bh * g1), Sg1 = fma(mv.bh, g0, mv.S11 * g1);
      N[0] = fma(q0, mv.mx, q1 * mv.my) - fma(g0, Sq0, g1 * Sq1);
      N[1] = fma(q1, mv.mx, -q0 * mv.my) - fma(q1, Sg0, -q0 * Sg1);
      N[2] = q0; N[3] = q1;
    } else {
      const double a0 = mv.xr - fma(mv.M00, q0, fma(mv.M01, q1, mv.M02 * q2));
      const double a1 = mv.yr - fma(mv.M01, q0, fma(mv.M11, q1, mv.M12 * q2));
      const double nt = fma(q1, a0, -q0 * a1);
      if (VARIANT == 0) { N[0] = nt; N[1] = q0; N[2] = q1; }      // (theta, x, y)
      else              { N[0] = q0; N[1] = q1; N[2] = nt; }      // (x, y, theta)
    }
  }
  return dd;
}
template <int VARIANT, bool WANT_JAC>
__device__ __forceinline__ double pair_core(const PoseConst& k, const RawCell& m, const RawCell& f, double* __restrict__ N) {
  Moving mv;
  moving_part<VARIANT>(k, cellc_from_raw(m), mv);
  return fixed_part<VARIANT, WANT_JAC>(k, mv, cellc_from_raw(f), N);
}

__device__ __forceinline__ bool dd_valid(double dd) { return (dd >= 0.0) && (dd < 1.0e300); }   // false for NaN, inf, negative

// ---- loss: w = weight*rho'(s), hrho = weight*rho(s)/2, wd = w/s (finite garbage when s == 0: it only multiplies N = 0) ----
// POS: the caller guarantees s > 0 (the select-free fast path)
template <int LOSS, bool POS = false>
__device__ __forceinline__ void loss_eval(double s, const LossConst& k, double& w, double& hrho, double& wd) {
  const bool pos = POS || s > 0.0;
  if (LOSS == L_BARRON_M2) {                    // alpha = -2: rho = pre (1/u - 1), rho' = pre e ts / u^2, u = s ts + 1
    const double u = fma(s, k.ts, 1.0);
    const double v = (u * u) * s;
    const double iv = rcp_fast(pos ? v : 1.0);  // 1 / (u^2 s)
    wd = k.K * iv;
    w = wd * s;
    const double iu = pos ? (iv * s) * u : 1.0;
    hrho = fma(k.preW, iu, -k.preW);
    return;
  }
  if (LOSS == L_NONE) { w = k.weight; hrho = 0.5 * k.weight * s; }
  else if (LOSS == L_WELSCH) {
    const double ex = exp(s * k.lc);            // lc = -1/b
    hrho = 0.5 * k.lb * (1.0 - ex) * k.weight; w = ex * k.weight;
  } else if (LOSS == L_BARRON_M1) {             // alpha = -1  -> exponent -1/2
    const double rs = rsqrt_fast(fma(s, k.ts, 1.0));
    hrho = fma(k.preW, rs, -k.preW);
    w = k.K * (rs * rs * rs);
  } else {                                      // generic Barron, same branches as BarronLoss::Evaluate
    const double alpha = 2.0 * k.e;
    if (alpha >= 2.0) { w = k.weight; hrho = 0.5 * k.weight * s; }
    else if (fabs(alpha) <= 0.05) {
      const double sum = 1.0 + s * k.lc, inv = 1.0 / sum;
      hrho = 0.5 * k.lb * log(sum) * k.weight;
      w = fmax(2.2250738585072014e-308, inv) * k.weight;
    } else {
      const double to_exp = fma(s, k.ts, 1.0);
      const double p1 = pow(to_exp, k.e - 1.0);
      hrho = k.preW * (p1 * to_exp - 1.0);
      w = k.K * p1;
    }
  }
  wd = w * rcp_fast(pos ? s : 1.0);
}

template <int VARIANT>
__device__ __forceinline__ void make_pose_const(const double* __restrict__ pose, PoseConst& k) {
  k.n2 = 1.0; k.ja = 0.0; k.jb = 0.0;
  if (VARIANT == 0) {
    // theta = atan2(s, c): rotate by the normalised complex; dtheta/dc = -s/|z|^2, dtheta/ds = c/|z|^2
    const double c = pose[0], s = pose[1];
    const double n2 = fma(c, c, s * s);
    const double rn = rsqrt_fast(n2), in2 = rn * rn;
    k.c = c * rn; k.s = s * rn; k.tx = pose[2]; k.ty = pose[3];
    k.ja = -s * in2; k.jb = c * in2;
  } else if (VARIANT == 1) {
    k.c = pose[0]; k.s = pose[1]; k.tx = pose[2]; k.ty = pose[3];
    k.n2 = k.c * k.c + k.s * k.s;
  } else {
    // NormalizeAngle (R/include/ndt_registration/state_manifold.h:17-23) then cos/sin
    const double two_pi = 2.0 * 3.14159265358979323846;
    const double th = pose[2] - two_pi * floor((pose[2] + 3.14159265358979323846) / two_pi);
    double sn, cs;
    sincos(th, &sn, &cs);
    k.c = cs; k.s = sn; k.tx = pose[0]; k.ty = pose[1];
  }
  k.cc = k.c * k.c; k.ss = k.s * k.s; k.cs = k.c * k.s; k.hd = 0.5 * (k.cc - k.ss);
  k.ch = 0.5 * k.c; k.sh = 0.5 * k.s;
}

// loss constants for one mu.  lp carries the mu-independent factors (host-computed, LossParams in common.cuh)
__device__ __forceinline__ void make_loss_const(const LossParams& lp, double mu, LossConst& k) {
  k.weight = lp.weight;
  const double b = mu * lp.a2;
  const double c = rcp_fast(b);
  k.lb = b; k.e = 0.5 * lp.alpha;
  if (lp.kind == RANDT_LOSS_WELSCH) { k.lc = -c; k.pre = 0; k.ts = 0; k.preW = 0; k.K = 0; return; }
  k.lc = c; k.pre = b * lp.fa; k.ts = c * lp.tf;
  k.preW = 0.5 * k.pre * k.weight;
  k.K = k.pre * k.e * k.ts * k.weight;
}

// ---- warp butterfly reduction ----------------------------------------------------------------------------------------
// N per-lane partial sums -> every lane ends up with the 32-lane total of ONE slot; ceil(N/2) + ceil(N/4) + ... shuffles
// instead of 5 N.  Fixed exchange pattern => bitwise reproducible.
template <int N, int XOR>
__device__ __forceinline__ void bfly_reduce(double* a, int lane) {
  if constexpr (XOR >= 1) {
    if constexpr (N > 1) {
      constexpr int H = (N + 1) / 2;
      const bool up = (lane & XOR) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const double lo = a[i];
        const double hi = (i + H < N) ? a[i + H] : 0.0;
        const double send = up ? lo : hi;
        const double keep = up ? hi : lo;
        a[i] = keep + __shfl_xor_sync(kFull, send, XOR);
      }
      bfly_reduce<H, XOR / 2>(a, lane);
    } else {
      a[0] += __shfl_xor_sync(kFull, a[0], XOR);
      bfly_reduce<1, XOR / 2>(a, lane);
    }
  }
}
// the lane (with bit 0 clear) that holds slot `slot` after bfly_reduce<N, 16>
template <int N>
__host__ __device__ constexpr int bfly_lane_of_slot(int slot) {
  int lane = 0, n = N, s = slot;
  for (int x = 16; x >= 2; x >>= 1) {
    if (n > 1) {
      const int h = (n + 1) / 2;
      if (s >= h) { lane |= x; s -= h; }
      n = h;
    }
  }
  return lane;
}
// slot -> owner lane after bfly_reduce<N, 16>, packed 5 bits per slot (N <= 12 per word)
template <int N>
__host__ __device__ constexpr unsigned long long bfly_lane_table(int first) {
  unsigned long long t = 0;
  for (int s = 0; s < 12 && first + s < N; ++s) t |= (unsigned long long)bfly_lane_of_slot<N>(first + s) << (5 * s);
  return t;
}
template <int N>
__device__ __forceinline__ int bfly_owner(int slot) {
  constexpr unsigned long long t0 = bfly_lane_table<N>(0), t1 = bfly_lane_table<N>(12);
  return (int)(((slot < 12 ? t0 : t1) >> (5 * (slot < 12 ? slot : slot - 12))) & 31ull);
}
// Warp reduction of NS per-lane sums through shared memory (the just-consumed stage's record area is free at a tile's end):
// every lane stores its NS accumulators column-wise, then lane (2 s + h) adds half h of row s — sixteen values, as two
// independent chains — and one shuffle joins the halves.  ~45 instructions instead of ~150 for the register butterfly, fixed
// summation order (bitwise reproducible).  Afterwards slot s's total is in lanes 2 s and 2 s + 1.
template <int NS>
__device__ __forceinline__ double smem_reduce(const double* acc, double* scratch /* >= NS * 34 doubles, 16-byte aligned */, int lane) {
  static_assert(NS <= 16, "two lanes per slot");
  constexpr int LD = 34;                     // row stride (doubles): even, so that 16-byte loads stay aligned; 34 spreads rows over banks
  __syncwarp();                              // every lane has read its records out of this buffer
#pragma unroll
  for (int e = 0; e < NS; ++e) scratch[e * LD + lane] = acc[e];
  __syncwarp();
  const int s = lane >> 1, h = lane & 1;
  double t0 = 0.0, t1 = 0.0;
  if (s < NS) {
    const double2* row = reinterpret_cast<const double2*>(scratch + s * LD + 16 * h);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 v = row[i]; t0 += v.x; t1 += v.y; }
  }
  double t = t0 + t1;
  t += __shfl_xor_sync(kFull, t, 1);
  __syncwarp();                              // the buffer may be restaged once everybody is done reading
  return t;
}

// max of non-negative finite doubles over the warp: two integer REDUX (the IEEE order of non-negative doubles is the order of
// their bit patterns) instead of five 64-bit shuffle + compare rounds
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// Write the 24-double record of one segment cooperatively: lane e < 24 fetches the basis total its entry depends on from the lane
// that owns it, scales it, and the warp stores the record with one coalesced 8-byte-per-lane store.
// Slots: [NH upper triangle (i <= j, row-major)] [NB gradient] [cost] [sum dd].  What entry e needs does not depend on the tile, so
// every lane derives it once per kernel (out_map) and keeps it packed in one register:
//   bits 0..4 slot | 5..6 row factor | 7..8 column factor (0: 1, 1: ja, 2: jb, 3: 0) | 9..13 index in the packed layout | 14 stored in the
//   packed layout | 15 entry scales a slot total (else: max r, n, or nothing)
template <int VARIANT, bool WANT_JAC>
__device__ __forceinline__ uint32_t out_map(int lane) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NJ = WANT_JAC ? NH + NB : 0;
  const int e = lane;
  // entry e: H (e < 16: row r = e / 4, column c = e % 4), g (16..19), cost (20), max r (21), sum r^2 (22), n (23)
  int slot = 0, fr = 3, fc = 0, scaled = 0;
  if (e < 20) {
    if (WANT_JAC) {
      const int r = e < 16 ? (e >> 2) : (e - 16), c = e & 3;
      int tr, tc, sr, sc;
      if (VARIANT == 0) {   // ambient (c, s, tx, ty) from the tangent basis (theta, x, y): d theta/d c = ja, d theta/d s = jb
        tr = r < 2 ? 0 : r - 1; tc = c < 2 ? 0 : c - 1;
        sr = r == 0 ? 1 : (r == 1 ? 2 : 0); sc = c == 0 ? 1 : (c == 1 ? 2 : 0);
      } else {
        tr = r; tc = c; sr = r < NB ? 0 : 3; sc = c < NB ? 0 : 3;
        if (tr >= NB) tr = 0;
        if (tc >= NB) tc = 0;
      }
      if (e < 16) {
        const int i = tr < tc ? tr : tc, j = tr < tc ? tc : tr;
        slot = i * NB - (i * (i - 1)) / 2 + (j - i);
        fr = sr; fc = sc;
      } else { slot = NH + tr; fr = sr; fc = 0; }
      scaled = 1;
    }
  } else if (e == 20) { slot = NJ; fr = 0; fc = 0; scaled = 1; }
  else if (e == 22) { slot = NJ + 1; fr = 0; fc = 0; scaled = 1; }
  // RANDT_PACKED_*: of H only row <= column (index in the row-major upper triangle), everything after H moves up by six
  const int r = e >> 2, c = e & 3;
  const int pk = e < 16 ? r * 4 - (r * (r - 1)) / 2 + (c - r) : e - 6;
  const int in_packed = (e < RANDT_FUSED_STRIDE && (e >= 16 || c >= r)) ? 1 : 0;
  return (uint32_t)slot | ((uint32_t)fr << 5) | ((uint32_t)fc << 7) | ((uint32_t)(pk & 31) << 9) | ((uint32_t)in_packed << 14) | ((uint32_t)scaled << 15);
}
__device__ __forceinline__ double out_factor(uint32_t sel, double ja, double jb) {
  return sel == 0u ? 1.0 : (sel == 1u ? ja : (sel == 2u ? jb : 0.0));
}
// `mine` = the caller's own slot total; src_lane = the lane that owns the slot this lane's entry needs
__device__ __forceinline__ void write_segment_out_from(double mine, double max_dd, double ja, double jb, uint32_t n_pairs, double* __restrict__ out_base,
                                                       uint32_t seg, uint32_t packed, int lane, uint32_t omap, int src_lane);
// slot s is owned by lane s * owner_mul (smem_reduce: 2, the partial fold: 1)
__device__ __forceinline__ void write_segment_out(double mine, double max_dd, double ja, double jb, uint32_t n_pairs, double* __restrict__ out_base,
                                                  uint32_t seg, uint32_t packed, int lane, uint32_t omap, int owner_mul) {
  if (packed == 3u) {
    // RANDT_BASIS_*: the slot totals as they are (three-dimensional bases, Jacobian evaluations only: slots 0..5 upper triangle of H in
    // the kernel's own basis, 6..8 gradient, 9 cost) — no chain-rule factors, 80 bytes per segment
    const double v = __shfl_sync(kFull, mine, (lane < RANDT_BASIS_STRIDE ? lane : 0) * owner_mul);
    if (lane < RANDT_BASIS_STRIDE) out_base[(size_t)seg * RANDT_BASIS_STRIDE + lane] = v;
    return;
  }
  write_segment_out_from(mine, max_dd, ja, jb, n_pairs, out_base, seg, packed, lane, omap, (int)(omap & 31u) * owner_mul);
}
__device__ __forceinline__ void write_segment_out_from(double mine, double max_dd, double ja, double jb, uint32_t n_pairs, double* __restrict__ out_base,
                                                       uint32_t seg, uint32_t packed, int lane, uint32_t omap, int src_lane) {
  const int e = lane;
  const double v = __shfl_sync(kFull, mine, src_lane);
  double val = (omap & 0x8000u) ? v * (out_factor((omap >> 5) & 3u, ja, jb) * out_factor((omap >> 7) & 3u, ja, jb)) : 0.0;
  if (e == 21) val = max_dd > 0.0 ? max_dd * rsqrt_fast(max_dd) : 0.0;   // max raw residual
  if (e == 23) val = (double)n_pairs;
  if (!packed) {
    if (e < RANDT_FUSED_STRIDE) out_base[(size_t)seg * RANDT_FUSED_STRIDE + e] = val;
  } else if (omap & 0x4000u) {
    const uint32_t pk = (omap >> 9) & 31u;
    if (packed == 1u) out_base[(size_t)seg * RANDT_PACKED_STRIDE + pk] = val;
    else if (pk < (uint32_t)RANDT_CORE_STRIDE) out_base[(size_t)seg * RANDT_CORE_STRIDE + pk] = val;     // packed == 2: H, g, cost only
  }
}

// ---- the software-pipelined tile stream of one warp ------------------------------------------------------------------
// Every warp walks its chunk list (P.warp_off) as a stream of 32-duo chunks (a lane owns one duo = up to two pairs that share their
// moving cell).  Lane 0 issues ONE bulk copy per chunk (32 x 144 B, completion on the stage's mbarrier) and, for the first chunk of a
// tile, cp.async copies of the tile's pose and mu.  kStages chunks are in flight per warp, so registers hold nothing but the
// accumulators while HBM latency is covered.  The two pairs of a duo are evaluated as two independent instruction streams (ILP for
// the half-rate fp64 pipe).
constexpr int kWarpsPerCta = kK3Threads / 32;
#ifndef RANDT_K3_STAGES
#define RANDT_K3_STAGES 2
#endif
constexpr int kStages = RANDT_K3_STAGES;
constexpr int kMinCtas = RANDT_K3_MIN_CTAS;   // CTAs per SM the register allocation is bounded for

constexpr int kRecF4 = (int)(sizeof(DuoRec) / sizeof(float4));   // float4 per compact record (7)
// SCRATCH_BYTES: what the tile epilogue needs of the record area when it is reused as reduction scratch (NS * 34 doubles)
template <int SCRATCH_BYTES>
struct __align__(128) StageBufT {
  static constexpr int kAreaF4 = (SCRATCH_BYTES > 32 * (int)sizeof(DuoRec) ? SCRATCH_BYTES : 32 * (int)sizeof(DuoRec)) / 16;
  float4 rec[kAreaF4];     // the chunk's duo records (112 B each, see DuoRec in common.cuh), landed by ONE bulk copy
  double pose[2][4];       // [0]: pose of the tile at lane 0 (valid for the first chunk of a tile); [1]: pose of the tile that starts
  double mu[2];            //      inside a split chunk; mu likewise
  uint32_t segoff[2][2];   // seg_off[seg], seg_off[seg + 1] of those tiles (pairs of the segment, entry 23 of its record)
  unsigned long long bar;  // mbarrier the bulk copy completes on
};

// One lane's duo out of the landed chunk: seven conflict-free LDS.128, then the couples' sums are rebuilt from their codes.  A duo
// flagged kRecEscape fetches its three cells as stored from the overflow table instead (rare: divergent global loads).
__device__ __forceinline__ void load_duo(const float4* __restrict__ rec, int lane, const DuoRecFull* __restrict__ ovf, CellC& m, CellC& f0,
                                         CellC& f1, bool& two) {
  const float4* r = rec + lane * kRecF4;
  const float4 v0 = r[0], v1 = r[1], v2 = r[2], v3 = r[3], v4 = r[4], v5 = r[5], v6 = r[6];
  const uint32_t w = __float_as_uint(v6.w);
  two = (w & kRecNoSecond) == 0u;
  if ((w & kRecEscape) == 0u) {
    m.mx = v0.x; m.my = v0.y; m.mi = v0.z; m.s00 = v0.w; m.s11 = v1.x; m.s22 = v1.y;
    m.b2 = sym_decode(v1.z, w << 27); m.e2 = sym_decode(v1.w, w << 25); m.f2 = sym_decode(v2.x, w << 23);
    f0.mx = v2.y; f0.my = v2.z; f0.mi = v2.w; f0.s00 = v3.x; f0.s11 = v3.y; f0.s22 = v3.z;
    f0.b2 = sym_decode(v3.w, w << 21); f0.e2 = sym_decode(v4.x, w << 19); f0.f2 = sym_decode(v4.y, w << 17);
    f1.mx = v4.z; f1.my = v4.w; f1.mi = v5.x; f1.s00 = v5.y; f1.s11 = v5.z; f1.s22 = v5.w;
    f1.b2 = sym_decode(v6.x, w << 15); f1.e2 = sym_decode(v6.y, w << 13); f1.f2 = sym_decode(v6.z, w << 11);
  } else {
    const float4* q = ovf[__float_as_uint(v0.x)].v;
    RawCell c;
    c.a = __ldg(q + 0); c.b = __ldg(q + 1); c.c = __ldg(q + 2); m = cellc_from_raw(c);
    c.a = __ldg(q + 3); c.b = __ldg(q + 4); c.c = __ldg(q + 5); f0 = cellc_from_raw(c);
    c.a = __ldg(q + 6); c.b = __ldg(q + 7); c.c = __ldg(q + 8); f1 = cellc_from_raw(c);
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA 1-D bulk copy + mbarrier (sm_90+ PTX; SASS UBLKCP / SYNCS) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, unsigned long long* bar) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gmem), "r"(bytes), "r"(b)
               : "memory");
}
// programmatic dependent launch (sm_90+): let the next kernel of the stream start its prologue while this grid drains, and wait for
// the previous grid's memory before touching anything it may have produced
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- the warp's work queue -------------------------------------------------------------------------------------------
// The host cuts every tile into chunk descriptors (common.cuh: ChunkDesc) and stores them warp after warp in the balanced order, so
// the device side of the schedule is a flat list walk: 32 descriptors at a time are pulled into shared memory with one coalesced
// load (plus their segments' active flags, folded into a ballot mask); per chunk, lane 0 arms the stage's mbarrier and issues ONE
// bulk (TMA) copy of the chunk's duo records plus, for the first chunk of a tile, 16-byte cp.async copies of the segment's pose
// and mu.  No per-chunk index arithmetic, no gathers, no LSU traffic for the cell data.
struct WarpQueue {
  ChunkDesc* q;        // [64] in shared memory: two halves of 32; chunk j of the warp's list sits in q[j & 63]
  uint32_t c_begin, c_end, act[2];
  __device__ __forceinline__ uint4 load_desc(const DeviceProblem& P, uint32_t block, int lane) const {   // (duo_begin, meta, seg, part); meta 0 past the end
    const uint32_t idx = c_begin + block * 32u + (uint32_t)lane;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (idx < c_end) v = __ldg(reinterpret_cast<const uint4*>(P.chunks) + idx);
    return v;
  }
  __device__ __forceinline__ void publish(const DeviceProblem& P, uint4 v, uint32_t block, int lane) {   // active flags, queue, warp barrier
    uint32_t on = 0;
    if (v.y & kChunkCountMask) on = P.seg_active ? P.seg_active[v.z] : 1u;
    *reinterpret_cast<uint4*>(&q[(block & 1u) * 32u + (uint32_t)lane]) = v;
    const uint32_t mask = __ballot_sync(kFull, on != 0u);
    if (block & 1u) act[1] = mask; else act[0] = mask;
    __syncwarp();
  }
  // Block b (chunks 32 b .. 32 b + 31) replaces block b - 2 in its half: by the time chunk 32 b is staged, the chunk being consumed is
  // at most kStages - 1 <= 32 behind it, i.e. in block b - 1 or later.
  __device__ __forceinline__ void refill(const DeviceProblem& P, uint32_t block, int lane) { publish(P, load_desc(P, block, lane), block, lane); }
  __device__ __forceinline__ ChunkDesc get(uint32_t j) const {
    const uint4 v = *reinterpret_cast<const uint4*>(&q[j & 63u]);
    ChunkDesc d; d.duo_begin = v.x; d.meta = v.y; d.seg = v.z; d.part = v.w;
    return d;
  }
  __device__ __forceinline__ bool live(uint32_t j) const { return (((j & 32u) ? act[1] : act[0]) >> (j & 31u)) & 1u; }
};

template <int NP, bool WANT_SEGOFF, typename SB>
__device__ __forceinline__ void stage_issue(const DeviceProblem& P, const WarpQueue& wq, uint32_t j, int lane, SB* sb,
                                            const double* __restrict__ poses, const double* __restrict__ mu_per_seg) {
  if (lane == 0 && wq.live(j)) {
    const ChunkDesc d = wq.get(j);
    const uint32_t n_here = d.meta & kChunkCountMask;
    if (n_here) {
      const uint32_t bytes = n_here * (uint32_t)sizeof(DuoRec);
      mbar_expect_tx(&sb->bar, bytes);
      bulk_g2s(&sb->rec[0], P.duo_recs + d.duo_begin, bytes, &sb->bar);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (d.meta & (h == 0 ? kChunkFirst : kChunkSplit)) {
          const uint32_t sg = h == 0 ? d.seg : d.part;          // a split chunk carries its second segment in `part`
          const double* ps = poses + (size_t)sg * NP;
          if (NP == 4) { cp_async16(&sb->pose[h][0], ps); cp_async16(&sb->pose[h][2], ps + 2); }
          else { cp_async8(&sb->pose[h][0], ps); cp_async8(&sb->pose[h][1], ps + 1); cp_async8(&sb->pose[h][2], ps + 2); }
          if (mu_per_seg) cp_async8(&sb->mu[h], mu_per_seg + sg);
          if (WANT_SEGOFF) { cp_async4(&sb->segoff[h][0], P.seg_off + sg); cp_async4(&sb->segoff[h][1], P.seg_off + sg + 1); }
        }
      }
    }
  }
  cp_async_commit();
}

// Tile finished: reduce the per-lane sums across the warp (through `scratch`, the just-consumed stage buffer) and emit the
// segment's record — directly when the tile is its segment's only one, else as a partial that the last tile to finish folds.
template <int VARIANT, bool WANT_JAC, int NS>
__device__ __forceinline__ void finish_tile(const DeviceProblem& P, const double* vals, double max_dd, uint32_t n_bad, uint32_t seg, bool solo,
                                            uint32_t part_slot, const PoseConst& kc, uint32_t n_pairs_seg, uint32_t omap, double* scratch,
                                            double* __restrict__ out, unsigned long long* __restrict__ bad_counter, int lane) {
  const double mine = smem_reduce<NS>(vals, scratch, lane);   // slot s total: lanes 2s, 2s+1
  const double mx = warp_max_nonneg(max_dd);
  const uint32_t bad = __reduce_add_sync(kFull, n_bad);
  if (solo) {
    write_segment_out(mine, mx, kc.ja, kc.jb, n_pairs_seg, out, seg, P.out_packed, lane, omap, 2);
    if (lane == 0 && bad) atomicAdd(bad_counter, (unsigned long long)bad);
  } else {
    // partial record of this tile: [NS sums][max dd][bad], one entry per lane
    const uint32_t first = P.seg_first_tile[seg], seg_tiles = P.seg_first_tile[seg + 1] - first;
    double* part = P.partials + (size_t)part_slot * kMaxAcc;
    double pv = __shfl_sync(kFull, mine, 2 * (lane < NS ? lane : 0));
    if (lane == NS) pv = mx;
    if (lane == NS + 1) pv = (double)bad;
    if (lane < NS + 2) part[lane] = pv;
    __threadfence();
    __syncwarp();
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(&P.seg_counters[seg], 1u);
    ticket = __shfl_sync(kFull, ticket, 0);
    if (ticket == seg_tiles - 1) {   // last tile of this segment to finish: fold the partials in tile order
      __threadfence();
      double v = 0.0;
      if (lane < NS + 2) {
        for (uint32_t u = 0; u < seg_tiles; ++u) {
          const double x = __ldcg(P.partials + (size_t)(first + u) * kMaxAcc + lane);
          v = (lane == NS) ? fmax(v, x) : v + x;
        }
      }
      const double mx_all = __shfl_sync(kFull, v, NS);
      const double bad_all = __shfl_sync(kFull, v, NS + 1);
      write_segment_out(v, mx_all, kc.ja, kc.jb, n_pairs_seg, out, seg, P.out_packed, lane, omap, 1);
      if (lane == 0) {
        if (bad_all != 0.0) atomicAdd(bad_counter, (unsigned long long)bad_all);
        P.seg_counters[seg] = 0u;   // re-arm for the next launch
      }
    }
  }
}

// One lane's duo (two pairs sharing their moving cell) added into `acc` (H upper triangle, g, cost, sum dd), max dd, bad count.
template <int VARIANT, int LOSS, bool WANT_JAC, int NS>
__device__ __forceinline__ void accumulate_duo(const PoseConst& kc, const LossConst& lc, const float4* __restrict__ rec, const DuoRecFull* __restrict__ ovf,
                                               int lane, double* acc, double& max_dd, uint32_t& n_bad) {
  constexpr int NB = VarTraits<VARIANT>::NB;
  constexpr int NH = NB * (NB + 1) / 2;
  constexpr int NJ = WANT_JAC ? NH + NB : 0;
  CellC m, f[2];
  bool two;
  load_duo(rec, lane, ovf, m, f[0], f[1], two);
  Moving mv;
  moving_part<VARIANT>(kc, m, mv);
  double dd[2], N[2][4], wgt[2], hrho[2], wd[2];
  bool ok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) dd[j] = fixed_part<VARIANT, WANT_JAC>(kc, mv, f[j], N[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) ok[j] = dd_valid(dd[j]);
  const bool use1 = two && ok[1];
  // Nearly every duo is complete, well-conditioned and off the r = 0 corner: when that holds for all lanes that are here, loss and sums
  // run without the guards and per-accumulator selects of the general path (a warp-uniform branch).
  if (__all_sync(__activemask(), ok[0] && use1 && dd[0] > 0.0 && dd[1] > 0.0)) {
#pragma unroll
    for (int j = 0; j < 2; ++j) loss_eval<LOSS, true>(dd[j], lc, wgt[j], hrho[j], wd[j]);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (WANT_JAC) {
        int q = 0;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          const double wa = wd[j] * N[j][a];
#pragma unroll
          for (int b2 = a; b2 < NB; ++b2) { acc[q] = fma(wa, N[j][b2], acc[q]); ++q; }
          acc[NH + a] = fma(wgt[j], N[j][a], acc[NH + a]);
        }
      }
      acc[NJ] += hrho[j];
      acc[NJ + 1] += dd[j];
    }
    max_dd = fmax(max_dd, fmax(dd[0], dd[1]));
    return;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) loss_eval<LOSS>(dd[j], lc, wgt[j], hrho[j], wd[j]);
  n_bad += (ok[0] ? 0u : 1u) + ((two && !ok[1]) ? 1u : 0u);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (j == 0 ? ok[0] : use1) {
      if (WANT_JAC) {
        int q = 0;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          const double wa = wd[j] * N[j][a];
#pragma unroll
          for (int b2 = a; b2 < NB; ++b2) { acc[q] = fma(wa, N[j][b2], acc[q]); ++q; }
          acc[NH + a] = fma(wgt[j], N[j][a], acc[NH + a]);
        }
      }
      acc[NJ] += hrho[j];
      acc[NJ + 1] += dd[j];
      max_dd = fmax(max_dd, dd[j]);
    }
  }
}


}  // namespace

}  // namespace randt
