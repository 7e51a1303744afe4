// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or call anything under oracle/.
//
// PARITY UNPINNED: the reference (IGMR-RWTH/RaNDT-SLAM @ 1d995a5) has no tests,
// golden vectors or fixtures for this path and cannot be compiled here (Eigen,
// Ceres 2.1.0, Sophus 1.22.10, PCL, ROS are absent).  This file restates the
// forward-mode dual number that Ceres 2.1.0 (pinned in /root/reference/Dockerfile:11-17)
// uses for `AutoDiffCostFunction<..., 1, 4>` (reference call site:
// ros/ndt_radar_slam/src/ndt_registration/ndt_matcher.cpp:229-233), from the
// published semantics of ceres/jet.h: a value `a` plus an N-vector `v` of partials,
// with the operator formulas listed beside each function below.
#pragma once
#include <cmath>

namespace orc {

template <int N>
struct Jet {
  double a;
  double v[N];

  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT (implicit, like ceres)
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

// f + g, f - g : componentwise.
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h;
}
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h;
}
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h;
}
// f * g = (f.a g.a, f.a g.v + f.v g.a)
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h;
}
// f / g : with gi = 1/g.a and q = f.a*gi -> (q, (f.v - q g.v) * gi)
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi;
  h.a = q; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h;
}
// mixed scalar forms
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) {
  Jet<N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h;
}
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) {
  Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h;
}
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) {
  const double si = 1.0 / s; Jet<N> h; h.a = f.a * si; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * si; return h;
}
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  // s / g = (s/g.a, -s g.v / g.a^2)
  Jet<N> h; const double m = -s / (g.a * g.a); h.a = s / g.a; for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * m; return h;
}
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N>& operator-=(Jet<N>& f, const Jet<N>& g) { f = f - g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, const Jet<N>& g) { f = f * g; return f; }

// sqrt(f) = (t, f.v / (2 t)),  t = sqrt(f.a).  (t == 0 -> inf/NaN partials, as in ceres.)
template <int N> inline Jet<N> sqrt(const Jet<N>& f) {
  Jet<N> h; const double t = std::sqrt(f.a); const double w = 1.0 / (2.0 * t);
  h.a = t; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * w; return h;
}
template <int N> inline Jet<N> sin(const Jet<N>& f) {
  Jet<N> h; const double c = std::cos(f.a); h.a = std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h;
}
template <int N> inline Jet<N> cos(const Jet<N>& f) {
  Jet<N> h; const double s = -std::sin(f.a); h.a = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h;
}
// atan2(g, f) : (atan2(g.a, f.a), (-g.a f.v + f.a g.v) / (f.a^2 + g.a^2))
template <int N> inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
  Jet<N> h; const double t = 1.0 / (f.a * f.a + g.a * g.a); h.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; ++i) h.v[i] = t * (-g.a * f.v[i] + f.a * g.v[i]); return h;
}
// floor: value floor, zero partials (ceres defines it this way; used by NormalizeAngle).
template <int N> inline Jet<N> floor(const Jet<N>& f) { return Jet<N>(std::floor(f.a)); }

// double overloads so templated code can call orc::sqrt etc. on plain doubles
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double floor(double x) { return std::floor(x); }

inline double value_of(double x) { return x; }
template <int N> inline double value_of(const Jet<N>& f) { return f.a; }

}  // namespace orc
