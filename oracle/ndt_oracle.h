// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or call anything under oracle/.
//
// PARITY UNPINNED.  The reference (IGMR-RWTH/RaNDT-SLAM @ 1d995a5) ships no tests, golden
// vectors or fixtures for this path, and cannot be compiled in this image: every hot-path
// file includes Eigen + PCL (ndt_cell.h:4-8) or Ceres 2.1.0 + Sophus 1.22.10
// (ceres_residuals.h:9-18), none of which are on disk (versions: /root/reference/Dockerfile:11,20;
// Eigen 3.3.7 from ros:noetic-perception).  This header is a CPU restatement of the
// reference's algorithm for the NDT voxelise -> associate -> residual/Jacobian path, written
// against flat arrays (no Eigen/PCL types).  Its own pinning is (i) three-way agreement of the
// dual-number path, the closed form and central finite differences, and (ii) the independent
// mpmath fixtures under tests/golden/ (generator committed beside them).
//
// Path shorthand: R/ = /root/reference/ros/ndt_radar_slam/
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

#include "jet.h"

namespace orc {

// ------------------------------------------------------------------------------------------
// Plain data
// ------------------------------------------------------------------------------------------
struct Pt4 { float x, y, z, i; };            // PointXYZI as float4 (z unused by the path)
struct Cell12 { float mu[3]; float cov[9]; };  // mean (x, y, intensity) + row-major 3x3 covariance

struct MapGeom {
  // R/src/ndt_representation/ndt_map.cpp:7-21 (Map::initialize)
  int size_x = 0, size_y = 0;   // cells per side, AFTER the `size /= resolution` truncation (R/src/ndt_slam/ndt_slam.cpp:653-654)
  double res = 1.0;
  double max_linf = 0.0;        // max_neighbour_manhattan_distance
  double off_x = 0.0, off_y = 0.0;
  uint32_t n_slots() const { return (uint32_t)size_x * (uint32_t)size_y; }
  void set(int sx, int sy, double r, double linf, double cx = 0.0, double cy = 0.0) {
    size_x = sx; size_y = sy; res = r; max_linf = linf;
    off_x = -static_cast<double>((unsigned)sx) / 2.0 * r + cx;
    off_y = -static_cast<double>((unsigned)sy) / 2.0 * r + cy;
  }
};

// static_cast<unsigned int>(double) of a negative value is UB in C++; x86-64 gcc emits cvttsd2si
// (64-bit) and keeps the low 32 bits.  The oracle (and the CUDA path) define it that way.
inline uint32_t to_u32_trunc(double v) {
  if (!(v > -9.2e18 && v < 9.2e18)) return 0u;  // NaN / out of int64 range: defined as 0
  return (uint32_t)(int64_t)v;
}

// R/include/ndt_representation/ndt_map.h:87-90,181-184 (coordinateToIndex / getIndex)
inline uint32_t coord_to_index(const MapGeom& g, float x, float y) {
  const uint32_t mx = to_u32_trunc(((double)x - g.off_x) / g.res);
  const uint32_t my = to_u32_trunc(((double)y - g.off_y) / g.res);
  return my * (uint32_t)g.size_x + mx;  // unsigned wrap, no per-axis bound check (quirk B.8)
}

// ------------------------------------------------------------------------------------------
// a1  Grid::cluster            R/src/radar_preprocessing/grid.cpp:7-14
// ------------------------------------------------------------------------------------------
inline int n_clusters_from_params(double max_range, double resolution) {
  // R/src/ndt_slam/ndt_slam.cpp:691 — double pow truncated into an int field
  return (int)std::pow(2.0 * max_range / resolution, 2);
}
inline int grid_row_size(size_t n_clusters) { return static_cast<int>(std::sqrt((double)n_clusters)); }
inline float grid_resolution(size_t n_clusters, float max_range) {
  return max_range * 2 / (grid_row_size(n_clusters));
}
inline void grid_labels(const Pt4* pts, size_t n, size_t n_clusters, float max_range, int32_t* labels) {
  const int row = grid_row_size(n_clusters);
  const float res = max_range * 2 / row;
  for (size_t k = 0; k < n; ++k)
    labels[k] = static_cast<int>(pts[k].x / res) + row * static_cast<int>(pts[k].y / res);  // trunc toward zero (quirk B.1)
}

// ------------------------------------------------------------------------------------------
// Eigen 3.3.7 SelfAdjointEigenSolver<Matrix2f>(m) restated (float).  The constructor runs
// compute(m, ComputeEigenvectors): scale the lower triangle by its max |coeff|, Householder
// tridiagonalisation (a no-op for n = 2: diag = (a00, a11), sub = a10, Q = I), implicit
// symmetric QR steps with Wilkinson shift, ascending sort.  Used by R/src/ndt_representation/ndt_cell.cpp:104-110.
// V is returned as V[row][col]; column j is the eigenvector of ev[j].
// ------------------------------------------------------------------------------------------
inline float eig_hypot_f(float x, float y) {
  float ax = std::fabs(x), ay = std::fabs(y), p, qp;
  if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
  if (p == 0.0f) return 0.0f;
  return p * std::sqrt(1.0f + qp * qp);
}
inline void givens_f(float p, float q, float& c, float& s) {
  if (q == 0.0f) { c = p < 0.0f ? -1.0f : 1.0f; s = 0.0f; }
  else if (p == 0.0f) { c = 0.0f; s = q < 0.0f ? 1.0f : -1.0f; }
  else if (std::fabs(p) > std::fabs(q)) {
    float t = q / p; float u = std::sqrt(1.0f + t * t); if (p < 0.0f) u = -u;
    c = 1.0f / u; s = -t * c;
  } else {
    float t = p / q; float u = std::sqrt(1.0f + t * t); if (q < 0.0f) u = -u;
    s = -1.0f / u; c = -t * s;
  }
}
inline void sym2_eigen_f(float a00, float a10, float a11, float ev[2], float V[2][2]) {
  float scale = std::max(std::max(std::fabs(a00), std::fabs(a10)), std::fabs(a11));
  if (scale == 0.0f) scale = 1.0f;
  float d0 = a00 / scale, d1 = a11 / scale, e = a10 / scale;
  V[0][0] = 1.0f; V[0][1] = 0.0f; V[1][0] = 0.0f; V[1][1] = 1.0f;
  const float tiny = std::numeric_limits<float>::min();
  const float prec = 2.0f * std::numeric_limits<float>::epsilon();
  int iter = 0;
  while (true) {
    if (std::fabs(e) <= (std::fabs(d0) + std::fabs(d1)) * prec || std::fabs(e) <= tiny) e = 0.0f;
    if (e == 0.0f) break;
    ++iter;
    if (iter > 30 * 2) break;
    // one implicit QR step on the 2x2 block
    const float td = (d0 - d1) * 0.5f;
    float mu = d1;
    if (td == 0.0f) {
      mu -= std::fabs(e);
    } else {
      const float e2 = e * e;
      const float h = eig_hypot_f(td, e);
      if (e2 == 0.0f) mu -= (e / (td + (td > 0.0f ? 1.0f : -1.0f))) * (e / h);
      else            mu -= e2 / (td + (td > 0.0f ? h : -h));
    }
    const float x = d0 - mu, z = e;
    float c, s;
    givens_f(x, z, c, s);
    const float sdk = s * d0 + c * e;
    const float dkp1 = s * e + c * d1;
    const float nd0 = c * (c * d0 - s * e) - s * (c * e - s * d1);
    const float nd1 = s * sdk + c * dkp1;
    const float ne = c * sdk - s * dkp1;
    d0 = nd0; d1 = nd1; e = ne;
    if (!(c == 1.0f && s == 0.0f)) {
      for (int r = 0; r < 2; ++r) {
        const float xi = V[r][0], yi = V[r][1];
        V[r][0] = c * xi + (-s) * yi;
        V[r][1] = s * xi + c * yi;
      }
    }
  }
  if (d1 < d0) {  // ascending sort (minCoeff picks the first minimum on ties)
    std::swap(d0, d1);
    std::swap(V[0][0], V[0][1]);
    std::swap(V[1][0], V[1][1]);
  }
  ev[0] = d0 * scale; ev[1] = d1 * scale;
}

// ------------------------------------------------------------------------------------------
// a4  Cell::updateCell          R/src/ndt_representation/ndt_cell.cpp:36-114
// One cluster -> (mean, cov) in float32, sequential accumulation in point order, population
// covariance (/n), eigenvalue floor on the xy block, +1e-6 on the intensity variance.
// pNDT (use_pndt) is off in every shipped config and is not restated.
// Returns false when the cluster is rejected (n <= min_points, strict: ndt_cell.cpp:26,37).
// ------------------------------------------------------------------------------------------
inline void regularize_cell(Cell12& c) {
  // ndt_cell.cpp:102-112
  float ev[2], V[2][2];
  sym2_eigen_f(c.cov[0], c.cov[3], c.cov[4], ev, V);  // lower triangle of the xy block: (0,0),(1,0),(1,1)
  ev[0] = std::max(ev[0], 0.001f * ev[1]);
  // eig_vectors * diag(ev) * eig_vectors.inverse(), evaluated left to right
  float T[2][2];
  for (int r = 0; r < 2; ++r) {
    T[r][0] = V[r][0] * ev[0] + V[r][1] * 0.0f;
    T[r][1] = V[r][0] * 0.0f + V[r][1] * ev[1];
  }
  const float det = V[0][0] * V[1][1] - V[1][0] * V[0][1];
  const float invdet = 1.0f / det;
  float I[2][2];
  I[0][0] = V[1][1] * invdet;  I[1][0] = -V[1][0] * invdet;
  I[0][1] = -V[0][1] * invdet; I[1][1] = V[0][0] * invdet;
  for (int r = 0; r < 2; ++r)
    for (int q = 0; q < 2; ++q)
      c.cov[r * 3 + q] = T[r][0] * I[0][q] + T[r][1] * I[1][q];
  c.cov[8] = (float)((double)c.cov[8] + 0.000001);
}

inline bool cell_from_points(const Pt4* pts, const uint32_t* idx, size_t n, int min_points, Cell12& out, float* max_intensity = nullptr) {
  if (!((size_t)n > (size_t)min_points) || n == 0) return false;
  float sx = 0.f, sy = 0.f, si = 0.f;
  double mi = 0.0;
  for (size_t k = 0; k < n; ++k) {
    const Pt4& p = pts[idx ? idx[k] : k];
    sx += p.x; sy += p.y; si += p.i;
    mi = std::max(mi, (double)p.i);
  }
  const float nf = (float)(double)n;  // `/= static_cast<double>(n)` on a float vector converts to float first
  const float mx = sx / nf, my = sy / nf, mz = si / nf;
  float c00 = 0.f, c11 = 0.f, c22 = 0.f, c01 = 0.f, c02 = 0.f, c12 = 0.f;
  for (size_t k = 0; k < n; ++k) {
    const Pt4& p = pts[idx ? idx[k] : k];
    const float dx = p.x - mx, dy = p.y - my, di = p.i - mz;
    c00 += dx * dx; c11 += dy * dy; c22 += di * di;
    c01 += dx * dy; c02 += dx * di; c12 += dy * di;
  }
  out.mu[0] = mx; out.mu[1] = my; out.mu[2] = mz;
  out.cov[0] = c00 / nf; out.cov[1] = c01 / nf; out.cov[2] = c02 / nf;
  out.cov[3] = c01 / nf; out.cov[4] = c11 / nf; out.cov[5] = c12 / nf;
  out.cov[6] = c02 / nf; out.cov[7] = c12 / nf; out.cov[8] = c22 / nf;
  regularize_cell(out);
  if (max_intensity) *max_intensity = (float)mi;
  return true;
}

// ------------------------------------------------------------------------------------------
// a2 + a3   ClusterGenerator::labelClouds  R/src/radar_preprocessing/radar_preprocessor.cpp:151-169
//           Map::insertCluster             R/src/ndt_representation/ndt_map.cpp:238-245
// A voxelised scan ("NDT map"): compact cell vector in ascending-label order + dense slot table.
// ------------------------------------------------------------------------------------------
struct NdtMap {
  MapGeom geom;
  std::vector<Cell12> cells;
  std::vector<uint32_t> npts;     // Cell::n_points_
  std::vector<int32_t> slot;      // grid_indizes_: -1 = empty
  int dropped_out_of_map = 0;     // reference would throw std::out_of_range (vector::at); oracle drops + counts
  void init(const MapGeom& g) { geom = g; cells.clear(); npts.clear(); slot.assign(g.n_slots(), -1); dropped_out_of_map = 0; }
};

inline void voxelize(const Pt4* pts, size_t n, size_t n_clusters, float max_range, int min_points,
                     const MapGeom& geom, NdtMap& map, std::vector<int32_t>* cell_labels = nullptr) {
  map.init(geom);
  std::vector<int32_t> labels(n);
  grid_labels(pts, n, n_clusters, max_range, labels.data());
  // sort + unique -> cluster rank (ascending signed label); points keep original order inside a cluster
  std::vector<int32_t> uniq(labels);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  std::vector<std::vector<uint32_t>> members(uniq.size());
  for (size_t k = 0; k < n; ++k) {
    const size_t c = std::lower_bound(uniq.begin(), uniq.end(), labels[k]) - uniq.begin();
    members[c].push_back((uint32_t)k);
  }
  if (cell_labels) cell_labels->clear();
  for (size_t c = 0; c < uniq.size(); ++c) {
    Cell12 cell;
    if (!cell_from_points(pts, members[c].data(), members[c].size(), min_points, cell)) continue;
    const uint32_t s = coord_to_index(geom, cell.mu[0], cell.mu[1]);
    if (s >= geom.n_slots()) { map.dropped_out_of_map++; continue; }
    map.slot[s] = (int32_t)map.cells.size();  // a later cluster overwrites an earlier one (quirk B.7)
    map.cells.push_back(cell);
    map.npts.push_back((uint32_t)members[c].size());
    if (cell_labels) cell_labels->push_back(uniq[c]);
  }
}

// ------------------------------------------------------------------------------------------
// Sophus 1.22.10 SE2d::cast<float>() (sophus/se2.hpp:cast, so2.hpp: SO2(complex) constructor -> normalize()):
// the unit complex is cast to float and re-normalised (length = hypot(re, im); complex /= length), the translation is cast.
// This is what every `Eigen::Affine2f(x.cast<float>().matrix())` in the reference holds
// (R/src/ndt_registration/ndt_matcher.cpp:208, R/src/local_fuser/local_fuser.cpp:175,280,338).  out = (re, im, tx, ty).
// ------------------------------------------------------------------------------------------
inline void se2d_cast_float(const double pose[4], float out[4]) {
  float re = static_cast<float>(pose[0]), im = static_cast<float>(pose[1]);
  const float length = std::hypot(re, im);   // glibc hypotf: (float)sqrt((double)re*re + (double)im*im)
  re /= length; im /= length;
  out[0] = re; out[1] = im; out[2] = static_cast<float>(pose[2]); out[3] = static_cast<float>(pose[3]);
}

// ------------------------------------------------------------------------------------------
// Eigen 3.3.7 Transform<float,3,Affine>::rotation() of the lift Cell::transformCell builds (ndt_cell.cpp:118-122):
// linear part L = [[c,-s,0],[s,c,0],[0,0,1]].  rotation() = computeRotationScaling(&R, 0) (Geometry/Transform.h): JacobiSVD<Matrix3f>
// (two-sided Jacobi, SVD/JacobiSVD.h compute() + real_2x2_jacobi_svd + Jacobi.h makeJacobi / apply_rotation_in_the_plane),
// x = det(U V^T), U.col(0) /= x, R = U V^T.  R equals L only up to float rounding (and differs from it for about half of all angles).
// ------------------------------------------------------------------------------------------
namespace eig3f {
struct Rot { float c, s; };
inline bool make_jacobi(float x, float y, float z, Rot& r) {
  const float deno = 2.0f * std::fabs(y);
  if (deno < std::numeric_limits<float>::min()) { r.c = 1.0f; r.s = 0.0f; return false; }
  const float tau = (x - z) / deno;
  const float w = std::sqrt(tau * tau + 1.0f);
  float t;
  if (tau > 0.0f) t = 1.0f / (tau + w); else t = 1.0f / (tau - w);
  const float sign_t = t > 0.0f ? 1.0f : -1.0f;
  const float n = 1.0f / std::sqrt(t * t + 1.0f);
  r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r.c = n;
  return true;
}
// apply_rotation_in_the_plane on rows p, q (applyOnTheLeft) / columns p, q with the transposed rotation (applyOnTheRight)
inline void rot_rows(float m[3][3], int p, int q, Rot j, int ncols = 3) {
  if (j.c == 1.0f && j.s == 0.0f) return;
  for (int k = 0; k < ncols; ++k) { const float xi = m[p][k], yi = m[q][k]; m[p][k] = j.c * xi + j.s * yi; m[q][k] = -j.s * xi + j.c * yi; }
}
inline void rot_cols(float m[3][3], int p, int q, Rot j) {
  const float c = j.c, s = -j.s;     // j.transpose()
  if (c == 1.0f && s == 0.0f) return;
  for (int k = 0; k < 3; ++k) { const float xi = m[k][p], yi = m[k][q]; m[k][p] = c * xi + s * yi; m[k][q] = -s * xi + c * yi; }
}
inline float det3(const float m[3][3]) {
  auto h = [&](int a, int b, int c) { return m[0][a] * (m[1][b] * m[2][c] - m[1][c] * m[2][b]); };
  return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}
inline void jacobi_svd3(const float a[3][3], float U[3][3], float V[3][3], float sv[3]) {
  const float precision = 2.0f * std::numeric_limits<float>::epsilon();
  const float consider_as_zero = std::numeric_limits<float>::min();
  float scale = 0.0f;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(a[i][j]));
  if (scale == 0.0f) scale = 1.0f;
  float w[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { w[i][j] = a[i][j] / scale; U[i][j] = V[i][j] = (i == j) ? 1.0f : 0.0f; }
  float max_diag = 0.0f;
  for (int i = 0; i < 3; ++i) max_diag = std::max(max_diag, std::fabs(w[i][i]));
  bool finished = false;
  while (!finished) {
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        const float threshold = std::max(consider_as_zero, precision * max_diag);
        if (std::fabs(w[p][q]) > threshold || std::fabs(w[q][p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd
          float m[3][3] = {{w[p][p], w[p][q], 0.f}, {w[q][p], w[q][q], 0.f}, {0.f, 0.f, 0.f}};
          Rot rot1;
          const float t = m[0][0] + m[1][1];
          const float d = m[1][0] - m[0][1];
          if (std::fabs(d) < std::numeric_limits<float>::min()) { rot1.s = 0.0f; rot1.c = 1.0f; }
          else { const float u = t / d; const float tmp = std::sqrt(1.0f + u * u); rot1.s = 1.0f / tmp; rot1.c = u / tmp; }
          rot_rows(m, 0, 1, rot1, 2);
          Rot j_right;
          make_jacobi(m[0][0], m[0][1], m[1][1], j_right);
          const Rot jrt = {j_right.c, -j_right.s};
          const Rot j_left = {rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};   // rot1 * j_right.transpose()
          rot_rows(w, p, q, j_left);
          const Rot jlt = {j_left.c, -j_left.s};
          rot_cols(U, p, q, jlt);
          rot_cols(w, p, q, j_right);
          rot_cols(V, p, q, j_right);
          max_diag = std::max(max_diag, std::max(std::fabs(w[p][p]), std::fabs(w[q][q])));
        }
      }
  }
  for (int i = 0; i < 3; ++i) {
    sv[i] = std::fabs(w[i][i]);
    if (w[i][i] < 0.0f) for (int k = 0; k < 3; ++k) U[k][i] = -U[k][i];
  }
  for (int i = 0; i < 3; ++i) sv[i] = sv[i] * scale;
  for (int i = 0; i < 3; ++i) {      // descending sort, first maximum wins
    int pos = 0; float best = sv[i];
    for (int k = 1; k < 3 - i; ++k) if (sv[i + k] > best) { best = sv[i + k]; pos = k; }
    if (best == 0.0f) break;
    if (pos) {
      pos += i;
      std::swap(sv[i], sv[pos]);
      for (int k = 0; k < 3; ++k) { std::swap(U[k][pos], U[k][i]); std::swap(V[k][pos], V[k][i]); }
    }
  }
}
// 3x3 float product, coefficient = x0 + (x1 + x2)  (Eigen's unrolled reduction of three terms)
inline void mul3(const float A[3][3], const float B[3][3], float C[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + (A[i][1] * B[1][j] + A[i][2] * B[2][j]);
}
inline void transpose3(const float A[3][3], float T[3][3]) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[i][j] = A[j][i]; }
}  // namespace eig3f

inline void affine_rotation_f(float c, float s, float R[3][3]) {
  const float L[3][3] = {{c, -s, 0.f}, {s, c, 0.f}, {0.f, 0.f, 1.f}};
  float U[3][3], V[3][3], Vt[3][3], UVt[3][3], sv[3];
  eig3f::jacobi_svd3(L, U, V, sv);
  eig3f::transpose3(V, Vt);
  eig3f::mul3(U, Vt, UVt);
  const float x = eig3f::det3(UVt);
  for (int k = 0; k < 3; ++k) U[k][0] = U[k][0] / x;
  eig3f::mul3(U, Vt, R);
}

// ------------------------------------------------------------------------------------------
// a5  Cell::transformCell / Map::transformMap   R/src/ndt_representation/ndt_cell.cpp:117-123, ndt_map.cpp:177-182
// float32.  trans = [[c,-s,tx],[s,c,ty]] (the Affine2f the caller built).  mean <- trans_3d * mean (translation + linear * mean, the
// linear part as given); cov <- R cov R^T with R = trans_3d.rotation(), Eigen's SVD polar factor of the linear part (above).
// ------------------------------------------------------------------------------------------
inline void transform_cell(Cell12& cell, float c, float s, float tx, float ty, const float R[3][3]) {
  const float x = cell.mu[0], y = cell.mu[1], in = cell.mu[2];
  cell.mu[0] = tx + (c * x + ((-s) * y + 0.f * in));
  cell.mu[1] = ty + (s * x + (c * y + 0.f * in));
  cell.mu[2] = 0.f + (0.f * x + (0.f * y + 1.f * in));
  float S[3][3], T[3][3], Rt[3][3], O[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = cell.cov[i * 3 + j];
  eig3f::mul3(R, S, T);
  eig3f::transpose3(R, Rt);
  eig3f::mul3(T, Rt, O);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cell.cov[i * 3 + j] = O[i][j];
}
inline void transform_cell(Cell12& cell, float c, float s, float tx, float ty) {
  float R[3][3];
  affine_rotation_f(c, s, R);
  transform_cell(cell, c, s, tx, ty, R);
}

// ------------------------------------------------------------------------------------------
// a6  Cell::operator+= / Map::mergeMapCell   R/include/ndt_representation/ndt_cell.h:133-142, ndt_map.cpp:191-207
// ------------------------------------------------------------------------------------------
inline void merge_cell(Cell12& a, uint32_t& na, const Cell12& b, uint32_t nb) {
  const float w1 = (float)(na - 1u);                         // unsigned int -> float
  const float w2 = (float)((size_t)nb - 1);                  // size_t -> float
  const float w3 = (float)(((size_t)na * (size_t)nb) / ((size_t)na + (size_t)nb));  // integer division (quirk B.4)
  float d[3] = {a.mu[0] - b.mu[0], a.mu[1] - b.mu[1], a.mu[2] - b.mu[2]};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      a.cov[i * 3 + j] = (w1 * a.cov[i * 3 + j] + w2 * b.cov[i * 3 + j]) + w3 * (d[i] * d[j]);
  const float n1 = (float)na, n2 = (float)(size_t)nb, nn = (float)((size_t)na + (size_t)nb);
  for (int i = 0; i < 3; ++i) a.mu[i] = ((a.mu[i] * n1) + (b.mu[i] * n2)) / nn;
  na += nb;
  const float dn = (float)(na - 1u);
  for (int i = 0; i < 9; ++i) a.cov[i] /= dn;
}

inline void merge_map_cell(NdtMap& fixed, const NdtMap& moving) {
  for (size_t i = 0; i < moving.cells.size(); ++i) {
    const Cell12& m = moving.cells[i];
    const uint32_t s = coord_to_index(fixed.geom, m.mu[0], m.mu[1]);
    if (s < fixed.slot.size()) {
      const int32_t idx = fixed.slot[s];
      if (idx >= 0) {
        merge_cell(fixed.cells[idx], fixed.npts[idx], m, moving.npts[i]);
      } else {
        fixed.cells.push_back(m);
        fixed.npts.push_back(moving.npts[i]);
        fixed.slot[s] = (int32_t)fixed.cells.size() - 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// a7  Map::getClosestCells (+ getAdjacentIndizes, mahalanobisSquaredIntensity)
//     R/src/ndt_representation/ndt_map.cpp:101-151,163-175 ; ndt_cell.cpp:172-176
// ------------------------------------------------------------------------------------------
inline float inv3_quadform_f(const float m[9], const float v[3]) {
  // Eigen fixed-size 3x3 float inverse by cofactors (compute_inverse_size3), then (v^T * inv) * v
  auto M = [&](int r, int c) { return m[r * 3 + c]; };
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
  };
  const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const float det = c0 * M(0, 0) + (c1 * M(1, 0) + c2 * M(2, 0));
  const float invdet = 1.0f / det;
  float inv[3][3];
  inv[0][0] = c0 * invdet; inv[0][1] = c1 * invdet; inv[0][2] = c2 * invdet;
  inv[1][0] = cof(0, 1) * invdet; inv[1][1] = cof(1, 1) * invdet; inv[1][2] = cof(2, 1) * invdet;
  inv[2][0] = cof(0, 2) * invdet; inv[2][1] = cof(1, 2) * invdet; inv[2][2] = cof(2, 2) * invdet;
  float t[3];
  for (int j = 0; j < 3; ++j) t[j] = v[0] * inv[0][j] + (v[1] * inv[1][j] + v[2] * inv[2][j]);
  return t[0] * v[0] + (t[1] * v[1] + t[2] * v[2]);
}

inline double mahalanobis_sq_intensity(const Cell12& query, const Cell12& other) {
  float S[9], d[3];
  for (int i = 0; i < 9; ++i) S[i] = other.cov[i] + query.cov[i];
  for (int i = 0; i < 3; ++i) d[i] = other.mu[i] - query.mu[i];
  return (double)inv3_quadform_f(S, d);
}

// window of side 2r+1 around `index`, reference traversal order, unsigned wrap, de-duplicated
inline void adjacent_indices(const MapGeom& g, uint32_t index, int r, std::vector<uint32_t>& out) {
  out.clear();
  const uint32_t n_slots = g.n_slots();
  for (int i = -r; i <= r; ++i)
    for (int j = -r; j <= r; ++j) {
      const uint32_t ni = index + (uint32_t)i + (uint32_t)j * (uint32_t)g.size_x;
      if (ni < n_slots && std::find(out.begin(), out.end(), ni) == out.end()) out.push_back(ni);
    }
}

enum LookupMetric { LOOKUP_MAHALANOBIS_INTENSITY = 0, LOOKUP_EUCLID_XY = 1 };

// query is already expressed in the fixed map's frame
inline void closest_cells(const NdtMap& fixed, const Cell12& query, int k, int metric, std::vector<uint32_t>& out) {
  std::vector<std::pair<double, size_t>> targets;
  const uint32_t center = coord_to_index(fixed.geom, query.mu[0], query.mu[1]);
  int r = 0;
  std::vector<uint32_t> adj;
  const int r_stop = static_cast<int>(fixed.geom.max_linf / fixed.geom.res);
  while (targets.size() < (size_t)k && adj.size() < fixed.geom.n_slots()) {
    targets.clear();
    adjacent_indices(fixed.geom, center, r, adj);
    for (size_t a = 0; a < adj.size(); ++a) {
      const int32_t ci = fixed.slot[adj[a]];
      if (ci >= 0) {
        double dist;
        if (metric == LOOKUP_MAHALANOBIS_INTENSITY) {
          dist = mahalanobis_sq_intensity(query, fixed.cells[ci]);
        } else {
          const float dx = query.mu[0] - fixed.cells[ci].mu[0], dy = query.mu[1] - fixed.cells[ci].mu[1];
          dist = (double)std::sqrt(dx * dx + dy * dy);
        }
        targets.push_back(std::make_pair(dist, (size_t)ci));
      }
    }
    ++r;
    if (r >= r_stop) break;
  }
  std::sort(targets.begin(), targets.end());
  const size_t m = std::min((size_t)std::max(k, 0), targets.size());
  out.clear();
  for (size_t t = 0; t < m; ++t) out.push_back((uint32_t)targets[t].second);
}

// ------------------------------------------------------------------------------------------
// a8  Matcher::addNDTFactor — association half      R/src/ndt_registration/ndt_matcher.cpp:183-288
// pose = Sophus SE2d storage [cos, sin, tx, ty].  Emits the residual-block list in reference
// order: for each moving cell (ascending), for each neighbour (ascending (dist, index)).
// ------------------------------------------------------------------------------------------
struct PairList {
  std::vector<uint32_t> im, jf;
};
inline void associate(const NdtMap& fixed, const NdtMap& moving, const double pose[4], int k, int metric, PairList& out) {
  float a[4];
  se2d_cast_float(pose, a);        // initial_guess.cast<float>() (ndt_matcher.cpp:208,213)
  const float c = a[0], s = a[1], tx = a[2], ty = a[3];
  float R[3][3];
  affine_rotation_f(c, s, R);
  std::vector<uint32_t> nn;
  for (size_t i = 0; i < moving.cells.size(); ++i) {
    Cell12 q = moving.cells[i];
    if (metric == LOOKUP_MAHALANOBIS_INTENSITY) {
      transform_cell(q, c, s, tx, ty, R);
    } else {
      // initial_guess.cast<float>() * mean_xy  (Sophus SE2f * point = R p + t)
      const float x = q.mu[0], y = q.mu[1];
      q.mu[0] = (c * x - s * y) + tx;
      q.mu[1] = (s * x + c * y) + ty;
    }
    closest_cells(fixed, q, k, metric, nn);
    for (uint32_t j : nn) { out.im.push_back((uint32_t)i); out.jf.push_back(j); }
  }
}

// ------------------------------------------------------------------------------------------
// a9/a10  NDT residual functors          R/include/ndt_registration/ceres_residuals.h:421-552
// Templated on T = double or Jet<N> exactly like the reference functors are templated for ceres.
// ------------------------------------------------------------------------------------------
template <typename T>
inline void inv3(const T m[3][3], T inv[3][3]) {
  // Eigen cofactor inverse (same structure as the float version above, generic scalar)
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  const T c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const T det = c0 * m[0][0] + (c1 * m[1][0] + c2 * m[2][0]);
  const T invdet = T(1.0) / det;
  inv[0][0] = c0 * invdet; inv[0][1] = c1 * invdet; inv[0][2] = c2 * invdet;
  inv[1][0] = cof(0, 1) * invdet; inv[1][1] = cof(1, 1) * invdet; inv[1][2] = cof(2, 1) * invdet;
  inv[2][0] = cof(0, 2) * invdet; inv[2][1] = cof(1, 2) * invdet; inv[2][2] = cof(2, 2) * invdet;
}
template <typename T>
inline void inv2(const T m[2][2], T inv[2][2]) {
  const T det = m[0][0] * m[1][1] - m[1][0] * m[0][1];
  const T invdet = T(1.0) / det;
  inv[0][0] = m[1][1] * invdet; inv[1][0] = -m[1][0] * invdet;
  inv[0][1] = -m[0][1] * invdet; inv[1][1] = m[0][0] * invdet;
}

// NormalizeAngle  R/include/ndt_registration/state_manifold.h:17-23
template <typename T>
inline T normalize_angle(const T& a) {
  const T two_pi(2.0 * M_PI);
  return a - two_pi * floor((a + T(M_PI)) / two_pi);
}

// Eigen AngleAxis<T>(angle, z).toRotationMatrix()
template <typename T>
inline void rot_z3(const T& angle, T R[3][3]) {
  const T sn = sin(angle), c = cos(angle);
  const T one_c = T(1.0) - c;
  R[0][0] = T(0.0) + c; R[0][1] = T(0.0) - sn; R[0][2] = T(0.0);
  R[1][0] = T(0.0) + sn; R[1][1] = T(0.0) + c; R[1][2] = T(0.0);
  R[2][0] = T(0.0); R[2][1] = T(0.0); R[2][2] = one_c * T(1.0) + c;
}

struct PairConst3 {  // what the reference functor snapshots: float -> double casts (ndt_matcher.cpp:231)
  double mm[3], mc[3][3], fm[3], fc[3][3];
  PairConst3(const Cell12& m, const Cell12& f) {
    for (int i = 0; i < 3; ++i) { mm[i] = (double)m.mu[i]; fm[i] = (double)f.mu[i]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { mc[i][j] = (double)m.cov[i * 3 + j]; fc[i][j] = (double)f.cov[i * 3 + j]; }
  }
};

// shared tail: r = sqrt(delta^T (R mc R^T + fc)^-1 delta), 3-D
template <typename T>
inline T ndt_residual3(const T R[3][3], const T t[3], const PairConst3& k) {
  T d[3];
  for (int i = 0; i < 3; ++i) d[i] = (R[i][0] * T(k.mm[0]) + (R[i][1] * T(k.mm[1]) + R[i][2] * T(k.mm[2]))) + t[i] - T(k.fm[i]);
  T RS[3][3], B[3][3], Bi[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    RS[i][j] = R[i][0] * T(k.mc[0][j]) + (R[i][1] * T(k.mc[1][j]) + R[i][2] * T(k.mc[2][j]));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    B[i][j] = (RS[i][0] * R[j][0] + (RS[i][1] * R[j][1] + RS[i][2] * R[j][2])) + T(k.fc[i][j]);
  inv3(B, Bi);
  T row[3];
  for (int j = 0; j < 3; ++j) row[j] = d[0] * Bi[0][j] + (d[1] * Bi[1][j] + d[2] * Bi[2][j]);
  const T q = row[0] * d[0] + (row[1] * d[1] + row[2] * d[2]);
  return sqrt(q);
}

enum ResidualVariant {
  VAR_SE2_INTENSITY = 0,   // NDTFrameToMapIntensityFactorResidualSE2  ceres_residuals.h:520-552 (LIVE in all shipped configs)
  VAR_SE2_XY = 1,          // NDTFrameToMapFactorResidualSE2           ceres_residuals.h:454-484
  VAR_VEC_INTENSITY = 2,   // NDTFrameToMapIntensityFactorResidual     ceres_residuals.h:486-518   params (x, y | theta)
  VAR_VEC_XY = 3           // NDTFrameToMapFactorResidual              ceres_residuals.h:421-451   params (x, y | theta)
};

// pose = [cos, sin, tx, ty]; theta = so2().log() = atan2(sin, cos)
template <typename T>
inline T residual_se2_intensity(const T pose[4], const PairConst3& k) {
  const T t[3] = {pose[2], pose[3], T(0.0)};
  const T theta = atan2(pose[1], pose[0]);
  T R[3][3];
  rot_z3(theta, R);
  return ndt_residual3(R, t, k);
}
// params = [x, y, theta]
template <typename T>
inline T residual_vec_intensity(const T p[3], const PairConst3& k) {
  const T t[3] = {p[0], p[1], T(0.0)};
  T R[3][3];
  rot_z3(normalize_angle(p[2]), R);
  return ndt_residual3(R, t, k);
}
template <typename T>
inline T ndt_residual2(const T R[2][2], const T t[2], const PairConst3& k) {
  T d[2];
  for (int i = 0; i < 2; ++i) d[i] = (R[i][0] * T(k.mm[0]) + R[i][1] * T(k.mm[1])) + t[i] - T(k.fm[i]);
  T RS[2][2], B[2][2], Bi[2][2];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) RS[i][j] = R[i][0] * T(k.mc[0][j]) + R[i][1] * T(k.mc[1][j]);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) B[i][j] = (RS[i][0] * R[j][0] + RS[i][1] * R[j][1]) + T(k.fc[i][j]);
  inv2(B, Bi);
  const T r0 = d[0] * Bi[0][0] + d[1] * Bi[1][0];
  const T r1 = d[0] * Bi[0][1] + d[1] * Bi[1][1];
  return sqrt(r0 * d[0] + r1 * d[1]);
}
// Sophus SE2 * point and rotationMatrix() use the stored (un-normalised) complex [c -s; s c]
template <typename T>
inline T residual_se2_xy(const T pose[4], const PairConst3& k) {
  const T R[2][2] = {{pose[0], -pose[1]}, {pose[1], pose[0]}};
  const T t[2] = {pose[2], pose[3]};
  return ndt_residual2(R, t, k);
}
template <typename T>
inline T residual_vec_xy(const T p[3], const PairConst3& k) {
  const T a = normalize_angle(p[2]);
  const T sn = sin(a), c = cos(a);
  const T R[2][2] = {{c, -sn}, {sn, c}};   // Eigen::Rotation2D::toRotationMatrix
  const T t[2] = {p[0], p[1]};
  return ndt_residual2(R, t, k);
}

inline int variant_num_params(int variant) { return (variant == VAR_SE2_INTENSITY || variant == VAR_SE2_XY) ? 4 : 3; }

// What ceres::AutoDiffCostFunction<F,1,...>::Evaluate returns for one residual block:
// residual value and the 1 x n_params row (ambient parameters).  Autodiff = Jet path.
inline double eval_pair_value(int variant, const double* params, const Cell12& m, const Cell12& f);
inline void eval_pair_autodiff(int variant, const double* params, const Cell12& m, const Cell12& f, double* r, double* J) {
  // ceres::AutoDiffCostFunction::Evaluate calls the functor with plain doubles when no Jacobian is requested
  if (!J) { *r = eval_pair_value(variant, params, m, f); return; }
  const PairConst3 k(m, f);
  if (variant == VAR_SE2_INTENSITY || variant == VAR_SE2_XY) {
    Jet<4> p[4];
    for (int i = 0; i < 4; ++i) p[i] = Jet<4>(params[i], i);
    const Jet<4> res = (variant == VAR_SE2_INTENSITY) ? residual_se2_intensity(p, k) : residual_se2_xy(p, k);
    *r = res.a;
    if (J) for (int i = 0; i < 4; ++i) J[i] = res.v[i];
  } else {
    Jet<3> p[3];
    for (int i = 0; i < 3; ++i) p[i] = Jet<3>(params[i], i);
    const Jet<3> res = (variant == VAR_VEC_INTENSITY) ? residual_vec_intensity(p, k) : residual_vec_xy(p, k);
    *r = res.a;
    if (J) for (int i = 0; i < 3; ++i) J[i] = res.v[i];
  }
}
inline double eval_pair_value(int variant, const double* params, const Cell12& m, const Cell12& f) {
  const PairConst3 k(m, f);
  switch (variant) {
    case VAR_SE2_INTENSITY: return residual_se2_intensity<double>(params, k);
    case VAR_SE2_XY:        return residual_se2_xy<double>(params, k);
    case VAR_VEC_INTENSITY: return residual_vec_intensity<double>(params, k);
    default:                return residual_vec_xy<double>(params, k);
  }
}

// Closed form (generalised to a non-symmetric covariance), VAR_SE2_INTENSITY only.
// Independent of the Jet path; used for the three-way self check and as the "hand-optimised CPU" baseline.
inline void eval_pair_closed_form(const double pose[4], const Cell12& m, const Cell12& f, double* r, double* J) {
  const PairConst3 k(m, f);
  const double n2 = pose[0] * pose[0] + pose[1] * pose[1];
  const double n = std::sqrt(n2);
  const double c = pose[0] / n, s = pose[1] / n;
  const double xr = c * k.mm[0] - s * k.mm[1], yr = s * k.mm[0] + c * k.mm[1];
  const double d[3] = {xr + pose[2] - k.fm[0], yr + pose[3] - k.fm[1], k.mm[2] - k.fm[2]};
  const double R[3][3] = {{c, -s, 0}, {s, c, 0}, {0, 0, 1}};
  double RS[3][3], M[3][3], B[3][3], Bi[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { RS[i][j] = 0; for (int l = 0; l < 3; ++l) RS[i][j] += R[i][l] * k.mc[l][j]; }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { M[i][j] = 0; for (int l = 0; l < 3; ++l) M[i][j] += RS[i][l] * R[j][l]; B[i][j] = M[i][j] + k.fc[i][j]; }
  inv3(B, Bi);
  double q[3], p[3];  // q = B^-1 d, p = B^-T d
  for (int i = 0; i < 3; ++i) { q[i] = 0; p[i] = 0; for (int j = 0; j < 3; ++j) { q[i] += Bi[i][j] * d[j]; p[i] += Bi[j][i] * d[j]; } }
  const double dd = d[0] * q[0] + d[1] * q[1] + d[2] * q[2];
  const double rr = std::sqrt(dd);
  *r = rr;
  if (!J) return;
  // d(dd) = (q+p)^T d(delta) - p^T dB q ;  d(delta)/dtheta = (-yr, xr, 0) ; dB/dtheta = S M + M S^T
  double Mq[3], Mtp[3];
  for (int i = 0; i < 3; ++i) { Mq[i] = 0; Mtp[i] = 0; for (int j = 0; j < 3; ++j) { Mq[i] += M[i][j] * q[j]; Mtp[i] += M[j][i] * p[j]; } }
  // S^T v = (v1, -v0, 0)
  const double pSMq = (p[1] * Mq[0] - p[0] * Mq[1]) + (Mtp[0] * q[1] - Mtp[1] * q[0]);
  const double dth = ((q[0] + p[0]) * (-yr) + (q[1] + p[1]) * xr - pSMq) / (2.0 * rr);
  J[0] = dth * (-pose[1] / n2);
  J[1] = dth * (pose[0] / n2);
  J[2] = (q[0] + p[0]) / (2.0 * rr);
  J[3] = (q[1] + p[1]) / (2.0 * rr);
}

// ------------------------------------------------------------------------------------------
// a11  Robust losses     R/src/ndt_registration/ceres_loss_functions.cpp:10-39, ceres_loss_functions.h:9-48
//      + ceres::ScaledLoss, ceres::Corrector (Ceres 2.1.0 loss_function.cc / corrector.cc, restated)
// ------------------------------------------------------------------------------------------
enum LossKind { LOSS_NONE = 0, LOSS_BARRON = 1, LOSS_WELSCH = 2 };
struct Loss {
  int kind = LOSS_NONE;
  double a = 1.0;       // loss_function_scale
  double alpha = -2.0;  // loss_function_convexity (Barron only)
  double mu = 1.0;      // GNC control parameter
  double weight = 1.0;  // ScaledLoss factor; applied to rho, rho', rho''
  void evaluate(double s, double rho[3]) const {
    if (kind == LOSS_NONE) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
    else if (kind == LOSS_WELSCH) {
      const double b = mu * a * a, c = -1.0 / b;
      const double ex = std::exp(s * c);
      rho[0] = b * (1 - ex); rho[1] = ex; rho[2] = c * ex;
    } else {
      const double b = mu * a * a, c = 1 / b, factor = std::abs(alpha - 2.0), exponent = 0.5 * alpha;
      const double pre_factor = b * factor / alpha, times_s = 2 * c / factor;
      if (alpha >= 2.0) { rho[0] = s; rho[1] = 1; rho[2] = 0; }
      else if (std::abs(alpha) <= 0.05) {
        const double sum = 1.0 + s * c, inv = 1.0 / sum;
        rho[0] = b * std::log(sum);
        rho[1] = std::max(std::numeric_limits<double>::min(), inv);
        rho[2] = -c * (inv * inv);
      } else {
        const double to_exp = s * times_s + 1.0;
        rho[0] = pre_factor * (std::pow(to_exp, exponent) - 1.);
        rho[1] = pre_factor * exponent * std::pow(to_exp, exponent - 1.) * times_s;
        rho[2] = pre_factor * exponent * (exponent - 1) * std::pow(to_exp, exponent - 2.) * times_s * times_s;
      }
    }
    rho[0] *= weight; rho[1] *= weight; rho[2] *= weight;
  }
};

struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if ((sq_norm == 0.0) || (rho[2] <= 0.0)) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; return; }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  // single scalar residual with an n-wide row
  void correct(double& r, double* J, int n) const {
    if (J) {
      if (alpha_sq_norm == 0.0) { for (int i = 0; i < n; ++i) J[i] *= sqrt_rho1; }
      else { for (int i = 0; i < n; ++i) { const double rtj = J[i] * r; J[i] = sqrt_rho1 * (J[i] - alpha_sq_norm * r * rtj); } }
    }
    r *= residual_scaling;
  }
};

// One residual block as ceres::ResidualBlock::Evaluate would return it (cost, corrected r, corrected J).
inline void eval_block(int variant, const double* params, const Cell12& m, const Cell12& f, const Loss& loss,
                       bool apply_loss, double* cost, double* r, double* J) {
  eval_pair_autodiff(variant, params, m, f, r, J);
  const double sq = (*r) * (*r);
  if (!apply_loss || (loss.kind == LOSS_NONE && loss.weight == 1.0)) { *cost = 0.5 * sq; return; }
  double rho[3];
  loss.evaluate(sq, rho);
  *cost = 0.5 * rho[0];
  Corrector corr(sq, rho);
  corr.correct(*r, J, variant_num_params(variant));
}

// ------------------------------------------------------------------------------------------
// a12  GNC schedule       R/src/ndt_registration/ndt_matcher.cpp:386-397 / 472-483
// ------------------------------------------------------------------------------------------
inline double gnc_initial_mu(double max_residual, double loss_scale, double divisor, int steps) {
  double mu = 2.0 * std::pow(max_residual, 2) / std::pow(loss_scale, 2);
  return std::min(mu, std::pow(divisor, steps - 1));
}

// ------------------------------------------------------------------------------------------
// Map::calculateCSDivergence     R/src/ndt_representation/ndt_map.cpp:42-99   
// float32 3x3 algebra in Eigen's evaluation order, fp64 accumulation in the reference's loop order.  The reference leaves
// interaction_term / fixed_term / moving_term uninitialised (quirk B.15); the oracle starts them at 0.
// `moving` must already be expressed in the fixed map's frame (R/src/local_fuser/local_fuser.cpp:338).
// ------------------------------------------------------------------------------------------
inline float det3_f(const float m[9]) {   // Eigen determinant_impl<Matrix3f>: helper(0,1,2) - helper(1,0,2) + helper(2,0,1)
  auto M = [&](int r, int c) { return m[r * 3 + c]; };
  const float h0 = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1));
  const float h1 = M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0));
  const float h2 = M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
  return (h0 - h1) + h2;
}
inline void inv3_f(const float m[9], float inv[9]) {   // Eigen compute_inverse_size3
  auto M = [&](int r, int c) { return m[r * 3 + c]; };
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
  };
  const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const float det = c0 * M(0, 0) + (c1 * M(1, 0) + c2 * M(2, 0));
  const float invdet = 1.0f / det;
  inv[0] = c0 * invdet; inv[1] = c1 * invdet; inv[2] = c2 * invdet;
  inv[3] = cof(0, 1) * invdet; inv[4] = cof(1, 1) * invdet; inv[5] = cof(2, 1) * invdet;
  inv[6] = cof(0, 2) * invdet; inv[7] = cof(1, 2) * invdet; inv[8] = cof(2, 2) * invdet;
}
inline double cs_gauss_overlap(const Cell12& a, const Cell12& b) {
  float d[3], S[9], Si[9];
  for (int i = 0; i < 3; ++i) d[i] = a.mu[i] - b.mu[i];
  for (int i = 0; i < 9; ++i) S[i] = a.cov[i] + b.cov[i];
  inv3_f(S, Si);
  float t[3];
  for (int j = 0; j < 3; ++j) t[j] = d[0] * Si[j] + (d[1] * Si[3 + j] + d[2] * Si[6 + j]);
  const double e = (double)(t[0] * d[0] + (t[1] * d[1] + t[2] * d[2]));
  return (0.5 / std::sqrt(M_PI * M_PI * (double)det3_f(S))) * std::exp(-0.5 * e);
}
inline double cs_divergence(const Cell12* fixed, size_t nf, const Cell12* moving, size_t nm, double terms[3] = nullptr) {
  double interaction_term = 0.0, fixed_term = 0.0, moving_term = 0.0;
  for (size_t f = 0; f < nf; ++f) {
    if (det3_f(fixed[f].cov) < 0.00001) continue;
    for (size_t q = 0; q < nm; ++q) interaction_term += cs_gauss_overlap(fixed[f], moving[q]);
    float inv[9]; inv3_f(fixed[f].cov, inv);
    fixed_term += std::sqrt((double)det3_f(inv)) / (2 * M_PI);
    for (size_t q = 0; q < f; ++q) fixed_term += 2 * cs_gauss_overlap(fixed[f], fixed[q]);
  }
  for (size_t f = 0; f < nm; ++f) {
    if (det3_f(moving[f].cov) < 0.00001) continue;
    float inv[9]; inv3_f(moving[f].cov, inv);
    moving_term += std::sqrt((double)det3_f(inv)) / (2 * M_PI);
    for (size_t q = 0; q < f; ++q) moving_term += 2 * cs_gauss_overlap(moving[f], moving[q]);
  }
  if (terms) { terms[0] = interaction_term; terms[1] = fixed_term; terms[2] = moving_term; }
  return -std::log(interaction_term) + 0.5 * std::log(fixed_term) + 0.5 * std::log(moving_term);
}

// ------------------------------------------------------------------------------------------
// RadarPreprocessor::filterScan   R/src/radar_preprocessing/radar_preprocessor.cpp:45-125   
// Restated sequentially as written.  The two `while(true)` walks leave closer_idx / further_idx uninitialised when their
// size_t wrap-around guard fires first (UB in the reference); the oracle defines them as the position reached.
// tf: row-major 3x4 of initial_transform_radar_baselink (Affine3f); pcl::transformPointCloud's SSE2 path (PCL 1.10
// detail::Transformer<float>::se3: _mm_add_ps(p0, _mm_add_ps(p1, _mm_add_ps(p2, c[3])))) evaluates each row as x c0 + (y c1 + (z c2 + c3)).
// ------------------------------------------------------------------------------------------
struct FilterParams { float min_distance, max_distance, min_intensity; double beam_thr; float tf[12]; };
inline void filter_scan(const Pt4* raw, size_t n, const FilterParams& fp, std::vector<Pt4>& out, std::vector<size_t>* peaks = nullptr) {
  out.clear();
  float current_angle = 1000;
  float max_intensity = 0;
  size_t current_max_idx = 0;
  std::vector<size_t> max_idzs;
  for (size_t i = 0; i < n; ++i) {
    const float dist = std::hypot(raw[i].x, raw[i].y);
    const float angle = std::atan2(raw[i].y, raw[i].x);
    const float intensity = raw[i].i;
    if (std::abs(angle - current_angle) > 0.0001) {
      if (current_angle < 3 * M_PI) {
        if (max_idzs.empty() || max_idzs.back() != current_max_idx) max_idzs.push_back(current_max_idx);
        max_intensity = 0;
      }
      current_angle = angle;
    }
    if (dist > fp.min_distance && dist < fp.max_distance && intensity > max_intensity) { max_intensity = intensity; current_max_idx = i; }
  }
  if (peaks) *peaks = max_idzs;
  auto hyp = [&](size_t k) { return std::hypot(raw[k].x, raw[k].y); };
  for (size_t m = 0; m < max_idzs.size(); ++m) {
    const size_t s = max_idzs[m];
    size_t closer = s, further = s, k = 0;
    while (true) {
      if (s - k - 1 > n - 1) { closer = s - k; break; }                       // size_t wrap: reached index 0
      if (((hyp(s - k) - hyp(s - k - 1)) > fp.beam_thr) || (raw[s - k].i <= raw[s - k - 1].i) || (hyp(s - k) < fp.min_distance)) { closer = s - k; break; }
      ++k;
    }
    k = 0;
    while (true) {
      if (s + k + 1 > n - 1) { further = s + k; break; }
      if (((hyp(s + k) - hyp(s + k + 1)) > fp.beam_thr) || (raw[s + k].i <= raw[s + k + 1].i) || (hyp(s + k) < fp.min_distance)) { further = s + k; break; }
      ++k;
    }
    for (size_t j = closer; j <= further; ++j) {
      const float dist = std::hypot(raw[j].x, raw[j].y);
      if (dist > fp.min_distance && dist < fp.max_distance && raw[j].i > fp.min_intensity) {
        Pt4 o;
        o.x = raw[j].x * fp.tf[0] + (raw[j].y * fp.tf[1] + (raw[j].z * fp.tf[2] + fp.tf[3]));
        o.y = raw[j].x * fp.tf[4] + (raw[j].y * fp.tf[5] + (raw[j].z * fp.tf[6] + fp.tf[7]));
        o.z = raw[j].x * fp.tf[8] + (raw[j].y * fp.tf[9] + (raw[j].z * fp.tf[10] + fp.tf[11]));
        o.i = raw[j].i;
        out.push_back(o);
      }
    }
  }
}

}  // namespace orc
