// Eigen's Transform::rotation() of the float affine Cell::transformCell builds, and the AffineRec derived from a pose: shared by
// k1_voxelize.cu (prepare_affine_kernel) and k2_associate.cu (the single-map fused association).  Only meaningful in translation units
// compiled with -fmad=false: every product / sum is a separate IEEE operation in Eigen's order.
#pragma once
#include <float.h>

#include "common.cuh"

namespace randt {
namespace {

// ---- Eigen 3.3.7 Transform<float,3,Affine>::rotation() of the lift Cell::transformCell builds (ndt_cell.cpp:118-122) ----
// rotation() = computeRotationScaling (Geometry/Transform.h): JacobiSVD<Matrix3f> of L = [[c,-s,0],[s,c,0],[0,0,1]] (two-sided Jacobi sweeps,
// SVD/JacobiSVD.h + Jacobi.h), x = det(U V^T), U.col(0) /= x, R = U V^T.  R equals L only up to float rounding, and the covariances the
// reference transforms see R, not L.  One thread per map/pose runs this once; the cell kernels read the result.
struct JRot { float c, s; };
__device__ __forceinline__ void jrot_rows(float m[3][3], int p, int q, JRot j, int ncols) {
  if (j.c == 1.0f && j.s == 0.0f) return;
  for (int k = 0; k < ncols; ++k) { const float xi = m[p][k], yi = m[q][k]; m[p][k] = j.c * xi + j.s * yi; m[q][k] = -j.s * xi + j.c * yi; }
}
__device__ __forceinline__ void jrot_cols(float m[3][3], int p, int q, JRot j) {
  const float c = j.c, s = -j.s;
  if (c == 1.0f && s == 0.0f) return;
  for (int k = 0; k < 3; ++k) { const float xi = m[k][p], yi = m[k][q]; m[k][p] = c * xi + s * yi; m[k][q] = -s * xi + c * yi; }
}
__device__ __forceinline__ void mul3_eigen(const float A[3][3], const float B[3][3], float C[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + (A[i][1] * B[1][j] + A[i][2] * B[2][j]);
}
__device__ void affine_rotation_f(float c, float s, float R[3][3]) {
  const float precision = 2.0f * FLT_EPSILON, consider_as_zero = FLT_MIN;
  float w[3][3] = {{c, -s, 0.f}, {s, c, 0.f}, {0.f, 0.f, 1.f}};
  float U[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}}, V[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
  float scale = 0.0f;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = fabsf(w[i][j]) > scale ? fabsf(w[i][j]) : scale;
  if (scale == 0.0f) scale = 1.0f;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w[i][j] = w[i][j] / scale;
  float max_diag = 0.0f;
  for (int i = 0; i < 3; ++i) max_diag = fabsf(w[i][i]) > max_diag ? fabsf(w[i][i]) : max_diag;
  bool finished = false;
  for (int sweep = 0; !finished && sweep < 64; ++sweep) {      // Eigen has no sweep bound; 64 is never reached for a near-rotation
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        const float pm = precision * max_diag;
        const float threshold = consider_as_zero < pm ? pm : consider_as_zero;
        if (fabsf(w[p][q]) > threshold || fabsf(w[q][p]) > threshold) {
          finished = false;
          float m[3][3] = {{w[p][p], w[p][q], 0.f}, {w[q][p], w[q][q], 0.f}, {0.f, 0.f, 0.f}};
          JRot rot1;
          const float t = m[0][0] + m[1][1];
          const float d = m[1][0] - m[0][1];
          if (fabsf(d) < FLT_MIN) { rot1.s = 0.0f; rot1.c = 1.0f; }
          else { const float u = t / d; const float tmp = sqrtf(1.0f + u * u); rot1.s = 1.0f / tmp; rot1.c = u / tmp; }
          jrot_rows(m, 0, 1, rot1, 2);
          JRot jr;     // makeJacobi(m(0,0), m(0,1), m(1,1))
          {
            const float x = m[0][0], y = m[0][1], z = m[1][1];
            const float deno = 2.0f * fabsf(y);
            if (deno < FLT_MIN) { jr.c = 1.0f; jr.s = 0.0f; }
            else {
              const float tau = (x - z) / deno;
              const float ww = sqrtf(tau * tau + 1.0f);
              float tt;
              if (tau > 0.0f) tt = 1.0f / (tau + ww); else tt = 1.0f / (tau - ww);
              const float sign_t = tt > 0.0f ? 1.0f : -1.0f;
              const float n = 1.0f / sqrtf(tt * tt + 1.0f);
              jr.s = -sign_t * (y / fabsf(y)) * fabsf(tt) * n;
              jr.c = n;
            }
          }
          const JRot jrt = {jr.c, -jr.s};
          const JRot jl = {rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
          jrot_rows(w, p, q, jl, 3);
          const JRot jlt = {jl.c, -jl.s};
          jrot_cols(U, p, q, jlt);
          jrot_cols(w, p, q, jr);
          jrot_cols(V, p, q, jr);
          const float a1 = fabsf(w[p][p]), a2 = fabsf(w[q][q]);
          const float mx = a1 < a2 ? a2 : a1;
          max_diag = max_diag < mx ? mx : max_diag;
        }
      }
  }
  float sv[3];
  for (int i = 0; i < 3; ++i) {
    sv[i] = fabsf(w[i][i]);
    if (w[i][i] < 0.0f) for (int k = 0; k < 3; ++k) U[k][i] = -U[k][i];
  }
  for (int i = 0; i < 3; ++i) sv[i] = sv[i] * scale;
  for (int i = 0; i < 3; ++i) {
    int pos = 0; float best = sv[i];
    for (int k = 1; k < 3 - i; ++k) if (sv[i + k] > best) { best = sv[i + k]; pos = k; }
    if (best == 0.0f) break;
    if (pos) {
      pos += i;
      float t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
      for (int k = 0; k < 3; ++k) { t = U[k][pos]; U[k][pos] = U[k][i]; U[k][i] = t; t = V[k][pos]; V[k][pos] = V[k][i]; V[k][i] = t; }
    }
  }
  float Vt[3][3], UVt[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[i][j] = V[j][i];
  mul3_eigen(U, Vt, UVt);
  const float x = (UVt[0][0] * (UVt[1][1] * UVt[2][2] - UVt[1][2] * UVt[2][1]) - UVt[0][1] * (UVt[1][0] * UVt[2][2] - UVt[1][2] * UVt[2][0])) +
                  UVt[0][2] * (UVt[1][0] * UVt[2][1] - UVt[1][1] * UVt[2][0]);
  for (int k = 0; k < 3; ++k) U[k][0] = U[k][0] / x;
  mul3_eigen(U, Vt, R);
}


// AffineRec (4 x float4, see common.cuh) of the float affine (c, s, tx, ty)
__device__ __forceinline__ void make_affine_rec(float c, float s, float tx, float ty, float4* __restrict__ out) {
  float R[3][3];
  affine_rotation_f(c, s, R);
  out[0] = make_float4(c, s, tx, ty);
  out[1] = make_float4(R[0][0], R[0][1], R[0][2], R[1][0]);
  out[2] = make_float4(R[1][1], R[1][2], R[2][0], R[2][1]);
  out[3] = make_float4(R[2][2], 0.f, 0.f, 0.f);
}
// from float64 Sophus SE2d storage: `Eigen::Affine2f(pose.cast<float>().matrix())` (ndt_matcher.cpp:208, local_fuser.cpp:338): Sophus' cast
// re-normalises the float complex (length = hypot(re, im), glibc hypotf = double sqrt rounded once)
__device__ __forceinline__ void make_affine_rec_se2d(const double* __restrict__ p, float4* __restrict__ out) {
  const float re = (float)p[0], im = (float)p[1];
  const float length = (float)sqrt((double)re * (double)re + (double)im * (double)im);
  make_affine_rec(re / length, im / length, (float)p[2], (float)p[3], out);
}

}  // namespace
}  // namespace randt
