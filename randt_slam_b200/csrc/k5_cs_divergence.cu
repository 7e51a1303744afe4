// K5 — Cauchy-Schwarz divergence between two NDT maps (loop-closure verification), sm_100a.  Compiled with -fmad=false: the
// float32 3x3 algebra reproduces the reference's Eigen evaluation order; sums are fp64.
//
// Replaces Map::calculateCSDivergence   R/src/ndt_representation/ndt_map.cpp:42-99, called right after every
// Matcher::estimateLoopConstraint (R/src/local_fuser/local_fuser.cpp:338-340, 396-402) on the submap and the moving map
// transformed by the refined estimate.  With G(a, b) = 0.5 / sqrt(pi^2 det(S_a + S_b)) * exp(-0.5 (mu_a - mu_b)^T (S_a + S_b)^-1 (mu_a - mu_b)):
//   interaction = sum over fixed f with det(S_f) >= 1e-5, over all moving q, of G(f, q)
//   fixed_term  = sum over fixed f with det(S_f) >= 1e-5 of [ sqrt(det(S_f^-1)) / (2 pi) + 2 sum_{q < f} G(f, q) ]   (q over ALL earlier cells)
//   moving_term = the same over the moving map
//   cs = -log(interaction) + 0.5 log(fixed_term) + 0.5 log(moving_term)
// The reference never initialises its three accumulators; they are zero here.
//
// Work: all-pairs, O(N_f N_m + N_f^2/2 + N_m^2/2) 3x3 inverses per map pair.  grid = (kSplit, B): the rows of map pair b are dealt
// round-robin to the warps of its kSplit CTAs, lanes stride a row's columns, per-lane fp64 partial sums are folded in a fixed
// order (warp shuffle tree -> shared memory -> one partial record per CTA); the last CTA of a map pair to finish (ticket) adds
// the kSplit partials in index order and writes the divergence: deterministic, no atomics on the sums.
#include <math.h>

#include "common.cuh"

namespace randt {
namespace {

constexpr int kCsThreads = 256;
constexpr int kCsWarps = kCsThreads / 32;
constexpr int kCsSplit = 8;

struct CellF { float mu[3]; float cov[9]; };
__device__ __forceinline__ CellF load_cell(const float4* __restrict__ tab, uint32_t idx) {
  const float4 a = __ldg(tab + 3 * (size_t)idx), b = __ldg(tab + 3 * (size_t)idx + 1), c = __ldg(tab + 3 * (size_t)idx + 2);
  CellF o;
  o.mu[0] = a.x; o.mu[1] = a.y; o.mu[2] = a.z;
  o.cov[0] = a.w; o.cov[1] = b.x; o.cov[2] = b.y; o.cov[3] = b.z; o.cov[4] = b.w; o.cov[5] = c.x; o.cov[6] = c.y; o.cov[7] = c.z; o.cov[8] = c.w;
  return o;
}
#define M_(m, r, c) (m)[(r) * 3 + (c)]
// Eigen determinant_impl<Matrix3f>: bruteforce_det3_helper(0,1,2) - helper(1,0,2) + helper(2,0,1)
__device__ __forceinline__ float det3(const float* m) {
  const float h0 = M_(m, 0, 0) * (M_(m, 1, 1) * M_(m, 2, 2) - M_(m, 1, 2) * M_(m, 2, 1));
  const float h1 = M_(m, 0, 1) * (M_(m, 1, 0) * M_(m, 2, 2) - M_(m, 1, 2) * M_(m, 2, 0));
  const float h2 = M_(m, 0, 2) * (M_(m, 1, 0) * M_(m, 2, 1) - M_(m, 1, 1) * M_(m, 2, 0));
  return (h0 - h1) + h2;
}
// Eigen compute_inverse_size3 (cofactors, determinant from column 0 with the unrolled-redux association p0 + (p1 + p2))
__device__ __forceinline__ void inv3(const float* m, float* inv) {
#define COF_(i, j) (M_(m, ((i) + 1) % 3, ((j) + 1) % 3) * M_(m, ((i) + 2) % 3, ((j) + 2) % 3) - M_(m, ((i) + 1) % 3, ((j) + 2) % 3) * M_(m, ((i) + 2) % 3, ((j) + 1) % 3))
  const float c0 = COF_(0, 0), c1 = COF_(1, 0), c2 = COF_(2, 0);
  const float det = c0 * M_(m, 0, 0) + (c1 * M_(m, 1, 0) + c2 * M_(m, 2, 0));
  const float invdet = 1.0f / det;
  inv[0] = c0 * invdet; inv[1] = c1 * invdet; inv[2] = c2 * invdet;
  inv[3] = COF_(0, 1) * invdet; inv[4] = COF_(1, 1) * invdet; inv[5] = COF_(2, 1) * invdet;
  inv[6] = COF_(0, 2) * invdet; inv[7] = COF_(1, 2) * invdet; inv[8] = COF_(2, 2) * invdet;
#undef COF_
}
#undef M_
// G(a, b) of the header comment
__device__ __forceinline__ double gauss_overlap(const CellF& a, const CellF& b) {
  float d[3], S[9], Si[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = a.mu[i] - b.mu[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) S[i] = a.cov[i] + b.cov[i];
  inv3(S, Si);
  float t[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = d[0] * Si[j] + (d[1] * Si[3 + j] + d[2] * Si[6 + j]);
  const double e = (double)(t[0] * d[0] + (t[1] * d[1] + t[2] * d[2]));
  const double pi = 3.14159265358979323846;
  return (0.5 / sqrt(pi * pi * (double)det3(S))) * exp(-0.5 * e);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kCsThreads) k5_cs_divergence_kernel(const float4* __restrict__ cells_f, const uint32_t* __restrict__ off_f,
                                                                     const float4* __restrict__ cells_m, const uint32_t* __restrict__ off_m,
                                                                     double* __restrict__ partials /*[B][kCsSplit][3]*/, uint32_t* __restrict__ tickets,
                                                                     double* __restrict__ out) {
  const uint32_t b = blockIdx.y, split = blockIdx.x;
  const uint32_t f0 = off_f[b], nf = off_f[b + 1] - f0, m0 = off_m[b], nm = off_m[b + 1] - m0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = split * kCsWarps + warp, n_gw = kCsSplit * kCsWarps;
  double s_int = 0.0, s_fix = 0.0, s_mov = 0.0;
  const double two_pi = 2.0 * 3.14159265358979323846;
  // rows of the fixed map: interaction with every moving cell + pairs with earlier fixed cells
  for (uint32_t i = gw; i < nf; i += n_gw) {
    const CellF f = load_cell(cells_f, f0 + i);
    if (det3(f.cov) < 0.00001f) continue;                      // `continue` skips the whole row (ndt_map.cpp:56-58)
    for (uint32_t j = lane; j < nm; j += 32) s_int += gauss_overlap(f, load_cell(cells_m, m0 + j));
    if (lane == 0) { float inv[9]; inv3(f.cov, inv); s_fix += sqrt((double)det3(inv)) / two_pi; }
    for (uint32_t j = lane; j < i; j += 32) s_fix += 2.0 * gauss_overlap(f, load_cell(cells_f, f0 + j));
  }
  // rows of the moving map
  for (uint32_t i = gw; i < nm; i += n_gw) {
    const CellF f = load_cell(cells_m, m0 + i);
    if (det3(f.cov) < 0.00001f) continue;
    if (lane == 0) { float inv[9]; inv3(f.cov, inv); s_mov += sqrt((double)det3(inv)) / two_pi; }
    for (uint32_t j = lane; j < i; j += 32) s_mov += 2.0 * gauss_overlap(f, load_cell(cells_m, m0 + j));
  }
  __shared__ double sm[kCsWarps][3];
  __shared__ uint32_t s_ticket;
  s_int = warp_sum(s_int); s_fix = warp_sum(s_fix); s_mov = warp_sum(s_mov);
  if (lane == 0) { sm[warp][0] = s_int; sm[warp][1] = s_fix; sm[warp][2] = s_mov; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < kCsWarps; ++w) v += sm[w][threadIdx.x];
    partials[((size_t)b * kCsSplit + split) * 3 + threadIdx.x] = v;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[b], 1u);
  __syncthreads();
  if (s_ticket == kCsSplit - 1 && threadIdx.x == 0) {            // last CTA of this map pair: fold the partials in index order
    __threadfence();
    double t[3] = {0.0, 0.0, 0.0};
    for (int sp = 0; sp < kCsSplit; ++sp)
      for (int k = 0; k < 3; ++k) t[k] += __ldcg(partials + ((size_t)b * kCsSplit + sp) * 3 + k);
    out[b] = -log(t[0]) + 0.5 * log(t[1]) + 0.5 * log(t[2]);
    tickets[b] = 0u;                                             // re-arm
  }
}

}  // namespace

cudaError_t launch_cs_divergence(const float4* cells_f, const uint32_t* off_f, const float4* cells_m, const uint32_t* off_m, uint32_t n_maps,
                                 double* d_partials, uint32_t* d_tickets, double* d_out, cudaStream_t s, int* n_launches) {
  if (n_maps == 0) return cudaSuccess;
  k5_cs_divergence_kernel<<<dim3(kCsSplit, n_maps), kCsThreads, 0, s>>>(cells_f, off_f, cells_m, off_m, d_partials, d_tickets, d_out);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}
int cs_divergence_split() { return kCsSplit; }

}  // namespace randt
