// K6 — per-azimuth peak filter of a raw polar radar scan (sm_100a).  Compiled with -fmad=false.
//
// Replaces RadarPreprocessor::filterScan   R/src/radar_preprocessing/radar_preprocessor.cpp:45-125  (the only
// HBM-scale input of the path — 400 azimuths x ~3000 range bins x 16 B = 19 MB per Oxford scan against 80 KB of filtered points):
//   pass 1  for every azimuth, the first strongest return with min_range < range < max_range (strict `>` against a running
//           maximum that starts at 0); the reference emits an azimuth's peak when the NEXT azimuth starts, so the last azimuth of
//           a scan never contributes, and an azimuth without a valid return contributes nothing (except that a scan whose very
//           first azimuth is empty emits point 0);
//   pass 2  from each peak, walk inwards and outwards while the intensity keeps falling, the range step stays below
//           beam_distance_increment_threshold and the range stays above min_range; keep the points of that run that pass the
//           range / min_intensity gates; transform them to the base frame (pcl::transformPointCloud with an Affine3f).
// The reference finds azimuth boundaries by comparing atan2(y, x) of every point with the first point of the current azimuth
// (tolerance 1e-4 rad).  This kernel takes the organised shape (n_azimuths x n_bins, what cloud_in->height/width carry) and
// VERIFIES that the angle rule would cut the scan at exactly those rows; if not it reports RANDT_E_INVALID instead of guessing.
//
// k6_row_peak_kernel: one CTA per azimuth, coalesced float4 loads, block arg-max with lowest-index tie-break (= first-wins).
// k6_runs_kernel: one thread per emitted peak walks its run (tens of dependent loads), block scan of the kept counts, ordered
// write of the transformed points.  HBM-bound: 16 B per raw point read once, algorithmic bytes 16 n_az n_bins + 16 n_out.
#include <math.h>

#include "common.cuh"

namespace randt {
namespace {

constexpr int kRowThreads = 256;
constexpr uint32_t kNoPeak = 0xffffffffu;

// std::hypot(float, float): glibc evaluates it in double and rounds once
__device__ __forceinline__ float hypot_ref(float x, float y) { return (float)sqrt((double)x * (double)x + (double)y * (double)y); }

// Per point the reference evaluates hypot (range gate) and atan2 (azimuth cut).  Both are only DECISIONS against thresholds, so
// the kernel takes them with cheap float32 bounds and falls back to the exact evaluation in the narrow bands where the bound
// cannot decide (relative width 1e-5 around the range limits; azimuth deviations between 0.45e-4 and the 1e-4 cut):
//   range:  d2 = x^2 + y^2 in float carries a relative error < 3e-7, and fl32(sqrt) adds < 6e-8, so d2 outside (1 -+ 1e-5) lim^2 decides;
//   angle:  |atan2(p) - atan2(p0)| <= 1e-4 is certain when |p0 x p| <= 0.45e-4 (p0 . p) with p0 . p > 0 (the true angle is then < 0.45e-4 and
//           the two float roundings add < 1e-6) — unless the row lies on the +-pi branch cut of atan2, which rows flag up front.
__global__ void __launch_bounds__(kRowThreads) k6_row_peak_kernel(const float4* __restrict__ raw, uint32_t n_az, uint32_t n_bins, float min_d,
                                                                  float max_d, uint32_t* __restrict__ peak_idx, float* __restrict__ row_angle,
                                                                  int* __restrict__ status) {
  const uint32_t row = blockIdx.x, scan = blockIdx.y;          // a batch of scans: one grid row per scan, per-scan outputs
  raw += (size_t)scan * n_az * n_bins; peak_idx += (size_t)scan * n_az; row_angle += (size_t)scan * n_az; status += scan;
  const float4* p = raw + (size_t)row * n_bins;
  const float4 first = __ldg(p);
  const float x0 = first.x, y0 = first.y;
  const float a0 = atan2f(y0, x0);
  const bool near_cut = !(x0 > 0.0f || fabsf(y0) > 1e-3f * fabsf(x0));    // within ~1e-3 rad of +-pi (or at the origin): always exact
  const float lo2 = min_d * min_d, hi2 = max_d * max_d;
  const float lo2_in = lo2 * 1.00001f, lo2_out = lo2 * 0.99999f, hi2_in = hi2 * 0.99999f, hi2_out = hi2 * 1.00001f;
  float best = 0.0f; uint32_t best_i = kNoPeak; bool bad = false;
  constexpr int U = 4;                                                      // independent 16-byte loads in flight per thread (12 was slower: 12.8 vs 9.9 us)
  for (uint32_t b0 = threadIdx.x; b0 < n_bins; b0 += kRowThreads * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const uint32_t b = b0 + u * kRowThreads; v[u] = b < n_bins ? __ldg(p + b) : first; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t b = b0 + u * kRowThreads;
      if (b >= n_bins) continue;
      const float x = v[u].x, y = v[u].y;
      // azimuth rule
      const float cr = x0 * y - y0 * x, dt = x0 * x + y0 * y;
      if (near_cut || !(fabsf(cr) <= 0.45e-4f * dt)) { if (fabsf(atan2f(y, x) - a0) > 0.0001f) bad = true; }
      // range gate
      const float d2 = x * x + y * y;
      bool in_range;
      if (d2 > lo2_in && d2 < hi2_in) in_range = true;
      else if (d2 < lo2_out || d2 > hi2_out) in_range = false;
      else { const float dist = hypot_ref(x, y); in_range = dist > min_d && dist < max_d; }
      if (in_range && v[u].w > best) { best = v[u].w; best_i = b; }       // ascending b per thread: first-wins inside the thread
    }
  }
  __shared__ float s_val[kRowThreads];
  __shared__ uint32_t s_idx[kRowThreads];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_i;
  __syncthreads();
  if (bad) s_bad = 1;
  for (int o = kRowThreads / 2; o > 0; o >>= 1) {
    __syncthreads();
    if ((int)threadIdx.x < o) {
      const float v2 = s_val[threadIdx.x + o]; const uint32_t i2 = s_idx[threadIdx.x + o];
      if (v2 > s_val[threadIdx.x] || (v2 == s_val[threadIdx.x] && i2 < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v2; s_idx[threadIdx.x] = i2; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    peak_idx[row] = s_idx[0] == kNoPeak ? kNoPeak : row * n_bins + s_idx[0];
    row_angle[row] = a0;
    if (s_bad) atomicExch(status, 1);
  }
}

struct FilterDev { float min_d, max_d, min_i; double beam_thr; float tf[12]; };

// MODE 0: one scan, count and write in one pass (n_out[0] = kept points).  A batch (one CTA per scan) takes two passes around an
// exclusive scan of the per-scan counts: MODE 1 counts into n_out[scan], MODE 2 writes scan `scan` at out[scan_off[scan] ...].
template <int MODE>
__global__ void __launch_bounds__(1024) k6_runs_kernel(const float4* __restrict__ raw, uint32_t n_az, uint32_t n_bins, FilterDev f,
                                                       const uint32_t* __restrict__ peak_idx, const float* __restrict__ row_angle,
                                                       float4* __restrict__ out, uint32_t cap, uint32_t* __restrict__ n_out,
                                                       const uint32_t* __restrict__ scan_off, int* __restrict__ status) {
  const uint32_t scan = blockIdx.x;
  raw += (size_t)scan * n_az * n_bins; peak_idx += (size_t)scan * n_az; row_angle += (size_t)scan * n_az; status += scan; n_out += scan;
  const uint32_t w_base = MODE == 2 ? scan_off[scan] : 0u;
  // the reference's cut rule between consecutive rows: |angle(first of row r+1) - angle(first of row r)| must exceed 1e-4
  __shared__ uint32_t s_scan[1024];
  __shared__ uint32_t s_carry;
  const uint32_t n_pts = n_az * n_bins;
  for (uint32_t r = threadIdx.x; r + 1 < n_az; r += blockDim.x)
    if (!(fabsf(row_angle[r + 1] - row_angle[r]) > 0.0001f)) atomicExch(status, 1);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  // rows 0 .. n_az-2 emit (the last azimuth is never closed); row 0 without a valid return emits point 0
  const uint32_t n_rows = n_az > 0 ? n_az - 1 : 0;
  for (uint32_t base = 0; base < n_rows; base += blockDim.x) {
    const uint32_t r = base + threadIdx.x;
    uint32_t peak = kNoPeak, lo = 0, hi = 0, cnt = 0;
    if (r < n_rows) {
      peak = peak_idx[r];
      if (peak == kNoPeak && r == 0) peak = 0;
    }
    if (peak != kNoPeak) {
      uint32_t k = 0;
      while (true) {        // towards the sensor
        const uint32_t cur = peak - k;
        if (cur == 0) { lo = cur; break; }              // size_t wrap check of the reference (closer_idx is then left as is: defined here)
        const float4 a = __ldg(raw + cur), b = __ldg(raw + cur - 1);
        const float da = hypot_ref(a.x, a.y), db = hypot_ref(b.x, b.y);
        if (((double)(da - db) > f.beam_thr) || (a.w <= b.w) || (da < f.min_d)) { lo = cur; break; }
        ++k;
      }
      k = 0;
      while (true) {        // away from the sensor
        const uint32_t cur = peak + k;
        if (cur + 1 > n_pts - 1) { hi = cur; break; }
        const float4 a = __ldg(raw + cur), b = __ldg(raw + cur + 1);
        const float da = hypot_ref(a.x, a.y), db = hypot_ref(b.x, b.y);
        if (((double)(da - db) > f.beam_thr) || (a.w <= b.w) || (da < f.min_d)) { hi = cur; break; }
        ++k;
      }
      for (uint32_t j = lo; j <= hi; ++j) {
        const float4 v = __ldg(raw + j);
        const float dist = hypot_ref(v.x, v.y);
        if (dist > f.min_d && dist < f.max_d && v.w > f.min_i) ++cnt;
      }
    }
    // block exclusive scan of cnt (Hillis-Steele in shared memory)
    s_scan[threadIdx.x] = cnt;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {
      uint32_t v = 0;
      if (threadIdx.x >= o) v = s_scan[threadIdx.x - o];
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    const uint32_t incl = s_scan[threadIdx.x], total = s_scan[blockDim.x - 1];
    uint32_t w = w_base + s_carry + incl - cnt;
    if (MODE != 1 && peak != kNoPeak) {
      for (uint32_t j = lo; j <= hi; ++j) {
        const float4 v = __ldg(raw + j);
        const float dist = hypot_ref(v.x, v.y);
        if (dist > f.min_d && dist < f.max_d && v.w > f.min_i) {
          if (w < cap) {
            float4 o;
            // pcl::detail::Transformer<float>::se3 (PCL 1.10, SSE2): p0 + (p1 + (p2 + c3)), pk = src[k] * column k
            o.x = v.x * f.tf[0] + (v.y * f.tf[1] + (v.z * f.tf[2] + f.tf[3]));
            o.y = v.x * f.tf[4] + (v.y * f.tf[5] + (v.z * f.tf[6] + f.tf[7]));
            o.z = v.x * f.tf[8] + (v.y * f.tf[9] + (v.z * f.tf[10] + f.tf[11]));
            o.w = v.w;
            out[w] = o;
          }
          ++w;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && MODE != 2) { *n_out = s_carry; if (MODE == 0 && s_carry > cap) atomicExch(status, 2); }
}

}  // namespace

// pcl::PointXYZI as PCL lays it out (32 bytes: x, y, z, 1.0f | intensity, 3 pad floats; pcl/impl/point_types.hpp) -> the path's float4
// (x, y, 0, intensity): one thread per point, two 16-byte loads, one 16-byte store.  Lets the caller hand cloud.points.data() over as it is.
__global__ void __launch_bounds__(256) pcl_xyzi_to_float4_kernel(const float4* __restrict__ in, uint32_t n, float4* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = __ldg(in + 2 * (size_t)i), b = __ldg(in + 2 * (size_t)i + 1);
  out[i] = make_float4(a.x, a.y, 0.0f, b.x);
}
cudaError_t launch_pcl_xyzi_to_float4(const void* d_in32, uint32_t n, float4* d_out, cudaStream_t s, int* n_launches) {
  if (n == 0) return cudaSuccess;
  pcl_xyzi_to_float4_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(static_cast<const float4*>(d_in32), n, d_out);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_filter_scan(const float4* d_raw, uint32_t n_az, uint32_t n_bins, const randt_filter_params& fp, uint32_t* d_peak, float* d_angle,
                               float4* d_out, uint32_t cap, uint32_t* d_n_out, int* d_status, cudaStream_t s, int* n_launches) {
  if (n_az == 0 || n_bins == 0) return cudaMemsetAsync(d_n_out, 0, sizeof(uint32_t), s);
  FilterDev f; f.min_d = fp.min_range; f.max_d = fp.max_range; f.min_i = fp.min_intensity; f.beam_thr = fp.beam_distance_increment_threshold;
  for (int i = 0; i < 12; ++i) f.tf[i] = fp.sensor_to_base[i];
  k6_row_peak_kernel<<<n_az, kRowThreads, 0, s>>>(d_raw, n_az, n_bins, f.min_d, f.max_d, d_peak, d_angle, d_status);
  k6_runs_kernel<0><<<1, 1024, 0, s>>>(d_raw, n_az, n_bins, f, d_peak, d_angle, d_out, cap, d_n_out, nullptr, d_status);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

// n_scans scans of the same shape, back to back in d_raw.  d_peak / d_angle: [n_scans * n_az]; d_counts: [n_scans]; d_scan_off: [n_scans + 1]
// (exclusive scan of the counts, the offsets randt_voxelize takes); d_block_sums: >= n_scans / 1024 + 2; d_status: [n_scans].
cudaError_t launch_filter_scans(const float4* d_raw, uint32_t n_scans, uint32_t n_az, uint32_t n_bins, const randt_filter_params& fp, uint32_t* d_peak,
                                float* d_angle, float4* d_out, uint32_t cap, uint32_t* d_counts, uint32_t* d_scan_off, uint32_t* d_block_sums,
                                int* d_status, cudaStream_t s, int* n_launches) {
  if (n_scans == 0) return cudaMemsetAsync(d_scan_off, 0, sizeof(uint32_t), s);
  if (n_az == 0 || n_bins == 0) return cudaMemsetAsync(d_scan_off, 0, ((size_t)n_scans + 1) * sizeof(uint32_t), s);
  FilterDev f; f.min_d = fp.min_range; f.max_d = fp.max_range; f.min_i = fp.min_intensity; f.beam_thr = fp.beam_distance_increment_threshold;
  for (int i = 0; i < 12; ++i) f.tf[i] = fp.sensor_to_base[i];
  k6_row_peak_kernel<<<dim3(n_az, n_scans), kRowThreads, 0, s>>>(d_raw, n_az, n_bins, f.min_d, f.max_d, d_peak, d_angle, d_status);
  k6_runs_kernel<1><<<n_scans, 1024, 0, s>>>(d_raw, n_az, n_bins, f, d_peak, d_angle, d_out, cap, d_counts, nullptr, d_status);
  if (n_launches) *n_launches += 2;
  cudaError_t e = launch_exclusive_scan_u32(d_counts, d_scan_off, n_scans, d_block_sums, s, n_launches);
  if (e != cudaSuccess) return e;
  k6_runs_kernel<2><<<n_scans, 1024, 0, s>>>(d_raw, n_az, n_bins, f, d_peak, d_angle, d_out, cap, d_counts, d_scan_off, d_status);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace randt
