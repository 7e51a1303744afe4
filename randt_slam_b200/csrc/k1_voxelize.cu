// K1 — radar point cloud -> NDT cells (sm_100a).  Compiled with -fmad=false (see k2_associate.cu): the reference's cell
// statistics are sequential float32 sums in point order, and this kernel reproduces that order and rounding exactly.
//
// Replaces, per scan:
//   Grid::cluster                    R/src/radar_preprocessing/grid.cpp:7-14            label = int(x/res) + row*int(y/res)
//   ClusterGenerator::labelClouds    R/src/radar_preprocessing/radar_preprocessor.cpp:151-169   clusters in ascending-label order
//   Map::insertCluster               R/src/ndt_representation/ndt_map.cpp:238-245       keep n > min_points; slot table, later wins
//   Cell::addPointCloud/updateCell   R/src/ndt_representation/ndt_cell.cpp:25-114       mean, population covariance, regularisation
//
// One CTA per scan.  Points are read with coalesced float4 loads; labels are binned with a STABLE parallel counting sort
// (each warp owns a contiguous slice of the scan, per-warp histograms in shared memory, warp-level __match_any_sync ranking)
// so that every cell sees its points in original order; one warp per kept cell then accumulates the two passes in float32 in exactly
// that order (coalesced loads of 32 points, the sequential sum carried through shuffles).  Kept cells are written in ascending-label order into a per-scan padded region and compacted
// across the batch by a second kernel.
#include <float.h>

#include "common.cuh"
#include "affine_device.cuh"

namespace randt {
namespace {

constexpr int kVoxThreads = 1024;  // most threads of the one CTA a scan gets: 32 warps shorten every block-wide phase (the bin scans, the slot-table
                                   // clear) of a lone scan; batches launch 512 so that two scans share an SM (56 registers per thread)
constexpr int kVoxWarps = kVoxThreads / 32;
constexpr int kVoxCountWarps = 8;  // warps that own a slice of the scan in the stable counting sort (one 16-bit histogram each)

// exclusive block scan of a 64-bit value (two 32-bit counters packed: both are scanned in one pass)
__device__ __forceinline__ unsigned long long block_scan_excl(unsigned long long v, unsigned long long* warp_sums, unsigned long long* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n_warps = (int)(blockDim.x >> 5);
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  if (w == 0) {
    unsigned long long s = (lane < n_warps) ? warp_sums[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    warp_sums[lane] = s;
  }
  __syncthreads();
  const unsigned long long base = (w > 0) ? warp_sums[w - 1] : 0ull;
  *total = warp_sums[n_warps - 1];
  __syncthreads();
  return base + x - v;
}

// ---- Eigen 3.3.7 SelfAdjointEigenSolver<Matrix2f> (constructor -> compute(): scale, trivial 2x2 tridiagonalisation, implicit
// QR with Wilkinson shift, ascending sort), as used by Cell::updateCell's regularisation (ndt_cell.cpp:104-110) ----
__device__ __forceinline__ float hypot_eigen(float x, float y) {
  const float ax = fabsf(x), ay = fabsf(y);
  float p, qp;
  if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
  if (p == 0.0f) return 0.0f;
  return p * sqrtf(1.0f + qp * qp);
}
__device__ __forceinline__ void make_givens(float p, float q, float& c, float& s) {
  if (q == 0.0f) { c = p < 0.0f ? -1.0f : 1.0f; s = 0.0f; }
  else if (p == 0.0f) { c = 0.0f; s = q < 0.0f ? 1.0f : -1.0f; }
  else if (fabsf(p) > fabsf(q)) {
    const float t = q / p; float u = sqrtf(1.0f + t * t); if (p < 0.0f) u = -u;
    c = 1.0f / u; s = -t * c;
  } else {
    const float t = p / q; float u = sqrtf(1.0f + t * t); if (q < 0.0f) u = -u;
    s = -1.0f / u; c = -t * s;
  }
}
__device__ void selfadjoint_eig2(float a00, float a10, float a11, float& l0, float& l1, float& v00, float& v01, float& v10, float& v11) {
  float scale = fmaxf(fmaxf(fabsf(a00), fabsf(a10)), fabsf(a11));
  if (scale == 0.0f) scale = 1.0f;
  float d0 = a00 / scale, d1 = a11 / scale, e = a10 / scale;
  v00 = 1.0f; v01 = 0.0f; v10 = 0.0f; v11 = 1.0f;
  const float prec = 2.0f * FLT_EPSILON;
  for (int iter = 0;;) {
    if (fabsf(e) <= (fabsf(d0) + fabsf(d1)) * prec || fabsf(e) <= FLT_MIN) e = 0.0f;
    if (e == 0.0f) break;
    if (++iter > 60) break;
    const float td = (d0 - d1) * 0.5f;
    float mu = d1;
    if (td == 0.0f) mu -= fabsf(e);
    else {
      const float e2 = e * e;
      const float h = hypot_eigen(td, e);
      if (e2 == 0.0f) mu -= (e / (td + (td > 0.0f ? 1.0f : -1.0f))) * (e / h);
      else            mu -= e2 / (td + (td > 0.0f ? h : -h));
    }
    float c, s;
    make_givens(d0 - mu, e, c, s);
    const float sdk = s * d0 + c * e;
    const float dkp1 = s * e + c * d1;
    const float nd0 = c * (c * d0 - s * e) - s * (c * e - s * d1);
    const float nd1 = s * sdk + c * dkp1;
    const float ne = c * sdk - s * dkp1;
    d0 = nd0; d1 = nd1; e = ne;
    if (!(c == 1.0f && s == 0.0f)) {
      const float ns = -s;
      float xi = v00, yi = v01; v00 = c * xi + ns * yi; v01 = s * xi + c * yi;
      xi = v10; yi = v11;       v10 = c * xi + ns * yi; v11 = s * xi + c * yi;
    }
  }
  if (d1 < d0) { float t = d0; d0 = d1; d1 = t; t = v00; v00 = v01; v01 = t; t = v10; v10 = v11; v11 = t; }
  l0 = d0 * scale; l1 = d1 * scale;
}

struct CellOut { float mu[3]; float cov[9]; };

// Cell::updateCell for a fresh cell: two sequential float32 passes over the cell's points (ndt_cell.cpp:43-65), then the
// xy eigenvalue floor and the +1e-6 on the intensity variance (ndt_cell.cpp:102-112).
// One THREAD per cell: the stable counting sort has laid the cell's points (x, y, intensity) out contiguously in scan order in SHARED
// memory, so the thread walks its own run at shared-memory latency, adding in point order: the order (and with it every rounding) is
// exactly the reference's sequential loop, and the 32 lanes of a warp work on 32 cells at once.
__device__ void cell_stats_thread(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pi, uint32_t n, CellOut& o) {
  float sx = 0.f, sy = 0.f, si = 0.f;
#pragma unroll 4
  for (uint32_t k = 0; k < n; ++k) { sx += px[k]; sy += py[k]; si += pi[k]; }
  const float nf = (float)n;
  const float mx = sx / nf, my = sy / nf, mz = si / nf;
  float c00 = 0.f, c11 = 0.f, c22 = 0.f, c01 = 0.f, c02 = 0.f, c12 = 0.f;
#pragma unroll 4
  for (uint32_t k = 0; k < n; ++k) {
    const float dx = px[k] - mx, dy = py[k] - my, di = pi[k] - mz;
    c00 += dx * dx; c11 += dy * dy; c22 += di * di;
    c01 += dx * dy; c02 += dx * di; c12 += dy * di;
  }
  o.mu[0] = mx; o.mu[1] = my; o.mu[2] = mz;
  o.cov[0] = c00 / nf; o.cov[1] = c01 / nf; o.cov[2] = c02 / nf;
  o.cov[3] = c01 / nf; o.cov[4] = c11 / nf; o.cov[5] = c12 / nf;
  o.cov[6] = c02 / nf; o.cov[7] = c12 / nf; o.cov[8] = c22 / nf;
  float l0, l1, v00, v01, v10, v11;
  selfadjoint_eig2(o.cov[0], o.cov[3], o.cov[4], l0, l1, v00, v01, v10, v11);
  l0 = fmaxf(l0, 0.001f * l1);
  // V * diag(l) * V^-1, left to right (the explicit zeros of the diagonal matrix take part in the sums)
  const float t00 = v00 * l0 + v01 * 0.0f, t01 = v00 * 0.0f + v01 * l1;
  const float t10 = v10 * l0 + v11 * 0.0f, t11 = v10 * 0.0f + v11 * l1;
  const float det = v00 * v11 - v10 * v01;
  const float invdet = 1.0f / det;
  const float i00 = v11 * invdet, i10 = -v10 * invdet, i01 = -v01 * invdet, i11 = v00 * invdet;
  o.cov[0] = t00 * i00 + t01 * i10; o.cov[1] = t00 * i01 + t01 * i11;
  o.cov[3] = t10 * i00 + t11 * i10; o.cov[4] = t10 * i01 + t11 * i11;
  o.cov[8] = (float)((double)o.cov[8] + 0.000001);
}

// per-scan status codes: VOX_* in common.cuh

// dynamic shared memory: uint32 bin_start[span_cap] | uint32 bin_keep[span_cap] | uint16 whist[n_cnt_warps][span_cap] | float sx[pt_cap] sy[pt_cap] si[pt_cap]
//   bin_start[bin] = kept points before the bin (where the bin's run starts in sx / sy / si)
//   bin_keep[bin]  = index of the bin's cell among the kept cells of the scan, or kNotKept
constexpr uint32_t kNotKept = 0xffffffffu;
__global__ void __launch_bounds__(kVoxThreads) k1_voxelize_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ scan_off,
                                                                 int row, float label_res, int min_points, MapGeomDev geom,
                                                                 uint32_t span_cap, int n_cnt_warps, uint32_t cell_cap, float4* __restrict__ cells_out,
                                                                 uint32_t* __restrict__ npts_out, int32_t* __restrict__ labels_out,
                                                                 uint32_t* __restrict__ cell_count, int32_t* __restrict__ slot_out,
                                                                 int32_t* __restrict__ labels_scratch, uint32_t pt_cap, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* bin_start = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* bin_keep = bin_start + span_cap;
  unsigned short* whist = reinterpret_cast<unsigned short*>(bin_keep + span_cap);
  float* sx = reinterpret_cast<float*>(whist + (size_t)n_cnt_warps * span_cap);
  float* sy = sx + pt_cap;
  float* si = sy + pt_cap;
  __shared__ unsigned long long warp_sums[32];
  __shared__ int s_min[kVoxWarps], s_max[kVoxWarps];
  __shared__ int s_lab_min, s_lab_max;

  const uint32_t b = blockIdx.x;
  const uint32_t p0 = scan_off[b], p1 = scan_off[b + 1];
  const uint32_t n = p1 - p0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_thr = blockDim.x; const int n_warps = (int)(blockDim.x >> 5);
  int32_t* __restrict__ slot = slot_out + (size_t)b * geom.n_slots;
  {
    // the slot table starts out empty (-1): 16-byte stores where the table is aligned for them
    int4* s4 = reinterpret_cast<int4*>(slot);
    const uint32_t n4 = ((reinterpret_cast<uintptr_t>(slot) & 15u) == 0u) ? geom.n_slots / 4u : 0u;
    for (uint32_t e = tid; e < n4; e += n_thr) s4[e] = make_int4(-1, -1, -1, -1);
    for (uint32_t e = n4 * 4u + tid; e < geom.n_slots; e += n_thr) slot[e] = -1;
  }
  if (tid == 0) { cell_count[b] = 0; status[b] = VOX_OK; }
  if (n == 0) return;
  if (n > pt_cap) { if (tid == 0) status[b] = VOX_SPAN; return; }

  // ---- phase 0: labels (Grid::cluster), min/max label ----
  int lmin = INT_MAX, lmax = INT_MIN;
  for (uint32_t e = tid; e < n; e += n_thr) {
    const float4 p = __ldg(pts + p0 + e);
    const int lab = (int)(p.x / label_res) + row * (int)(p.y / label_res);   // C++ float->int conversion truncates toward zero
    labels_scratch[p0 + e] = lab;
    lmin = min(lmin, lab); lmax = max(lmax, lab);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
  if (lane == 0) { s_min[warp] = lmin; s_max[warp] = lmax; }
  // per-warp histograms cleared while the labels settle
  for (uint32_t e = tid; e < (uint32_t)n_cnt_warps * span_cap / 2u; e += n_thr) reinterpret_cast<uint32_t*>(whist)[e] = 0u;
  __syncthreads();
  if (warp == 0) {
    int a = lane < n_warps ? s_min[lane] : INT_MAX, z = lane < n_warps ? s_max[lane] : INT_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a = min(a, __shfl_xor_sync(0xffffffffu, a, o)); z = max(z, __shfl_xor_sync(0xffffffffu, z, o)); }
    if (lane == 0) { s_lab_min = a; s_lab_max = z; }
  }
  __syncthreads();
  const int lab_min = s_lab_min;
  const long long span_ll = (long long)s_lab_max - (long long)lab_min + 1;
  if (span_ll > (long long)span_cap) { if (tid == 0) status[b] = VOX_SPAN; return; }
  const uint32_t span = (uint32_t)span_ll;

  // ---- phase 1: per-warp histograms over contiguous slices (stable counting sort, pass 1) ----
  const uint32_t slice = (n + n_cnt_warps - 1) / n_cnt_warps;
  if (slice > 65535u) { if (tid == 0) status[b] = VOX_SPAN; return; }
  if (warp < n_cnt_warps) {
    const uint32_t s0 = warp * slice, s1 = min(n, s0 + slice);
    unsigned short* h = whist + (size_t)warp * span_cap;
    // the slice is walked 32 points at a time in scan order (each step updates the histogram the next one reads), but the labels of
    // eight steps are fetched up front: the global-memory latency is paid once per eight steps instead of once per step
    for (uint32_t base = s0; base < s1; base += 32u * 8u) {
      uint32_t bins[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t e = base + 32u * (uint32_t)u + lane;
        bins[u] = e < s1 ? (uint32_t)(labels_scratch[p0 + e] - lab_min) : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (base + 32u * (uint32_t)u >= s1) break;          // warp-uniform
        const uint32_t bin = bins[u];
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 0xffffffffu && (__ffs(peers) - 1) == lane) h[bin] = (unsigned short)(h[bin] + __popc(peers));
        __syncwarp();
      }
    }
  }
  __syncthreads();

  // ---- phase 2: totals per bin -> exclusive scans: points before the bin (bin_start) and kept cells before the bin (ascending label =
  //      cluster order; keep test count > min_points, Cell::addPointCloud) — both counters in one 64-bit scan ----
  unsigned long long carry = 0ull;
  for (uint32_t base = 0; base < span; base += n_thr) {
    const uint32_t bin = base + tid;
    uint32_t tot = 0;
    if (bin < span) for (int w = 0; w < n_cnt_warps; ++w) tot += whist[(size_t)w * span_cap + bin];
    const bool keep = tot > 0u && (long long)tot > (long long)min_points;
    unsigned long long total;
    const unsigned long long ex = carry + block_scan_excl((unsigned long long)(keep ? tot : 0u) | ((unsigned long long)(keep ? 1u : 0u) << 32), warp_sums, &total);
    if (bin < span) {
      bin_start[bin] = (uint32_t)ex;
      const uint32_t ci = (uint32_t)(ex >> 32);
      bin_keep[bin] = keep ? ci : kNotKept;
      if (keep) {
        if (ci < cell_cap) { npts_out[(size_t)b * cell_cap + ci] = tot; labels_out[(size_t)b * cell_cap + ci] = (int32_t)bin + lab_min; }
        else status[b] = VOX_CELL_CAP;
      }
      // turn per-warp counts into per-warp offsets inside the bin
      uint32_t run = 0;
      for (int w = 0; w < n_cnt_warps; ++w) {
        const uint32_t c = whist[(size_t)w * span_cap + bin];
        whist[(size_t)w * span_cap + bin] = (unsigned short)run;
        run += c;
      }
      if (run > 65535u) status[b] = VOX_SPAN;   // a single cell with > 65535 points does not fit the 16-bit offsets
    }
    carry += total;
  }
  __syncthreads();
  const uint32_t n_keep = min((uint32_t)(carry >> 32), cell_cap);

  // ---- phase 3: stable scatter of the points (x, y, intensity) of kept cells into cell-major, scan-ordered runs in shared memory ----
  if (warp < n_cnt_warps) {
    const uint32_t s0 = warp * slice, s1 = min(n, s0 + slice);
    unsigned short* h = whist + (size_t)warp * span_cap;
    for (uint32_t base = s0; base < s1; base += 32u * 4u) {       // labels and points of four steps fetched up front (see phase 1)
      uint32_t bins[4]; float4 pp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t e = base + 32u * (uint32_t)u + lane;
        const bool act = e < s1;
        bins[u] = act ? (uint32_t)(labels_scratch[p0 + e] - lab_min) : 0xffffffffu;
        pp[u] = act ? __ldg(pts + p0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (base + 32u * (uint32_t)u >= s1) break;          // warp-uniform
        const uint32_t bin = bins[u];
        const bool act = bin != 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        const int leader = __ffs(peers) - 1;
        uint32_t off = 0;
        if (act && leader == lane) { off = h[bin]; h[bin] = (unsigned short)(off + __popc(peers)); }
        off = __shfl_sync(0xffffffffu, off, leader);
        if (act && bin_keep[bin] != kNotKept) {
          const uint32_t at = bin_start[bin] + off + __popc(peers & ((1u << lane) - 1u));
          sx[at] = pp[u].x; sy[at] = pp[u].y; si[at] = pp[u].w;
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();

  if (tid == 0) cell_count[b] = n_keep;
  // ---- phase 4: one thread per kept cell: sequential float32 statistics, slot table ----
  for (uint32_t ci = tid; ci < n_keep; ci += n_thr) {
    const uint32_t bin = (uint32_t)(labels_out[(size_t)b * cell_cap + ci] - lab_min);
    const uint32_t cnt = npts_out[(size_t)b * cell_cap + ci];
    const uint32_t at = bin_start[bin];
    CellOut o;
    cell_stats_thread(sx + at, sy + at, si + at, cnt, o);
    const uint32_t s = coord_to_index(geom, o.mu[0], o.mu[1]);
    float4* dst = cells_out + 3 * ((size_t)b * cell_cap + ci);
    dst[0] = make_float4(o.mu[0], o.mu[1], o.mu[2], o.cov[0]);
    dst[1] = make_float4(o.cov[1], o.cov[2], o.cov[3], o.cov[4]);
    dst[2] = make_float4(o.cov[5], o.cov[6], o.cov[7], o.cov[8]);
    // the reference would throw (vector::at) for a mean outside the map; flagged instead, cell kept without a slot
    if (s < geom.n_slots) atomicMax(&slot[s], (int32_t)ci);   // "later cluster wins" == largest kept index
    else status[b] = VOX_OUT_OF_MAP;
  }
}

__global__ void k1_compact_cells_kernel(const float4* __restrict__ cells_p, const uint32_t* __restrict__ npts_p, const int32_t* __restrict__ labels_p,
                                        const uint32_t* __restrict__ cell_off, uint32_t cell_cap, float4* __restrict__ cells,
                                        uint32_t* __restrict__ npts, int32_t* __restrict__ labels) {
  const uint32_t b = blockIdx.y;
  const uint32_t c0 = cell_off[b], nc = cell_off[b + 1] - c0;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const size_t src = (size_t)b * cell_cap + i;
#pragma unroll
  for (int e = 0; e < 3; ++e) cells[3 * (size_t)(c0 + i) + e] = cells_p[3 * src + e];
  npts[c0 + i] = npts_p[src];
  if (labels) labels[c0 + i] = labels_p[src];
}

// grid_indizes_ rebuilt from cell means, cell order, later cell wins (Map::insertCluster semantics)
__global__ void build_slots_kernel(const float4* __restrict__ cells, const uint32_t* __restrict__ cell_off, MapGeomDev geom, int32_t* __restrict__ slot_out) {
  const uint32_t b = blockIdx.y;
  const uint32_t c0 = cell_off[b], nc = cell_off[b + 1] - c0;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const float4 a = cells[3 * (size_t)(c0 + i)];
  const uint32_t s = coord_to_index(geom, a.x, a.y);
  if (s < geom.n_slots) atomicMax(&slot_out[(size_t)b * geom.n_slots + s], (int32_t)i);
}

// One thread per map: the float affine of the reference + the rotation its covariances see -> AffineRec (4 x float4).
//   from_se2d != 0: in = float64 [B][4] Sophus SE2d storage; `Eigen::Affine2f(pose.cast<float>().matrix())` (ndt_matcher.cpp:208, local_fuser.cpp:338):
//                   Sophus' cast re-normalises the float complex (length = hypot(re, im), glibc hypotf = double sqrt rounded once).
//   else          : in = float32 [B][4] (cos, sin, tx, ty) of an Affine2f the caller already holds.
__global__ void prepare_affine_kernel(const void* __restrict__ in, int from_se2d, uint32_t B, float4* __restrict__ out) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (from_se2d) make_affine_rec_se2d(static_cast<const double*>(in) + 4 * (size_t)b, out + 4 * (size_t)b);
  else {
    const float4 a = static_cast<const float4*>(in)[b];
    make_affine_rec(a.x, a.y, a.z, a.w, out + 4 * (size_t)b);
  }
}

// Map::transformMap: every cell of map b by the affine of map b (Cell::transformCell, ndt_cell.cpp:117-123): mean <- t + L mean with the
// linear part as given, covariance <- R cov R^T with R = Eigen's rotation() of it (AffineRec, prepare_affine_kernel)
__global__ void transform_cells_kernel(float4* __restrict__ cells, const uint32_t* __restrict__ cell_off, const float4* __restrict__ aff) {
  const uint32_t b = blockIdx.y;
  const uint32_t c0 = cell_off[b], nc = cell_off[b + 1] - c0;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  float4* p = cells + 3 * (size_t)(c0 + i);
  CellRaw q; q.a = p[0]; q.b = p[1]; q.c = p[2];
  transform_cell_affine(q, aff + 4 * (size_t)b);
  p[0] = q.a; p[1] = q.b; p[2] = q.c;
}

// Cell::operator+= (ndt_cell.h:133-142) of the cell (bmu, bcov, nb) into the accumulated cell (amu, acov, na): unsigned / size_t weights
// converted to float, (na * nb) / (na + nb) in integer arithmetic, population covariances weighted by n - 1.
__device__ __forceinline__ void merge_cell(float* amu, float* acov, uint32_t& na, const float* bmu, const float* bcov, uint32_t nb) {
  const float w1 = (float)((unsigned long long)na - 1ull);
  const float w2 = (float)((unsigned long long)nb - 1ull);
  const float w3 = (float)(((unsigned long long)na * (unsigned long long)nb) / ((unsigned long long)na + (unsigned long long)nb));
  const float d[3] = {amu[0] - bmu[0], amu[1] - bmu[1], amu[2] - bmu[2]};
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) acov[r * 3 + c] = (w1 * acov[r * 3 + c] + w2 * bcov[r * 3 + c]) + w3 * (d[r] * d[c]);
  const float n1 = (float)na, n2 = (float)(unsigned long long)nb, nn = (float)((unsigned long long)na + (unsigned long long)nb);
  for (int r = 0; r < 3; ++r) amu[r] = ((amu[r] * n1) + (bmu[r] * n2)) / nn;
  const uint32_t nsum = na + nb;
  const float dn = (float)(nsum - 1u);
  for (int e = 0; e < 9; ++e) acov[e] /= dn;
  na = nsum;
}
__device__ __forceinline__ void load_cell(const float4* __restrict__ p, float* mu, float* cov) {
  const float4 A = p[0], B = p[1], Cq = p[2];
  mu[0] = A.x; mu[1] = A.y; mu[2] = A.z;
  cov[0] = A.w; cov[1] = B.x; cov[2] = B.y; cov[3] = B.z; cov[4] = B.w; cov[5] = Cq.x; cov[6] = Cq.y; cov[7] = Cq.z; cov[8] = Cq.w;
}

// Map::mergeMapCell + Cell::operator+=, one CTA per map.  The reference walks the moving cells in order; a mover either merges into
// the cell its slot holds or founds a new cell there (which later movers of the same slot then merge into).  Movers that target
// DIFFERENT slots never interact, so the walk splits into independent chains — all movers of one slot, in mover order — and one
// thread runs each chain with the reference's arithmetic; new cells are numbered in the order their founders appear.  Bit-identical
// to the sequential walk.  Output region of map b has room for n_f(b) + n_m(b) cells; the fixed slot table is updated in place.
// dynamic shared memory: uint32 target[nm_cap] (slot of every mover, 0xffffffff: outside the map)
constexpr int kMergeThreads = 256;
__global__ void __launch_bounds__(kMergeThreads) merge_maps_kernel(const float4* __restrict__ f_cells, const uint32_t* __restrict__ f_npts, const uint32_t* __restrict__ f_off,
                                  int32_t* __restrict__ f_slot, const float4* __restrict__ m_cells, const uint32_t* __restrict__ m_npts,
                                  const uint32_t* __restrict__ m_off, MapGeomDev geom, const uint32_t* __restrict__ o_off, float4* __restrict__ o_cells,
                                  uint32_t* __restrict__ o_npts, uint32_t* __restrict__ o_count) {
  extern __shared__ uint32_t target[];
  __shared__ unsigned long long warp_sums[32];
  const uint32_t b = blockIdx.x;
  const uint32_t f0 = f_off[b], nf = f_off[b + 1] - f0, m0 = m_off[b], nm = m_off[b + 1] - m0, o0 = o_off[b];
  const int tid = threadIdx.x;
  // copy the existing fixed cells (all threads)
  for (uint32_t e = tid; e < nf * 3; e += kMergeThreads) o_cells[3 * (size_t)o0 + e] = f_cells[3 * (size_t)f0 + e];
  for (uint32_t e = tid; e < nf; e += kMergeThreads) o_npts[o0 + e] = f_npts[f0 + e];
  for (uint32_t i = tid; i < nm; i += kMergeThreads) {
    const float4 A = m_cells[3 * (size_t)(m0 + i)];
    const uint32_t s = coord_to_index(geom, A.x, A.y);
    target[i] = s < geom.n_slots ? s : 0xffffffffu;
  }
  __syncthreads();
  int32_t* slot = f_slot + (size_t)b * geom.n_slots;
  unsigned long long carry = 0ull;
  for (uint32_t base = 0; base < nm; base += kMergeThreads) {
    const uint32_t i = base + tid;
    bool head = false, founds = false;
    uint32_t s = 0xffffffffu;
    if (i < nm) {
      s = target[i];
      head = s != 0xffffffffu;
      for (uint32_t j = 0; head && j < i; ++j) head = target[j] != s;       // the first mover of its slot runs the slot's chain
      founds = head && slot[s] < 0;
    }
    unsigned long long total;
    const uint32_t rank = (uint32_t)(carry + block_scan_excl(founds ? 1ull : 0ull, warp_sums, &total));
    carry += total;
    if (head) {
      float amu[3], acov[9], bmu[3], bcov[9];
      uint32_t na, idx, j0;
      if (founds) {
        idx = nf + rank;
        load_cell(m_cells + 3 * (size_t)(m0 + i), amu, acov); na = m_npts[m0 + i];
        j0 = i + 1;
      } else {
        idx = (uint32_t)slot[s];
        load_cell(o_cells + 3 * (size_t)(o0 + idx), amu, acov); na = o_npts[o0 + idx];
        j0 = i;
      }
      for (uint32_t j = j0; j < nm; ++j) {
        if (target[j] != s) continue;
        load_cell(m_cells + 3 * (size_t)(m0 + j), bmu, bcov);
        merge_cell(amu, acov, na, bmu, bcov, m_npts[m0 + j]);
      }
      float4* dst = o_cells + 3 * (size_t)(o0 + idx);
      dst[0] = make_float4(amu[0], amu[1], amu[2], acov[0]);
      dst[1] = make_float4(acov[1], acov[2], acov[3], acov[4]);
      dst[2] = make_float4(acov[5], acov[6], acov[7], acov[8]);
      o_npts[o0 + idx] = na;
      if (founds) slot[s] = (int32_t)idx;
    }
  }
  if (tid == 0) o_count[b] = nf + (uint32_t)carry;
}

}  // namespace

cudaError_t launch_voxelize(const float4* d_pts, const uint32_t* d_scan_off, uint32_t n_scans, uint32_t max_pts_per_scan,
                            const randt_grid_params& gp, const MapGeomDev& geom, uint32_t cell_cap_per_scan, float4* d_cells_p,
                            uint32_t* d_npts_p, int32_t* d_labels_p, uint32_t* d_cell_count, int32_t* d_slot, int32_t* d_labels_scratch,
                            int* d_status, cudaStream_t s, int* n_launches) {
  if (n_scans == 0) return cudaSuccess;
  const int row = static_cast<int>(sqrt((double)(size_t)gp.n_clusters));
  if (row <= 0) return cudaErrorInvalidValue;
  const float label_res = gp.max_range * 2 / row;
  // label span bound for points within +-max_range (plus margin); a scan whose labels span more reports VOX_SPAN
  const long long bound = (long long)(row + 4) * (long long)(row + 4);
  uint32_t span_cap = (uint32_t)((bound + 255) / 256 * 256);
  const size_t smem_max = 220 * 1024;
  const uint32_t pt_cap = (max_pts_per_scan + 31u) / 32u * 32u;        // the kept points of a scan are staged in shared memory
  const size_t fixed = (size_t)pt_cap * 12;
  if (fixed + 1024 > smem_max) return cudaErrorInvalidValue;           // > 18 k points in one scan (randt_voxelize refuses such scans)
  if ((size_t)span_cap * (8 + 2) + fixed > smem_max) span_cap = (uint32_t)((smem_max - fixed) / 10 / 256 * 256);
  if (span_cap == 0) return cudaErrorInvalidValue;
  // counting warps: each owns a contiguous slice of the scan and a 16-bit histogram.  As many as leave room for a second CTA on the
  // SM when the batch has more scans than SMs; small scans do not need all of them (fewer histograms to clear and fold).
  const size_t budget = n_scans > (uint32_t)kSmCount ? (size_t)110 * 1024 : smem_max;
  int n_cnt_warps = kVoxCountWarps;
  while (n_cnt_warps > 1 && ((size_t)span_cap * 8 + (size_t)n_cnt_warps * span_cap * 2 + fixed > budget ||
                             (max_pts_per_scan + n_cnt_warps - 1) / n_cnt_warps < 256)) n_cnt_warps >>= 1;
  const size_t smem = (size_t)span_cap * 8 + (size_t)n_cnt_warps * span_cap * 2 + fixed;
  if (smem > smem_max) return cudaErrorInvalidValue;
  cudaError_t e = cudaSuccess;
  if (smem > 48u * 1024u) e = cudaFuncSetAttribute(k1_voxelize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int threads = n_scans > (uint32_t)kSmCount ? 512 : kVoxThreads;     // a batch: two scans per SM; a lone scan: all 32 warps
  k1_voxelize_kernel<<<n_scans, threads, smem, s>>>(d_pts, d_scan_off, row, label_res, gp.min_points, geom, span_cap, n_cnt_warps,
                                                    cell_cap_per_scan, d_cells_p, d_npts_p, d_labels_p, d_cell_count, d_slot,
                                                    d_labels_scratch, pt_cap, d_status);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_compact_cells(const float4* d_cells_p, const uint32_t* d_npts_p, const int32_t* d_labels_p, const uint32_t* d_cell_off,
                                 uint32_t n_scans, uint32_t cell_cap_per_scan, uint32_t max_cells_per_scan, float4* d_cells,
                                 uint32_t* d_npts, int32_t* d_labels, cudaStream_t s, int* n_launches) {
  if (n_scans == 0 || max_cells_per_scan == 0) return cudaSuccess;
  dim3 grid((max_cells_per_scan + 127) / 128, n_scans);
  k1_compact_cells_kernel<<<grid, 128, 0, s>>>(d_cells_p, d_npts_p, d_labels_p, d_cell_off, cell_cap_per_scan, d_cells, d_npts, d_labels);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_build_slots(const float4* d_cells, const uint32_t* d_cell_off, uint32_t n_maps, uint32_t max_per_map,
                               const MapGeomDev& geom, int32_t* d_slot, cudaStream_t s, int* n_launches) {
  if (n_maps == 0 || max_per_map == 0) return cudaSuccess;
  dim3 grid((max_per_map + 127) / 128, n_maps);
  build_slots_kernel<<<grid, 128, 0, s>>>(d_cells, d_cell_off, geom, d_slot);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_prepare_affine(const void* d_in, int from_se2d, uint32_t n_maps, float4* d_aff, cudaStream_t s, int* n_launches) {
  if (n_maps == 0) return cudaSuccess;
  prepare_affine_kernel<<<(n_maps + 127u) / 128u, 128, 0, s>>>(d_in, from_se2d, n_maps, d_aff);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_transform_cells(float4* d_cells, const uint32_t* d_cell_off, uint32_t n_maps, uint32_t max_per_map, const float4* d_aff,
                                   cudaStream_t s, int* n_launches) {
  if (n_maps == 0 || max_per_map == 0) return cudaSuccess;
  dim3 grid((max_per_map + 127) / 128, n_maps);
  transform_cells_kernel<<<grid, 128, 0, s>>>(d_cells, d_cell_off, d_aff);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_merge_maps(const float4* f_cells, const uint32_t* f_npts, const uint32_t* f_off, int32_t* f_slot, const float4* m_cells,
                              const uint32_t* m_npts, const uint32_t* m_off, uint32_t n_maps, const MapGeomDev& geom, const uint32_t* o_off,
                              float4* o_cells, uint32_t* o_npts, uint32_t* o_count, uint32_t max_m_per_map, cudaStream_t s, int* n_launches) {
  if (n_maps == 0) return cudaSuccess;
  const size_t smem = (size_t)std::max(1u, max_m_per_map) * sizeof(uint32_t);
  if (smem > 200u * 1024u) return cudaErrorInvalidValue;       // > 51 200 moving cells in one map
  if (smem > 48u * 1024u) { cudaError_t e = cudaFuncSetAttribute(merge_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
  merge_maps_kernel<<<n_maps, kMergeThreads, smem, s>>>(f_cells, f_npts, f_off, f_slot, m_cells, m_npts, m_off, geom, o_off, o_cells, o_npts, o_count);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace randt
