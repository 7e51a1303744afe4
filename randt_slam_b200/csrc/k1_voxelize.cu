// K1 — radar point cloud -> NDT cells (sm_100a).  Compiled with -fmad=false (see k2_associate.cu): the reference's cell
// statistics are sequential float32 sums in point order, and this kernel reproduces that order and rounding exactly.
//
// Replaces, per scan:
//   Grid::cluster                    R/src/radar_preprocessing/grid.cpp:7-14            label = int(x/res) + row*int(y/res)
//   ClusterGenerator::labelClouds    R/src/radar_preprocessing/radar_preprocessor.cpp:151-169   clusters in ascending-label order
//   Map::insertCluster               R/src/ndt_representation/ndt_map.cpp:238-245       keep n > min_points; slot table, later wins
//   Cell::addPointCloud/updateCell   R/src/ndt_representation/ndt_cell.cpp:25-114       mean, population covariance, regularisation
//
// One CTA per scan.  Labels are binned (bin = label - smallest label of the scan) with warp-aggregated shared-memory atomics
// (__match_any_sync: one update per distinct bin of a 32-point step) that give every bin its point count and a 32-bit mask of the
// parts of the scan its points lie in.  The STABLE sort the reference's order-dependent float32 sums need is then done cell by cell:
// one warp per kept cell walks only the parts its mask names, ballots the points that carry its bin and writes them — in scan order —
// into the cell's run of a cell-major copy of the scan in shared memory (a filtered radar scan is azimuth-major, so a cell's points
// lie within a few beams of each other and a walk is one or two parts; any other point order works, only slower).  One THREAD per kept
// cell then runs the reference's two sequential passes over its run.  Kept cells are written in ascending-label order into a per-scan
// padded region and compacted across the batch by a second kernel.
#include <float.h>

#include "common.cuh"
#include "affine_device.cuh"

namespace randt {
namespace {

constexpr int kVoxThreads = 1024;  // most threads of the one CTA a scan gets: 32 warps shorten every block-wide phase of a lone scan (and
                                   // leave its ~100 cell walks 3-4 to a warp); batches launch 512 so that two scans share an SM
constexpr int kVoxWarps = kVoxThreads / 32;

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
// One THREAD per cell: the cell's points (x, y, intensity) lie contiguously in scan order in SHARED memory, so the thread walks its
// own run at shared-memory latency, adding in point order: the order (and with it every rounding) is exactly the reference's
// sequential loop, and the 32 lanes of a warp work on 32 cells at once.  Eight points are fetched per round so that the loads do
// not sit in the dependent chain of the sums.
__device__ void cell_stats_thread(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pi, uint32_t n, CellOut& o) {
  float sx = 0.f, sy = 0.f, si = 0.f;
  uint32_t k = 0;
  for (; k + 8u <= n; k += 8u) {
    float x[8], y[8], z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = px[k + u]; y[u] = py[k + u]; z[u] = pi[k + u]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) { sx += x[u]; sy += y[u]; si += z[u]; }
  }
  for (; k < n; ++k) { sx += px[k]; sy += py[k]; si += pi[k]; }
  const float nf = (float)n;
  const float mx = sx / nf, my = sy / nf, mz = si / nf;
  float c00 = 0.f, c11 = 0.f, c22 = 0.f, c01 = 0.f, c02 = 0.f, c12 = 0.f;
  k = 0;
  for (; k + 4u <= n; k += 4u) {
    float x[4], y[4], z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { x[u] = px[k + u]; y[u] = py[k + u]; z[u] = pi[k + u]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float dx = x[u] - mx, dy = y[u] - my, di = z[u] - mz;
      c00 += dx * dx; c11 += dy * dy; c22 += di * di;
      c01 += dx * dy; c02 += dx * di; c12 += dy * di;
    }
  }
  for (; k < n; ++k) {
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

// dynamic shared memory: uint32 mask[span_cap] | uint16 cnt[span_cap] | uint16 start[span_cap] | float sx[pt_cap] sy[pt_cap] si[pt_cap] | uint16 bins[pt_cap] | uint16 order[cell_cap]
//   mask[bin]:  bit j set = the bin has points among the scan's points [j part, (j + 1) part)  (part = a whole number of 32-point steps)
//   cnt[bin]:   the bin's point count (two 16-bit counters to a word, added to as one 32-bit atomic: a scan has at most 16 384 points,
//               so a counter cannot carry into its neighbour)
//   start[bin]: where the (kept) bin's run starts in sx / sy / si;  until the sort, the labels (int32) live in the sx array
//   (staged launches append float ox[pt_cap] oy[pt_cap] oi[pt_cap]: the scan in its original order)
//   order[i]:   the kept cells by falling size class (so that the 32 cells a warp takes in phase 4 are about equally long)
__global__ void __launch_bounds__(kVoxThreads) k1_voxelize_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ scan_off,
                                                                 int row, float label_res, int min_points, MapGeomDev geom,
                                                                 uint32_t span_cap, uint32_t cell_cap, float4* __restrict__ cells_out,
                                                                 uint32_t* __restrict__ npts_out, int32_t* __restrict__ labels_out,
                                                                 uint32_t* __restrict__ cell_count, int32_t* __restrict__ slot_out,
                                                                 uint32_t pt_cap, int staged, unsigned short* __restrict__ bins_scratch, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* mask = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* cnt2 = mask + span_cap;                                         // span_cap / 2 words
  unsigned short* start = reinterpret_cast<unsigned short*>(cnt2 + span_cap / 2u);
  float* sx = reinterpret_cast<float*>(start + span_cap);
  float* sy = sx + pt_cap;
  float* si = sy + pt_cap;
  // (a scan too long for 14 bytes of shared memory per point keeps its bins in global scratch: bins_scratch != NULL)
  unsigned short* smem_bins = reinterpret_cast<unsigned short*>(si + pt_cap);
  unsigned short* bins = bins_scratch ? bins_scratch + scan_off[blockIdx.x] : smem_bins;
  unsigned short* order = bins_scratch ? smem_bins : smem_bins + pt_cap;
  // staged (few scans, shared memory to spare): the scan itself also sits in shared memory, so the sort reads it at shared-memory
  // latency instead of going back to L2 once per round
  float* ox = reinterpret_cast<float*>(smem_raw + (((size_t)(reinterpret_cast<unsigned char*>(order + cell_cap) - smem_raw) + 15) & ~(size_t)15));
  float* oy = ox + pt_cap;
  float* oi = oy + pt_cap;
  int* labs = reinterpret_cast<int*>(sx);
  __shared__ uint32_t s_class[32], s_class_at[32];
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

  // ---- phase 0: labels (Grid::cluster), smallest / largest label ----
  int lmin = INT_MAX, lmax = INT_MIN;
  for (uint32_t e = tid; e < n; e += 4u * n_thr) {        // four loads in flight per thread
    float4 p[4];
#pragma unroll
    for (uint32_t u = 0; u < 4u; ++u) if (e + u * n_thr < n) p[u] = __ldg(pts + p0 + e + u * n_thr);
#pragma unroll
    for (uint32_t u = 0; u < 4u; ++u) {
      if (e + u * n_thr < n) {
        const int lab = (int)(p[u].x / label_res) + row * (int)(p[u].y / label_res);   // C++ float->int conversion truncates toward zero
        labs[e + u * n_thr] = lab;
        if (staged) { ox[e + u * n_thr] = p[u].x; oy[e + u * n_thr] = p[u].y; oi[e + u * n_thr] = p[u].w; }
        lmin = min(lmin, lab); lmax = max(lmax, lab);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
  if (lane == 0) { s_min[warp] = lmin; s_max[warp] = lmax; }
  for (uint32_t e = tid; e < span_cap + span_cap / 2u; e += n_thr) mask[e] = 0u;      // mask and cnt2 are adjacent
  if (tid < 32) s_class[tid] = 0u;
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
  const uint32_t n_steps = (n + 31u) >> 5;
  const uint32_t part_steps = (n_steps + 31u) >> 5;          // 32-point steps per mask bit

  // ---- phase 1: bin of every point; per bin the point count and the parts of the scan it occurs in (one shared-memory update per
  //      distinct bin of a 32-point step: points along a beam share their cell) ----
  for (uint32_t e0 = (uint32_t)warp * 32u; e0 < n; e0 += n_thr) {
    const uint32_t e = e0 + (uint32_t)lane;
    const bool act = e < n;
    uint32_t bin = 0xffffffffu;
    if (act) bin = (uint32_t)(labs[e] - lab_min);
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (act) {
      bins[e] = (unsigned short)bin;
      if ((__ffs(peers) - 1) == lane) {
        atomicAdd(&cnt2[bin >> 1], (uint32_t)__popc(peers) << (16u * (bin & 1u)));
        atomicOr(&mask[bin], 1u << ((e0 >> 5) / part_steps));
      }
    }
  }
  __syncthreads();

  // ---- phase 2: kept cells in ascending label (= cluster) order and where their runs start; keep test count > min_points
  //      (Cell::addPointCloud).  Every thread owns a run of consecutive bins; one 64-bit block scan (kept cells | kept points) ----
  const uint32_t per = (span + n_thr - 1u) / n_thr;
  const uint32_t b0 = min(span, (uint32_t)tid * per), b1 = min(span, b0 + per);
  unsigned long long mine_sum = 0ull;
  for (uint32_t bin = b0; bin < b1; ++bin) {
    const uint32_t tot = (cnt2[bin >> 1] >> (16u * (bin & 1u))) & 0xffffu;
    if (tot > 0u && (long long)tot > (long long)min_points) mine_sum += (unsigned long long)tot | (1ull << 32);
  }
  unsigned long long total;
  unsigned long long ex = block_scan_excl(mine_sum, warp_sums, &total);
  for (uint32_t bin = b0; bin < b1; ++bin) {
    const uint32_t tot = (cnt2[bin >> 1] >> (16u * (bin & 1u))) & 0xffffu;
    if (tot > 0u && (long long)tot > (long long)min_points) {
      const uint32_t ci = (uint32_t)(ex >> 32);
      start[bin] = (unsigned short)(uint32_t)ex;
      if (ci < cell_cap) {
        npts_out[(size_t)b * cell_cap + ci] = tot; labels_out[(size_t)b * cell_cap + ci] = (int32_t)bin + lab_min;
        atomicAdd(&s_class[31 - __clz(tot)], 1u);       // size class = floor(log2(points))
      } else status[b] = VOX_CELL_CAP;
      ex += (unsigned long long)tot | (1ull << 32);
    }
  }
  __syncthreads();          // (also: every thread has read its labels; the sx array is free for the sorted points)
  const uint32_t n_keep = min((uint32_t)(total >> 32), cell_cap);
  if (tid == 0) cell_count[b] = n_keep;
  if (warp == 0) {       // where every size class starts in `order`, largest class first
    const uint32_t mine_n = s_class[31 - lane];
    uint32_t x = mine_n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    s_class_at[31 - lane] = x - mine_n;
  }
  __syncthreads();
  for (uint32_t c = tid; c < n_keep; c += n_thr)
    order[atomicAdd(&s_class_at[31 - __clz(npts_out[(size_t)b * cell_cap + c])], 1u)] = (unsigned short)c;

  // ---- phase 3: stable sort, one warp per kept cell: walk the parts of the scan the cell occurs in, ballot the points that carry its
  //      bin, write them in scan order into the cell's run ----
  const unsigned lt = (1u << lane) - 1u;
  for (uint32_t c = (uint32_t)warp; c < n_keep; c += (uint32_t)n_warps) {
    const uint32_t bin = (uint32_t)(labels_out[(size_t)b * cell_cap + c] - lab_min);
    const uint32_t cnt = npts_out[(size_t)b * cell_cap + c];
    uint32_t at = start[bin], left = cnt;
    unsigned m = mask[bin];
    while (m != 0u && left != 0u) {
      const uint32_t j = (uint32_t)__ffs(m) - 1u;
      m &= m - 1u;
      const uint32_t s_end = min(n_steps, (j + 1u) * part_steps);
      for (uint32_t st = j * part_steps; st < s_end && left != 0u; st += 4u) {
        // four steps per round: the point loads of all four are in flight before the first is stored
        bool hit[4]; unsigned bal[4]; float4 p[4];
#pragma unroll
        for (uint32_t u = 0; u < 4u; ++u) {
          const uint32_t e = ((st + u) << 5) + (uint32_t)lane;
          hit[u] = (st + u) < s_end && e < n && bins[e] == (unsigned short)bin;
          bal[u] = __ballot_sync(0xffffffffu, hit[u]);
          if (hit[u]) p[u] = staged ? make_float4(ox[e], oy[e], 0.f, oi[e]) : __ldg(pts + p0 + e);
        }
#pragma unroll
        for (uint32_t u = 0; u < 4u; ++u) {
          if (hit[u]) {
            const uint32_t pos = at + (uint32_t)__popc(bal[u] & lt);
            sx[pos] = p[u].x; sy[pos] = p[u].y; si[pos] = p[u].w;
          }
          const uint32_t k = (uint32_t)__popc(bal[u]);
          at += k; left -= k;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 4: one thread per kept cell: sequential float32 statistics, slot table.  Cells are taken 32 to a warp in falling size
  //      class: a warp runs as long as its longest cell, and the fewest warps are busy (the code is issue-bound, not latency-bound) ----
  for (uint32_t i = tid; i < n_keep; i += n_thr) {
    const uint32_t c = order[i];
    const uint32_t bin = (uint32_t)(labels_out[(size_t)b * cell_cap + c] - lab_min);
    const uint32_t cnt = npts_out[(size_t)b * cell_cap + c];
    const uint32_t at = start[bin];
    CellOut o;
    cell_stats_thread(sx + at, sy + at, si + at, cnt, o);
    const uint32_t s = coord_to_index(geom, o.mu[0], o.mu[1]);
    float4* dst = cells_out + 3 * ((size_t)b * cell_cap + c);
    dst[0] = make_float4(o.mu[0], o.mu[1], o.mu[2], o.cov[0]);
    dst[1] = make_float4(o.cov[1], o.cov[2], o.cov[3], o.cov[4]);
    dst[2] = make_float4(o.cov[5], o.cov[6], o.cov[7], o.cov[8]);
    // the reference would throw (vector::at) for a mean outside the map; flagged instead, cell kept without a slot
    if (s < geom.n_slots) atomicMax(&slot[s], (int32_t)c);   // "later cluster wins" == largest kept index
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

// a scan of this many points no longer fits its bins (2 bytes per point) into shared memory next to its sorted copy: the caller then
// passes n_points x 2 bytes of global scratch
bool voxelize_needs_bins_scratch(uint32_t max_pts_per_scan, uint32_t cell_cap_per_scan, const randt_grid_params& gp) {
  const int row = static_cast<int>(sqrt((double)(size_t)gp.n_clusters));
  const long long bound = (long long)(row + 4) * (long long)(row + 4);
  const size_t span_cap = (size_t)((bound + 255) / 256 * 256);
  const uint32_t pt_cap = (max_pts_per_scan + 31u) / 32u * 32u;
  const size_t fixed = (size_t)pt_cap * 14 + (((size_t)std::min(cell_cap_per_scan, pt_cap) * 2 + 15) & ~(size_t)15);
  return span_cap * 8 + fixed > (size_t)220 * 1024;
}

cudaError_t launch_voxelize(const float4* d_pts, const uint32_t* d_scan_off, uint32_t n_scans, uint32_t max_pts_per_scan,
                            const randt_grid_params& gp, const MapGeomDev& geom, uint32_t cell_cap_per_scan, float4* d_cells_p,
                            uint32_t* d_npts_p, int32_t* d_labels_p, uint32_t* d_cell_count, int32_t* d_slot, unsigned short* d_bins_scratch, int* d_status,
                            cudaStream_t s, int* n_launches) {
  if (n_scans == 0) return cudaSuccess;
  const int row = static_cast<int>(sqrt((double)(size_t)gp.n_clusters));
  if (row <= 0) return cudaErrorInvalidValue;
  const float label_res = gp.max_range * 2 / row;
  // label span bound for points within +-max_range (plus margin); a scan whose labels span more reports VOX_SPAN
  const long long bound = (long long)(row + 4) * (long long)(row + 4);
  uint32_t span_cap = (uint32_t)((bound + 255) / 256 * 256);
  const size_t smem_max = 220 * 1024;
  const uint32_t pt_cap = (max_pts_per_scan + 31u) / 32u * 32u;        // a scan is staged in shared memory: 14 bytes per point (+ 2 per cell)
  const bool bins_global = voxelize_needs_bins_scratch(max_pts_per_scan, cell_cap_per_scan, gp);
  if (bins_global && !d_bins_scratch) return cudaErrorInvalidValue;
  const size_t fixed = (size_t)pt_cap * (bins_global ? 12 : 14) + (((size_t)std::min(cell_cap_per_scan, pt_cap) * 2 + 15) & ~(size_t)15);
  if (fixed + 2048 > smem_max) return cudaErrorInvalidValue;           // > 18 k points in one scan (randt_voxelize refuses such scans)
  if ((size_t)span_cap * 8 + fixed > smem_max) span_cap = (uint32_t)((smem_max - fixed) / 8 / 256 * 256);
  span_cap = std::min<uint32_t>(span_cap, 65280u);                     // bins are 16 bit (0xffff marks "no point")
  if (span_cap == 0) return cudaErrorInvalidValue;
  const bool batch = n_scans > (uint32_t)kSmCount;
  size_t smem = (size_t)span_cap * 8 + fixed;
  const bool staged = !batch && !bins_global && smem + (size_t)pt_cap * 12 + 16 <= smem_max;
  if (staged) smem += (size_t)pt_cap * 12 + 16;
  if (smem > 48u * 1024u) {
    const cudaError_t e = cudaFuncSetAttribute(k1_voxelize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int threads = batch ? 512 : kVoxThreads;     // (256 and 1024 threads per scan measured 20 % and 35 % slower on a 4096-scan batch)     // a batch: two (or more) scans per SM; a lone scan: all 32 warps
  k1_voxelize_kernel<<<n_scans, threads, smem, s>>>(d_pts, d_scan_off, row, label_res, gp.min_points, geom, span_cap, cell_cap_per_scan, d_cells_p,
                                                    d_npts_p, d_labels_p, d_cell_count, d_slot, pt_cap, staged ? 1 : 0, bins_global ? d_bins_scratch : nullptr, d_status);
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
