// K2 — moving-cell -> fixed-cell association (sm_100a).  Compiled with -fmad=false: every float32 operation below is a
// separate IEEE multiply/add in the order the reference's Eigen expressions evaluate them, so distances (and therefore
// the chosen neighbour sets) are reproducible bit for bit.
//
// Replaces, per moving cell, the association half of Matcher::addNDTFactor
//   R/src/ndt_registration/ndt_matcher.cpp:200-217 (Mahalanobis lookup) / :249-253 (Euclidean lookup)
// = Cell::transformCell (R/src/ndt_representation/ndt_cell.cpp:117-123) by the initial guess cast to float,
//   Map::getClosestCells (R/src/ndt_representation/ndt_map.cpp:101-151) with Map::getAdjacentIndizes (:163-175) windows and
//   Cell::mahalanobisSquaredIntensity (ndt_cell.cpp:172-176), first k by (distance, index).
// One thread per moving cell; the fixed map's dense slot table and cells are read through the read-only path.
#include "common.cuh"
#include "affine_device.cuh"
#include "k3_device.cuh"   // RawCell + encode_duo_record (the fused single-map path builds K3's records itself)

namespace randt {
namespace {

struct CellF { float mu[3]; float cov[9]; };

__device__ __forceinline__ CellF load_cell_f(const float4* __restrict__ tab, uint32_t idx) {
  const float4 a = __ldg(tab + 3 * (size_t)idx), b = __ldg(tab + 3 * (size_t)idx + 1), c = __ldg(tab + 3 * (size_t)idx + 2);
  CellF o;
  o.mu[0] = a.x; o.mu[1] = a.y; o.mu[2] = a.z;
  o.cov[0] = a.w; o.cov[1] = b.x; o.cov[2] = b.y; o.cov[3] = b.z; o.cov[4] = b.w; o.cov[5] = c.x; o.cov[6] = c.y; o.cov[7] = c.z; o.cov[8] = c.w;
  return o;
}

// (v^T inv(m)) v with Eigen's cofactor inverse evaluation order (float)
__device__ __forceinline__ float inv3_quadform_f(const float m[9], const float v[3]) {
#define M_(r, c) m[(r) * 3 + (c)]
#define COF_(i, j) (M_(((i) + 1) % 3, ((j) + 1) % 3) * M_(((i) + 2) % 3, ((j) + 2) % 3) - M_(((i) + 1) % 3, ((j) + 2) % 3) * M_(((i) + 2) % 3, ((j) + 1) % 3))
  const float c0 = COF_(0, 0), c1 = COF_(1, 0), c2 = COF_(2, 0);
  const float det = c0 * M_(0, 0) + (c1 * M_(1, 0) + c2 * M_(2, 0));
  const float invdet = 1.0f / det;
  float inv[3][3];
  inv[0][0] = c0 * invdet; inv[0][1] = c1 * invdet; inv[0][2] = c2 * invdet;
  inv[1][0] = COF_(0, 1) * invdet; inv[1][1] = COF_(1, 1) * invdet; inv[1][2] = COF_(2, 1) * invdet;
  inv[2][0] = COF_(0, 2) * invdet; inv[2][1] = COF_(1, 2) * invdet; inv[2][2] = COF_(2, 2) * invdet;
#undef COF_
#undef M_
  float t[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = v[0] * inv[0][j] + (v[1] * inv[1][j] + v[2] * inv[2][j]);
  return t[0] * v[0] + (t[1] * v[1] + t[2] * v[2]);
}

__device__ __forceinline__ bool closer(double d, uint32_t i, double bd, uint32_t bi) {
  // std::sort on pair<double, size_t>: ascending distance, ties by index.  NaN never precedes anything.
  return (d < bd) || (d == bd && i < bi);
}

// Association of moving cell i (absolute index into cells_m) with the fixed map whose cells start at f0 and whose slot table is `slot`:
// the first k occupied slots of the expanding window by (distance, index).  aff: the AffineRec of the initial guess.  -> number found.
__device__ __forceinline__ int associate_cell(const float4* __restrict__ cells_f, uint32_t f0, const int32_t* __restrict__ slot,
                                              const float4* __restrict__ cells_m, uint32_t i, const MapGeomDev& geom, const float4* __restrict__ aff,
                                              int k, int metric, uint32_t* best_i) {
  const float4 T = aff[0];   // AffineRec: (c, s, tx, ty) of initial_guess.cast<float>(), then its Eigen rotation()
  CellF q;
  if (metric == RANDT_LOOKUP_MAHALANOBIS_INTENSITY) {
    CellRaw raw;
    raw.a = __ldg(cells_m + 3 * (size_t)i); raw.b = __ldg(cells_m + 3 * (size_t)i + 1); raw.c = __ldg(cells_m + 3 * (size_t)i + 2);
    transform_cell_affine(raw, aff);
    q.mu[0] = raw.a.x; q.mu[1] = raw.a.y; q.mu[2] = raw.a.z;
    q.cov[0] = raw.a.w; q.cov[1] = raw.b.x; q.cov[2] = raw.b.y; q.cov[3] = raw.b.z; q.cov[4] = raw.b.w; q.cov[5] = raw.c.x; q.cov[6] = raw.c.y; q.cov[7] = raw.c.z; q.cov[8] = raw.c.w;
  } else {
    q = load_cell_f(cells_m, i);
    const float x = q.mu[0], y = q.mu[1];
    q.mu[0] = (T.x * x - T.y * y) + T.z;
    q.mu[1] = (T.y * x + T.x * y) + T.w;
  }
  const uint32_t center = coord_to_index(geom, q.mu[0], q.mu[1]);

  double best_d[kMaxNeighbours];
  int n_best = 0;
  uint32_t total = 0;
  int r = 0;
  while (total < (uint32_t)(k > 0 ? k : 0)) {
    // ring r of the window (all (i,j) with max(|i|,|j|) == r); earlier rings were visited in earlier iterations, which is
    // equivalent to the reference re-collecting the whole (2r+1)^2 window because the final candidate set is sorted.
    for (int di = -r; di <= r; ++di) {
      const int step = (di == -r || di == r) ? 1 : 2 * r;   // full column at the ring's ends, else only top & bottom
      for (int dj = -r; dj <= r; dj += (step > 0 ? step : 1)) {
        const uint32_t ni = center + (uint32_t)di + (uint32_t)dj * (uint32_t)geom.size_x;
        if (ni < geom.n_slots) {
          const int32_t ci = __ldg(slot + ni);
          if (ci >= 0) {
            const CellF f = load_cell_f(cells_f, f0 + (uint32_t)ci);
            double dist;
            if (metric == RANDT_LOOKUP_MAHALANOBIS_INTENSITY) {
              float S[9], d[3];
#pragma unroll
              for (int e = 0; e < 9; ++e) S[e] = f.cov[e] + q.cov[e];
#pragma unroll
              for (int e = 0; e < 3; ++e) d[e] = f.mu[e] - q.mu[e];
              dist = (double)inv3_quadform_f(S, d);
            } else {
              const float dx = q.mu[0] - f.mu[0], dy = q.mu[1] - f.mu[1];
              dist = (double)sqrtf(dx * dx + dy * dy);
            }
            ++total;
            // insert into the sorted top-k list
            int pos = n_best;
            if (n_best == k && !closer(dist, (uint32_t)ci, best_d[k - 1], best_i[k - 1])) pos = -1;
            if (pos >= 0) {
              if (n_best < k) ++n_best;
              int p = n_best - 1;
              while (p > 0 && closer(dist, (uint32_t)ci, best_d[p - 1], best_i[p - 1])) {
                best_d[p] = best_d[p - 1]; best_i[p] = best_i[p - 1]; --p;
              }
              best_d[p] = dist; best_i[p] = (uint32_t)ci;
            }
          }
        }
        if (r == 0) break;
      }
    }
    ++r;
    if (r >= geom.r_stop) break;
  }
  return n_best;
}

__global__ void __launch_bounds__(128) k2_associate_kernel(const float4* __restrict__ cells_f, const uint32_t* __restrict__ cell_off_f,
                                                          const int32_t* __restrict__ slot_f, const float4* __restrict__ cells_m,
                                                          const uint32_t* __restrict__ cell_off_m, MapGeomDev geom,
                                                          const float4* __restrict__ pose_f, int k, int metric, uint32_t* __restrict__ nn,
                                                          uint32_t* __restrict__ cnt) {
  const uint32_t b = blockIdx.y;
  const uint32_t m0 = cell_off_m[b], m1 = cell_off_m[b + 1];
  const uint32_t i = m0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m1) return;
  uint32_t best_i[kMaxNeighbours];
  const int n_best = associate_cell(cells_f, cell_off_f[b], slot_f + (size_t)b * geom.n_slots, cells_m, i, geom, pose_f + 4 * (size_t)b, k, metric, best_i);
  cnt[i] = (uint32_t)n_best;
  for (int t = 0; t < n_best; ++t) nn[(size_t)i * k + t] = best_i[t];
}

// ---- single map pair, everything in one launch ------------------------------------------------------------------------------
// The per-scan call of a live stream (one moving scan against one submap): the AffineRec of the initial guess, the association,
// the pair / duo lists in the reference's block order, K3's compact records (a lone registration's tiles sit in duo order: one tile
// per warp), the snapshots of both cell tables and the totals the host needs — one CTA, one launch, one read-back, instead of ~15
// launches and three host synchronisations.
constexpr int kSingleThreads = 1024;
__global__ void __launch_bounds__(kSingleThreads) k2_associate_single_kernel(const float4* __restrict__ cells_f, uint32_t n_f, const int32_t* __restrict__ slot,
                                                                            const float4* __restrict__ cells_m, uint32_t n_m, MapGeomDev geom,
                                                                            const double* __restrict__ pose0, int k, int metric, uint2* __restrict__ pairs,
                                                                            Duo* __restrict__ duos, DuoRec* __restrict__ recs, uint32_t* __restrict__ duo_p0,
                                                                            DuoRecFull* __restrict__ overflow, uint32_t overflow_cap, float4* __restrict__ snap_m,
                                                                            float4* __restrict__ snap_f, uint32_t* __restrict__ totals /* P, n_duos, n_overflow */,
                                                                            uint32_t* __restrict__ layout /* or NULL: the one-registration layout K7 walks */,
                                                                            const uint32_t* __restrict__ n_m_dev /* or NULL: the moving map's size, still on the device */,
                                                                            double ndt_weight, double* __restrict__ weight_out) {
  // chained behind K1 without a host round trip, the size of the moving map is read from where K1 left it, and the ScaledLoss
  // weight of the scan, ndt_weight / (n_cells k) (ndt_matcher.cpp:392), is left for K7 next to the records
  if (n_m_dev) n_m = *n_m_dev;
  if (weight_out && threadIdx.x == 0 && n_m > 0u) *weight_out = ndt_weight / ((double)n_m * (double)k);
  __shared__ float4 aff[4];
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t warp_sums2[32];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (tid == 0) make_affine_rec_se2d(pose0, aff);
  // the snapshots (the reference's functors copy their cells by value)
  if (snap_m) {
    for (uint32_t e = tid; e < 3u * n_m; e += kSingleThreads) snap_m[e] = __ldg(cells_m + e);
    for (uint32_t e = tid; e < 3u * n_f; e += kSingleThreads) snap_f[e] = __ldg(cells_f + e);
  }
  __syncthreads();
  uint32_t carry_p = 0, carry_d = 0;
  for (uint32_t base = 0; base < n_m; base += kSingleThreads) {
    const uint32_t i = base + tid;
    uint32_t best_i[kMaxNeighbours];
    int n = 0;
    if (i < n_m) n = associate_cell(cells_f, 0u, slot, cells_m, i, geom, aff, k, metric, best_i);
    // exclusive block scans of the pair and duo counts (two warp-shuffle scans, one round of barriers)
    const uint32_t cp = (uint32_t)n, cd = ((uint32_t)n + 1u) >> 1;
    uint32_t xp = cp, xd = cd;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t yp = __shfl_up_sync(0xffffffffu, xp, o), yd = __shfl_up_sync(0xffffffffu, xd, o);
      if (lane >= o) { xp += yp; xd += yd; }
    }
    if (lane == 31) { warp_sums[wp] = xp; warp_sums2[wp] = xd; }
    __syncthreads();
    if (wp == 0) {
      uint32_t sp = warp_sums[lane], sd = warp_sums2[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t yp = __shfl_up_sync(0xffffffffu, sp, o), yd = __shfl_up_sync(0xffffffffu, sd, o);
        if (lane >= o) { sp += yp; sd += yd; }
      }
      warp_sums[lane] = sp; warp_sums2[lane] = sd;
    }
    __syncthreads();
    const uint32_t pbase = carry_p + (wp > 0 ? warp_sums[wp - 1] : 0u) + xp - cp;
    const uint32_t dbase = carry_d + (wp > 0 ? warp_sums2[wp - 1] : 0u) + xd - cd;
    carry_p += warp_sums[31]; carry_d += warp_sums2[31];
    __syncthreads();
    if (i < n_m) {
      RawCell c[3];
      c[0].a = __ldg(cells_m + 3 * (size_t)i); c[0].b = __ldg(cells_m + 3 * (size_t)i + 1); c[0].c = __ldg(cells_m + 3 * (size_t)i + 2);
      if (pairs) for (int t = 0; t < n; ++t) pairs[pbase + t] = make_uint2(i, best_i[t]);
      for (int t = 0; t < n; t += 2) {
        const bool two = t + 1 < n;
        Duo d;
        d.im = i; d.jf0 = best_i[t]; d.jf1 = two ? best_i[t + 1] : kNoCell; d.p0 = pbase + t;
        const uint32_t di = dbase + (uint32_t)(t >> 1);
        if (duos) { duos[di] = d; duo_p0[di] = d.p0; }
        c[1].a = __ldg(cells_f + 3 * (size_t)d.jf0); c[1].b = __ldg(cells_f + 3 * (size_t)d.jf0 + 1); c[1].c = __ldg(cells_f + 3 * (size_t)d.jf0 + 2);
        c[2] = c[1];
        if (two) { c[2].a = __ldg(cells_f + 3 * (size_t)d.jf1); c[2].b = __ldg(cells_f + 3 * (size_t)d.jf1 + 1); c[2].c = __ldg(cells_f + 3 * (size_t)d.jf1 + 2); }
        encode_duo_record(c, two, recs + di, overflow, overflow_cap, totals + 2);
      }
    }
  }
  if (tid == 0) {
    totals[0] = carry_p; totals[1] = carry_d;
    if (layout) {     // seg_off[2], seg_duo_off[2], seg_first_tile[1], tile_rec_begin[1]: one registration, one tile, records in duo order
      layout[0] = 0u; layout[1] = carry_p; layout[2] = 0u; layout[3] = carry_d; layout[4] = 0u; layout[5] = 0u;
    }
  }
}

__global__ void k2_compact_pairs_kernel(const uint32_t* __restrict__ nn, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ scan,
                                        const uint32_t* __restrict__ cell_off_m, const uint32_t* __restrict__ cell_off_f, int k,
                                        uint2* __restrict__ pairs) {
  const uint32_t b = blockIdx.y;
  const uint32_t m0 = cell_off_m[b], m1 = cell_off_m[b + 1];
  const uint32_t i = m0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m1) return;
  const uint32_t f0 = cell_off_f[b];
  const uint32_t base = scan[i], n = cnt[i];
  for (uint32_t t = 0; t < n; ++t) pairs[base + t] = make_uint2(i, f0 + nn[(size_t)i * k + t]);
}

// duos of moving cell i: its neighbours two by two (the grouping K3 evaluates, see common.cuh)
__global__ void k2_duo_counts_kernel(const uint32_t* __restrict__ cnt, uint32_t n, uint32_t* __restrict__ cnt2) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt2[i] = (cnt[i] + 1u) >> 1;
}
// out[b] = scan[cell_off[b]], out[n_off + b] = scan2[cell_off[b]]: the per-map pair / duo offsets, so that the host reads B+1 values
// per table instead of the whole scans
__global__ void k2_gather_offsets_kernel(const uint32_t* __restrict__ scan, const uint32_t* __restrict__ scan2,
                                         const uint32_t* __restrict__ cell_off, uint32_t n_off, uint32_t* __restrict__ out) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_off) return;
  const uint32_t c = cell_off[b];
  out[b] = scan[c]; out[n_off + b] = scan2[c];
}
__global__ void k2_compact_duos_kernel(const uint32_t* __restrict__ nn, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ scan,
                                       const uint32_t* __restrict__ scan2, const uint32_t* __restrict__ cell_off_m,
                                       const uint32_t* __restrict__ cell_off_f, int k, Duo* __restrict__ duos) {
  const uint32_t b = blockIdx.y;
  const uint32_t m0 = cell_off_m[b], m1 = cell_off_m[b + 1];
  const uint32_t i = m0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m1) return;
  const uint32_t f0 = cell_off_f[b];
  const uint32_t base = scan[i], n = cnt[i], dbase = scan2[i];
  for (uint32_t t = 0; t < n; t += 2) {
    Duo d;
    d.im = i; d.jf0 = f0 + nn[(size_t)i * k + t];
    d.jf1 = (t + 1 < n) ? f0 + nn[(size_t)i * k + t + 1] : kNoCell;
    d.p0 = base + t;
    duos[dbase + (t >> 1)] = d;
  }
}

// ---- exclusive scan of u32 (three small kernels; construction-time only) -------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    warp_sums[lane] = s;
  }
  __syncthreads();
  const uint32_t base = (w > 0) ? warp_sums[w - 1] : 0u;
  if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
  __syncthreads();
  return base + x - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_local_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                                                  uint32_t* __restrict__ block_sums) {
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  uint32_t v[kScanItems], s = 0;
#pragma unroll
  for (int e = 0; e < kScanItems; ++e) { v[e] = (base + e < n) ? in[base + e] : 0u; s += v[e]; }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, &total);
#pragma unroll
  for (int e = 0; e < kScanItems; ++e) { if (base + e < n) out[base + e] = ex; ex += v[e]; }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(uint32_t* __restrict__ block_sums, uint32_t n_blocks, uint32_t* __restrict__ total_out) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_blocks; base += kScanThreads) {
    const uint32_t idx = base + threadIdx.x;
    const uint32_t v = idx < n_blocks ? block_sums[idx] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, &total);
    if (idx < n_blocks) block_sums[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ block_sums) {
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
  for (int e = 0; e < kScanItems; ++e) if (base + e < n) out[base + e] += add;
}

// randt_problem_concat: the pairs and duos of one part, shifted into the joint tables (moving / fixed cell indices by the rows already
// there, pair indices by the pairs already there)
__global__ void __launch_bounds__(256) shift_part_kernel(const uint2* __restrict__ pairs_in, uint32_t n_pairs, const Duo* __restrict__ duos_in, uint32_t n_duos,
                                                         uint32_t m_base, uint32_t f_base, uint32_t p_base, uint2* __restrict__ pairs_out,
                                                         Duo* __restrict__ duos_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pairs) { const uint2 v = pairs_in[i]; pairs_out[i] = make_uint2(v.x + m_base, v.y + f_base); }
  if (i < n_duos) {
    Duo d = duos_in[i];
    d.im += m_base; d.jf0 += f_base; if (d.jf1 != kNoCell) d.jf1 += f_base; d.p0 += p_base;
    duos_out[i] = d;
  }
}

}  // namespace

cudaError_t launch_associate(const float4* cells_f, const uint32_t* cell_off_f, const int32_t* slot_f, const float4* cells_m,
                             const uint32_t* cell_off_m, uint32_t n_maps, uint32_t n_m_total, uint32_t max_m_per_map,
                             const MapGeomDev& geom, const float4* d_pose_f, int k, int metric, uint32_t* d_nn, uint32_t* d_cnt,
                             cudaStream_t s, int* n_launches) {
  if (n_maps == 0 || max_m_per_map == 0) return cudaSuccess;
  dim3 grid((max_m_per_map + 127) / 128, n_maps);
  k2_associate_kernel<<<grid, 128, 0, s>>>(cells_f, cell_off_f, slot_f, cells_m, cell_off_m, geom, d_pose_f, k, metric, d_nn, d_cnt);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_associate_single(const float4* cells_f, uint32_t n_f, const int32_t* slot_f, const float4* cells_m, uint32_t n_m,
                                    const MapGeomDev& geom, const double* d_pose0, int k, int metric, uint2* d_pairs, Duo* d_duos, DuoRec* d_recs,
                                    uint32_t* d_duo_p0, DuoRecFull* d_overflow, uint32_t overflow_cap, float4* d_snap_m, float4* d_snap_f,
                                    uint32_t* d_totals, uint32_t* d_layout, const uint32_t* d_n_m, double ndt_weight, double* d_weight, cudaStream_t s,
                                    int* n_launches) {
  k2_associate_single_kernel<<<1, kSingleThreads, 0, s>>>(cells_f, n_f, slot_f, cells_m, n_m, geom, d_pose0, k, metric, d_pairs, d_duos, d_recs, d_duo_p0,
                                                         d_overflow, overflow_cap, d_snap_m, d_snap_f, d_totals, d_layout, d_n_m, ndt_weight, d_weight);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_compact_pairs(const uint32_t* d_nn, const uint32_t* d_cnt, const uint32_t* d_scan, const uint32_t* cell_off_m,
                                 const uint32_t* cell_off_f, uint32_t n_maps, uint32_t n_m_total, uint32_t max_m_per_map, int k,
                                 uint2* d_pairs, cudaStream_t s, int* n_launches) {
  if (n_maps == 0 || max_m_per_map == 0) return cudaSuccess;
  dim3 grid((max_m_per_map + 127) / 128, n_maps);
  k2_compact_pairs_kernel<<<grid, 128, 0, s>>>(d_nn, d_cnt, d_scan, cell_off_m, cell_off_f, k, d_pairs);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_duo_counts(const uint32_t* d_cnt, uint32_t n, uint32_t* d_cnt2, cudaStream_t s, int* n_launches) {
  if (n == 0) return cudaSuccess;
  k2_duo_counts_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_cnt, n, d_cnt2);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_gather_offsets(const uint32_t* d_scan, const uint32_t* d_scan2, const uint32_t* cell_off, uint32_t n_off, uint32_t* d_out,
                                  cudaStream_t s, int* n_launches) {
  if (n_off == 0) return cudaSuccess;
  k2_gather_offsets_kernel<<<(n_off + 255) / 256, 256, 0, s>>>(d_scan, d_scan2, cell_off, n_off, d_out);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_compact_duos(const uint32_t* d_nn, const uint32_t* d_cnt, const uint32_t* d_scan, const uint32_t* d_scan2,
                                const uint32_t* cell_off_m, const uint32_t* cell_off_f, uint32_t n_maps, uint32_t n_m_total,
                                uint32_t max_m_per_map, int k, Duo* d_duos, cudaStream_t s, int* n_launches) {
  if (n_maps == 0 || max_m_per_map == 0) return cudaSuccess;
  dim3 grid((max_m_per_map + 127) / 128, n_maps);
  k2_compact_duos_kernel<<<grid, 128, 0, s>>>(d_nn, d_cnt, d_scan, d_scan2, cell_off_m, cell_off_f, k, d_duos);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

// d_out has n+1 entries: exclusive scan plus the grand total at d_out[n].  d_block_sums: >= ceil(n/1024) entries.
cudaError_t launch_exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, uint32_t n, uint32_t* d_block_sums, cudaStream_t s,
                                      int* n_launches) {
  if (n == 0) return cudaMemsetAsync(d_out, 0, sizeof(uint32_t), s);
  const uint32_t n_blocks = (n + kScanBlock - 1) / kScanBlock;
  scan_local_kernel<<<n_blocks, kScanThreads, 0, s>>>(d_in, d_out, n, d_block_sums);
  scan_sums_kernel<<<1, kScanThreads, 0, s>>>(d_block_sums, n_blocks, d_out + n);
  scan_add_kernel<<<n_blocks, kScanThreads, 0, s>>>(d_out, n, d_block_sums);
  if (n_launches) *n_launches += 3;
  return cudaGetLastError();
}

cudaError_t launch_shift_part(const uint2* pairs_in, uint32_t n_pairs, const Duo* duos_in, uint32_t n_duos, uint32_t m_base, uint32_t f_base,
                              uint32_t p_base, uint2* pairs_out, Duo* duos_out, cudaStream_t s, int* n_launches) {
  const uint32_t n = n_pairs > n_duos ? n_pairs : n_duos;
  if (n == 0) return cudaSuccess;
  shift_part_kernel<<<(n + 255) / 256, 256, 0, s>>>(pairs_in, n_pairs, duos_in, n_duos, m_base, f_base, p_base, pairs_out, duos_out);
  if (n_launches) ++*n_launches;
  return cudaGetLastError();
}

}  // namespace randt
