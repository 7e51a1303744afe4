// C-ABI layer: contexts, device-resident maps and problems, and the extern "C" entry points declared in include/randt_gpu.h.
// Host code only orchestrates (stream-ordered allocation, the per-map offsets, the tile schedule and chunk lists, the polling of the
// batched solver, the copy pipeline of the async evaluation); all per-point / per-cell / per-pair / per-registration work runs in the
// kernels of k1_voxelize.cu, k2_associate.cu, k3_pair_eval.cu, k4_lm_step.cu, k5_cs_divergence.cu and k6_filter_scan.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <new>
#include <queue>
#include <utility>

#include "common.cuh"

using namespace randt;

// The stream of a context, shared with the maps and problems created on it: an owned stream outlives its context for as long as such
// objects exist, so that their destroy functions can release device memory in stream order (cudaFreeAsync) instead of through the
// device-wide synchronisation of cudaFree.  Also carries the small pinned/event resources the batched solver polls with.
struct StreamRef {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own = false;
  cudaMemPool_t pool = nullptr;              // this context's own stream-ordered pool (release threshold: never trim)
  uint32_t* h_n_active = nullptr;            // pinned [2]
  cudaEvent_t ev[2] = {nullptr, nullptr};
  // input pipeline of randt_eval_fused_async: a copy stream and a two-slot device ring for the poses of the calls in flight
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  double* ring[2] = {nullptr, nullptr}; size_t ring_cap[2] = {0, 0}; uint64_t ring_n = 0;
  const void* known_pinned[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; unsigned known_n = 0;   // buffers already checked
  cudaStream_t copy_out = nullptr; cudaEvent_t ev_d2h[2] = {nullptr, nullptr};
  cudaEvent_t ev_call[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // delivery of call t: ev_call[t & 7]
  double* oring[2] = {nullptr, nullptr}; size_t oring_cap[2] = {0, 0};
  ~StreamRef() {
    cudaSetDevice(device);
    if (copy) cudaStreamSynchronize(copy);
    if (copy_out) cudaStreamSynchronize(copy_out);
    if (own && stream) cudaStreamSynchronize(stream);
    for (int i = 0; i < 2; ++i) {
      if (ev_d2h[i]) cudaEventDestroy(ev_d2h[i]);
      for (int q = i; q < 8; q += 2) if (ev_call[q]) cudaEventDestroy(ev_call[q]);
      if (oring[i]) cudaFree(oring[i]);
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
      if (ring[i]) cudaFree(ring[i]);
    }
    if (copy) cudaStreamDestroy(copy);
    if (copy_out) cudaStreamDestroy(copy_out);
    if (h_n_active) cudaFreeHost(h_n_active);
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    if (pool) cudaMemPoolDestroy(pool);      // every map / problem that allocated from it held a reference to this object
    if (own && stream) cudaStreamDestroy(stream);
  }
};

struct randt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::shared_ptr<StreamRef> sref;
  std::string err;
  uint64_t launches = 0;
  unsigned long long* d_bad = nullptr;   // count of degenerate pairs seen by K3
  // scratch of finish_problem / randt_associate, kept between calls: the schedule builder's vectors, a pinned staging block for the
  // one upload a problem's schedule needs (stage_ev: that upload has left the block) and a pinned block for offset read-backs
  Schedule sched;
  uint32_t* h_stage = nullptr; size_t stage_cap = 0; cudaEvent_t stage_ev = nullptr; bool stage_busy = false;
  uint32_t* h_offs = nullptr; size_t offs_cap = 0; cudaEvent_t offs_ev = nullptr;
};

struct randt_map {
  int device = 0;
  std::shared_ptr<StreamRef> sref;   // stream of the creating context
  randt_grid_params gp{};
  MapGeomDev geom{};
  uint32_t B = 0, n_cells = 0, max_per_map = 0;
  bool has_npts = true;          // false: uploaded without point counts (cannot take part in a merge)
  float4* cells = nullptr;       // [n_cells][3]
  uint32_t* npts = nullptr;      // [n_cells]
  int32_t* labels = nullptr;     // [n_cells] voxel labels (voxelised maps only)
  uint32_t* cell_off = nullptr;  // [B+1] device
  int32_t* slot = nullptr;       // [B][n_slots]
  std::vector<uint32_t> h_cell_off;
};

struct randt_problem {
  int device = 0;
  std::shared_ptr<StreamRef> sref;   // stream of the creating context
  uint32_t S = 0, P = 0, n_m = 0, n_f = 0;
  float4 *cells_m = nullptr, *cells_f = nullptr;
  uint2* pairs = nullptr;
  Duo* duos = nullptr; uint32_t n_duos = 0;
  DuoRec* duo_recs = nullptr; uint32_t* duo_p0 = nullptr;
  DuoRecFull* duo_overflow = nullptr; uint32_t n_overflow = 0;   // full-precision records of the duos the compact form cannot hold
  std::vector<uint32_t> h_duo_off;   // [S+1] duo offsets per segment
  uint32_t n_tiles = 0;
  ChunkDesc* chunks = nullptr; uint32_t n_chunks = 0;             // plan B: one tile per chunk (solver, EMIT)
  ChunkDesc* chunks_full = nullptr; uint32_t n_chunks_full = 0;   // plan A: split chunks allowed (full FUSED evaluation)
  uint32_t *warp_off = nullptr, *warp_off_full = nullptr; uint32_t n_warps = 0;
  uint32_t* seg_first_tile = nullptr;
  uint32_t* seg_off = nullptr;
  uint32_t *seg_duo_off = nullptr, *rec_of_tile = nullptr; uint32_t tile_duos = 0;   // what the persistent solver walks
  uint32_t *solve_items = nullptr, *solve_counter = nullptr; uint32_t n_solve_items = 0; bool solve_all = true;   // segments short enough for it
  uint32_t* sched_blob = nullptr;     // one allocation behind warp_off .. solve_items (views into it)
  double* work_blob = nullptr;        // one allocation behind partials, d_poses, d_out, d_mu, seg_counters, solve_counter
  unsigned char* assoc_blob = nullptr;
  double* partials = nullptr;
  uint32_t* seg_counters = nullptr;
  std::vector<uint32_t> h_seg_off;
  bool has_empty_segment = false;
  // scratch for the host-pointer entry points
  double *d_poses = nullptr, *d_out = nullptr, *d_mu = nullptr, *d_r = nullptr, *d_J = nullptr;
  double* d_sweep = nullptr; size_t sweep_cap = 0;
  double* lm_weight = nullptr;   // [S] per-registration loss weights (randt_register_batch_weighted)
  // workspace of the batched solver (randt_register_batch), allocated on first use
  LmState* lm_state = nullptr; double *lm_eval_pose = nullptr, *lm_mu = nullptr, *lm_rec = nullptr, *lm_poses = nullptr, *lm_result = nullptr;
  uint32_t *lm_active = nullptr, *lm_n_active = nullptr;
  ChunkDesc* lm_chunks = nullptr; uint32_t *lm_flags = nullptr, *lm_scan = nullptr, *lm_bs = nullptr, *lm_warp_off = nullptr;   // re-planned schedule
};

namespace {

constexpr uint32_t kSingleMaxCells = 8192;   // the fused one-CTA association serves map pairs up to this many cells a side

int fail(randt_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (ctx) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof(buf), "%s", what);
    ctx->err = buf;
  }
  return code;
}

#define CK(call)                                                                 \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) return fail(ctx, RANDT_E_CUDA, #call, e__);          \
  } while (0)

// Device memory comes from the stream-ordered pool (cudaMallocAsync) while an API call is running on a context: allocation and
// release are then queue operations on the context's stream instead of device-wide synchronisations, which is what the per-scan
// calls (voxelise, associate, merge) are made of.  Outside a call (the destroy functions) plain cudaFree is used, which is valid
// for pool memory as well.
thread_local cudaStream_t t_stream = nullptr;
thread_local cudaMemPool_t t_pool = nullptr;
thread_local bool t_in_call = false;
struct StreamScope {
  cudaStream_t prev; cudaMemPool_t prev_pool; bool prev_in;
  StreamScope(cudaStream_t s, cudaMemPool_t pool) : prev(t_stream), prev_pool(t_pool), prev_in(t_in_call) { t_stream = s; t_pool = pool; t_in_call = true; }
  ~StreamScope() { t_stream = prev; t_pool = prev_pool; t_in_call = prev_in; }
};
template <typename T>
cudaError_t dev_alloc(T** p, size_t n) {
  *p = nullptr; if (n == 0) n = 1;
  if (t_in_call && t_pool) return cudaMallocFromPoolAsync(reinterpret_cast<void**>(p), n * sizeof(T), t_pool, t_stream);
  if (t_in_call) return cudaMallocAsync(reinterpret_cast<void**>(p), n * sizeof(T), t_stream);
  return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
}
inline void dev_free(void* p) {
  if (!p) return;
  if (t_in_call) cudaFreeAsync(p, t_stream); else cudaFree(p);
}

bool same_geom(const randt_grid_params& a, const randt_grid_params& b) {
  return a.size_x == b.size_x && a.size_y == b.size_y && a.resolution == b.resolution;
}

void free_map(randt_map* m) {
  if (!m) return;
  dev_free(m->cells); dev_free(m->npts); dev_free(m->labels); dev_free(m->cell_off); dev_free(m->slot);
  delete m;
}
void free_problem(randt_problem* p) {
  if (!p) return;
  dev_free(p->cells_m); dev_free(p->cells_f); dev_free(p->pairs); dev_free(p->duos); dev_free(p->duo_recs); dev_free(p->duo_p0); dev_free(p->duo_overflow);
  dev_free(p->chunks); dev_free(p->chunks_full);
  dev_free(p->sched_blob);     // warp_off, warp_off_full, seg_first_tile, seg_off, seg_duo_off, rec_of_tile, solve_items
  dev_free(p->work_blob);      // partials, d_poses, d_out, d_mu, seg_counters, solve_counter
  dev_free(p->d_r); dev_free(p->d_J); dev_free(p->d_sweep); dev_free(p->lm_weight);
  dev_free(p->lm_state); dev_free(p->lm_eval_pose); dev_free(p->lm_mu); dev_free(p->lm_rec); dev_free(p->lm_poses); dev_free(p->lm_result);
  dev_free(p->lm_active); dev_free(p->lm_n_active);
  dev_free(p->lm_chunks); dev_free(p->lm_flags); dev_free(p->lm_scan); dev_free(p->lm_bs); dev_free(p->lm_warp_off);
  delete p;
}

// a pinned host block of the context, grown on demand
int pinned_reserve(randt_ctx* ctx, uint32_t** buf, size_t* cap, size_t words) {
  if (*cap >= words) return RANDT_OK;
  if (*buf) { cudaFreeHost(*buf); *buf = nullptr; *cap = 0; }
  const size_t want = std::max<size_t>(words + words / 2, 4096);
  if (cudaHostAlloc(reinterpret_cast<void**>(buf), want * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) return fail(ctx, RANDT_E_NOMEM, "pinned staging block");
  *cap = want;
  return RANDT_OK;
}

// tile list, balanced schedule, record table and per-segment bookkeeping from the host offsets of a problem.  The host computes the
// tile assignment only (schedule.hpp build_schedule_core, vectors reused between calls); everything the device needs goes up in ONE
// copy from a pinned block, and the chunk lists of both plans are written by a kernel.
// Resident warps the K3 schedule is built for.  RANDT_K3_MAX_WARPS (diagnostic, read per problem) lowers it, so that a small problem gives
// every warp a long chunk list (tests walk the descriptor queue through several refills that way).
uint32_t k3_warp_budget() {
  const char* e = getenv("RANDT_K3_MAX_WARPS");
  if (e) { const long v = strtol(e, nullptr, 10); if (v >= 1 && v < (long)kK3MaxWarps) return (uint32_t)v; }
  return (uint32_t)kK3MaxWarps;
}

int finish_problem(randt_ctx* ctx, randt_problem* p, bool records_ready = false) {
  static const bool trace = getenv("RANDT_DEBUG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[randt] finish_problem %-18s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_prev).count());
    t_prev = t;
  };
  Schedule& sch = ctx->sched;
  const uint32_t S = p->S;
  build_schedule_core(p->h_duo_off.data(), S, k3_warp_budget(), sch);     // schedule.hpp: tiles, LPT assignment, record layout, chunk ranges
  const uint32_t W = sch.n_warps, T = (uint32_t)sch.tiles.size();
  lap("schedule");
  p->n_warps = W; p->n_tiles = T;
  p->n_chunks = sch.woffB[W]; p->n_chunks_full = sch.woffA[W];
  p->tile_duos = sch.tile_duos;
  for (uint32_t s = 0; s < S; ++s) if (p->h_seg_off[s + 1] == p->h_seg_off[s]) p->has_empty_segment = true;
  // the fused single-map association wrote its records in duo order: valid when the schedule streams the tiles in that order
  if (records_ready && !std::equal(sch.tile_duo_begin.begin(), sch.tile_duo_begin.end(), sch.tile_rec_begin.begin())) {
    records_ready = false;
    dev_free(p->duo_recs); dev_free(p->duo_p0); dev_free(p->duo_overflow);
    p->duo_recs = nullptr; p->duo_p0 = nullptr; p->duo_overflow = nullptr;
  }
  // segments the persistent solver (K7) takes: short enough for one warp
  uint32_t n_items = 0;
  for (uint32_t s = 0; s < S; ++s) if (p->h_duo_off[s + 1] - p->h_duo_off[s] <= kSolveMaxDuos) ++n_items;
  p->n_solve_items = n_items; p->solve_all = n_items == S;
  // ---- one staging block, one upload ----
  const size_t o_tseg = 0, o_first = o_tseg + T, o_mine = o_first + S + 1, o_moff = o_mine + T, o_wrb = o_moff + W + 1, o_woffA = o_wrb + W,
               o_woffB = o_woffA + W + 1, o_trb = o_woffB + W + 1, o_tdb = o_trb + T + 1, o_rot = o_tdb + T, o_soff = o_rot + T, o_doff = o_soff + S + 1,
               o_items = o_doff + S + 1, words = o_items + (p->solve_all ? 0 : n_items);
  if (ctx->stage_busy) { CK(cudaEventSynchronize(ctx->stage_ev)); ctx->stage_busy = false; }
  if (int rc = pinned_reserve(ctx, &ctx->h_stage, &ctx->stage_cap, words)) return rc;
  if (!ctx->stage_ev) CK(cudaEventCreateWithFlags(&ctx->stage_ev, cudaEventDisableTiming));
  uint32_t* h = ctx->h_stage;
  for (uint32_t t = 0; t < T; ++t) h[o_tseg + t] = sch.tiles[t].seg;
  auto put = [&](size_t off, const std::vector<uint32_t>& v) { if (!v.empty()) memcpy(h + off, v.data(), v.size() * sizeof(uint32_t)); };
  put(o_first, sch.first); put(o_mine, sch.mine); put(o_moff, sch.mine_off); put(o_wrb, sch.warp_rec_begin); put(o_woffA, sch.woffA); put(o_woffB, sch.woffB);
  put(o_trb, sch.tile_rec_begin); put(o_tdb, sch.tile_duo_begin); put(o_rot, sch.rec_of_tile); put(o_soff, p->h_seg_off); put(o_doff, p->h_duo_off);
  if (!p->solve_all) { uint32_t i = 0; for (uint32_t s = 0; s < S; ++s) if (p->h_duo_off[s + 1] - p->h_duo_off[s] <= kSolveMaxDuos) h[o_items + i++] = s; }
  CK(dev_alloc(&p->sched_blob, words));
  CK(cudaMemcpyAsync(p->sched_blob, h, words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->stage_ev, ctx->stream)); ctx->stage_busy = true;
  uint32_t* d = p->sched_blob;
  p->seg_first_tile = d + o_first; p->warp_off_full = d + o_woffA; p->warp_off = d + o_woffB; p->rec_of_tile = d + o_rot; p->seg_off = d + o_soff;
  p->seg_duo_off = d + o_doff; p->solve_items = p->solve_all ? nullptr : d + o_items;
  int nl = 0;
  CK(dev_alloc(&p->chunks, p->n_chunks)); CK(dev_alloc(&p->chunks_full, p->n_chunks_full));
  CK(launch_emit_chunks(d + o_tseg, d + o_first, d + o_doff, sch.tile_duos, d + o_mine, d + o_moff, d + o_wrb, d + o_woffA, d + o_woffB, W, p->chunks_full, p->chunks,
                        ctx->stream, &nl));
  lap("upload + chunks");
  // record-major duo table in schedule order (what K3 streams)
  if (!records_ready) {
    // compact records; the few duos that do not fit go to an overflow table sized by a first guess and, if that was too small, rebuilt
    uint32_t* d_novf = nullptr; uint32_t h_novf = 0, ovf_cap = std::max<uint32_t>(64u, p->n_duos / 512u);
    cudaError_t e = dev_alloc(&p->duo_recs, p->n_duos);
    if (e == cudaSuccess) e = dev_alloc(&p->duo_p0, p->n_duos);
    if (e == cudaSuccess) e = dev_alloc(&d_novf, 1);
    for (int attempt = 0; attempt < 2 && e == cudaSuccess; ++attempt) {
      e = dev_alloc(&p->duo_overflow, ovf_cap);
      if (e == cudaSuccess) e = cudaMemsetAsync(d_novf, 0, sizeof(uint32_t), ctx->stream);
      if (e == cudaSuccess) e = launch_build_duo_records(p->cells_m, p->cells_f, p->duos, p->n_duos, d + o_trb, d + o_tdb, T, p->duo_recs, p->duo_p0, p->duo_overflow,
                                                         ovf_cap, d_novf, ctx->stream, &nl);
      if (e == cudaSuccess) e = cudaMemcpyAsync(&h_novf, d_novf, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess || h_novf <= ovf_cap) break;
      dev_free(p->duo_overflow); p->duo_overflow = nullptr; ovf_cap = h_novf;
    }
    p->n_overflow = h_novf;
    dev_free(d_novf);
    lap("records");
    if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "finish_problem: record table", e);
  }
  ctx->launches += nl;
  // ---- work arrays of the evaluation / solver entry points: one allocation ----
  const size_t n_part = (size_t)T * kMaxAcc, n_dbl = n_part + (size_t)S * 4 + (size_t)S * RANDT_FUSED_STRIDE + S, n_cnt = (size_t)S + 2;
  CK(dev_alloc(&p->work_blob, n_dbl + (n_cnt + 1) / 2));
  p->partials = p->work_blob; p->d_poses = p->partials + n_part; p->d_out = p->d_poses + (size_t)S * 4; p->d_mu = p->d_out + (size_t)S * RANDT_FUSED_STRIDE;
  p->seg_counters = reinterpret_cast<uint32_t*>(p->d_mu + S); p->solve_counter = p->seg_counters + S;
  CK(cudaMemsetAsync(p->seg_counters, 0, n_cnt * sizeof(uint32_t), ctx->stream));
  lap("tail");
  return RANDT_OK;
}

DeviceProblem view(const randt_problem* p) {
  DeviceProblem d;
  d.cells_m = p->cells_m; d.cells_f = p->cells_f; d.pairs = p->pairs; d.duos = p->duos; d.duo_recs = p->duo_recs; d.duo_overflow = p->duo_overflow; d.duo_p0 = p->duo_p0; d.seg_off = p->seg_off; d.chunks = p->chunks; d.n_chunks = p->n_chunks; d.warp_off = p->warp_off; d.n_warps = p->n_warps;
  d.seg_first_tile = p->seg_first_tile; d.n_segments = p->S; d.n_pairs = p->P; d.partials = p->partials; d.seg_counters = p->seg_counters;
  d.seg_active = nullptr; d.plan_static = 1u; d.out_packed = 0u;
  return d;
}

int make_loss(randt_ctx* ctx, const randt_loss* l, LossParams* out) {
  LossParams lp; lp.kind = RANDT_LOSS_NONE; lp.a2 = 1.0; lp.alpha = 2.0; lp.weight = 1.0; lp.mu = 1.0; lp.fa = 0.0; lp.tf = 0.0;
  if (l) {
    if (l->kind < RANDT_LOSS_NONE || l->kind > RANDT_LOSS_WELSCH) return fail(ctx, RANDT_E_INVALID, "unknown loss kind");
    lp.kind = l->kind; lp.a2 = l->scale * l->scale; lp.alpha = l->alpha; lp.weight = l->weight; lp.mu = l->mu;
    if (l->kind != RANDT_LOSS_NONE && !(l->scale > 0.0 && l->mu > 0.0)) return fail(ctx, RANDT_E_INVALID, "loss scale and mu must be > 0");
    if (l->kind == RANDT_LOSS_BARRON && l->alpha == 0.0) { /* handled by the |alpha| <= 0.05 branch */ }
  }
  if (lp.kind == RANDT_LOSS_BARRON) {
    const double factor = fabs(lp.alpha - 2.0);
    lp.fa = factor / lp.alpha;     // inf for alpha == 0: that case takes the log branch and never reads it
    lp.tf = 2.0 / factor;
  }
  *out = lp;
  return RANDT_OK;
}

int check_variant(randt_ctx* ctx, int v) { return (v < 0 || v > 3) ? fail(ctx, RANDT_E_INVALID, "variant must be 0..3") : RANDT_OK; }
int np_of(int variant) { return variant <= 1 ? 4 : 3; }

}  // namespace

extern "C" {

int randt_version(void) { return 100; }

int randt_ctx_create(int device, void* stream, randt_ctx** out) {
  if (!out) return RANDT_E_INVALID;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return RANDT_E_CUDA;
  randt_ctx* ctx = new (std::nothrow) randt_ctx();
  if (!ctx) return RANDT_E_NOMEM;
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) {
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else { e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking); ctx->own_stream = true; }
    if (e == cudaSuccess) {
      ctx->sref = std::make_shared<StreamRef>();
      ctx->sref->device = device; ctx->sref->stream = ctx->stream; ctx->sref->own = ctx->own_stream;
    }
  }
  if (e == cudaSuccess) {
    // a pool of this context's own (the device's default pool and its release threshold belong to the host application)
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned; props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice; props.location.id = device;
    e = cudaMemPoolCreate(&ctx->sref->pool, &props);
    if (e == cudaSuccess) {
      unsigned long long keep = ~0ull;     // never trim: the per-scan calls reuse the same few buffers
      e = cudaMemPoolSetAttribute(ctx->sref->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_bad), sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_bad, 0, sizeof(unsigned long long), ctx->stream);
  if (e != cudaSuccess) { delete ctx; return RANDT_E_CUDA; }
  *out = ctx;
  return RANDT_OK;
}

void randt_ctx_destroy(randt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->sref && ctx->sref->copy) cudaStreamSynchronize(ctx->sref->copy);           // async evaluations still uploading / delivering
  if (ctx->sref && ctx->sref->copy_out) cudaStreamSynchronize(ctx->sref->copy_out);
  dev_free(ctx->d_bad);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_offs) cudaFreeHost(ctx->h_offs);
  if (ctx->stage_ev) cudaEventDestroy(ctx->stage_ev);
  if (ctx->offs_ev) cudaEventDestroy(ctx->offs_ev);
  delete ctx;      // the stream itself goes with the last object that was created on this context (StreamRef)
}

const char* randt_last_error(const randt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* randt_ctx_stream(const randt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int randt_ctx_sync(randt_ctx* ctx) {
  if (!ctx) return RANDT_E_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->sref && ctx->sref->copy_out) CK(cudaStreamSynchronize(ctx->sref->copy_out));
  return RANDT_OK;
}
uint64_t randt_ctx_launch_count(const randt_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* randt_host_alloc(size_t bytes) { void* p = nullptr; return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
void randt_host_free(void* p) { if (p) cudaFreeHost(p); }
void* randt_dev_alloc(size_t bytes) { void* p = nullptr; return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
void randt_dev_free(void* p) { if (p) cudaFree(p); }
int randt_memcpy_h2d(randt_ctx* ctx, void* dst, const void* src, size_t bytes) { if (!ctx) return RANDT_E_INVALID; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); return RANDT_OK; }
int randt_memcpy_d2h(randt_ctx* ctx, void* dst, const void* src, size_t bytes) { if (!ctx) return RANDT_E_INVALID; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); return RANDT_OK; }

// returns (and clears) the number of degenerate pairs (non-finite or negative d^T B^-1 d) K3 has seen on this context
int randt_ctx_take_bad_pairs(randt_ctx* ctx, uint64_t* count) {
  if (!ctx || !count) return RANDT_E_INVALID;
  unsigned long long h = 0;
  CK(cudaMemcpyAsync(&h, ctx->d_bad, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_bad, 0, sizeof(h), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *count = h;
  return RANDT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------------------------
int randt_voxelize(randt_ctx* ctx, const float* pts4, const uint32_t* scan_off, uint32_t n_scans, const randt_grid_params* gp,
                   int pts_on_device, randt_map** out) {
  if (!ctx || !out || !gp || !scan_off || (!pts4 && n_scans && scan_off[n_scans] > 0)) return fail(ctx, RANDT_E_INVALID, "randt_voxelize: null argument");
  *out = nullptr;
  if (gp->n_clusters <= 0 || gp->size_x <= 0 || gp->size_y <= 0 || !(gp->resolution > 0) || !(gp->max_range > 0))
    return fail(ctx, RANDT_E_INVALID, "randt_voxelize: bad grid parameters");
  for (uint32_t b = 0; b < n_scans; ++b)
    if (scan_off[b + 1] >= scan_off[b] && scan_off[b + 1] - scan_off[b] > 16384u)
      return fail(ctx, RANDT_E_CAPACITY, "randt_voxelize: more than 16384 points in one scan (a scan is sorted in shared memory; the shipped sensors deliver ~3-5 k filtered points)");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t B = n_scans;
  if (scan_off[0] != 0) return fail(ctx, RANDT_E_INVALID, "randt_voxelize: scan_off[0] must be 0");
  const uint32_t n_pts = scan_off[B];
  uint32_t max_pts = 0;
  for (uint32_t b = 0; b < B; ++b) { if (scan_off[b + 1] < scan_off[b]) return fail(ctx, RANDT_E_INVALID, "scan_off not monotone"); max_pts = std::max(max_pts, scan_off[b + 1] - scan_off[b]); }
  const uint32_t div = (uint32_t)std::max(gp->min_points, 0) + 1u;
  const uint32_t cell_cap = std::max<uint32_t>(1, max_pts / div);   // a kept cell has > min_points points
  randt_map* m = new (std::nothrow) randt_map();
  if (!m) return RANDT_E_NOMEM;
  m->device = ctx->device; m->sref = ctx->sref; m->gp = *gp; m->geom = make_geom(*gp); m->B = B;
  float4* d_pts = nullptr; bool own_pts = false;
  uint32_t *d_scan_off = nullptr, *d_cnt = nullptr, *d_npts_p = nullptr;
  int32_t* d_labels_p = nullptr; float4* d_cells_p = nullptr; int* d_status = nullptr;
  int rc = RANDT_OK;
  auto cleanup = [&]() {
    if (own_pts) dev_free(d_pts);
    dev_free(d_scan_off); dev_free(d_cnt); dev_free(d_npts_p); dev_free(d_labels_p);
    dev_free(d_cells_p);
  };
#define CKV(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, RANDT_E_CUDA, #call, e__); cleanup(); free_map(m); return rc; } } while (0)
  static const bool trace = getenv("RANDT_DEBUG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[randt] voxelize %-18s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_prev).count());
    t_prev = t;
  };
  if (pts_on_device) d_pts = const_cast<float4*>(reinterpret_cast<const float4*>(pts4));
  else { CKV(dev_alloc(&d_pts, n_pts)); own_pts = true; if (n_pts) CKV(cudaMemcpyAsync(d_pts, pts4, (size_t)n_pts * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream)); }
  lap("points h2d");
  CKV(dev_alloc(&d_scan_off, B + 1));
  CKV(cudaMemcpyAsync(d_scan_off, scan_off, (size_t)(B + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CKV(dev_alloc(&d_cnt, 2 * (size_t)B)); d_status = reinterpret_cast<int*>(d_cnt + B);   // counts | status codes
  CKV(dev_alloc(&d_cells_p, (size_t)B * cell_cap * 3)); CKV(dev_alloc(&d_npts_p, (size_t)B * cell_cap)); CKV(dev_alloc(&d_labels_p, (size_t)B * cell_cap));
  CKV(dev_alloc(&m->slot, (size_t)B * m->geom.n_slots));
  int nl = 0;
  unsigned short* d_bins = nullptr;
  if (voxelize_needs_bins_scratch(max_pts, cell_cap, *gp)) CKV(dev_alloc(&d_bins, n_pts));
  { const cudaError_t ev__ = launch_voxelize(d_pts, d_scan_off, B, max_pts, *gp, m->geom, cell_cap, d_cells_p, d_npts_p, d_labels_p, d_cnt, m->slot, d_bins, d_status, ctx->stream, &nl);
    dev_free(d_bins);      // stream-ordered: released after the kernel
    CKV(ev__); }
  lap("queued");
  // counts and status codes sit side by side and come back in one copy into the context's pinned block
  if (pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, 2 * (size_t)B + 2) != RANDT_OK) { cleanup(); free_map(m); return RANDT_E_NOMEM; }
  if (B) CKV(cudaMemcpyAsync(ctx->h_offs, d_cnt, 2 * (size_t)B * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CKV(cudaStreamSynchronize(ctx->stream));
  const uint32_t* h_cnt = ctx->h_offs; const int* h_status = reinterpret_cast<const int*>(ctx->h_offs + B);
  lap("counts back");
  for (uint32_t b = 0; b < B; ++b) {
    if (h_status[b] == VOX_SPAN || h_status[b] == VOX_CELL_CAP) {
      cleanup(); free_map(m);
      return fail(ctx, RANDT_E_CAPACITY, "randt_voxelize: label span / cell capacity exceeded (points far outside +-max_range?)");
    }
    if (h_status[b] == VOX_OUT_OF_MAP) {
      cleanup(); free_map(m);
      return fail(ctx, RANDT_E_INVALID, "randt_voxelize: a cell mean falls outside the map (the reference throws std::out_of_range here)");
    }
  }
  m->h_cell_off.assign(B + 1, 0);
  for (uint32_t b = 0; b < B; ++b) { m->h_cell_off[b + 1] = m->h_cell_off[b] + h_cnt[b]; m->max_per_map = std::max(m->max_per_map, h_cnt[b]); }
  m->n_cells = m->h_cell_off[B];
  CKV(dev_alloc(&m->cell_off, B + 1));
  CKV(cudaMemcpyAsync(m->cell_off, m->h_cell_off.data(), (size_t)(B + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (B == 1) {
    // a single scan's padded region already is its compact table: hand the buffers over (the per-scan call of a live stream)
    m->cells = d_cells_p; m->npts = d_npts_p; m->labels = d_labels_p;
    d_cells_p = nullptr; d_npts_p = nullptr; d_labels_p = nullptr;
  } else {
    CKV(dev_alloc(&m->cells, (size_t)m->n_cells * 3)); CKV(dev_alloc(&m->npts, m->n_cells)); CKV(dev_alloc(&m->labels, m->n_cells));
    CKV(launch_compact_cells(d_cells_p, d_npts_p, d_labels_p, m->cell_off, B, cell_cap, m->max_per_map, m->cells, m->npts, m->labels, ctx->stream, &nl));
    // (no wait: the padded buffers are released in stream order, and every later use of the map is queued on the same stream)
  }
#undef CKV
  ctx->launches += nl;
  cleanup();
  *out = m;
  return RANDT_OK;
}

// pcl::PointXYZI records (32 bytes each, cloud.points.data()) -> the path's float4 points on the device
int randt_points_from_pcl_xyzi(randt_ctx* ctx, const void* pcl_points, uint32_t n, int on_device, float* d_out4) {
  if (!ctx || (!pcl_points && n) || (!d_out4 && n)) return fail(ctx, RANDT_E_INVALID, "randt_points_from_pcl_xyzi: null argument");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  if (n == 0) return RANDT_OK;
  const void* d_in = pcl_points; void* d_tmp = nullptr;
  int nl = 0;
  cudaError_t e = cudaSuccess;
  if (!on_device) {
    unsigned char* t = nullptr;
    e = dev_alloc(&t, (size_t)n * 32);
    d_tmp = t; d_in = t;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tmp, pcl_points, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess) e = launch_pcl_xyzi_to_float4(d_in, n, reinterpret_cast<float4*>(d_out4), ctx->stream, &nl);
  dev_free(d_tmp);
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_points_from_pcl_xyzi", e);
  ctx->launches += nl;
  return RANDT_OK;
}

int randt_filter_scan(randt_ctx* ctx, const float* raw4, uint32_t n_az, uint32_t n_bins, const randt_filter_params* fp, int raw_on_device,
                      float* out4, int out_on_device, uint32_t cap, uint32_t* n_out) {
  if (!ctx || !fp || !n_out || (!raw4 && n_az && n_bins) || (!out4 && cap)) return fail(ctx, RANDT_E_INVALID, "randt_filter_scan: null argument");
  if ((unsigned long long)n_az * n_bins > 0x7fffffffull) return fail(ctx, RANDT_E_CAPACITY, "randt_filter_scan: scan too large");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  *n_out = 0;
  const size_t n = (size_t)n_az * n_bins;
  float4 *d_raw = nullptr, *d_out = nullptr; uint32_t *d_peak = nullptr, *d_n = nullptr; float* d_angle = nullptr; int* d_status = nullptr;
  bool own_raw = false, own_out = false;
  int nl = 0;
  auto cleanup = [&]() { if (own_raw) dev_free(d_raw); if (own_out) dev_free(d_out); dev_free(d_peak); dev_free(d_n); dev_free(d_angle); dev_free(d_status); };
  cudaError_t e = cudaSuccess;
  if (raw_on_device) d_raw = const_cast<float4*>(reinterpret_cast<const float4*>(raw4));
  else { own_raw = true; e = dev_alloc(&d_raw, n); if (e == cudaSuccess && n) e = cudaMemcpyAsync(d_raw, raw4, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream); }
  if (out_on_device) d_out = reinterpret_cast<float4*>(out4);
  else if (e == cudaSuccess) { own_out = true; e = dev_alloc(&d_out, cap); }
  if (e == cudaSuccess) e = dev_alloc(&d_peak, n_az);
  if (e == cudaSuccess) e = dev_alloc(&d_angle, n_az);
  if (e == cudaSuccess) e = dev_alloc(&d_n, 1);
  if (e == cudaSuccess) e = dev_alloc(&d_status, 1);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = launch_filter_scan(d_raw, n_az, n_bins, *fp, d_peak, d_angle, d_out, cap, d_n, d_status, ctx->stream, &nl);
  int h_status = 0; uint32_t h_n = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_n, d_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && !out_on_device && h_status == 0 && h_n)
    e = cudaMemcpy(out4, d_out, (size_t)std::min(h_n, cap) * sizeof(float4), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_filter_scan", e);
  ctx->launches += nl;
  *n_out = h_n;
  if (h_status == 1) return fail(ctx, RANDT_E_INVALID, "randt_filter_scan: the scan is not organised by azimuth (the reference's angle rule cuts it elsewhere)");
  if (h_status == 2) return fail(ctx, RANDT_E_CAPACITY, "randt_filter_scan: output capacity too small");
  return RANDT_OK;
}

int randt_filter_scans(randt_ctx* ctx, const float* raw4, uint32_t n_scans, uint32_t n_az, uint32_t n_bins, const randt_filter_params* fp,
                       int raw_on_device, float* out4, int out_on_device, uint32_t cap, uint32_t* scan_off) {
  if (!ctx || !fp || !scan_off || (!raw4 && n_scans && n_az && n_bins) || (!out4 && cap)) return fail(ctx, RANDT_E_INVALID, "randt_filter_scans: null argument");
  if ((unsigned long long)n_az * n_bins > 0x7fffffffull) return fail(ctx, RANDT_E_CAPACITY, "randt_filter_scans: scan too large");
  if ((unsigned long long)n_scans * n_az > 0x7fffffffull || n_scans > 65535u) return fail(ctx, RANDT_E_CAPACITY, "randt_filter_scans: too many scans per call");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  for (uint32_t i = 0; i <= n_scans; ++i) scan_off[i] = 0;
  if (n_scans == 0) return RANDT_OK;
  const size_t n = (size_t)n_scans * n_az * n_bins;
  float4 *d_raw = nullptr, *d_out = nullptr; uint32_t *d_peak = nullptr, *d_cnt = nullptr, *d_off = nullptr, *d_bs = nullptr; float* d_angle = nullptr;
  int* d_status = nullptr;
  bool own_raw = false, own_out = false;
  int nl = 0;
  auto cleanup = [&]() {
    if (own_raw) dev_free(d_raw); if (own_out) dev_free(d_out);
    dev_free(d_peak); dev_free(d_cnt); dev_free(d_off); dev_free(d_bs); dev_free(d_angle); dev_free(d_status);
  };
  cudaError_t e = cudaSuccess;
  if (raw_on_device) d_raw = const_cast<float4*>(reinterpret_cast<const float4*>(raw4));
  else { own_raw = true; e = dev_alloc(&d_raw, n); if (e == cudaSuccess && n) e = cudaMemcpyAsync(d_raw, raw4, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream); }
  if (out_on_device) d_out = reinterpret_cast<float4*>(out4);
  else if (e == cudaSuccess) { own_out = true; e = dev_alloc(&d_out, cap); }
  if (e == cudaSuccess) e = dev_alloc(&d_peak, (size_t)n_scans * n_az);
  if (e == cudaSuccess) e = dev_alloc(&d_angle, (size_t)n_scans * n_az);
  if (e == cudaSuccess) e = dev_alloc(&d_cnt, n_scans);
  if (e == cudaSuccess) e = dev_alloc(&d_off, (size_t)n_scans + 1);
  if (e == cudaSuccess) e = dev_alloc(&d_bs, n_scans / 1024 + 2);
  if (e == cudaSuccess) e = dev_alloc(&d_status, n_scans);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, (size_t)n_scans * sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = launch_filter_scans(d_raw, n_scans, n_az, n_bins, *fp, d_peak, d_angle, d_out, cap, d_cnt, d_off, d_bs, d_status, ctx->stream, &nl);
  std::vector<int> h_status(n_scans, 0);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_status.data(), d_status, (size_t)n_scans * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(scan_off, d_off, ((size_t)n_scans + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  bool bad_shape = false;
  for (uint32_t i = 0; i < n_scans; ++i) bad_shape = bad_shape || h_status[i] == 1;
  const bool over = scan_off[n_scans] > cap;
  if (e == cudaSuccess && !out_on_device && !bad_shape && !over && scan_off[n_scans])
    e = cudaMemcpy(out4, d_out, (size_t)scan_off[n_scans] * sizeof(float4), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_filter_scans", e);
  ctx->launches += nl;
  if (bad_shape) return fail(ctx, RANDT_E_INVALID, "randt_filter_scans: a scan is not organised by azimuth (the reference's angle rule cuts it elsewhere)");
  if (over) return fail(ctx, RANDT_E_CAPACITY, "randt_filter_scans: output capacity too small");
  return RANDT_OK;
}

int randt_map_upload(randt_ctx* ctx, const float* cells, const uint32_t* npts, const uint32_t* cell_off, uint32_t n_maps, const int32_t* slot,
                     const randt_grid_params* gp, randt_map** out) {
  if (!ctx || !out || !gp || !cell_off) return fail(ctx, RANDT_E_INVALID, "randt_map_upload: null argument");
  *out = nullptr;
  if (gp->size_x <= 0 || gp->size_y <= 0 || !(gp->resolution > 0)) return fail(ctx, RANDT_E_INVALID, "randt_map_upload: bad map geometry");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  randt_map* m = new (std::nothrow) randt_map();
  if (!m) return RANDT_E_NOMEM;
  m->device = ctx->device; m->sref = ctx->sref; m->gp = *gp; m->geom = make_geom(*gp); m->B = n_maps;
  m->h_cell_off.assign(cell_off, cell_off + n_maps + 1);
  for (uint32_t b = 0; b < n_maps; ++b) {
    if (cell_off[b + 1] < cell_off[b]) { free_map(m); return fail(ctx, RANDT_E_INVALID, "cell_off not monotone"); }
    m->max_per_map = std::max(m->max_per_map, cell_off[b + 1] - cell_off[b]);
  }
  m->n_cells = cell_off[n_maps];
  if (m->n_cells && !cells) { free_map(m); return fail(ctx, RANDT_E_INVALID, "randt_map_upload: cells == NULL"); }
  int rc = RANDT_OK; int nl = 0;
#define CKM(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, RANDT_E_CUDA, #call, e__); free_map(m); return rc; } } while (0)
  CKM(dev_alloc(&m->cells, (size_t)m->n_cells * 3)); CKM(dev_alloc(&m->npts, m->n_cells)); CKM(dev_alloc(&m->cell_off, n_maps + 1));
  CKM(dev_alloc(&m->slot, (size_t)n_maps * m->geom.n_slots));
  if (m->n_cells) CKM(cudaMemcpyAsync(m->cells, cells, (size_t)m->n_cells * 48, cudaMemcpyHostToDevice, ctx->stream));
  if (npts && m->n_cells) CKM(cudaMemcpyAsync(m->npts, npts, (size_t)m->n_cells * 4, cudaMemcpyHostToDevice, ctx->stream));
  else { CKM(cudaMemsetAsync(m->npts, 0, std::max<size_t>(1, m->n_cells) * 4, ctx->stream)); m->has_npts = m->n_cells == 0; }
  CKM(cudaMemcpyAsync(m->cell_off, cell_off, (size_t)(n_maps + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (slot) CKM(cudaMemcpyAsync(m->slot, slot, (size_t)n_maps * m->geom.n_slots * 4, cudaMemcpyHostToDevice, ctx->stream));
  else {
    CKM(cudaMemsetAsync(m->slot, 0xFF, std::max<size_t>(1, (size_t)n_maps * m->geom.n_slots) * 4, ctx->stream));
    CKM(launch_build_slots(m->cells, m->cell_off, n_maps, m->max_per_map, m->geom, m->slot, ctx->stream, &nl));
  }
  CKM(cudaStreamSynchronize(ctx->stream));
#undef CKM
  ctx->launches += nl;
  *out = m;
  return RANDT_OK;
}

int randt_map_info(const randt_map* m, uint32_t* n_maps, uint32_t* n_cells_total, uint32_t* n_slots) {
  if (!m) return RANDT_E_INVALID;
  if (n_maps) *n_maps = m->B;
  if (n_cells_total) *n_cells_total = m->n_cells;
  if (n_slots) *n_slots = m->geom.n_slots;
  return RANDT_OK;
}

int randt_map_download(randt_ctx* ctx, const randt_map* m, float* cells, uint32_t* npts, int32_t* labels, uint32_t* cell_off, int32_t* slot) {
  if (!ctx || !m) return fail(ctx, RANDT_E_INVALID, "randt_map_download: null argument");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  if (cells && m->n_cells) CK(cudaMemcpyAsync(cells, m->cells, (size_t)m->n_cells * 48, cudaMemcpyDeviceToHost, ctx->stream));
  if (npts && m->n_cells) CK(cudaMemcpyAsync(npts, m->npts, (size_t)m->n_cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (labels) {
    if (!m->labels) return fail(ctx, RANDT_E_INVALID, "randt_map_download: this map has no voxel labels");
    if (m->n_cells) CK(cudaMemcpyAsync(labels, m->labels, (size_t)m->n_cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (cell_off) memcpy(cell_off, m->h_cell_off.data(), m->h_cell_off.size() * 4);
  if (slot) CK(cudaMemcpyAsync(slot, m->slot, (size_t)m->B * m->geom.n_slots * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RANDT_OK;
}

namespace {
// in: host float32 [B][4] (cos, sin, tx, ty) when from_se2d == 0, host float64 [B][4] Sophus SE2d storage otherwise
int map_transform_impl(randt_ctx* ctx, randt_map* m, const void* in, int from_se2d) {
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const size_t in_bytes = (size_t)m->B * 4 * (from_se2d ? sizeof(double) : sizeof(float));
  unsigned char* d_in = nullptr; float4* d_aff = nullptr;
  CK(dev_alloc(&d_in, in_bytes));
  cudaError_t e = dev_alloc(&d_aff, (size_t)m->B * 4);
  int nl = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = launch_prepare_affine(d_in, from_se2d, m->B, d_aff, ctx->stream, &nl);
  if (e == cudaSuccess) e = launch_transform_cells(m->cells, m->cell_off, m->B, m->max_per_map, d_aff, ctx->stream, &nl);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(d_in); dev_free(d_aff);
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_map_transform", e);
  ctx->launches += nl;
  return RANDT_OK;
}
}  // namespace

int randt_map_transform(randt_ctx* ctx, randt_map* m, const float* trans) {
  if (!ctx || !m || !trans) return fail(ctx, RANDT_E_INVALID, "randt_map_transform: null argument");
  return map_transform_impl(ctx, m, trans, 0);
}
int randt_map_transform_se2d(randt_ctx* ctx, randt_map* m, const double* poses) {
  if (!ctx || !m || !poses) return fail(ctx, RANDT_E_INVALID, "randt_map_transform_se2d: null argument");
  return map_transform_impl(ctx, m, poses, 1);
}

int randt_map_merge(randt_ctx* ctx, randt_map* F, const randt_map* M) {
  if (!ctx || !F || !M) return fail(ctx, RANDT_E_INVALID, "randt_map_merge: null argument");
  if (F->B != M->B || !same_geom(F->gp, M->gp)) return fail(ctx, RANDT_E_INVALID, "randt_map_merge: batch size / geometry mismatch");
  if (!F->has_npts || !M->has_npts)
    return fail(ctx, RANDT_E_INVALID, "randt_map_merge: a map was uploaded without point counts (Cell::operator+= weights by n - 1, ndt_cell.h:133-142)");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t B = F->B;
  uint32_t cap = 1;
  for (uint32_t b = 0; b < B; ++b) cap = std::max(cap, (F->h_cell_off[b + 1] - F->h_cell_off[b]) + (M->h_cell_off[b + 1] - M->h_cell_off[b]));
  std::vector<uint32_t> h_ooff(B + 1);
  for (uint32_t b = 0; b <= B; ++b) h_ooff[b] = b * cap;
  float4 *o_cells = nullptr, *n_cells = nullptr; uint32_t *o_npts = nullptr, *o_cnt = nullptr, *d_ooff = nullptr, *n_npts = nullptr, *n_off = nullptr;
  int nl = 0; int rc = RANDT_OK;
  auto cleanup = [&]() { dev_free(o_cells); dev_free(o_npts); dev_free(o_cnt); dev_free(d_ooff); };
#define CKG(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, RANDT_E_CUDA, #call, e__); cleanup(); dev_free(n_cells); dev_free(n_npts); dev_free(n_off); return rc; } } while (0)
  CKG(dev_alloc(&o_cells, (size_t)B * cap * 3)); CKG(dev_alloc(&o_npts, (size_t)B * cap)); CKG(dev_alloc(&o_cnt, B)); CKG(dev_alloc(&d_ooff, B + 1));
  CKG(cudaMemcpyAsync(d_ooff, h_ooff.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CKG(launch_merge_maps(F->cells, F->npts, F->cell_off, F->slot, M->cells, M->npts, M->cell_off, B, F->geom, d_ooff, o_cells, o_npts, o_cnt, M->max_per_map, ctx->stream, &nl));
  if (pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, (size_t)B + 1) != RANDT_OK) { cleanup(); return RANDT_E_NOMEM; }
  const uint32_t* h_cnt = ctx->h_offs;
  if (B) CKG(cudaMemcpyAsync(ctx->h_offs, o_cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CKG(cudaStreamSynchronize(ctx->stream));
  std::vector<uint32_t> new_off(B + 1, 0); uint32_t max_per = 0;
  for (uint32_t b = 0; b < B; ++b) { new_off[b + 1] = new_off[b] + h_cnt[b]; max_per = std::max(max_per, h_cnt[b]); }
  CKG(dev_alloc(&n_off, B + 1));
  CKG(cudaMemcpyAsync(n_off, new_off.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (B == 1) {
    // a single map's padded region already is its compact table: hand the buffers over (the keyframe insertion of a live stream)
    n_cells = o_cells; n_npts = o_npts; o_cells = nullptr; o_npts = nullptr;
  } else {
    CKG(dev_alloc(&n_cells, (size_t)new_off[B] * 3)); CKG(dev_alloc(&n_npts, new_off[B]));
    CKG(launch_compact_cells(o_cells, o_npts, nullptr, n_off, B, cap, max_per, n_cells, n_npts, nullptr, ctx->stream, &nl));
    // (no wait: the padded buffers are released in stream order, and every later use of the map is queued on the same stream)
  }
#undef CKG
  cleanup();
  dev_free(F->cells); dev_free(F->npts); dev_free(F->cell_off); dev_free(F->labels);
  F->cells = n_cells; F->npts = n_npts; F->cell_off = n_off; F->labels = nullptr;
  F->h_cell_off = new_off; F->n_cells = new_off[B]; F->max_per_map = max_per;
  ctx->launches += nl;
  return RANDT_OK;
}

int randt_cs_divergence(randt_ctx* ctx, const randt_map* F, const randt_map* M, double* out) {
  if (!ctx || !F || !M || !out) return fail(ctx, RANDT_E_INVALID, "randt_cs_divergence: null argument");
  if (F->B != M->B) return fail(ctx, RANDT_E_INVALID, "randt_cs_divergence: fixed and moving batches differ in size");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t B = F->B;
  if (B == 0) return RANDT_OK;
  double *d_part = nullptr, *d_out = nullptr; uint32_t* d_tick = nullptr;
  int nl = 0;
  cudaError_t e = dev_alloc(&d_part, (size_t)B * cs_divergence_split() * 3);
  if (e == cudaSuccess) e = dev_alloc(&d_out, B);
  if (e == cudaSuccess) e = dev_alloc(&d_tick, B);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_tick, 0, (size_t)B * sizeof(uint32_t), ctx->stream);
  if (e == cudaSuccess) e = launch_cs_divergence(F->cells, F->cell_off, M->cells, M->cell_off, B, d_part, d_tick, d_out, ctx->stream, &nl);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(d_part); dev_free(d_out); dev_free(d_tick);
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_cs_divergence", e);
  ctx->launches += nl;
  return RANDT_OK;
}

void randt_map_destroy(randt_map* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  std::shared_ptr<StreamRef> sr = m->sref;      // keeps the stream alive until the frees are queued
  if (sr && sr->own) { StreamScope scope__(sr->stream, sr->pool); free_map(m); }
  else free_map(m);                             // borrowed stream (may be gone by now): cudaFree
}

// ---------------------------------------------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------------------------------------------
int randt_associate(randt_ctx* ctx, const randt_map* F, const randt_map* M, const double* pose0, int k, int metric, randt_problem** out) {
  if (!ctx || !F || !M || !pose0 || !out) return fail(ctx, RANDT_E_INVALID, "randt_associate: null argument");
  *out = nullptr;
  if (F->B != M->B) return fail(ctx, RANDT_E_INVALID, "randt_associate: fixed and moving batches differ in size");
  if (k < 1 || k > kMaxNeighbours) return fail(ctx, RANDT_E_INVALID, "randt_associate: k must be in 1..8");
  if (metric != RANDT_LOOKUP_MAHALANOBIS_INTENSITY && metric != RANDT_LOOKUP_EUCLID_XY) return fail(ctx, RANDT_E_INVALID, "randt_associate: bad metric");
  MapGeomDev geom = F->geom;
  geom.r_stop = static_cast<int>(F->gp.max_linf / F->gp.resolution);
  if (2 * (geom.r_stop > 0 ? geom.r_stop - 1 : 0) + 1 > geom.size_x)
    return fail(ctx, RANDT_E_CAPACITY, "randt_associate: search window wider than the map (duplicate window slots are not supported)");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t B = F->B, n_m = M->n_cells;
  randt_problem* p = new (std::nothrow) randt_problem();
  if (!p) return RANDT_E_NOMEM;
  p->device = ctx->device; p->sref = ctx->sref; p->S = B; p->n_m = n_m; p->n_f = F->n_cells;
  // one scan against one submap (the live stream's call): one fused launch and one read-back
  static const bool no_fused = getenv("RANDT_NO_FUSED_ASSOC") != nullptr;   // A/B switch for tests and profiling
  if (B == 1 && !no_fused && n_m <= kSingleMaxCells && F->n_cells <= kSingleMaxCells) {
    const uint32_t ovf_cap = 64;
    double* d_p0 = nullptr; uint32_t* d_tot = nullptr; int nl1 = 0;
    const size_t max_duos = (size_t)n_m * ((k + 1) / 2);
    cudaError_t e = dev_alloc(&d_p0, 4);
    if (e == cudaSuccess) e = dev_alloc(&d_tot, 3);
    if (e == cudaSuccess) e = dev_alloc(&p->pairs, (size_t)n_m * k);
    if (e == cudaSuccess) e = dev_alloc(&p->duos, max_duos);
    if (e == cudaSuccess) e = dev_alloc(&p->duo_recs, max_duos);
    if (e == cudaSuccess) e = dev_alloc(&p->duo_p0, max_duos);
    if (e == cudaSuccess) e = dev_alloc(&p->duo_overflow, ovf_cap);
    if (e == cudaSuccess) e = dev_alloc(&p->cells_m, (size_t)n_m * 3);
    if (e == cudaSuccess) e = dev_alloc(&p->cells_f, (size_t)F->n_cells * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_p0, pose0, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_tot, 0, 3 * sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess) e = launch_associate_single(F->cells, F->n_cells, F->slot, M->cells, n_m, geom, d_p0, k, metric, p->pairs, p->duos, p->duo_recs, p->duo_p0,
                                                      p->duo_overflow, ovf_cap, p->cells_m, p->cells_f, d_tot, nullptr, nullptr, 0.0, nullptr, ctx->stream, &nl1);
    if (e == cudaSuccess && pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, 4) != RANDT_OK) e = cudaErrorMemoryAllocation;
    uint32_t* h_tot = ctx->h_offs;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_tot, d_tot, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(d_p0); dev_free(d_tot);
    if (e != cudaSuccess) { int rc1 = fail(ctx, RANDT_E_CUDA, "randt_associate (single map)", e); free_problem(p); return rc1; }
    ctx->launches += nl1;
    p->P = h_tot[0]; p->n_duos = h_tot[1]; p->n_overflow = h_tot[2];
    p->h_seg_off = {0u, p->P}; p->h_duo_off = {0u, p->n_duos};
    bool ready = true;
    if (h_tot[2] > ovf_cap) {      // more escapes than the first guess holds: let the general record builder size the table
      ready = false;
      dev_free(p->duo_recs); dev_free(p->duo_p0); dev_free(p->duo_overflow);
      p->duo_recs = nullptr; p->duo_p0 = nullptr; p->duo_overflow = nullptr;
    }
    int rc1 = finish_problem(ctx, p, ready);
    if (rc1 != RANDT_OK) { free_problem(p); return rc1; }
    *out = p;
    return RANDT_OK;
  }
  static const bool trace = getenv("RANDT_DEBUG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[randt] associate %-18s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_prev).count());
    t_prev = t;
  };
  float4* d_pose = nullptr; double* d_pose0 = nullptr; uint32_t *d_nn = nullptr, *d_cnt = nullptr, *d_scan = nullptr, *d_bs = nullptr, *d_cnt2 = nullptr, *d_scan2 = nullptr, *d_offs = nullptr;
  int rc = RANDT_OK; int nl = 0;
  auto cleanup = [&]() { dev_free(d_pose); dev_free(d_pose0); dev_free(d_nn); dev_free(d_cnt); dev_free(d_scan); dev_free(d_bs); dev_free(d_cnt2); dev_free(d_scan2); dev_free(d_offs); };
#define CKA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, RANDT_E_CUDA, #call, e__); cleanup(); free_problem(p); return rc; } } while (0)
  // initial_guess.cast<float>() and the rotation its transformCell applies, once per map on the device (AffineRec)
  CKA(dev_alloc(&d_pose, (size_t)B * 4)); CKA(dev_alloc(&d_pose0, (size_t)B * 4));
  if (B) CKA(cudaMemcpyAsync(d_pose0, pose0, (size_t)B * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CKA(launch_prepare_affine(d_pose0, 1, B, d_pose, ctx->stream, &nl));
  CKA(dev_alloc(&d_nn, (size_t)n_m * k)); CKA(dev_alloc(&d_cnt, n_m)); CKA(dev_alloc(&d_scan, (size_t)n_m + 1)); CKA(dev_alloc(&d_bs, n_m / 1024 + 2));
  CKA(launch_associate(F->cells, F->cell_off, F->slot, M->cells, M->cell_off, B, n_m, M->max_per_map, geom, d_pose, k, metric, d_nn, d_cnt, ctx->stream, &nl));
  CKA(launch_exclusive_scan_u32(d_cnt, d_scan, n_m, d_bs, ctx->stream, &nl));
  // duos (K3's grouping): ceil(cnt / 2) per moving cell
  CKA(dev_alloc(&d_cnt2, n_m)); CKA(dev_alloc(&d_scan2, (size_t)n_m + 1));
  CKA(launch_duo_counts(d_cnt, n_m, d_cnt2, ctx->stream, &nl));
  CKA(launch_exclusive_scan_u32(d_cnt2, d_scan2, n_m, d_bs, ctx->stream, &nl));
  // The host only needs the B+1 per-map offsets of both scans: gather them on the device and start their copy into pinned memory
  // now; the compaction kernels and the snapshots are queued behind it, and the host waits for the offsets only (it then builds the
  // schedule while the device is still compacting).  The pair / duo tables are sized by their bounds (k and ceil(k/2) per moving
  // cell; nearly every cell finds its k neighbours), so no size has to come back before the compaction kernels run.
  CKA(dev_alloc(&d_offs, 2 * ((size_t)B + 1)));
  CKA(launch_gather_offsets(d_scan, d_scan2, M->cell_off, B + 1, d_offs, ctx->stream, &nl));
  if (pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, 2 * ((size_t)B + 1)) != RANDT_OK) { cleanup(); free_problem(p); return RANDT_E_NOMEM; }
  if (!ctx->offs_ev) CKA(cudaEventCreateWithFlags(&ctx->offs_ev, cudaEventDisableTiming));
  CKA(cudaMemcpyAsync(ctx->h_offs, d_offs, 2 * ((size_t)B + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CKA(cudaEventRecord(ctx->offs_ev, ctx->stream));
  CKA(dev_alloc(&p->pairs, (size_t)n_m * k));
  CKA(launch_compact_pairs(d_nn, d_cnt, d_scan, M->cell_off, F->cell_off, B, n_m, M->max_per_map, k, p->pairs, ctx->stream, &nl));
  CKA(dev_alloc(&p->duos, (size_t)n_m * ((k + 1) / 2)));
  CKA(launch_compact_duos(d_nn, d_cnt, d_scan, d_scan2, M->cell_off, F->cell_off, B, n_m, M->max_per_map, k, p->duos, ctx->stream, &nl));
  // snapshot the cell tables (the reference's functors copy their cells; maps may be merged/transformed afterwards)
  CKA(dev_alloc(&p->cells_m, (size_t)n_m * 3)); CKA(dev_alloc(&p->cells_f, (size_t)F->n_cells * 3));
  if (n_m) CKA(cudaMemcpyAsync(p->cells_m, M->cells, (size_t)n_m * 48, cudaMemcpyDeviceToDevice, ctx->stream));
  if (F->n_cells) CKA(cudaMemcpyAsync(p->cells_f, F->cells, (size_t)F->n_cells * 48, cudaMemcpyDeviceToDevice, ctx->stream));
  lap("queued");
  CKA(cudaEventSynchronize(ctx->offs_ev));
  lap("offsets back");
  p->h_seg_off.assign(ctx->h_offs, ctx->h_offs + B + 1);
  p->h_duo_off.assign(ctx->h_offs + B + 1, ctx->h_offs + 2 * ((size_t)B + 1));
  p->P = p->h_seg_off[B]; p->n_duos = p->h_duo_off[B];
#undef CKA
  cleanup();
  ctx->launches += nl;
  rc = finish_problem(ctx, p);
  if (rc != RANDT_OK) { free_problem(p); return rc; }
  *out = p;
  return RANDT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// problems
// ---------------------------------------------------------------------------------------------------------------
int randt_problem_create(randt_ctx* ctx, const float* cells_m, uint32_t n_m, const float* cells_f, uint32_t n_f, const uint32_t* pair_m,
                         const uint32_t* pair_f, uint32_t n_pairs, const uint32_t* seg_off, uint32_t n_segments, randt_problem** out) {
  if (!ctx || !out || !seg_off || (n_pairs && (!pair_m || !pair_f)) || (n_m && !cells_m) || (n_f && !cells_f))
    return fail(ctx, RANDT_E_INVALID, "randt_problem_create: null argument");
  *out = nullptr;
  if (seg_off[0] != 0 || seg_off[n_segments] != n_pairs) return fail(ctx, RANDT_E_INVALID, "randt_problem_create: seg_off must span [0, n_pairs]");
  for (uint32_t s = 0; s < n_segments; ++s) if (seg_off[s + 1] < seg_off[s]) return fail(ctx, RANDT_E_INVALID, "randt_problem_create: seg_off not monotone");
  for (uint32_t i = 0; i < n_pairs; ++i) if (pair_m[i] >= n_m || pair_f[i] >= n_f) return fail(ctx, RANDT_E_INVALID, "randt_problem_create: pair index out of range");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  randt_problem* p = new (std::nothrow) randt_problem();
  if (!p) return RANDT_E_NOMEM;
  p->device = ctx->device; p->sref = ctx->sref; p->S = n_segments; p->P = n_pairs; p->n_m = n_m; p->n_f = n_f;
  p->h_seg_off.assign(seg_off, seg_off + n_segments + 1);
  std::vector<uint2> h_pairs(n_pairs);
  for (uint32_t i = 0; i < n_pairs; ++i) h_pairs[i] = make_uint2(pair_m[i], pair_f[i]);
  int rc = RANDT_OK;
#define CKP(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, RANDT_E_CUDA, #call, e__); free_problem(p); return rc; } } while (0)
  // duos: consecutive pairs of one segment that share their moving cell, two by two
  std::vector<Duo> h_duos; h_duos.reserve(n_pairs / 2 + n_segments + 1);
  p->h_duo_off.assign(n_segments + 1, 0);
  for (uint32_t sg = 0; sg < n_segments; ++sg) {
    for (uint32_t i = seg_off[sg]; i < seg_off[sg + 1];) {
      Duo d; d.im = pair_m[i]; d.jf0 = pair_f[i]; d.jf1 = kNoCell; d.p0 = i;
      if (i + 1 < seg_off[sg + 1] && pair_m[i + 1] == pair_m[i]) { d.jf1 = pair_f[i + 1]; i += 2; } else { i += 1; }
      h_duos.push_back(d);
    }
    p->h_duo_off[sg + 1] = (uint32_t)h_duos.size();
  }
  p->n_duos = (uint32_t)h_duos.size();
  CKP(dev_alloc(&p->cells_m, (size_t)n_m * 3)); CKP(dev_alloc(&p->cells_f, (size_t)n_f * 3)); CKP(dev_alloc(&p->pairs, n_pairs));
  CKP(dev_alloc(&p->duos, p->n_duos));
  if (p->n_duos) CKP(cudaMemcpyAsync(p->duos, h_duos.data(), (size_t)p->n_duos * sizeof(Duo), cudaMemcpyHostToDevice, ctx->stream));
  if (n_m) CKP(cudaMemcpyAsync(p->cells_m, cells_m, (size_t)n_m * 48, cudaMemcpyHostToDevice, ctx->stream));
  if (n_f) CKP(cudaMemcpyAsync(p->cells_f, cells_f, (size_t)n_f * 48, cudaMemcpyHostToDevice, ctx->stream));
  if (n_pairs) CKP(cudaMemcpyAsync(p->pairs, h_pairs.data(), (size_t)n_pairs * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaStreamSynchronize(ctx->stream));
#undef CKP
  rc = finish_problem(ctx, p);
  if (rc != RANDT_OK) { free_problem(p); return rc; }
  *out = p;
  return RANDT_OK;
}

int randt_problem_concat(randt_ctx* ctx, const randt_problem* const* parts, uint32_t n_parts, const uint32_t* seg_of_part, uint32_t n_segments,
                         randt_problem** out) {
  if (!ctx || !out || (n_parts && (!parts || !seg_of_part))) return fail(ctx, RANDT_E_INVALID, "randt_problem_concat: null argument");
  *out = nullptr;
  uint64_t n_m = 0, n_f = 0, n_p = 0, n_d = 0;
  for (uint32_t i = 0; i < n_parts; ++i) {
    const randt_problem* q = parts[i];
    if (!q || q->S != 1 || q->device != ctx->device) return fail(ctx, RANDT_E_INVALID, "randt_problem_concat: parts must be single-segment problems of this device");
    if (seg_of_part[i] >= n_segments || (i && seg_of_part[i] < seg_of_part[i - 1])) return fail(ctx, RANDT_E_INVALID, "randt_problem_concat: seg_of_part must be non-decreasing and below n_segments");
    n_m += q->n_m; n_f += q->n_f; n_p += q->P; n_d += q->n_duos;
  }
  if (n_m > 0xffffffffull || n_f > 0xffffffffull || n_p > 0xffffffffull) return fail(ctx, RANDT_E_CAPACITY, "randt_problem_concat: joint tables exceed 32-bit indices");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  randt_problem* p = new (std::nothrow) randt_problem();
  if (!p) return RANDT_E_NOMEM;
  p->device = ctx->device; p->sref = ctx->sref; p->S = n_segments; p->P = (uint32_t)n_p; p->n_m = (uint32_t)n_m; p->n_f = (uint32_t)n_f; p->n_duos = (uint32_t)n_d;
  p->h_seg_off.assign((size_t)n_segments + 1, 0u); p->h_duo_off.assign((size_t)n_segments + 1, 0u);
  int rc = RANDT_OK, nl = 0;
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, RANDT_E_CUDA, #call, e__); free_problem(p); return rc; } } while (0)
  CKC(dev_alloc(&p->cells_m, (size_t)n_m * 3)); CKC(dev_alloc(&p->cells_f, (size_t)n_f * 3)); CKC(dev_alloc(&p->pairs, n_p)); CKC(dev_alloc(&p->duos, n_d));
  uint32_t mb = 0, fb = 0, pb = 0, db = 0;
  for (uint32_t i = 0; i < n_parts; ++i) {
    const randt_problem* q = parts[i];
    // the parts' tables were written on their creating context's stream; the parts of one caller share it (ordering is the caller's otherwise)
    if (q->n_m) CKC(cudaMemcpyAsync(p->cells_m + 3 * (size_t)mb, q->cells_m, (size_t)q->n_m * 48, cudaMemcpyDeviceToDevice, ctx->stream));
    if (q->n_f) CKC(cudaMemcpyAsync(p->cells_f + 3 * (size_t)fb, q->cells_f, (size_t)q->n_f * 48, cudaMemcpyDeviceToDevice, ctx->stream));
    CKC(launch_shift_part(q->pairs, q->P, q->duos, q->n_duos, mb, fb, pb, p->pairs + pb, p->duos + db, ctx->stream, &nl));
    mb += q->n_m; fb += q->n_f; pb += q->P; db += q->n_duos;
    for (uint32_t s = seg_of_part[i] + 1; s <= n_segments; ++s) { p->h_seg_off[s] = pb; p->h_duo_off[s] = db; }
  }
#undef CKC
  ctx->launches += nl;
  rc = finish_problem(ctx, p);
  if (rc != RANDT_OK) { free_problem(p); return rc; }
  *out = p;
  return RANDT_OK;
}

int randt_problem_info(const randt_problem* p, uint32_t* n_segments, uint32_t* n_pairs, uint32_t* n_m, uint32_t* n_f) {
  if (!p) return RANDT_E_INVALID;
  if (n_segments) *n_segments = p->S;
  if (n_pairs) *n_pairs = p->P;
  if (n_m) *n_m = p->n_m;
  if (n_f) *n_f = p->n_f;
  return RANDT_OK;
}

int randt_problem_schedule(randt_ctx* ctx, const randt_problem* p, uint32_t* counts, uint32_t* plan_a4, uint32_t* plan_b4, uint32_t cap_chunks, uint32_t* woff_a,
                           uint32_t* woff_b, uint32_t cap_warps, uint32_t* duo_off) {
  if (!ctx || !p || !counts) return fail(ctx, RANDT_E_INVALID, "randt_problem_schedule: null argument");
  counts[0] = p->n_warps; counts[1] = p->n_tiles; counts[2] = p->n_chunks_full; counts[3] = p->n_chunks; counts[4] = (uint32_t)kK3MaxWarps;
  if (((plan_a4 || plan_b4) && (p->n_chunks_full > cap_chunks || p->n_chunks > cap_chunks)) || ((woff_a || woff_b) && p->n_warps > cap_warps))
    return fail(ctx, RANDT_E_CAPACITY, "randt_problem_schedule: output arrays too small");
  CK(cudaSetDevice(ctx->device));
  static_assert(sizeof(ChunkDesc) == 16, "4 x uint32 records");
  if (plan_a4 && p->n_chunks_full) CK(cudaMemcpyAsync(plan_a4, p->chunks_full, (size_t)p->n_chunks_full * 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (plan_b4 && p->n_chunks) CK(cudaMemcpyAsync(plan_b4, p->chunks, (size_t)p->n_chunks * 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (woff_a) CK(cudaMemcpyAsync(woff_a, p->warp_off_full, ((size_t)p->n_warps + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (woff_b) CK(cudaMemcpyAsync(woff_b, p->warp_off, ((size_t)p->n_warps + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (duo_off) memcpy(duo_off, p->h_duo_off.data(), p->h_duo_off.size() * sizeof(uint32_t));
  return RANDT_OK;
}

int randt_problem_layout(const randt_problem* p, uint32_t* n_duos, uint32_t* record_bytes, uint32_t* n_overflow) {
  if (!p) return RANDT_E_INVALID;
  if (n_duos) *n_duos = p->n_duos;
  if (record_bytes) *record_bytes = (uint32_t)sizeof(DuoRec);
  if (n_overflow) *n_overflow = p->n_overflow;
  return RANDT_OK;
}

int randt_problem_download(randt_ctx* ctx, const randt_problem* p, uint32_t* pair_m, uint32_t* pair_f, uint32_t* seg_off) {
  if (!ctx || !p) return fail(ctx, RANDT_E_INVALID, "randt_problem_download: null argument");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  if ((pair_m || pair_f) && p->P) {
    std::vector<uint2> h(p->P);
    CK(cudaMemcpyAsync(h.data(), p->pairs, (size_t)p->P * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < p->P; ++i) { if (pair_m) pair_m[i] = h[i].x; if (pair_f) pair_f[i] = h[i].y; }
  }
  if (seg_off) memcpy(seg_off, p->h_seg_off.data(), p->h_seg_off.size() * 4);
  return RANDT_OK;
}

// download the snapshotted cell tables (tests)
int randt_problem_download_cells(randt_ctx* ctx, const randt_problem* p, float* cells_m, float* cells_f) {
  if (!ctx || !p) return fail(ctx, RANDT_E_INVALID, "randt_problem_download_cells: null argument");
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  if (cells_m && p->n_m) CK(cudaMemcpyAsync(cells_m, p->cells_m, (size_t)p->n_m * 48, cudaMemcpyDeviceToHost, ctx->stream));
  if (cells_f && p->n_f) CK(cudaMemcpyAsync(cells_f, p->cells_f, (size_t)p->n_f * 48, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RANDT_OK;
}

void randt_problem_destroy(randt_problem* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  std::shared_ptr<StreamRef> sr = p->sref;
  if (sr && sr->own) { StreamScope scope__(sr->stream, sr->pool); free_problem(p); }
  else free_problem(p);
}

// ---------------------------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------------------------
int randt_eval_emit_dev(randt_ctx* ctx, const randt_problem* p, int variant, const double* d_poses, double* d_r, double* d_J) {
  if (!ctx || !p || !d_poses || !d_r) return fail(ctx, RANDT_E_INVALID, "randt_eval_emit_dev: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  int nl = 0;
  CK(launch_eval_emit(view(p), variant, d_poses, d_r, d_J, ctx->d_bad, ctx->stream, &nl));
  ctx->launches += nl;
  return RANDT_OK;
}

int randt_eval_emit(randt_ctx* ctx, const randt_problem* cp, int variant, const double* poses, double* r, double* J) {
  if (!ctx || !cp || !poses || !r) return fail(ctx, RANDT_E_INVALID, "randt_eval_emit: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  randt_problem* p = const_cast<randt_problem*>(cp);
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const int np = np_of(variant);
  if (!p->d_r) CK(dev_alloc(&p->d_r, p->P));
  if (J && !p->d_J) CK(dev_alloc(&p->d_J, (size_t)p->P * 4));
  CK(cudaMemcpyAsync(p->d_poses, poses, (size_t)p->S * np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int rc = randt_eval_emit_dev(ctx, p, variant, p->d_poses, p->d_r, J ? p->d_J : nullptr);
  if (rc) return rc;
  // A single problem's residuals and Jacobian rows (the ceres::CostFunction::Evaluate shape: a few KB) come back through the
  // context's pinned block: two copies into pageable memory would each be a blocking staged transfer (~10 us apiece on this host).
  const size_t n_r = p->P, n_J = J ? (size_t)p->P * np : 0;
  if ((n_r + n_J) * sizeof(double) <= ((size_t)1 << 20) && n_r) {
    if (int rc2 = pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, 2 * (n_r + n_J))) return rc2;
    double* h = reinterpret_cast<double*>(ctx->h_offs);
    CK(cudaMemcpyAsync(h, p->d_r, n_r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_J) CK(cudaMemcpyAsync(h + n_r, p->d_J, n_J * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(r, h, n_r * sizeof(double));
    if (n_J) memcpy(J, h + n_r, n_J * sizeof(double));
    return RANDT_OK;
  }
  if (p->P) CK(cudaMemcpyAsync(r, p->d_r, (size_t)p->P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (J && p->P) CK(cudaMemcpyAsync(J, p->d_J, (size_t)p->P * np * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RANDT_OK;
}

namespace {
int eval_fused_dev_impl(randt_ctx* ctx, const randt_problem* p, int variant, const double* d_poses, const randt_loss* loss,
                        const double* d_mu_per_seg, int want_jac, double* d_out, int packed);
inline size_t out_stride(int packed) { return packed == 1 ? RANDT_PACKED_STRIDE : (packed == 2 ? RANDT_CORE_STRIDE : (packed == 3 ? RANDT_BASIS_STRIDE : RANDT_FUSED_STRIDE)); }
}
int randt_eval_fused_dev(randt_ctx* ctx, const randt_problem* p, int variant, const double* d_poses, const randt_loss* loss,
                         const double* d_mu_per_seg, int want_jac, double* d_out) {
  return eval_fused_dev_impl(ctx, p, variant, d_poses, loss, d_mu_per_seg, want_jac, d_out, 0);
}
namespace {
int eval_fused_dev_impl(randt_ctx* ctx, const randt_problem* p, int variant, const double* d_poses, const randt_loss* loss,
                        const double* d_mu_per_seg, int want_jac, double* d_out, int packed) {
  if (!ctx || !p || !d_poses || !d_out) return fail(ctx, RANDT_E_INVALID, "randt_eval_fused_dev: null argument");
  if (packed < 0 || packed > 3) return fail(ctx, RANDT_E_INVALID, "randt_eval_fused: packed must be 0 .. 3");
  if (packed == 3 && (variant == RANDT_VAR_SE2_XY || !want_jac))
    return fail(ctx, RANDT_E_INVALID, "randt_eval_fused: the basis record (packed == 3) needs a three-dimensional basis and a Jacobian evaluation");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  LossParams lp;
  if (int rc = make_loss(ctx, loss, &lp)) return rc;
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  // segments without pairs produce no tile: clear their records up front
  if (p->has_empty_segment) CK(cudaMemsetAsync(d_out, 0, (size_t)p->S * out_stride(packed) * sizeof(double), ctx->stream));
  int nl = 0;
  DeviceProblem v = view(p);
  v.chunks = p->chunks_full; v.n_chunks = p->n_chunks_full; v.warp_off = p->warp_off_full;   // every segment is evaluated: packed schedule
  v.out_packed = (uint32_t)packed;
  CK(launch_eval_fused(v, variant, d_poses, lp, d_mu_per_seg, want_jac != 0, d_out, ctx->d_bad, ctx->stream, &nl));
  ctx->launches += nl;
  return RANDT_OK;
}
}  // namespace

int randt_eval_fused(randt_ctx* ctx, const randt_problem* cp, int variant, const double* poses, const randt_loss* loss,
                     const double* mu_per_seg, int want_jac, double* out) {
  if (!ctx || !cp || !poses || !out) return fail(ctx, RANDT_E_INVALID, "randt_eval_fused: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  randt_problem* p = const_cast<randt_problem*>(cp);
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const int np = np_of(variant);
  CK(cudaMemcpyAsync(p->d_poses, poses, (size_t)p->S * np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (mu_per_seg) CK(cudaMemcpyAsync(p->d_mu, mu_per_seg, (size_t)p->S * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // When the caller's result buffer is pinned (mapped) host memory, K3 stores its 192-byte records straight into it: the writes
  // travel over PCIe while the rest of the batch is still being evaluated, instead of a 192 S byte copy after the kernel.
  double* d_out = p->d_out;
  bool direct = false;
  if (!p->has_empty_segment) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) { d_out = static_cast<double*>(at.devicePointer); direct = true; }
    else cudaGetLastError();
  }
  int rc = randt_eval_fused_dev(ctx, p, variant, p->d_poses, loss, mu_per_seg ? p->d_mu : nullptr, want_jac, d_out);
  if (rc) return rc;
  const size_t n_out = (size_t)p->S * RANDT_FUSED_STRIDE;
  if (p->S && !direct && n_out * sizeof(double) <= ((size_t)1 << 20)) {      // small result into pageable memory: through the pinned block
    if (int rc2 = pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, 2 * n_out)) return rc2;
    CK(cudaMemcpyAsync(ctx->h_offs, p->d_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->h_offs, n_out * sizeof(double));
    return RANDT_OK;
  }
  if (p->S && !direct) CK(cudaMemcpyAsync(out, p->d_out, (size_t)p->S * RANDT_FUSED_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RANDT_OK;
}

// Pipelined form of randt_eval_fused for callers that evaluate one pose set after another (BnB levels, pose-grid scoring): the call
// only enqueues.  The poses travel on a copy stream into one of two device slots, so that the upload of call i+1 overlaps the
// kernel of call i, and the records of call i leave on a third stream while the kernel of call i+1 runs.  randt_ctx_sync() orders the
// host with all results; short of that, the records of call i are in host memory once call i+2 has left the context's stream.
int randt_eval_fused_async(randt_ctx* ctx, const randt_problem* cp, int variant, const double* poses, const randt_loss* loss,
                           const double* mu_per_seg, int want_jac, int packed, double* out) {
  if (!ctx || !cp || !poses || !out) return fail(ctx, RANDT_E_INVALID, "randt_eval_fused_async: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  randt_problem* p = const_cast<randt_problem*>(cp);
  CK(cudaSetDevice(ctx->device));
  StreamRef& r = *ctx->sref;
  // a pageable buffer would turn the copies synchronous: insist on pinned memory (checked once per buffer: callers cycle through a few;
  // a stale entry after the caller freed and reallocated the address only costs that synchronous copy)
  auto is_pinned = [&r](const void* h) -> bool {
    for (const void* k : r.known_pinned) if (k == h) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost) { r.known_pinned[r.known_n++ & 7u] = h; return true; }
    cudaGetLastError();
    return false;
  };
  if (!is_pinned(out) || !is_pinned(poses) || (mu_per_seg && !is_pinned(mu_per_seg)))
    return fail(ctx, RANDT_E_INVALID, "randt_eval_fused_async: poses, mu_per_seg and out must be pinned host memory (randt_host_alloc)");
  if (p->S == 0) {      // nothing to do, but the call still has a ticket (delivered as soon as everything before it is)
    if (r.copy_out) CK(cudaEventRecord(r.ev_call[(r.ring_n + 1) & 7u], r.copy_out));
    r.ring_n++;
    return RANDT_OK;
  }
  if (!r.copy) {
    CK(cudaStreamCreateWithFlags(&r.copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&r.copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&r.ev_in[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&r.ev_done[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&r.ev_d2h[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&r.ev_call[i], cudaEventDisableTiming));
  }
  const int np = np_of(variant);
  const uint64_t ticket = r.ring_n + 1;      // committed only once every step below has been enqueued: a failed call has no ticket
  const int slot = (int)(r.ring_n & 1);
  const size_t n_pose = (size_t)p->S * np, need = n_pose + (mu_per_seg ? p->S : 0);
  // the slot is free once the kernel that read it two calls ago has finished
  CK(cudaStreamWaitEvent(r.copy, r.ev_done[slot], 0));
  if (r.ring_cap[slot] < need) {
    if (r.ring[slot]) CK(cudaFreeAsync(r.ring[slot], r.copy));
    r.ring[slot] = nullptr; r.ring_cap[slot] = 0;
    CK(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&r.ring[slot]), need * sizeof(double), r.pool, r.copy));
    r.ring_cap[slot] = need;
  }
  double* d_in = r.ring[slot];
  CK(cudaMemcpyAsync(d_in, poses, n_pose * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  if (mu_per_seg) CK(cudaMemcpyAsync(d_in + n_pose, mu_per_seg, (size_t)p->S * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  CK(cudaEventRecord(r.ev_in[slot], r.copy));
  CK(cudaStreamWaitEvent(ctx->stream, r.ev_in[slot], 0));
  // Records go to a device slot and leave on a third stream: a DMA copy moves the 192 S bytes at the full PCIe rate while the next
  // call's kernel runs (stores from the kernel straight into mapped host memory, as the blocking call does, reach ~3/4 of that rate).
  if (packed < 0 || packed > 3) return fail(ctx, RANDT_E_INVALID, "randt_eval_fused_async: packed must be 0 .. 3");
  const size_t n_out = (size_t)p->S * out_stride(packed);   // K3 writes the layout itself
  CK(cudaStreamWaitEvent(ctx->stream, r.ev_d2h[slot], 0));     // the copy-out of two calls ago has drained this slot
  if (r.oring_cap[slot] < n_out) {
    if (r.oring[slot]) CK(cudaFreeAsync(r.oring[slot], ctx->stream));
    r.oring[slot] = nullptr; r.oring_cap[slot] = 0;
    CK(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&r.oring[slot]), n_out * sizeof(double), r.pool, ctx->stream));
    r.oring_cap[slot] = n_out;
  }
  double* d_out = r.oring[slot];
  int rc = eval_fused_dev_impl(ctx, p, variant, d_in, loss, mu_per_seg ? d_in + n_pose : nullptr, want_jac, d_out, packed);
  if (rc) return rc;
  const double* d_ship = d_out; const size_t n_ship = n_out;
  CK(cudaEventRecord(r.ev_done[slot], ctx->stream));
  CK(cudaStreamWaitEvent(r.copy_out, r.ev_done[slot], 0));
  CK(cudaMemcpyAsync(out, d_ship, n_ship * sizeof(double), cudaMemcpyDeviceToHost, r.copy_out));
  CK(cudaEventRecord(r.ev_d2h[slot], r.copy_out));             // randt_ctx_sync() waits for the copy-out stream as well
  CK(cudaEventRecord(r.ev_call[ticket & 7u], r.copy_out));     // delivery of this call (randt_ctx_wait_async)
  r.ring_n = ticket;
  return RANDT_OK;
}

uint64_t randt_ctx_async_count(const randt_ctx* ctx) { return (ctx && ctx->sref) ? ctx->sref->ring_n : 0; }

int randt_ctx_wait_async(randt_ctx* ctx, uint64_t ticket) {
  if (!ctx) return RANDT_E_INVALID;
  StreamRef& r = *ctx->sref;
  if (ticket == 0 || ticket > r.ring_n) return fail(ctx, RANDT_E_INVALID, "randt_ctx_wait_async: no such call");
  if (!r.copy_out) return RANDT_OK;
  // the event belongs to call `ticket`, or — if the caller ran more than eight calls ahead — to a later call, which completes after it
  CK(cudaEventSynchronize(r.ev_call[ticket & 7u]));
  return RANDT_OK;
}

// K8: all pairs of (moving cell, fixed cell) of every map pair of the batch, within `window` metres (L-infinity, <= 0: unbounded) of
// the transformed moving mean.  d_poses [B][np], d_out [B][RANDT_FUSED_STRIDE] on the device.
int randt_eval_allpairs_dev(randt_ctx* ctx, const randt_map* F, const randt_map* M, int variant, const double* d_poses, const randt_loss* loss,
                            double window, double* d_out) {
  if (!ctx || !F || !M || !d_poses || !d_out) return fail(ctx, RANDT_E_INVALID, "randt_eval_allpairs: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  if (F->B != M->B) return fail(ctx, RANDT_E_INVALID, "randt_eval_allpairs: fixed and moving batches differ in size");
  LossParams lp;
  if (int rc = make_loss(ctx, loss, &lp)) return rc;
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t B = F->B;
  if (B == 0) return RANDT_OK;
  uint32_t tm, tf;
  allpairs_tiles(M->max_per_map, F->max_per_map, &tm, &tf);
  double* d_part = nullptr; uint32_t* d_tick = nullptr;
  int nl = 0;
  cudaError_t e = dev_alloc(&d_part, (size_t)B * tm * tf * kMaxAcc);
  if (e == cudaSuccess) e = dev_alloc(&d_tick, B);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_tick, 0, (size_t)B * sizeof(uint32_t), ctx->stream);
  if (e == cudaSuccess) e = launch_allpairs(F->cells, F->cell_off, F->max_per_map, M->cells, M->cell_off, M->max_per_map, B, variant, d_poses, lp, window,
                                            d_part, d_tick, d_out, ctx->d_bad, ctx->stream, &nl);
  dev_free(d_part); dev_free(d_tick);
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_eval_allpairs", e);
  ctx->launches += nl;
  return RANDT_OK;
}

int randt_eval_allpairs(randt_ctx* ctx, const randt_map* F, const randt_map* M, int variant, const double* poses, const randt_loss* loss,
                        double window, double* out) {
  if (!ctx || !F || !M || !poses || !out) return fail(ctx, RANDT_E_INVALID, "randt_eval_allpairs: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t B = F->B;
  if (B == 0) return RANDT_OK;
  const int np = np_of(variant);
  double *d_p = nullptr, *d_o = nullptr;
  CK(dev_alloc(&d_p, (size_t)B * np));
  cudaError_t e = dev_alloc(&d_o, (size_t)B * RANDT_FUSED_STRIDE);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_p, poses, (size_t)B * np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  int rc = RANDT_OK;
  if (e == cudaSuccess) rc = randt_eval_allpairs_dev(ctx, F, M, variant, d_p, loss, window, d_o);
  if (e == cudaSuccess && rc == RANDT_OK) e = cudaMemcpyAsync(out, d_o, (size_t)B * RANDT_FUSED_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && rc == RANDT_OK) e = cudaStreamSynchronize(ctx->stream);
  dev_free(d_p); dev_free(d_o);
  if (rc != RANDT_OK) return rc;
  if (e != cudaSuccess) return fail(ctx, RANDT_E_CUDA, "randt_eval_allpairs", e);
  return RANDT_OK;
}

int randt_sweep_costs(randt_ctx* ctx, const randt_problem* cp, uint32_t seg, int variant, const double* poses, uint32_t n_poses,
                      const randt_loss* loss, double* cost) {
  if (!ctx || !cp || !poses || !cost) return fail(ctx, RANDT_E_INVALID, "randt_sweep_costs: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  randt_problem* p = const_cast<randt_problem*>(cp);
  if (seg >= p->S) return fail(ctx, RANDT_E_INVALID, "randt_sweep_costs: segment out of range");
  LossParams lp;
  if (int rc = make_loss(ctx, loss, &lp)) return rc;
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const int np = np_of(variant);
  const size_t need = (size_t)n_poses * (np + 1);
  if (p->sweep_cap < need) { dev_free(p->d_sweep); p->d_sweep = nullptr; p->sweep_cap = 0; CK(dev_alloc(&p->d_sweep, need)); p->sweep_cap = need; }
  double* d_p = p->d_sweep; double* d_c = p->d_sweep + (size_t)n_poses * np;
  if (n_poses) CK(cudaMemcpyAsync(d_p, poses, (size_t)n_poses * np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int nl = 0;
  CK(launch_sweep_costs(view(p), p->h_seg_off[seg], p->h_seg_off[seg + 1], variant, d_p, n_poses, lp, d_c, ctx->stream, &nl));
  ctx->launches += nl;
  if (n_poses) CK(cudaMemcpyAsync(cost, d_c, (size_t)n_poses * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RANDT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K4: batched registration
// ---------------------------------------------------------------------------------------------------------------
void randt_solver_options_default(randt_solver_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->max_num_iterations = 200; o->use_manifold = 0; o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->min_relative_decrease = 1e-3;
  o->gnc_loss_scale = 1.0; o->gnc_divisor = 1.1; o->gnc_max_steps = 2; o->poll_interval = 0;
}

namespace {
int register_batch_impl(randt_ctx* ctx, const randt_problem* cp, int variant, double* d_poses, const randt_loss* loss, const double* d_weight_per_seg,
                        const randt_solver_options* opt, double* d_result);
}
int randt_register_batch_dev(randt_ctx* ctx, const randt_problem* cp, int variant, double* d_poses, const randt_loss* loss,
                             const randt_solver_options* opt, double* d_result) {
  return register_batch_impl(ctx, cp, variant, d_poses, loss, nullptr, opt, d_result);
}
namespace {
int register_batch_impl(randt_ctx* ctx, const randt_problem* cp, int variant, double* d_poses, const randt_loss* loss, const double* d_weight_per_seg,
                        const randt_solver_options* opt, double* d_result) {
  if (!ctx || !cp || !d_poses || !opt || !d_result) return fail(ctx, RANDT_E_INVALID, "randt_register_batch: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  if (!(opt->gnc_divisor > 1.0) || !(opt->gnc_loss_scale > 0.0) || opt->gnc_max_steps < 1 || opt->max_num_iterations < 0 ||
      opt->max_num_consecutive_invalid_steps < 1 || !(opt->initial_trust_region_radius > 0.0))
    return fail(ctx, RANDT_E_INVALID, "randt_register_batch: bad solver options (gnc_divisor must be > 1, scales > 0, steps >= 1)");
  randt_loss l0; l0.kind = RANDT_LOSS_NONE; l0.scale = 1.0; l0.alpha = 2.0; l0.mu = 1.0; l0.weight = 1.0;
  if (loss) { l0 = *loss; l0.mu = 1.0; }
  LossParams lp;
  if (int rc = make_loss(ctx, &l0, &lp)) return rc;
  randt_problem* p = const_cast<randt_problem*>(cp);
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t S = p->S;
  const int np = np_of(variant);
  if (S == 0) return RANDT_OK;
  int nl = 0;
  // Registrations one warp can take (<= kSolveMaxDuos duos) are solved start to finish by the persistent kernel K7 in ONE launch; the
  // rest (and everything, when poll_interval < 0) goes through the stepwise loop: one K3 fused launch + one K4 launch per LM iteration.
  const bool stepwise_only = opt->poll_interval < 0;
  const uint32_t n_k7 = stepwise_only ? 0u : p->n_solve_items;
  if (d_weight_per_seg && n_k7 != S)
    return fail(ctx, RANDT_E_INVALID, "randt_register_batch_weighted: per-registration weights need the persistent solver (registrations of <= 1024 pairs, poll_interval >= 0)");
  if (n_k7) {
    SolveLayout L;
    L.seg_duo_off = p->seg_duo_off; L.tile_rec_begin = p->rec_of_tile; L.tile_duos = p->tile_duos;
    L.items = p->solve_all ? nullptr : p->solve_items; L.n_items = n_k7; L.next_item = p->solve_counter;
    CK(cudaMemsetAsync(p->solve_counter, 0, sizeof(uint32_t), ctx->stream));
    CK(launch_solve_persistent(view(p), L, variant, opt->use_manifold, lp, d_weight_per_seg, *opt, d_poses, d_poses, d_result, ctx->d_bad, ctx->stream, &nl));
    if (n_k7 == S) { ctx->launches += nl; return RANDT_OK; }
  }
  if (!p->lm_state) {
    CK(dev_alloc(&p->lm_state, S)); CK(dev_alloc(&p->lm_eval_pose, (size_t)S * 4)); CK(dev_alloc(&p->lm_mu, S));
    CK(dev_alloc(&p->lm_rec, (size_t)S * RANDT_FUSED_STRIDE)); CK(dev_alloc(&p->lm_active, S)); CK(dev_alloc(&p->lm_n_active, 1));
    CK(dev_alloc(&p->lm_chunks, p->n_chunks)); CK(dev_alloc(&p->lm_flags, p->n_chunks)); CK(dev_alloc(&p->lm_scan, (size_t)p->n_chunks + 1));
    CK(dev_alloc(&p->lm_bs, p->n_chunks / 1024 + 2)); CK(dev_alloc(&p->lm_warp_off, (size_t)p->n_warps + 1));
  }
  StreamRef& sr = *ctx->sref;
  if (!sr.h_n_active) {
    CK(cudaHostAlloc(reinterpret_cast<void**>(&sr.h_n_active), 2 * sizeof(uint32_t), cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&sr.ev[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&sr.ev[1], cudaEventDisableTiming));
  }
  CK(launch_lm_init(S, np, d_poses, p->lm_state, p->lm_eval_pose, p->lm_mu, p->lm_active, p->lm_rec, p->lm_n_active, ctx->stream, &nl));
  uint32_t planned_for = S;     // active segments when the schedule in use was made
  if (n_k7) {
    // only the long registrations are left: everything K7 solved starts out inactive
    std::vector<uint32_t> mask(S, 0u);
    uint32_t n_left = 0;
    for (uint32_t s = 0; s < S; ++s) if (p->h_duo_off[s + 1] - p->h_duo_off[s] > kSolveMaxDuos) { mask[s] = 1u; ++n_left; }
    CK(cudaMemcpyAsync(p->lm_active, mask.data(), (size_t)S * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(p->lm_n_active, &n_left, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    planned_for = S;
  }
  DeviceProblem v = view(p);
  v.seg_active = p->lm_active;
  const int poll = opt->poll_interval > 0 ? opt->poll_interval : 4;
  // every solve needs at most max_num_iterations candidate evaluations + its start evaluation; one more launch seeds the GNC
  const long long cap = 2 + (long long)(opt->gnc_max_steps + 1) * ((long long)opt->max_num_iterations + 2);
  // Iterations are enqueued in groups of `poll`; the active counter of group g is read back while group g + 1 already runs, so the
  // GPU never waits for the host (the price: up to one group of no-op launches after the last segment has finished).
  bool done = false;
  const auto t_solver0 = std::chrono::steady_clock::now();
  const long long n_groups = (cap + poll - 1) / poll;
  for (long long g = 0; g < n_groups && !done; ++g) {
    for (int i = 0; i < poll; ++i) {
      CK(launch_eval_fused(v, variant, p->lm_eval_pose, lp, p->lm_mu, true, p->lm_rec, ctx->d_bad, ctx->stream, &nl));
      CK(launch_lm_step(S, np, opt->use_manifold, *opt, p->lm_rec, p->lm_state, p->lm_eval_pose, p->lm_mu, p->lm_active, p->lm_n_active, d_poses,
                        d_result, ctx->stream, &nl));
    }
    CK(cudaMemcpyAsync(&sr.h_n_active[g & 1], p->lm_n_active, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(sr.ev[g & 1], ctx->stream));
    if (g >= 1) {
      CK(cudaEventSynchronize(sr.ev[(g - 1) & 1]));
      const uint32_t n_act = sr.h_n_active[(g - 1) & 1];
      done = n_act == 0u;
      if (getenv("RANDT_DEBUG_SOLVER"))
        fprintf(stderr, "[randt solver] group %lld (iterations <= %lld): %u active  t=%.1f us\n", g - 1, g * poll, n_act,
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_solver0).count());
      // finished segments leave holes K3 has to step over: compact the schedule once a quarter of its segments are gone
      if (!done && (unsigned long long)n_act * 4ull <= (unsigned long long)planned_for * 3ull) {
        CK(launch_replan(p->chunks, p->n_chunks, p->lm_active, p->n_warps, p->lm_flags, p->lm_scan, p->lm_bs, p->lm_chunks, p->lm_warp_off,
                         ctx->stream, &nl));
        v.chunks = p->lm_chunks; v.warp_off = p->lm_warp_off; v.plan_static = 0u;
        planned_for = n_act;
      }
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (!done) done = sr.h_n_active[(n_groups - 1) & 1] == 0u;
  ctx->launches += nl;
  if (!done) return fail(ctx, RANDT_E_NONFINITE, "randt_register_batch: iteration cap reached with active segments (non-finite evaluations?)");
  return RANDT_OK;
}
}  // namespace

int randt_register_batch(randt_ctx* ctx, const randt_problem* cp, int variant, double* poses, const randt_loss* loss,
                         const randt_solver_options* opt, double* result) {
  return randt_register_batch_weighted(ctx, cp, variant, poses, loss, nullptr, opt, result);
}

int randt_register_batch_weighted(randt_ctx* ctx, const randt_problem* cp, int variant, double* poses, const randt_loss* loss,
                                  const double* weight_per_seg, const randt_solver_options* opt, double* result) {
  if (!ctx || !cp || !poses || !opt || !result) return fail(ctx, RANDT_E_INVALID, "randt_register_batch: null argument");
  if (check_variant(ctx, variant)) return RANDT_E_INVALID;
  randt_problem* p = const_cast<randt_problem*>(cp);
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  const uint32_t S = p->S;
  const int np = np_of(variant);
  if (S == 0) return RANDT_OK;
  if (!p->lm_poses) { CK(dev_alloc(&p->lm_poses, (size_t)S * 4)); CK(dev_alloc(&p->lm_result, (size_t)S * RANDT_REG_STRIDE)); }
  CK(cudaMemcpyAsync(p->lm_poses, poses, (size_t)S * np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (weight_per_seg) {
    if (!p->lm_weight) CK(dev_alloc(&p->lm_weight, S));
    CK(cudaMemcpyAsync(p->lm_weight, weight_per_seg, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = register_batch_impl(ctx, p, variant, p->lm_poses, loss, weight_per_seg ? p->lm_weight : nullptr, opt, p->lm_result);
  if (rc) return rc;
  const size_t n_po = (size_t)S * np, n_re = (size_t)S * RANDT_REG_STRIDE;
  if ((n_po + n_re) * sizeof(double) <= ((size_t)1 << 20)) {       // small batches: both tables through the pinned block, one wait
    if (int rc2 = pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, 2 * (n_po + n_re))) return rc2;
    double* h = reinterpret_cast<double*>(ctx->h_offs);
    CK(cudaMemcpyAsync(h, p->lm_poses, n_po * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h + n_po, p->lm_result, n_re * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(poses, h, n_po * sizeof(double)); memcpy(result, h + n_po, n_re * sizeof(double));
    return RANDT_OK;
  }
  CK(cudaMemcpyAsync(poses, p->lm_poses, (size_t)S * np * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(result, p->lm_result, (size_t)S * RANDT_REG_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RANDT_OK;
}

namespace {
// One scan against one submap as ONE chain on the device: points up, K1, the fused association (which reads the scan's cell count from
// where K1 left it and leaves K3's records, a one-registration layout and the scan's ScaledLoss weight in device memory), K7 — and one
// copy back with pose, result, totals, cell count and K1's status.  One synchronisation; no pair lists, no K3 schedule.  Same kernels
// and the same arithmetic as randt_voxelize + randt_associate + randt_register_batch.
// -> *scan_out = the scan's map (for the keyframe insertion); *handled = false when the registration has to be redone through the
//    general path (more duos than the persistent solver takes in a batch call, escapes beyond the table): the map is valid either way.
int scan_chain(randt_ctx* ctx, const randt_map* F, const float* pts4, uint32_t n_pts, const randt_grid_params* gp, int k, int metric, const randt_loss* loss,
               double ndt_weight, const randt_solver_options* opt, double* pose_io, double* res, randt_map** scan_out, bool* handled) {
  *handled = false; *scan_out = nullptr;
  MapGeomDev geom = F->geom;
  geom.r_stop = static_cast<int>(F->gp.max_linf / F->gp.resolution);
  randt_loss l0 = *loss; l0.mu = 1.0;
  LossParams lp;
  if (int rc = make_loss(ctx, &l0, &lp)) return rc;
  CK(cudaSetDevice(ctx->device));
  StreamScope scope__(ctx->stream, ctx->sref->pool);
  static const bool trace = getenv("RANDT_DEBUG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[randt] scan_chain %-14s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_prev).count());
    t_prev = t;
  };
  const uint32_t div = (uint32_t)std::max(gp->min_points, 0) + 1u;
  const uint32_t cell_cap = std::max<uint32_t>(1, n_pts / div);
  const size_t max_duos = (size_t)cell_cap * ((k + 1) / 2);
  randt_map* m = new (std::nothrow) randt_map();
  if (!m) return RANDT_E_NOMEM;
  m->device = ctx->device; m->sref = ctx->sref; m->gp = *gp; m->geom = make_geom(*gp); m->B = 1;
  // one block: [pose 4 | result 8 | weight 1 | pad 1] doubles, [totals 3 | counter 1 | layout 6 | cells 1 | status 1 | scan_off 2 | pad 2] words,
  // records, escapes
  constexpr uint32_t ovf_cap = 64;
  constexpr size_t io_doubles = 4 + RANDT_REG_STRIDE + 2, io_words = 16;
  const size_t off_rec = (io_doubles * 8 + io_words * 4 + 127) & ~(size_t)127;
  const size_t off_ovf = off_rec + ((max_duos * sizeof(DuoRec) + 127) & ~(size_t)127);
  const size_t bytes = off_ovf + ovf_cap * sizeof(DuoRecFull);
  unsigned char* blk = nullptr; float4* d_pts = nullptr; unsigned short* d_bins = nullptr;
  struct Back { double pose[4]; double res[RANDT_REG_STRIDE]; double weight[2]; uint32_t words[io_words]; };
  static_assert(sizeof(Back) == io_doubles * 8 + io_words * 4, "the read-back mirrors the head of the block");
  int nl = 0;
  cudaError_t e = dev_alloc(&blk, bytes);
  if (e == cudaSuccess) e = dev_alloc(&d_pts, n_pts);
  if (e == cudaSuccess) e = dev_alloc(&m->cells, (size_t)cell_cap * 3);
  if (e == cudaSuccess) e = dev_alloc(&m->npts, cell_cap);
  if (e == cudaSuccess) e = dev_alloc(&m->labels, cell_cap);
  if (e == cudaSuccess) e = dev_alloc(&m->slot, m->geom.n_slots);
  if (e == cudaSuccess) e = dev_alloc(&m->cell_off, 2);
  if (e == cudaSuccess && voxelize_needs_bins_scratch(n_pts, cell_cap, *gp)) e = dev_alloc(&d_bins, n_pts);
  if (e == cudaSuccess && pinned_reserve(ctx, &ctx->h_offs, &ctx->offs_cap, sizeof(Back) / sizeof(uint32_t)) != RANDT_OK) e = cudaErrorMemoryAllocation;
  double* d_pose = reinterpret_cast<double*>(blk); double* d_res = d_pose + 4; double* d_weight = d_res + RANDT_REG_STRIDE;
  uint32_t* d_words = reinterpret_cast<uint32_t*>(blk + io_doubles * 8);
  uint32_t *d_tot = d_words, *d_counter = d_words + 3, *d_layout = d_words + 4, *d_cnt = d_words + 10, *d_scan_off = d_words + 12;
  int* d_status = reinterpret_cast<int*>(d_words + 11);
  DuoRec* d_recs = reinterpret_cast<DuoRec*>(blk + off_rec);
  DuoRecFull* d_ovf = reinterpret_cast<DuoRecFull*>(blk + off_ovf);
  Back& h = *reinterpret_cast<Back*>(ctx->h_offs);
  if (e == cudaSuccess) {
    // the head of the block goes up in one copy: pose, zeros, the loss weight for scans without ndt_weight, the scan's point range
    memset(&h, 0, sizeof(Back));
    memcpy(h.pose, pose_io, sizeof(h.pose));
    h.weight[0] = lp.weight;
    h.words[12] = 0u; h.words[13] = n_pts;
    e = cudaMemcpyAsync(blk, &h, sizeof(Back), cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess && n_pts) e = cudaMemcpyAsync(d_pts, pts4, (size_t)n_pts * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = launch_voxelize(d_pts, d_scan_off, 1, n_pts, *gp, m->geom, cell_cap, m->cells, m->npts, m->labels, d_cnt, m->slot, d_bins, d_status, ctx->stream, &nl);
  if (e == cudaSuccess) e = launch_associate_single(F->cells, F->n_cells, F->slot, m->cells, 0u, geom, d_pose, k, metric, nullptr, nullptr, d_recs, nullptr, d_ovf, ovf_cap,
                                                    nullptr, nullptr, d_tot, d_layout, d_cnt, ndt_weight > 0.0 ? ndt_weight : 0.0, ndt_weight > 0.0 ? d_weight : nullptr,
                                                    ctx->stream, &nl);
  if (e == cudaSuccess) {
    DeviceProblem v;
    v.duo_recs = d_recs; v.duo_overflow = d_ovf; v.seg_off = d_layout; v.seg_first_tile = d_layout + 4; v.n_segments = 1;
    SolveLayout L;
    // one tile that covers any registration: the chunk index never leaves it
    L.seg_duo_off = d_layout + 2; L.tile_rec_begin = d_layout + 5; L.tile_duos = 1u << 20; L.items = nullptr; L.n_items = 1; L.next_item = d_counter;
    e = launch_solve_persistent(v, L, 0, opt->use_manifold, lp, d_weight, *opt, d_pose, d_pose, d_res, ctx->d_bad, ctx->stream, &nl);
  }
  lap("queued");
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h, blk, sizeof(Back), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);      // (the pinned block is free again: the upload has long left it)
  lap("result back");
  dev_free(blk); dev_free(d_pts); dev_free(d_bins);
  if (e != cudaSuccess) { free_map(m); return fail(ctx, RANDT_E_CUDA, "randt_scan_step (voxelise + associate + solve chain)", e); }
  ctx->launches += nl;
  const int status = (int)h.words[11];
  if (status == VOX_SPAN || status == VOX_CELL_CAP) {
    free_map(m);
    return fail(ctx, RANDT_E_CAPACITY, "randt_voxelize: label span / cell capacity exceeded (points far outside +-max_range?)");
  }
  if (status == VOX_OUT_OF_MAP) {
    free_map(m);
    return fail(ctx, RANDT_E_INVALID, "randt_voxelize: a cell mean falls outside the map (the reference throws std::out_of_range here)");
  }
  const uint32_t n_cells = h.words[10];
  m->h_cell_off = {0u, n_cells}; m->n_cells = n_cells; m->max_per_map = n_cells;
  if (cudaMemcpyAsync(m->cell_off, m->h_cell_off.data(), 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    free_map(m);
    return fail(ctx, RANDT_E_CUDA, "randt_scan_step: cell offsets");
  }
  *scan_out = m;
  // the general path solves registrations above kSolveMaxDuos stepwise (other summation order): keep the composite equal to it
  if (h.words[2] > ovf_cap || h.words[1] > kSolveMaxDuos) return RANDT_OK;
  memcpy(pose_io, h.pose, sizeof(h.pose));
  memcpy(res, h.res, sizeof(h.res));
  *handled = true;
  return RANDT_OK;
}

// may the chain take this call?  (anything it cannot take is left to the separate entry points, which also report the errors)
bool scan_chain_applies(const randt_map* F, uint32_t n_pts, const randt_grid_params* gp, int k, int metric, const randt_solver_options* opt) {
  static const bool no_fused = getenv("RANDT_NO_FUSED_ASSOC") != nullptr;
  if (no_fused || opt->poll_interval < 0 || F->n_cells == 0 || F->n_cells > kSingleMaxCells || n_pts == 0 || n_pts > 16384u) return false;
  if (!same_geom(F->gp, *gp) || gp->n_clusters <= 0 || gp->size_x <= 0 || gp->size_y <= 0 || !(gp->resolution > 0) || !(gp->max_range > 0)) return false;
  if (k < 1 || k > kMaxNeighbours || (metric != RANDT_LOOKUP_MAHALANOBIS_INTENSITY && metric != RANDT_LOOKUP_EUCLID_XY)) return false;
  if (!(opt->gnc_divisor > 1.0) || !(opt->gnc_loss_scale > 0.0) || opt->gnc_max_steps < 1 || opt->max_num_iterations < 0 ||
      opt->max_num_consecutive_invalid_steps < 1 || !(opt->initial_trust_region_radius > 0.0)) return false;
  const int r_stop = static_cast<int>(F->gp.max_linf / F->gp.resolution);
  if (2 * (r_stop > 0 ? r_stop - 1 : 0) + 1 > F->geom.size_x) return false;
  const uint32_t cell_cap = std::max<uint32_t>(1, n_pts / ((uint32_t)std::max(gp->min_points, 0) + 1u));
  return (size_t)cell_cap * ((k + 1) / 2) <= 16384u;
}
}  // namespace

// the per-scan chain as one call: see randt_gpu.h
int randt_scan_step(randt_ctx* ctx, randt_map* submap, const float* pts4, uint32_t n_pts, const randt_grid_params* gp, int k, int metric,
                    const randt_loss* loss, double ndt_weight, const randt_solver_options* opt, int insert_keyframe, double* pose_io, double* result,
                    uint32_t* n_cells_out) {
  if (!ctx || !submap || !gp || !opt || !pose_io || (!pts4 && n_pts)) return fail(ctx, RANDT_E_INVALID, "randt_scan_step: null argument");
  if (submap->B != 1) return fail(ctx, RANDT_E_INVALID, "randt_scan_step: the submap must be a single map");
  static const bool trace = getenv("RANDT_DEBUG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[randt] scan_step %-12s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_prev).count());
    t_prev = t;
  };
  randt_loss l; l.kind = RANDT_LOSS_NONE; l.scale = 1.0; l.alpha = 2.0; l.mu = 1.0; l.weight = 1.0;
  if (loss) l = *loss;
  double res[RANDT_REG_STRIDE] = {0};
  randt_map* scan = nullptr;
  int rc = RANDT_OK;
  bool handled = false;
  if (scan_chain_applies(submap, n_pts, gp, k, metric, opt)) {
    rc = scan_chain(ctx, submap, pts4, n_pts, gp, k, metric, &l, ndt_weight, opt, pose_io, res, &scan, &handled);
    if (rc != RANDT_OK) return rc;
    lap("chain");
  } else {
    const uint32_t off[2] = {0u, n_pts};
    rc = randt_voxelize(ctx, pts4, off, 1, gp, 0, &scan);
    if (rc != RANDT_OK) return rc;
    lap("voxelize");
  }
  if (n_cells_out) *n_cells_out = scan->n_cells;
  if (submap->n_cells > 0 && !handled) {
    if (ndt_weight > 0.0 && scan->n_cells > 0) l.weight = ndt_weight / ((double)scan->n_cells * (double)k);
    randt_problem* prob = nullptr;
    rc = randt_associate(ctx, submap, scan, pose_io, k, metric, &prob);
    lap("associate");
    if (rc == RANDT_OK) {
      rc = randt_register_batch(ctx, prob, 0, pose_io, &l, opt, res);
      lap("register");
    }
    randt_problem_destroy(prob);
    lap("destroy");
  }
  if (rc == RANDT_OK && (insert_keyframe || submap->n_cells == 0)) {
    rc = randt_map_transform_se2d(ctx, scan, pose_io);
    if (rc == RANDT_OK) rc = randt_map_merge(ctx, submap, scan);
    lap("keyframe");
  }
  randt_map_destroy(scan);
  if (result) memcpy(result, res, sizeof(res));
  return rc;
}

}  // extern "C"
