// The K3 work schedule: tiles, their balanced assignment to the resident warps of the persistent grid, and the two chunk-descriptor
// lists the kernels walk.  Plain C++ (no CUDA types) so that the host logic can be exercised without a device
// (tests/test_schedule_cpu.py drives it through librandt_host.so); capi.cu uploads what build_schedule() produces.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace randt {

constexpr int kSmCount = 148;     // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kTileDuos = 256;    // most duos per tile (one warp owns a tile)
constexpr int kMinTileDuos = 32;  // tile length for small problems

// A work tile: duos [begin, end) of segment `seg`; `part` = index of the tile inside its segment.  One warp owns a tile.
struct Tile { uint32_t seg, begin, end, part; };
// What the kernels walk: a tile cut into chunks of <= 32 duos (one per lane).  meta: bits 0..5 = duos in the chunk (0 marks the end of
// a warp's list), kChunkFirst / kChunkLast = first / last chunk of its tile, kChunkSolo = the tile is its segment's only tile.
// part = index of the tile's partial record (seg_first_tile[seg] + tile.part), used when a segment spans several tiles.
struct alignas(16) ChunkDesc { uint32_t duo_begin, meta, seg, part; };
constexpr uint32_t kChunkCountMask = 0x3fu, kChunkFirst = 0x100u, kChunkLast = 0x200u, kChunkSolo = 0x400u;
// Split chunks (full-evaluation schedule only): lanes [0, sp) finish one solo tile (segment `seg`, which therefore has kChunkLast) and
// lanes [sp, count) start the next solo tile of the same warp, whose segment is carried in `part`; sp sits in bits 16..21.
// kChunkNewLast: that second tile also ends inside this chunk.  The record table is laid out in schedule order so that the two
// tiles' records are adjacent.
constexpr uint32_t kChunkSplit = 0x800u, kChunkNewLast = 0x1000u;
constexpr int kChunkSplitShift = 16;

struct Schedule {
  std::vector<Tile> tiles;                       // segment after segment
  std::vector<uint32_t> first;                   // [S+1] first tile of every segment
  uint32_t n_warps = 0;
  uint32_t tile_duos = 0;                        // duos per tile (a multiple of 32; a segment's last tile may be shorter)
  std::vector<uint32_t> mine, mine_off;          // tiles of warp w, ascending: mine[mine_off[w] .. mine_off[w + 1])
  std::vector<uint32_t> warp_rec_begin;          // [n_warps] record offset at which warp w's tiles start
  std::vector<ChunkDesc> planA, planB;           // plan A: split chunks allowed (full evaluation); plan B: one tile per chunk (solver, EMIT)
  std::vector<uint32_t> woffA, woffB;            // [n_warps+1] chunk ranges of every warp in the two plans
  std::vector<uint32_t> tile_rec_begin;          // [n_tiles+1] record offsets of the tiles in schedule order (warp after warp)
  std::vector<uint32_t> tile_duo_begin;          // [n_tiles]   where each of those tiles starts in the segment-ordered duo list
  std::vector<uint32_t> rec_of_tile;             // [n_tiles]   record offset of tile t, tiles in SEGMENT order (what the persistent solver walks)
  // scratch of the assignment, kept so that a schedule object that is used again does not allocate
  std::vector<uint32_t> bucket_, order_, warp_of_, fill_;
  std::vector<int32_t> head_, next_;
};

#if defined(__CUDACC__)
#define RANDT_HD __host__ __device__
#else
#define RANDT_HD
#endif

// The chunk lists of ONE warp of the schedule: its tiles mine[q0 .. q1) in order, records from rec0 on.  outB(i, desc) / outA(i, desc)
// receive the warp's chunks of the two plans (i counts from 0); onTile(q, tile index, record offset) is called once per tile;
// tile_of(t) returns tile t.
// Returns the record offset after the warp's last tile; *nA / *nB = chunks emitted.  The same code counts (empty callbacks, host),
// writes the host vectors (build_schedule) and runs one thread per warp on the device (capi.cu uploads only the assignment).
//   plan A (full evaluation): consecutive solo tiles of a warp are packed — the last, partly filled chunk of a tile takes the first
//           duos of the next tile (kChunkSplit), which keeps the lanes busy when problems are only ~3 chunks long;
//   plan B (solver with active flags, EMIT): every chunk belongs to one tile.
template <class TileOf, class OutA, class OutB, class OnTile>
RANDT_HD inline uint32_t walk_warp(TileOf&& tile_of, const uint32_t* first, const uint32_t* mine, uint32_t q0, uint32_t q1, uint32_t rec0,
                                   OutA&& outA, OutB&& outB, OnTile&& onTile, uint32_t* nA_out, uint32_t* nB_out) {
  uint32_t rec = rec0, nA = 0, nB = 0;
  ChunkDesc pend; pend.duo_begin = 0; pend.meta = 0; pend.seg = 0; pend.part = 0;
  bool have = false;      // pend holds plan A's latest chunk (a later solo tile may still move into its free lanes)
  bool open = false;      // ... and it ends a solo tile, is not split yet and has free lanes
  for (uint32_t q = q0; q < q1; ++q) {
    const uint32_t t = mine[q];
    const Tile tl = tile_of(t);
    const bool solo = first[tl.seg + 1] - first[tl.seg] == 1u;
    const uint32_t len = tl.end - tl.begin, rb = rec;
    onTile(q, t, rb);
    rec += len;
    const uint32_t part = first[tl.seg] + tl.part;
    for (uint32_t o = 0; o < len; o += 32u) {
      ChunkDesc c;
      c.duo_begin = rb + o; c.seg = tl.seg; c.part = part;
      const uint32_t rest = len - o;
      c.meta = (rest < 32u ? rest : 32u) | (o == 0 ? kChunkFirst : 0u) | (o + 32u >= len ? kChunkLast : 0u) | (solo ? kChunkSolo : 0u);
      outB(nB++, c);
    }
    uint32_t o = 0;
    if (solo && open) {   // this tile starts in the free lanes of the previous tile's last chunk
      const uint32_t n_old = pend.meta & kChunkCountMask, room = 32u - n_old, take = room < len ? room : len;
      pend.meta = (pend.meta & ~kChunkCountMask) | (n_old + take) | kChunkSplit | (n_old << kChunkSplitShift) | (take == len ? kChunkNewLast : 0u);
      pend.part = tl.seg;
      o = take;
    }
    open = false;
    for (; o < len; o += 32u) {
      if (have) outA(nA++, pend);
      const uint32_t rest = len - o, n = rest < 32u ? rest : 32u;
      pend.duo_begin = rb + o; pend.seg = tl.seg; pend.part = part;
      pend.meta = n | (o == 0 ? kChunkFirst : 0u) | (o + 32u >= len ? kChunkLast : 0u) | (solo ? kChunkSolo : 0u);
      have = true;
      open = solo && (o + 32u >= len) && n < 32u;
    }
  }
  if (have) outA(nA++, pend);
  *nA_out = nA; *nB_out = nB;
  return rec;
}

// tile t of segment seg (what the device rebuilds from tile_seg, first and the duo offsets instead of reading a tile table)
RANDT_HD inline Tile tile_from_seg(uint32_t t, uint32_t seg, const uint32_t* first, const uint32_t* duo_off, uint32_t tile_duos) {
  Tile tl;
  tl.seg = seg; tl.part = t - first[seg];
  tl.begin = duo_off[seg] + tl.part * tile_duos;
  const uint32_t e = duo_off[seg + 1];
  tl.end = e - tl.begin < tile_duos ? e : tl.begin + tile_duos;
  return tl;
}

// Everything of the schedule except the chunk lists themselves: tiles, their balanced assignment, the record layout and the chunk
// ranges of every warp.  duo_off: [S+1] duo offsets per segment; max_warps: resident warps of the persistent grid (kK3MaxWarps)
inline void build_schedule_core(const uint32_t* duo_off, uint32_t S, uint32_t max_warps, Schedule& out) {
  const uint32_t n_duos = duo_off[S];
  std::vector<Tile>& tiles = out.tiles;
  std::vector<uint32_t>& first = out.first;
  tiles.clear();
  first.assign((size_t)S + 1, 0);
  // one warp owns a tile.  Big batches: tiles of up to kTileDuos duos (a whole ~200-pair registration per warp, no partials);
  // small problems: shorter tiles so that the pairs still spread over the SMs (each extra tile costs one partial record).
  uint32_t tile_duos = (n_duos / (uint32_t)(kSmCount * 4) + 31u) / 32u * 32u;
  tile_duos = std::max<uint32_t>(kMinTileDuos, std::min<uint32_t>(tile_duos, kTileDuos));
  out.tile_duos = tile_duos;
  {
    uint32_t T = 0;
    for (uint32_t s = 0; s < S; ++s) { first[s] = T; T += (duo_off[s + 1] - duo_off[s] + tile_duos - 1u) / tile_duos; }
    tiles.resize(T);
    Tile* tp = tiles.data();
    for (uint32_t s = 0; s < S; ++s) {
      uint32_t part = 0;
      const uint32_t e = duo_off[s + 1];
      for (uint32_t b = duo_off[s]; b < e; b += tile_duos, ++part) {
        Tile& t = tp[first[s] + part];
        t.seg = s; t.begin = b; t.end = e - b < tile_duos ? e : b + tile_duos; t.part = part;
      }
    }
  }
  first[S] = (uint32_t)tiles.size();
  // Balanced static schedule: longest-processing-time assignment of tiles to the resident warps of the persistent grid (a tile
  // costs its duos plus a fixed prologue/reduce/emit overhead).  Registration problems differ in size, and one warp walks only ~7
  // of them per launch at the bench size, so round-robin striding leaves warps (and whole SMs) idle at the tail.
  const uint32_t n_warps = std::max<uint32_t>(1u, std::min<uint32_t>(max_warps, (uint32_t)tiles.size()));
  out.n_warps = n_warps;
  // (tile costs are small integers: counting sort and a bucket queue make this linear in the number of tiles)
  const uint32_t T = (uint32_t)tiles.size();
  std::vector<uint32_t>&mine_off = out.mine_off, &mine = out.mine;
  mine_off.assign((size_t)n_warps + 1, 0); mine.resize(T);
  {
    auto cost = [&](uint32_t t) { return (tiles[t].end - tiles[t].begin) + 24u; };
    const uint32_t max_cost = tile_duos + 24u;
    std::vector<uint32_t>&bucket = out.bucket_, &order = out.order_, &warp_of = out.warp_of_;
    bucket.assign((size_t)max_cost + 2, 0); order.resize(T); warp_of.resize(T);
    for (uint32_t t = 0; t < T; ++t) ++bucket[max_cost - cost(t) + 1];
    for (uint32_t c = 0; c <= max_cost; ++c) bucket[c + 1] += bucket[c];
    for (uint32_t t = 0; t < T; ++t) order[bucket[max_cost - cost(t)]++] = t;      // descending cost, ties in tile order
    // least-loaded warp through a bucket queue over the (integer) loads: the minimum load never decreases, and greedy assignment
    // keeps every load below average + max_cost
    uint64_t total = 0;
    for (uint32_t t = 0; t < T; ++t) total += cost(t);
    const uint32_t n_loads = (uint32_t)(total / n_warps) + 2u * max_cost + 2u;
    std::vector<int32_t>&head = out.head_, &next = out.next_;
    head.assign(n_loads, -1); next.assign(n_warps, -1);
    for (uint32_t w = n_warps; w-- > 0;) { next[w] = head[0]; head[0] = (int32_t)w; }
    uint32_t cur = 0;
    for (uint32_t i = 0; i < T; ++i) {
      const uint32_t t = order[i];
      while (head[cur] < 0) ++cur;
      const uint32_t w = (uint32_t)head[cur];
      head[cur] = next[w];
      warp_of[t] = w; ++mine_off[w + 1];
      const uint32_t nl = cur + cost(t);
      next[w] = head[nl]; head[nl] = (int32_t)w;
    }
    for (uint32_t w = 0; w < n_warps; ++w) mine_off[w + 1] += mine_off[w];
    std::vector<uint32_t>& fill = out.fill_;
    fill.assign(mine_off.begin(), mine_off.end() - 1);
    for (uint32_t t = 0; t < T; ++t) mine[fill[warp_of[t]]++] = t;
  }
  // Records are laid out in schedule order (warp after warp, tile after tile), so that a warp streams one contiguous range and the
  // tail of one tile and the head of the next can share a chunk.  Counting pass of walk_warp: record offsets and chunk ranges.
  out.tile_rec_begin.resize((size_t)T + 1); out.tile_duo_begin.resize(T); out.rec_of_tile.resize(T);
  out.woffA.assign((size_t)n_warps + 1, 0); out.woffB.assign((size_t)n_warps + 1, 0);
  out.warp_rec_begin.resize(n_warps);
  uint32_t* trb = out.tile_rec_begin.data(); uint32_t* tdb = out.tile_duo_begin.data(); uint32_t* rot = out.rec_of_tile.data();
  const Tile* tp = tiles.data();
  uint32_t rec = 0;
  const uint32_t* fp = first.data(); const uint32_t* mp = mine.data();
  for (uint32_t w = 0; w < n_warps; ++w) {
    // (walk_warp's bookkeeping in closed form: a tile of len duos has ceil(len / 32) chunks in plan B; in plan A a solo tile first
    //  fills the free lanes of an open chunk)
    uint32_t nA = 0, nB = 0, room = 0;      // room: free lanes of plan A's open chunk (0 = none open)
    out.warp_rec_begin[w] = rec;
    for (uint32_t q = mine_off[w]; q < mine_off[w + 1]; ++q) {
      const uint32_t t = mp[q];
      const Tile& tl = tp[t];
      const bool solo = fp[tl.seg + 1] - fp[tl.seg] == 1u;
      const uint32_t len = tl.end - tl.begin;
      trb[q] = rec; tdb[q] = tl.begin; rot[t] = rec;
      rec += len;
      nB += (len + 31u) >> 5;
      const uint32_t take = (solo && room) ? (room < len ? room : len) : 0u;
      const uint32_t rest = len - take;
      nA += (rest + 31u) >> 5;
      room = (solo && rest && (rest & 31u)) ? 32u - (rest & 31u) : 0u;
    }
    out.woffA[w + 1] = out.woffA[w] + nA; out.woffB[w + 1] = out.woffB[w] + nB;
  }
  trb[T] = rec;
}

// the whole schedule on the host (tests, and the reference the device-emitted chunk lists are compared with)
inline void build_schedule(const uint32_t* duo_off, uint32_t S, uint32_t max_warps, Schedule& out) {
  build_schedule_core(duo_off, S, max_warps, out);
  out.planA.resize(out.woffA[out.n_warps]); out.planB.resize(out.woffB[out.n_warps]);
  for (uint32_t w = 0; w < out.n_warps; ++w) {
    ChunkDesc* a = out.planA.data() + out.woffA[w]; ChunkDesc* b = out.planB.data() + out.woffB[w];
    uint32_t nA = 0, nB = 0;
    const Tile* tp = out.tiles.data();
    walk_warp([tp](uint32_t t) { return tp[t]; }, out.first.data(), out.mine.data(), out.mine_off[w], out.mine_off[w + 1], out.warp_rec_begin[w],
              [a](uint32_t i, const ChunkDesc& c) { a[i] = c; }, [b](uint32_t i, const ChunkDesc& c) { b[i] = c; }, [](uint32_t, uint32_t, uint32_t) {}, &nA, &nB);
  }
}

}  // namespace randt
