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
  std::vector<ChunkDesc> planA, planB;           // plan A: split chunks allowed (full evaluation); plan B: one tile per chunk (solver, EMIT)
  std::vector<uint32_t> woffA, woffB;            // [n_warps+1] chunk ranges of every warp in the two plans
  std::vector<uint32_t> tile_rec_begin;          // [n_tiles+1] record offsets of the tiles in schedule order (warp after warp)
  std::vector<uint32_t> tile_duo_begin;          // [n_tiles]   where each of those tiles starts in the segment-ordered duo list
  std::vector<uint32_t> rec_of_tile;             // [n_tiles]   record offset of tile t, tiles in SEGMENT order (what the persistent solver walks)
};

// duo_off: [S+1] duo offsets per segment; max_warps: resident warps of the persistent grid (kK3MaxWarps)
inline void build_schedule(const uint32_t* duo_off, uint32_t S, uint32_t max_warps, Schedule& out) {
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
  for (uint32_t s = 0; s < S; ++s) {
    first[s] = (uint32_t)tiles.size();
    uint32_t part = 0;
    for (uint32_t b = duo_off[s]; b < duo_off[s + 1]; b += tile_duos) {
      Tile t; t.seg = s; t.begin = b; t.end = std::min(duo_off[s + 1], b + tile_duos); t.part = part++;
      tiles.push_back(t);
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
  std::vector<uint32_t> mine_off(n_warps + 1, 0), mine(T);    // tiles of warp w, ascending: mine[mine_off[w] .. mine_off[w + 1])
  {
    auto cost = [&](uint32_t t) { return (tiles[t].end - tiles[t].begin) + 24u; };
    const uint32_t max_cost = tile_duos + 24u;
    std::vector<uint32_t> bucket(max_cost + 2, 0), order(T), warp_of(T);
    for (uint32_t t = 0; t < T; ++t) ++bucket[max_cost - cost(t) + 1];
    for (uint32_t c = 0; c <= max_cost; ++c) bucket[c + 1] += bucket[c];
    for (uint32_t t = 0; t < T; ++t) order[bucket[max_cost - cost(t)]++] = t;      // descending cost, ties in tile order
    // least-loaded warp through a bucket queue over the (integer) loads: the minimum load never decreases, and greedy assignment
    // keeps every load below average + max_cost
    uint64_t total = 0;
    for (uint32_t t = 0; t < T; ++t) total += cost(t);
    const uint32_t n_loads = (uint32_t)(total / n_warps) + 2u * max_cost + 2u;
    std::vector<int32_t> head(n_loads, -1), next(n_warps, -1);
    for (uint32_t w = n_warps; w-- > 0;) { next[w] = head[0]; head[0] = (int32_t)w; }
    uint32_t cur = 0;
    for (uint32_t t : order) {
      while (head[cur] < 0) ++cur;
      const uint32_t w = (uint32_t)head[cur];
      head[cur] = next[w];
      warp_of[t] = w; ++mine_off[w + 1];
      const uint32_t nl = cur + cost(t);
      next[w] = head[nl]; head[nl] = (int32_t)w;
    }
    for (uint32_t w = 0; w < n_warps; ++w) mine_off[w + 1] += mine_off[w];
    std::vector<uint32_t> fill(mine_off.begin(), mine_off.end() - 1);
    for (uint32_t t = 0; t < T; ++t) mine[fill[warp_of[t]]++] = t;
  }
  // Records are laid out in schedule order (warp after warp, tile after tile), so that a warp streams one contiguous range and the
  // tail of one tile and the head of the next can share a chunk.  Two chunk lists over the same records:
  //   plan A (full evaluation): consecutive solo tiles of a warp are packed — the last, partly filled chunk of a tile takes the first
  //           duos of the next tile (kChunkSplit), which keeps the lanes busy when problems are only ~3 chunks long;
  //   plan B (solver with active flags, EMIT): every chunk belongs to one tile.
  std::vector<uint32_t>& tile_rec_begin = out.tile_rec_begin; std::vector<uint32_t>& tile_duo_begin = out.tile_duo_begin;
  tile_rec_begin.clear(); tile_duo_begin.clear();
  out.rec_of_tile.assign(tiles.size(), 0u);
  tile_rec_begin.reserve(tiles.size() + 1); tile_duo_begin.reserve(tiles.size());
  std::vector<ChunkDesc>& planA = out.planA; std::vector<ChunkDesc>& planB = out.planB;
  planA.clear(); planB.clear();
  planA.reserve(n_duos / 32 + tiles.size() + 1); planB.reserve(n_duos / 32 + tiles.size() + 1);
  std::vector<uint32_t>& woffA = out.woffA; std::vector<uint32_t>& woffB = out.woffB;
  woffA.assign(n_warps + 1, 0); woffB.assign(n_warps + 1, 0);
  uint32_t rec = 0;
  for (uint32_t w = 0; w < n_warps; ++w) {
    bool open = false;      // the last chunk of plan A ends a solo tile, is not split yet and has free lanes
    for (uint32_t q = mine_off[w]; q < mine_off[w + 1]; ++q) {
      const uint32_t t = mine[q];
      const Tile& tl = tiles[t];
      const bool solo = first[tl.seg + 1] - first[tl.seg] == 1u;
      const uint32_t len = tl.end - tl.begin, rb = rec;
      tile_rec_begin.push_back(rb); tile_duo_begin.push_back(tl.begin);
      out.rec_of_tile[t] = rb;
      rec += len;
      const uint32_t part = first[tl.seg] + tl.part;
      for (uint32_t o = 0; o < len; o += 32u) {
        ChunkDesc c;
        c.duo_begin = rb + o; c.seg = tl.seg; c.part = part;
        c.meta = std::min(32u, len - o) | (o == 0 ? kChunkFirst : 0u) | (o + 32u >= len ? kChunkLast : 0u) | (solo ? kChunkSolo : 0u);
        planB.push_back(c);
      }
      uint32_t o = 0;
      if (solo && open) {   // this tile starts in the free lanes of the previous tile's last chunk
        ChunkDesc& pc = planA.back();
        const uint32_t n_old = pc.meta & kChunkCountMask, take = std::min(32u - n_old, len);
        pc.meta = (pc.meta & ~kChunkCountMask) | (n_old + take) | kChunkSplit | (n_old << kChunkSplitShift) | (take == len ? kChunkNewLast : 0u);
        pc.part = tl.seg;
        o = take;
      }
      open = false;
      for (; o < len; o += 32u) {
        ChunkDesc c;
        c.duo_begin = rb + o; c.seg = tl.seg; c.part = part;
        const uint32_t n = std::min(32u, len - o);
        c.meta = n | (o == 0 ? kChunkFirst : 0u) | (o + 32u >= len ? kChunkLast : 0u) | (solo ? kChunkSolo : 0u);
        planA.push_back(c);
        open = solo && (o + 32u >= len) && n < 32u;
      }
    }
    woffA[w + 1] = (uint32_t)planA.size(); woffB[w + 1] = (uint32_t)planB.size();
  }
  tile_rec_begin.push_back(rec);
}

}  // namespace randt
