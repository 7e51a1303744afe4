"""The K3 work schedule (randt_slam_b200/csrc/schedule.hpp) checked on the CPU: tiles, their longest-processing-time assignment to the
resident warps, the record layout in schedule order, and the two chunk-descriptor lists the kernels walk (plan A with split chunks for
full evaluations, plan B with one tile per chunk for the solver and EMIT).  The builder is the code capi.cu runs at problem
construction; it is driven here through librandt_host.so, which loads without a GPU.

Besides structural invariants, both plans are "executed" by a Python model of how k3_fused_kernel consumes descriptors (accumulate
per lane, finish a tile on kChunkLast, hand over inside a split chunk): every segment must receive each of its duos exactly once."""
import numpy as np
import pytest

from randt_slam_b200 import hostapi



@pytest.fixture(scope="module", autouse=True)
def _libraries_built():
    """the schedule builder is reached through librandt_host.so: make sure it exists (nvcc / g++ cross-compile without a GPU)"""
    from randt_slam_b200 import build
    build.build_all()


COUNT, FIRST, LAST, SOLO, SPLIT, NEWLAST, SPLIT_SHIFT = 0x3F, 0x100, 0x200, 0x400, 0x800, 0x1000, 16
MAX_WARPS = 148 * 4 * 4


def offsets(sizes):
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)


def cases():
    rng = np.random.default_rng(0)
    yield "bench-like", rng.integers(60, 130, 16384)
    yield "ragged", np.array([1, 2, 31, 32, 33, 5, 63, 64, 65, 1, 1, 1, 40, 0, 255, 256, 257, 3, 600, 17, 96, 0, 700, 29, 30, 31, 32, 33, 34, 7])
    yield "single small", np.array([5])
    yield "single huge", np.array([150000])
    yield "all empty", np.zeros(7, np.int64)
    yield "no segments", np.zeros(0, np.int64)
    yield "tiny many", rng.integers(0, 4, 5000)
    yield "few long", rng.integers(2000, 9000, 40)
    yield "mixed", np.concatenate([rng.integers(1, 40, 300), rng.integers(200, 3000, 30), np.zeros(11, np.int64)])[rng.permutation(341)]


@pytest.mark.parametrize("name,sizes", list(cases()), ids=[c[0] for c in cases()])
@pytest.mark.parametrize("max_warps", [MAX_WARPS, 7])
def test_schedule_invariants(name, sizes, max_warps):
    sizes = np.asarray(sizes, np.int64)
    S = len(sizes)
    off = offsets(sizes)
    n_duos = int(off[-1])
    sc = hostapi.build_schedule(off, max_warps)
    tiles, first = sc["tiles"].astype(np.int64), sc["first"].astype(np.int64)
    T = len(tiles)
    # ---- tiles: each segment cut in order into pieces of one common length (a multiple of 32 in [32, 256]) ----
    assert first[0] == 0 and first[-1] == T and (np.diff(first) >= 0).all()
    for s in range(S):
        ts = tiles[first[s]:first[s + 1]]
        if sizes[s] == 0:
            assert len(ts) == 0
            continue
        assert ts[0, 1] == off[s] and ts[-1, 2] == off[s + 1]
        assert (ts[:, 0] == s).all() and (ts[1:, 1] == ts[:-1, 2]).all() and (ts[:, 3] == np.arange(len(ts))).all()
    lens = tiles[:, 2] - tiles[:, 1]
    if T:
        tile_len = lens.max()
        assert lens.min() >= 1 and tile_len <= 256
        per_seg_last = np.zeros(T, bool); per_seg_last[first[1:][np.diff(first) > 0] - 1] = True
        full = lens[~per_seg_last]
        assert len(full) == 0 or (len(set(full.tolist())) == 1 and full[0] % 32 == 0 and 32 <= full[0] <= 256)
    # ---- assignment: every tile belongs to exactly one warp; records laid out warp after warp, tile after tile ----
    n_warps = sc["n_warps"]
    assert n_warps == max(1, min(max_warps, T))
    trb, tdb = sc["tile_rec_begin"].astype(np.int64), sc["tile_duo_begin"].astype(np.int64)
    assert len(trb) == T + 1 and trb[0] == 0 and trb[-1] == n_duos == sc["n_records"]
    begin_to_tile = {int(b): i for i, b in enumerate(tiles[:, 1])}
    sched_tiles = np.array([begin_to_tile[int(b)] for b in tdb], np.int64)            # schedule position -> tile id
    assert sorted(sched_tiles.tolist()) == list(range(T))                               # a permutation of the tiles
    assert np.array_equal(np.diff(trb), lens[sched_tiles])
    # ---- plan B: one tile per chunk ----
    pb, wb = sc["plan_b"].astype(np.int64), sc["woff_b"].astype(np.int64)
    assert wb[0] == 0 and wb[-1] == len(pb) and (np.diff(wb) >= 0).all() and len(wb) == n_warps + 1
    seen = np.zeros(n_duos, np.int32)
    pos = 0                                                                             # position in schedule order
    loads = []
    for w in range(n_warps):
        load = 0
        j = wb[w]
        while j < wb[w + 1]:
            t = sched_tiles[pos]; seg = tiles[t, 0]; ln = lens[t]
            solo = first[seg + 1] - first[seg] == 1
            n_chunks = -(-ln // 32)
            for c in range(n_chunks):
                db, meta, cseg, part = pb[j + c]
                cnt = meta & COUNT
                assert cnt == min(32, ln - 32 * c) and db == trb[pos] + 32 * c and cseg == seg and part == first[seg] + tiles[t, 3]
                assert bool(meta & FIRST) == (c == 0) and bool(meta & LAST) == (c == n_chunks - 1) and bool(meta & SOLO) == solo
                assert not meta & (SPLIT | NEWLAST)
                seen[db:db + cnt] += 1
            j += n_chunks; pos += 1; load += ln + 24
        loads.append(load)
    assert pos == T and (seen == 1).all()
    # longest-processing-time bound: no warp exceeds the average load by more than one (largest) tile
    if T:
        assert max(loads) <= sum(loads) / n_warps + lens.max() + 24
        if n_warps < T and name == "bench-like":
            assert max(loads) <= 1.05 * sum(loads) / n_warps
    # ---- plan A: the same records; a tile may start in the free lanes of the previous solo tile's last chunk ----
    pa, wa = sc["plan_a"].astype(np.int64), sc["woff_a"].astype(np.int64)
    assert wa[0] == 0 and wa[-1] == len(pa) and len(wa) == n_warps + 1 and len(pa) <= len(pb)
    got = run_plan(pa, wa, first, np.diff(off.astype(np.int64)), n_duos)
    want = run_plan(pb, wb, first, np.diff(off.astype(np.int64)), n_duos)
    assert got == want == {s: int(sizes[s]) for s in range(S) if sizes[s] > 0}
    # deterministic
    sc2 = hostapi.build_schedule(off, max_warps)
    assert all(np.array_equal(sc[k], sc2[k]) for k in ("tiles", "plan_a", "plan_b", "woff_a", "woff_b", "tile_rec_begin", "tile_duo_begin"))


def run_plan(plan, woff, first, seg_sizes, n_records):
    """Model of the fused kernel's walk: per warp, per chunk, lanes accumulate; a tile is finished at its last chunk (for a solo tile
    straight into the segment's record, otherwise as a partial that is folded when all parts have arrived); in a split chunk lanes
    [0, sp) finish the old tile and lanes [sp, n) start the tile of segment `part`.  Returns duos credited per segment."""
    covered = np.zeros(n_records, np.int32)
    credited, partial_seen = {}, {}
    for w in range(len(woff) - 1):
        acc, cur_seg = 0, None
        for j in range(woff[w], woff[w + 1]):
            db, meta, seg, part = (int(x) for x in plan[j])
            n = meta & COUNT
            assert 1 <= n <= 32
            covered[db:db + n] += 1
            if meta & SPLIT:
                sp = (meta >> SPLIT_SHIFT) & 63
                assert meta & LAST and meta & SOLO and 0 < sp < n, "a split chunk ends a solo tile and starts another"
                assert cur_seg == seg or (meta & FIRST and cur_seg is None)
                acc += sp
                credited[seg] = credited.get(seg, 0) + acc           # finish the old (solo) tile
                assert first[seg + 1] - first[seg] == 1 and first[part + 1] - first[part] == 1
                acc, cur_seg = n - sp, part
                if meta & NEWLAST:
                    credited[part] = credited.get(part, 0) + acc
                    acc, cur_seg = 0, None
                else:
                    assert acc < seg_sizes[part]
                continue
            assert not meta & NEWLAST
            if meta & FIRST:
                assert acc == 0 and cur_seg is None
                cur_seg = seg
            assert cur_seg == seg
            acc += n
            if meta & LAST:
                if meta & SOLO:
                    credited[seg] = credited.get(seg, 0) + acc
                else:                                                # partial record `part`, folded once per segment
                    assert first[seg] <= part < first[seg + 1] and part not in partial_seen
                    partial_seen[part] = acc
                acc, cur_seg = 0, None
        assert acc == 0 and cur_seg is None, "a warp's list ends on a tile boundary"
    for part, a in partial_seen.items():
        seg = int(np.searchsorted(first, part, side="right") - 1)
        credited[seg] = credited.get(seg, 0) + a
    assert (covered == 1).all()
    return credited


def test_split_chunks_actually_pack_small_problems():
    """on the bench shape (registrations of ~2.8 chunks) plan A needs clearly fewer chunk iterations than plan B"""
    rng = np.random.default_rng(1)
    sc = hostapi.build_schedule(offsets(rng.integers(60, 130, 16384)), MAX_WARPS)
    assert len(sc["plan_a"]) < 0.9 * len(sc["plan_b"])
    assert ((sc["plan_a"][:, 1] & SPLIT) != 0).sum() > 1000


def test_schedule_randomised_shapes():
    """hypothesis: arbitrary mixes of empty / tiny / chunk-aligned / multi-tile segments and any warp count keep every invariant"""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    sizes = st.lists(st.one_of(st.integers(0, 3), st.integers(28, 36), st.integers(60, 70), st.integers(250, 262), st.integers(0, 1500)), min_size=0, max_size=60)

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(sizes, st.integers(1, 64))
    def check(sz, max_warps):
        test_schedule_invariants("random", np.array(sz, np.int64), max_warps)
    check()
