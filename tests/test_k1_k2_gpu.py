"""K1 (voxelise), K2 (associate) and map maintenance parity: GPU through the C-ABI vs the CPU oracle.

Index work (labels, point counts, slot tables, neighbour lists, pair order) must be identical.  Cell statistics are float32 in
both paths with the same operation order (K1/K2 are built with -fmad=false), so they are asserted bit-identical as well;
the documented tolerance for the float path is 1e-5 relative (north_star), the observed difference is 0 ulp.
"""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

PRESETS = {"oxford": P.OXFORD, "c1": P.C1, "indoor": P.INDOOR, "outdoor": P.OUTDOOR}


def batch_scans(p, seeds):
    scans = [H.make_scan(p, s, (0.3 * i, -0.2 * i, 0.01 * i), 50 + s) for i, s in enumerate(seeds)]
    scans.insert(1, np.zeros((0, 4), np.float32))                       # empty scan
    scans.insert(3, scans[0][: p.min_points_per_cell].copy())          # too few points for any cell
    off = np.zeros(len(scans) + 1, np.uint32)
    off[1:] = np.cumsum([len(s) for s in scans])
    return scans, off, np.concatenate(scans, 0)


@pytest.mark.parametrize("preset", list(PRESETS))
def test_voxelize_matches_oracle(oracle, gpu_ctx, preset):
    p = PRESETS[preset]
    scans, off, pts = batch_scans(p, [1, 2, 3])
    m = gpu_ctx.voxelize(pts, off, capi.grid_params(p))
    d = m.download(want_labels=True)
    assert len(d["cell_off"]) == len(scans) + 1
    total = 0
    for b, sc in enumerate(scans):
        v = oracle.voxelize(sc, *H.vox_args(p))
        a, z = d["cell_off"][b], d["cell_off"][b + 1]
        assert z - a == len(v["cells"]), "cell count differs for scan %d" % b
        assert np.array_equal(d["labels"][a:z], v["labels"])
        assert np.array_equal(d["npts"][a:z], v["npts"])
        assert np.array_equal(d["slot"][b], v["slot"])
        if z > a:
            assert H.rel_err(d["cells"][a:z, :3], v["cells"][:, :3]) < 1e-5
            assert np.array_equal(d["cells"][a:z].view(np.uint32), v["cells"].view(np.uint32)), "cell statistics not bit-identical"
        total += z - a
    assert total > 100


def test_voxelize_device_pointer_and_large_batch(oracle, gpu_ctx):
    p = P.OXFORD
    base = [H.make_scan(p, 4, (0.1 * i, 0.0, 0.0), 300 + i) for i in range(4)]
    scans = [base[i % 4] for i in range(64)]
    off = np.zeros(65, np.uint32); off[1:] = np.cumsum([len(s) for s in scans])
    pts = np.concatenate(scans, 0)
    m = gpu_ctx.voxelize(pts, off, capi.grid_params(p))
    d = m.download()
    ref = [oracle.voxelize(s, *H.vox_args(p)) for s in base]
    for b in range(64):
        a, z = d["cell_off"][b], d["cell_off"][b + 1]
        assert np.array_equal(d["cells"][a:z].view(np.uint32), ref[b % 4]["cells"].view(np.uint32))


def _same_cells(d, b, v):
    a, z = d["cell_off"][b], d["cell_off"][b + 1]
    assert z - a == len(v["cells"])
    assert np.array_equal(d["labels"][a:z], v["labels"]) and np.array_equal(d["npts"][a:z], v["npts"])
    assert np.array_equal(d["slot"][b], v["slot"])
    assert np.array_equal(d["cells"][a:z].view(np.uint32), v["cells"].view(np.uint32))


@pytest.mark.parametrize("n_scans", [1, 200])       # a lone scan (staged in shared memory) and a batch (two scans per SM)
def test_voxelize_any_point_order_and_extreme_cells(oracle, gpu_ctx, n_scans):
    """K1 sorts cell by cell along the parts of the scan a cell occurs in; an azimuth-major scan keeps those walks short, but nothing may
    depend on it.  Shuffled points (every cell occurs in every part), a scan that is one single cell, the largest scan the call
    accepts (16 384 points), and a cell that straddles the end of the scan — all bit-identical to the oracle's sequential sums."""
    p = P.OXFORD
    rng = np.random.default_rng(9)
    base = H.make_scan(p, 6, (0.2, 0.1, 0.05), 77)
    shuffled = base[rng.permutation(len(base))]
    one_cell = np.zeros((3000, 4), np.float32)
    one_cell[:, 0] = 10.0 + rng.uniform(0, 1.0, 3000); one_cell[:, 1] = 20.0 + rng.uniform(0, 1.0, 3000); one_cell[:, 3] = rng.uniform(80, 200, 3000)
    big = np.concatenate([base, base[::-1], shuffled, base])[:16384].copy()
    big[:, :2] += rng.normal(0, 0.02, (len(big), 2)).astype(np.float32)
    wrap = np.concatenate([base[len(base) // 2:], base[:len(base) // 2]])       # the scan starts in the middle of the sweep
    cases = [shuffled, one_cell, big, wrap]
    scans = [cases[i % len(cases)] for i in range(max(n_scans, len(cases)))] if n_scans > 1 else None
    gp = capi.grid_params(p)
    ref = [oracle.voxelize(c, *H.vox_args(p)) for c in cases]
    assert len(ref[1]["cells"]) >= 1 and ref[1]["npts"].max() > 1000 and len(big) == 16384
    if n_scans == 1:
        for c, v in zip(cases, ref):
            m = gpu_ctx.voxelize(c, [0, len(c)], gp)
            _same_cells(m.download(want_labels=True), 0, v)
            m.close()
    else:
        off = np.zeros(len(scans) + 1, np.uint32); off[1:] = np.cumsum([len(x) for x in scans])
        m = gpu_ctx.voxelize(np.concatenate(scans), off, gp)
        d = m.download(want_labels=True)
        for b in range(len(scans)):
            _same_cells(d, b, ref[b % len(cases)])
        m.close()


def test_voxelize_rejects_far_points(gpu_ctx):
    p = P.OXFORD
    pts = np.zeros((50, 4), np.float32); pts[:, 0] = 1e6; pts[:, 3] = 90
    pts[:25, 0] = -1e6
    with pytest.raises(capi.RandtError) as e:
        gpu_ctx.voxelize(pts, [0, 50], capi.grid_params(p))
    assert e.value.code in (capi.E_CAPACITY, capi.E_INVALID)


@pytest.mark.parametrize("preset", ["oxford", "c1", "indoor"])
@pytest.mark.parametrize("metric", [capi.LOOKUP_MAHALANOBIS, capi.LOOKUP_EUCLID])
def test_associate_matches_oracle(oracle, gpu_ctx, preset, metric):
    p = PRESETS[preset]
    B = 3
    fixed, moving, poses = [], [], []
    for b in range(B):
        f = H.build_submap(oracle, p, 20 + b, n_scans=3)
        pts = H.make_scan(p, 20 + b, (0.5, -0.3, 0.02), 700 + b)
        mv = oracle.voxelize(pts, *H.vox_args(p))
        fixed.append(f); moving.append(mv)
        poses.append(synth.pose_to_se2(0.45 + 0.1 * b, -0.25, 0.015 * (b + 1)))
    gp = capi.grid_params(p)
    f_off = np.concatenate([[0], np.cumsum([len(f["cells"]) for f in fixed])]).astype(np.uint32)
    m_off = np.concatenate([[0], np.cumsum([len(m["cells"]) for m in moving])]).astype(np.uint32)
    fm = gpu_ctx.map_upload(np.concatenate([f["cells"] for f in fixed]), f_off, gp, npts=np.concatenate([f["npts"] for f in fixed]),
                            slot=np.stack([f["slot"] for f in fixed]))
    mm = gpu_ctx.map_upload(np.concatenate([m["cells"] for m in moving]), m_off, gp)
    prob = gpu_ctx.associate(fm, mm, np.stack(poses), p.n_results_nn_lookup, metric)
    pm, pf, seg = prob.download()
    assert prob.n_segments == B
    for b in range(B):
        im, jf = oracle.associate(fixed[b]["cells"], fixed[b]["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance,
                                  moving[b]["cells"], poses[b], p.n_results_nn_lookup, metric)
        a, z = seg[b], seg[b + 1]
        assert z - a == len(im), "pair count differs in problem %d" % b
        assert np.array_equal(pm[a:z] - m_off[b], im)
        assert np.array_equal(pf[a:z] - f_off[b], jf)
        assert len(im) > 20
    # the problem snapshots its cell tables
    cm, cf = prob.download_cells()
    assert np.array_equal(cm, np.concatenate([m["cells"] for m in moving]))


@pytest.mark.parametrize("preset,n_submap_scans", [("oxford", 3), ("indoor", 2), ("c1", 6)])
@pytest.mark.parametrize("metric", [capi.LOOKUP_MAHALANOBIS, capi.LOOKUP_EUCLID])
def test_associate_single_map_equals_batched_path(oracle, gpu_ctx, preset, n_submap_scans, metric):
    """One scan against one submap takes the fused one-launch path (association, pair / duo lists, K3's records and the snapshots in
    one kernel).  Its problem must be the one the batched path builds for the same pair: same pairs as the oracle's
    getCellsAndNeighbors restatement, and equal fused blocks and bit-identical solver results."""
    p = PRESETS[preset]
    f = H.build_submap(oracle, p, 31, n_scans=n_submap_scans)
    pts = H.make_scan(p, 31, (0.4, -0.2, 0.03), 911)
    mv = oracle.voxelize(pts, *H.vox_args(p))
    pose = synth.pose_to_se2(0.38, -0.22, 0.025)
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    nf, nm = len(f["cells"]), len(mv["cells"])
    f1 = gpu_ctx.map_upload(f["cells"], np.array([0, nf], np.uint32), gp, npts=f["npts"], slot=f["slot"][None])
    m1 = gpu_ctx.map_upload(mv["cells"], np.array([0, nm], np.uint32), gp)
    single = gpu_ctx.associate(f1, m1, pose[None], k, metric)
    f2 = gpu_ctx.map_upload(np.concatenate([f["cells"]] * 2), np.array([0, nf, 2 * nf], np.uint32), gp, npts=np.concatenate([f["npts"]] * 2),
                            slot=np.stack([f["slot"]] * 2))
    m2 = gpu_ctx.map_upload(np.concatenate([mv["cells"]] * 2), np.array([0, nm, 2 * nm], np.uint32), gp)
    both = gpu_ctx.associate(f2, m2, np.stack([pose, pose]), k, metric)
    im, jf = oracle.associate(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, mv["cells"], pose, k, metric)
    pm, pf, seg = single.download()
    assert list(seg) == [0, len(im)] and len(im) > 20
    assert np.array_equal(pm, im) and np.array_equal(pf, jf)
    bm, bf, bseg = both.download()
    assert np.array_equal(bm[:bseg[1]], pm) and np.array_equal(bf[:bseg[1]], pf)
    cm, cf = single.download_cells()
    assert np.array_equal(cm, mv["cells"]) and np.array_equal(cf, f["cells"])
    assert single.layout() == (both.layout()[0] // 2, both.layout()[1], both.layout()[2] // 2)
    loss = capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0)
    for q in (pose, synth.pose_to_se2(0.5, -0.1, -0.04)):
        a = single.eval_fused(q[None], loss)
        b = both.eval_fused(np.stack([q, q]), loss)      # (tile sizes differ between the two schedules: same sums, other grouping)
        assert np.allclose(a[0], b[0], rtol=1e-11, atol=1e-13) and np.array_equal(b[0], b[1])
    opt = capi.solver_options()
    ra = single.register_batch(pose[None], loss, opt)
    rb = both.register_batch(np.stack([pose, pose]), loss, opt)
    assert np.array_equal(np.asarray(ra[0])[0], np.asarray(rb[0])[0])


@pytest.mark.parametrize("metric", [capi.LOOKUP_MAHALANOBIS, capi.LOOKUP_EUCLID])
def test_associate_single_map_with_thousands_of_cells(oracle, gpu_ctx, metric):
    """the one-CTA association walks a large moving map 1 024 cells at a time, carrying the pair / duo offsets from round to round:
    3 000 moving against 6 000 fixed cells (drawn directly, slot table built on upload), against the oracle and the batched path"""
    p = P.INDOOR
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    rng = np.random.default_rng(77)
    ext = 0.45 * p.size_x * p.resolution
    cf = H.random_cells(rng, 6000, extent=ext); cm = H.random_cells(rng, 3000, extent=ext)
    pose = synth.pose_to_se2(0.3, -0.2, 0.05)
    F1 = gpu_ctx.map_upload(cf, np.array([0, 6000], np.uint32), gp)
    M1 = gpu_ctx.map_upload(cm, np.array([0, 3000], np.uint32), gp)
    slot = F1.download()["slot"][0]
    single = gpu_ctx.associate(F1, M1, pose[None], k, metric)
    pm, pf, seg = single.download()
    im, jf = oracle.associate(cf, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, cm, pose, k, metric)
    assert seg[1] == len(im) and len(im) > 3000
    assert np.array_equal(pm, im) and np.array_equal(pf, jf)
    F2 = gpu_ctx.map_upload(np.concatenate([cf, cf]), np.array([0, 6000, 12000], np.uint32), gp)
    M2 = gpu_ctx.map_upload(np.concatenate([cm, cm]), np.array([0, 3000, 6000], np.uint32), gp)
    both = gpu_ctx.associate(F2, M2, np.stack([pose, pose]), k, metric)
    bm, bf, bseg = both.download()
    assert np.array_equal(bm[:bseg[1]], pm) and np.array_equal(bf[:bseg[1]], pf)
    loss = capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0)
    a = single.eval_fused(pose[None], loss); b = both.eval_fused(np.stack([pose, pose]), loss)
    assert np.allclose(a[0], b[0], rtol=1e-11, atol=1e-13)
    assert single.layout()[0] * 2 == both.layout()[0]


def test_slot_table_rebuild_matches_insert_order(oracle, gpu_ctx):
    p = P.OXFORD
    v = oracle.voxelize(H.make_scan(p, 9, (0, 0, 0), 9), *H.vox_args(p))
    m = gpu_ctx.map_upload(v["cells"], [0, len(v["cells"])], capi.grid_params(p))   # slot=None -> rebuilt on device
    assert np.array_equal(m.download()["slot"][0], v["slot"])


def test_transform_and_merge_match_oracle(oracle, gpu_ctx):
    p = P.OXFORD
    gp = capi.grid_params(p)
    B = 2
    subs = [dict(cells=np.zeros((0, 12), np.float32), npts=np.zeros(0, np.uint32), slot=np.full(p.size_x * p.size_y, -1, np.int32)) for _ in range(B)]
    gm = gpu_ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(B + 1, np.uint32), gp)
    for step in range(4):
        scans, trans = [], []
        for b in range(B):
            pose = (0.5 * step, 0.1 * b * step, 0.01 * step)
            scans.append(H.make_scan(p, 30 + b, pose, 40 + 10 * b + step))
            trans.append([math.cos(pose[2]), math.sin(pose[2]), pose[0], pose[1]])
        off = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint32)
        mv = gpu_ctx.voxelize(np.concatenate(scans), off, gp)
        mv.transform(np.array(trans, np.float32))
        dmv = mv.download()
        gm.merge(mv)
        d = gm.download()
        for b in range(B):
            v = oracle.voxelize(scans[b], *H.vox_args(p))
            t = np.array(trans[b], np.float32)
            mc = oracle.transform_cells(v["cells"], *t)
            a, z = dmv["cell_off"][b], dmv["cell_off"][b + 1]
            assert np.array_equal(dmv["cells"][a:z].view(np.uint32), mc.view(np.uint32)), "transformCell differs"
            c, n, s = oracle.merge_map_cell(subs[b]["cells"], subs[b]["npts"], subs[b]["slot"], p.size_x, p.size_y, p.resolution, mc, v["npts"])
            subs[b] = dict(cells=c, npts=n, slot=s)
            a, z = d["cell_off"][b], d["cell_off"][b + 1]
            assert z - a == len(c)
            assert np.array_equal(d["npts"][a:z], n)
            assert np.array_equal(d["slot"][b], s)
            assert np.array_equal(d["cells"][a:z].view(np.uint32), c.view(np.uint32)), "mergeMapCell differs at step %d" % step
    assert len(subs[0]["cells"]) > 60


def test_cs_divergence_matches_oracle(oracle, gpu_ctx):
    """Map::calculateCSDivergence (ndt_map.cpp:42-99) for a batch of (submap, transformed scan) pairs; fp64 sums of float32 terms:
    asserted 1e-9 absolute on the divergence (a log of sums ~1e2)"""
    p = P.OXFORD
    gp = capi.grid_params(p)
    B = 3
    fixed = [H.build_submap(oracle, p, 50 + b, n_scans=3) for b in range(B)]
    moving, trans = [], []
    for b in range(B):
        pts = H.make_scan(p, 50 + b, (0.5, -0.3, 0.02), 800 + b)
        mv = oracle.voxelize(pts, *H.vox_args(p))
        t = np.array([math.cos(0.02 + 0.01 * b), math.sin(0.02 + 0.01 * b), 0.5, -0.3 + 0.1 * b], np.float32)
        moving.append(oracle.transform_cells(mv["cells"], *t)); trans.append(t)
    f_off = np.concatenate([[0], np.cumsum([len(f["cells"]) for f in fixed])]).astype(np.uint32)
    m_off = np.concatenate([[0], np.cumsum([len(m) for m in moving])]).astype(np.uint32)
    fm = gpu_ctx.map_upload(np.concatenate([f["cells"] for f in fixed]), f_off, gp)
    mm = gpu_ctx.map_upload(np.concatenate(moving), m_off, gp)
    cs = fm.cs_divergence(mm)
    cs2 = fm.cs_divergence(mm)
    assert np.array_equal(cs, cs2), "deterministic reduction"
    for b in range(B):
        want, terms = oracle.cs_divergence(fixed[b]["cells"], moving[b])
        assert np.isfinite(want) and np.all(terms > 0)
        assert abs(cs[b] - want) < 1e-9, (b, cs[b], want)
    # a far-away scan overlaps less: larger divergence
    far = oracle.transform_cells(moving[0], 1.0, 0.0, 30.0, 0.0)
    mm2 = gpu_ctx.map_upload(far, [0, len(far)], gp)
    fm0 = gpu_ctx.map_upload(fixed[0]["cells"], [0, len(fixed[0]["cells"])], gp)
    assert fm0.cs_divergence(mm2)[0] > cs[0] + 1.0
