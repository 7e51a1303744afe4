"""Pins the CPU oracle to the REFERENCE'S OWN code.

oracle/_ref/libref.so is the reference's ceres_loss_functions.cpp, grid.cpp, radar_preprocessor.cpp, ndt_cell.cpp and ndt_map.cpp compiled
unmodified from /root/reference (oracle/Makefile `ref`, shim headers under oracle/shim/); tests/golden/ref_golden.npz holds what it
produced on the seeded inputs of tests/golden/ref_cases.py (generator: tests/golden/gen_ref_golden.py).  Bars: labels, counts, slot
tables, neighbour lists bit-exact; float32 cell statistics bit-exact; the loss (rho, rho', rho'') bit-exact (same libm, same formula).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_cases as RC  # noqa: E402

from oracle import ref_py as R  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))
LIVE = R.available()
live = pytest.mark.skipif(not LIVE, reason="oracle/_ref/libref.so is not built and /root/reference is not here")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64) if a.dtype == np.float64 else a


def dense(sp, n):
    s = np.full(n, -1, np.int32)
    s[sp[:, 0]] = sp[:, 1]
    return s


def inputs(name):
    p, pf, pm = RC.scans(name)
    if RC.digest(pf) != str(G[name + "/sha_pf"]) or RC.digest(pm) != str(G[name + "/sha_pm"]):
        pytest.skip("the synthetic generator produces different points here than when the fixture was made")
    return p, pf, pm


@live
def test_ref_library_builds_from_the_reference_tree():
    assert R.build() and os.path.exists(R._LIB_PATH)
    assert R.lib().ref_version() == 1


# ---- a11 ---------------------------------------------------------------------------------------------------------------------
def test_oracle_loss_equals_reference_loss(oracle):
    O = oracle
    s = G["loss_s"]
    for ci, (a, al, mu) in enumerate(G["loss_cases"]):
        got = np.array([O.loss_eval(O.LOSS_BARRON, a, al, mu, 1.0, x) for x in s])
        assert np.array_equal(bits(got), bits(G["loss_barron"][ci])), (a, al, mu)
        got = np.array([O.loss_eval(O.LOSS_BARRON, a, al, 1.0, 1.0, x) for x in s])
        assert np.array_equal(bits(got), bits(G["loss_barron_nomu"][ci])), (a, al)
        got = np.array([O.loss_eval(O.LOSS_WELSCH, a, -2.0, mu, 1.0, x) for x in s])
        assert np.array_equal(bits(got), bits(G["loss_welsch"][ci])), (a, mu)


@live
def test_fixture_loss_is_what_the_reference_library_returns():
    for ci, (a, al, mu) in enumerate(G["loss_cases"]):
        assert np.array_equal(bits(R.barron(a, al, mu, G["loss_s"])), bits(G["loss_barron"][ci]))
        assert np.array_equal(bits(R.welsch(a, mu, G["loss_s"])), bits(G["loss_welsch"][ci]))


# ---- Sophus cast + Eigen rotation() ----------------------------------------------------------------------------------------
def test_oracle_affine_and_rotation_equal_reference(oracle):
    th = RC.rot_angles()
    poses = np.stack([np.cos(th), np.sin(th), 0.25 * th, -0.5 * th], 1)
    differs = 0
    for i, q in enumerate(poses):
        a = oracle.se2d_cast_float(q)
        assert np.array_equal(bits(a), bits(G["rot_affine"][i]))
        Rm = oracle.affine_rotation(a[0], a[1])
        assert np.array_equal(bits(Rm), bits(G["rot_R"][i])), th[i]
        differs += not np.array_equal(Rm, np.array([[a[0], -a[1], 0], [a[1], a[0], 0], [0, 0, 1]], np.float32))
    assert differs > 100      # the polar factor really is not the raw linear part (what round 1 assumed)


# ---- a1-a7 -------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RC.PRESETS)
def test_oracle_voxelise_transform_merge_associate_equal_reference(oracle, name):
    O = oracle
    p, pf, pm = inputs(name)
    va = RC.vox_args(p)
    assert np.array_equal(O.grid_labels(pf, p.n_clusters, p.max_range), G[name + "/labels_f"])
    vf, vm = O.voxelize(pf, *va), O.voxelize(pm, *va)
    for tag, v in (("f", vf), ("m", vm)):
        assert np.array_equal(bits(v["cells"]), bits(G["%s/cells_%s" % (name, tag)]))
        assert np.array_equal(v["npts"], G["%s/npts_%s" % (name, tag)])
        assert np.array_equal(v["slot"], dense(G["%s/slot_%s" % (name, tag)], p.size_x * p.size_y))
    assert len(vf["cells"]) > 20
    for ai, theta in enumerate(RC.ANGLES):
        pose = RC.pose_for(theta)
        a = O.se2d_cast_float(pose)
        mt = O.transform_cells(vm["cells"], *a)
        assert np.array_equal(bits(mt), bits(G["%s/a%d/cells_t" % (name, ai)]))
        mc, mn, ms = O.merge_map_cell(vf["cells"], vf["npts"], vf["slot"], p.size_x, p.size_y, p.resolution, mt, vm["npts"])
        assert np.array_equal(bits(mc), bits(G["%s/a%d/merged_cells" % (name, ai)]))
        assert np.array_equal(mn, G["%s/a%d/merged_npts" % (name, ai)])
        assert np.array_equal(ms, dense(G["%s/a%d/merged_slot" % (name, ai)], p.size_x * p.size_y))
        ft = O.transform_cells(vf["cells"], *a)
        mc2, mn2, _ = O.merge_map_cell(mc, mn, ms, p.size_x, p.size_y, p.resolution, ft, vf["npts"])
        assert np.array_equal(bits(mc2), bits(G["%s/a%d/merged2_cells" % (name, ai)]))
        assert np.array_equal(mn2, G["%s/a%d/merged2_npts" % (name, ai)])
        for metric in (0, 1):
            im, jf = O.associate(vf["cells"], vf["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, vm["cells"], pose,
                                 p.n_results_nn_lookup, metric)
            want = G["%s/a%d/pairs_m%d" % (name, ai, metric)]
            assert np.array_equal(im, want[:, 0]) and np.array_equal(jf, want[:, 1])


@live
@pytest.mark.parametrize("name", RC.PRESETS)
def test_fixture_cells_are_what_the_reference_library_returns(name):
    p, pf, pm = inputs(name)
    F = R.RefMap.from_scan(pf, *RC.vox_args(p))
    g = F.get()
    assert np.array_equal(bits(g["cells"]), bits(G[name + "/cells_f"])) and np.array_equal(g["npts"], G[name + "/npts_f"])
    M = R.RefMap.from_scan(pm, *RC.vox_args(p))
    pose = RC.pose_for(RC.ANGLES[1])
    im, jf = F.associate(M, pose, p.n_results_nn_lookup, 0)
    assert np.array_equal(np.stack([im, jf], 1), G["%s/a1/pairs_m0" % name])


# ---- f1 ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RC.PRESETS)
def test_oracle_filter_scan_equals_reference(oracle, name):
    p, raw, n_az, n_bins = RC.raw_scan(name)
    if RC.digest(raw) != str(G[name + "/sha_raw"]):
        pytest.skip("the synthetic generator produces a different raw scan here than when the fixture was made")
    for tag, tf in (("id", None), ("tf", G[name + "/filter_tfmat"])):
        kept, npk = oracle.filter_scan(raw, p.min_range, p.max_range, p.min_intensity, p.beam_distance_increment_threshold, tf)
        want = G["%s/filter_%s" % (name, tag)]
        assert kept.shape == want.shape and np.array_equal(bits(kept), bits(want))
        assert npk == int(G["%s/filter_%s_npeaks" % (name, tag)])
        assert len(kept) > 100


# ---- live sweep: more seeds / angles than the fixture holds -------------------------------------------------------------------
@live
@pytest.mark.parametrize("name", RC.PRESETS)
def test_oracle_equals_reference_library_on_fresh_seeds(oracle, name):
    from randt_slam_b200 import params as P, synth
    O = oracle
    p = P.PRESETS[name]
    va = RC.vox_args(p)
    rng = np.random.default_rng(1234)
    for seed in (11, 12):
        sc = synth.scene_for(p, seed); kw = synth.preset_scan_kwargs(p); kw["n_azimuth"] = 160
        pf = synth.make_scan(sc, (0.0, 0.0, 0.0), p, seed, **kw)
        F = R.RefMap.from_scan(pf, *va); vf = O.voxelize(pf, *va)
        cur = dict(cells=vf["cells"], npts=vf["npts"], slot=vf["slot"])
        for step in range(3):          # a small keyframe chain: voxelise, transform, associate, merge
            th = float(rng.uniform(-3.1, 3.1)); pose = synth.pose_to_se2(float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)), th)
            pm = synth.make_scan(sc, (0.4 * step, 0.1, 0.02 * step), p, 100 * seed + step, **kw)
            M = R.RefMap.from_scan(pm, *va); vm = O.voxelize(pm, *va)
            assert np.array_equal(bits(M.get()["cells"]), bits(vm["cells"]))
            for metric in (0, 1):
                im, jf = F.associate(M, pose, p.n_results_nn_lookup, metric)
                io, jo = O.associate(cur["cells"], cur["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, vm["cells"], pose,
                                     p.n_results_nn_lookup, metric)
                assert np.array_equal(im, io) and np.array_equal(jf, jo)
            M.transform(pose); F.merge(M)
            mt = O.transform_cells(vm["cells"], *O.se2d_cast_float(pose))
            c, n, s = O.merge_map_cell(cur["cells"], cur["npts"], cur["slot"], p.size_x, p.size_y, p.resolution, mt, vm["npts"])
            cur = dict(cells=c, npts=n, slot=s)
            g = F.get()
            assert np.array_equal(bits(g["cells"]), bits(c)) and np.array_equal(g["npts"], n) and np.array_equal(g["slot"], s)
