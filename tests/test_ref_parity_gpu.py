"""The CUDA path (through the C-ABI) held to what the REFERENCE'S OWN compiled sources produced.

tests/golden/ref_golden.npz was written by oracle/_ref/libref.so — the reference's ceres_loss_functions.cpp, grid.cpp, radar_preprocessor.cpp,
ndt_cell.cpp and ndt_map.cpp compiled unmodified (tests/golden/gen_ref_golden.py) — on the seeded inputs of tests/golden/ref_cases.py.
Bars: cell statistics (float32), point counts, slot tables, neighbour lists and filtered points bit-exact; the robust loss as it enters
K3's normal equations 1e-12 relative (K3 evaluates rho through one reciprocal / rsqrt instead of pow: a few ulp).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_cases as RC  # noqa: E402

from randt_slam_b200 import capi  # noqa: E402
from tests import helpers as H  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def dense(sp, n):
    s = np.full(n, -1, np.int32)
    s[sp[:, 0]] = sp[:, 1]
    return s


def inputs(name):
    p, pf, pm = RC.scans(name)
    if RC.digest(pf) != str(G[name + "/sha_pf"]) or RC.digest(pm) != str(G[name + "/sha_pm"]):
        pytest.skip("the synthetic generator produces different points here than when the fixture was made")
    return p, pf, pm


@pytest.mark.parametrize("name", RC.PRESETS)
def test_voxelise_transform_merge_associate_equal_reference(gpu_ctx, name):
    p, pf, pm = inputs(name)
    gp = capi.grid_params(p)
    ns = p.size_x * p.size_y
    both = gpu_ctx.voxelize(np.concatenate([pf, pm]), [0, len(pf), len(pf) + len(pm)], gp)
    d = both.download()
    for b, tag in enumerate("fm"):
        a, z = d["cell_off"][b], d["cell_off"][b + 1]
        assert np.array_equal(bits(d["cells"][a:z]), bits(G["%s/cells_%s" % (name, tag)])), "Cell::updateCell differs from the reference"
        assert np.array_equal(d["npts"][a:z], G["%s/npts_%s" % (name, tag)])
        assert np.array_equal(d["slot"][b], dense(G["%s/slot_%s" % (name, tag)], ns))
    both.close()
    cf, nf, sf = G[name + "/cells_f"], G[name + "/npts_f"], dense(G[name + "/slot_f"], ns)
    cm, nm = G[name + "/cells_m"], G[name + "/npts_m"]
    for ai, theta in enumerate(RC.ANGLES):
        pose = RC.pose_for(theta)
        M = gpu_ctx.map_upload(cm, [0, len(cm)], gp, npts=nm)
        F = gpu_ctx.map_upload(cf, [0, len(cf)], gp, npts=nf, slot=sf[None])
        for metric in (capi.LOOKUP_MAHALANOBIS, capi.LOOKUP_EUCLID):
            prob = gpu_ctx.associate(F, M, pose[None], p.n_results_nn_lookup, metric)
            im, jf, _ = prob.download()
            want = G["%s/a%d/pairs_m%d" % (name, ai, metric)]
            assert np.array_equal(im, want[:, 0]) and np.array_equal(jf, want[:, 1]), "getClosestCells differs from the reference (angle %g)" % theta
            prob.close()
        M.transform_se2d(pose[None])
        assert np.array_equal(bits(M.download()["cells"]), bits(G["%s/a%d/cells_t" % (name, ai)])), "Cell::transformCell differs (angle %g)" % theta
        F.merge(M)
        g = F.download()
        assert np.array_equal(bits(g["cells"]), bits(G["%s/a%d/merged_cells" % (name, ai)])), "Cell::operator+= differs"
        assert np.array_equal(g["npts"], G["%s/a%d/merged_npts" % (name, ai)])
        assert np.array_equal(g["slot"][0], dense(G["%s/a%d/merged_slot" % (name, ai)], ns))
        F2 = gpu_ctx.map_upload(cf, [0, len(cf)], gp, npts=nf, slot=sf[None])
        F2.transform_se2d(pose[None])
        F.merge(F2)
        g = F.download()
        assert np.array_equal(bits(g["cells"]), bits(G["%s/a%d/merged2_cells" % (name, ai)])), "second merge differs"
        assert np.array_equal(g["npts"], G["%s/a%d/merged2_npts" % (name, ai)])
        for m_ in (M, F, F2):
            m_.close()


def test_rotation_of_the_float_affine_equals_reference(gpu_ctx):
    """Affine2f(pose.cast<float>().matrix()) and its Transform::rotation() for 512 headings, read back through one probe cell per pose:
    mean = affine * mean, covariance = R diag(1, 2, 3) R^T, both bit-exact against the reference's affine / rotation (fixture)."""
    th = RC.rot_angles()
    poses = np.stack([np.cos(th), np.sin(th), 0.25 * th, -0.5 * th], 1)
    B = len(poses)
    cell = np.zeros((B, 12), np.float32); cell[:, 0] = 1.5; cell[:, 1] = -2.25; cell[:, 2] = 90.0
    cell[:, 3] = 1.0; cell[:, 7] = 2.0; cell[:, 11] = 3.0
    from randt_slam_b200 import params as P
    gp = capi.grid_params(P.OXFORD)
    M = gpu_ctx.map_upload(cell, np.arange(B + 1, dtype=np.uint32), gp)
    M.transform_se2d(poses)
    got = M.download()["cells"]
    M.close()
    aff, R = G["rot_affine"], G["rot_R"]
    S = np.diag(np.array([1.0, 2.0, 3.0], np.float32))
    for b in range(B):
        c, s, tx, ty = aff[b]
        x, y, i = cell[b, :3]
        f32 = np.float32
        want_mu = np.array([tx + (c * x + ((-s) * y + f32(0) * i)), ty + (s * x + (c * y + f32(0) * i)), i], np.float32)
        T = np.zeros((3, 3), np.float32); O = np.zeros((3, 3), np.float32)
        for r in range(3):
            for q in range(3):
                T[r, q] = R[b][r, 0] * S[0, q] + (R[b][r, 1] * S[1, q] + R[b][r, 2] * S[2, q])
        for r in range(3):
            for q in range(3):
                O[r, q] = T[r, 0] * R[b][q, 0] + (T[r, 1] * R[b][q, 1] + T[r, 2] * R[b][q, 2])
        assert np.array_equal(bits(got[b, :3]), bits(want_mu)), th[b]
        assert np.array_equal(bits(got[b, 3:]), bits(O.reshape(9))), th[b]


@pytest.mark.parametrize("name", RC.PRESETS)
def test_filter_scan_equals_reference(gpu_ctx, name):
    p, raw, n_az, n_bins = RC.raw_scan(name)
    if RC.digest(raw) != str(G[name + "/sha_raw"]):
        pytest.skip("the synthetic generator produces a different raw scan here than when the fixture was made")
    for tag, tf in (("id", None), ("tf", G[name + "/filter_tfmat"])):
        kept = gpu_ctx.filter_scan(raw, n_az, n_bins, capi.filter_params(p, tf))
        want = G["%s/filter_%s" % (name, tag)]
        assert kept.shape == want.shape and np.array_equal(bits(kept), bits(want)), "filterScan differs from the reference (%s)" % tag


def test_loss_in_the_fused_normal_equations_equals_reference_loss(gpu_ctx):
    """One pair per segment: FUSED cost = weight rho(r^2) / 2, g = weight rho' r J^T, H = weight rho' J^T J (ceres' Corrector for rho'' <= 0) with
    rho, rho' from the reference's BarronLoss / WelschLoss (fixture), r and J from EMIT."""
    rng = np.random.default_rng(5)
    n = len(G["loss_s"])
    cm = H.random_cells(rng, n, extent=5.0); cf = H.random_cells(rng, n, extent=5.0)
    cf[:, :2] = cm[:, :2] + rng.normal(0, 0.4, (n, 2)).astype(np.float32); cf[:, 2] = cm[:, 2] + rng.normal(0, 5.0, n).astype(np.float32)
    idx = np.arange(n, dtype=np.uint32)
    prob = gpu_ctx.problem_create(cm, cf, idx, idx, np.arange(n + 1, dtype=np.uint32))
    pose = np.tile(np.array([np.cos(0.1), np.sin(0.1), 0.05, -0.02]), (n, 1))
    r, J = prob.eval_emit(pose)
    s = r * r
    # The fixture tabulates the reference's rho, rho' on its own s grid, not at our s = r^2: the formulas of ceres_loss_functions.cpp:19-39 are
    # restated inline below, pinned against the fixture on the fixture's grid first, and then evaluated at our s.
    def barron(a, alpha, mu, s_):
        b = mu * a * a; c = 1.0 / b
        if alpha >= 2.0:
            return s_, np.ones_like(s_)
        if abs(alpha) <= 0.05:
            sm = 1.0 + s_ * c
            return b * np.log(sm), np.maximum(np.finfo(np.float64).tiny, 1.0 / sm)
        f = abs(alpha - 2.0); e = 0.5 * alpha; pre = b * f / alpha; ts = 2 * c / f
        te = s_ * ts + 1.0
        return pre * (te ** e - 1.0), pre * e * te ** (e - 1.0) * ts

    for ci, (a, alpha, mu) in enumerate(G["loss_cases"]):
        rho_fix, drho_fix = barron(a, alpha, mu, G["loss_s"])
        ref = G["loss_barron"][ci]
        # (rho = pre (pow(..) - 1) cancels for tiny s: an ulp of pow is a large relative error of rho there, hence the absolute floor)
        assert np.allclose(rho_fix, ref[:, 0], rtol=1e-12, atol=1e-13 * a * a * mu) and np.allclose(drho_fix, ref[:, 1], rtol=1e-12, atol=0)
        rho, drho = barron(a, alpha, mu, s)
        for weight in (1.0, 0.37):
            out = capi.unpack_fused(prob.eval_fused(pose, capi.make_loss(capi.LOSS_BARRON, a, alpha, mu, weight)))
            assert H.rel_err(out["cost"], 0.5 * weight * rho) < 1e-12, (a, alpha, mu)
            want_g = (weight * drho * r)[:, None] * J
            want_H = (weight * drho)[:, None, None] * J[:, :, None] * J[:, None, :]
            assert H.rel_err(out["g"], want_g) < 1e-11, (a, alpha, mu)
            assert H.rel_err(out["H"], want_H) < 1e-11, (a, alpha, mu)
    prob.close()
