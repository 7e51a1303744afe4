"""Moving cells with a varying number of neighbours (1 ... 4, the k = 4 default of the indoor / outdoor / mixed presets): duos that are
full, half-empty, and cells that span two duos, in ragged segments, through EMIT, FUSED (all four variants) and the batched solver,
against the oracle."""
import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _mixed_problem(rng, sizes, n_f, max_k):
    im, jf, seg = [], [], [0]
    m = 0
    for sz in sizes:
        for _ in range(sz):
            k = int(rng.integers(1, max_k + 1))
            for j in rng.choice(n_f, k, replace=False):
                im.append(m); jf.append(int(j))
            m += 1
        seg.append(len(im))
    return np.array(im, np.uint32), np.array(jf, np.uint32), np.array(seg, np.uint32), m


@pytest.mark.parametrize("max_k", [3, 4])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_mixed_neighbour_counts_every_duo_pattern(oracle, gpu_ctx, max_k, variant):
    rng = np.random.default_rng(10 * max_k + variant)
    sizes = [1, 2, 15, 16, 17, 3, 31, 32, 33, 1, 1, 64, 0, 127, 128, 129, 300, 9, 48, 0, 350, 5]
    n_f = 700
    im, jf, seg, n_m = _mixed_problem(rng, sizes, n_f, max_k)
    assert max(np.bincount(im)) == max_k and min(np.bincount(im)) == 1
    cm = H.random_cells(rng, n_m, extent=6.0); cf = H.random_cells(rng, n_f, extent=6.0)
    S = len(sizes)
    npar = 4 if variant <= 1 else 3
    if variant <= 1:
        poses = np.stack([synth.pose_to_se2(*rng.uniform(-0.3, 0.3, 3)) for _ in range(S)])
    else:
        poses = rng.uniform(-0.3, 0.3, (S, 3))
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    # EMIT: every pair's residual and Jacobian row lands in its own row
    r, J = prob.eval_emit(poses, variant=variant)
    for s in range(S):
        a, z = seg[s], seg[s + 1]
        if z == a:
            continue
        ro, Jo = oracle.eval_pairs(variant, cm, cf, im[a:z], jf[a:z], poses[s], mode=0)
        assert np.max(np.abs(r[a:z] - ro) / (1 + np.abs(ro))) < 1e-10
        assert np.max(np.abs(J[a:z] - Jo[:, :npar]) / (1 + np.abs(Jo[:, :npar]))) < 1e-8
    # FUSED with a loss and per-segment mu
    mus = rng.uniform(1.0, 3.0, S)
    lt = (capi.LOSS_BARRON, 1.5, -1.0, 1.0, 0.05)
    got = prob.eval_fused(poses, capi.make_loss(*lt), mu_per_seg=mus, variant=variant)
    want, _ = oracle.fused_batch(variant, cm, cf, im, jf, seg, poses, lt, mu_per_seg=mus)
    sc = np.abs(want["H"]).max(axis=(1, 2), keepdims=True) + 1e-300
    assert np.max(np.abs(got[:, :16].reshape(-1, 4, 4) - want["H"]) / sc) < 1e-9
    assert np.max(np.abs(got[:, 16:20] - want["g"]) / (np.abs(want["g"]).max(axis=1, keepdims=True) + 1e-300)) < 1e-9
    assert np.allclose(got[:, capi.FUSED_COST], want["cost"], rtol=1e-10, atol=1e-300)
    assert np.allclose(got[:, capi.FUSED_MAXR], want["max_r"], rtol=1e-12, atol=0)
    assert np.array_equal(got[:, capi.FUSED_N], np.diff(seg).astype(np.float64))
    prob.close()


def test_k4_preset_from_associate_through_solver(oracle, gpu_ctx):
    """k = 4 (the indoor / outdoor / mixed presets): device-built duo records, registration against the oracle's restatement"""
    p = P.OUTDOOR
    k = p.n_results_nn_lookup
    assert k == 4
    case = H.make_registration_case(oracle, p, seed=5)
    gp = capi.grid_params(p)
    f, mv = case["fixed"], case["moving"]
    F = gpu_ctx.map_upload(f["cells"], [0, len(f["cells"])], gp, npts=f["npts"], slot=f["slot"][None])
    M = gpu_ctx.map_upload(mv["cells"], [0, len(mv["cells"])], gp)
    prob = gpu_ctx.associate(F, M, case["pose0"][None], k)
    pm, pf, seg = prob.download()
    assert np.array_equal(pm, case["im"]) and np.array_equal(pf, case["jf"])
    assert np.bincount(pm).max() > 2                                     # cells spanning two duos
    lt = (capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0)
    got = prob.eval_fused(case["pose0"][None], capi.make_loss(*lt))
    want = oracle.fused(0, mv["cells"], f["cells"], pm, pf, case["pose0"], lt)
    assert np.max(np.abs(got[0, :16].reshape(4, 4) - want["H"])) / np.abs(want["H"]).max() < 1e-10
    assert abs(got[0, capi.FUSED_COST] - want["cost"]) <= 1e-11 * abs(want["cost"])
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.gnc_steps, max_num_iterations=p.max_iteration)
    pose, res = prob.register_batch(case["pose0"][None], capi.make_loss(*lt), opt)
    o = oracle.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, mv["cells"], case["pose0"], k,
                               matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                               divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True)
    assert np.max(np.abs(pose[0] - o["pose"])) < 1e-7
    assert res[0, capi.REG_ITERATIONS] == o["iterations"]
    F.close(); M.close(); prob.close()
