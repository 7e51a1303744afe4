"""Batched GNC + LM registration (K3 + K4 through randt_register_batch) vs the CPU oracle's restatement of
Matcher::estimateLoopConstraint (R/src/ndt_registration/ndt_matcher.cpp:426-493) on identical pair lists.

Tolerance: the two sides run the same minimiser on evaluations that agree to ~1e-13, so the accepted-step sequences coincide;
poses are asserted to 1e-7 absolute (north_star: 1e-5 relative), the score to 1e-7 relative, and iteration / solve counts equal.
"""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def build_batch(oracle, p, seeds, guesses):
    cases = []
    for sd, g in zip(seeds, guesses):
        cases.append(H.make_registration_case(oracle, p, seed=sd, n_fixed_scans=4, true_pose=(0.6, -0.4, 0.03), guess=g))
    cm = np.concatenate([c["moving"]["cells"] for c in cases]); cf = np.concatenate([c["fixed"]["cells"] for c in cases])
    om = np.cumsum([0] + [len(c["moving"]["cells"]) for c in cases]); of = np.cumsum([0] + [len(c["fixed"]["cells"]) for c in cases])
    pm = np.concatenate([c["im"] + om[i] for i, c in enumerate(cases)]).astype(np.uint32)
    pf = np.concatenate([c["jf"] + of[i] for i, c in enumerate(cases)]).astype(np.uint32)
    seg = np.cumsum([0] + [len(c["im"]) for c in cases]).astype(np.uint32)
    poses = np.stack([c["pose0"] for c in cases])
    return cases, cm, cf, pm, pf, seg, poses


def oracle_solve(oracle, p, c, on_manifold, loop_scale, weight=1.0):
    f = c["fixed"]
    return oracle.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, c["moving"]["cells"],
                                  c["pose0"], p.n_results_nn_lookup, matcher_loss_scale=p.loss_function_scale, loop_scale=loop_scale,
                                  alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.loop_closure_gnc_steps,
                                  on_manifold=on_manifold, loss_weight=weight, pairs=(c["im"], c["jf"]))


@pytest.mark.parametrize("preset", ["oxford", "c1"])
@pytest.mark.parametrize("on_manifold", [False, True])
def test_register_batch_matches_oracle(oracle, gpu_ctx, preset, on_manifold):
    p = {"oxford": P.OXFORD, "c1": P.C1}[preset]
    guesses = [(0.5, -0.3, 0.02), (0.2, -0.1, 0.0), (0.9, -0.6, 0.05), (0.6, -0.4, 0.03), (0.0, 0.0, 0.0)]
    cases, cm, cf, pm, pf, seg, poses = build_batch(oracle, p, [3, 4, 5, 6, 7], guesses)
    # an extra segment without any pair in the middle of the batch ("NO RESIDUALS ADDED")
    seg2 = np.concatenate([seg[:3], [seg[2]], seg[3:]]).astype(np.uint32)
    poses2 = np.concatenate([poses[:2], poses[:1], poses[2:]])
    prob = gpu_ctx.problem_create(cm, cf, pm, pf, seg2)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)
    opt = capi.solver_options(use_manifold=int(on_manifold), gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.loop_closure_gnc_steps, max_num_iterations=p.max_iteration)
    out, res = prob.register_batch(poses2, loss, opt)
    assert res[2, capi.REG_STATUS] == 1 and np.array_equal(out[2], poses2[2])
    idx = [0, 1, 3, 4, 5]
    for k, c in zip(idx, cases):
        o = oracle_solve(oracle, p, c, on_manifold, p.loop_closure_scale)
        assert o["status"] == 0 and res[k, capi.REG_STATUS] == 0
        assert int(res[k, capi.REG_GNC_SOLVES]) == o["gnc_solves"]
        assert abs(res[k, capi.REG_MU_FIRST] - o["mu_first"]) <= 1e-9 * abs(o["mu_first"])
        if on_manifold:
            # well-posed problem (3 tangent unknowns): identical accepted-step sequence, poses to 1e-7
            assert np.max(np.abs(out[k] - o["pose"])) < 1e-7, (k, out[k], o["pose"])
            assert abs(res[k, capi.REG_SCORE] - o["score"]) <= 1e-7 * abs(o["score"])
            assert int(res[k, capi.REG_ITERATIONS]) == o["iterations"]
            assert int(res[k, capi.REG_EVALS]) == o["evals"]
            assert abs(math.hypot(out[k, 0], out[k, 1]) - 1.0) < 1e-12
        else:
            # raw ambient parameters: |(c, s)| is a gauge direction with a singular J^T J; its LM step is rounding noise
            # (reference included), so only the gauge-invariant pose is comparable, to the accuracy ceres' function tolerance leaves
            th, tho = math.atan2(out[k, 1], out[k, 0]), math.atan2(o["pose"][1], o["pose"][0])
            assert abs(th - tho) < 2e-3 and np.max(np.abs(out[k, 2:] - o["pose"][2:])) < 5e-3
            assert abs(res[k, capi.REG_SCORE] - o["score"]) <= 1e-3 * abs(o["score"])
            assert abs(int(res[k, capi.REG_ITERATIONS]) - o["iterations"]) <= 0.25 * o["iterations"]
    # converges towards the true pose from every guess
    th = np.arctan2(out[idx, 1], out[idx, 0])
    assert np.all(np.abs(out[idx, 2] - 0.6) < 0.25) and np.all(np.abs(out[idx, 3] + 0.4) < 0.25) and np.all(np.abs(th - 0.03) < 0.02)


def test_register_batch_is_deterministic_and_independent_of_batching(oracle, gpu_ctx):
    """a registration gives the same bits alone and inside a batch (segments never interact)"""
    p = P.OXFORD
    guesses = [(0.5, -0.3, 0.02), (0.2, -0.1, 0.0), (0.9, -0.6, 0.05)]
    cases, cm, cf, pm, pf, seg, poses = build_batch(oracle, p, [11, 12, 13], guesses)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)
    opt = capi.solver_options(gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=p.loop_closure_gnc_steps)
    prob = gpu_ctx.problem_create(cm, cf, pm, pf, seg)
    out1, res1 = prob.register_batch(poses, loss, opt)
    out2, res2 = prob.register_batch(poses, loss, opt)
    assert np.array_equal(out1, out2) and np.array_equal(res1, res2)
    c = cases[1]
    solo = gpu_ctx.problem_create(c["moving"]["cells"], c["fixed"]["cells"], c["im"], c["jf"], [0, len(c["im"])])
    o, r = solo.register_batch(c["pose0"][None], loss, opt)
    assert np.array_equal(o[0], out1[1]) and np.array_equal(r[0], res1[1])


def test_register_batch_vector_variant_and_scaled_loss(oracle, gpu_ctx):
    """odometry-style loss (ScaledLoss weight ndt_weight / (n_cells k), ndt_matcher.cpp:392) and the (x, y | theta) functor"""
    p = P.OXFORD
    c = H.make_registration_case(oracle, p, seed=21, n_fixed_scans=4, true_pose=(0.6, -0.4, 0.03), guess=(0.4, -0.2, 0.01))
    w = p.ndt_weight / (len(c["moving"]["cells"]) * p.n_results_nn_lookup)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, w)
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=p.gnc_steps)
    prob = gpu_ctx.problem_create(c["moving"]["cells"], c["fixed"]["cells"], c["im"], c["jf"], [0, len(c["im"])])
    out, res = prob.register_batch(c["pose0"][None], loss, opt)
    f = c["fixed"]
    o = oracle.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, c["moving"]["cells"], c["pose0"],
                               p.n_results_nn_lookup, matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale,
                               alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True,
                               loss_weight=w, pairs=(c["im"], c["jf"]))
    assert np.max(np.abs(out[0] - o["pose"])) < 1e-7 and int(res[0, capi.REG_ITERATIONS]) == o["iterations"]
    # vector functor: parameters (x, y, theta)
    x0 = np.array([[0.4, -0.2, 0.01]])
    out3, res3 = prob.register_batch(x0, loss, opt, variant=capi.VAR_VEC_INTENSITY)
    o3 = oracle.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, c["moving"]["cells"], np.append(x0[0], 0.0),
                                p.n_results_nn_lookup, variant=2, matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale,
                                alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=False,
                                loss_weight=w, pairs=(c["im"], c["jf"]))
    assert np.max(np.abs(out3[0] - o3["pose"][:3])) < 1e-7
    assert abs(out3[0, 2] - math.atan2(out[0, 1], out[0, 0])) < 1e-3


def test_weighted_batch_equals_single_registrations(oracle, gpu_ctx):
    """the odometry weight ndt_weight / (n_cells k) depends on each scan's own cell count (ndt_matcher.cpp:367,392): a batch solved with
    per-registration weights gives every registration the bits it gets alone with that weight as loss->weight"""
    p = P.OXFORD
    guesses = [(0.5, -0.3, 0.02), (0.2, -0.1, 0.0), (0.9, -0.6, 0.05)]
    cases, cm, cf, pm, pf, seg, poses = build_batch(oracle, p, [41, 42, 43], guesses)
    w = np.array([p.ndt_weight / (len(c["moving"]["cells"]) * p.n_results_nn_lookup) for c in cases])
    assert len(set(w)) == 3
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=p.gnc_steps)
    prob = gpu_ctx.problem_create(cm, cf, pm, pf, seg)
    out, res = prob.register_batch(poses, capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0), opt, weights=w)
    for j, c in enumerate(cases):
        solo = gpu_ctx.problem_create(c["moving"]["cells"], c["fixed"]["cells"], c["im"], c["jf"], [0, len(c["im"])])
        o, r = solo.register_batch(c["pose0"][None], capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, w[j]), opt)
        assert np.array_equal(o[0], out[j]) and np.array_equal(r[0], res[j])
    with pytest.raises(capi.RandtError):     # the stepwise path has one weight per call
        prob.register_batch(poses, None, capi.solver_options(poll_interval=-1), weights=w)


def test_register_batch_rejects_bad_options(gpu_ctx):
    cm = np.zeros((1, 12), np.float32); cm[0, 3:] = np.eye(3).reshape(9)
    prob = gpu_ctx.problem_create(cm, cm, [0], [0], [0, 1])
    with pytest.raises(capi.RandtError):
        prob.register_batch(np.array([[1.0, 0, 0, 0]]), None, capi.solver_options(gnc_divisor=1.0))


def test_register_batch_many_segments(oracle, gpu_ctx):
    """a batch of 4500+ segments (whole-segment tiles, schedule re-planned as segments finish): every replica of a (problem, guess)
    pair gives the same bits wherever it sits in the batch, and agrees with the solo solve (whose 32-duo tiles add the same terms in
    another order) to 1e-9"""
    p = P.OXFORD
    guesses = [(0.5, -0.3, 0.02), (0.2, -0.1, 0.0), (0.9, -0.6, 0.05)]
    cases, cm, cf, pm, pf, seg, poses = build_batch(oracle, p, [31, 32, 33], guesses)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=3)
    solo = gpu_ctx.problem_create(cm, cf, pm, pf, seg)
    want_pose, want_res = solo.register_batch(poses, loss, opt)
    # 4500 segments over the same cell tables: segment s is a replica of problem s % 3 (an occasional empty segment in between)
    reps = 1500
    big_pm, big_pf, big_seg, big_pose, kind = [], [], [0], [], []
    for r in range(reps):
        for j in range(3):
            a, b = int(seg[j]), int(seg[j + 1])
            big_pm.append(pm[a:b]); big_pf.append(pf[a:b]); big_seg.append(big_seg[-1] + (b - a)); big_pose.append(poses[j]); kind.append(j)
        if r % 400 == 7:
            big_seg.append(big_seg[-1]); big_pose.append(poses[0]); kind.append(-1)
    prob = gpu_ctx.problem_create(cm, cf, np.concatenate(big_pm), np.concatenate(big_pf), np.array(big_seg, np.uint32))
    assert prob.n_segments >= 4096
    out, res = prob.register_batch(np.stack(big_pose), loss, opt)
    kind = np.array(kind)
    for j in range(3):
        sel = kind == j
        first = int(np.nonzero(sel)[0][0])
        assert np.array_equal(out[sel], np.broadcast_to(out[first], out[sel].shape)), j
        assert np.array_equal(res[sel], np.broadcast_to(res[first], res[sel].shape)), j
        assert np.max(np.abs(out[first] - want_pose[j])) < 1e-9
        assert int(res[first, capi.REG_ITERATIONS]) == int(want_res[j, capi.REG_ITERATIONS])
    assert np.all(res[kind == -1, capi.REG_STATUS] == 1)
    o = oracle.loop_constraint(cases[1]["fixed"]["cells"], cases[1]["fixed"]["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance,
                               cases[1]["moving"]["cells"], cases[1]["pose0"], p.n_results_nn_lookup, matcher_loss_scale=p.loss_function_scale,
                               loop_scale=p.loop_closure_scale, alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, max_gnc_steps=3,
                               on_manifold=True, pairs=(cases[1]["im"], cases[1]["jf"]))
    assert np.max(np.abs(want_pose[1] - o["pose"])) < 1e-7 and int(want_res[1, capi.REG_ITERATIONS]) == o["iterations"]
