"""Matcher::estimateTransformCeres (R/src/ndt_registration/ndt_matcher.cpp:322-424) through the product's host layer
(randt::Matcher::solveWindow: every NDT residual block of the window evaluated by K3 in one launch per LM evaluation, the motion-model /
IMU factors on the host, ceres' trust-region loop over the joint normal equations) against the CPU oracle's restatement of the same joint
problem (oracle/window_oracle.h: dual numbers for every block, lm_oracle.h's minimiser).  Tolerance: window states 1e-6 (north_star: 1e-5)."""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, hostapi, params as P, synth
from randt_slam_b200 import workloads as W
from tests import helpers as H

pytestmark = pytest.mark.gpu

DT = 0.25


def drive_pose(i):
    """a gentle left curve, ~1 m per scan"""
    return (1.0 * i, 0.02 * i * i, 0.01 * i)


def make_case(p, scene, n_fixed, W, rng, perturb=(0.15, 0.15, 0.01)):
    fixed_pose = [drive_pose(i) for i in range(n_fixed)]
    fixed = [H.make_scan(p, scene, fp, 300 + i) for i, fp in enumerate(fixed_pose)]
    fixed_se2 = np.stack([synth.pose_to_se2(*fp) for fp in fixed_pose])
    first = n_fixed            # the constant state is the last fixed pose's successor - 1
    window = [H.make_scan(p, scene, drive_pose(first + j), 400 + j) for j in range(1, W + 1)]
    st = np.zeros((W + 1, 14))
    for j in range(W + 1):
        x, y, th = drive_pose(first + j)
        if j >= 1:
            d = rng.normal(0, perturb)
            x, y, th = x + d[0], y + d[1], th + d[2]
        st[j, :4] = synth.pose_to_se2(x, y, th); st[j, 4:7] = [x, y, th]
        st[j, 7:9] = [4.0 + rng.normal(0, 0.2), rng.normal(0, 0.2)]; st[j, 9] = 0.04 + rng.normal(0, 0.01)
        st[j, 13] = DT * (first + j)
    return fixed, fixed_se2, window, st


def oracle_window(oracle, p, fixed, fixed_se2, window, st, q, trans, imu, tolerances, k):
    """the same blocks on the CPU: per free state, per fixed map, the association at the state's own pose (ndt_matcher.cpp:356-359)"""
    w = W.oracle_window_problem(oracle, p, fixed, fixed_se2, window, st, k)
    qo = q.copy(); qo[14] = 0 if q[7] else 2
    return oracle.window_solve(st, qo, trans, w["cells_m"], w["cells_f"], w["im"], w["jf"], w["seg_off"], w["n_cells"], imu=imu, tolerances=tolerances), w["n_cells"]


CASES = [
    # manifold, constant velocity, imu, fixed maps, window
    (True, True, False, 1, 3),     # parameters_oxford.yaml
    (True, False, True, 2, 3),     # constant acceleration + IMU, current and previous submap
    (False, True, True, 1, 2),     # vector parametrisation (pos, rot)
    (True, True, False, 1, 1),     # the second scan of a run: one free state
]


@pytest.mark.parametrize("manifold,cv,use_imu,n_fixed,W", CASES)
@pytest.mark.parametrize("tight", [False, True])
def test_window_solve_matches_oracle(oracle, manifold, cv, use_imu, n_fixed, W, tight):
    """tight = False: ceres' default tolerances, as the reference runs.  tight = True: tolerances that never fire and 12 iterations per
    solve, so that both sides take exactly the same number of trust-region steps and every step is compared (the joint problem creeps
    along a weakly determined valley — the velocities under covariance_scaling_factor 0.01 — so runs to full convergence amplify
    rounding differences into different iteration counts)."""
    p = P.OXFORD
    k = p.n_results_nn_lookup
    rng = np.random.default_rng(11 + 7 * W + n_fixed)
    fixed, fixed_se2, window, st = make_case(p, 120, n_fixed, W, rng)
    imu = np.array([0.01 + rng.normal(0, 0.002) for _ in range(W)])
    q = hostapi.window_params(k=k, gnc_steps=p.gnc_steps, loss_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                              divisor=p.gnc_control_parameter_divisor, ndt_weight=p.ndt_weight, manifold=manifold, constant_velocity=cv,
                              use_imu=use_imu, weight_imu=64.0, weight_imu_bias=750.0, covariance_scaling_factor=0.01,
                              max_iteration=12 if tight else p.max_iteration)
    tol = (1e-30, 1e-30, 1e-30) if tight else None
    trans = st[-2, :4].copy()          # current_transform_: the previous estimate
    s1, t1, info = hostapi.window_solve(capi.grid_params(p), fixed, fixed_se2, window, st, q, trans, imu=imu, tolerances=tol)
    (s0, t0, ref), n_cells = oracle_window(oracle, p, fixed, fixed_se2, window, st, q, trans, imu, tol, k)
    assert info["status"] == 0 and ref["status"] == 0 and info["rejected"] == 0 and ref["rejected"] == 0
    assert info["n_cells"] == n_cells and info["n_tangent"] == ref["n_tangent"] == 3 + (0 if cv else 2) + W * (6 + (0 if cv else 2) + (1 if use_imu else 0))
    assert abs(info["max_residual"] - ref["max_residual"]) <= 1e-9 * ref["max_residual"] and info["mu_first"] == pytest.approx(ref["mu_first"], rel=1e-9)
    assert info["gnc_solves"] == ref["gnc_solves"]
    assert np.max(np.abs(s1 - s0)) < 1e-6, (np.max(np.abs(s1 - s0)), info, ref)    # observed: <= 8e-8 after 26 steps, <= 1e-7 at the defaults
    assert np.max(np.abs(t1 - t0)) < 1e-6 and np.array_equal(t1, s1[-1, :4])
    assert abs(info["final_cost"] - ref["final_cost"]) <= 1e-7 * ref["final_cost"]
    if tight:
        assert info["total_iterations"] == ref["total_iterations"] == 2 * 13
    else:
        assert abs(info["total_iterations"] - ref["total_iterations"]) <= 1
    # one K3 launch per LM evaluation: the raw pass + per solve iteration 0 and one per candidate
    assert info["evaluations"] <= 1 + info["total_iterations"] + info["gnc_solves"]
    # the solve pulled the newest state towards the truth it was perturbed from
    x, y, th = drive_pose(n_fixed + W)
    if not tight:
        assert abs(s1[-1, 2] - x) < 0.12 and abs(s1[-1, 3] - y) < 0.12 and abs(math.atan2(s1[-1, 1], s1[-1, 0]) - th) < 0.01
    # only the newest state has both representations synchronised (ndt_matcher.cpp:399-406)
    if manifold:
        assert np.array_equal(s1[-1, 4:6], s1[-1, 2:4]) and s1[-1, 6] == math.atan2(s1[-1, 1], s1[-1, 0])
        assert np.array_equal(s1[1:-1, 4:7], st[1:-1, 4:7])
    else:
        assert s1[-1, 0] == math.cos(s1[-1, 6]) and np.array_equal(s1[-1, 2:4], s1[-1, 4:6])
    assert np.array_equal(s1[0, :7], st[0, :7]) and s1[0, 12] == st[0, 12]       # the oldest state's pose and bias are constant ...
    assert not np.array_equal(s1[0, 7:10], st[0, 7:10])                           # ... its velocities are not (ndt_matcher.cpp:304-312)
    if cv:
        assert np.array_equal(s1[:, 10:12], st[:, 10:12])


def test_window_rejection_gate(oracle):
    """an estimate further than pose_reject_translation from the prior: the newest state falls back to the one before it with zero
    velocities (ndt_matcher.cpp:408-422)"""
    p = P.OXFORD
    rng = np.random.default_rng(5)
    fixed, fixed_se2, window, st = make_case(p, 121, 1, 3, rng)
    q = hostapi.window_params(k=p.n_results_nn_lookup, gnc_steps=p.gnc_steps, divisor=p.gnc_control_parameter_divisor, ndt_weight=p.ndt_weight,
                              reject_translation=1e-3)
    trans = st[-2, :4].copy()
    s1, t1, info = hostapi.window_solve(capi.grid_params(p), fixed, fixed_se2, window, st, q, trans)
    (s0, t0, ref), _ = oracle_window(oracle, p, fixed, fixed_se2, window, st, q, trans, None, None, p.n_results_nn_lookup)
    assert info["rejected"] == 1 and ref["rejected"] == 1
    assert np.array_equal(s1[-1, :7], s1[-2, :7]) and np.all(s1[-1, 7:12] == 0.0) and np.array_equal(t1, s1[-2, :4])
    assert np.max(np.abs(s1 - s0)) < 1e-6


def test_window_without_ndt_blocks_is_reported(oracle):
    p = P.OXFORD
    rng = np.random.default_rng(6)
    fixed, fixed_se2, window, st = make_case(p, 122, 1, 2, rng)
    far = [w[:8] for w in window]      # too few points for a single cell (min_points_per_cell): the scans voxelise to nothing
    q = hostapi.window_params(k=p.n_results_nn_lookup)
    s1, t1, info = hostapi.window_solve(capi.grid_params(p), fixed, fixed_se2, far, st, q, st[-2, :4])
    assert info["status"] == 1 and np.array_equal(s1, st) and np.array_equal(t1, st[-2, :4])


# ---- a drive through the window odometry, scan by scan -------------------------------------------------------------------------------
def test_window_odometry_drive_matches_the_oracle_chain(oracle):
    """14 scans of a curve through randt_hostapi_window_replay (predictTransform -> estimateTransformCeres over the 3-scan window ->
    delayed keyframe insertion at the smoothed pose) and through the same loop on the oracle: both chains run free"""
    p = P.OXFORD
    n = 14
    truth = [(0.9 * i, 0.015 * i * i, 0.008 * i) for i in range(n)]
    scans = [H.make_scan(p, 130, truth[i], 600 + i) for i in range(n)]
    stamps = 0.25 * np.arange(n)
    q = W.window_odometry_params(hostapi, p)
    poses, states, stats, totals = hostapi.window_replay(capi.grid_params(p), scans, stamps, q, smoothing_steps=3, insertion_step=2)
    o_poses, o_states, o_cells, _ = W.oracle_window_replay(oracle, p, scans, stamps, q, 3, 2)
    assert totals["keyframes"] == 5 and totals["submap_cells"] == o_cells      # trajectory sizes 6, 8, 10, 12, 14
    # both chains run free at ceres' default tolerances (a solve stops where the relative cost change drops under 1e-6): observed 1e-5
    assert np.max(np.abs(poses - o_poses)) < 1e-4, np.max(np.abs(poses - o_poses), axis=1)
    assert np.max(np.abs(states - o_states)) < 1e-4
    est = np.stack([states[:, 2], states[:, 3], np.arctan2(states[:, 1], states[:, 0])], 1)
    assert np.max(np.abs(est[:, :2] - np.array(truth)[:, :2])) < 0.35 and np.max(np.abs(est[:, 2] - np.array(truth)[:, 2])) < 0.02
    assert np.all(stats[1:, 1] >= 3) and np.all(stats[1:, 2] == 0)
    # velocities were learnt from the motion-model factors: ~0.9 m per 0.25 s
    assert abs(states[-1, 7] - 3.6) < 0.8


def test_problem_concat_equals_the_host_built_joint_problem(oracle, gpu_ctx):
    """randt_problem_concat (device-side join of the per-state, per-fixed-map associations) == the same tables downloaded, joined on the
    host and handed to randt_problem_create: pair lists, cell snapshots and the fused records, bit for bit"""
    p = P.OXFORD
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    rng = np.random.default_rng(3)
    fixed, fixed_se2, window, st = make_case(p, 123, 2, 3, rng)
    f_maps = []
    for pts, T in zip(fixed, fixed_se2):
        m = gpu_ctx.voxelize(pts, [0, len(pts)], gp); m.transform_se2d(T[None]); f_maps.append(m)
    parts, seg_of_part = [], []
    for j in range(3):
        mv = gpu_ctx.voxelize(window[j], [0, len(window[j])], gp)
        for fm in f_maps:
            parts.append(gpu_ctx.associate(fm, mv, st[j + 1, :4][None], k)); seg_of_part.append(j)
        mv.close()
    joint = gpu_ctx.problem_concat(parts, seg_of_part, 3)
    cm, cf, pm, pf, so, mb, fb = [], [], [], [], [0], 0, 0
    for q, sg in zip(parts, seg_of_part):
        a, b, _ = q.download(); m_, f_ = q.download_cells()
        pm.append(a + mb); pf.append(b + fb); cm.append(m_); cf.append(f_); mb += len(m_); fb += len(f_)
        if len(so) == sg + 1:
            so.append(0)
        so[sg + 1] = sum(len(x) for x in pm)
    ref = gpu_ctx.problem_create(np.concatenate(cm), np.concatenate(cf), np.concatenate(pm), np.concatenate(pf), np.array(so, np.uint32))
    assert joint.n_segments == 3 and joint.n_pairs == ref.n_pairs > 100
    for x, y in zip(joint.download(), ref.download()):
        assert np.array_equal(x, y)
    for x, y in zip(joint.download_cells(), ref.download_cells()):
        assert np.array_equal(x, y)
    loss = capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0, 1.3, 0.02)
    poses = st[1:, :4].copy()
    assert np.array_equal(joint.eval_fused(poses, loss), ref.eval_fused(poses, loss))
    r0, J0 = joint.eval_emit(poses); r1, J1 = ref.eval_emit(poses)
    assert np.array_equal(r0, r1) and np.array_equal(J0, J1)
    # an empty segment in the middle (a window scan without cells) and misuse
    gap = gpu_ctx.problem_concat([parts[0], parts[5]], [0, 2], 3)
    out = gap.eval_fused(poses, loss)
    assert np.all(out[1] == 0.0)
    assert np.array_equal(out[0], parts[0].eval_fused(poses[0:1], loss)[0]) and np.array_equal(out[2], parts[5].eval_fused(poses[2:3], loss)[0])
    with pytest.raises(capi.RandtError):
        gpu_ctx.problem_concat([parts[1], parts[0]], [1, 0], 3)
    with pytest.raises(capi.RandtError):
        gpu_ctx.problem_concat([joint], [0], 1)
    for q in parts + [joint, ref, gap] + f_maps:
        q.close()
