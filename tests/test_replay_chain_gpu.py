"""BASELINE configs[4]: the full-length replay (8 609 scans, the Oxford sequence length) through randt_scan_step, against the CPU
oracle chain.

Two comparisons.  (1) Teacher-forced, every scan of the drive: the oracle chain drives (its submap and its previous pose are the
inputs), and the device registers the same scan against the same submap from the same guess; each single step must agree to 1e-6 in
pose with equal LM iteration counts for (nearly) all steps.  Where the two minimisers stop at different iterates — ceres' default
function_tolerance ends a solve while a weakly constrained direction (a corridor) is still centimetres from its minimum — the step is
re-solved on both sides with the tolerances tightened until the stopping point is the minimum itself, and must then reach the same
minimal cost (1e-6) at the same pose (1e-3: a nearly flat direction leaves the minimiser itself that loose); such steps are counted.  Run on a stride of the drive so the test stays within a minute.  (2) Free-running,
the whole drive: the device chain feeds its own poses back; both chains must track the same trajectory (the difference stays far
below the odometry's own drift) and the device chain must stay on the ground truth."""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth, workloads as W
from tests import helpers as H

pytestmark = pytest.mark.gpu
N_SCANS = 8609


@pytest.fixture(scope="module")
def drive():
    return W.make_loop_drive(P.OXFORD, W.REPLAY_SCENE_SEED, N_SCANS)


def test_full_length_replay_tracks_and_matches_oracle_chain(oracle, gpu_ctx, drive):
    p = P.OXFORD
    truth, scans = drive
    poses_g, dt, its = W.device_replay(gpu_ctx, capi, p, scans)
    est = np.stack([poses_g[:, 2], poses_g[:, 3]], 1)
    err = np.hypot(est[:, 0] - truth[:, 0], est[:, 1] - truth[:, 1])
    # 20 laps of a 30 m circle, odometry only (no loop closure), on the scene where the reference's algorithm itself tracks
    # (workloads.REPLAY_SCENE_SEED: the oracle chain stays within 0.71 m over the whole drive)
    assert np.all(np.isfinite(poses_g)) and err.max() < 1.5, err.max()
    assert abs(np.hypot(poses_g[:, 0], poses_g[:, 1]) - 1.0).max() < 1e-9      # manifold mode keeps the complex number normalised
    # the oracle chain over a prefix, free-running: same trajectory
    n_o = 600
    poses_o, _ = W.oracle_replay(oracle, p, scans[:n_o])
    d = np.abs(poses_o - poses_g[:n_o]).max(axis=1)
    assert d.max() < 0.05, d.max()
    assert np.median(d) < 1e-3
    print("replay: %d scans, %.3f ms/scan, %.1f LM iterations/scan, max error vs truth %.2f m, chains differ by max %.2e (median %.1e) over %d scans"
          % (len(scans), dt * 1e3 / (len(scans) - 1), its, err.max(), d.max(), np.median(d), n_o))


def test_every_step_matches_oracle_teacher_forced(oracle, gpu_ctx, drive):
    p = P.OXFORD
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    _, scans = drive
    opt = W.odometry_solver(capi, p)
    va = H.vox_args(p)
    n = 1200
    v0 = oracle.voxelize(scans[0], *va)
    cells, npts, slot = oracle.merge_map_cell(np.zeros((0, 12), np.float32), np.zeros(0, np.uint32), np.full(p.size_x * p.size_y, -1, np.int32), p.size_x, p.size_y,
                                              p.resolution, oracle.transform_cells(v0["cells"], 1.0, 0.0, 0.0, 0.0), v0["npts"])
    pose = synth.pose_to_se2(0, 0, 0)
    tight = loose = 0
    worst = worst_tight = 0.0
    for i in range(1, n):
        v = oracle.voxelize(scans[i], *va)
        w = p.ndt_weight / (len(v["cells"]) * k)
        o = oracle.loop_constraint(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, v["cells"], pose, k,
                                   matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                                   divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True, loss_weight=w)
        if i % 4 == 0:      # the device takes the oracle chain's state as its input
            F = gpu_ctx.map_upload(cells, [0, len(cells)], gp, npts=npts, slot=slot[None])
            M = gpu_ctx.voxelize(scans[i], [0, len(scans[i])], gp)
            prob = gpu_ctx.associate(F, M, pose[None], k)
            out, res = prob.register_batch(pose[None], capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, w), opt)
            dpose = float(np.max(np.abs(out[0] - o["pose"])))
            if dpose < 1e-6 and int(res[0, capi.REG_ITERATIONS]) == o["iterations"]:
                tight += 1
            else:
                # same problem, both sides run to the minimum itself
                ot = oracle.loop_constraint(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, v["cells"], pose, k,
                                            matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                                            divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True, loss_weight=w,
                                            max_iterations=2000, tolerances=(1e-14, 1e-13, 1e-14))
                opt_t = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=p.gnc_steps,
                                            max_num_iterations=2000, function_tolerance=1e-14, parameter_tolerance=1e-13, gradient_tolerance=1e-14)
                out_t, res_t = prob.register_batch(pose[None], capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, w), opt_t)
                d_t = float(np.max(np.abs(out_t[0] - ot["pose"])))
                # the same minimum: equal minimal cost to 1e-6; along a nearly flat direction the minimiser itself is only determined to
                # sqrt(tolerance / curvature), so the poses are held to 1e-3 there
                assert abs(res_t[0, capi.REG_SCORE] - ot["score"]) <= 1e-6 * abs(ot["score"]), (i, res_t[0, capi.REG_SCORE], ot["score"])
                assert d_t < 1e-3, (i, dpose, d_t)
                worst_tight = max(worst_tight, d_t)
                # at the default tolerances both stopped on a slow crawl (each step gaining < 1e-6 of the cost): costs within 2e-3
                assert abs(res[0, capi.REG_SCORE] - o["score"]) <= 2e-3 * abs(o["score"]), (i, res[0, capi.REG_SCORE], o["score"])
                loose += 1
            worst = max(worst, dpose)
            F.close(); M.close(); prob.close()
        pose = o["pose"]
        if i % 2 == 0:
            mc = oracle.transform_cells(v["cells"], *o["pose"].astype(np.float32))
            cells, npts, slot = oracle.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, v["npts"])
    assert tight >= 0.9 * (tight + loose), (tight, loose)
    print("teacher-forced: %d steps at 1e-6 with equal iteration counts, %d stopped at different iterates (same minimum at tight tolerances, poses within %.1e there), worst %.2e" % (tight, loose, worst_tight, worst))


@pytest.mark.parametrize("preset", ["oxford", "indoor"])
def test_scan_step_equals_the_separate_calls(gpu_ctx, preset):
    """randt_scan_step associates and solves back to back on the device (no host round trip in between, no pair lists, no K3 schedule).
    A user of the separate entry points (randt_voxelize -> randt_associate -> randt_register_batch -> randt_map_transform_se2d ->
    randt_map_merge) must get the same bits: poses, solver records and the submap after every keyframe."""
    p = P.PRESETS[preset]
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    opt = W.odometry_solver(capi, p)
    _, scans = W.make_loop_drive(p, W.REPLAY_SCENE_SEED, 40)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0)
    empty = lambda: gpu_ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(2, np.uint32), gp)
    sub_a, sub_b = empty(), empty()
    pose_a = synth.pose_to_se2(0, 0, 0)
    pose_a, _, _ = sub_a.scan_step(scans[0], gp, k, loss, p.ndt_weight, opt, True, pose_a)
    pose_b = synth.pose_to_se2(0, 0, 0)
    m0 = gpu_ctx.voxelize(scans[0], [0, len(scans[0])], gp)
    m0.transform_se2d(pose_b[None]); sub_b.merge(m0); m0.close()
    for i in range(1, len(scans)):
        pose_a, res_a, nc = sub_a.scan_step(scans[i], gp, k, loss, p.ndt_weight, opt, i % 2 == 0, pose_a)
        M = gpu_ctx.voxelize(scans[i], [0, len(scans[i])], gp)
        assert M.info()[1] == nc
        prob = gpu_ctx.associate(sub_b, M, pose_b[None], k)
        out, res_b = prob.register_batch(pose_b[None], capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, p.ndt_weight / (nc * k)), opt)
        pose_b = out[0]
        assert np.array_equal(pose_a, pose_b), i
        assert np.array_equal(res_a, res_b[0]), i
        if i % 2 == 0:
            M.transform_se2d(pose_b[None]); sub_b.merge(M)
        M.close(); prob.close()
    ca, cb = sub_a.download(), sub_b.download()
    assert np.array_equal(ca["cells"], cb["cells"]) and np.array_equal(ca["npts"], cb["npts"]) and np.array_equal(ca["slot"], cb["slot"])
    sub_a.close(); sub_b.close()


def test_scan_step_with_a_scan_that_yields_no_cells(gpu_ctx):
    """A scan whose points all fall into cells with too few points has no NDT cells: the reference's matcher adds no residual blocks
    ("WARNING: NO RESIDUALS ADDED!", ndt_matcher.cpp:454-456) and leaves the pose alone.  The composite must do the same — through
    its fused path and through the separate calls — and carry on with the next scan."""
    p = P.OXFORD
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    opt = W.odometry_solver(capi, p)
    _, scans = W.make_loop_drive(p, W.REPLAY_SCENE_SEED, 6)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0)
    sub = gpu_ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(2, np.uint32), gp)
    pose, _, _ = sub.scan_step(scans[0], gp, k, loss, p.ndt_weight, opt, True, synth.pose_to_se2(0, 0, 0))
    n_before = sub.info()[1]
    sparse = scans[1][:: max(1, len(scans[1]) // 8)][:8].copy()           # eight scattered points: no cell reaches min_points
    pose2, res, nc = sub.scan_step(sparse, gp, k, loss, p.ndt_weight, opt, True, pose)
    assert nc == 0 and np.array_equal(pose2, pose)
    assert res[capi.REG_STATUS] == 1 and res[capi.REG_ITERATIONS] == 0
    assert sub.info()[1] == n_before                                       # nothing to merge
    # the separate calls agree
    M = gpu_ctx.voxelize(sparse, [0, len(sparse)], gp)
    assert M.info()[1] == 0
    prob = gpu_ctx.associate(sub, M, pose[None], k)
    out, r = prob.register_batch(pose[None], loss, opt)
    assert prob.n_pairs == 0 and np.array_equal(out[0], pose) and r[0, capi.REG_STATUS] == 1
    M.close(); prob.close()
    # and the stream carries on
    pose3, res3, nc3 = sub.scan_step(scans[2], gp, k, loss, p.ndt_weight, opt, True, pose2)
    assert nc3 > 20 and res3[capi.REG_STATUS] == 0 and np.all(np.isfinite(pose3))
    sub.close()


def test_scan_step_chain_hands_long_registrations_to_the_general_path(gpu_ctx):
    """The chain solves on a four-warp team whatever the registration's length; the batch call solves registrations above 512 duos
    stepwise (K3 + K4), in another summation order.  So that the composite always equals the separate calls, the chain's result is
    dropped for such a scan and the registration is redone through the general path.  A fine clustering (16 384 clusters, cells kept
    from 2 points on) makes scans of ~300 cells = ~600 duos at k = 4."""
    p = P.INDOOR
    gp = capi.GridParams(float(p.max_range), 16384, 1, int(p.size_x), int(p.size_y), float(p.resolution), float(p.max_neighbor_linf_distance))
    k = p.n_results_nn_lookup
    opt = W.odometry_solver(capi, p)
    _, scans = W.make_loop_drive(p, W.REPLAY_SCENE_SEED, 8)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0)
    empty = lambda: gpu_ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(2, np.uint32), gp)
    sub_a, sub_b = empty(), empty()
    pose_a = synth.pose_to_se2(0, 0, 0)
    pose_a, _, _ = sub_a.scan_step(scans[0], gp, k, loss, p.ndt_weight, opt, True, pose_a)
    pose_b = synth.pose_to_se2(0, 0, 0)
    m0 = gpu_ctx.voxelize(scans[0], [0, len(scans[0])], gp)
    m0.transform_se2d(pose_b[None]); sub_b.merge(m0); m0.close()
    longest = 0
    for i in range(1, len(scans)):
        pose_a, res_a, nc = sub_a.scan_step(scans[i], gp, k, loss, p.ndt_weight, opt, i % 2 == 0, pose_a)
        M = gpu_ctx.voxelize(scans[i], [0, len(scans[i])], gp)
        prob = gpu_ctx.associate(sub_b, M, pose_b[None], k)
        longest = max(longest, prob.layout()[0])
        out, res_b = prob.register_batch(pose_b[None], capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, p.ndt_weight / (nc * k)), opt)
        pose_b = out[0]
        assert np.array_equal(pose_a, pose_b) and np.array_equal(res_a, res_b[0]), i
        if i % 2 == 0:
            M.transform_se2d(pose_b[None]); sub_b.merge(M)
        M.close(); prob.close()
    assert longest > 512, longest          # the case this test is about did occur
    sub_a.close(); sub_b.close()
