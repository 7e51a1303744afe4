"""Sequence replay (BASELINE configs[4] in miniature): a synthetic drive processed scan by scan through the whole device path —
voxelise (K1) -> associate against the growing submap (K2) -> GNC + LM registration (K3 + K4, manifold mode, odometry loss
ScaledLoss(Barron(a, alpha, mu), ndt_weight / (n_cells k)), R/src/ndt_registration/ndt_matcher.cpp:392) -> keyframe insertion every
second scan (transformMap + mergeMapCell, R/src/local_fuser/local_fuser.cpp:164-190) — against the same chain on the CPU oracle.

The two chains are independent: each feeds its own poses back into its own submap.  Poses are asserted to 1e-6 (the float32 map
maintenance sees poses that agree to ~1e-9, which leaves the merged cells bit-identical on this sequence)."""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_drive_replay_matches_oracle_chain(oracle, gpu_ctx):
    p = P.OXFORD
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    scene = synth.scene_for(p, 77)
    kw = synth.preset_scan_kwargs(p)
    n_scans = 9
    truth = [(0.45 * i, 0.05 * i, 0.004 * i) for i in range(n_scans)]
    scans = [synth.make_scan(scene, truth[i], p, 500 + i, **kw) for i in range(n_scans)]
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.gnc_steps, max_num_iterations=p.max_iteration)

    # ---- device chain ----
    g_poses = [synth.pose_to_se2(0, 0, 0)]
    sub = gpu_ctx.voxelize(scans[0], [0, len(scans[0])], gp)
    for i in range(1, n_scans):
        mv = gpu_ctx.voxelize(scans[i], [0, len(scans[i])], gp)
        guess = g_poses[-1]
        prob = gpu_ctx.associate(sub, mv, guess[None], k)
        n_cells = mv.info()[1]
        loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, p.ndt_weight / (n_cells * k))
        pose, res = prob.register_batch(guess[None], loss, opt)
        assert res[0, capi.REG_STATUS] == 0
        g_poses.append(pose[0])
        if i % 2 == 0:      # insertion_step: 2 (parameters_oxford.yaml:43)
            mv.transform(pose[0].astype(np.float32)[None])
            sub.merge(mv)
        mv.close(); prob.close()
    g_sub = sub.download()

    # ---- oracle chain ----
    o_poses = [synth.pose_to_se2(0, 0, 0)]
    v0 = oracle.voxelize(scans[0], *H.vox_args(p))
    cells, npts, slot = v0["cells"], v0["npts"], v0["slot"]
    for i in range(1, n_scans):
        v = oracle.voxelize(scans[i], *H.vox_args(p))
        guess = o_poses[-1]
        w = p.ndt_weight / (len(v["cells"]) * k)
        o = oracle.loop_constraint(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, v["cells"], guess, k,
                                   matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                                   divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True, loss_weight=w)
        assert o["status"] == 0
        o_poses.append(o["pose"])
        if i % 2 == 0:
            t = o["pose"].astype(np.float32)
            mc = oracle.transform_cells(v["cells"], *t)
            cells, npts, slot = oracle.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, v["npts"])

    g = np.array(g_poses); o = np.array(o_poses)
    assert np.max(np.abs(g - o)) < 1e-6, np.max(np.abs(g - o), axis=1)
    # the drive is actually tracked (each registration starts from the previous pose)
    est = np.stack([g[:, 2], g[:, 3], np.arctan2(g[:, 1], g[:, 0])], 1)
    assert np.max(np.abs(est[:, :2] - np.array(truth)[:, :2])) < 0.3 and np.max(np.abs(est[:, 2] - np.array(truth)[:, 2])) < 0.02
    # and the submaps agree
    assert len(g_sub["cells"]) == len(cells)
    assert np.array_equal(g_sub["npts"], npts)
    assert H.rel_err(g_sub["cells"], cells) < 1e-5
