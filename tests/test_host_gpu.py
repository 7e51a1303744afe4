"""The C++ host mirror of the reference surface (include/randt_host.hpp: randt::Map, randt::Matcher, randt::NdtCostFunction), driven
through its C hooks, vs the CPU oracle.  These read like the tests the reference never had for Matcher::estimateLoopConstraint,
Matcher::addNDTFactor + CostFunction::Evaluate and Matcher::estimateTransformGlobalBNB."""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, hostapi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def oracle_maps(oracle, p, fixed_pts, moving_pts):
    f = oracle.voxelize(fixed_pts, *H.vox_args(p)); m = oracle.voxelize(moving_pts, *H.vox_args(p))
    return f, m


@pytest.mark.parametrize("n", [1, 3])
def test_estimate_loop_constraint_matches_oracle(oracle, n):
    p = P.OXFORD
    fixed = [H.make_scan(p, 40 + b, (0.0, 0.0, 0.0), 1 + b) for b in range(n)]
    moving = [H.make_scan(p, 40 + b, (0.6, -0.4, 0.03), 50 + b) for b in range(n)]
    guess = np.stack([synth.pose_to_se2(0.4 + 0.05 * b, -0.25, 0.02) for b in range(n)])
    poses, scores = hostapi.loop_constraints(capi.grid_params(p), fixed, moving, guess, p.n_results_nn_lookup, p.loss_function_scale,
                                             p.loss_function_convexity, p.gnc_control_parameter_divisor, p.loop_closure_gnc_steps, p.loop_closure_scale)
    for b in range(n):
        f, m = oracle_maps(oracle, p, fixed[b], moving[b])
        o = oracle.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, m["cells"], guess[b],
                                   p.n_results_nn_lookup, matcher_loss_scale=p.loss_function_scale, loop_scale=p.loop_closure_scale,
                                   alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.loop_closure_gnc_steps,
                                   on_manifold=False)   # quirk B.13: the loop-closure pose block carries no manifold
        # Raw-ambient mode: the scale of (cos, sin) is a gauge direction the residual cannot see (theta = atan2), so J^T J is singular
        # along it and the LM step there is rounding noise over the 1e-6 minimum diagonal — in the reference as much as here.  Two
        # evaluations that agree to 1e-13 therefore drift apart in |(c, s)| and may stop an iteration apart; what is comparable is the
        # gauge-invariant pose to the accuracy of ceres' function tolerance (1e-6 relative cost change ~ 1e-3 in the pose).
        th, tho = math.atan2(poses[b, 1], poses[b, 0]), math.atan2(o["pose"][1], o["pose"][0])
        assert abs(th - tho) < 2e-3 and np.max(np.abs(poses[b, 2:] - o["pose"][2:])) < 5e-3
        assert abs(scores[b] - o["score"]) <= 1e-3 * abs(o["score"])
        assert abs(poses[b, 2] - 0.6) < 0.3 and abs(poses[b, 3] + 0.4) < 0.3


@pytest.mark.parametrize("with_loss", [False, True])
def test_cost_function_evaluate_matches_per_block_ceres_semantics(oracle, with_loss):
    """one batched ceres::CostFunction == the reference's P residual blocks: same corrected residuals / Jacobian rows, same cost"""
    p = P.OXFORD
    fixed = H.make_scan(p, 60, (0.0, 0.0, 0.0), 3); moving = H.make_scan(p, 60, (0.6, -0.4, 0.03), 4)
    guess = synth.pose_to_se2(0.5, -0.3, 0.02)
    pose = synth.pose_to_se2(0.52, -0.33, 0.025) * [1.0005, 1.0005, 1, 1]
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.7, 5000.0 / 200) if with_loss else None
    res, J, max_raw = hostapi.cost_function(capi.grid_params(p), fixed, moving, p.n_results_nn_lookup, guess, pose, loss)
    f, m = oracle_maps(oracle, p, fixed, moving)
    im, jf = oracle.associate(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, m["cells"], guess,
                              p.n_results_nn_lookup)
    r0, J0 = oracle.eval_pairs(0, m["cells"], f["cells"], im, jf, pose, 0)
    assert len(res) == len(im) + 1 and abs(max_raw - r0.max()) < 1e-10 * r0.max()
    if not with_loss:
        assert np.max(np.abs(res[:-1] - r0) / r0) < 1e-9 and res[-1] == 0.0
        assert np.max(np.abs(J[:-1] - J0)) < 1e-8 * np.max(np.abs(J0)) and np.all(J[-1] == 0)
    else:
        lt = (loss.kind, loss.scale, loss.alpha, loss.mu, loss.weight)
        fo = oracle.fused(0, m["cells"], f["cells"], im, jf, pose, lt, True)
        # 1/2 |residuals|^2 is the robustified cost ceres reports, J^T J and J^T r are what ceres accumulates block by block
        assert abs(0.5 * np.sum(res ** 2) - fo["cost"]) < 1e-9 * fo["cost"]
        assert np.max(np.abs(J.T @ J - fo["H"])) < 1e-8 * np.max(np.abs(fo["H"]))
        assert np.max(np.abs(J.T @ res - fo["g"])) < 1e-8 * np.max(np.abs(fo["g"]))
        rho = np.array([oracle.loss_eval(*lt[:4], lt[4], s)[1] for s in r0 ** 2])
        assert np.max(np.abs(res[:-1] - np.sqrt(rho) * r0)) < 1e-9 * np.max(r0)
    # residual-only call (jacobians == nullptr)
    res2, J2, _ = hostapi.cost_function(capi.grid_params(p), fixed, moving, p.n_results_nn_lookup, guess, pose, loss, want_jac=False)
    assert J2 is None and np.array_equal(res, res2)


def test_global_bnb_matches_sequential_reference_search(oracle):
    p = P.OXFORD
    fixed = H.make_scan(p, 70, (0.0, 0.0, 0.0), 5); moving = H.make_scan(p, 70, (1.3, -0.9, 0.06), 6)
    start = synth.pose_to_se2(0.3, 0.1, 0.0)
    pose, min_cost, launches = hostapi.bnb(capi.grid_params(p), fixed, moving, start, p.loss_function_convexity, p.loop_closure_scale)
    f, m = oracle_maps(oracle, p, fixed, moving)
    o = oracle.bnb(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, m["cells"], start,
                   p.loss_function_convexity, p.loop_closure_scale)
    assert o["n_evaluated"] > 100
    assert abs(min_cost - o["min_cost"]) <= 1e-9 * abs(o["min_cost"])
    assert np.max(np.abs(pose - o["pose"])) < 1e-12
    assert launches <= 16          # association (6 kernels + record build) + one sweep per tree level, not one launch per pose


def test_host_layer_reports_errors(oracle):
    p = P.OXFORD
    gp = capi.grid_params(p)
    pts = np.zeros((50, 4), np.float32); pts[:, 0] = 1e6; pts[:25, 0] = -1e6; pts[:, 3] = 90
    with pytest.raises(capi.RandtError):
        hostapi.loop_constraints(gp, [pts], [pts], [synth.pose_to_se2(0, 0, 0)], 2, 1.0, -2.0, 1.1, 2, 0.5)


def test_export_normal_distributions_wire_format(oracle):
    """ndt_msgs Mean (x, y, i) + Covariance (xx, xy, xi, yy, yi, ii) as NDTSlam::createVisualizationMsg fills them (ndt_slam.cpp:370-393)"""
    p = P.OXFORD
    pts = H.make_scan(p, 80, (0.0, 0.0, 0.0), 7)
    mean, cov = hostapi.export_normal_distributions(capi.grid_params(p), pts)
    v = oracle.voxelize(pts, *H.vox_args(p))["cells"]
    assert mean.shape == (len(v), 3) and cov.shape == (len(v), 6)
    assert np.array_equal(mean, v[:, :3].astype(np.float64))
    c = v[:, 3:].reshape(-1, 3, 3).astype(np.float64)
    assert np.array_equal(cov, np.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], 1))


@pytest.mark.parametrize("n_fixed", [1, 2])
def test_estimate_transform_ndt_matches_oracle(oracle, n_fixed):
    """the NDT part of Matcher::estimateTransformCeres: blocks against every fixed map (current submap + the previous one while they
    overlap), ScaledLoss(Barron, ndt_weight / (n_cells k)), GNC, SE2 manifold; and the rejection gate (ndt_matcher.cpp:408-422)"""
    p = P.OXFORD
    k = p.n_results_nn_lookup
    fixed_pose = [(0.0, 0.0, 0.0), (0.5, 0.1, 0.01)][:n_fixed]
    fixed = [H.make_scan(p, 90, fp, 11 + i) for i, fp in enumerate(fixed_pose)]
    fixed_se2 = np.stack([synth.pose_to_se2(*fp) for fp in fixed_pose])
    moving = H.make_scan(p, 90, (0.9, -0.3, 0.03), 21)
    prior = synth.pose_to_se2(0.8, -0.2, 0.02)
    args = (k, p.loss_function_scale, p.loss_function_convexity, p.gnc_control_parameter_divisor, p.gnc_steps, p.ndt_weight)
    pose, ok = hostapi.odometry(capi.grid_params(p), fixed, fixed_se2, moving, prior, *args)
    assert ok
    # oracle: the same blocks, fixed tables one after the other
    mv = oracle.voxelize(moving, *H.vox_args(p))
    f_cells, im_all, jf_all, base = [], [], [], 0
    for pts, T in zip(fixed, fixed_se2):
        v = oracle.voxelize(pts, *H.vox_args(p))
        cells = oracle.transform_cells(v["cells"], *T.astype(np.float32))
        slot = v["slot"]          # Map::transformMap moves the cells, not grid_indizes_ (ndt_map.cpp:177-182): lookups use the old slots
        im, jf = oracle.associate(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, mv["cells"], prior, k)
        f_cells.append(cells); im_all.append(im); jf_all.append(jf + base); base += len(cells)
    w = p.ndt_weight / (len(mv["cells"]) * k)
    o = oracle.loop_constraint(np.concatenate(f_cells), np.full(p.size_x * p.size_y, -1, np.int32), p.size_x, p.size_y, p.resolution,
                               p.max_neighbor_linf_distance, mv["cells"], prior, k, matcher_loss_scale=p.loss_function_scale,
                               loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor,
                               max_gnc_steps=p.gnc_steps, on_manifold=True, loss_weight=w,
                               pairs=(np.concatenate(im_all).astype(np.uint32), np.concatenate(jf_all).astype(np.uint32)))
    assert o["status"] == 0 and len(np.concatenate(im_all)) > 50 * n_fixed
    assert np.max(np.abs(pose - o["pose"])) < 1e-7
    est = (pose[2], pose[3], np.arctan2(pose[1], pose[0]))
    assert abs(est[0] - 0.9) < 0.3 and abs(est[1] + 0.3) < 0.3 and abs(est[2] - 0.03) < 0.02
    # the gate: an estimate further than pose_reject_translation from the prior is refused and the prior comes back untouched
    pose_r, ok_r = hostapi.odometry(capi.grid_params(p), fixed, fixed_se2, moving, prior, *args, reject_translation=1e-4)
    assert not ok_r and np.array_equal(pose_r, prior)
