"""BASELINE configs[3], literally: a batch of 256 independent keyframe-pair registrations (Matcher::estimateLoopConstraint,
R/src/ndt_registration/ndt_matcher.cpp:426-493, one fresh problem per pair) partitioned over N ranks with shard.partition.

SURVEY §4 (iv): the same batch on 1 vs N ranks gives identical poses.  Every block is built and solved here on the one GPU the test
box has, exactly as a rank would build and solve its shard (randt_slam_b200/workloads.py: build_literal_block -> K1, transform/merge,
K2, K7); the gathered table must be BIT-identical for N = 1, 2, 4 and 8, because the persistent solver's arithmetic for one
registration depends on nothing but that registration's pair list.  (bench.py runs the same batch on real ranks over NCCL and prints
the table's checksum, which must not change with --gpus.)
"""
import hashlib
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, shard, workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def batch():
    return W.literal_batch(P.OXFORD, n=256, seed=40, scenes=16)


def solve_sharded(ctx, batch, world, weights=None):
    table = np.zeros((batch["n"], shard.ROW))
    for begin, end in shard.partition(batch["n"], world, weights):
        rows, _ = W.solve_literal_block(ctx, capi, batch, begin, end)
        table[begin:end] = rows
        table[begin:end, 7] = np.arange(begin, end)
    return table


def test_literal_256_batch_identical_for_1_2_4_8_ranks(gpu_ctx, batch):
    t1 = solve_sharded(gpu_ctx, batch, 1)
    assert np.all(t1[:, 6] == 0), "every registration of the batch has residual blocks"
    for world in (2, 4, 8):
        tn = solve_sharded(gpu_ctx, batch, world)
        assert np.array_equal(tn.view(np.uint64), t1.view(np.uint64)), "world %d" % world
    # balanced by a per-problem weight (uneven blocks) as well
    w = 1.0 + (np.arange(batch["n"]) % 7)
    tw = solve_sharded(gpu_ctx, batch, 4, w)
    assert np.array_equal(tw.view(np.uint64), t1.view(np.uint64))
    # the registrations found their poses: translation within 0.3 m, heading within 0.02 rad of the truth for nearly all of them
    th = np.arctan2(t1[:, 1], t1[:, 0])
    ok = (np.hypot(t1[:, 2] - batch["truth"][:, 0], t1[:, 3] - batch["truth"][:, 1]) < 0.3) & (np.abs(th - batch["truth"][:, 2]) < 0.02)
    assert ok.mean() > 0.9, ok.mean()
    assert hashlib.sha256(t1[:, :7].tobytes()).hexdigest() == hashlib.sha256(tw[:, :7].tobytes()).hexdigest()


def test_team_and_one_warp_solver_modes_give_identical_bits(gpu_ctx):
    """Small batches are solved in team mode (a CTA of four warps shares one registration: three prepare pair terms, warp 0 adds them in
    the one-warp kernel's order), larger ones with one warp per registration.  A 600-registration batch solved in one call (one-warp
    mode) and in blocks of 150 (team mode) must give the same table, bit for bit."""
    big = W.literal_batch(P.OXFORD, n=600, seed=41, scenes=16)
    one = np.zeros((600, shard.ROW)); team = np.zeros((600, shard.ROW))
    rows, _ = W.solve_literal_block(gpu_ctx, capi, big, 0, 600)
    one[:] = rows
    for begin in range(0, 600, 150):
        rows, _ = W.solve_literal_block(gpu_ctx, capi, big, begin, begin + 150)
        team[begin:begin + 150] = rows
    assert np.array_equal(one[:, :7].view(np.uint64), team[:, :7].view(np.uint64))
    # a lone registration (the per-scan call) as well
    rows, _ = W.solve_literal_block(gpu_ctx, capi, big, 77, 78)
    assert np.array_equal(rows[0, :7].view(np.uint64), one[77, :7].view(np.uint64))


def test_literal_batch_matches_oracle(oracle, gpu_ctx, batch):
    """a sample of the batch against the CPU restatement of estimateLoopConstraint on the pair lists the device built.  Raw ambient pose
    block: at ceres' default function_tolerance two implementations may stop an iteration apart, so the comparison is made
    (a) at the defaults, on the gauge-invariant pose, to what that tolerance leaves, and (b) with the tolerances tightened on both sides
    until the stopping point is the minimum itself: there the gauge-invariant poses agree to 1e-6."""
    p = batch["p"]
    begin, end = 0, 24
    prob, poses = W.build_literal_block(gpu_ctx, capi, batch, begin, end)
    loss, opt = W.literal_solver(capi, p)
    out, res = prob.register_batch(poses, loss, opt)
    tight = capi.solver_options(use_manifold=0, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                                gnc_max_steps=p.loop_closure_gnc_steps, max_num_iterations=2000, function_tolerance=1e-14,
                                parameter_tolerance=1e-13, gradient_tolerance=1e-14)
    out_t, res_t = prob.register_batch(poses, loss, tight)
    pm, pf, seg = prob.download()
    cm, cf = prob.download_cells()
    dummy = np.full(p.size_x * p.size_y, -1, np.int32)
    worst = worst_t = 0.0
    for s in range(end - begin):
        a, b = int(seg[s]), int(seg[s + 1])
        m0, f0 = int(pm[a:b].min()), int(pf[a:b].min())
        kw = dict(matcher_loss_scale=p.loss_function_scale, loop_scale=p.loop_closure_scale, alpha=p.loss_function_convexity,
                  divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.loop_closure_gnc_steps, on_manifold=False,
                  pairs=(pm[a:b] - m0, pf[a:b] - f0))
        args = (cf[f0:int(pf[a:b].max()) + 1], dummy, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, cm[m0:int(pm[a:b].max()) + 1],
                poses[s], p.n_results_nn_lookup)
        o = oracle.loop_constraint(*args, max_iterations=p.max_iteration, **kw)
        th, tho = math.atan2(out[s, 1], out[s, 0]), math.atan2(o["pose"][1], o["pose"][0])
        assert abs(th - tho) < 2e-3 and np.max(np.abs(out[s, 2:] - o["pose"][2:])) < 5e-3
        assert int(res[s, capi.REG_GNC_SOLVES]) == o["gnc_solves"]
        worst = max(worst, abs(th - tho), float(np.max(np.abs(out[s, 2:] - o["pose"][2:]))))
        ot = oracle.loop_constraint(*args, max_iterations=2000, tolerances=(1e-14, 1e-13, 1e-14), **kw)
        tht, thot = math.atan2(out_t[s, 1], out_t[s, 0]), math.atan2(ot["pose"][1], ot["pose"][0])
        d = max(abs(tht - thot), float(np.max(np.abs(out_t[s, 2:] - ot["pose"][2:]))))
        worst_t = max(worst_t, d)
        assert d < 1e-6, (s, d)
    print("literal batch vs oracle: default tolerances %.2e, tight tolerances %.2e" % (worst, worst_t))
    prob.close()


def test_persistent_and_stepwise_solvers_agree(gpu_ctx, batch):
    """K7 (one warp per registration, one launch) against the stepwise path (K3 fused + K4 per LM iteration, poll_interval < 0): same
    minimiser on evaluations that add the same terms in another order (the stepwise path cuts a registration into 32-duo tiles here);
    manifold mode, where the problem is well posed: equal iteration counts, poses to 1e-9."""
    p = batch["p"]
    prob, poses = W.build_literal_block(gpu_ctx, capi, batch, 32, 96)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0)
    kw = dict(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=p.gnc_steps,
              max_num_iterations=p.max_iteration)
    l0 = gpu_ctx.launch_count
    a, ra = prob.register_batch(poses, loss, capi.solver_options(**kw))
    l1 = gpu_ctx.launch_count
    b, rb = prob.register_batch(poses, loss, capi.solver_options(poll_interval=-1, **kw))
    l2 = gpu_ctx.launch_count
    assert l1 - l0 == 1 and l2 - l1 > 20
    assert np.array_equal(ra[:, capi.REG_ITERATIONS], rb[:, capi.REG_ITERATIONS])
    assert np.max(np.abs(a - b)) < 1e-9
    assert np.max(np.abs(ra[:, capi.REG_SCORE] - rb[:, capi.REG_SCORE]) / ra[:, capi.REG_SCORE]) < 1e-9
    prob.close()
