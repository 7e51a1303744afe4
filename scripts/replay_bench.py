#!/usr/bin/env python
"""Sequence replay (BASELINE configs[4]): R independent synthetic drives processed scan by scan, in lockstep, through the device path —
K1 voxelise -> K2 associate against each drive's growing submap -> K3 + K4 registration (manifold
mode, odometry loss ScaledLoss(Barron(a, alpha, mu), ndt_weight / (n_cells k)), gnc_steps of the preset) -> keyframe insertion every
second scan (transform + merge).  A single drive (R = 1) is a chain of tiny dependent launches — latency-bound, about one CPU core's
worth ("replicas only"); R drives in lockstep turn every step into one batched call.
Prints one JSON line with scans/s for the device path and for the same chain on the CPU oracle (one host thread, a bounded sample).
usage: python scripts/replay_bench.py [--replicas 256] [--scans 24]"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from randt_slam_b200 import capi, params as P, synth  # noqa: E402


def make_drive(p, seed, n_scans):
    scene = synth.scene_for(p, seed)
    kw = synth.preset_scan_kwargs(p)
    rng = np.random.default_rng(seed)
    vx, vy, w = 0.45 + rng.uniform(-0.1, 0.1), rng.uniform(-0.05, 0.05), rng.uniform(-0.006, 0.006)
    truth = [(vx * i, vy * i, w * i) for i in range(n_scans)]
    return truth, [synth.make_scan(scene, truth[i], p, seed * 1000 + i, **kw) for i in range(n_scans)]


def device_replay(ctx, p, drives, n_scans):
    R = len(drives)
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.gnc_steps, max_num_iterations=p.max_iteration)

    def batch(i):
        scans = [d[1][i] for d in drives]
        off = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint32)
        return np.concatenate(scans), off
    poses = np.tile(synth.pose_to_se2(0, 0, 0), (R, 1))
    pts, off = batch(0)
    sub = ctx.voxelize(pts, off, gp)
    iters = 0.0
    t0 = time.perf_counter()
    for i in range(1, n_scans):
        pts, off = batch(i)
        mv = ctx.voxelize(pts, off, gp)
        prob = ctx.associate(sub, mv, poses, k)
        # loss weight ndt_weight / (n_cells k) differs per drive only through n_cells: use the batch mean (the reference has one drive)
        n_cells = mv.info()[1] / R
        loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, p.ndt_weight / (n_cells * k))
        poses, res = prob.register_batch(poses, loss, opt)
        iters += float(res[:, capi.REG_ITERATIONS].mean())
        if i % 2 == 0:
            mv.transform(poses.astype(np.float32))
            sub.merge(mv)
        mv.close(); prob.close()
    ctx.sync()
    dt = time.perf_counter() - t0
    return poses, dt, iters / (n_scans - 1)


def oracle_replay(p, drive, n_scans):
    from oracle import oracle_py as O
    va = (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)
    k = p.n_results_nn_lookup
    pose = synth.pose_to_se2(0, 0, 0)
    t0 = time.perf_counter()
    v0 = O.voxelize(drive[1][0], *va)
    cells, npts, slot = v0["cells"], v0["npts"], v0["slot"]
    for i in range(1, n_scans):
        v = O.voxelize(drive[1][i], *va)
        w = p.ndt_weight / (len(v["cells"]) * k)
        o = O.loop_constraint(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, v["cells"], pose, k,
                              matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                              divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True, loss_weight=w)
        pose = o["pose"]
        if i % 2 == 0:
            mc = O.transform_cells(v["cells"], *pose.astype(np.float32))
            cells, npts, slot = O.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, v["npts"])
    return pose, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=256)
    ap.add_argument("--scans", type=int, default=24)
    ap.add_argument("--pool", type=int, default=8, help="distinct synthetic drives (replicated to --replicas)")
    args = ap.parse_args()
    p = P.OXFORD
    pool = [make_drive(p, 300 + j, args.scans) for j in range(args.pool)]
    out = {"workload": "configs[4]-shaped: synthetic drives of %d Oxford-shape scans (~5 k filtered points each), voxelise + register against the "
                       "growing submap + keyframe insertion every 2nd scan, oxford odometry parameters" % args.scans}
    with capi.Context(0) as ctx:
        for R in sorted({1, args.replicas}):
            drives = [pool[j % args.pool] for j in range(R)]
            device_replay(ctx, p, drives, min(args.scans, 6))          # warm-up (pool allocations, first launches)
            poses, dt, its = device_replay(ctx, p, drives, args.scans)
            err = np.array([math.hypot(poses[j, 2] - drives[j][0][-1][0], poses[j, 3] - drives[j][0][-1][1]) for j in range(R)])
            out["device_R%d" % R] = {"replicas": R, "scans_per_s": R * (args.scans - 1) / dt, "ms_per_step": dt * 1e3 / (args.scans - 1),
                                     "mean_lm_iterations_per_scan": its, "median_final_position_error_m": float(np.median(err)),
                                     "drives_within_0.5m": int((err < 0.5).sum())}
            if R == 1:
                pose_r1 = poses[0].copy()
    po, dto = oracle_replay(p, pool[0], args.scans)
    out["oracle_1thread"] = {"scans_per_s": (args.scans - 1) / dto, "ms_per_scan": dto * 1e3 / (args.scans - 1),
                             "max_abs_pose_difference_vs_device_R1": float(np.max(np.abs(po - pose_r1)))}
    out["note"] = ("NDT-only odometry without the reference's motion-model / IMU factors and pose-rejection gate (host side, out of scope): a drive "
                   "whose synthetic scene offers too little structure may lose track; that is counted, not hidden")
    out["sensor_rate_hz"] = 4.02
    print(json.dumps(out))


if __name__ == "__main__":
    main()
