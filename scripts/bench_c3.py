#!/usr/bin/env python
"""Second roofline data point for K3: BASELINE configs[2]-shaped problems (0.5 m cells, ~2 k moving x ~8 k fixed cells, k = 4, outdoor
loss), cells drawn directly as SURVEY §8d describes, reference SE(2)+intensity semantics.  Segments are long here (~8 k pairs = 16 tiles
of 256 duos), so the per-tile overheads of the small-problem bench vanish and segments are folded across tiles.
usage: python scripts/bench_c3.py [--problems 512] [--steps 100]   -> one JSON line"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from randt_slam_b200 import capi, params as P, synth  # noqa: E402


def draw_cells(rng, n, half, res):
    """one cell per occupied slot: mean uniform in its slot, SPD covariance with eigenvalues logU(1e-3, 0.1), intensity variance U(20, 400)"""
    side = int(2 * half / res)
    slots = rng.choice(side * side, size=n, replace=False)
    gx, gy = slots % side, slots // side
    mu = np.zeros((n, 3)); mu[:, 0] = -half + (gx + rng.uniform(0.05, 0.95, n)) * res; mu[:, 1] = -half + (gy + rng.uniform(0.05, 0.95, n)) * res
    mu[:, 2] = rng.uniform(70, 200, n)
    ang = rng.uniform(0, math.pi, n)
    l1 = np.exp(rng.uniform(math.log(1e-3), math.log(0.1), n)); l2 = np.exp(rng.uniform(math.log(1e-3), math.log(0.1), n))
    c, s = np.cos(ang), np.sin(ang)
    cov = np.zeros((n, 3, 3))
    cov[:, 0, 0] = c * c * l1 + s * s * l2; cov[:, 1, 1] = s * s * l1 + c * c * l2; cov[:, 0, 1] = cov[:, 1, 0] = c * s * (l1 - l2)
    cov[:, 2, 2] = rng.uniform(20, 400, n)
    # xy-intensity cross terms with correlations in (-0.4, 0.4) of the smaller xy eigenvalue (SURVEY's N(0, 0.1) is not PSD for thin cells)
    x = rng.uniform(-0.4, 0.4, (n, 2)) * np.sqrt(np.minimum(l1, l2) * cov[:, 2, 2])[:, None]
    cov[:, 0, 2] = cov[:, 2, 0] = x[:, 0]; cov[:, 1, 2] = cov[:, 2, 1] = x[:, 1]
    out = np.zeros((n, 12), np.float32); out[:, :3] = mu; out[:, 3:] = cov.reshape(n, 9)
    order = np.argsort(slots)          # grid order, like a voxelised map
    return out[order]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=384)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--pool", type=int, default=8)
    args = ap.parse_args()
    import torch
    p = P.C3
    gp = capi.grid_params(p)
    rng = np.random.default_rng(3)
    half = 45.0
    stream = torch.cuda.Stream()
    ctx = capi.Context(0, stream=stream.cuda_stream)
    fixed = [draw_cells(rng, 8000, half, p.resolution) for _ in range(args.pool)]
    moving = []
    for j in range(args.pool):                       # moving cells = perturbed subset of the fixed map seen from a displaced pose
        idx = rng.choice(8000, 2000, replace=False); idx.sort()
        m = fixed[j][idx].copy()
        m[:, :2] += rng.normal(0, 0.05, (2000, 2)).astype(np.float32)
        th, tx, ty = 0.02, 0.15, -0.1                # express in the moving frame: p_m = R^T (p_f - t)
        c, s = math.cos(th), math.sin(th)
        x, y = m[:, 0] - tx, m[:, 1] - ty
        m[:, 0], m[:, 1] = c * x + s * y, -s * x + c * y
        moving.append(m)
    sel = np.arange(args.problems) % args.pool
    f_off = np.arange(args.problems + 1, dtype=np.uint32) * 8000
    m_off = np.arange(args.problems + 1, dtype=np.uint32) * 2000
    F = ctx.map_upload(np.concatenate([fixed[j] for j in sel]), f_off, gp)
    M = ctx.map_upload(np.concatenate([moving[j] for j in sel]), m_off, gp)
    poses = np.stack([synth.pose_to_se2(0.15 + rng.uniform(-0.1, 0.1), -0.1 + rng.uniform(-0.1, 0.1), 0.02 + rng.uniform(-0.01, 0.01)) for _ in sel])
    prob = ctx.associate(F, M, poses, 4, capi.LOOKUP_MAHALANOBIS)
    pm, pf, seg = prob.download()
    F.close(); M.close()
    S, Pn = prob.n_segments, prob.n_pairs
    n_f_ref = len(np.unique(pf))
    loss = capi.make_loss(capi.LOSS_BARRON, P.OUTDOOR.loss_function_scale, P.OUTDOOR.loss_function_convexity, 1.0, 1.0)
    d_poses = torch.from_numpy(poses).cuda(); d_out = torch.zeros((S, capi.FUSED_STRIDE), dtype=torch.float64, device="cuda")
    with torch.cuda.stream(stream):
        for _ in range(10):
            prob.eval_fused_dev(d_poses.data_ptr(), d_out.data_ptr(), loss)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            prob.eval_fused_dev(d_poses.data_ptr(), d_out.data_ptr(), loss)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    alg = 48 * (prob.n_m + n_f_ref) + 8 * Pn + 32 * S + 192 * S
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except OSError:
        pk = 6650.0
    ach = alg / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": "configs[2]-shaped: 2k moving x 8k fixed cells, 0.5 m, k=4, outdoor loss (alpha=-1), SE(2)+intensity", "problems": S,
                      "pairs": Pn, "pairs_per_problem": Pn / S, "moving_cells": prob.n_m, "fixed_cells_referenced": n_f_ref, "ms_per_launch": ms,
                      "pairs_per_s": Pn / (ms * 1e-3), "algorithmic_bytes": alg, "bytes_per_pair": alg / Pn,
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk},
                      "finite": bool(np.isfinite(d_out.cpu().numpy()).all()), "bad_pairs": ctx.take_bad_pairs()}))


if __name__ == "__main__":
    main()
