#!/usr/bin/env python
"""K1 on a resident batch of scans (default 256 Oxford-shape scans), a few calls: the target of `ncu -k regex:k1_voxelize` when tuning K1.
usage: python scripts/k1_profile.py [n_scans]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from randt_slam_b200 import capi, params as P, workloads as W  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = P.OXFORD
sub, _, mov, _ = W.pool_scans(p, 1)
base = (sub[:16] + mov) * ((n_scans + 31) // 32)
base = base[:n_scans]
pts = np.concatenate(base)
off = np.concatenate([[0], np.cumsum([len(s) for s in base])]).astype(np.uint32)
gp = capi.grid_params(p)
with capi.Context(0) as ctx:
    import torch
    d = torch.from_numpy(pts).cuda()
    for _ in range(3):
        ctx.voxelize(d.data_ptr(), off, gp, pts_on_device=True).close()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.voxelize(d.data_ptr(), off, gp, pts_on_device=True).close()
    ctx.sync()
    print("%d scans, %d points: %.1f us per randt_voxelize call" % (n_scans, len(pts), (time.perf_counter() - t0) / 10 * 1e6))
