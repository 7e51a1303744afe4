"""Shared input builders for the tests (inputs are produced with the synthetic generator + the CPU oracle)."""
import math

import numpy as np

from randt_slam_b200 import synth


def vox_args(p):
    return (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)


def make_scan(p, scene_seed, pose, scan_seed, **over):
    sc = synth.scene_for(p, scene_seed)
    kw = synth.preset_scan_kwargs(p)
    kw.update(over)
    return synth.make_scan(sc, pose, p, scan_seed, **kw)


def build_submap(O, p, scene_seed, n_scans=10, step=0.5, scan_seed0=100, **over):
    """fixed map = mergeMapCell of n_scans consecutive scans along a straight path."""
    cells = np.zeros((0, 12), np.float32); npts = np.zeros(0, np.uint32); slot = np.full(p.size_x * p.size_y, -1, np.int32)
    for i in range(n_scans):
        pose = (step * i, 0.0, 0.0)
        pts = make_scan(p, scene_seed, pose, scan_seed0 + i, **over)
        v = O.voxelize(pts, *vox_args(p))
        mc = O.transform_cells(v["cells"], math.cos(pose[2]), math.sin(pose[2]), pose[0], pose[1])
        cells, npts, slot = O.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, v["npts"])
    return dict(cells=cells, npts=npts, slot=slot)


def make_registration_case(O, p, seed, n_fixed_scans=3, true_pose=(0.6, -0.4, 0.03), guess=(0.5, -0.3, 0.02), **over):
    """(fixed map, moving scan cells, initial-guess pose[4], pair list)"""
    fixed = build_submap(O, p, seed, n_scans=n_fixed_scans, **over)
    pts = make_scan(p, seed, true_pose, 999 + seed, **over)
    mov = O.voxelize(pts, *vox_args(p))
    pose0 = synth.pose_to_se2(*guess)
    im, jf = O.associate(fixed["cells"], fixed["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance,
                         mov["cells"], pose0, p.n_results_nn_lookup)
    return dict(fixed=fixed, moving=mov, pose0=pose0, im=im, jf=jf, pts=pts)


def random_cells(rng, n, extent=40.0, thin=True):
    """cells drawn directly: random means, SPD covariances incl. xy-intensity cross terms."""
    mu = np.zeros((n, 3)); mu[:, :2] = rng.uniform(-extent, extent, (n, 2)); mu[:, 2] = rng.uniform(70, 200, n)
    ang = rng.uniform(0, math.pi, n)
    l1 = np.exp(rng.uniform(math.log(1e-3), math.log(0.1), n)); l2 = np.exp(rng.uniform(math.log(1e-3), math.log(0.1), n))
    if thin:
        l1 = np.maximum(l1, 1e-3 * l2); l2 = np.maximum(l2, 1e-3 * l1)
    c, s = np.cos(ang), np.sin(ang)
    cov = np.zeros((n, 3, 3))
    cov[:, 0, 0] = c * c * l1 + s * s * l2; cov[:, 1, 1] = s * s * l1 + c * c * l2
    cov[:, 0, 1] = cov[:, 1, 0] = c * s * (l1 - l2)
    cov[:, 2, 2] = rng.uniform(20, 400, n)
    x = rng.normal(0, 0.02, (n, 2))
    cov[:, 0, 2] = cov[:, 2, 0] = x[:, 0]; cov[:, 1, 2] = cov[:, 2, 1] = x[:, 1]
    out = np.zeros((n, 12), np.float32)
    out[:, :3] = mu; out[:, 3:] = cov.reshape(n, 9)
    return out


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, floor)
    return float(np.max(np.abs(a - b))) / scale if a.size else 0.0
