"""Seeded inputs of the reference-pinned fixtures (tests/golden/ref_golden.npz).  Shared by the generator (gen_ref_golden.py, which
runs the reference's own compiled sources, oracle/_ref) and by the tests that replay them through the oracle and the CUDA path."""
import hashlib

import numpy as np

from randt_slam_b200 import params as P
from randt_slam_b200 import synth

PRESETS = ("oxford", "indoor", "outdoor", "mixed")
ANGLES = (0.02, 0.86, -2.88)            # initial-guess headings; for the last two Eigen's rotation() differs from the raw linear part
LOSS_CASES = [(1.0, -2.0, 1.0), (1.5, -2.0, 3.3), (0.5, -2.0, 1.21), (2.0, -1.0, 1.0), (2.0, -1.0, 1.1), (2.0, -1.5, 2.0), (1.0, 0.03, 1.0),
              (1.0, -0.05, 1.0), (1.0, 0.0, 2.0), (1.0, 2.0, 1.0), (0.5, 1.0, 7.0), (1.0, 2.5, 1.0)]
LOSS_S = np.concatenate([[0.0, 1e-300, 1e-12, 1e-3], np.logspace(-3, 6, 60)])
N_ROT = 512


def rot_angles():
    return np.concatenate([[0.0, 0.02, 0.7, np.pi / 2, np.pi, -np.pi / 2, 3.0, -2.2], np.random.default_rng(17).uniform(-np.pi, np.pi, N_ROT - 8)])


def vox_args(p):
    return (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)


def scans(name):
    """(preset, fixed scan, moving scan): ~half-density scans keep the fixture small"""
    p = P.PRESETS[name]
    sc = synth.scene_for(p, 3)
    kw = synth.preset_scan_kwargs(p)
    kw["n_azimuth"] = 240
    return p, synth.make_scan(sc, (0.3, -0.2, 0.1), p, 5, **kw), synth.make_scan(sc, (0.9, 0.1, 0.12), p, 6, **kw)


def raw_scan(name, n_az=96, n_bins=500):
    p = P.PRESETS[name]
    sc = synth.scene_for(p, 3)
    return p, synth.make_raw_scan(sc, (0.0, 0.0, 0.0), p, 3, n_azimuth=n_az, n_bins=n_bins, bin_size=p.max_range / (n_bins - 40)), n_az, n_bins


def pose_for(theta):
    return synth.pose_to_se2(0.5, -0.3, theta)


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
