"""Host-side logic that needs neither GPU nor oracle: the parameter derivations the reference performs in
NDTSlam::readParameters (R/src/ndt_slam/ndt_slam.cpp:653-654,691) and the seeded synthetic scan generator."""
import numpy as np

from randt_slam_b200 import params as P, synth


def test_map_size_truncation_and_cluster_grid():
    # `size_x /= resolution` on an int truncates: 50/0.5=100, 50/1.2->41, 50/1.0=50, 400/3.5->114  
    assert (P.INDOOR.size_x, P.OUTDOOR.size_x, P.MIXED.size_x, P.OXFORD.size_x) == (100, 41, 50, 114)
    # n_clusters = int((2 max_range / resolution)^2); row = int(sqrt(n_clusters))           
    assert [p.grid_row_size for p in (P.INDOOR, P.OUTDOOR, P.MIXED, P.OXFORD)] == [48, 26, 32, 57]
    assert P.OXFORD.n_clusters == 3265 and P.C1.n_clusters == 1024
    # window radius bound int(max_linf / res): r_max = r_stop - 1 -> indoor 7, outdoor 2, mixed 3, oxford 1
    assert [p.r_stop - 1 for p in (P.INDOOR, P.OUTDOOR, P.MIXED, P.OXFORD)] == [7, 2, 3, 1]


def test_shipped_loss_and_lookup_parameters():
    assert (P.INDOOR.loss_function_convexity, P.OUTDOOR.loss_function_convexity, P.MIXED.loss_function_convexity,
            P.OXFORD.loss_function_convexity) == (-2.0, -1.0, -1.5, -2.0)
    assert (P.INDOOR.min_points_per_cell, P.OUTDOOR.min_points_per_cell, P.MIXED.min_points_per_cell, P.OXFORD.min_points_per_cell) == (5, 3, 3, 10)
    assert P.OXFORD.n_results_nn_lookup == 2 and P.INDOOR.n_results_nn_lookup == 4
    for p in P.PRESETS.values():
        assert p.optimize_on_manifold and p.use_intensity_as_dimension and p.lookup_distribution


def test_synthetic_scan_is_seeded_and_oxford_shaped():
    p = P.OXFORD
    sc = synth.scene_for(p, 3)
    kw = synth.preset_scan_kwargs(p)
    a = synth.make_scan(sc, (0.5, -0.2, 0.01), p, 17, **kw)
    b = synth.make_scan(sc, (0.5, -0.2, 0.01), p, 17, **kw)
    assert a.dtype == np.float32 and a.shape[1] == 4 and np.array_equal(a, b)
    assert 3000 < len(a) < 9000                       # "~5k points" (BASELINE configs[1])
    assert np.all(a[:, 2] == 0) and a[:, 3].min() >= p.min_intensity and a[:, 3].max() <= 255
    rng = np.hypot(a[:, 0], a[:, 1])
    assert rng.min() >= p.min_range * 0.9 and rng.max() <= p.max_range
    c = synth.make_scan(sc, (0.5, -0.2, 0.01), p, 18, **kw)
    assert not np.array_equal(a[: min(len(a), len(c))], c[: min(len(a), len(c))])


def test_pose_to_se2_is_sophus_storage_order():
    q = synth.pose_to_se2(1.0, 2.0, 0.5)
    assert np.allclose(q, [np.cos(0.5), np.sin(0.5), 1.0, 2.0])
