"""K6 (raw-scan peak filter, RadarPreprocessor::filterScan, R/src/radar_preprocessing/radar_preprocessor.cpp:45-125) vs the oracle's
sequential restatement: the kept points, their order and their transformed coordinates must be bit-identical (index work + float32
arithmetic in the same order)."""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

TF = np.array([[math.cos(0.3), -math.sin(0.3), 0.0, 1.25], [math.sin(0.3), math.cos(0.3), 0.0, -0.4], [0.0, 0.0, 1.0, 0.0]], np.float32)


def oracle_filter(oracle, raw, p, tf=None):
    return oracle.filter_scan(raw, p.min_range, p.max_range, p.min_intensity, p.beam_distance_increment_threshold, tf)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_filter_scan_matches_oracle(oracle, gpu_ctx, seed):
    p = P.OXFORD
    n_az, n_bins = 400, 1200
    raw = synth.make_raw_scan(synth.scene_for(p, 10 + seed), (0.2 * seed, -0.1, 0.01 * seed), p, seed, n_azimuth=n_az, n_bins=n_bins, bin_size=0.09)
    want, n_peaks = oracle_filter(oracle, raw, p, TF)
    got = gpu_ctx.filter_scan(raw, n_az, n_bins, capi.filter_params(p, TF))
    assert n_peaks > 200 and len(want) > 1000
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_filter_scan_edge_cases(oracle, gpu_ctx):
    p = P.OXFORD
    n_az, n_bins = 64, 300
    raw = synth.make_raw_scan(synth.scene_for(p, 5), (0, 0, 0), p, 5, n_azimuth=n_az, n_bins=n_bins, bin_size=0.3).reshape(n_az, n_bins, 4)
    raw[0, :, 3] = 0.0            # first azimuth without any valid return: the reference emits point 0
    raw[7, :, 3] = 0.0            # an empty azimuth in the middle contributes nothing
    raw[9, :, 3] = 80.0           # plateau: first strongest return wins, the walk stops immediately (intensity not falling)
    raw[-1, :, 3] = 250.0         # the last azimuth is never emitted
    raw = raw.reshape(-1, 4)
    want, _ = oracle_filter(oracle, raw, p)
    got = gpu_ctx.filter_scan(raw, n_az, n_bins, capi.filter_params(p))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert not np.any(np.isclose(got[:, 3], 250.0))
    # capacity
    with pytest.raises(capi.RandtError) as e:
        gpu_ctx.filter_scan(raw, n_az, n_bins, capi.filter_params(p), cap=4)
    assert e.value.code == capi.E_CAPACITY


def test_filter_scan_rejects_unorganised_input(gpu_ctx):
    p = P.OXFORD
    n_az, n_bins = 32, 200
    raw = synth.make_raw_scan(synth.scene_for(p, 6), (0, 0, 0), p, 6, n_azimuth=n_az, n_bins=n_bins, bin_size=0.4).reshape(n_az, n_bins, 4)
    raw[3, 100:] = raw[4, 100:]          # half of a row belongs to the next azimuth
    with pytest.raises(capi.RandtError) as e:
        gpu_ctx.filter_scan(raw.reshape(-1, 4), n_az, n_bins, capi.filter_params(p))
    assert e.value.code == capi.E_INVALID


def test_filter_then_voxelize_equals_oracle_chain(oracle, gpu_ctx):
    """raw scan -> filterScan -> Grid::cluster ... Cell::updateCell, all on the device, against the oracle's chain"""
    p = P.OXFORD
    n_az, n_bins = 400, 1200
    raw = synth.make_raw_scan(synth.scene_for(p, 21), (0.3, 0.2, 0.02), p, 21, n_azimuth=n_az, n_bins=n_bins, bin_size=0.09)
    pts = gpu_ctx.filter_scan(raw, n_az, n_bins, capi.filter_params(p))
    m = gpu_ctx.voxelize(pts, [0, len(pts)], capi.grid_params(p)).download()
    want_pts, _ = oracle_filter(oracle, raw, p)
    v = oracle.voxelize(want_pts, *H.vox_args(p))
    assert len(v["cells"]) > 30
    assert np.array_equal(m["cells"].view(np.uint32), v["cells"].view(np.uint32))


def test_filter_scan_threshold_bands_take_the_exact_path(oracle, gpu_ctx):
    """returns within a few ulp of min_range / max_range and azimuth deviations between the kernel's fast bound (0.45e-4 rad) and the
    reference's 1e-4 rad cut must be decided exactly like the reference decides them"""
    p = P.OXFORD
    n_az, n_bins = 48, 256
    raw = synth.make_raw_scan(synth.scene_for(p, 8), (0, 0, 0), p, 8, n_azimuth=n_az, n_bins=n_bins, bin_size=0.5).reshape(n_az, n_bins, 4)
    rng = np.random.default_rng(0)
    for row in range(2, 40, 3):
        az = math.atan2(float(raw[row, 10, 1]), float(raw[row, 10, 0]))
        for j, lim in enumerate((p.min_range, p.max_range)):
            for k in range(6):               # ranges lim * (1 + {-3..2} * 6e-8): float neighbours of the limit
                r = np.float32(lim) * np.float32(1.0 + (k - 3) * 6e-8)
                b = 20 + 8 * j + k
                raw[row, b, 0] = np.float32(r * math.cos(az)); raw[row, b, 1] = np.float32(r * math.sin(az)); raw[row, b, 3] = 200.0 + k
        # an azimuth wobble of 0.7e-4 rad: above the fast bound, below the reference's cut -> must not split the row
        r = 30.0
        raw[row, 100, 0] = np.float32(r * math.cos(az + 0.7e-4)); raw[row, 100, 1] = np.float32(r * math.sin(az + 0.7e-4))
    flat = raw.reshape(-1, 4)
    want, _ = oracle_filter(oracle, flat, p)
    got = gpu_ctx.filter_scan(flat, n_az, n_bins, capi.filter_params(p))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a wobble of 1.3e-4 rad does split the row in the reference: rejected
    raw[5, 100, 0] = np.float32(30.0 * math.cos(math.atan2(float(raw[5, 10, 1]), float(raw[5, 10, 0])) + 1.3e-4))
    raw[5, 100, 1] = np.float32(30.0 * math.sin(math.atan2(float(raw[5, 10, 1]), float(raw[5, 10, 0])) + 1.3e-4))
    with pytest.raises(capi.RandtError) as e:
        gpu_ctx.filter_scan(raw.reshape(-1, 4), n_az, n_bins, capi.filter_params(p))
    assert e.value.code == capi.E_INVALID


def test_filter_scans_batch_equals_scan_by_scan(oracle, gpu_ctx):
    """randt_filter_scans: a batch of scans (incl. one whose first azimuth is empty and one that keeps nothing) gives, scan after scan,
    exactly what randt_filter_scan / the oracle give for each, plus the offsets randt_voxelize takes"""
    p = P.OXFORD
    n_az, n_bins, B = 96, 500, 7
    raws = []
    for b in range(B):
        r = synth.make_raw_scan(synth.scene_for(p, 30 + b), (0.1 * b, -0.05 * b, 0.01 * b), p, 40 + b, n_azimuth=n_az, n_bins=n_bins,
                                bin_size=0.2).reshape(n_az, n_bins, 4)
        if b == 2:
            r[0, :, 3] = 0.0
        if b == 4:
            r[:, :, 3] = 0.0      # nothing above any gate: only the reference's "point 0" rule can fire, and its point fails the gates
        raws.append(r.reshape(-1, 4))
    fp = capi.filter_params(p, TF)
    pts, off = gpu_ctx.filter_scans(np.concatenate(raws), B, n_az, n_bins, fp)
    assert off[0] == 0 and off[-1] == len(pts) and (np.diff(off.astype(np.int64)) >= 0).all()
    for b in range(B):
        want, _ = oracle_filter(oracle, raws[b], p, TF)
        one = gpu_ctx.filter_scan(raws[b], n_az, n_bins, fp)
        got = pts[off[b]:off[b + 1]]
        assert got.shape == want.shape, b
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), b
        assert np.array_equal(one.view(np.uint32), want.view(np.uint32)), b
    assert off[5] == off[4]                         # the silent scan keeps nothing
    # straight into the voxeliser: same maps as voxelising every filtered scan on its own
    gp = capi.grid_params(p)
    m = gpu_ctx.voxelize(pts, off, gp).download()
    for b in (0, 3, 6):
        v = oracle.voxelize(pts[off[b]:off[b + 1]], *H.vox_args(p))
        a, z = m["cell_off"][b], m["cell_off"][b + 1]
        assert np.array_equal(m["cells"][a:z].view(np.uint32), v["cells"].view(np.uint32))
    # capacity over the whole batch, and a malformed scan anywhere in the batch
    with pytest.raises(capi.RandtError) as e:
        gpu_ctx.filter_scans(np.concatenate(raws), B, n_az, n_bins, fp, cap=int(off[-1]) - 1)
    assert e.value.code == capi.E_CAPACITY
    bad = [r.copy() for r in raws]
    bad[5].reshape(n_az, n_bins, 4)[3, 250:] = bad[5].reshape(n_az, n_bins, 4)[4, 250:]
    with pytest.raises(capi.RandtError) as e:
        gpu_ctx.filter_scans(np.concatenate(bad), B, n_az, n_bins, fp)
    assert e.value.code == capi.E_INVALID


def test_pcl_xyzi_records_feed_the_filter_unchanged(oracle, gpu_ctx):
    """the reference's filterScan takes a pcl::PointCloud<PointXYZI> (32-byte records); randt_points_from_pcl_xyzi turns cloud.points.data()
    into the path's float4 points on the device: filtering those equals filtering the host-repacked scan bit for bit"""
    import torch
    p = P.OXFORD
    raw = synth.make_raw_scan(synth.scene_for(p, 9), (0.0, 0.0, 0.0), p, 9, n_azimuth=64, n_bins=500, bin_size=0.2)
    n = len(raw)
    pcl = np.zeros((n, 8), np.float32)
    pcl[:, 0] = raw[:, 0]; pcl[:, 1] = raw[:, 1]; pcl[:, 2] = 0.37; pcl[:, 3] = 1.0      # z is ignored by the path, w = 1 as PCL sets it
    pcl[:, 4] = raw[:, 3]; pcl[:, 5:] = np.float32(-7.0)                                  # padding holds garbage
    d_pts = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    d_out = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    gpu_ctx.points_from_pcl_xyzi(pcl, d_pts.data_ptr())
    fp = capi.filter_params(p)
    kept = gpu_ctx.filter_scan_dev(d_pts.data_ptr(), 64, 500, fp, d_out.data_ptr(), n)
    gpu_ctx.sync()
    assert np.array_equal(d_pts.cpu().numpy(), raw)
    want = gpu_ctx.filter_scan(raw, 64, 500, fp)
    assert kept == len(want) and np.array_equal(d_out[:kept].cpu().numpy().view(np.uint32), want.view(np.uint32))
