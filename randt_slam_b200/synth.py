"""Seeded synthetic radar scans for tests and bench (SURVEY.md §8d "Synthetic inputs").

The Oxford rosbags are not available offline, so every input is generated: a 2-D scene of random wall
segments plus point reflectors inside a square of half-width `max_range`; a scan casts `n_azimuth` beams from a
sensor pose, keeps the nearest hit per beam and emits the short run of range bins around it that the reference's
per-azimuth peak filter (R/src/radar_preprocessing/radar_preprocessor.cpp:45-125) would keep: a few points per
beam, intensity peaked at the hit.  Output layout is the path's input layout: float32 [N, 4] = (x, y, 0, intensity)
in the sensor/base frame.
"""
import math

import numpy as np


class Scene:
    def __init__(self, seed, half_width, n_walls=60, n_reflectors=80, wall_len=(0.08, 0.45), boundary=0.85):
        rng = np.random.default_rng(seed)
        hw = half_width
        c = rng.uniform(-hw, hw, size=(n_walls, 2))
        ang = rng.uniform(0, math.pi, size=n_walls)
        ln = rng.uniform(wall_len[0] * hw, wall_len[1] * hw, size=n_walls)
        d = np.stack([np.cos(ang), np.sin(ang)], 1) * ln[:, None] * 0.5
        # keep the neighbourhood of the origin free so the sensor is not inside a wall
        self.a = c - d
        self.b = c + d
        refl = rng.uniform(-hw, hw, size=(n_reflectors, 2))
        r = 0.004 * hw
        self.a = np.concatenate([self.a, refl - [r, 0]], 0)
        self.b = np.concatenate([self.b, refl + [r, 0]], 0)
        if boundary:
            # a jagged closed outline (street canyon / room) so that most beams return something
            k = 24
            ang = np.sort(rng.uniform(0, 2 * math.pi, k))
            rad = boundary * hw * rng.uniform(0.55, 1.0, k)
            v = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
            self.a = np.concatenate([self.a, v], 0)
            self.b = np.concatenate([self.b, np.roll(v, -1, 0)], 0)
        self.half_width = hw


def cast(scene, pose, n_azimuth, max_range, rng, az_jitter_deg=0.45):
    """nearest wall hit per beam.  pose = (x, y, theta) of the sensor in the scene frame.  -> (angle[n], range[n]) (range nan = miss)"""
    th = pose[2] + np.arange(n_azimuth) * (2 * math.pi / n_azimuth) + np.deg2rad(rng.uniform(-az_jitter_deg, az_jitter_deg, n_azimuth)) * 0.0
    o = np.array(pose[:2], dtype=np.float64)
    d = np.stack([np.cos(th), np.sin(th)], 1)                 # [n,2]
    a, b = scene.a, scene.b
    e = b - a                                                  # [m,2]
    # ray o + t d  vs segment a + u e :  t = cross(a-o, e)/cross(d, e), u = cross(a-o, d)/cross(d, e)
    ao = a - o                                                 # [m,2]
    den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (ao[None, :, 0] * e[None, :, 1] - ao[None, :, 1] * e[None, :, 0]) / den
        u = (ao[None, :, 0] * d[:, None, 1] - ao[None, :, 1] * d[:, None, 0]) / den
    ok = (t > 0) & (u >= 0) & (u <= 1) & np.isfinite(t)
    t = np.where(ok, t, np.inf)
    rr = t.min(1)
    rr = np.where(rr < max_range, rr, np.nan)
    return th - pose[2], rr


def make_scan(scene, pose, params, seed, n_azimuth=400, bin_size=0.0438, half_bins=(2, 7), range_sigma=0.06,
              az_jitter_deg=0.45, intensity_mean=90.0, intensity_sigma=15.0, intensity_clip=(70.0, 255.0)):
    """One filtered scan in the sensor frame: float32 [N,4] (x, y, 0, intensity), beams in azimuth order."""
    rng = np.random.default_rng(seed)
    ang, rr = cast(scene, pose, n_azimuth, params.max_range, rng)
    hit = np.isfinite(rr) & (rr > params.min_range + half_bins[1] * bin_size)
    ang = ang[hit] + np.deg2rad(rng.uniform(-az_jitter_deg, az_jitter_deg, hit.sum()))
    rr = rr[hit] + rng.normal(0.0, range_sigma, hit.sum())
    nb = rng.integers(half_bins[0], half_bins[1] + 1, size=len(rr))          # half-width of the kept bin run
    peak = np.clip(rng.normal(intensity_mean + 40.0, intensity_sigma, len(rr)), intensity_clip[0] + 10, intensity_clip[1])
    pts = []
    for k in range(-half_bins[1], half_bins[1] + 1):
        m = np.abs(k) <= nb
        r_k = rr[m] + k * bin_size
        # triangular intensity profile falling from the peak towards the gate value
        w = 1.0 - np.abs(k) / (nb[m] + 1.0)
        inten = intensity_clip[0] + 1.0 + (peak[m] - intensity_clip[0] - 1.0) * w + rng.normal(0, 1.0, m.sum())
        inten = np.clip(inten, intensity_clip[0] + 0.5, intensity_clip[1])
        ok = (r_k > params.min_range) & (r_k < params.max_range)
        beam = np.nonzero(m)[0][ok]
        pts.append(np.stack([beam.astype(np.float64), np.full(ok.sum(), k, np.float64), r_k[ok] * np.cos(ang[m][ok]),
                             r_k[ok] * np.sin(ang[m][ok]), inten[ok]], 1))
    p = np.concatenate(pts, 0)
    order = np.lexsort((p[:, 1], p[:, 0]))     # beam-major, bin-minor: the order a polar scan is stored in
    p = p[order]
    out = np.zeros((len(p), 4), np.float32)
    out[:, 0] = p[:, 2]; out[:, 1] = p[:, 3]; out[:, 3] = p[:, 4]
    return out


def make_raw_scan(scene, pose, params, seed, n_azimuth=400, n_bins=3000, bin_size=0.0438, floor=(5.0, 45.0), peak=(90.0, 220.0), width_bins=4.0):
    """One RAW polar scan as the organised point cloud RadarPreprocessor::filterScan receives: float32 [n_azimuth * n_bins, 4] =
    (x, y, 0, intensity) in the sensor frame, azimuth-major, every range bin present (noise floor plus a peak around each wall hit)."""
    rng = np.random.default_rng(seed)
    _, rr = cast(scene, pose, n_azimuth, params.max_range, rng)
    az = (np.arange(n_azimuth) + 0.5) * (2 * math.pi / n_azimuth)
    az = np.where(az > math.pi, az - 2 * math.pi, az)                  # atan2 range; the half-step offset keeps rows away from +-pi
    r = (np.arange(n_bins) + 1.0) * bin_size                            # bin 0 would sit on the sensor (atan2(0, 0) = 0 for every azimuth)
    inten = rng.uniform(floor[0], floor[1], (n_azimuth, n_bins))
    hit = np.isfinite(rr)
    amp = rng.uniform(peak[0], peak[1], n_azimuth)
    bump = amp[:, None] * np.exp(-0.5 * ((r[None, :] - np.where(hit, rr, -1e9)[:, None]) / (width_bins * bin_size)) ** 2)
    inten = np.clip(inten + bump, 0.0, 255.0)
    out = np.zeros((n_azimuth, n_bins, 4), np.float32)
    out[:, :, 0] = r[None, :] * np.cos(az)[:, None]
    out[:, :, 1] = r[None, :] * np.sin(az)[:, None]
    out[:, :, 3] = inten
    return out.reshape(-1, 4)


def pose_to_se2(x, y, theta):
    """Sophus SE2d storage order [cos, sin, tx, ty]"""
    return np.array([math.cos(theta), math.sin(theta), x, y], np.float64)


def preset_scan_kwargs(params):
    """Beam geometry per preset so that cell occupancy lands where SURVEY §8a says it does."""
    if params.max_range >= 50.0:      # oxford-shape: 400 beams, 4.38 cm bins
        return dict(n_azimuth=400, bin_size=0.0438, half_bins=(3, 9), range_sigma=0.06, intensity_mean=90.0,
                    intensity_sigma=15.0, intensity_clip=(70.0, 255.0))
    # indoor/outdoor/mixed radar: dense short-range scans
    return dict(n_azimuth=400, bin_size=0.04, half_bins=(2, 5), range_sigma=0.02, intensity_mean=30.0, intensity_sigma=8.0,
                intensity_clip=(6.0, 120.0))


def scene_for(params, seed):
    hw = params.max_range
    if params.max_range >= 50.0:
        return Scene(seed, hw, n_walls=70, n_reflectors=60)
    return Scene(seed, hw, n_walls=45, n_reflectors=40, wall_len=(0.15, 0.7))
