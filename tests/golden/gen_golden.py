#!/usr/bin/env python
"""Generates tests/golden/*.json — known-answer vectors for the NDT hot path, computed INDEPENDENTLY of oracle/ and of the
CUDA kernels: 50-digit mpmath arithmetic straight from the mathematical definition in the reference functors, derivatives by
mpmath's high-order numerical differentiation of that definition (no closed form, no dual numbers).

The reference (IGMR-RWTH/RaNDT-SLAM) ships no golden vectors or tests for this path and cannot be built in this image
(SURVEY.md §8c), so these vectors pin the oracle ("parity unpinned" with respect to the reference's own binaries remains true
and is stated in DESIGN.md).  Definitions restated here, R/ = ros/ndt_radar_slam/ in the reference tree:

  residual   R/include/ndt_registration/ceres_residuals.h:541-547 (variant 0), :475-479 (1), :507-513 (2), :441-446 (3)
             r = sqrt(d^T B^-1 d), d = R mu_m + t - mu_f, B = R S_m R^T + S_f, inputs are the float32 cell statistics
  Barron     R/src/ndt_registration/ceres_loss_functions.cpp:19-39
  Welsch     R/src/ndt_registration/ceres_loss_functions.cpp:10-17
  labels     R/src/radar_preprocessing/grid.cpp:7-14  (float32 divide, int truncation toward zero)
  merge      R/include/ndt_representation/ndt_cell.h:133-142 (integer division of n1*n2/(n1+n2))

Run:  python tests/golden/gen_golden.py      (rewrites the json files next to this script; seeded, deterministic)
"""
import json
import math
import os

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 50


def f32(x):
    return float(np.float32(x))


def random_cell(rng):
    """float32 cell statistics in the range the voxeliser produces (xy covariance metres^2, intensity variance large)"""
    ang = rng.uniform(0, math.pi)
    l1, l2 = np.exp(rng.uniform(math.log(1e-3), math.log(2.0), 2))
    c, s = math.cos(ang), math.sin(ang)
    cov = np.zeros((3, 3))
    cov[0, 0] = c * c * l1 + s * s * l2
    cov[1, 1] = s * s * l1 + c * c * l2
    cov[0, 1] = cov[1, 0] = c * s * (l1 - l2)
    cov[2, 2] = rng.uniform(20, 400)
    cov[0, 2] = cov[2, 0] = rng.normal(0, 0.05)
    cov[1, 2] = cov[2, 1] = rng.normal(0, 0.05)
    cov32 = cov.astype(np.float32)
    # the reference's regularised covariances are asymmetric at float-ulp level: perturb one off-diagonal by one ulp
    cov32[0, 1] = np.nextafter(cov32[0, 1], np.float32(np.inf))
    mu = np.array([rng.uniform(-40, 40), rng.uniform(-40, 40), rng.uniform(70, 200)], np.float32)
    return np.concatenate([mu, cov32.reshape(9)]).astype(np.float32)


def mat3(v):
    return mp.matrix(3, 3, ) if v is None else mp.matrix([[mp.mpf(float(v[3 * i + j])) for j in range(3)] for i in range(3)])


def residual(variant, params, cm, cf):
    """params: mpf list.  variant 0/1: [c, s, tx, ty]; 2/3: [x, y, theta]"""
    mu_m = mp.matrix([mp.mpf(float(cm[i])) for i in range(3)])
    mu_f = mp.matrix([mp.mpf(float(cf[i])) for i in range(3)])
    Sm, Sf = mat3(cm[3:]), mat3(cf[3:])
    if variant == 0:
        th = mp.atan2(params[1], params[0]); c, s = mp.cos(th), mp.sin(th); tx, ty = params[2], params[3]
    elif variant == 1:
        c, s, tx, ty = params            # un-normalised complex number used as the rotation matrix
    else:
        th = params[2]
        th = th - 2 * mp.pi * mp.floor((th + mp.pi) / (2 * mp.pi))
        c, s = mp.cos(th), mp.sin(th); tx, ty = params[0], params[1]
    if variant in (0, 2):
        R = mp.matrix([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        t = mp.matrix([tx, ty, 0])
        d = R * mu_m + t - mu_f
        B = R * Sm * R.T + Sf
        q = mp.lu_solve(B, d)
        return mp.sqrt((d.T * q)[0])
    R = mp.matrix([[c, -s], [s, c]])
    d = R * mu_m[0:2, 0] + mp.matrix([tx, ty]) - mu_f[0:2, 0]
    B = R * Sm[0:2, 0:2] * R.T + Sf[0:2, 0:2]
    q = mp.lu_solve(B, d)
    return mp.sqrt((d.T * q)[0])


def jacobian(variant, params, cm, cf):
    n = len(params)
    out = []
    for i in range(n):
        def f(x, i=i):
            p = list(params); p[i] = x
            return residual(variant, p, cm, cf)
        out.append(mp.diff(f, params[i], h=mp.mpf(10) ** -12))
    return out


def gen_pairs(seed=2026, n=24):
    rng = np.random.default_rng(seed)
    cases = []
    for k in range(n):
        variant = k % 4
        cm = random_cell(rng)
        cf = random_cell(rng)
        # put the fixed cell near where the moving cell lands so that r is O(1..10)
        th = rng.uniform(-math.pi, math.pi) if k % 3 else rng.uniform(-0.1, 0.1)
        tx, ty = rng.uniform(-3, 3, 2)
        x = math.cos(th) * float(cm[0]) - math.sin(th) * float(cm[1]) + tx
        y = math.sin(th) * float(cm[0]) + math.cos(th) * float(cm[1]) + ty
        cf[0] = np.float32(x + rng.normal(0, 0.8)); cf[1] = np.float32(y + rng.normal(0, 0.8))
        cf[2] = np.float32(float(cm[2]) + rng.normal(0, 15))
        if variant <= 1:
            scale = 1.0 + (0.01 * rng.uniform(-1, 1) if k % 2 else 0.0)     # un-normalised (c, s) as ceres may hand over
            params = [scale * math.cos(th), scale * math.sin(th), tx, ty]
        else:
            params = [tx, ty, th + (2 * math.pi if k % 5 == 0 else 0.0)]
        pm = [mp.mpf(p) for p in params]
        r = residual(variant, pm, cm, cf)
        J = jacobian(variant, pm, cm, cf)
        cases.append(dict(variant=variant, params=[float(p) for p in params], cell_m=[float(v) for v in cm], cell_f=[float(v) for v in cf],
                          r=float(r), J=[float(j) for j in J]))
    return cases


def barron(s, a, alpha, mu):
    b = mu * a * a; c = 1 / b
    if alpha >= 2:
        return s, mp.mpf(1), mp.mpf(0)
    if abs(alpha) <= mp.mpf("0.05"):
        return b * mp.log(1 + s * c), 1 / (1 + s * c), -c / (1 + s * c) ** 2
    factor = abs(alpha - 2); e = alpha / 2; pre = b * factor / alpha; ts = 2 * c / factor
    u = s * ts + 1
    return pre * (u ** e - 1), pre * e * u ** (e - 1) * ts, pre * e * (e - 1) * u ** (e - 2) * ts * ts


def welsch(s, a, mu):
    b = mu * a * a; c = -1 / b
    ex = mp.exp(s * c)
    return b * (1 - ex), ex, c * ex


def gen_losses():
    out = []
    for (a, alpha, mu) in [(1.0, -2.0, 1.0), (1.5, -2.0, 1.69), (2.0, -1.0, 1.1), (2.0, -1.5, 1.21), (0.5, -2.0, 2.357947691),
                           (1.5, 0.03, 1.0), (1.5, 1.0, 2.0), (1.5, 2.5, 2.0)]:
        for s in (0.0, 1e-6, 0.3, 2.0, 17.5, 400.0):
            rho = barron(mp.mpf(s), mp.mpf(a), mp.mpf(alpha), mp.mpf(mu))
            out.append(dict(kind="barron", a=a, alpha=alpha, mu=mu, s=s, rho=[float(v) for v in rho]))
    for (a, mu) in [(1.5, 1.0), (2.0, 3.0)]:
        for s in (0.0, 0.3, 2.0, 17.5):
            rho = welsch(mp.mpf(s), mp.mpf(a), mp.mpf(mu))
            out.append(dict(kind="welsch", a=a, alpha=0.0, mu=mu, s=s, rho=[float(v) for v in rho]))
    return out


def gen_labels(seed=7):
    """Grid::cluster in numpy float32 (independent of the C++ restatement): label = int(x/res) + row*int(y/res)."""
    rng = np.random.default_rng(seed)
    cases = []
    for (max_range, resolution) in [(12.0, 0.5), (16.0, 1.2), (16.0, 1.0), (100.0, 3.5)]:
        n_clusters = int(math.pow(2.0 * max_range / resolution, 2))
        row = int(math.sqrt(n_clusters))
        res = np.float32(np.float32(max_range) * np.float32(2) / np.float32(row))
        pts = rng.uniform(-max_range, max_range, (40, 2)).astype(np.float32)
        pts[:6] = [[0.1, 0.1], [-0.1, 0.1], [0.1, -0.1], [-0.1, -0.1], [float(res), -float(res)], [-float(res) * 0.999, float(res) * 1.001]]
        lx = np.trunc(pts[:, 0] / res).astype(np.int64); ly = np.trunc(pts[:, 1] / res).astype(np.int64)
        labels = (lx + row * ly).astype(int)
        cases.append(dict(max_range=max_range, resolution=resolution, n_clusters=n_clusters, row=row, xy=pts.astype(float).tolist(),
                          labels=labels.tolist()))
    return cases


def gen_cell_stats(seed=11):
    """Cell::updateCell before regularisation is a population mean/covariance; after it the xy block has lambda0 >= 1e-3 lambda1
    and cov[2][2] += 1e-6.  High-precision values; the float32 paths must agree to 1e-5 relative (north_star tolerance)."""
    rng = np.random.default_rng(seed)
    cases = []
    for n, thin in [(12, False), (30, False), (25, True), (200, False)]:
        base = rng.uniform(-30, 30, 2)
        d = rng.normal(0, 1.0, (n, 2)) * ([1.5, 0.4] if not thin else [2.0, 1e-4])
        ang = rng.uniform(0, math.pi)
        Rm = np.array([[math.cos(ang), -math.sin(ang)], [math.sin(ang), math.cos(ang)]])
        xy = (d @ Rm.T + base).astype(np.float32)
        inten = rng.uniform(70, 200, n).astype(np.float32)
        P = [[mp.mpf(float(xy[i, 0])), mp.mpf(float(xy[i, 1])), mp.mpf(float(inten[i]))] for i in range(n)]
        mu = [sum(p[j] for p in P) / n for j in range(3)]
        cov = [[sum((p[a] - mu[a]) * (p[b] - mu[b]) for p in P) / n for b in range(3)] for a in range(3)]
        A = mp.matrix([[cov[0][0], cov[0][1]], [cov[1][0], cov[1][1]]])
        ev, V = mp.eigsy(A)
        lam = sorted([ev[0], ev[1]])
        order = [0, 1] if ev[0] <= ev[1] else [1, 0]
        l0 = max(lam[0], mp.mpf("0.001") * lam[1]); l1 = lam[1]
        v0 = V[:, order[0]]; v1 = V[:, order[1]]
        A2 = l0 * (v0 * v0.T) + l1 * (v1 * v1.T)
        reg = [[A2[0, 0], A2[0, 1], cov[0][2]], [A2[1, 0], A2[1, 1], cov[1][2]], [cov[2][0], cov[2][1], cov[2][2] + mp.mpf("0.000001")]]
        pts4 = np.zeros((n, 4), np.float32); pts4[:, :2] = xy; pts4[:, 3] = inten
        cases.append(dict(points=pts4.astype(float).tolist(), mean=[float(v) for v in mu], cov=[[float(v) for v in r] for r in reg],
                          floor_active=bool(lam[0] < mp.mpf("0.001") * lam[1])))
    return cases


def gen_merge(seed=13):
    """Cell::operator+= with the unsigned integer division (n1*n2)/(n1+n2) — exact rational arithmetic on float32 inputs."""
    rng = np.random.default_rng(seed)
    cases = []
    for (n1, n2) in [(11, 13), (40, 7), (3, 3), (100, 101)]:
        a = random_cell(rng); b = random_cell(rng)
        b[:3] = a[:3] + rng.normal(0, 0.5, 3).astype(np.float32)
        A = [mp.mpf(float(v)) for v in a]; Bc = [mp.mpf(float(v)) for v in b]
        w3 = (n1 * n2) // (n1 + n2)
        d = [A[i] - Bc[i] for i in range(3)]
        cov = [((n1 - 1) * A[3 + 3 * i + j] + (n2 - 1) * Bc[3 + 3 * i + j] + w3 * d[i] * d[j]) / (n1 + n2 - 1) for i in range(3) for j in range(3)]
        mu = [(A[i] * n1 + Bc[i] * n2) / (n1 + n2) for i in range(3)]
        cases.append(dict(n1=n1, n2=n2, a=[float(v) for v in a], b=[float(v) for v in b], mean=[float(v) for v in mu], cov=[float(v) for v in cov]))
    return cases


def gen_fused(seed=17):
    """Per-pose robustified normal equations as ceres accumulates them for P residual blocks that share one pose, each block with
    ScaledLoss(rho, weight) and the Corrector of the rho'' <= 0 branch (residual and Jacobian row scaled by sqrt(weight rho')):
    H = sum weight rho'(r^2) J^T J,  g = sum weight rho'(r^2) r J^T,  cost = 1/2 sum weight rho(r^2);  plus max r and sum r^2."""
    rng = np.random.default_rng(seed)
    cases = []
    for (variant, kind, a, alpha, mu, weight, n_pairs) in [(0, "barron", 1.0, -2.0, 1.0, 1.0, 7), (0, "barron", 2.0, -1.0, 1.21, 0.37, 9),
                                                             (2, "welsch", 1.5, 0.0, 2.0, 1.0, 6), (1, "barron", 1.5, -1.5, 1.0, 2.5, 5),
                                                             (3, "none", 1.0, 2.0, 1.0, 1.0, 4)]:
        th = rng.uniform(-0.4, 0.4); tx, ty = rng.uniform(-1, 1, 2)
        params = [math.cos(th), math.sin(th), tx, ty] if variant <= 1 else [tx, ty, th]
        pm = [mp.mpf(p) for p in params]
        n = len(params)
        H = [[mp.mpf(0)] * n for _ in range(n)]; g = [mp.mpf(0)] * n; cost = mp.mpf(0); max_r = mp.mpf(0); sum_sq = mp.mpf(0)
        cells_m, cells_f = [], []
        for _ in range(n_pairs):
            cm = random_cell(rng); cf = random_cell(rng)
            x = math.cos(th) * float(cm[0]) - math.sin(th) * float(cm[1]) + tx
            y = math.sin(th) * float(cm[0]) + math.cos(th) * float(cm[1]) + ty
            cf[0] = np.float32(x + rng.normal(0, 0.5)); cf[1] = np.float32(y + rng.normal(0, 0.5)); cf[2] = np.float32(float(cm[2]) + rng.normal(0, 10))
            r = residual(variant, pm, cm, cf); J = jacobian(variant, pm, cm, cf)
            sq = r * r
            if kind == "barron":
                rho = barron(sq, mp.mpf(a), mp.mpf(alpha), mp.mpf(mu))
            elif kind == "welsch":
                rho = welsch(sq, mp.mpf(a), mp.mpf(mu))
            else:
                rho = (sq, mp.mpf(1), mp.mpf(0))
            assert rho[2] <= 0
            w1 = mp.mpf(weight) * rho[1]
            for i in range(n):
                g[i] += w1 * r * J[i]
                for j in range(n):
                    H[i][j] += w1 * J[i] * J[j]
            cost += mp.mpf(weight) * rho[0] / 2
            max_r = max(max_r, r); sum_sq += sq
            cells_m.append([float(v) for v in cm]); cells_f.append([float(v) for v in cf])
        cases.append(dict(variant=variant, kind=kind, a=a, alpha=alpha, mu=mu, weight=weight, params=[float(p) for p in params], cells_m=cells_m,
                          cells_f=cells_f, H=[[float(v) for v in row] for row in H], g=[float(v) for v in g], cost=float(cost), max_r=float(max_r),
                          sum_sq=float(sum_sq)))
    return cases


def main():
    data = dict(
        note="generated by tests/golden/gen_golden.py (mpmath %s, 50 digits); independent of oracle/ and of the CUDA kernels" % mp.__version__,
        pairs=gen_pairs(), losses=gen_losses(), labels=gen_labels(), cell_stats=gen_cell_stats(), merges=gen_merge(), fused=gen_fused())
    with open(os.path.join(HERE, "ndt_golden.json"), "w") as f:
        json.dump(data, f, indent=0)
    print("wrote", os.path.join(HERE, "ndt_golden.json"), {k: len(v) for k, v in data.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
