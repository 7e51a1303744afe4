#!/usr/bin/env python
"""Writes tests/golden/ref_full_inputs.json: the seeded inputs the full-reference harness (oracle/ref_full/gen_fixtures.cpp) evaluates with
the reference's own headers on a machine that has its dependencies (Eigen 3.3.7, Ceres 2.1.0, Sophus 1.22.10: /root/reference/Dockerfile).
Inputs are generated here (numpy, CPU oracle for the cell tables and pair lists) so that the C++ side needs no generator of its own.
usage: python oracle/ref_full/gen_inputs.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from randt_slam_b200 import params as P, synth  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    cm = H.random_cells(rng, 48); cf = H.random_cells(rng, 48)
    cf[:, :2] = cm[:, :2] + rng.normal(0, 0.3, (48, 2)).astype(np.float32)          # residuals of O(1)
    poses4 = [synth.pose_to_se2(*rng.uniform(-1, 1, 3) * [0.5, 0.5, 0.3]).tolist() for _ in range(4)]
    poses4[1] = (np.array(poses4[1]) * [1.002, 1.002, 1, 1]).tolist()                  # un-normalised complex part (raw ambient block)
    poses3 = rng.uniform(-1, 1, (4, 3)).tolist()
    regs = []
    p = P.OXFORD
    for seed, guess in ((3, (0.5, -0.3, 0.02)), (4, (0.2, -0.1, 0.0)), (5, (0.9, -0.6, 0.05))):
        c = H.make_registration_case(O, p, seed=seed, n_fixed_scans=4, true_pose=(0.6, -0.4, 0.03), guess=guess)
        regs.append(dict(cells_m=c["moving"]["cells"].astype(np.float64).tolist(), cells_f=c["fixed"]["cells"].astype(np.float64).tolist(),
                         pair_m=c["im"].tolist(), pair_f=c["jf"].tolist(), pose0=c["pose0"].tolist()))
    doc = dict(
        note="cells are float32 values written as doubles: [mean x, y, i, cov row-major 9]; poses4 = Sophus SE2d::data() order [cos, sin, tx, ty]",
        pairs=dict(cells_m=cm.astype(np.float64).tolist(), cells_f=cf.astype(np.float64).tolist()),
        poses4=poses4, poses3=poses3,
        loss=dict(s=[0.0, 1e-6, 0.03, 0.5, 1.0, 4.0, 37.5, 900.0], settings=[[1.0, -2.0, 1.0], [2.0, -1.0, 3.3], [2.0, -1.5, 1.21], [1.5, 0.03, 1.0], [1.5, 1.0, 2.0], [1.5, 2.5, 2.0]]),
        registrations=regs,
        solver=dict(loss_function_scale=p.loss_function_scale, loop_closure_scale=p.loop_closure_scale, convexity=p.loss_function_convexity,
                    divisor=p.gnc_control_parameter_divisor, loop_closure_gnc_steps=p.loop_closure_gnc_steps, gnc_steps=p.gnc_steps, max_iteration=p.max_iteration))
    out = os.path.join(ROOT, "tests", "golden", "ref_full_inputs.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print(out, os.path.getsize(out), "bytes")
    # the same as whitespace-separated text (what the C++ harness reads: no JSON parser needed there)
    txt = os.path.join(ROOT, "tests", "golden", "ref_full_inputs.txt")
    r17 = lambda v: " ".join("%.17g" % x for x in v)
    with open(txt, "w") as f:
        f.write("pairs %d\n" % len(cm))
        for a, b in zip(doc["pairs"]["cells_m"], doc["pairs"]["cells_f"]):
            f.write(r17(a) + " " + r17(b) + "\n")
        f.write("poses4 %d\n" % len(poses4)); [f.write(r17(q) + "\n") for q in poses4]
        f.write("poses3 %d\n" % len(poses3)); [f.write(r17(q) + "\n") for q in poses3]
        f.write("loss_s %d\n%s\n" % (len(doc["loss"]["s"]), r17(doc["loss"]["s"])))
        f.write("loss_settings %d\n" % len(doc["loss"]["settings"])); [f.write(r17(q) + "\n") for q in doc["loss"]["settings"]]
        sv = doc["solver"]
        f.write("solver %s\n" % r17([sv["loss_function_scale"], sv["loop_closure_scale"], sv["convexity"], sv["divisor"], sv["loop_closure_gnc_steps"], sv["gnc_steps"], sv["max_iteration"]]))
        f.write("registrations %d\n" % len(regs))
        for g in regs:
            f.write("registration %d %d %d\n" % (len(g["cells_m"]), len(g["cells_f"]), len(g["pair_m"])))
            [f.write(r17(c) + "\n") for c in g["cells_m"]]; [f.write(r17(c) + "\n") for c in g["cells_f"]]
            f.write(" ".join(str(i) for i in g["pair_m"]) + "\n" + " ".join(str(i) for i in g["pair_f"]) + "\n" + r17(g["pose0"]) + "\n")
    print(txt, os.path.getsize(txt), "bytes")
    window_inputs(rng)


def window_inputs(rng):
    """seeded state pairs for the host factors of estimateTransformCeres' window problem (gen_window_fixtures.cpp): poses near each other
    along a drive (small and large rotation differences, identical stamps for the 0.2 s clamp, a rotation difference exactly 0 for the
    series branches of SE2::log / exp), velocities, accelerations, IMU biases"""
    import math
    sqrtI = 0.3 * np.diag([1.0, 1.0, 10.0, 1.0, 3.0, 0.1, 20.0, 60.0])
    sqrtI[0, 3] = 0.05; sqrtI[4, 1] = -0.02            # off-diagonal entries: applyOnTheLeft is a full matrix product
    cases = []
    for i in range(24):
        th = rng.uniform(-3.0, 3.0); x, y = rng.uniform(-20, 20, 2)
        v = rng.normal([4.0, 0.2], [1.5, 0.5]); om = rng.normal(0.1, 0.2); acc = rng.normal(0, 0.5, 2)
        dt = [0.25, 0.2486, 0.0, 0.05, 0.4][i % 5]
        d = rng.normal(0, [0.3, 0.3, 0.05])
        if i == 3:
            d[2] = 0.0; om = 0.0                          # pose_pred^-1 * pose_1 with rotation exactly 0, exp() of a screw without rotation
        if i == 7:
            d[2] = 2.5                                    # a large rotation residual
        dte = max(dt, 0.2)
        th1 = th + om * dte + d[2]; x1 = x + math.cos(th) * v[0] * dte - math.sin(th) * v[1] * dte + d[0]; y1 = y + math.sin(th) * v[0] * dte + math.cos(th) * v[1] * dte + d[1]
        a = [math.cos(th), math.sin(th), x, y, x, y, th, v[0], v[1], om, acc[0], acc[1], rng.normal(0, 0.01), 10.0]
        b = [math.cos(th1), math.sin(th1), x1, y1, x1, y1, th1, v[0] + rng.normal(0, 0.3), v[1] + rng.normal(0, 0.3), om + rng.normal(0, 0.05),
             acc[0] + rng.normal(0, 0.1), acc[1] + rng.normal(0, 0.1), rng.normal(0, 0.01), 10.0 + dt]
        if i == 11:
            a[0] *= 1.001; a[1] *= 1.001                  # a complex part that is not exactly of unit length (what LM steps leave behind)
        cases.append(dict(a=a, b=b, imu_rot=float(om * dt + rng.normal(0, 0.01))))
    doc = dict(note="state14 = cos, sin, tx, ty, pos_x, pos_y, rot, vx, vy, omega, ax, ay, imu_bias, stamp", sqrtI=sqrtI.tolist(), weight_imu=64.0,
               weight_imu_bias=750.0, cases=cases)
    solves = window_solves()      # text file only (the cell tables are float32 values: nine digits round-trip them)
    out = os.path.join(ROOT, "tests", "golden", "ref_full_window_inputs.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    txt = os.path.join(ROOT, "tests", "golden", "ref_full_window_inputs.txt")
    r17 = lambda v: " ".join("%.17g" % x for x in v)
    with open(txt, "w") as f:
        f.write("sqrtI\n" + "\n".join(r17(row) for row in sqrtI.tolist()) + "\n")
        f.write("imu_weights %s\n" % r17([doc["weight_imu"], doc["weight_imu_bias"]]))
        f.write("cases %d\n" % len(cases))
        for c in cases:
            f.write("case\n" + r17(c["a"]) + "\n" + r17(c["b"]) + "\n" + r17([c["imu_rot"]]) + "\n")
        r9 = lambda v: " ".join("%.9g" % x for x in v)
        f.write("solves %d\n" % len(solves))
        for g in solves:
            f.write("solve %d %d %d %d\n" % (g["W"], len(g["cells_m"]), len(g["cells_f"]), len(g["pair_m"])))
            f.write("params " + r17(g["params16"]) + "\n")
            f.write("sqrtI\n" + "\n".join(r17(row) for row in g["sqrtI"]) + "\n")
            f.write("tolerances " + r17(g["tolerances"]) + "\n")
            f.write("n_cells %d\n" % g["n_cells"])
            f.write("states\n" + "\n".join(r17(row) for row in g["states"]) + "\n")
            f.write("imu " + r17(g["imu"]) + "\n")
            f.write("cells_m\n" + "\n".join(r9(c) for c in g["cells_m"]) + "\n")
            f.write("cells_f\n" + "\n".join(r9(c) for c in g["cells_f"]) + "\n")
            f.write("pair_m " + " ".join(str(i) for i in g["pair_m"]) + "\n")
            f.write("pair_f " + " ".join(str(i) for i in g["pair_f"]) + "\n")
            f.write("seg_off " + " ".join(str(i) for i in g["seg_off"]) + "\n")
    print(txt, os.path.getsize(txt), "bytes")


def window_solves():
    """whole window problems for gen_window_fixtures.cpp's "solves" section: the cases of tests/test_window_gpu.py (Oxford as shipped at
    ceres' default tolerances and with a fixed number of steps; constant acceleration + IMU + two fixed maps; the vector parametrisation)
    with their NDT pair lists frozen by the oracle's association"""
    from randt_slam_b200 import hostapi, workloads as W
    from tests.test_window_gpu import make_case
    p = P.OXFORD
    k = p.n_results_nn_lookup
    out = []
    for manifold, cv, use_imu, n_fixed, Wn, fixed_steps in ((True, True, False, 1, 3, False), (True, True, False, 1, 3, True), (True, False, True, 2, 3, True),
                                                           (False, True, True, 1, 2, True)):
        rng = np.random.default_rng(11 + 7 * Wn + n_fixed)
        fixed, fixed_se2, window, st = make_case(p, 120, n_fixed, Wn, rng)
        imu = [0.01 + rng.normal(0, 0.002) for _ in range(Wn)]
        q = hostapi.window_params(k=k, gnc_steps=p.gnc_steps, loss_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                                  divisor=p.gnc_control_parameter_divisor, ndt_weight=p.ndt_weight, manifold=manifold, constant_velocity=cv,
                                  use_imu=use_imu, weight_imu=64.0, weight_imu_bias=750.0, covariance_scaling_factor=0.01,
                                  max_iteration=12 if fixed_steps else p.max_iteration)
        w = W.oracle_window_problem(O, p, fixed, fixed_se2, window, st, k)
        out.append(dict(W=Wn, params16=q[:16].tolist(), sqrtI=q[16:].reshape(8, 8).tolist(), tolerances=[1e-30] * 3 if fixed_steps else [0.0] * 3,
                        n_cells=int(w["n_cells"]), states=st.tolist(), imu=imu, cells_m=w["cells_m"].astype(np.float64).tolist(),
                        cells_f=w["cells_f"].astype(np.float64).tolist(), pair_m=w["im"].tolist(), pair_f=w["jf"].tolist(), seg_off=w["seg_off"].tolist()))
    return out


if __name__ == "__main__":
    main()
