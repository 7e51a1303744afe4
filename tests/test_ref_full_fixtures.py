"""Consumes tests/golden/ref_full_outputs.txt — the four NDT functors through ceres' autodiff, BarronLoss and estimateLoopConstraint's
solve, produced by the REFERENCE'S OWN headers with the real Eigen / Ceres / Sophus (oracle/ref_full/gen_fixtures.cpp, run where those
exist; see oracle/ref_full/README.md).  The file cannot be produced in the development image, so the comparisons are skipped until it is
present; what always runs is a self-check of the file format and of the comparison code on an oracle-written stand-in (tmp dir).

Bars once the file exists: r and J 1e-9 relative (north_star: 1e-5), rho/rho'/rho'' 1e-12, registration poses 1e-6 on the manifold and
on the gauge-invariant pose in raw-ambient mode to what ceres' function tolerance leaves (2e-3 rad / 5e-3 m), equal solve counts."""
import json
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUTS = os.path.join(ROOT, "tests", "golden", "ref_full_inputs.json")
OUTPUTS = os.path.join(ROOT, "tests", "golden", "ref_full_outputs.txt")


def load_inputs():
    return json.load(open(INPUTS))


def parse_outputs(path):
    tok = open(path).read().split()
    pos = 0

    def take(n):
        nonlocal pos
        v = tok[pos:pos + n]; pos += n
        return v
    out = {"functor": {}, "loss": None, "registrations": []}
    for _ in range(4):
        assert take(1) == ["functor"]
        variant, n_pose, n_pair = (int(x) for x in take(3))
        rows = np.array(take(6 * n_pose * n_pair), np.float64).reshape(n_pose, n_pair, 6)
        out["functor"][variant] = rows
    assert take(1) == ["loss"]
    nl, ns = (int(x) for x in take(2))
    out["loss"] = np.array(take(3 * nl * ns), np.float64).reshape(nl, ns, 3)
    assert take(1) == ["registrations"]
    nr = int(take(1)[0])
    for _ in range(nr):
        pair = []
        for _ in range(2):
            row = take(9)
            pair.append(dict(on_manifold=int(row[0]), pose=np.array(row[1:5], np.float64), score=float(row[5]), solves=int(row[6]), iterations=int(row[7]),
                             mu_first=float(row[8])))
        out["registrations"].append(pair)
    assert pos == len(tok)
    return out


def oracle_outputs(O, doc):
    """what the oracle computes for the same inputs, in the parsed layout"""
    cm = np.array(doc["pairs"]["cells_m"], np.float32); cf = np.array(doc["pairs"]["cells_f"], np.float32)
    n = len(cm)
    idx = np.arange(n, dtype=np.uint32)
    res = {"functor": {}, "registrations": []}
    for variant in range(4):
        poses = doc["poses4"] if variant <= 1 else doc["poses3"]
        rows = np.zeros((len(poses), n, 6))
        for s, pose in enumerate(poses):
            r, J = O.eval_pairs(variant, cm, cf, idx, idx, np.array(pose, np.float64), 0)
            rows[s, :, 0] = 1.0; rows[s, :, 1] = r; rows[s, :, 2:2 + J.shape[1]] = J
        res["functor"][variant] = rows
    ls = np.zeros((len(doc["loss"]["settings"]), len(doc["loss"]["s"]), 3))
    for l, (a, alpha, mu) in enumerate(doc["loss"]["settings"]):
        for k, s in enumerate(doc["loss"]["s"]):
            ls[l, k] = O.loss_eval(O.LOSS_BARRON, a, alpha, mu, 1.0, s)
    res["loss"] = ls
    sv = doc["solver"]
    for g in doc["registrations"]:
        gm = np.array(g["cells_m"], np.float32); gf = np.array(g["cells_f"], np.float32)
        pair = []
        for on_manifold in (0, 1):
            o = O.loop_constraint(gf, np.full(1, -1, np.int32), 1, 1, 1.0, 1.0, gm, np.array(g["pose0"]), 2, matcher_loss_scale=sv["loss_function_scale"],
                                  loop_scale=sv["loop_closure_scale"], alpha=sv["convexity"], divisor=sv["divisor"], max_gnc_steps=sv["loop_closure_gnc_steps"],
                                  max_iterations=sv["max_iteration"], on_manifold=bool(on_manifold),
                                  pairs=(np.array(g["pair_m"], np.uint32), np.array(g["pair_f"], np.uint32)))
            pair.append(dict(on_manifold=on_manifold, pose=o["pose"], score=o["score"], solves=o["gnc_solves"], iterations=o["iterations"], mu_first=o["mu_first"]))
        res["registrations"].append(pair)
    return res


def compare(ref, got, what):
    for variant in range(4):
        a, b = ref["functor"][variant], got["functor"][variant]
        assert a.shape == b.shape and np.all(a[..., 0] == 1.0), "%s: functor %d evaluation failed in the reference" % (what, variant)
        r_ref, r_got = a[..., 1], b[..., 1]
        assert np.max(np.abs(r_got - r_ref) / np.maximum(np.abs(r_ref), 1e-300)) < 1e-9, (what, variant)
        scale = np.maximum(np.max(np.abs(a[..., 2:]), axis=-1, keepdims=True), 1e-30)
        assert np.max(np.abs(b[..., 2:] - a[..., 2:]) / scale) < 1e-8, (what, variant)
    if got.get("loss") is not None:
        assert np.max(np.abs(got["loss"] - ref["loss"]) / np.maximum(np.abs(ref["loss"]), 1e-300)) < 1e-12
    for pr, pg in zip(ref["registrations"], got["registrations"]):
        for r, g in zip(pr, pg):
            assert r["solves"] == g["solves"] and abs(r["mu_first"] - g["mu_first"]) <= 1e-9 * abs(r["mu_first"])
            th_r, th_g = math.atan2(r["pose"][1], r["pose"][0]), math.atan2(g["pose"][1], g["pose"][0])
            if r["on_manifold"]:
                assert abs(th_r - th_g) < 1e-6 and np.max(np.abs(r["pose"][2:] - g["pose"][2:])) < 1e-6, what
                assert abs(r["score"] - g["score"]) <= 1e-6 * abs(r["score"])
            else:
                assert abs(th_r - th_g) < 2e-3 and np.max(np.abs(r["pose"][2:] - g["pose"][2:])) < 5e-3, what


def write_outputs(path, res):
    with open(path, "w") as f:
        for variant in range(4):
            rows = res["functor"][variant]
            f.write("functor %d %d %d\n" % (variant, rows.shape[0], rows.shape[1]))
            for row in rows.reshape(-1, 6):
                f.write("%d %s\n" % (int(row[0]), " ".join("%.17g" % x for x in row[1:])))
        ls = res["loss"]
        f.write("loss %d %d\n" % ls.shape[:2])
        for row in ls.reshape(-1, 3):
            f.write(" ".join("%.17g" % x for x in row) + "\n")
        f.write("registrations %d\n" % len(res["registrations"]))
        for pair in res["registrations"]:
            for r in pair:
                f.write("%d %s %.17g %d %d %.17g\n" % (r["on_manifold"], " ".join("%.17g" % x for x in r["pose"]), r["score"], r["solves"], r["iterations"], r["mu_first"]))


def test_fixture_format_and_comparison_self_check(oracle, tmp_path):
    doc = load_inputs()
    res = oracle_outputs(oracle, doc)
    p = tmp_path / "stand_in_outputs.txt"
    write_outputs(str(p), res)
    back = parse_outputs(str(p))
    compare(back, res, "self-check")
    assert back["functor"][0].shape == (4, 48, 6) and len(back["registrations"]) == 3


@pytest.mark.skipif(not os.path.exists(OUTPUTS), reason="tests/golden/ref_full_outputs.txt not generated (needs the reference's Eigen/Ceres/Sophus: oracle/ref_full/README.md)")
def test_oracle_matches_full_reference_fixtures(oracle):
    compare(parse_outputs(OUTPUTS), oracle_outputs(oracle, load_inputs()), "oracle vs reference")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(OUTPUTS), reason="tests/golden/ref_full_outputs.txt not generated (needs the reference's Eigen/Ceres/Sophus: oracle/ref_full/README.md)")
def test_cuda_matches_full_reference_fixtures(gpu_ctx):
    from randt_slam_b200 import capi
    doc = load_inputs()
    ref = parse_outputs(OUTPUTS)
    cm = np.array(doc["pairs"]["cells_m"], np.float32); cf = np.array(doc["pairs"]["cells_f"], np.float32)
    n = len(cm)
    idx = np.arange(n, dtype=np.uint32)
    got = {"functor": {}, "loss": None, "registrations": []}
    for variant in range(4):
        poses = np.array(doc["poses4"] if variant <= 1 else doc["poses3"], np.float64)
        S = len(poses)
        prob = gpu_ctx.problem_create(cm, cf, np.tile(idx, S), np.tile(idx, S), np.arange(S + 1, dtype=np.uint32) * n)
        r, J = prob.eval_emit(poses, variant=variant)
        rows = np.zeros((S, n, 6)); rows[..., 0] = 1.0; rows[..., 1] = r.reshape(S, n); rows[..., 2:2 + J.shape[1]] = J.reshape(S, n, -1)
        got["functor"][variant] = rows
        prob.close()
    sv = doc["solver"]
    for g in doc["registrations"]:
        gm = np.array(g["cells_m"], np.float32); gf = np.array(g["cells_f"], np.float32)
        prob = gpu_ctx.problem_create(gm, gf, np.array(g["pair_m"], np.uint32), np.array(g["pair_f"], np.uint32), [0, len(g["pair_m"])])
        pair = []
        for on_manifold in (0, 1):
            loss = capi.make_loss(capi.LOSS_BARRON, sv["loop_closure_scale"], sv["convexity"], 1.0, 1.0)
            opt = capi.solver_options(use_manifold=on_manifold, gnc_loss_scale=sv["loss_function_scale"], gnc_divisor=sv["divisor"],
                                      gnc_max_steps=sv["loop_closure_gnc_steps"], max_num_iterations=sv["max_iteration"])
            out, res = prob.register_batch(np.array(g["pose0"])[None], loss, opt)
            pair.append(dict(on_manifold=on_manifold, pose=out[0], score=res[0, capi.REG_SCORE], solves=int(res[0, capi.REG_GNC_SOLVES]),
                             iterations=int(res[0, capi.REG_ITERATIONS]), mu_first=res[0, capi.REG_MU_FIRST]))
        got["registrations"].append(pair)
        prob.close()
    compare(ref, got, "cuda vs reference")


# ---- the host factors of estimateTransformCeres' window problem (oracle/ref_full/gen_window_fixtures.cpp) -------------------------------
W_INPUTS = os.path.join(ROOT, "tests", "golden", "ref_full_window_inputs.json")
W_OUTPUTS = os.path.join(ROOT, "tests", "golden", "ref_full_window_outputs.txt")
W_ORDER = [(0, 1), (0, 0), (1, 1), (1, 0)]      # (kind, manifold): motion SE2, motion vector, imu SE2, imu vector


def parse_window_outputs(path):
    tok = open(path).read().split()
    pos = 0

    def take(n):
        nonlocal pos
        v = tok[pos:pos + n]; pos += n
        return v
    assert take(1) == ["cases"]
    n = int(take(1)[0])
    cases = []
    for _ in range(n):
        blocks = {}
        for _ in range(4):
            assert take(1) == ["block"]
            kind, manifold, nres, ok = (int(x) for x in take(4))
            rows = np.array(take(21 * nres), np.float64).reshape(nres, 21)
            blocks[(kind, manifold)] = dict(ok=ok, res=rows[:, 0].copy(), jac=rows[:, 1:].copy())
        cases.append(blocks)
    solves = []
    if pos < len(tok):
        assert take(1) == ["solves"]
        ns = int(take(1)[0])
        for _ in range(ns):
            assert take(1) == ["solve"]
            row = take(6)
            Wn = int(row[0])
            states = np.array(take(14 * (Wn + 1)), np.float64).reshape(Wn + 1, 14)
            solves.append(dict(W=Wn, solves=int(row[1]), iterations=int(row[2]), final_cost=float(row[3]), mu_first=float(row[4]), max_residual=float(row[5]),
                               states=states))
    assert pos == len(tok)
    return cases, solves


def parse_window_solve_inputs(path):
    """the "solves" section of ref_full_window_inputs.txt (whole window problems with frozen pair lists)"""
    tok = open(path).read().split()
    pos = tok.index("solves")

    def take(n):
        nonlocal pos
        v = tok[pos:pos + n]; pos += n
        return v

    def key(k):
        assert take(1) == [k], k
    key("solves")
    out = []
    for _ in range(int(take(1)[0])):
        key("solve")
        Wn, nm, nf, Pn = (int(x) for x in take(4))
        g = dict(W=Wn)
        key("params"); g["params16"] = np.array(take(16), np.float64)
        key("sqrtI"); g["sqrtI"] = np.array(take(64), np.float64)
        key("tolerances"); g["tolerances"] = np.array(take(3), np.float64)
        key("n_cells"); g["n_cells"] = int(take(1)[0])
        key("states"); g["states"] = np.array(take(14 * (Wn + 1)), np.float64).reshape(Wn + 1, 14)
        key("imu"); g["imu"] = np.array(take(Wn), np.float64)
        key("cells_m"); g["cells_m"] = np.array(take(12 * nm), np.float64).astype(np.float32).reshape(nm, 12)
        key("cells_f"); g["cells_f"] = np.array(take(12 * nf), np.float64).astype(np.float32).reshape(nf, 12)
        key("pair_m"); g["pair_m"] = np.array(take(Pn), np.uint32)
        key("pair_f"); g["pair_f"] = np.array(take(Pn), np.uint32)
        key("seg_off"); g["seg_off"] = np.array(take(Wn + 1), np.uint32)
        out.append(g)
    assert pos == len(tok)
    return out


def oracle_window_solves(O, problems):
    res = []
    for g in problems:
        q = np.concatenate([g["params16"], g["sqrtI"]])
        q[14] = 0 if q[7] else 2                       # the oracle reads the functor variant there
        tol = tuple(g["tolerances"]) if g["tolerances"][0] > 0 else None
        st, _, info = O.window_solve(g["states"], q, g["states"][-2, :4], g["cells_m"], g["cells_f"], g["pair_m"], g["pair_f"], g["seg_off"], g["n_cells"],
                                     imu=g["imu"], tolerances=tol)
        if info["rejected"]:
            raise AssertionError("fixture case hits the rejection gate")
        # estimateTransformCeres leaves the stale representation of the newest state to its caller: compare what ceres wrote
        res.append(dict(W=g["W"], solves=int(info["gnc_solves"]), iterations=int(info["total_iterations"]), final_cost=info["final_cost"],
                        mu_first=info["mu_first"], max_residual=info["max_residual"], states=st))
    return res


def compare_window_solves(ref, got, problems, what):
    assert len(ref) == len(got) == len(problems)
    for r, g, prob in zip(ref, got, problems):
        manifold = prob["params16"][7] != 0
        fixed_steps = prob["tolerances"][0] > 0
        assert r["solves"] == g["solves"] and abs(r["mu_first"] - g["mu_first"]) <= 1e-9 * abs(r["mu_first"]), what
        assert abs(r["max_residual"] - g["max_residual"]) <= 1e-9 * r["max_residual"], what
        cols = list(range(0, 4)) + list(range(7, 13)) if manifold else list(range(4, 13))      # what ceres optimised (pose | pos, rot), velocities, bias
        d = np.max(np.abs(r["states"][:, cols] - g["states"][:, cols]))
        if fixed_steps:
            # the same number of trust-region steps on both sides: DENSE_QR against the damped normal equations, nothing else differs
            assert r["iterations"] == g["iterations"] and d < 1e-5, (what, d)
            assert abs(r["final_cost"] - g["final_cost"]) <= 1e-7 * r["final_cost"]
        else:
            # ceres' default tolerances: the solve stops along a weakly determined valley (Window)
            assert d < 5e-2 and abs(r["final_cost"] - g["final_cost"]) <= 1e-4 * r["final_cost"], (what, d)


def oracle_window_outputs(O, doc):
    sq = np.array(doc["sqrtI"], np.float64).reshape(64)
    out = []
    for c in doc["cases"]:
        blocks = {}
        for kind, manifold in W_ORDER:
            r, J = O.factor_block(kind, manifold, c["a"], c["b"], sq, c["imu_rot"], doc["weight_imu"], doc["weight_imu_bias"])
            blocks[(kind, manifold)] = dict(ok=1, res=r, jac=J)
        out.append(blocks)
    return out


def write_window_outputs(path, cases, solves=()):
    with open(path, "w") as f:
        f.write("cases %d\n" % len(cases))
        for blocks in cases:
            for kind, manifold in W_ORDER:
                b = blocks[(kind, manifold)]
                f.write("block %d %d %d %d\n" % (kind, manifold, len(b["res"]), b["ok"]))
                for r, row in zip(b["res"], b["jac"]):
                    f.write("%.17g %s\n" % (r, " ".join("%.17g" % x for x in row)))
        if solves:
            f.write("solves %d\n" % len(solves))
            for g in solves:
                f.write("solve %d %d %d %.17g %.17g %.17g\n" % (g["W"], g["solves"], g["iterations"], g["final_cost"], g["mu_first"], g["max_residual"]))
                for row in g["states"]:
                    f.write(" ".join("%.17g" % x for x in row) + "\n")


def compare_window(ref, got, what, tol=1e-9):
    assert len(ref) == len(got)
    for i, (a, b) in enumerate(zip(ref, got)):
        for key in W_ORDER:
            assert a[key]["ok"] == 1, "%s: the reference failed to evaluate case %d block %s" % (what, i, key)
            sr = max(np.max(np.abs(a[key]["res"])), 1e-12); sj = max(np.max(np.abs(a[key]["jac"])), 1e-12)
            assert np.max(np.abs(a[key]["res"] - b[key]["res"])) <= tol * sr, (what, i, key)
            assert np.max(np.abs(a[key]["jac"] - b[key]["jac"])) <= tol * sj, (what, i, key)


def tangent_system(blocks, b_pose, manifold, cv, use_imu):
    """cost, g, H of the two-state window [a, b] from the per-block residuals / ambient Jacobians, in the product's parameter order
    (a: lin_vel, rot_vel, [lin_acc]; b: pose (through Sophus::Manifold<SE2>'s PlusJacobian), lin_vel, rot_vel, [lin_acc], [imu_bias])"""
    cols = []                                     # tangent column -> ambient combination
    def unit(slot):
        e = np.zeros(20); e[slot] = 1.0; return e
    cols += [unit(4), unit(5), unit(6)] + ([] if cv else [unit(7), unit(8)])
    if manifold:
        c, s = b_pose[0], b_pose[1]
        Pj = np.array([[0, 0, -s], [0, 0, c], [c, -s, 0], [s, c, 0]], np.float64)
        for t in range(3):
            e = np.zeros(20); e[10:14] = Pj[:, t]; cols.append(e)
    else:
        cols += [unit(10), unit(11), unit(12)]
    cols += [unit(14), unit(15), unit(16)] + ([] if cv else [unit(17), unit(18)]) + ([unit(19)] if use_imu else [])
    T = np.array(cols).T                          # [20, nt]
    keys = [(0, int(manifold))] + ([(1, int(manifold))] if use_imu else [])
    res = np.concatenate([blocks[k]["res"] for k in keys]); Jt = np.concatenate([blocks[k]["jac"] @ T for k in keys])
    return 0.5 * float(res @ res), Jt.T @ res, Jt.T @ Jt


def check_host_factors(cases, doc, tol):
    """the product's host layer (randt_hostapi_window_factors) on every case as a two-state window, against the blocks"""
    from randt_slam_b200 import hostapi
    for c, blocks in zip(doc["cases"], cases):
        for manifold in (True, False):
            for cv in (True, False):
                for use_imu in (False, True):
                    q = hostapi.window_params(manifold=manifold, constant_velocity=cv, use_imu=use_imu, weight_imu=doc["weight_imu"],
                                              weight_imu_bias=doc["weight_imu_bias"], covariance_scaling_factor=1.0)
                    q[16:] = np.array(doc["sqrtI"], np.float64).reshape(64)
                    cost, g, H = hostapi.window_factors(np.array([c["a"], c["b"]]), q, [c["imu_rot"]])
                    c0, g0, H0 = tangent_system(blocks, c["b"], manifold, cv, use_imu)
                    assert abs(cost - c0) <= tol * max(c0, 1e-30)
                    assert np.max(np.abs(g - g0)) <= tol * max(np.max(np.abs(g0)), 1e-30) and np.max(np.abs(H - H0)) <= tol * max(np.max(np.abs(H0)), 1e-30)


def test_window_fixture_format_and_comparison_self_check(oracle, tmp_path):
    doc = json.load(open(W_INPUTS))
    res = oracle_window_outputs(oracle, doc)
    problems = parse_window_solve_inputs(W_INPUTS.replace(".json", ".txt"))
    solved = oracle_window_solves(oracle, problems)
    p = tmp_path / "stand_in_window_outputs.txt"
    write_window_outputs(str(p), res, solved)
    back, back_solves = parse_window_outputs(str(p))
    compare_window(back, res, "self-check", tol=1e-15)
    compare_window_solves(back_solves, solved, problems, "self-check")
    assert len(problems) == 4 and [g["W"] for g in problems] == [3, 3, 3, 2] and solved[1]["iterations"] == 2 * 13
    assert all(np.max(np.abs(g["states"] - q["states"])) > 1e-3 for g, q in zip(solved, problems))      # the solves moved the states
    assert len(back) == 24 and back[0][(0, 1)]["jac"].shape == (8, 20) and back[0][(1, 0)]["jac"].shape == (2, 20)
    # the vector functors never see the pose block, the SE(2) functors never see pos / rot ... both share the velocity columns
    assert np.all(back[0][(1, 0)]["jac"][:, [0, 1, 3, 4, 5, 6, 7, 8]] == 0.0)
    check_host_factors(back, doc, 1e-10)        # (g sums terms of opposite sign: a few 1e-12 of its largest entry)


@pytest.mark.skipif(not os.path.exists(W_OUTPUTS), reason="tests/golden/ref_full_window_outputs.txt not generated (needs the reference's Eigen/Ceres/Sophus: oracle/ref_full/README.md)")
def test_window_factors_match_full_reference_fixtures(oracle):
    doc = json.load(open(W_INPUTS))
    ref, ref_solves = parse_window_outputs(W_OUTPUTS)
    compare_window(ref, oracle_window_outputs(oracle, doc), "oracle vs reference")
    check_host_factors(ref, doc, 1e-9)
    problems = parse_window_solve_inputs(W_INPUTS.replace(".json", ".txt"))
    compare_window_solves(ref_solves, oracle_window_solves(oracle, problems), problems, "oracle vs reference (estimateTransformCeres)")
