"""bench.py's bookkeeping that needs no GPU: the algorithmic-byte formula of SURVEY §8d, the DRAM traffic figure taken from the committed
ncu capture, the workload naming, and the committed bench line itself (contract keys, roofline arithmetic)."""
import json
import os

import bench
from randt_slam_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_formula():
    st = dict(n_m=1000, n_f_referenced=3000, pairs=2000, segments=10)
    assert bench.algorithmic_bytes(st) == 48 * (1000 + 3000) + 8 * 2000 + 32 * 10 + 192 * 10


def test_traffic_comes_from_the_final_capture():
    t = bench.measured_traffic(2941151)
    txt = open(os.path.join(ROOT, "profiles", "r02_k3_fused_final_ncu_summary.txt")).read()     # the latest round's `final` capture wins
    import re
    rd = float(re.search(r"dram__bytes_read\.sum\s+Mbyte\s+([\d.]+)", txt).group(1)); wr = float(re.search(r"dram__bytes_write\.sum\s+Mbyte\s+([\d.]+)", txt).group(1))
    assert abs(t - (rd + wr) * 1e6) < 1.0
    assert bench.measured_traffic(12345) is None


def test_workload_names():
    assert "Oxford-shape" in bench.workload_name(P.OXFORD) and "oxford params" in bench.workload_name(P.OXFORD)
    assert "indoor params" in bench.workload_name(P.INDOOR)


def test_round2_bench_line_carries_every_config():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_final.json")))
    assert d["metric"] == bench.METRIC and d["roofline"]["frac"] >= 0.60 and d["gpu_launches"] == d["steps"]
    # DRAM traffic of the dominant kernel no longer exceeds its algorithmic bytes (round 1: 1.25 x)
    assert d["roofline"]["traffic"] <= 1.05 * d["roofline"]["algorithmic_bytes_per_launch"]
    c = d["configs"]
    assert set(c) >= {"c0", "c1", "c2", "c3", "c4"}
    assert c["c0"]["max_rel_err_r_vs_oracle"] < 1e-9 and c["c2"]["gpu_vs_oracle_max_rel_err_first_problem"] < 1e-9
    assert c["c2"]["all_pairs"]["pairs_used_first_problem"] == 2000 * 8000 and "not run" in c["c2"]["se3_imu"]
    assert c["c3"]["registrations"] == 256 and c["c3"]["failed"] == 0 and c["c3"]["scaling"] == "strong" and len(c["c3"]["table_sha256"]) == 64
    assert c["c4"]["scans"] == 8609 and c["c4"]["scans_per_s"] > 3 * c["c4"]["cpu_baseline"]["scans_per_s"] * 0.9
    r = d["registrations"]
    assert r["launches_per_batch"] == 1 and r["failed"] == 0 and r["value"] > 2.5e6
    e = d["e2e"]
    assert e["d2h_bytes_per_step"] == d["config"]["problems_per_gpu"] * 8 * 10 and e["results_equal_blocking_call"]      # 80-byte basis records
    assert e["max_rel_difference_vs_blocking_call"] < 1e-13 and e["value"] > 0.9 * d["value"]
    # the drive through Matcher::estimateTransformCeres (window of three states, motion-model factors)
    w = c["c4"]["window_odometry"]
    assert w["scans"] == 1000 and w["rejected_estimates"] == 0 and w["max_position_error_m"] < 1.0
    assert w["scans_per_s"] > 3 * w["cpu_baseline"]["scans_per_s"] and abs(w["final_speed_m_per_s"] - w["true_speed_m_per_s"]) < 0.1
    # the literal batch gives the same table on 1, 2, 4 and 8 GPUs
    s = json.load(open(os.path.join(ROOT, "profiles", "r02_scaling_1_2_4_8.json")))
    assert len({s[n]["configs_c3"]["table_sha256"] for n in ("1", "2", "4", "8")}) == 1
    assert all(s[n]["e2e"]["d2h_bytes_per_step"] == 16384 * 80 for n in ("1", "2", "4", "8"))


def test_committed_bench_line_keeps_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["vs_baseline"] is None and d["warmup"] >= 3
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["frac"] >= 0.60
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(d["value"] - d["config"]["pairs_per_gpu"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"] and e["results_equal_blocking_call"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and d["value"] / c["single_thread_value"] >= 100      # north_star: >= 100x one CPU thread
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_basis_record_expansion_is_the_chain_rule():
    """capi.basis_to_core: the 80-byte basis record (H_b, g_b, cost w.r.t. theta = atan2(s, c), tx, ty) -> the ambient core record; checked on
    a synthetic residual r(c, s, tx, ty) = a . (theta, tx, ty): J_ambient = a F, so H = F^T (a a^T) F and g = F^T a r"""
    import numpy as np
    from randt_slam_b200 import capi
    rng = np.random.default_rng(0)
    S = 5
    poses = np.stack([np.array([np.cos(t), np.sin(t), x, y]) * [n, n, 1, 1] for t, x, y, n in zip(rng.uniform(-3, 3, S), rng.normal(0, 5, S), rng.normal(0, 5, S),
                                                                                                 [1.0, 1.01, 0.97, 1.0, 1.2])])
    a = rng.normal(0, 1, (S, 3)); r = rng.normal(0, 1, S)
    basis = np.zeros((S, capi.BASIS_STRIDE))
    iu = np.triu_indices(3)
    for s in range(S):
        basis[s, :6] = np.outer(a[s], a[s])[iu]; basis[s, 6:9] = a[s] * r[s]; basis[s, 9] = 0.5 * r[s] ** 2
    core = capi.basis_to_core(basis, poses)
    iu4 = np.triu_indices(4)
    for s in range(S):
        c, sn = poses[s, 0], poses[s, 1]
        n2 = c * c + sn * sn
        J = np.array([a[s, 0] * (-sn / n2), a[s, 0] * (c / n2), a[s, 1], a[s, 2]])      # d theta / d c = -s / n2, d theta / d s = c / n2
        assert np.allclose(core[s, :10], np.outer(J, J)[iu4], rtol=1e-13, atol=1e-15)
        assert np.allclose(core[s, 10:14], J * r[s], rtol=1e-13, atol=1e-15) and core[s, 14] == basis[s, 9]
    # finite differences of theta = atan2(s, c) confirm the chain-rule row
    h = 1e-7
    c, sn = poses[1, 0], poses[1, 1]
    assert abs((np.arctan2(sn, c + h) - np.arctan2(sn, c - h)) / (2 * h) - (-sn / (c * c + sn * sn))) < 1e-8
