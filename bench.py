#!/usr/bin/env python
"""bench.py — NDT cell-pair residual+Jacobian evaluation throughput on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (K3, FUSED mode: every cell pair's residual, its SE(2) Jacobian, the Barron-loss
correction and the per-pose J^T J / J^T r reduction) over one batch of synthetic registration problems.

Workload (config.workload): BASELINE configs[1]/[3] — Oxford-shape radar scans (400 beams, ~5k filtered points) registered
against a 10-scan submap with the shipped oxford parameters; `problems_per_gpu` independent (scan, submap) problems are
evaluated per step (the loop-closure candidate batch of configs[3]), sized so the resident tables exceed the 126 MB L2 and
every step streams them from HBM.  The problems are built by the product path itself: synthetic points -> K1 voxelise ->
map transform/merge -> K2 associate; a pool of distinct scenes is replicated (with distinct poses) to reach the batch size.

  value : pairs evaluated / s, tables + poses resident in HBM, CUDA-event timed on the context's stream
  e2e   : same metric through the host-pointer C-ABI call randt_eval_fused(): pinned-host poses H2D and the per-pose
          normal equations D2H inside the timed region, every step (the cell tables are construction-time state of the
          cost function, exactly as the reference's functors hold copies of their cells: ceres_residuals.h:528-535)
  --impl reference : the CPU restatement of the reference's path (oracle/, Jet<4> autodiff like ceres) on all host threads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from randt_slam_b200 import params as P  # noqa: E402
from randt_slam_b200 import synth  # noqa: E402
from randt_slam_b200 import workloads as W  # noqa: E402
from randt_slam_b200.workloads import POOL, SUBMAP_SCANS, pool_scans  # noqa: E402

METRIC = "ndt_cell_pair_residual_jacobian_evals_per_s"
UNIT = "pairs/s"
DEFAULT_PROBLEMS = 16384  # per GPU; ~22 KB of tables each -> ~360 MB resident, > 126 MB L2


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def build_problem(ctx, capi, p, n_problems, seed0):
    """Builds the batched problem with the product kernels; returns (problem, poses[S,4], stats)."""
    gp = capi.grid_params(p)
    sub, sub_pose, mov, true_pose = pool_scans(p, seed0)
    # submaps: merge the transformed keyframe maps scan by scan (LocalFuser keyframe insertion order)
    fixed = ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(POOL + 1, np.uint32), gp)
    for i in range(SUBMAP_SCANS):
        scans = [sub[j * SUBMAP_SCANS + i] for j in range(POOL)]
        off = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint32)
        m = ctx.voxelize(np.concatenate(scans), off, gp)
        pose = sub_pose[i]
        m.transform(np.tile(np.array([math.cos(pose[2]), math.sin(pose[2]), pose[0], pose[1]], np.float32), (POOL, 1)))
        fixed.merge(m)
        m.close()
    off = np.concatenate([[0], np.cumsum([len(s) for s in mov])]).astype(np.uint32)
    moving = ctx.voxelize(np.concatenate(mov), off, gp)
    fd, md = fixed.download(), moving.download()
    fixed.close(); moving.close()
    # replicate the pool to n_problems distinct problems (distinct initial guesses), tables laid out problem after problem
    rng = np.random.default_rng(seed0 + 7)
    idx = np.arange(n_problems) % POOL
    f_cnt = np.diff(fd["cell_off"]).astype(np.int64); m_cnt = np.diff(md["cell_off"]).astype(np.int64)
    f_off = np.concatenate([[0], np.cumsum(f_cnt[idx])]).astype(np.uint32)
    m_off = np.concatenate([[0], np.cumsum(m_cnt[idx])]).astype(np.uint32)
    f_cells = np.concatenate([fd["cells"][fd["cell_off"][j]:fd["cell_off"][j + 1]] for j in idx])
    m_cells = np.concatenate([md["cells"][md["cell_off"][j]:md["cell_off"][j + 1]] for j in idx])
    f_npts = np.concatenate([fd["npts"][fd["cell_off"][j]:fd["cell_off"][j + 1]] for j in idx])
    slot = fd["slot"][idx]
    guess = np.array([true_pose[j] for j in idx]) + rng.uniform(-1, 1, (n_problems, 3)) * [0.3, 0.3, 0.02]
    poses = np.stack([synth.pose_to_se2(*g) for g in guess])
    F = ctx.map_upload(f_cells, f_off, gp, npts=f_npts, slot=slot)
    M = ctx.map_upload(m_cells, m_off, gp)
    prob = ctx.associate(F, M, poses, p.n_results_nn_lookup, capi.LOOKUP_MAHALANOBIS)
    F.close(); M.close()
    pm, pf, seg = prob.download()
    stats = dict(n_m=int(prob.n_m), n_f=int(prob.n_f), pairs=int(prob.n_pairs), segments=int(prob.n_segments),
                 n_f_referenced=int(len(np.unique(pf))))
    host = dict(cells_m=m_cells, cells_f=f_cells, pm=pm, pf=pf, seg=seg, f_off=f_off, m_off=m_off, slot=slot)
    return prob, poses, stats, host


def measured_traffic(pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of one K3 fused launch from the committed `ncu --set full` capture of this very
    workload (profiles/r*_k3_fused_*_ncu_summary.txt: the `final` capture of the latest round, else the last by name); None when no
    capture was taken on this batch size."""
    import glob
    import re
    best = None
    paths = glob.glob(os.path.join(ROOT, "profiles", "r*_k3_fused_*_ncu_summary.txt"))
    for path in sorted(paths, key=lambda q: (os.path.basename(q).split("_")[0], "_final_" in os.path.basename(q), os.path.basename(q))):
        txt = open(path).read()
        m_pairs = re.search(r"([\d ]+) pairs per launch", txt)
        rd = re.search(r"dram__bytes_read\.sum\s+Mbyte\s+([\d.]+)", txt); wr = re.search(r"dram__bytes_write\.sum\s+Mbyte\s+([\d.]+)", txt)
        if m_pairs and rd and wr and int(m_pairs.group(1).replace(" ", "")) == pairs:
            best = (float(rd.group(1)) + float(wr.group(1))) * 1e6
    return best


def algorithmic_bytes(st, fused=True):
    """48 (N_m + N_f) + 8 P + 32 S + OUT, with N_f = fixed cells the pair list references; OUT = 192 S (fused)."""
    return 48 * (st["n_m"] + st["n_f_referenced"]) + 8 * st["pairs"] + 32 * st["segments"] + 192 * st["segments"]


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill(); out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_fused(host, poses, loss, n_problems, threads, repeats):
    """oracle leg: the same fused evaluation on the CPU over the first n_problems segments"""
    from oracle import oracle_py as O
    seg = host["seg"][: n_problems + 1]
    P_ = int(seg[-1])
    out, t = O.fused_batch(0, host["cells_m"], host["cells_f"], host["pm"][:P_], host["pf"][:P_], seg, poses[:n_problems],
                           (loss.kind, loss.scale, loss.alpha, loss.mu, loss.weight), want_jac=True, n_threads=threads, repeats=repeats)
    return P_ * repeats / t, t, out


def cpu_registrations(p, host, poses, n_sample, threads):
    """oracle leg: Matcher::estimateLoopConstraint restated (GNC + ceres-LM, Jet<4> evaluation) on the first n_sample problems.
    -> (registrations/s on one thread, registrations/s on `threads` threads, mean minimiser iterations)"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_py as O
    seg = host["seg"]

    def solve(s):
        a, b = int(seg[s]), int(seg[s + 1])
        pm = host["pm"][a:b]; pf = host["pf"][a:b]
        m0, f0 = int(pm.min()), int(pf.min())
        cm = host["cells_m"][m0:int(pm.max()) + 1]; cf = host["cells_f"][f0:int(pf.max()) + 1]
        dummy_slot = np.full(p.size_x * p.size_y, -1, np.int32)     # pairs are given: the slot table is not consulted
        r = O.loop_constraint(cf, dummy_slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, cm, poses[s], p.n_results_nn_lookup,
                              matcher_loss_scale=p.loss_function_scale, loop_scale=p.loop_closure_scale, alpha=p.loss_function_convexity,
                              divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.loop_closure_gnc_steps, max_iterations=p.max_iteration,
                              on_manifold=False, pairs=(pm - m0, pf - f0))
        return r["iterations"]
    t0 = time.perf_counter()
    its = [solve(s) for s in range(n_sample)]
    t1 = time.perf_counter() - t0
    with ThreadPoolExecutor(max_workers=threads) as ex:       # the ctypes call releases the GIL
        t0 = time.perf_counter()
        reps = max(1, (threads * 2) // n_sample + 1)
        list(ex.map(solve, [s % n_sample for s in range(n_sample * reps)]))
        tN = time.perf_counter() - t0
    return n_sample / t1, n_sample * reps / tN, float(np.mean(its))


def run_reference(args, p, loss):
    """--impl reference: CPU restatement of the reference path (the reference itself is not buildable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    from randt_slam_b200 import capi
    # the workload is built with the product kernels when a GPU is present (identical inputs to the product arm);
    # otherwise (CPU-only container) with the oracle's own voxeliser.  Either way only the oracle is timed.
    n_sample = min(args.ref_problems, args.problems)
    host, poses = build_host_workload(p, args.problems, args.seed)      # the GPU arm's problem set (same seeds, same construction rules)
    threads = O.hw_threads()
    P_ = int(host["seg"][n_sample]); P_all = int(host["seg"][args.problems])
    for _ in range(args.warmup):
        cpu_fused(host, poses, loss, n_sample, threads, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_fused(host, poses, loss, n_sample, threads, 1)
    dt = time.perf_counter() - t0
    value = P_ * args.steps / dt
    reg1, regN, reg_it = cpu_registrations(p, host, poses, min(args.reg_cpu_sample, n_sample), threads) if args.reg_steps > 0 else (None, None, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(p), "problems_per_gpu": args.problems, "pairs_per_gpu": P_all, "k": p.n_results_nn_lookup,
                   "mode": "fused (r, J, Barron corrector, per-pose J^T J / J^T r)", "preset": p.name,
                   "step": "bounded sample: the first %d problems (%d pairs) of the workload per step" % (n_sample, P_)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "first %d of %d problems (%d pairs) per step, Jet<4> autodiff + Barron corrector + J^T J accumulation" % (n_sample, args.problems, P_)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "registrations": {"value": regN, "unit": "registrations/s", "single_thread_value": reg1, "cores": threads, "mean_iterations": reg_it,
                          "what": "Matcher::estimateLoopConstraint restated (GNC + ceres-LM, Jet<4> evaluation), oxford loop-closure parameters"},
    }
    emit(line)


def workload_name(p):
    shape = "Oxford-shape scan (400 beams, ~5k pts)" if p.max_range >= 50.0 else "%s-shape scan (400 beams)" % p.name
    return "configs[1]/[3]: %s vs 10-scan submap, %s params, SE(2)+intensity, batched independent problems" % (shape, p.name)


def build_host_workload(p, n_problems, seed0):
    """CPU-only construction of the same workload shape (used by the reference arm, which must not need a GPU)."""
    from oracle import oracle_py as O
    sub, sub_pose, mov, true_pose = pool_scans(p, seed0)
    va = (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)
    rng = np.random.default_rng(seed0 + 7)
    fixed, moving = [], []
    for j in range(POOL):
        cells = np.zeros((0, 12), np.float32); npts = np.zeros(0, np.uint32); slot = np.full(p.size_x * p.size_y, -1, np.int32)
        for i in range(SUBMAP_SCANS):
            v = O.voxelize(sub[j * SUBMAP_SCANS + i], *va)
            pose = sub_pose[i]
            mc = O.transform_cells(v["cells"], math.cos(pose[2]), math.sin(pose[2]), pose[0], pose[1])
            cells, npts, slot = O.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, v["npts"])
        fixed.append(dict(cells=cells, slot=slot))
        moving.append(O.voxelize(mov[j], *va)["cells"])
    idx = np.arange(n_problems) % POOL
    guess = np.array([true_pose[j] for j in idx]) + rng.uniform(-1, 1, (n_problems, 3)) * [0.3, 0.3, 0.02]
    poses = np.stack([synth.pose_to_se2(*g) for g in guess])
    cm, cf, pm, pf, seg = [], [], [], [], [0]
    om = of = 0
    for s, j in enumerate(idx):
        im, jf = O.associate(fixed[j]["cells"], fixed[j]["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, moving[j],
                             poses[s], p.n_results_nn_lookup)
        cm.append(moving[j]); cf.append(fixed[j]["cells"]); pm.append(im + om); pf.append(jf + of)
        om += len(moving[j]); of += len(fixed[j]["cells"]); seg.append(seg[-1] + len(im))
    host = dict(cells_m=np.concatenate(cm), cells_f=np.concatenate(cf), pm=np.concatenate(pm).astype(np.uint32),
                pf=np.concatenate(pf).astype(np.uint32), seg=np.array(seg, np.uint32))
    return host, poses


def run_configs(args, ctx, capi, stream, rank, world, local, barrier):
    """Sub-records for the BASELINE configs that are not the headline workload: c0 (plumbing case), c2 (2 k x 8 k shape), c3 (the literal
    256-registration batch, sharded over the ranks: strong scaling) and c4 (sequence replay, rank 0; a single drive does not shard)."""
    import hashlib
    import torch
    import torch.distributed as dist
    from randt_slam_b200 import shard
    dev = "cuda:%d" % local
    out = {}
    pk, _ = peaks()
    oracle = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle_py as oracle
    # ---- c0 ----
    if rank == 0:
        out["c0"] = W.run_c0(ctx, capi, oracle)
    # ---- c1 (indoor reading): BASELINE.json words configs[1] as "indoor params"; the headline follows SURVEY 8d (oxford parameters) ----
    if args.c1_indoor_problems > 0 and args.preset != "indoor":
        pi = P.PRESETS["indoor"]
        prob_i, poses_i, st_i, _ = build_problem(ctx, capi, pi, args.c1_indoor_problems, args.seed)
        loss_i = capi.make_loss(capi.LOSS_BARRON, pi.loop_closure_scale, pi.loss_function_convexity, 1.0, 1.0)
        d_pi = torch.from_numpy(poses_i).to(dev); d_oi = torch.zeros((st_i["segments"], capi.FUSED_STRIDE), dtype=torch.float64, device=dev)
        for _ in range(5):
            prob_i.eval_fused_dev(d_pi.data_ptr(), d_oi.data_ptr(), loss_i)
        barrier()
        ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ei0.record(stream)
        for _ in range(50):
            prob_i.eval_fused_dev(d_pi.data_ptr(), d_oi.data_ptr(), loss_i)
        ei1.record(stream)
        barrier()
        ms_i = ei0.elapsed_time(ei1) / 50
        alg_i = algorithmic_bytes(st_i)
        out["c1_indoor"] = {"workload": workload_name(pi), "problems_per_gpu": st_i["segments"], "pairs_per_gpu": st_i["pairs"], "k": pi.n_results_nn_lookup,
                            "ms_per_step": ms_i, "pairs_per_s_per_gpu": st_i["pairs"] / (ms_i * 1e-3),
                            "roofline": {"bound": "hbm", "achieved": alg_i / (ms_i * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                         "frac": alg_i / (ms_i * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes_per_launch": alg_i, "bytes_per_pair": alg_i / st_i["pairs"]},
                            "note": "same kernel, k = 4: every moving cell is shared by four pairs, so the algorithmic bytes per pair are lower; the pair rate is what the kernel is bound by"}
        prob_i.close()
        del d_pi, d_oi
    # ---- c2: configs[2] shape with the reference's SE(2) + intensity functor and kNN pair list ----
    if args.c2_problems > 0:
        prob, poses, host = W.build_c2(ctx, capi, args.c2_problems)
        S, Pn = prob.n_segments, prob.n_pairs
        loss = W.c2_loss(capi)
        d_poses = torch.from_numpy(poses).to(dev); d_out = torch.zeros((S, capi.FUSED_STRIDE), dtype=torch.float64, device=dev)
        for _ in range(5):
            prob.eval_fused_dev(d_poses.data_ptr(), d_out.data_ptr(), loss)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 50
        e0.record(stream)
        for _ in range(n_rep):
            prob.eval_fused_dev(d_poses.data_ptr(), d_out.data_ptr(), loss)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / n_rep
        stc = dict(n_m=int(prob.n_m), n_f_referenced=int(len(np.unique(host["pf"]))), pairs=int(Pn), segments=int(S))
        alg = algorithmic_bytes(stc)
        rec = {"workload": "configs[2] shape: %d problems of 2 k moving x 8 k fixed cells, 0.5 m, k = 4, outdoor loss (alpha = -1), reference SE(2) + intensity "
                           "functor, kNN pair list (reference semantics)" % S,
               "se3_imu": "not run: the reference is strictly SE(2) (SURVEY fact 2); there is no reference implementation or oracle for an SE(3)/IMU variant",
               "pairs_per_gpu": int(Pn), "pairs_per_problem": Pn / S, "ms_per_step": ms, "pairs_per_s_per_gpu": Pn / (ms * 1e-3),
               "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": alg / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes_per_launch": alg, "bytes_per_pair": alg / Pn},
               "finite": bool(torch.isfinite(d_out).all().item())}
        if oracle is not None:      # parity spot check: first problem against the oracle
            a, b = int(host["seg"][0]), int(host["seg"][1])
            fo = oracle.fused(0, host["cells_m"], host["cells_f"], host["pm"][a:b], host["pf"][a:b], poses[0], (loss.kind, loss.scale, loss.alpha, loss.mu, loss.weight), True)
            g = capi.unpack_fused(d_out[0].cpu().numpy())
            rec["gpu_vs_oracle_max_rel_err_first_problem"] = float(max(np.max(np.abs(g["H"] - fo["H"])) / np.max(np.abs(fo["H"])), abs(g["cost"] - fo["cost"]) / abs(fo["cost"])))
        # all-pairs variant (K8): every moving cell against every fixed cell of its map pair, N_m x N_f = 16 M pairs per problem; inputs are
        # cache resident, so this one is bound by the fp64 pipe: reported in pairs/s and fp64 flop/s (~200 flop per pair closed form)
        n_ap = min(S, args.c2_allpairs_problems)
        if n_ap > 0:
            gp = capi.grid_params(P.C3)
            nm_, nf_ = int(host["m_off"][1]), int(host["f_off"][1])
            Fa = ctx.map_upload(host["cells_f"][:n_ap * nf_], host["f_off"][:n_ap + 1], gp); Ma = ctx.map_upload(host["cells_m"][:n_ap * nm_], host["m_off"][:n_ap + 1], gp)
            d_oa = torch.zeros((n_ap, capi.FUSED_STRIDE), dtype=torch.float64, device=dev)
            for _ in range(2):
                Fa.eval_allpairs_dev(Ma, d_poses.data_ptr(), d_oa.data_ptr(), loss)
            barrier()
            n_rep = 5
            e0.record(stream)
            for _ in range(n_rep):
                Fa.eval_allpairs_dev(Ma, d_poses.data_ptr(), d_oa.data_ptr(), loss)
            e1.record(stream)
            barrier()
            ms_ap = e0.elapsed_time(e1) / n_rep
            pairs_ap = float(n_ap) * nm_ * nf_
            FLOP_PER_PAIR = 2 * 62 + 36          # measured op mix of the per-pair closed form: ~62 DFMA (2 flop) + ~36 DMUL / DADD
            rec["all_pairs"] = {"workload": "%d problems x (%d moving x %d fixed cells) = %.3g pairs per step, no neighbour search, no window" % (n_ap, nm_, nf_, pairs_ap),
                                "bound": "fp64 pipe (inputs cache resident: %.2f MB of cells per problem)" % (48 * (nm_ + nf_) / 1e6),
                                "ms_per_step": ms_ap, "pairs_per_s_per_gpu": pairs_ap / (ms_ap * 1e-3), "fp64_flops_per_pair": FLOP_PER_PAIR,
                                "fp64_tflops": pairs_ap * FLOP_PER_PAIR / (ms_ap * 1e-3) / 1e12, "pairs_used_first_problem": float(d_oa[0, capi.FUSED_STRIDE - 1].item()),
                                "finite": bool(torch.isfinite(d_oa).all().item())}
            Fa.close(); Ma.close()
            del d_oa
        host["F"].close(); host["M"].close()
        prob.close()
        del d_poses, d_out
        out["c2"] = rec
    # ---- c3: the literal batch, strong scaling ----
    if args.c3_batch > 0:
        batch = W.literal_batch(P.OXFORD, n=args.c3_batch, seed=40, scenes=16)
        begin, end = shard.partition(batch["n"], world)[rank]
        W.solve_literal_block(ctx, capi, batch, begin, end)                 # warm-up (allocations, first launches)
        barrier()
        rows, dt = W.solve_literal_block(ctx, capi, batch, begin, end)      # dt: this rank's randt_register_batch call (host poses in, poses + records out)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        table = shard.gather_results(rows, begin, batch["n"], rank, world, device=torch.device("cuda", local))    # the job's only exchange (NCCL all_gather)
        if rank == 0:
            makespan = float(t[0])
            th = np.arctan2(table[:, 1], table[:, 0])
            ok = (np.hypot(table[:, 2] - batch["truth"][:, 0], table[:, 3] - batch["truth"][:, 1]) < 0.3) & (np.abs(th - batch["truth"][:, 2]) < 0.02)
            out["c3"] = {"workload": "configs[3]: %d independent keyframe-pair registrations (estimateLoopConstraint semantics, oxford loop-closure parameters), "
                                     "contiguous blocks over %d rank(s), one all_gather of the result rows" % (batch["n"], world),
                         "scaling": "strong", "registrations": batch["n"], "n_gpus": world, "makespan_ms": makespan * 1e3,
                         "registrations_per_s": batch["n"] / makespan, "mean_iterations": float(table[:, 5].mean()), "max_iterations": float(table[:, 5].max()),
                         "failed": int((table[:, 6] != 0).sum()), "within_0.3m_0.02rad_of_truth": int(ok.sum()),
                         "table_sha256": hashlib.sha256(np.ascontiguousarray(table[:, :7]).tobytes()).hexdigest(),
                         "what": "makespan = max over ranks of the wall time of one randt_register_batch call on the rank's block (host initial guesses in, poses + result "
                                 "records out); table_sha256 covers [cos, sin, tx, ty, score, iterations, status] of all registrations and must not depend on n_gpus"}
    # ---- c4: sequence replay (a single drive has a scan -> scan dependency: replicas only) ----
    if args.replay_scans > 1 and rank == 0:
        p4 = P.OXFORD
        t0 = time.perf_counter()
        truth, scans = W.make_loop_drive(p4, W.REPLAY_SCENE_SEED, args.replay_scans)
        t_gen = time.perf_counter() - t0
        W.device_replay(ctx, capi, p4, scans[:12])                          # warm-up
        l0 = ctx.launch_count
        poses_g, dt, its = W.device_replay(ctx, capi, p4, scans)
        launches = ctx.launch_count - l0
        est = np.stack([poses_g[:, 2], poses_g[:, 3], np.unwrap(np.arctan2(poses_g[:, 1], poses_g[:, 0]))], 1)
        err = np.hypot(est[:, 0] - truth[:, 0], est[:, 1] - truth[:, 1])
        rec = {"workload": "configs[4]: one synthetic drive of %d Oxford-shape scans (~5 k filtered points each) around a 30 m circle, per scan: K1 voxelise -> K2 "
                           "associate against the submap -> K7 GNC + LM registration -> every 2nd scan transform + merge into the submap" % len(scans),
               "scans": len(scans), "scans_per_s": (len(scans) - 1) / dt, "ms_per_scan": dt * 1e3 / (len(scans) - 1), "kernel_launches_per_scan": launches / (len(scans) - 1),
               "mean_lm_iterations_per_scan": its, "max_position_error_m": float(err.max()), "final_position_error_m": float(err[-1]),
               "sensor_rate_hz": 4.02, "scan_generation_s": t_gen, "n_gpus_used": 1,
               "note": "replicas only: a drive is a chain of dependent scans; NDT-only odometry without the reference's motion-model / IMU factors (host side)"}
        if oracle is not None and args.replay_oracle_scans > 1:
            n_o = min(args.replay_oracle_scans, len(scans))
            poses_o, dto = W.oracle_replay(oracle, p4, scans[:n_o])
            rec["cpu_baseline"] = {"scans_per_s": (n_o - 1) / dto, "cores": 1, "kind": "port", "sample": "first %d scans of the drive" % n_o,
                                   "max_abs_pose_difference_vs_device": float(np.max(np.abs(poses_o - poses_g[:n_o])))}
        if args.window_scans > 1:
            # the same drive through the reference-shaped odometry: Matcher::predictTransform -> Matcher::estimateTransformCeres over the
            # smoothing window (NDT blocks of 3 states in one K3 launch per LM evaluation + motion-model factors on the host) -> delayed
            # keyframe insertion at the smoothed pose (LocalFuser::processScan, local_fuser.cpp:99-300)
            from randt_slam_b200 import hostapi as HA
            n_w = min(args.window_scans, len(scans))
            q = W.window_odometry_params(HA, p4)
            stamps = 0.2486 * np.arange(n_w)                                 # Oxford scan period
            gp4 = capi.grid_params(p4)
            HA.window_replay(gp4, scans[:12], stamps[:12], q)                # warm-up
            poses_w, states_w, stats_w, tot = HA.window_replay(gp4, scans[:n_w], stamps, q)
            est_w = np.stack([states_w[:, 2], states_w[:, 3]], 1)
            err_w = np.hypot(est_w[:, 0] - truth[:n_w, 0], est_w[:, 1] - truth[:n_w, 1])
            wrec = {"workload": "first %d scans of the drive through randt::Matcher::estimateTransformCeres (window of 3 states, SE(2) manifold, constant-velocity "
                                "motion-model factors, parameters_oxford.yaml), keyframes inserted 4 scans late at their smoothed pose" % n_w,
                    "scans": n_w, "scans_per_s": (n_w - 1) / tot["seconds"], "ms_per_scan": tot["seconds"] * 1e3 / (n_w - 1),
                    "median_us_per_scan": float(np.median(stats_w[1:, 3])), "problem_setup_us_per_scan": tot["setup_seconds"] * 1e6 / (n_w - 1),
                    "gnc_lm_loop_us_per_scan": tot["solve_seconds"] * 1e6 / (n_w - 1), "mean_lm_iterations_per_scan": float(stats_w[1:, 0].mean()),
                    "mean_device_evaluations_per_scan": float(stats_w[1:, 1].mean()), "kernel_launches_per_scan": tot["launches"] / (n_w - 1),
                    "rejected_estimates": int(stats_w[:, 2].sum()), "keyframes": tot["keyframes"], "max_position_error_m": float(err_w.max()),
                    "final_speed_m_per_s": float(np.hypot(states_w[-1, 7], states_w[-1, 8])), "true_speed_m_per_s": 0.45 / 0.2486}
            if oracle is not None and args.window_oracle_scans > 1:
                n_o = min(args.window_oracle_scans, n_w)
                o_poses, o_states, _, dto = W.oracle_window_replay(oracle, p4, scans[:n_o], stamps[:n_o], q)
                wrec["cpu_baseline"] = {"scans_per_s": (n_o - 1) / dto, "cores": 1, "kind": "port", "sample": "first %d scans of the drive" % n_o,
                                        "max_abs_pose_difference_vs_device": float(np.max(np.abs(o_poses - poses_w[:n_o]))),
                                        "median_abs_pose_difference_vs_device": float(np.median(np.max(np.abs(o_poses - poses_w[:n_o]), axis=1))),
                                        "note": "both chains run free at ceres' default tolerances: single solves that creep along a weakly determined valley stop a step "
                                                "apart (DESIGN.md section 3, Window); with a fixed number of steps the iterates agree to 1e-7 (tests/test_window_gpu.py)"}
            rec["window_odometry"] = wrec
        out["c4"] = rec
    barrier()
    return out


def pin_rank_to_local_cores(local, world):
    """One process per GPU on one host: give every rank its own slice of the cores its GPU is attached to (nvidia-smi topo: CPU affinity
    of the GPU), so that the rank's caller thread, its copy-completion waits and torch's helper threads do not migrate across NUMA nodes
    or pile onto the cores of another rank.  -> description for the bench line (None when nothing was changed)."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    aff = {}
    for ln in out.splitlines():
        f = ln.split()
        if f and f[0].startswith("GPU") and f[0][3:].isdigit():
            spec = next((x for x in f[1:] if x[0].isdigit() and ("-" in x or "," in x) and not x.endswith("%")), None)
            if spec:
                cpus = []
                for part in spec.split(","):
                    a, _, b = part.partition("-")
                    cpus.extend(range(int(a), int(b or a) + 1))
                aff[int(f[0][3:])] = cpus
    if local not in aff:
        return None
    allowed = sorted(set(aff[local]) & set(os.sched_getaffinity(0)))
    peers = sorted(g for g in aff if g < world and aff[g] == aff[local])
    if not allowed or local not in peers:
        return None
    per = max(1, len(allowed) // len(peers))
    i = peers.index(local)
    mine = allowed[i * per:(i + 1) * per] or allowed
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return {"gpu_cpu_affinity": "%d-%d" % (aff[local][0], aff[local][-1]), "ranks_sharing_it": len(peers), "cores_of_this_rank": "%d-%d" % (mine[0], mine[-1])}


def d2h_bandwidth_probe(torch, dist, local, world, barrier, mb=64, reps=8):
    """every rank copies `mb` MB device -> pinned host at the same time; -> this rank's GB/s (the ranks' sum is what the host side sustains)"""
    d = torch.empty(mb * 1024 * 1024, dtype=torch.uint8, device="cuda:%d" % local)
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return mb * 1024 * 1024 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


_OUT = None


def emit(line):
    """the ONE JSON line, on the process's original stdout"""
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    # Libraries (NCCL's version banner, for one) write to file descriptor 1; keep the real stdout for the JSON line only.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--problems", type=int, default=DEFAULT_PROBLEMS, help="independent registration problems per GPU per step")
    ap.add_argument("--ref-problems", type=int, default=2048, help="problems each step of the CPU reference arm evaluates (bounded sample of the --problems workload)")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="problems in the cpu_baseline sample")
    ap.add_argument("--reg-steps", type=int, default=3, help="timed full-batch GNC+LM registration solves (0 disables the registrations leg)")
    ap.add_argument("--reg-streams", type=int, default=3, help="independent batches solved concurrently (own context/stream/thread each) in the pipelined leg; 0 disables")
    ap.add_argument("--reg-cpu-sample", type=int, default=48, help="registrations the oracle solves for the CPU comparison")
    ap.add_argument("--pre-scans", type=int, default=8, help="raw Oxford-size scans in the preprocessing leg (0 disables it)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-configs", action="store_true", help="skip the configs leg (c0, c2, c3, c4 sub-records)")
    ap.add_argument("--c1-indoor-problems", type=int, default=16384, help="problems of the indoor-parameter reading of configs[1] (c1_indoor sub-record; 0 disables)")
    ap.add_argument("--c2-problems", type=int, default=384, help="configs[2]-shaped problems (2 k x 8 k cells) evaluated per step in the c2 sub-record")
    ap.add_argument("--c2-allpairs-problems", type=int, default=16, help="problems of the c2 all-pairs variant (16 M pairs each)")
    ap.add_argument("--c3-batch", type=int, default=256, help="registrations of the literal configs[3] batch (sharded over the ranks)")
    ap.add_argument("--replay-scans", type=int, default=8609, help="scans of the configs[4] replay (Oxford sequence length); rank 0 only")
    ap.add_argument("--window-scans", type=int, default=1000, help="prefix of the drive replayed through estimateTransformCeres (window odometry); 0 skips it")
    ap.add_argument("--window-oracle-scans", type=int, default=120, help="prefix of that the CPU oracle's window chain replays for comparison")
    ap.add_argument("--replay-oracle-scans", type=int, default=300, help="prefix of the drive the CPU oracle chain replays for comparison")
    ap.add_argument("--preset", choices=sorted(P.PRESETS), default="oxford",
                    help="shipped parameter file the scans, maps and losses follow (SURVEY §8d C2 fixes oxford for the headline workload: an "
                         "Oxford-shape scan against parameters_oxford.yaml; BASELINE.json's configs[1] words it 'indoor params' — "
                         "`--preset indoor` runs that reading: 400-beam indoor-shape scans, 0.5 m cells, k = 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    p = P.PRESETS[args.preset]

    from randt_slam_b200 import capi
    # loss exactly as estimateLoopConstraint sets it (ndt_matcher.cpp:479): Barron(loop_closure_scale, alpha, mu), weight 1
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)

    if args.impl == "reference":
        run_reference(args, p, loss)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    pinned_to = pin_rank_to_local_cores(local, world)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream(device=local)
    ctx = capi.Context(local, stream=stream.cuda_stream)
    prob, poses, st, host = build_problem(ctx, capi, p, args.problems, args.seed)     # the same shard shape on every rank: equal pair counts
    S, Pn = st["segments"], st["pairs"]
    resident = 48 * (st["n_m"] + st["n_f"]) + 8 * Pn + 16 * prob.n_segments + 224 * S

    d_poses = torch.from_numpy(poses).to("cuda:%d" % local)
    d_out = torch.zeros((S, capi.FUSED_STRIDE), dtype=torch.float64, device="cuda:%d" % local)
    h_poses = torch.from_numpy(poses.copy()).pin_memory()
    h_out = torch.zeros((S, capi.FUSED_STRIDE), dtype=torch.float64).pin_memory()
    h_poses_np, h_out_np = h_poses.numpy(), h_out.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        prob.eval_fused_dev(d_poses.data_ptr(), d_out.data_ptr(), loss)

    def step_e2e():
        prob.eval_fused(h_poses_np, loss, out=h_out_np)

    # pipelined host-buffer calls (randt_eval_fused_async): every step uploads its own pinned poses and lands its own records in pinned
    # host memory; a buffer set is reused once the call that had it has delivered
    E2E_DEPTH = 4
    h_ring = [(capi.PinnedArray((S, 4)), capi.PinnedArray((S, capi.BASIS_STRIDE))) for _ in range(E2E_DEPTH)]
    for hp, _ in h_ring:
        hp.a[...] = poses
    from randt_slam_b200 import hostapi

    def run_e2e_pipelined(n):
        # the loop itself runs in C++ (librandt_host.so) over the public C-ABI: randt_eval_fused_async + randt_ctx_wait_async per step
        hostapi.eval_async_loop(ctx, prob, loss, [h[0].a for h in h_ring], [h[1].a for h in h_ring], n, packed=3)

    ev_ring = [torch.cuda.Event() for _ in range(E2E_DEPTH)]

    def run_e2e_pipelined_py(n):
        # the same loop driven from Python (ctypes call + torch events): kept as a cross-check of the C++ loop
        for i in range(n):
            j = i % E2E_DEPTH
            if i >= E2E_DEPTH:
                ev_ring[(i - 2) % E2E_DEPTH].synchronize()   # step i-2 has left the stream => the records of step i-4 are in host memory
            prob.eval_fused_async(h_ring[j][0].a, h_ring[j][1].a, loss, packed=3)
            ev_ring[j].record(stream)
        ctx.sync()

    with torch.cuda.stream(stream):
        # ---- value: device-resident ----
        for _ in range(args.warmup):
            step_dev()
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_dev()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count - l0
        # keep the same kernel running for >= 1.5 s so the 100 ms clock sampler sees it under load
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            for _ in range(50):
                step_dev()
            torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        # ---- e2e: host-pointer C-ABI, H2D poses + D2H normal equations every step ----
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_sync_s = time.perf_counter() - t0
        run_e2e_pipelined(args.warmup)
        barrier()
        t0 = time.perf_counter()
        run_e2e_pipelined(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        run_e2e_pipelined_py(args.warmup)
        barrier()
        t0 = time.perf_counter()
        run_e2e_pipelined_py(args.steps)
        barrier()
        e2e_py_s = time.perf_counter() - t0
        # the 80-byte basis records, expanded with the chain rule of the pose they were evaluated at, against the blocking call's full records
        core_ref = capi.pack_fused(h_out_np)[:, :capi.CORE_STRIDE]
        core_scale = np.max(np.abs(core_ref), axis=1, keepdims=True)
        e2e_max_err = max(float(np.max(np.abs(capi.basis_to_core(h_ring[j][1].a, poses) - core_ref) / core_scale)) for j in range(E2E_DEPTH))
        e2e_bits_equal = bool(e2e_max_err < 1e-13) and all(bool(np.array_equal(h_ring[j][1].a[:, 9], core_ref[:, 14])) for j in range(E2E_DEPTH))
        # ---- registrations: every problem of the batch solved to convergence (GNC + LM), K3 + K4, device resident ----
        reg = None
        if args.reg_steps > 0:
            # estimateLoopConstraint semantics (ndt_matcher.cpp:426-493): loop-closure scale and GNC steps, raw ambient pose block
            opt = capi.solver_options(use_manifold=0, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                                      gnc_max_steps=p.loop_closure_gnc_steps, max_num_iterations=p.max_iteration)
            d_reg = torch.empty_like(d_poses)
            d_res = torch.zeros((S, capi.REG_STRIDE), dtype=torch.float64, device="cuda:%d" % local)
            h_reg = torch.from_numpy(poses.copy()).pin_memory(); h_res = torch.zeros((S, capi.REG_STRIDE), dtype=torch.float64).pin_memory()

            def step_reg():
                d_reg.copy_(d_poses, non_blocking=True)
                prob.register_batch_dev(d_reg.data_ptr(), d_res.data_ptr(), loss, opt)

            step_reg()
            barrier()
            l1 = ctx.launch_count
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(args.reg_steps):
                step_reg()
            r1.record(stream)
            barrier()
            reg_ms = r0.elapsed_time(r1) / args.reg_steps
            reg_launches = (ctx.launch_count - l1) // args.reg_steps
            # end to end through the host-pointer call: initial guesses H2D, refined poses + result records D2H
            t0 = time.perf_counter()
            for _ in range(args.reg_steps):
                h_reg.numpy()[:] = poses
                lp_ = capi.C.byref(loss)
                ctx._check(capi.lib().randt_register_batch(ctx._h, prob._h, 0, capi._ptr(h_reg.numpy()), lp_, capi.C.byref(opt), capi._ptr(h_res.numpy())))
            barrier()
            reg_e2e_ms = (time.perf_counter() - t0) * 1e3 / args.reg_steps
            # Pipelined: a batch's makespan is set by its slowest registrations (~350 dependent K3 -> K4 iterations), during which the GPU is mostly
            # idle; independent batches on their own contexts (one stream and one host thread each) fill those gaps.
            pipe = None
            if args.reg_streams > 1:
                import threading
                workers = []
                for t_ in range(args.reg_streams):
                    c_ = capi.Context(local)
                    workers.append(dict(ctx=c_, prob=c_.problem_create(host["cells_m"], host["cells_f"], host["pm"], host["pf"], host["seg"]),
                                        poses=poses.copy(), res=np.zeros((S, capi.REG_STRIDE)), err=None))
                torch.cuda.synchronize()
                gate = threading.Barrier(args.reg_streams + 1)

                def run_(w_):
                    try:
                        for it_ in range(args.reg_steps + 1):
                            if it_ == 1:
                                gate.wait()                    # everybody has warmed up: the timed solves start together
                            w_["poses"][:] = poses
                            w_["ctx"]._check(capi.lib().randt_register_batch(w_["ctx"]._h, w_["prob"]._h, 0, capi._ptr(w_["poses"]), capi.C.byref(loss),
                                                                             capi.C.byref(opt), capi._ptr(w_["res"])))
                        gate.wait()
                    except Exception as e_:                     # never leave the other parties hanging at the barrier
                        w_["err"] = e_
                        gate.abort()
                ths = [threading.Thread(target=run_, args=(w_,)) for w_ in workers]
                for th_ in ths:
                    th_.start()
                try:
                    gate.wait(); t0 = time.perf_counter()
                    gate.wait(); dtp = time.perf_counter() - t0
                except threading.BrokenBarrierError:
                    dtp = None
                for th_ in ths:
                    th_.join()
                if dtp is not None:
                    same = all(np.array_equal(w_["res"][:, capi.REG_ITERATIONS], workers[0]["res"][:, capi.REG_ITERATIONS]) for w_ in workers)
                    pipe = {"value": args.reg_streams * args.reg_steps * S / dtp, "streams": args.reg_streams,
                            "ms_per_batch_effective": dtp * 1e3 / (args.reg_streams * args.reg_steps), "identical_results_across_streams": bool(same), "scope": "this rank's GPU only",
                            "what": "%d independent batches of %d registrations solved concurrently through randt_register_batch (host poses in, poses + "
                                    "records out), one context + stream + host thread each" % (args.reg_streams, S)}
                else:
                    pipe = {"error": str(next(w_["err"] for w_ in workers if w_["err"] is not None))}
                for w_ in workers:
                    w_["prob"].close(); w_["ctx"].close()
            res_np = d_res.cpu().numpy()
            reg = {"pipelined": pipe,"ms_per_batch": reg_ms, "e2e_ms_per_batch": reg_e2e_ms, "launches_per_batch": int(reg_launches),
                   "mean_iterations": float(res_np[:, capi.REG_ITERATIONS].mean()), "mean_gnc_solves": float(res_np[:, capi.REG_GNC_SOLVES].mean()),
                   "failed": int((res_np[:, capi.REG_STATUS] != 0).sum()), "rows": res_np, "poses": d_reg.cpu().numpy()}
        # ---- preprocessing: raw polar scan (400 x 3000 bins, resident) -> K6 peak filter -> K1 voxelise, per scan ----
        pre = None
        if args.pre_scans > 0:
            n_az, n_bins = 400, 3000
            fpar = capi.filter_params(p)
            raws = [torch.from_numpy(synth.make_raw_scan(synth.scene_for(p, args.seed + 50 + i), (0.3 * i, 0.0, 0.0), p, args.seed + 50 + i,
                                                         n_azimuth=n_az, n_bins=n_bins)).to("cuda:%d" % local) for i in range(args.pre_scans)]
            d_pts = torch.zeros((n_az * 64, 4), dtype=torch.float32, device="cuda:%d" % local)
            gp = capi.grid_params(p)

            def pre_pass():
                kept = cells = 0
                for r in raws:
                    n = ctx.filter_scan_dev(r.data_ptr(), n_az, n_bins, fpar, d_pts.data_ptr(), d_pts.shape[0])
                    m = ctx.voxelize(d_pts.data_ptr(), [0, n], gp, pts_on_device=True)
                    kept += n; cells += m.info()[1]
                    m.close()
                return kept, cells
            pre_pass()
            barrier()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                kept, cells = pre_pass()
            barrier()
            dt = (time.perf_counter() - t0) / (reps * len(raws))
            pre = {"scans_per_s": 1.0 / dt, "ms_per_scan": dt * 1e3, "raw_points_per_scan": n_az * n_bins, "raw_bytes_per_scan": 16 * n_az * n_bins,
                   "kept_points_per_scan": kept // len(raws), "cells_per_scan": cells // len(raws),
                   "what": "randt_filter_scan (K6, device in/out) + randt_voxelize (K1) per scan, wall clock incl. the two small D2H syncs; "
                           "the Oxford sensor delivers 4 scans/s"}
            # the same for a batch of scans per call (several sequences side by side / a backlog): randt_filter_scans -> randt_voxelize
            n_batch = 8 * len(raws)
            d_raw_b = torch.cat(raws * 8).contiguous()
            d_pts_b = torch.zeros((n_az * 64 * n_batch, 4), dtype=torch.float32, device="cuda:%d" % local)

            def pre_batch():
                off_b = ctx.filter_scans_dev(d_raw_b.data_ptr(), n_batch, n_az, n_bins, fpar, d_pts_b.data_ptr(), d_pts_b.shape[0])
                t1 = time.perf_counter()
                m = ctx.voxelize(d_pts_b.data_ptr(), off_b, gp, pts_on_device=True)
                nc = m.info()[1]
                m.close()
                return int(off_b[-1]), nc, t1
            pre_batch()
            barrier()
            t_f = t_all = 0.0
            for _ in range(reps):
                t0 = time.perf_counter()
                kept_b, cells_b, t1 = pre_batch()
                t2 = time.perf_counter()
                t_f += t1 - t0; t_all += t2 - t0
            t_f /= reps; t_all /= reps
            raw_bytes = 16.0 * n_az * n_bins * n_batch
            pk_, _ = peaks()
            pre["batched"] = {"scans_per_call": n_batch, "scans_per_s": n_batch / t_all, "ms_per_call": t_all * 1e3,
                              "filter_ms_per_call": t_f * 1e3, "filter_gbs": raw_bytes / t_f / 1e9, "filter_frac_of_hbm_peak": raw_bytes / t_f / 1e9 / pk_["hbm_gbs"],
                              "raw_bytes_per_call": int(raw_bytes), "kept_points_per_call": kept_b, "cells_per_call": int(cells_b),
                              "same_points_as_per_scan": bool(kept_b == 8 * kept),
                              "what": "randt_filter_scans (K6 over %d scans, device in/out) + randt_voxelize (K1) per call, wall clock incl. the host "
                                      "syncs; filter_gbs = raw bytes / wall time of the filter call" % n_batch}
            del d_raw_b, d_pts_b
        # ---- construction-time stages: K1 voxelise and K2 associate on resident batches ----
        stages = None
        if args.pre_scans > 0:
            sub_, _, mov_, _ = pool_scans(p, args.seed)
            base = sub_[:POOL] + mov_
            reps_ = 8                                        # 256 scans per call
            pts_np = np.concatenate(base * reps_)
            off_np = np.concatenate([[0], np.cumsum([len(s) for s in base * reps_])]).astype(np.uint32)
            d_pts_all = torch.from_numpy(pts_np).to("cuda:%d" % local)
            gp_ = capi.grid_params(p)
            ctx.voxelize(d_pts_all.data_ptr(), off_np, gp_, pts_on_device=True).close()
            barrier(); t0 = time.perf_counter()
            for _ in range(5):
                mvox = ctx.voxelize(d_pts_all.data_ptr(), off_np, gp_, pts_on_device=True)
                n_cells_vox = mvox.info()[1]; mvox.close()
            barrier(); t_vox = (time.perf_counter() - t0) / 5
            # K2 on the bench batch: re-associate the resident problem's maps
            Fm = ctx.map_upload(host["cells_f"], host["f_off"], gp_, slot=host["slot"])
            Mm = ctx.map_upload(host["cells_m"], host["m_off"], gp_)
            ctx.associate(Fm, Mm, poses, p.n_results_nn_lookup, capi.LOOKUP_MAHALANOBIS).close()
            barrier(); t0 = time.perf_counter()
            for _ in range(3):
                ctx.associate(Fm, Mm, poses, p.n_results_nn_lookup, capi.LOOKUP_MAHALANOBIS).close()
            barrier(); t_as = (time.perf_counter() - t0) / 3
            # K5: Cauchy-Schwarz divergence of every (submap, scan) pair of the batch (moving maps taken as already aligned)
            Fm.cs_divergence(Mm)
            barrier(); t0 = time.perf_counter()
            for _ in range(3):
                cs_all = Fm.cs_divergence(Mm)
            barrier(); t_cs = (time.perf_counter() - t0) / 3
            nf_ = np.diff(host["f_off"]).astype(np.float64); nm_ = np.diff(host["m_off"]).astype(np.float64)
            cs_terms = float(np.sum(nf_ * nm_ + nf_ * (nf_ - 1) / 2 + nm_ * (nm_ - 1) / 2))
            Fm.close(); Mm.close()
            pk_s, _ = peaks()
            vox_bytes = 16 * len(pts_np) + 52 * int(n_cells_vox)          # 16 N_pts + 48 N_cells + 4 N_cells
            # the same on 4096 scans per call (16 x the 256-scan batch)
            pts_big = d_pts_all.repeat(16, 1).contiguous()
            lens = np.diff(off_np.astype(np.int64))
            off_big = np.concatenate([[0], np.cumsum(np.tile(lens, 16))]).astype(np.uint32)
            ctx.voxelize(pts_big.data_ptr(), off_big, gp_, pts_on_device=True).close()
            barrier(); t0 = time.perf_counter()
            for _ in range(3):
                mbig = ctx.voxelize(pts_big.data_ptr(), off_big, gp_, pts_on_device=True)
                n_cells_big = mbig.info()[1]; mbig.close()
            barrier(); t_big = (time.perf_counter() - t0) / 3
            big_bytes = 16 * int(pts_big.shape[0]) + 52 * int(n_cells_big)
            del pts_big
            stages = {"voxelize": {"points_per_s": len(pts_np) / t_vox, "scans_per_s": (len(off_np) - 1) / t_vox, "ms_per_call": t_vox * 1e3, "scans_per_call": len(off_np) - 1,
                                   "points_per_call": int(len(pts_np)), "cells_per_call": int(n_cells_vox), "what": "randt_voxelize (K1), points resident, incl. per-call allocation and count readback",
                                   "roofline": {"bound": "hbm", "achieved": vox_bytes / t_vox / 1e9, "peak": pk_s["hbm_gbs"], "unit": "GB/s", "frac": vox_bytes / t_vox / 1e9 / pk_s["hbm_gbs"],
                                                "algorithmic_bytes_per_call": vox_bytes, "timing": "wall clock of the whole randt_voxelize call (allocation, K1, count readback, compaction)"},
                                   "scans_4096": {"ms_per_call": t_big * 1e3, "points_per_s": 16 * len(pts_np) / t_big, "cells_per_call": int(n_cells_big),
                                                  "roofline": {"bound": "hbm", "achieved": big_bytes / t_big / 1e9, "peak": pk_s["hbm_gbs"], "unit": "GB/s",
                                                               "frac": big_bytes / t_big / 1e9 / pk_s["hbm_gbs"], "algorithmic_bytes_per_call": big_bytes}}},
                      "cs_divergence": {"map_pairs_per_s": S / t_cs, "gaussian_overlaps_per_s": cs_terms / t_cs, "ms_per_call": t_cs * 1e3, "map_pairs_per_call": S,
                                        "finite": bool(np.isfinite(cs_all).all()), "what": "randt_cs_divergence (K5), maps resident: all-pairs 3x3 inverse + det + exp"},
                      "associate": {"queries_per_s": st["n_m"] / t_as, "ms_per_call": t_as * 1e3, "queries_per_call": st["n_m"],
                                    "what": "randt_associate (K2 + pair/duo compaction + record table + schedule), maps resident"}}
        configs = None if args.no_configs else run_configs(args, ctx, capi, stream, rank, world, local, barrier)
    bad = ctx.take_bad_pairs()

    t_ms = torch.tensor([ms, e2e_s * 1e3, reg["ms_per_batch"] if reg else 0.0, reg["e2e_ms_per_batch"] if reg else 0.0, e2e_sync_s * 1e3], dtype=torch.float64,
                        device="cuda:%d" % local)
    tot = torch.tensor([float(Pn), float(S)], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        # the only data exchange of the sharded job: gather every rank's per-problem result (here: cost) on all ranks
        if reg:
            from randt_slam_b200 import shard
            rows = np.zeros((S, shard.ROW)); rows[:, :4] = reg["poses"]; rows[:, 4] = reg["rows"][:, capi.REG_SCORE]
            rows[:, 5] = reg["rows"][:, capi.REG_ITERATIONS]; rows[:, 6] = reg["rows"][:, capi.REG_STATUS]
            table = shard.gather_results(rows, rank * S, world * S, rank, world, device=torch.device("cuda", local))   # NCCL all_gather
            assert table.shape[0] == world * S and not np.isnan(table[:, 4]).any()
    per_rank = None
    if world > 1:
        d2h_gbs = d2h_bandwidth_probe(torch, dist, local, world, barrier)
        mine = torch.tensor([ms, e2e_s * 1e3, reg["ms_per_batch"] if reg else 0.0, d2h_gbs], dtype=torch.float64, device="cuda:%d" % local)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank = {"ms_per_step": [float(g[0]) / args.steps for g in gathered], "e2e_ms_per_step": [float(g[1]) / args.steps for g in gathered],
                    "registrations_ms_per_batch": [float(g[2]) for g in gathered],
                    "concurrent_d2h_gbs": [float(g[3]) for g in gathered], "concurrent_d2h_gbs_sum": float(sum(float(g[3]) for g in gathered)),
                    "cpu_pinning_rank0": pinned_to}
    ms_all, e2e_ms_all, reg_ms_all, reg_e2e_ms_all, e2e_sync_ms_all = (float(t_ms[i]) for i in range(5))
    seg_all = float(tot[1])
    pairs_all = float(tot[0])

    if rank == 0:
        pk, pk_src = peaks()
        value = pairs_all * args.steps / (ms_all * 1e-3)
        e2e_value = pairs_all * args.steps / (e2e_ms_all * 1e-3)
        alg = algorithmic_bytes(st)
        kern_ms = ms / args.steps   # one launch per step, nothing else on the stream
        achieved = alg / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(p), "problems_per_gpu": S, "pairs_per_gpu": Pn, "moving_cells_per_gpu": st["n_m"],
                       "fixed_cells_per_gpu": st["n_f"], "fixed_cells_referenced": st["n_f_referenced"], "k": p.n_results_nn_lookup,
                       "mode": "fused (r, J, Barron corrector, per-pose J^T J / J^T r)", "resident_bytes_per_gpu": resident,
                       "l2_policy": "inputs larger than L2 (%.0f MB resident vs 126 MB), no flush" % (resident / 1e6),
                       "preset": p.name, "parallelism": "problems sharded across ranks, no data-path collective" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(S * 32), "d2h_bytes_per_step": int(S * 8 * capi.BASIS_STRIDE),
                    "ms_per_step": e2e_ms_all / args.steps, "in_flight": E2E_DEPTH,
                    "blocking_value": pairs_all * args.steps / (e2e_sync_ms_all * 1e-3), "blocking_ms_per_step": e2e_sync_ms_all / args.steps,
                    "results_equal_blocking_call": e2e_bits_equal, "max_rel_difference_vs_blocking_call": e2e_max_err, "python_loop_ms_per_step_rank0": e2e_py_s * 1e3 / args.steps,
                    "api": "randt_eval_fused_async from a C++ caller (host pointers, wall clock): every step uploads its own pinned poses (copy stream, two device "
                           "slots) and its per-pose records — the normal equations one LM iteration consumes, in the functor's own (theta, tx, ty) basis before the chain "
                           "rule to the ambient parameters: 3x3 upper triangle, gradient, cost = 80 B (packed == 3; the ambient 4x4 and the Sophus tangent system "
                           "follow from them and the pose on the host: capi.basis_to_core) — are copied out to "
                           "the caller's pinned result buffer (second copy stream) while the next step's kernel runs; four host buffer "
                           "sets, the host waits (randt_ctx_wait_async) for step i-4 before reusing its buffers for step i.  blocking_value: "
                           "the same through randt_eval_fused, one step at a time, K3 storing the full 192 B records straight into the "
                           "pinned result buffer.  results_equal_blocking_call: the basis records expanded to the ambient layout agree with the blocking call's "
                           "records to max_rel_difference_vs_blocking_call (< 1e-13) and the cost entries bit for bit"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": measured_traffic(Pn), "peak_source": pk_src, "kernel": "k3_fused_kernel<0,BARRON_M2,true>",
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": kern_ms},
            "clocks": clocks,
            "per_rank": per_rank,
            "degenerate_pairs": bad,
        }
        if configs is not None:
            configs["c1"] = {"workload": workload_name(p), "see": "the top-level value / e2e / roofline / registrations of this line"}
            line["configs"] = configs
        if pre:
            line["preprocess"] = pre
        if stages:
            if not args.no_cpu_baseline:
                from oracle import oracle_py as O
                va = (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)
                t0 = time.perf_counter()
                for sc_ in base[:8]:
                    O.voxelize(sc_, *va)
                stages["voxelize"]["cpu_points_per_s_1thread"] = sum(len(s_) for s_ in base[:8]) / (time.perf_counter() - t0)
                t0 = time.perf_counter(); nq = 0
                for s_ in range(8):
                    a_, b_ = int(host["f_off"][s_]), int(host["f_off"][s_ + 1]); c_, d_ = int(host["m_off"][s_]), int(host["m_off"][s_ + 1])
                    O.associate(host["cells_f"][a_:b_], host["slot"][s_], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance,
                                host["cells_m"][c_:d_], poses[s_], p.n_results_nn_lookup)
                    nq += d_ - c_
                stages["associate"]["cpu_queries_per_s_1thread"] = nq / (time.perf_counter() - t0)
                t0 = time.perf_counter(); nt = 0.0
                for s_ in range(4):
                    a_, b_ = int(host["f_off"][s_]), int(host["f_off"][s_ + 1]); c_, d_ = int(host["m_off"][s_]), int(host["m_off"][s_ + 1])
                    O.cs_divergence(host["cells_f"][a_:b_], host["cells_m"][c_:d_])
                    nt += (b_ - a_) * (d_ - c_) + (b_ - a_) * (b_ - a_ - 1) / 2 + (d_ - c_) * (d_ - c_ - 1) / 2
                stages["cs_divergence"]["cpu_gaussian_overlaps_per_s_1thread"] = nt / (time.perf_counter() - t0)
            line["stages"] = stages
        if reg:
            line["registrations"] = {
                "value": seg_all / (reg_ms_all * 1e-3), "unit": "registrations/s", "e2e_value": seg_all / (reg_e2e_ms_all * 1e-3),
                "ms_per_batch": reg_ms_all, "batch_per_gpu": S, "launches_per_batch": reg["launches_per_batch"],
                "mean_iterations": reg["mean_iterations"], "mean_gnc_solves": reg["mean_gnc_solves"], "failed": reg["failed"],
                "pipelined": reg["pipelined"],
                "what": "full GNC + Levenberg-Marquardt solve of every problem of the batch (randt_register_batch: K3 fused + K4 per LM iteration), "
                        "Matcher::estimateLoopConstraint semantics with the oxford loop-closure parameters"}
        if not args.no_cpu_baseline:
            from oracle import oracle_py as O
            n_s = min(args.cpu_sample, S)
            threads = O.hw_threads()
            v1, t1, out_cpu = cpu_fused(host, poses, loss, n_s, 1, 1)
            reps = max(1, int(10.0 / max(t1 / max(threads, 1), 1e-3)))
            reps = min(reps, 200)
            vN, tN, _ = cpu_fused(host, poses, loss, n_s, threads, reps)
            # parity spot check of the timed GPU result against the oracle on the sample
            g = d_out[:n_s].cpu().numpy()
            ref = np.concatenate([out_cpu["H"].reshape(n_s, 16), out_cpu["g"], out_cpu["cost"][:, None]], 1)
            got = np.concatenate([g[:, :16], g[:, 16:20], g[:, 20:21]], 1)
            err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
            if reg:
                c1, cN, cit = cpu_registrations(p, host, poses, min(args.reg_cpu_sample, S), threads)
                line["registrations"]["cpu_baseline"] = {"value": cN, "single_thread_value": c1, "cores": threads, "kind": "port", "mean_iterations": cit,
                                                         "sample": "first %d problems" % min(args.reg_cpu_sample, S)}
            # BASELINE.md B3 (informational): the closed form (what a hand-optimised CPU path would evaluate) on one thread,
            # residual + Jacobian per pair without loss or accumulation
            P_s = int(host["seg"][n_s])
            cf_reps = 0; t0 = time.perf_counter()
            while time.perf_counter() - t0 < 1.0:
                for s_ in range(min(n_s, 64)):
                    a_, b_ = int(host["seg"][s_]), int(host["seg"][s_ + 1])
                    O.eval_pairs(0, host["cells_m"], host["cells_f"], host["pm"][a_:b_], host["pf"][a_:b_], poses[s_], 1)
                cf_reps += 1
            cf_rate = cf_reps * int(host["seg"][min(n_s, 64)]) / (time.perf_counter() - t0)
            line["cpu_baseline"] = {"value": vN, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d problems (%d pairs) x %d passes, Jet<4> autodiff + corrector + J^T J" % (n_s, P_s, reps),
                                    "single_thread_value": v1, "closed_form_single_thread_value": cf_rate,
                                    "gpu_vs_oracle_max_rel_err_on_sample": err}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
