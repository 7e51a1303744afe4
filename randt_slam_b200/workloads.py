"""Synthetic workloads of the five BASELINE.json configs, shared by bench.py, the GPU parity tests and the scripts.

  c0  configs[0]  one synthetic 2-D scan pair on the 32 x 32 cluster grid (~200 cells): plumbing
  c1  configs[1]  Oxford-shape scan vs 10-scan submap, batched (bench.py's headline; built in bench.py from pool_scans below)
  c2  configs[2]  ~2 k moving x ~8 k fixed cells, 0.5 m, k = 4, outdoor loss.  The reference is strictly SE(2): the
                  SE(3) + IMU wording has no reference implementation and is not run; the shape is run with the reference's
                  SE(2) + intensity functor, once with the reference's kNN pair list and once all-pairs
  c3  configs[3]  the literal batch: 256 independent keyframe-pair registrations, sharded over the ranks
  c4  configs[4]  sequence replay: one drive, scan by scan (voxelise -> associate -> register -> keyframe merge)

Everything here is host-side construction (numpy) plus calls through the ctypes binding of the C-ABI; no compute fallback.
"""
import math
import os
import time

import numpy as np

from . import params as P
from . import synth

POOL = 16                 # distinct synthetic scenes of the headline workload
SUBMAP_SCANS = 10


# ---------------------------------------------------------------------------------------------------------------
# c1: scenes of the headline workload
# ---------------------------------------------------------------------------------------------------------------
def pool_scans(p, seed0, pool=POOL, submap_scans=SUBMAP_SCANS):
    """`pool` scenes x (submap_scans keyframe scans + 1 moving scan), numpy only."""
    kw = synth.preset_scan_kwargs(p)
    sub, mov, sub_pose, true_pose = [], [], [], []
    for j in range(pool):
        sc = synth.scene_for(p, seed0 + j)
        rng = np.random.default_rng(seed0 * 1000 + j)
        for i in range(submap_scans):
            pose = (0.5 * i, 0.0, 0.0)
            sub.append(synth.make_scan(sc, pose, p, seed0 * 100 + j * 20 + i, **kw)); sub_pose.append(pose)
        tp = (2.5 + rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.05, 0.05))
        mov.append(synth.make_scan(sc, tp, p, seed0 * 100 + j * 20 + 19, **kw)); true_pose.append(tp)
    return sub, sub_pose, mov, true_pose


def _offsets(scans):
    return np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint32)


def build_submaps(ctx, capi, p, sub, sub_pose, n_maps, submap_scans=SUBMAP_SCANS):
    """n_maps submaps, each the mergeMapCell of its submap_scans transformed keyframe scans (LocalFuser keyframe insertion order);
    sub[j * submap_scans + i] is scan i of map j.  One batched K1 / transform / merge call per scan index."""
    gp = capi.grid_params(p)
    fixed = ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(n_maps + 1, np.uint32), gp)
    for i in range(submap_scans):
        scans = [sub[j * submap_scans + i] for j in range(n_maps)]
        m = ctx.voxelize(np.concatenate(scans), _offsets(scans), gp)
        poses = np.array([[math.cos(sub_pose[j * submap_scans + i][2]), math.sin(sub_pose[j * submap_scans + i][2]),
                           sub_pose[j * submap_scans + i][0], sub_pose[j * submap_scans + i][1]] for j in range(n_maps)], np.float32)
        m.transform(poses)
        fixed.merge(m)
        m.close()
    return fixed


# ---------------------------------------------------------------------------------------------------------------
# c0: the 32 x 32 plumbing case
# ---------------------------------------------------------------------------------------------------------------
def c0_case(seed=5):
    """one scan pair on the mixed preset (its cluster grid is exactly 32 x 32): (params, fixed points, moving points, initial guess[4])"""
    p = P.C1
    sc = synth.scene_for(p, seed)
    kw = synth.preset_scan_kwargs(p)
    fixed = synth.make_scan(sc, (0.0, 0.0, 0.0), p, seed * 10 + 1, **kw)
    moving = synth.make_scan(sc, (0.30, -0.20, 0.04), p, seed * 10 + 2, **kw)
    return p, fixed, moving, synth.pose_to_se2(0.25, -0.15, 0.03)


def run_c0(ctx, capi, oracle=None, reps=20):
    """c0 through the product path: K1 x 2 -> K2 -> K3 EMIT (r, J) -> one Gauss-Newton/LM iteration (max_num_iterations = 1)."""
    p, fpts, mpts, pose0 = c0_case()
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    F = ctx.voxelize(fpts, [0, len(fpts)], gp); M = ctx.voxelize(mpts, [0, len(mpts)], gp)
    prob = ctx.associate(F, M, pose0[None], k)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, p.ndt_weight / (M.info()[1] * k))
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor, gnc_max_steps=1,
                              max_num_iterations=1)
    r, J = prob.eval_emit(pose0)
    pose1, res = prob.register_batch(pose0[None], loss, opt)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        prob.eval_emit(pose0)
    t_emit = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        prob.register_batch(pose0[None], loss, opt)
    t_gn = (time.perf_counter() - t0) / reps
    out = {"workload": "configs[0]: one synthetic 2-D scan pair, 32x32 cluster grid, %d moving / %d fixed cells, %d pairs, k = %d" %
                       (M.info()[1], F.info()[1], prob.n_pairs, k),
           "pairs": int(prob.n_pairs), "emit_us_per_call": t_emit * 1e6, "one_gn_iteration_us_per_call": t_gn * 1e6,
           "pairs_per_s_emit_call": prob.n_pairs / t_emit, "what": "host-pointer C-ABI calls, wall clock incl. H2D pose and D2H results"}
    if oracle is not None:
        fd, md = F.download(), M.download()
        pm, pf, _ = prob.download()
        ro, Jo = oracle.eval_pairs(0, md["cells"], fd["cells"], pm, pf, pose0, 0)
        out["max_rel_err_r_vs_oracle"] = float(np.max(np.abs(r - ro) / np.maximum(ro, 1e-300)))
        out["max_rel_err_J_vs_oracle"] = float(np.max(np.abs(J - Jo)) / np.max(np.abs(Jo)))
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.eval_pairs(0, md["cells"], fd["cells"], pm, pf, pose0, 0)
        out["oracle_emit_us_per_call_1thread"] = (time.perf_counter() - t0) / reps * 1e6
    F.close(); M.close(); prob.close()
    return out


# ---------------------------------------------------------------------------------------------------------------
# c2: the 2 k x 8 k shape
# ---------------------------------------------------------------------------------------------------------------
def draw_cells(rng, n, half, res):
    """one cell per occupied slot: mean uniform in its slot, SPD covariance with eigenvalues logU(1e-3, 0.1), intensity variance U(20, 400)
    (the xy-intensity cross terms are drawn as correlations in (-0.4, 0.4) so that thin cells stay positive definite)"""
    side = int(2 * half / res)
    slots = rng.choice(side * side, size=n, replace=False)
    gx, gy = slots % side, slots // side
    mu = np.zeros((n, 3)); mu[:, 0] = -half + (gx + rng.uniform(0.05, 0.95, n)) * res; mu[:, 1] = -half + (gy + rng.uniform(0.05, 0.95, n)) * res
    mu[:, 2] = rng.uniform(70, 200, n)
    ang = rng.uniform(0, math.pi, n)
    l1 = np.exp(rng.uniform(math.log(1e-3), math.log(0.1), n)); l2 = np.exp(rng.uniform(math.log(1e-3), math.log(0.1), n))
    c, s = np.cos(ang), np.sin(ang)
    cov = np.zeros((n, 3, 3))
    cov[:, 0, 0] = c * c * l1 + s * s * l2; cov[:, 1, 1] = s * s * l1 + c * c * l2; cov[:, 0, 1] = cov[:, 1, 0] = c * s * (l1 - l2)
    cov[:, 2, 2] = rng.uniform(20, 400, n)
    x = rng.uniform(-0.4, 0.4, (n, 2)) * np.sqrt(np.minimum(l1, l2) * cov[:, 2, 2])[:, None]
    cov[:, 0, 2] = cov[:, 2, 0] = x[:, 0]; cov[:, 1, 2] = cov[:, 2, 1] = x[:, 1]
    out = np.zeros((n, 12), np.float32); out[:, :3] = mu; out[:, 3:] = cov.reshape(n, 9)
    return out[np.argsort(slots)]          # grid order, like a voxelised map


def c2_pool(pool=8, seed=3, n_fixed=8000, n_moving=2000, half=45.0):
    """pool of (fixed cells [n_fixed,12], moving cells [n_moving,12]); the moving cells are a perturbed subset of the fixed map seen from a
    displaced pose (theta 0.02, t (0.15, -0.1))"""
    p = P.C3
    rng = np.random.default_rng(seed)
    fixed = [draw_cells(rng, n_fixed, half, p.resolution) for _ in range(pool)]
    moving = []
    th, tx, ty = 0.02, 0.15, -0.1
    c, s = math.cos(th), math.sin(th)
    for j in range(pool):
        idx = np.sort(rng.choice(n_fixed, n_moving, replace=False))
        m = fixed[j][idx].copy()
        m[:, :2] += rng.normal(0, 0.05, (n_moving, 2)).astype(np.float32)
        x, y = m[:, 0] - tx, m[:, 1] - ty                # express in the moving frame: p_m = R^T (p_f - t)
        m[:, 0], m[:, 1] = c * x + s * y, -s * x + c * y
        moving.append(m)
    return fixed, moving, rng


def build_c2(ctx, capi, n_problems, pool=8, seed=3, n_fixed=8000, n_moving=2000):
    """-> (problem with the reference's kNN pair list (k = 4, Mahalanobis + intensity lookup), poses [S,4], host dict)"""
    p = P.C3
    gp = capi.grid_params(p)
    fixed, moving, rng = c2_pool(pool, seed, n_fixed, n_moving)
    sel = np.arange(n_problems) % pool
    f_off = np.arange(n_problems + 1, dtype=np.uint32) * n_fixed
    m_off = np.arange(n_problems + 1, dtype=np.uint32) * n_moving
    f_cells = np.concatenate([fixed[j] for j in sel]); m_cells = np.concatenate([moving[j] for j in sel])
    F = ctx.map_upload(f_cells, f_off, gp)
    M = ctx.map_upload(m_cells, m_off, gp)
    poses = np.stack([synth.pose_to_se2(0.15 + rng.uniform(-0.1, 0.1), -0.1 + rng.uniform(-0.1, 0.1), 0.02 + rng.uniform(-0.01, 0.01)) for _ in sel])
    prob = ctx.associate(F, M, poses, 4, capi.LOOKUP_MAHALANOBIS)
    pm, pf, seg = prob.download()
    host = dict(cells_m=m_cells, cells_f=f_cells, pm=pm, pf=pf, seg=seg, f_off=f_off, m_off=m_off, F=F, M=M)
    return prob, poses, host


def c2_loss(capi):
    return capi.make_loss(capi.LOSS_BARRON, P.OUTDOOR.loss_function_scale, P.OUTDOOR.loss_function_convexity, 1.0, 1.0)


# ---------------------------------------------------------------------------------------------------------------
# c3: the literal 256-registration batch
# ---------------------------------------------------------------------------------------------------------------
def literal_batch(p=P.OXFORD, n=256, seed=40, scenes=32, submap_scans=SUBMAP_SCANS):
    """256 independent keyframe-pair registrations (a loop-closure candidate sweep): registration i pairs its OWN moving scan (own pose,
    own noise seed) with the submap of scene i % scenes.  -> dict(sub, sub_pose (per scene), moving[n], truth[n], guess[n,4])"""
    kw = synth.preset_scan_kwargs(p)
    sub, sub_pose = [], []
    scene_objs = [synth.scene_for(p, seed + j) for j in range(scenes)]
    for j in range(scenes):
        for i in range(submap_scans):
            pose = (0.5 * i, 0.0, 0.0)
            sub.append(synth.make_scan(scene_objs[j], pose, p, seed * 100 + j * 20 + i, **kw)); sub_pose.append(pose)
    rng = np.random.default_rng(seed + 7)
    moving, truth, guess = [], [], []
    for i in range(n):
        tp = (2.5 + rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(-0.08, 0.08))
        moving.append(synth.make_scan(scene_objs[i % scenes], tp, p, seed * 1000 + i, **kw)); truth.append(tp)
        g = np.array(tp) + rng.uniform(-1, 1, 3) * [0.3, 0.3, 0.02]
        guess.append(synth.pose_to_se2(*g))
    return dict(p=p, sub=sub, sub_pose=sub_pose, moving=moving, truth=np.array(truth), guess=np.array(guess), scenes=scenes, n=n,
                submap_scans=submap_scans)


def build_literal_block(ctx, capi, batch, begin, end):
    """the problems [begin, end) of a literal batch as ONE device problem (what a rank builds for its shard).  The submaps are built per
    scene and gathered per problem, so that a problem's tables do not depend on which other problems share the block."""
    p = batch["p"]; gp = capi.grid_params(p)
    idx = np.arange(begin, end)
    scenes = np.unique(idx % batch["scenes"])
    ss = batch["submap_scans"]
    sub = [batch["sub"][j * ss + i] for j in scenes for i in range(ss)]
    sub_pose = [batch["sub_pose"][j * ss + i] for j in scenes for i in range(ss)]
    fixed = build_submaps(ctx, capi, p, sub, sub_pose, len(scenes), ss)
    fd = fixed.download(); fixed.close()
    where = {int(j): q for q, j in enumerate(scenes)}
    sel = [where[int(i % batch["scenes"])] for i in idx]
    f_cnt = np.diff(fd["cell_off"]).astype(np.int64)
    f_off = np.concatenate([[0], np.cumsum(f_cnt[sel])]).astype(np.uint32)
    f_cells = np.concatenate([fd["cells"][fd["cell_off"][q]:fd["cell_off"][q + 1]] for q in sel]) if len(sel) else np.zeros((0, 12), np.float32)
    f_npts = np.concatenate([fd["npts"][fd["cell_off"][q]:fd["cell_off"][q + 1]] for q in sel]) if len(sel) else np.zeros(0, np.uint32)
    F = ctx.map_upload(f_cells, f_off, gp, npts=f_npts, slot=fd["slot"][sel])
    mov = [batch["moving"][i] for i in idx]
    M = ctx.voxelize(np.concatenate(mov), _offsets(mov), gp)
    poses = batch["guess"][begin:end].copy()
    prob = ctx.associate(F, M, poses, p.n_results_nn_lookup, capi.LOOKUP_MAHALANOBIS)
    F.close(); M.close()
    return prob, poses


def literal_solver(capi, p):
    """estimateLoopConstraint semantics (ndt_matcher.cpp:426-493): loop-closure scale and GNC steps, weight 1, raw ambient pose block"""
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)
    opt = capi.solver_options(use_manifold=0, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.loop_closure_gnc_steps, max_num_iterations=p.max_iteration)
    return loss, opt


def solve_literal_block(ctx, capi, batch, begin, end):
    """-> rows [end - begin, shard.ROW]: [cos, sin, tx, ty, score, iterations, status, (index)]"""
    from . import shard
    rows = np.zeros((end - begin, shard.ROW))
    if end <= begin:
        return rows, 0.0
    prob, poses = build_literal_block(ctx, capi, batch, begin, end)
    loss, opt = literal_solver(capi, batch["p"])
    ctx.sync()
    t0 = time.perf_counter()
    out, res = prob.register_batch(poses, loss, opt)
    dt = time.perf_counter() - t0
    prob.close()
    rows[:, :4] = out; rows[:, 4] = res[:, capi.REG_SCORE]; rows[:, 5] = res[:, capi.REG_ITERATIONS]; rows[:, 6] = res[:, capi.REG_STATUS]
    return rows, dt


# ---------------------------------------------------------------------------------------------------------------
# c4: sequence replay
# ---------------------------------------------------------------------------------------------------------------
def _scan_job(args):
    name, scene_seed, pose, scan_seed = args
    p = P.PRESETS[name]
    return synth.make_scan(synth.scene_for(p, scene_seed), pose, p, scan_seed, **synth.preset_scan_kwargs(p))


# Scene of the replay drive.  With the Oxford preset's 3.5 m cells, odometry alone (no loop closure) loses half a cell in a single
# step on many random scenes — the CPU oracle chain, i.e. the reference's algorithm, does so on 5 of 9 scenes tried (seeds 300-308;
# seed 300: a 1.7 m step at scan 272 that it undoes at scan 456) — and past such a step a free-running chain is a chaotic function
# of the last bits of its sums.  On this scene the oracle chain stays within 0.71 m of the ground truth over all 8 609 scans, so
# "both chains track the same trajectory" is a statement about the implementations and not about the scene.
REPLAY_SCENE_SEED = 304


def make_loop_drive(p, seed, n_scans, radius=30.0, step=0.45, workers=None):
    """A drive of n_scans scans around a circle inside one synthetic scene (an 8 609-scan straight drive would leave any finite scene):
    0.45 m per scan like the Oxford vehicle at 4 Hz.  -> (truth [n,3], scans list).  Scans are generated by a process pool."""
    dth = step / radius
    truth = np.array([(radius * math.sin(i * dth), radius * (1.0 - math.cos(i * dth)), i * dth) for i in range(n_scans)])
    jobs = [(p.name, seed, tuple(truth[i]), seed * 100000 + i) for i in range(n_scans)]
    workers = workers or min(32, os.cpu_count() or 1)
    if workers > 1 and n_scans >= 64:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            scans = pool.map(_scan_job, jobs, chunksize=16)
    else:
        scans = [_scan_job(j) for j in jobs]
    return truth, scans


def odometry_solver(capi, p):
    return capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                               gnc_max_steps=p.gnc_steps, max_num_iterations=p.max_iteration)


def device_replay(ctx, capi, p, scans, keyframe_every=2, submap_keyframes=None):
    """One drive through the device path, scan by scan: K1 voxelise -> K2 associate against the submap at the previous pose -> K7 GNC + LM
    (manifold mode, odometry loss ScaledLoss(Barron(a, alpha, mu), ndt_weight / (n_cells k)), ndt_matcher.cpp:392) -> every
    `keyframe_every`-th scan transformMap + mergeMapCell into the submap (local_fuser.cpp:164-190).  -> (poses [n,4], seconds, mean LM iterations)"""
    gp = capi.grid_params(p)
    k = p.n_results_nn_lookup
    opt = odometry_solver(capi, p)
    n = len(scans)
    poses = np.zeros((n, 4)); poses[0] = synth.pose_to_se2(0, 0, 0)
    loss = capi.make_loss(capi.LOSS_BARRON, p.loss_function_scale, p.loss_function_convexity, 1.0, 1.0)
    sub = ctx.map_upload(np.zeros((0, 12), np.float32), np.zeros(2, np.uint32), gp)
    sub.scan_step(scans[0], gp, k, loss, p.ndt_weight, opt, True, poses[0])          # the first scan founds the submap
    iters = 0.0
    ctx.sync()
    t0 = time.perf_counter()
    for i in range(1, n):
        poses[i], res, _ = sub.scan_step(scans[i], gp, k, loss, p.ndt_weight, opt, i % keyframe_every == 0, poses[i - 1])     # one C-ABI call per scan
        iters += float(res[capi.REG_ITERATIONS])
    ctx.sync()
    dt = time.perf_counter() - t0
    sub.close()
    return poses, dt, iters / max(1, n - 1)


def oracle_replay(O, p, scans, keyframe_every=2):
    """the same chain on the CPU oracle (one host thread) -> (poses [n,4], seconds)"""
    va = (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)
    k = p.n_results_nn_lookup
    n = len(scans)
    poses = np.zeros((n, 4)); poses[0] = synth.pose_to_se2(0, 0, 0)
    t0 = time.perf_counter()
    v0 = O.voxelize(scans[0], *va)
    # the first scan founds the submap as a keyframe at the identity (transformMap + mergeMapCell into the empty map, like every later one)
    cells, npts, slot = O.merge_map_cell(np.zeros((0, 12), np.float32), np.zeros(0, np.uint32), np.full(p.size_x * p.size_y, -1, np.int32), p.size_x, p.size_y,
                                         p.resolution, O.transform_cells(v0["cells"], 1.0, 0.0, 0.0, 0.0), v0["npts"])
    for i in range(1, n):
        v = O.voxelize(scans[i], *va)
        w = p.ndt_weight / (len(v["cells"]) * k)
        o = O.loop_constraint(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, v["cells"], poses[i - 1], k,
                              matcher_loss_scale=p.loss_function_scale, loop_scale=p.loss_function_scale, alpha=p.loss_function_convexity,
                              divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.gnc_steps, on_manifold=True, loss_weight=w)
        poses[i] = o["pose"]
        if i % keyframe_every == 0:
            mc = O.transform_cells(v["cells"], *o["pose"].astype(np.float32))
            cells, npts, slot = O.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, v["npts"])
    return poses, time.perf_counter() - t0


# ---------------------------------------------------------------------------------------------------------------
# c4, reference-shaped: Matcher::estimateTransformCeres over the smoothing window (motion-model factors), scan by scan
# ---------------------------------------------------------------------------------------------------------------
def window_odometry_params(hostapi, p, covariance_scaling_factor=0.01, motion_sqrtI_diag=(1.0, 1.0, 10.0, 1.0, 3.0, 0.1, 20.0, 60.0)):
    """parameters_oxford.yaml:60-87 for the joint window problem (SE(2) manifold, constant velocity, no IMU)"""
    return hostapi.window_params(k=p.n_results_nn_lookup, gnc_steps=p.gnc_steps, max_iteration=p.max_iteration, loss_scale=p.loss_function_scale,
                                 alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, ndt_weight=p.ndt_weight,
                                 covariance_scaling_factor=covariance_scaling_factor, motion_sqrtI_diag=motion_sqrtI_diag)



def oracle_window_problem(oracle, p, fixed, fixed_se2, window, st, k):
    """The NDT blocks of one estimateTransformCeres call on the CPU oracle: every fixed scan voxelised and moved by its pose (one fixed map
    each), per free window state (st[1:], oldest first) and fixed map the association at the state's own pose (ndt_matcher.cpp:356-359);
    -> dict(cells_m, cells_f, im, jf, seg_off, n_cells) with one segment per state (tables of the parts back to back)"""
    va = (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)
    f_tabs = []
    for pts, T in zip(fixed, fixed_se2):
        v = oracle.voxelize(pts, *va)
        f_tabs.append((oracle.transform_cells(v["cells"], *np.asarray(T).astype(np.float32)), v["slot"]))   # transformMap leaves grid_indizes_ alone
    cells_m, cells_f, im_all, jf_all, seg_off, n_cells = [], [], [], [], [0], 0
    mb = fb = 0
    for j in range(1, len(st)):
        mv = oracle.voxelize(window[j - 1], *va)
        n_cells += len(mv["cells"])
        for cells, slot in f_tabs:
            im, jf = oracle.associate(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, mv["cells"], st[j, :4], k)
            im_all.append(im + mb); jf_all.append(jf + fb)
            cells_m.append(mv["cells"]); cells_f.append(cells)
            mb += len(mv["cells"]); fb += len(cells)
        seg_off.append(sum(len(a) for a in im_all))
    return dict(cells_m=np.concatenate(cells_m), cells_f=np.concatenate(cells_f), im=np.concatenate(im_all).astype(np.uint32),
                jf=np.concatenate(jf_all).astype(np.uint32), seg_off=np.array(seg_off, np.uint32), n_cells=n_cells)


def _se2_exp(u):
    th = u[2]
    if abs(th) < 1e-10:
        a, b = 1.0 - th * th / 6.0, 0.5 * th - th ** 3 / 24.0
    else:
        a, b = math.sin(th) / th, (1.0 - math.cos(th)) / th
    return np.array([math.cos(th), math.sin(th), a * u[0] - b * u[1], b * u[0] + a * u[1]])


def _se2_mul(A, B):
    re, im = A[0] * B[0] - A[1] * B[1], A[0] * B[1] + A[1] * B[0]
    n2 = re * re + im * im
    if n2 != 1.0:
        sc = 2.0 / (1.0 + n2); re *= sc; im *= sc
    return np.array([re, im, A[2] + (A[0] * B[2] - A[1] * B[3]), A[3] + (A[1] * B[2] + A[0] * B[3])])


def oracle_window_replay(oracle, p, scans, stamps, q, smoothing_steps=3, insertion_step=2):
    """the same loop on the CPU oracle (LocalFuser::processScan, local_fuser.cpp:99-300, one submap)"""
    va = (p.n_clusters, p.max_range, p.min_points_per_cell, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance)
    k = p.n_results_nn_lookup
    insertion_delay = smoothing_steps + 1
    t_start = time.perf_counter()
    traj, window, to_insert, poses = [], [], [], []
    cells = npts = slot = None
    cur = np.array([1.0, 0.0, 0.0, 0.0])
    qo = q.copy(); qo[14] = 0
    for i, pts in enumerate(scans):
        v = oracle.voxelize(pts, *va)
        if cells is not None:
            last = traj[-1].copy()
            dt = max(stamps[i] - last[13], 0.2)
            nxt = last.copy()
            nxt[:4] = _se2_mul(last[:4], _se2_exp([last[7] * dt, last[8] * dt, last[9] * dt]))    # predictSE2 with zero acceleration
            nxt[4:6] = nxt[2:4]; nxt[6] = math.atan2(nxt[1], nxt[0]); nxt[10:12] = 0.0; nxt[12] = 0.0; nxt[13] = stamps[i]
            traj.append(nxt)
            window.append(v)
            W = min(len(traj) - 1, smoothing_steps)
            st = np.array(traj[-W - 1:])
            cm, cf, im_all, jf_all, seg_off, n_cells, mb, fb = [], [], [], [], [0], 0, 0, 0
            for j in range(1, W + 1):
                mv = window[len(window) - W + (j - 1)]
                n_cells += len(mv["cells"])
                im, jf = oracle.associate(cells, slot, p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, mv["cells"], st[j, :4], k)
                im_all.append(im + mb); jf_all.append(jf + fb); cm.append(mv["cells"]); cf.append(cells)
                mb += len(mv["cells"]); fb += len(cells)
                seg_off.append(sum(len(a) for a in im_all))
            s_out, t_out, info = oracle.window_solve(st, qo, cur, np.concatenate(cm), np.concatenate(cf), np.concatenate(im_all).astype(np.uint32),
                                                     np.concatenate(jf_all).astype(np.uint32), np.array(seg_off, np.uint32), n_cells)
            if info["status"] != 0:
                raise RuntimeError("oracle window solve: nothing to solve at scan %d" % i)
            for j in range(W + 1):
                traj[len(traj) - W - 1 + j] = s_out[j]
            cur = t_out
            for b in range(1, min(smoothing_steps, len(traj)) + 1):      # both representations of the window states
                X = traj[-b]; X[4:6] = X[2:4]; X[6] = math.atan2(X[1], X[0])
            n = len(traj)
            if len(window) >= smoothing_steps:
                window.pop(0)
            if n % insertion_step == 0:
                to_insert.append(v)
            if n >= insertion_delay + insertion_step and (n - insertion_delay) % insertion_step == 0 and to_insert:
                Xs = traj[n - insertion_delay - 1]
                kf = to_insert.pop(0)
                mc = oracle.transform_cells(kf["cells"], *Xs[:4].astype(np.float32))
                cells, npts, slot = oracle.merge_map_cell(cells, npts, slot, p.size_x, p.size_y, p.resolution, mc, kf["npts"])
        else:
            s0 = np.zeros(14); s0[:4] = cur; s0[4:6] = cur[2:4]; s0[6] = math.atan2(cur[1], cur[0]); s0[13] = stamps[i]
            traj.append(s0)
            cells, npts, slot = oracle.transform_cells(v["cells"], *cur.astype(np.float32)), v["npts"], v["slot"]
        poses.append(cur.copy())
    return np.array(poses), np.array(traj), len(cells), time.perf_counter() - t_start


