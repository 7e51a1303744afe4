"""ctypes view of librandt_host.so's test hooks (randt_hostapi_*): drives the C++ mirror of the reference's Map / Matcher /
ceres::CostFunction surface (include/randt_host.hpp) from the Python tests.  No compute here and no fallback."""
import ctypes as C
import os

import numpy as np

from . import capi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librandt_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise capi.RandtError(capi.E_INVALID, "host library %s is missing: run `python -m randt_slam_b200.build`" % LIB_PATH)
        capi.lib()   # librandt_gpu.so first (the host library links against it)
        L = C.CDLL(LIB_PATH)
        L.randt_hostapi_last_error.restype = C.c_char_p
        for name in ("randt_hostapi_loop_constraints", "randt_hostapi_cost_function", "randt_hostapi_bnb", "randt_hostapi_export", "randt_hostapi_odometry", "randt_hostapi_eval_async_loop", "randt_hostapi_build_schedule",
                     "randt_hostapi_window_factors", "randt_hostapi_window_solve", "randt_hostapi_window_replay",
                     "randt_hostapi_window_minimize_factors"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise capi.RandtError(rc, lib().randt_hostapi_last_error().decode())


def _pf(a):
    return a.ctypes.data_as(C.c_void_p)


def loop_constraints(gp, fixed_scans, moving_scans, poses, k, loss_function_scale, convexity, divisor, max_gnc_steps, loop_scale,
                     optimize_on_manifold=True, device=0):
    """Map::addClusters x2 -> Matcher::estimateLoopConstraint(s) for n (fixed scan, moving scan) pairs -> (poses [n,4], scores [n])"""
    n = len(fixed_scans)
    f = np.ascontiguousarray(np.concatenate(fixed_scans), np.float32); m = np.ascontiguousarray(np.concatenate(moving_scans), np.float32)
    fo = np.concatenate([[0], np.cumsum([len(s) for s in fixed_scans])]).astype(np.uint32)
    mo = np.concatenate([[0], np.cumsum([len(s) for s in moving_scans])]).astype(np.uint32)
    poses = np.ascontiguousarray(poses, np.float64).reshape(n, 4).copy()
    scores = np.zeros(n, np.float64)
    _check(lib().randt_hostapi_loop_constraints(C.c_int(device), C.byref(gp), _pf(f), _pf(fo), _pf(m), _pf(mo), C.c_uint32(n), C.c_int(k),
                                                C.c_double(loss_function_scale), C.c_double(convexity), C.c_double(divisor), C.c_int(max_gnc_steps),
                                                C.c_double(loop_scale), C.c_int(int(optimize_on_manifold)), _pf(poses), _pf(scores)))
    return poses, scores


def cost_function(gp, fixed_pts, moving_pts, k, guess, pose, loss=None, want_jac=True, device=0):
    """Matcher::addNDTFactor as one ceres::CostFunction, evaluated through CostFunction::Evaluate -> (residuals, J [n,4], max raw r)"""
    f = np.ascontiguousarray(fixed_pts, np.float32); m = np.ascontiguousarray(moving_pts, np.float32)
    cap = 4 * len(m) + 8
    res = np.zeros(cap, np.float64); J = np.zeros((cap, 4), np.float64) if want_jac else None
    n = C.c_uint32(0); mx = C.c_double(0)
    guess = np.ascontiguousarray(guess, np.float64); pose = np.ascontiguousarray(pose, np.float64)
    _check(lib().randt_hostapi_cost_function(C.c_int(device), C.byref(gp), _pf(f), C.c_uint32(len(f)), _pf(m), C.c_uint32(len(m)), C.c_int(k),
                                             _pf(guess), C.byref(loss) if loss is not None else None, _pf(pose), _pf(res),
                                             _pf(J) if want_jac else None, C.c_uint32(cap), C.byref(n), C.byref(mx)))
    return res[: n.value].copy(), (J[: n.value].copy() if want_jac else None), mx.value


def bnb(gp, fixed_pts, moving_pts, pose, convexity, scale, window_linear=4.5, window_angular=0.45, linear_step=0.4, max_px_range=4.0,
        cost_threshold=0.82, n_iter=2, device=0):
    f = np.ascontiguousarray(fixed_pts, np.float32); m = np.ascontiguousarray(moving_pts, np.float32)
    pose = np.ascontiguousarray(pose, np.float64).copy()
    mc = C.c_double(0); ne = C.c_uint32(0)
    _check(lib().randt_hostapi_bnb(C.c_int(device), C.byref(gp), _pf(f), C.c_uint32(len(f)), _pf(m), C.c_uint32(len(m)), C.c_double(convexity),
                                   C.c_double(scale), C.c_double(window_linear), C.c_double(window_angular), C.c_double(linear_step),
                                   C.c_double(max_px_range), C.c_double(cost_threshold), C.c_int(n_iter), _pf(pose), C.byref(mc), C.byref(ne)))
    return pose, mc.value, ne.value


def export_normal_distributions(gp, pts, device=0):
    """Map::addClusters + the ndt_msgs Mean/Covariance export -> (mean [n,3], cov [n,6]) float64"""
    pts = np.ascontiguousarray(pts, np.float32)
    cap = max(16, len(pts))
    mean = np.zeros((cap, 3)); cov = np.zeros((cap, 6)); n = C.c_uint32(0)
    _check(lib().randt_hostapi_export(C.c_int(device), C.byref(gp), _pf(pts), C.c_uint32(len(pts)), _pf(mean), _pf(cov), C.c_uint32(cap), C.byref(n)))
    return mean[: n.value].copy(), cov[: n.value].copy()


def odometry(gp, fixed_scans, fixed_poses, moving_pts, prior, k, loss_function_scale, convexity, divisor, gnc_steps, ndt_weight,
             optimize_on_manifold=True, reject_translation=5.0, reject_rotation=2.0, device=0):
    """Matcher::estimateTransformNDT: every fixed scan is voxelised and moved by its pose (a stand-in for a submap), the moving scan is
    registered against all of them at once -> (pose [4], accepted)"""
    fs = [np.ascontiguousarray(f, np.float32) for f in fixed_scans]
    ptrs = (C.c_void_p * len(fs))(*[f.ctypes.data for f in fs])
    n_pts = np.array([len(f) for f in fs], np.uint32)
    fp = np.ascontiguousarray(fixed_poses, np.float64).reshape(len(fs), 4)
    m = np.ascontiguousarray(moving_pts, np.float32)
    pose = np.ascontiguousarray(prior, np.float64).copy(); ok = C.c_int(0)
    _check(lib().randt_hostapi_odometry(C.c_int(device), C.byref(gp), ptrs, _pf(n_pts), _pf(fp), C.c_uint32(len(fs)), _pf(m), C.c_uint32(len(m)),
                                        C.c_int(k), C.c_double(loss_function_scale), C.c_double(convexity), C.c_double(divisor), C.c_int(gnc_steps),
                                        C.c_double(ndt_weight), C.c_int(int(optimize_on_manifold)), C.c_double(reject_translation),
                                        C.c_double(reject_rotation), _pf(pose), C.byref(ok)))
    return pose, bool(ok.value)


def eval_async_loop(ctx, prob, loss, poses_ring, out_ring, steps, packed=True, variant=0):
    """`steps` pipelined randt_eval_fused_async calls issued from C++ over caller-owned pinned buffer sets (capi.PinnedArray)"""
    d = len(poses_ring)
    pp = (C.c_void_p * d)(*[a.ctypes.data for a in poses_ring]); oo = (C.c_void_p * d)(*[a.ctypes.data for a in out_ring])
    _check(lib().randt_hostapi_eval_async_loop(ctx._h, prob._h, C.c_int(variant), C.byref(loss) if loss is not None else None, pp, oo, C.c_uint32(d),
                                               C.c_uint32(int(steps)), C.c_int(int(packed))))


def build_schedule(duo_off, max_warps):
    """The K3 schedule for per-segment duo offsets (no device needed) -> dict of numpy arrays"""
    duo_off = np.ascontiguousarray(duo_off, np.uint32)
    S = len(duo_off) - 1
    n_duos = int(duo_off[-1])
    cap_t = n_duos // 32 + S + 2; cap_c = n_duos // 32 + 2 * cap_t + 2
    counts = np.zeros(5, np.uint32)
    tiles = np.zeros((cap_t, 4), np.uint32); pa = np.zeros((cap_c, 4), np.uint32); pb = np.zeros((cap_c, 4), np.uint32)
    wa = np.zeros(max_warps + 1, np.uint32); wb = np.zeros(max_warps + 1, np.uint32)
    trb = np.zeros(cap_t + 1, np.uint32); tdb = np.zeros(cap_t, np.uint32); first = np.zeros(S + 1, np.uint32)
    _check(lib().randt_hostapi_build_schedule(_pf(duo_off), C.c_uint32(S), C.c_uint32(int(max_warps)), _pf(counts), _pf(tiles), C.c_uint32(cap_t), _pf(pa),
                                              _pf(pb), C.c_uint32(cap_c), _pf(wa), _pf(wb), _pf(trb), _pf(tdb), _pf(first)))
    nt, nw, na, nb, nr = (int(c) for c in counts)
    return dict(tiles=tiles[:nt], n_warps=nw, plan_a=pa[:na], plan_b=pb[:nb], woff_a=wa[: nw + 1], woff_b=wb[: nw + 1], tile_rec_begin=trb[: nt + 1],
                tile_duo_begin=tdb[:nt], first=first, n_records=nr)


def predict(state12, raw_dt, se2_model):
    """randt::predict / randt::predictSE2 on one state [cos, sin, tx, ty, pos_x, pos_y, rot, vx, vy, omega, ax, ay] -> predicted state"""
    s = np.ascontiguousarray(state12, np.float64).reshape(12)
    out = np.zeros(12, np.float64)
    lib().randt_hostapi_predict.restype = None
    lib().randt_hostapi_predict(C.c_int(int(se2_model)), _pf(s), C.c_double(float(raw_dt)), _pf(out))
    return out


def window_params(k=2, gnc_steps=2, max_iteration=200, loss_scale=1.0, alpha=-2.0, divisor=1.1, ndt_weight=5000.0, manifold=True,
                  constant_velocity=True, use_imu=False, weight_imu=64.0, weight_imu_bias=750000.1, reject_translation=5.0, reject_rotation=2.0,
                  use_intensity=True, covariance_scaling_factor=0.01, motion_sqrtI_diag=(1.0, 1.0, 10.0, 1.0, 3.0, 0.1, 20.0, 60.0)):
    """The 16 + 64 doubles randt_hostapi_window_solve / window_factors (and the oracle's orc_window_*) read; defaults = parameters_oxford.yaml.
    Slot 14 is use_intensity here and the functor variant (0 = SE2 + intensity ... 3 = vector xy) on the oracle side: see oracle_variant()."""
    q = np.zeros(80, np.float64)
    q[:15] = [k, gnc_steps, max_iteration, loss_scale, alpha, divisor, ndt_weight, float(manifold), float(constant_velocity), float(use_imu), weight_imu,
              weight_imu_bias, reject_translation, reject_rotation, float(use_intensity)]
    q[16:] = (covariance_scaling_factor * np.diag(np.asarray(motion_sqrtI_diag, np.float64))).reshape(64)
    return q


def window_factors(states14, params80, imu=None):
    """The motion-model (+ IMU) factors of the window over states [(W + 1), 14], no device -> (cost, g [nt], H [nt, nt])"""
    st = np.ascontiguousarray(states14, np.float64).reshape(-1, 14)
    W = len(st) - 1
    cap = 10 * (W + 1)
    g = np.zeros(cap); H = np.zeros(cap * cap); cost = C.c_double(0)
    im = None if imu is None else np.ascontiguousarray(imu, np.float64)
    nt = lib().randt_hostapi_window_factors(_pf(st), C.c_uint32(W), _pf(im) if im is not None else None, _pf(np.ascontiguousarray(params80, np.float64)),
                                            C.byref(cost), _pf(g), _pf(H))
    if nt < 0:
        _check(nt)
    return cost.value, g[:nt].copy(), H[: nt * nt].reshape(nt, nt).copy()


def window_solve(gp, fixed_scans, fixed_poses, window_scans, states14, params80, trans, imu=None, tolerances=None, device=0):
    """Matcher::estimateTransformCeres (joint window problem) over scans given as points -> (states [(W + 1), 14], trans [4], summary dict)"""
    fs = [np.ascontiguousarray(f, np.float32) for f in fixed_scans]
    ws = [np.ascontiguousarray(w, np.float32) for w in window_scans]
    fptr = (C.c_void_p * len(fs))(*[f.ctypes.data for f in fs]); wptr = (C.c_void_p * len(ws))(*[w.ctypes.data for w in ws])
    nf = np.array([len(f) for f in fs], np.uint32); nw = np.array([len(w) for w in ws], np.uint32)
    fp = np.ascontiguousarray(fixed_poses, np.float64).reshape(len(fs), 4)
    st = np.ascontiguousarray(states14, np.float64).reshape(len(ws) + 1, 14).copy()
    t = np.ascontiguousarray(trans, np.float64).reshape(4).copy()
    im = None if imu is None else np.ascontiguousarray(imu, np.float64)
    tol = None if tolerances is None else np.ascontiguousarray(tolerances, np.float64)
    out = np.zeros(10)
    _check(lib().randt_hostapi_window_solve(C.c_int(device), C.byref(gp), fptr, _pf(nf), _pf(fp), C.c_uint32(len(fs)), wptr, _pf(nw), C.c_uint32(len(ws)),
                                            _pf(st), _pf(im) if im is not None else None, _pf(np.ascontiguousarray(params80, np.float64)),
                                            _pf(tol) if tol is not None else None, _pf(t), _pf(out)))
    keys = ("status", "rejected", "gnc_solves", "total_iterations", "final_cost", "mu_first", "max_residual", "n_tangent", "evaluations", "n_cells")
    return st, t, dict(zip(keys, out.tolist()))


def window_replay(gp, scans, stamps, params80, smoothing_steps=3, insertion_step=2, yaw=None, device=0):
    """LocalFuser::processScan in miniature over a drive: voxelise -> predictTransform -> estimateTransformCeres over the window -> delayed
    keyframe insertion at the smoothed pose -> (poses [n, 4] as estimated on arrival, states [n, 14] at the end, stats [n, 4], totals dict)"""
    sc = [np.ascontiguousarray(s, np.float32) for s in scans]
    n = len(sc)
    pts = np.ascontiguousarray(np.concatenate(sc), np.float32)
    off = np.concatenate([[0], np.cumsum([len(s) for s in sc])]).astype(np.uint32)
    st = np.ascontiguousarray(stamps, np.float64)
    yw = None if yaw is None else np.ascontiguousarray(yaw, np.float64)
    poses = np.zeros((n, 4)); states = np.zeros((n, 14)); stats = np.zeros((n, 4)); totals = np.zeros(6)
    _check(lib().randt_hostapi_window_replay(C.c_int(device), C.byref(gp), _pf(pts), _pf(off), C.c_uint32(n), _pf(st), _pf(yw) if yw is not None else None,
                                             _pf(np.ascontiguousarray(params80, np.float64)), C.c_int(int(smoothing_steps)), C.c_int(int(insertion_step)),
                                             _pf(poses), _pf(states), _pf(stats), _pf(totals)))
    return poses, states, stats, dict(seconds=totals[0], submap_cells=int(totals[1]), launches=int(totals[2]), keyframes=int(totals[3]),
                                       setup_seconds=totals[4], solve_seconds=totals[5])


def window_minimize_factors(states14, params80, imu=None, tolerances=None, max_iterations=0):
    """window::minimize on the motion-model (+ IMU) factors alone, no device -> (states [(W + 1), 14], summary dict)"""
    st = np.ascontiguousarray(states14, np.float64).reshape(-1, 14).copy()
    W = len(st) - 1
    im = None if imu is None else np.ascontiguousarray(imu, np.float64)
    tol = None if tolerances is None else np.ascontiguousarray(tolerances, np.float64)
    out = np.zeros(4)
    _check(lib().randt_hostapi_window_minimize_factors(_pf(st), C.c_uint32(W), _pf(im) if im is not None else None, _pf(np.ascontiguousarray(params80, np.float64)),
                                                       _pf(tol) if tol is not None else None, C.c_int(int(max_iterations)), _pf(out)))
    return st, dict(initial_cost=out[0], final_cost=out[1], iterations=int(out[2]), termination=int(out[3]))
