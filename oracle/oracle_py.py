"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  PARITY UNPINNED: see oracle/ndt_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

VAR_SE2_INTENSITY, VAR_SE2_XY, VAR_VEC_INTENSITY, VAR_VEC_XY = 0, 1, 2, 3
LOSS_NONE, LOSS_BARRON, LOSS_WELSCH = 0, 1, 2
LOOKUP_MAHALANOBIS, LOOKUP_EUCLID = 0, 1


def _host_sig():
    """-march=native ties the binary to the CPU it was built on; rebuild when the snapshot lands on another host."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import zlib; return "%08x" % zlib.crc32(line.strip().encode())
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """Compile the C++ restatement (g++, seconds).  Safe to call repeatedly."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "ndt_oracle.h", "lm_oracle.h", "window_oracle.h", "jet.h", "Makefile")]
    sig_path = os.path.join(_HERE, "_build", "host_sig.txt")
    sig = _host_sig()
    try:
        same_host = open(sig_path).read() == sig
    except OSError:
        same_host = False
    if (not force) and same_host and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    with open(sig_path, "w") as f:
        f.write(sig)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _sig(_lib)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _sig(L):
    f, d, i, u, sz = C.c_float, C.c_double, C.c_int, C.c_uint32, C.c_size_t
    pf, pd, pi, pu = C.POINTER(f), C.POINTER(d), C.POINTER(C.c_int32), C.POINTER(u)
    L.orc_n_clusters.restype = i; L.orc_n_clusters.argtypes = [d, d]
    L.orc_grid_row_size.restype = i; L.orc_grid_row_size.argtypes = [sz]
    L.orc_grid_labels.restype = None; L.orc_grid_labels.argtypes = [pf, sz, sz, f, pi]
    L.orc_coord_to_index.restype = u; L.orc_coord_to_index.argtypes = [i, i, d, f, f]
    L.orc_sym2_eigen.restype = None; L.orc_sym2_eigen.argtypes = [f, f, f, pf, pf]
    L.orc_cell_from_points.restype = i; L.orc_cell_from_points.argtypes = [pf, sz, i, pf]
    L.orc_voxelize.restype = i
    L.orc_voxelize.argtypes = [pf, sz, sz, f, i, i, i, d, d, pf, pu, pi, pi, i, C.POINTER(i)]
    L.orc_se2d_cast_float.restype = None; L.orc_se2d_cast_float.argtypes = [pd, pf]
    L.orc_affine_rotation.restype = None; L.orc_affine_rotation.argtypes = [f, f, pf]
    L.orc_transform_cells.restype = None; L.orc_transform_cells.argtypes = [pf, sz, f, f, f, f]
    L.orc_merge_map_cell.restype = i; L.orc_merge_map_cell.argtypes = [pf, pu, i, i, pi, i, i, d, pf, pu, i]
    L.orc_associate.restype = i; L.orc_associate.argtypes = [pf, i, pi, i, i, d, d, pf, i, pd, i, i, pu, pu, i]
    L.orc_eval_pairs.restype = None; L.orc_eval_pairs.argtypes = [i, i, pf, pf, pu, pu, sz, pd, pd, pd]
    L.orc_loss_eval.restype = None; L.orc_loss_eval.argtypes = [i, d, d, d, d, d, pd]
    L.orc_corrector.restype = None; L.orc_corrector.argtypes = [d, pd, pd]
    L.orc_gnc_initial_mu.restype = d; L.orc_gnc_initial_mu.argtypes = [d, d, d, i]
    L.orc_fused.restype = None; L.orc_fused.argtypes = [i, pf, pf, pu, pu, sz, pd, i, d, d, d, d, i, pd]
    L.orc_hw_threads.restype = i; L.orc_hw_threads.argtypes = []
    L.orc_fused_batch.restype = d
    L.orc_fused_batch.argtypes = [i, pf, pf, pu, pu, pu, i, pd, i, d, d, d, d, pd, i, pd, i, i]
    L.orc_loop_constraint.restype = i
    L.orc_loop_constraint.argtypes = [pf, i, pi, i, i, d, d, pf, i, pd, i, i, i, d, d, d, d, i, i, i, d, pu, pu, i, pd]
    L.orc_sweep_costs.restype = None; L.orc_sweep_costs.argtypes = [i, pf, pf, pu, pu, sz, i, d, d, d, d, pd, sz, pd]
    L.orc_bnb.restype = i; L.orc_bnb.argtypes = [pf, i, pi, i, i, d, d, pf, i, pd, i, d, d, d, d, d, d, d, i, pd]
    L.orc_cs_divergence.restype = d; L.orc_cs_divergence.argtypes = [pf, sz, pf, sz, pd]
    L.orc_filter_scan.restype = i; L.orc_filter_scan.argtypes = [pf, sz, f, f, f, d, pf, pf, sz, C.POINTER(sz)]
    L.orc_se2_plus.restype = None; L.orc_se2_plus.argtypes = [pd, pd, pd]
    L.orc_se2_plus_jacobian.restype = None; L.orc_se2_plus_jacobian.argtypes = [pd, pd]
    L.orc_set_tolerances.restype = None; L.orc_set_tolerances.argtypes = [d, d, d]
    L.orc_window_evaluate.restype = i; L.orc_window_evaluate.argtypes = [pf, pf, pu, pu, pu, i, pd, pd, pd, i, d, d, pd, pd, pd, pd]
    L.orc_factor_block.restype = i; L.orc_factor_block.argtypes = [i, i, pd, pd, pd, d, d, d, pd, pd]
    L.orc_window_minimize_factors.restype = i; L.orc_window_minimize_factors.argtypes = [pd, i, pd, pd, i, pd]
    L.orc_window_solve.restype = i; L.orc_window_solve.argtypes = [pf, pf, pu, pu, pu, i, pd, pd, pd, sz, pd, pd]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def n_clusters(max_range, resolution):
    return lib().orc_n_clusters(float(max_range), float(resolution))


def grid_labels(pts, n_clusters_, max_range):
    pts = _f32(pts).reshape(-1, 4)
    out = np.empty(len(pts), np.int32)
    lib().orc_grid_labels(_p(pts, C.c_float), len(pts), int(n_clusters_), float(max_range), _p(out, C.c_int32))
    return out


def coord_to_index(size_x, size_y, res, x, y):
    return lib().orc_coord_to_index(size_x, size_y, float(res), float(x), float(y))


def sym2_eigen(a00, a10, a11):
    ev = np.empty(2, np.float32); V = np.empty(4, np.float32)
    lib().orc_sym2_eigen(float(a00), float(a10), float(a11), _p(ev, C.c_float), _p(V, C.c_float))
    return ev, V.reshape(2, 2)


def cell_from_points(pts, min_points):
    pts = _f32(pts).reshape(-1, 4)
    out = np.empty(12, np.float32)
    ok = lib().orc_cell_from_points(_p(pts, C.c_float), len(pts), int(min_points), _p(out, C.c_float))
    return out if ok else None


def voxelize(pts, n_clusters_, max_range, min_points, size_x, size_y, res, max_linf=0.0):
    """-> dict(cells [N,12] f32, npts [N] u32, labels [N] i32, slot [size_x*size_y] i32, dropped)"""
    pts = _f32(pts).reshape(-1, 4)
    cap = max(16, len(pts))
    cells = np.zeros((cap, 12), np.float32); npts = np.zeros(cap, np.uint32); labels = np.zeros(cap, np.int32)
    slot = np.full(size_x * size_y, -1, np.int32); dropped = C.c_int(0)
    n = lib().orc_voxelize(_p(pts, C.c_float), len(pts), int(n_clusters_), float(max_range), int(min_points), size_x, size_y,
                           float(res), float(max_linf), _p(cells, C.c_float), _p(npts, C.c_uint32), _p(labels, C.c_int32),
                           _p(slot, C.c_int32), cap, C.byref(dropped))
    assert n >= 0
    return dict(cells=cells[:n].copy(), npts=npts[:n].copy(), labels=labels[:n].copy(), slot=slot, dropped=dropped.value)


def se2d_cast_float(pose):
    """Sophus SE2d::cast<float>() restated -> float32 (re, im, tx, ty): what Eigen::Affine2f(pose.cast<float>().matrix()) holds"""
    pose = _f64(pose); out = np.zeros(4, np.float32)
    lib().orc_se2d_cast_float(_p(pose, C.c_double), _p(out, C.c_float))
    return out


def affine_rotation(c, s):
    """Eigen Transform<float,3,Affine>::rotation() of the lift of [[c,-s],[s,c]] restated -> [3, 3] float32"""
    out = np.zeros(9, np.float32)
    lib().orc_affine_rotation(float(c), float(s), _p(out, C.c_float))
    return out.reshape(3, 3)


def transform_cells(cells, c, s, tx, ty):
    out = _f32(cells).reshape(-1, 12).copy()
    lib().orc_transform_cells(_p(out, C.c_float), len(out), float(c), float(s), float(tx), float(ty))
    return out


def merge_map_cell(f_cells, f_npts, f_slot, size_x, size_y, res, m_cells, m_npts):
    n_f, n_m = len(f_cells), len(m_cells)
    cap = n_f + n_m + 1
    fc = np.zeros((cap, 12), np.float32); fc[:n_f] = f_cells
    fn = np.zeros(cap, np.uint32); fn[:n_f] = f_npts
    fs = _i32(f_slot).copy()
    mc = _f32(m_cells).reshape(-1, 12); mn = _u32(m_npts)
    n = lib().orc_merge_map_cell(_p(fc, C.c_float), _p(fn, C.c_uint32), n_f, cap, _p(fs, C.c_int32), size_x, size_y, float(res),
                                 _p(mc, C.c_float), _p(mn, C.c_uint32), n_m)
    assert n >= 0
    return fc[:n].copy(), fn[:n].copy(), fs


def associate(f_cells, f_slot, size_x, size_y, res, max_linf, m_cells, pose, k, metric=LOOKUP_MAHALANOBIS):
    fc = _f32(f_cells).reshape(-1, 12); mc = _f32(m_cells).reshape(-1, 12); fs = _i32(f_slot); pose = _f64(pose)
    cap = max(1, len(mc) * max(k, 1))
    im = np.zeros(cap, np.uint32); jf = np.zeros(cap, np.uint32)
    P = lib().orc_associate(_p(fc, C.c_float), len(fc), _p(fs, C.c_int32), size_x, size_y, float(res), float(max_linf),
                            _p(mc, C.c_float), len(mc), _p(pose, C.c_double), int(k), int(metric), _p(im, C.c_uint32),
                            _p(jf, C.c_uint32), cap)
    assert P >= 0
    return im[:P].copy(), jf[:P].copy()


def eval_pairs(variant, cells_m, cells_f, im, jf, params, mode=0):
    """mode 0 autodiff (Jet, = what ceres executes), 1 closed form, 2 value only -> r [P], J [P, np]"""
    cm = _f32(cells_m).reshape(-1, 12); cf = _f32(cells_f).reshape(-1, 12); im = _u32(im); jf = _u32(jf); params = _f64(params)
    npar = 4 if variant in (VAR_SE2_INTENSITY, VAR_SE2_XY) else 3
    r = np.zeros(len(im), np.float64); J = np.zeros((len(im), npar), np.float64)
    lib().orc_eval_pairs(variant, mode, _p(cm, C.c_float), _p(cf, C.c_float), _p(im, C.c_uint32), _p(jf, C.c_uint32), len(im),
                         _p(params, C.c_double), _p(r, C.c_double), _p(J, C.c_double))
    return r, J


def loss_eval(kind, a, alpha, mu, weight, s):
    rho = np.zeros(3, np.float64)
    lib().orc_loss_eval(kind, float(a), float(alpha), float(mu), float(weight), float(s), _p(rho, C.c_double))
    return rho


def corrector(sq_norm, rho):
    rho = _f64(rho); out = np.zeros(3, np.float64)
    lib().orc_corrector(float(sq_norm), _p(rho, C.c_double), _p(out, C.c_double))
    return out


def gnc_initial_mu(max_residual, loss_scale, divisor, steps):
    return lib().orc_gnc_initial_mu(float(max_residual), float(loss_scale), float(divisor), int(steps))


def _unpack_fused(o):
    return dict(H=o[..., :16].reshape(o.shape[:-1] + (4, 4)).copy(), g=o[..., 16:20].copy(), cost=o[..., 20].copy(),
                max_r=o[..., 21].copy(), sum_sq=o[..., 22].copy(), n=o[..., 23].copy())


def fused(variant, cells_m, cells_f, im, jf, params, loss=(LOSS_NONE, 1.0, -2.0, 1.0, 1.0), want_jac=True):
    cm = _f32(cells_m).reshape(-1, 12); cf = _f32(cells_f).reshape(-1, 12); im = _u32(im); jf = _u32(jf); params = _f64(params)
    out = np.zeros(24, np.float64)
    lib().orc_fused(variant, _p(cm, C.c_float), _p(cf, C.c_float), _p(im, C.c_uint32), _p(jf, C.c_uint32), len(im),
                    _p(params, C.c_double), int(loss[0]), float(loss[1]), float(loss[2]), float(loss[3]), float(loss[4]),
                    int(want_jac), _p(out, C.c_double))
    return _unpack_fused(out)


def hw_threads():
    return lib().orc_hw_threads()


def fused_batch(variant, cells_m, cells_f, im, jf, seg_off, poses, loss=(LOSS_NONE, 1.0, -2.0, 1.0, 1.0), mu_per_seg=None,
                want_jac=True, n_threads=1, repeats=1):
    """-> (dict of per-segment arrays, wall seconds for `repeats` passes)"""
    cm = _f32(cells_m).reshape(-1, 12); cf = _f32(cells_f).reshape(-1, 12); im = _u32(im); jf = _u32(jf)
    seg_off = _u32(seg_off); poses = _f64(poses)
    S = len(seg_off) - 1
    out = np.zeros((S, 24), np.float64)
    mu_p = None
    if mu_per_seg is not None:
        mu_arr = _f64(mu_per_seg); mu_p = _p(mu_arr, C.c_double)
    t = lib().orc_fused_batch(variant, _p(cm, C.c_float), _p(cf, C.c_float), _p(im, C.c_uint32), _p(jf, C.c_uint32),
                              _p(seg_off, C.c_uint32), S, _p(poses, C.c_double), int(loss[0]), float(loss[1]), float(loss[2]),
                              float(loss[3]), float(loss[4]), mu_p, int(want_jac), _p(out, C.c_double), int(n_threads), int(repeats))
    return _unpack_fused(out), t


def loop_constraint(f_cells, f_slot, size_x, size_y, res, max_linf, m_cells, pose, k, metric=LOOKUP_MAHALANOBIS,
                    variant=VAR_SE2_INTENSITY, matcher_loss_scale=1.0, loop_scale=1.0, alpha=-2.0, divisor=1.1, max_gnc_steps=2,
                    max_iterations=200, on_manifold=False, loss_weight=1.0, pairs=None, tolerances=None):
    """tolerances: (function, parameter, gradient) overriding ceres' defaults (1e-6, 1e-8, 1e-10) for this call"""
    fc = _f32(f_cells).reshape(-1, 12); mc = _f32(m_cells).reshape(-1, 12); fs = _i32(f_slot); pose = _f64(pose)
    out = np.zeros(9, np.float64)
    if pairs is not None:
        im, jf = _u32(pairs[0]), _u32(pairs[1]); pim, pjf, P = _p(im, C.c_uint32), _p(jf, C.c_uint32), len(im)
    else:
        pim, pjf, P = None, None, 0
    if tolerances is not None:
        lib().orc_set_tolerances(C.c_double(tolerances[0]), C.c_double(tolerances[1]), C.c_double(tolerances[2]))
    st = lib().orc_loop_constraint(_p(fc, C.c_float), len(fc), _p(fs, C.c_int32), size_x, size_y, float(res), float(max_linf),
                                   _p(mc, C.c_float), len(mc), _p(pose, C.c_double), int(k), int(metric), int(variant),
                                   float(matcher_loss_scale), float(loop_scale), float(alpha), float(divisor), int(max_gnc_steps),
                                   int(max_iterations), int(on_manifold), float(loss_weight), pim, pjf, P, _p(out, C.c_double))
    if tolerances is not None:
        lib().orc_set_tolerances(C.c_double(0.0), C.c_double(0.0), C.c_double(0.0))
    return dict(pose=out[:4].copy(), score=out[4], gnc_solves=int(out[5]), iterations=int(out[6]), evals=int(out[7]),
                mu_first=out[8], status=st)


def sweep_costs(variant, cells_m, cells_f, im, jf, loss, poses):
    cm = _f32(cells_m).reshape(-1, 12); cf = _f32(cells_f).reshape(-1, 12); im = _u32(im); jf = _u32(jf); poses = _f64(poses)
    npar = 4 if variant in (VAR_SE2_INTENSITY, VAR_SE2_XY) else 3
    S = poses.size // npar
    out = np.zeros(S, np.float64)
    lib().orc_sweep_costs(variant, _p(cm, C.c_float), _p(cf, C.c_float), _p(im, C.c_uint32), _p(jf, C.c_uint32), len(im),
                          int(loss[0]), float(loss[1]), float(loss[2]), float(loss[3]), float(loss[4]), _p(poses, C.c_double), S,
                          _p(out, C.c_double))
    return out


def se2_plus(T, d):
    T = _f64(T); d = _f64(d); out = np.zeros(4)
    lib().orc_se2_plus(_p(T, C.c_double), _p(d, C.c_double), _p(out, C.c_double))
    return out


def bnb(f_cells, f_slot, size_x, size_y, res, max_linf, m_cells, pose, alpha, scale, window_linear=4.5, window_angular=0.45, linear_step=0.4,
        max_px_range=4.0, cost_threshold=0.82, n_iter=2, variant=VAR_SE2_INTENSITY):
    """Matcher::estimateTransformGlobalBNB restated sequentially -> dict(pose, min_cost, n_evaluated)"""
    fc = _f32(f_cells).reshape(-1, 12); mc = _f32(m_cells).reshape(-1, 12); fs = _i32(f_slot); pose = _f64(pose)
    out = np.zeros(6, np.float64)
    lib().orc_bnb(_p(fc, C.c_float), len(fc), _p(fs, C.c_int32), size_x, size_y, float(res), float(max_linf), _p(mc, C.c_float), len(mc),
                  _p(pose, C.c_double), int(variant), float(alpha), float(scale), float(window_linear), float(window_angular), float(linear_step),
                  float(max_px_range), float(cost_threshold), int(n_iter), _p(out, C.c_double))
    return dict(pose=out[:4].copy(), min_cost=out[4], n_evaluated=int(out[5]))


def cs_divergence(f_cells, m_cells):
    """Map::calculateCSDivergence restated -> (divergence, [interaction, fixed, moving] terms)"""
    fc = _f32(f_cells).reshape(-1, 12); mc = _f32(m_cells).reshape(-1, 12)
    terms = np.zeros(3, np.float64)
    v = lib().orc_cs_divergence(_p(fc, C.c_float), len(fc), _p(mc, C.c_float), len(mc), _p(terms, C.c_double))
    return v, terms


def filter_scan(raw4, min_range, max_range, min_intensity, beam_thr, tf12=None):
    """RadarPreprocessor::filterScan restated -> (filtered points [n, 4] in the base frame, number of emitted peaks)"""
    raw = _f32(raw4).reshape(-1, 4)
    tf = _f32(np.eye(4)[:3] if tf12 is None else tf12).reshape(12)
    out = np.zeros((max(16, len(raw)), 4), np.float32)
    npk = C.c_size_t(0)
    n = lib().orc_filter_scan(_p(raw, C.c_float), len(raw), float(min_range), float(max_range), float(min_intensity), float(beam_thr),
                              _p(tf, C.c_float), _p(out, C.c_float), len(out), C.byref(npk))
    assert n >= 0
    return out[:n].copy(), int(npk.value)


def window_evaluate(states14, params80, imu=None, cells_m=None, cells_f=None, im=None, jf=None, seg_off=None, loss_kind=LOSS_NONE, mu=1.0, weight=1.0):
    """One evaluation of estimateTransformCeres' joint problem at the given states [(W + 1), 14] (window_oracle.h): cost, tangent gradient,
    tangent J^T J, max raw NDT residual.  Without pair lists only the motion-model / IMU factors are evaluated."""
    st = _f64(states14).reshape(-1, 14)
    W = len(st) - 1
    cm = _f32(cells_m if cells_m is not None else np.zeros((1, 12))); cf = _f32(cells_f if cells_f is not None else np.zeros((1, 12)))
    im_ = _u32(im if im is not None else np.zeros(1)); jf_ = _u32(jf if jf is not None else np.zeros(1))
    so = _u32(seg_off if seg_off is not None else np.zeros(W + 1))
    imu_ = _f64(imu if imu is not None else np.zeros(max(W, 1)))
    cap = 10 * (W + 1)
    cost = np.zeros(1); g = np.zeros(cap); H = np.zeros(cap * cap); mr = np.zeros(1)
    nt = lib().orc_window_evaluate(_p(cm, C.c_float), _p(cf, C.c_float), _p(im_, C.c_uint32), _p(jf_, C.c_uint32), _p(so, C.c_uint32), W,
                                   _p(st, C.c_double), _p(imu_, C.c_double), _p(_f64(params80), C.c_double), int(loss_kind), float(mu), float(weight),
                                   _p(cost, C.c_double), _p(g, C.c_double), _p(H, C.c_double), _p(mr, C.c_double))
    return float(cost[0]), g[:nt].copy(), H[: nt * nt].reshape(nt, nt).copy(), float(mr[0])


def window_solve(states14, params80, trans, cells_m, cells_f, im, jf, seg_off, n_cells_total, imu=None, tolerances=None):
    """Matcher::estimateTransformCeres restated (window_oracle.h) on pre-associated pair lists (segment w = the blocks of window state w + 1)
    -> (states [(W + 1), 14], trans [4], summary dict)"""
    st = _f64(states14).reshape(-1, 14).copy()
    W = len(st) - 1
    t = _f64(trans).reshape(4).copy()
    imu_ = _f64(imu if imu is not None else np.zeros(max(W, 1)))
    out = np.zeros(8)
    if tolerances is not None:
        lib().orc_set_tolerances(C.c_double(tolerances[0]), C.c_double(tolerances[1]), C.c_double(tolerances[2]))
    try:
        lib().orc_window_solve(_p(_f32(cells_m), C.c_float), _p(_f32(cells_f), C.c_float), _p(_u32(im), C.c_uint32), _p(_u32(jf), C.c_uint32),
                               _p(_u32(seg_off), C.c_uint32), W, _p(st, C.c_double), _p(imu_, C.c_double), _p(_f64(params80), C.c_double),
                               C.c_size_t(int(n_cells_total)), _p(t, C.c_double), _p(out, C.c_double))
    finally:
        if tolerances is not None:
            lib().orc_set_tolerances(C.c_double(0.0), C.c_double(0.0), C.c_double(0.0))
    keys = ("status", "rejected", "gnc_solves", "total_iterations", "final_cost", "mu_first", "max_residual", "n_tangent")
    return st, t, dict(zip(keys, out.tolist()))


def factor_block(kind, manifold, a14, b14, sqrtI64, imu_rot=0.0, weight_imu=0.0, weight_bias=0.0):
    """one motion-model (kind 0) or IMU (kind 1) factor between two states: residuals [n] and the ambient Jacobian [n, 20] over the
    parameter slots 10 * side + (pose 0-3 | pos 0-1, rot 2 | lin_vel 4-5 | rot_vel 6 | lin_acc 7-8 | imu_bias 9)"""
    res = np.zeros(8); jac = np.zeros((8, 20))
    n = lib().orc_factor_block(int(kind), int(bool(manifold)), _p(_f64(a14), C.c_double), _p(_f64(b14), C.c_double), _p(_f64(sqrtI64), C.c_double),
                               float(imu_rot), float(weight_imu), float(weight_bias), _p(res, C.c_double), _p(jac, C.c_double))
    return res[:n].copy(), jac[:n].copy()


def window_minimize_factors(states14, params80, imu=None, tolerances=None, max_iterations=0):
    """lm_oracle.h's minimiser on the motion-model (+ IMU) factors of the window alone -> (states [(W + 1), 14], summary dict)"""
    st = _f64(states14).reshape(-1, 14).copy()
    W = len(st) - 1
    imu_ = _f64(imu if imu is not None else np.zeros(max(W, 1)))
    out = np.zeros(4)
    if tolerances is not None:
        lib().orc_set_tolerances(C.c_double(tolerances[0]), C.c_double(tolerances[1]), C.c_double(tolerances[2]))
    try:
        lib().orc_window_minimize_factors(_p(st, C.c_double), W, _p(imu_, C.c_double), _p(_f64(params80), C.c_double), int(max_iterations), _p(out, C.c_double))
    finally:
        if tolerances is not None:
            lib().orc_set_tolerances(C.c_double(0.0), C.c_double(0.0), C.c_double(0.0))
    return st, dict(initial_cost=out[0], final_cost=out[1], iterations=int(out[2]), termination=int(out[3]))
