"""ctypes bindings over the C-ABI in include/randt_gpu.h (librandt_gpu.so, built in-tree by build.py).

This is plumbing for the tests and bench.py: every call goes straight to an exported `randt_*` symbol.  There is no CPU
fallback: if the shared library is missing or a call fails, a RandtError is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librandt_gpu.so")

VAR_SE2_INTENSITY, VAR_SE2_XY, VAR_VEC_INTENSITY, VAR_VEC_XY = 0, 1, 2, 3
LOSS_NONE, LOSS_BARRON, LOSS_WELSCH = 0, 1, 2
LOOKUP_MAHALANOBIS, LOOKUP_EUCLID = 0, 1
FUSED_STRIDE = 24
FUSED_H, FUSED_G, FUSED_COST, FUSED_MAXR, FUSED_SUMSQ, FUSED_N = 0, 16, 20, 21, 22, 23   # RANDT_FUSED_* of include/randt_gpu.h
CORE_STRIDE = 15                                                                         # RANDT_CORE_STRIDE: packed == 2 (H upper triangle, g, cost)
PACKED_STRIDE = 18                                                                       # RANDT_PACKED_*: H upper triangle (10), g (4), cost, max r, sum r^2, n
BASIS_STRIDE = 10                                                                        # RANDT_BASIS_*: packed == 3 (3x3 upper triangle, g, cost in the functor's basis)
_PACKED_SRC = [0, 1, 2, 3, 5, 6, 7, 10, 11, 15, 16, 17, 18, 19, 20, 21, 22, 23]


def basis_to_core(basis, poses):
    """[S, 10] basis records of RANDT_VAR_SE2_INTENSITY (packed == 3: derivatives w.r.t. theta = atan2(s, c), tx, ty) + the poses they were
    evaluated at -> [S, 15] core records (packed == 2: ambient H upper triangle, g, cost): H = F^T H_b F, g = F^T g_b with
    F = [[-s/n2, c/n2, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]"""
    b = np.asarray(basis, np.float64); q = np.asarray(poses, np.float64)
    S = len(b)
    n2 = q[:, 0] ** 2 + q[:, 1] ** 2
    F = np.zeros((S, 3, 4)); F[:, 0, 0] = -q[:, 1] / n2; F[:, 0, 1] = q[:, 0] / n2; F[:, 1, 2] = 1.0; F[:, 2, 3] = 1.0
    Hb = np.zeros((S, 3, 3))
    iu = np.triu_indices(3)
    Hb[:, iu[0], iu[1]] = b[:, 0:6]; Hb[:, iu[1], iu[0]] = b[:, 0:6]
    H = np.einsum("sia,sij,sjb->sab", F, Hb, F)
    g = np.einsum("sia,si->sa", F, b[:, 6:9])
    iu4 = np.triu_indices(4)
    return np.concatenate([H[:, iu4[0], iu4[1]], g, b[:, 9:10]], axis=1)


def pack_fused(full):
    """[S, 24] records -> [S, 18] (what randt_eval_fused_async(packed=1) delivers)"""
    return np.ascontiguousarray(np.asarray(full)[..., _PACKED_SRC])
E_INVALID, E_CUDA, E_CAPACITY, E_NONFINITE, E_NOMEM = -1, -2, -3, -4, -5


class RandtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("randt error %d: %s" % (code, msg))
        self.code = code


class Loss(C.Structure):
    _fields_ = [("kind", C.c_int32), ("scale", C.c_double), ("alpha", C.c_double), ("mu", C.c_double), ("weight", C.c_double)]


class GridParams(C.Structure):
    _fields_ = [("max_range", C.c_float), ("n_clusters", C.c_int32), ("min_points", C.c_int32), ("size_x", C.c_int32),
                ("size_y", C.c_int32), ("resolution", C.c_double), ("max_linf", C.c_double)]


class SolverOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("use_manifold", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
                ("jacobi_scaling", C.c_int32), ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("min_relative_decrease", C.c_double), ("gnc_loss_scale", C.c_double), ("gnc_divisor", C.c_double),
                ("gnc_max_steps", C.c_int32), ("poll_interval", C.c_int32)]


REG_SCORE, REG_FINAL_COST, REG_GNC_SOLVES, REG_ITERATIONS, REG_EVALS, REG_MU_FIRST, REG_STATUS, REG_TERMINATION, REG_STRIDE = range(9)


def solver_options(**kw):
    """ceres / reference defaults (randt_solver_options_default), overridden by keyword"""
    o = SolverOptions()
    lib().randt_solver_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class FilterParams(C.Structure):
    _fields_ = [("min_range", C.c_float), ("max_range", C.c_float), ("min_intensity", C.c_float), ("pad_", C.c_float),
                ("beam_distance_increment_threshold", C.c_double), ("sensor_to_base", C.c_float * 12)]


def filter_params(p, sensor_to_base=None):
    """from a params.NdtParams; sensor_to_base: 3x4 row-major Affine3f (identity by default)"""
    tf = np.eye(4, dtype=np.float32)[:3] if sensor_to_base is None else np.asarray(sensor_to_base, np.float32).reshape(3, 4)
    return FilterParams(float(p.min_range), float(p.max_range), float(p.min_intensity), 0.0, float(p.beam_distance_increment_threshold),
                        (C.c_float * 12)(*tf.reshape(12).tolist()))


def grid_params(p):
    """from a randt_slam_b200.params.NdtParams"""
    return GridParams(float(p.max_range), int(p.n_clusters), int(p.min_points_per_cell), int(p.size_x), int(p.size_y),
                      float(p.resolution), float(p.max_neighbor_linf_distance))


_lib = None

# every symbol include/randt_gpu.h declares (kept in sync by tests/test_abi.py)
_vp, _u32, _u64, _i, _d = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
_SIGS = {
    "randt_version": (_i, []),
    "randt_ctx_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "randt_ctx_destroy": (None, [_vp]),
    "randt_last_error": (C.c_char_p, [_vp]),
    "randt_ctx_stream": (_vp, [_vp]),
    "randt_ctx_sync": (_i, [_vp]),
    "randt_ctx_launch_count": (_u64, [_vp]),
    "randt_ctx_take_bad_pairs": (_i, [_vp, C.POINTER(_u64)]),
    "randt_host_alloc": (_vp, [C.c_size_t]),
    "randt_host_free": (None, [_vp]),
    "randt_dev_alloc": (_vp, [C.c_size_t]),
    "randt_dev_free": (None, [_vp]),
    "randt_memcpy_h2d": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "randt_memcpy_d2h": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "randt_filter_scan": (_i, [_vp, _vp, _u32, _u32, C.POINTER(FilterParams), _i, _vp, _i, _u32, C.POINTER(_u32)]),
    "randt_filter_scans": (_i, [_vp, _vp, _u32, _u32, _u32, C.POINTER(FilterParams), _i, _vp, _i, _u32, _vp]),
    "randt_voxelize": (_i, [_vp, _vp, _vp, _u32, C.POINTER(GridParams), _i, C.POINTER(_vp)]),
    "randt_map_upload": (_i, [_vp, _vp, _vp, _vp, _u32, _vp, C.POINTER(GridParams), C.POINTER(_vp)]),
    "randt_map_info": (_i, [_vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "randt_map_download": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "randt_map_transform": (_i, [_vp, _vp, _vp]),
    "randt_map_transform_se2d": (_i, [_vp, _vp, _vp]),
    "randt_map_merge": (_i, [_vp, _vp, _vp]),
    "randt_map_destroy": (None, [_vp]),
    "randt_cs_divergence": (_i, [_vp, _vp, _vp, _vp]),
    "randt_associate": (_i, [_vp, _vp, _vp, _vp, _i, _i, C.POINTER(_vp)]),
    "randt_problem_create": (_i, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _u32, _vp, _u32, C.POINTER(_vp)]),
    "randt_problem_info": (_i, [_vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "randt_problem_concat": (_i, [_vp, _vp, _u32, _vp, _u32, C.POINTER(_vp)]),
    "randt_register_batch_weighted": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "randt_points_from_pcl_xyzi": (_i, [_vp, _vp, _u32, _i, _vp]),
    "randt_scan_step": (_i, [_vp, _vp, _vp, _u32, _vp, _i, _i, _vp, C.c_double, _vp, _i, _vp, _vp, _vp]),
    "randt_eval_allpairs": (_i, [_vp, _vp, _vp, _i, _vp, _vp, C.c_double, _vp]),
    "randt_eval_allpairs_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp, C.c_double, _vp]),
    "randt_problem_layout": (_i, [_vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "randt_problem_schedule": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _u32, _vp]),
    "randt_problem_download": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "randt_problem_download_cells": (_i, [_vp, _vp, _vp, _vp]),
    "randt_problem_destroy": (None, [_vp]),
    "randt_eval_emit": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "randt_eval_emit_dev": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "randt_eval_fused": (_i, [_vp, _vp, _i, _vp, C.POINTER(Loss), _vp, _i, _vp]),
    "randt_ctx_async_count": (_u64, [_vp]),
    "randt_ctx_wait_async": (_i, [_vp, _u64]),
    "randt_eval_fused_async": (_i, [_vp, _vp, _i, _vp, C.POINTER(Loss), _vp, _i, _i, _vp]),
    "randt_eval_fused_dev": (_i, [_vp, _vp, _i, _vp, C.POINTER(Loss), _vp, _i, _vp]),
    "randt_sweep_costs": (_i, [_vp, _vp, _u32, _i, _vp, _u32, C.POINTER(Loss), _vp]),
    "randt_solver_options_default": (None, [_vp]),
    "randt_register_batch": (_i, [_vp, _vp, _i, _vp, C.POINTER(Loss), _vp, _vp]),
    "randt_register_batch_dev": (_i, [_vp, _vp, _i, _vp, C.POINTER(Loss), _vp, _vp]),
}


def lib():
    """Load librandt_gpu.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RandtError(E_INVALID, "CUDA extension %s is missing: run `python -m randt_slam_b200.build` "
                                        "(or __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32a(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def make_loss(kind=LOSS_NONE, scale=1.0, alpha=-2.0, mu=1.0, weight=1.0):
    return Loss(int(kind), float(scale), float(alpha), float(mu), float(weight))


class Context:
    """One randt_ctx (one CUDA stream).  stream: an int cudaStream_t to borrow, or None for an owned stream."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        rc = lib().randt_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h))
        if rc != 0:
            raise RandtError(rc, "randt_ctx_create failed (no CUDA device %d?)" % device)
        self.device = device

    def close(self):
        if self._h:
            lib().randt_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise RandtError(rc, lib().randt_last_error(self._h).decode())

    @property
    def stream(self):
        return lib().randt_ctx_stream(self._h)

    def sync(self):
        self._check(lib().randt_ctx_sync(self._h))

    @property
    def launch_count(self):
        return int(lib().randt_ctx_launch_count(self._h))

    def take_bad_pairs(self):
        n = C.c_uint64(0)
        self._check(lib().randt_ctx_take_bad_pairs(self._h, C.byref(n)))
        return int(n.value)

    # ---- K6 ----
    def filter_scan(self, raw4, n_azimuths, n_bins, fp, cap=None):
        """RadarPreprocessor::filterScan -> float32 [n, 4] filtered points in the base frame"""
        raw4 = _f32(raw4, (-1, 4))
        assert len(raw4) == n_azimuths * n_bins
        cap = int(cap if cap is not None else max(16, len(raw4)))
        out = np.zeros((cap, 4), np.float32)
        n = C.c_uint32(0)
        self._check(lib().randt_filter_scan(self._h, _ptr(raw4), int(n_azimuths), int(n_bins), C.byref(fp), 0, _ptr(out), 0, cap, C.byref(n)))
        return out[: n.value].copy()

    def points_from_pcl_xyzi(self, pcl_points, d_out4):
        """pcl::PointXYZI records (float32 [n, 8] host array: x, y, z, 1, intensity, pad x 3) -> float4 points at the device pointer d_out4"""
        a = _f32(pcl_points, (-1, 8))
        self._check(lib().randt_points_from_pcl_xyzi(self._h, _ptr(a), C.c_uint32(len(a)), 0, C.c_void_p(int(d_out4))))

    def filter_scan_dev(self, d_raw, n_azimuths, n_bins, fp, d_out, cap):
        """device pointers in and out (ints); returns the number of kept points"""
        n = C.c_uint32(0)
        self._check(lib().randt_filter_scan(self._h, _ptr(d_raw), int(n_azimuths), int(n_bins), C.byref(fp), 1, _ptr(d_out), 1, int(cap), C.byref(n)))
        return int(n.value)

    def filter_scans(self, raw4, n_scans, n_azimuths, n_bins, fp, cap=None):
        """n_scans scans of one shape back to back -> (float32 [n, 4] kept points scan after scan, uint32 [n_scans + 1] offsets)"""
        raw4 = _f32(raw4, (-1, 4))
        assert len(raw4) == n_scans * n_azimuths * n_bins
        cap = int(cap if cap is not None else max(16, len(raw4)))
        out = np.zeros((cap, 4), np.float32); off = np.zeros(n_scans + 1, np.uint32)
        self._check(lib().randt_filter_scans(self._h, _ptr(raw4), int(n_scans), int(n_azimuths), int(n_bins), C.byref(fp), 0, _ptr(out), 0, cap,
                                             _ptr(off)))
        return out[: off[-1]].copy(), off

    def filter_scans_dev(self, d_raw, n_scans, n_azimuths, n_bins, fp, d_out, cap):
        """device pointers in and out (ints); returns the uint32 [n_scans + 1] offsets (host)"""
        off = np.zeros(n_scans + 1, np.uint32)
        self._check(lib().randt_filter_scans(self._h, _ptr(d_raw), int(n_scans), int(n_azimuths), int(n_bins), C.byref(fp), 1, _ptr(d_out), 1,
                                             int(cap), _ptr(off)))
        return off

    def async_count(self):
        """ticket of the latest eval_fused_async call on this context"""
        return int(lib().randt_ctx_async_count(self._h))

    def wait_async(self, ticket):
        self._check(lib().randt_ctx_wait_async(self._h, C.c_uint64(int(ticket))))

    # ---- K1 ----
    def voxelize(self, pts, scan_off, gp, pts_on_device=False):
        """pts: float32 [N,4] host array (or int device pointer); scan_off: uint32 [B+1] -> Map"""
        scan_off = _u32a(scan_off)
        if not pts_on_device:
            pts = _f32(pts, (-1, 4))
        out = C.c_void_p()
        self._check(lib().randt_voxelize(self._h, _ptr(pts), _ptr(scan_off), len(scan_off) - 1, C.byref(gp), int(pts_on_device),
                                         C.byref(out)))
        return Map(self, out, gp)

    def map_upload(self, cells, cell_off, gp, npts=None, slot=None):
        cells = _f32(cells, (-1, 12)); cell_off = _u32a(cell_off)
        npts_a = _u32a(npts) if npts is not None else None
        slot_a = np.ascontiguousarray(slot, dtype=np.int32) if slot is not None else None
        out = C.c_void_p()
        self._check(lib().randt_map_upload(self._h, _ptr(cells), _ptr(npts_a), _ptr(cell_off), len(cell_off) - 1, _ptr(slot_a),
                                           C.byref(gp), C.byref(out)))
        return Map(self, out, gp)

    # ---- K2 ----
    def associate(self, fixed, moving, pose0, k, metric=LOOKUP_MAHALANOBIS):
        pose0 = _f64(pose0)
        out = C.c_void_p()
        self._check(lib().randt_associate(self._h, fixed._h, moving._h, _ptr(pose0), int(k), int(metric), C.byref(out)))
        return Problem(self, out)

    def problem_create(self, cells_m, cells_f, pair_m, pair_f, seg_off):
        cells_m = _f32(cells_m, (-1, 12)); cells_f = _f32(cells_f, (-1, 12))
        pair_m = _u32a(pair_m); pair_f = _u32a(pair_f); seg_off = _u32a(seg_off)
        out = C.c_void_p()
        self._check(lib().randt_problem_create(self._h, _ptr(cells_m), len(cells_m), _ptr(cells_f), len(cells_f), _ptr(pair_m),
                                               _ptr(pair_f), len(pair_m), _ptr(seg_off), len(seg_off) - 1, C.byref(out)))
        return Problem(self, out)

    def problem_concat(self, parts, seg_of_part, n_segments):
        """single-segment problems joined on the device: the pairs of part i go to segment seg_of_part[i] (non-decreasing)"""
        hs = (C.c_void_p * len(parts))(*[q._h.value for q in parts])
        so = _u32a(seg_of_part)
        out = C.c_void_p()
        self._check(lib().randt_problem_concat(self._h, hs, len(parts), _ptr(so), int(n_segments), C.byref(out)))
        return Problem(self, out)


class Map:
    def __init__(self, ctx, handle, gp):
        self.ctx, self._h, self.gp = ctx, handle, gp

    def close(self):
        if self._h:
            lib().randt_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        b, n, s = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().randt_map_info(self._h, C.byref(b), C.byref(n), C.byref(s))
        return b.value, n.value, s.value

    def download(self, want_labels=False, want_slot=True):
        B, n, ns = self.info()
        cells = np.zeros((n, 12), np.float32); npts = np.zeros(n, np.uint32); off = np.zeros(B + 1, np.uint32)
        labels = np.zeros(n, np.int32) if want_labels else None
        slot = np.zeros((B, ns), np.int32) if want_slot else None
        self.ctx._check(lib().randt_map_download(self.ctx._h, self._h, _ptr(cells), _ptr(npts), _ptr(labels), _ptr(off), _ptr(slot)))
        return dict(cells=cells, npts=npts, labels=labels, cell_off=off, slot=slot)

    def transform(self, trans):
        trans = _f32(trans, (-1, 4))
        assert len(trans) == self.info()[0]
        self.ctx._check(lib().randt_map_transform(self.ctx._h, self._h, _ptr(trans)))

    def transform_se2d(self, poses):
        """Map::transformMap(Eigen::Affine2f(pose.cast<float>().matrix())) for float64 [B, 4] Sophus SE2d poses"""
        poses = _f64(poses).reshape(-1, 4)
        assert len(poses) == self.info()[0]
        self.ctx._check(lib().randt_map_transform_se2d(self.ctx._h, self._h, _ptr(poses)))

    def merge(self, moving):
        self.ctx._check(lib().randt_map_merge(self.ctx._h, self._h, moving._h))

    def eval_allpairs(self, moving, poses, loss=None, window=0.0, variant=VAR_SE2_INTENSITY):
        """K8: self = fixed maps; every moving cell against every fixed cell (within `window` m, <= 0: all) -> fused records [B, 24]"""
        B = self.info()[0]
        npar = 4 if variant <= 1 else 3
        poses = _f64(poses).reshape(B, npar)
        out = np.zeros((B, FUSED_STRIDE), np.float64)
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_eval_allpairs(self.ctx._h, self._h, moving._h, int(variant), _ptr(poses), lp, float(window), _ptr(out)))
        return out

    def eval_allpairs_dev(self, moving, d_poses, d_out, loss=None, window=0.0, variant=VAR_SE2_INTENSITY):
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_eval_allpairs_dev(self.ctx._h, self._h, moving._h, int(variant), C.c_void_p(d_poses), lp, float(window), C.c_void_p(d_out)))

    def scan_step(self, pts, gp, k, loss, ndt_weight, opt, insert_keyframe, pose, metric=None):
        """randt_scan_step on this (single) submap: voxelise -> associate -> register -> optional keyframe merge.  -> (pose [4], result [REG_STRIDE], n_cells)"""
        pts = _f32(pts).reshape(-1, 4)
        pose = _f64(pose).reshape(4).copy()
        res = np.zeros(REG_STRIDE, np.float64)
        nc = C.c_uint32()
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_scan_step(self.ctx._h, self._h, _ptr(pts), C.c_uint32(len(pts)), C.byref(gp), int(k),
                                              int(LOOKUP_MAHALANOBIS if metric is None else metric), lp, float(ndt_weight), C.byref(opt), int(insert_keyframe),
                                              _ptr(pose), _ptr(res), C.byref(nc)))
        return pose, res, nc.value

    def cs_divergence(self, moving):
        """Map::calculateCSDivergence for every map pair of the batch -> float64 [B]"""
        out = np.zeros(self.info()[0], np.float64)
        self.ctx._check(lib().randt_cs_divergence(self.ctx._h, self._h, moving._h, _ptr(out)))
        return out


class Problem:
    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        s, p, m, f = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().randt_problem_info(handle, C.byref(s), C.byref(p), C.byref(m), C.byref(f))
        self.n_segments, self.n_pairs, self.n_m, self.n_f = s.value, p.value, m.value, f.value

    def close(self):
        if self._h:
            lib().randt_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def schedule(self):
        """the work schedule K3 walks, as built on the device -> dict (same keys as hostapi.build_schedule where they apply)"""
        counts = np.zeros(5, np.uint32)
        self.ctx._check(lib().randt_problem_schedule(self.ctx._h, self._h, _ptr(counts), None, None, 0, None, None, 0, None))
        nw, nt, na, nb, budget = (int(c) for c in counts)
        pa = np.zeros((max(na, 1), 4), np.uint32); pb = np.zeros((max(nb, 1), 4), np.uint32)
        wa = np.zeros(nw + 1, np.uint32); wb = np.zeros(nw + 1, np.uint32); doff = np.zeros(self.n_segments + 1, np.uint32)
        self.ctx._check(lib().randt_problem_schedule(self.ctx._h, self._h, _ptr(counts), _ptr(pa), _ptr(pb), max(na, nb, 1), _ptr(wa), _ptr(wb), nw, _ptr(doff)))
        return dict(n_warps=nw, n_tiles=nt, plan_a=pa[:na], plan_b=pb[:nb], woff_a=wa, woff_b=wb, duo_off=doff, warp_budget=budget)

    def layout(self):
        """-> (n_duos, record_bytes, n_overflow): the record table K3 streams"""
        d, b, o = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().randt_problem_layout(self._h, C.byref(d), C.byref(b), C.byref(o))
        return d.value, b.value, o.value

    def download(self):
        pm = np.zeros(self.n_pairs, np.uint32); pf = np.zeros(self.n_pairs, np.uint32); off = np.zeros(self.n_segments + 1, np.uint32)
        self.ctx._check(lib().randt_problem_download(self.ctx._h, self._h, _ptr(pm), _ptr(pf), _ptr(off)))
        return pm, pf, off

    def download_cells(self):
        cm = np.zeros((self.n_m, 12), np.float32); cf = np.zeros((self.n_f, 12), np.float32)
        self.ctx._check(lib().randt_problem_download_cells(self.ctx._h, self._h, _ptr(cm), _ptr(cf)))
        return cm, cf

    def eval_emit(self, poses, variant=VAR_SE2_INTENSITY, want_jac=True):
        npar = 4 if variant <= 1 else 3
        poses = _f64(poses).reshape(self.n_segments, npar)
        r = np.zeros(self.n_pairs, np.float64)
        J = np.zeros((self.n_pairs, npar), np.float64) if want_jac else None
        self.ctx._check(lib().randt_eval_emit(self.ctx._h, self._h, int(variant), _ptr(poses), _ptr(r), _ptr(J)))
        return r, J

    def eval_fused(self, poses, loss=None, mu_per_seg=None, want_jac=True, variant=VAR_SE2_INTENSITY, out=None):
        npar = 4 if variant <= 1 else 3
        poses = _f64(poses).reshape(self.n_segments, npar)
        if out is None:
            out = np.zeros((self.n_segments, FUSED_STRIDE), np.float64)
        mu = _f64(mu_per_seg) if mu_per_seg is not None else None
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_eval_fused(self.ctx._h, self._h, int(variant), _ptr(poses), lp, _ptr(mu), int(want_jac), _ptr(out)))
        return out

    def eval_fused_async(self, poses, out, loss=None, mu_per_seg=None, want_jac=True, variant=VAR_SE2_INTENSITY, packed=False):
        """enqueue-only evaluation: `poses`, `mu_per_seg` and `out` ([S, 24]; packed 1: [S, 18], 2: [S, 15], 3: [S, 10]) are float64 arrays in pinned host
        memory (PinnedArray); valid after Context.sync()"""
        for a in (poses, out, mu_per_seg):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
                raise TypeError("eval_fused_async needs C-contiguous float64 arrays")
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_eval_fused_async(self.ctx._h, self._h, int(variant), _ptr(poses), lp, _ptr(mu_per_seg), int(want_jac),
                                                     int(packed), _ptr(out)))

    def eval_fused_dev(self, d_poses, d_out, loss=None, d_mu=None, want_jac=True, variant=VAR_SE2_INTENSITY):
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_eval_fused_dev(self.ctx._h, self._h, int(variant), _ptr(d_poses), lp, _ptr(d_mu), int(want_jac),
                                                   _ptr(d_out)))

    def eval_emit_dev(self, d_poses, d_r, d_J, variant=VAR_SE2_INTENSITY):
        self.ctx._check(lib().randt_eval_emit_dev(self.ctx._h, self._h, int(variant), _ptr(d_poses), _ptr(d_r), _ptr(d_J)))

    def register_batch(self, poses, loss, opt, variant=VAR_SE2_INTENSITY, weights=None):
        """GNC + LM for every segment at once -> (poses [S, np], result [S, REG_STRIDE]); weights: one ScaledLoss weight per segment"""
        npar = 4 if variant <= 1 else 3
        poses = _f64(poses).reshape(self.n_segments, npar).copy()
        result = np.zeros((self.n_segments, REG_STRIDE), np.float64)
        lp = C.byref(loss) if loss is not None else None
        if weights is None:
            self.ctx._check(lib().randt_register_batch(self.ctx._h, self._h, int(variant), _ptr(poses), lp, C.byref(opt), _ptr(result)))
        else:
            w = _f64(weights).reshape(self.n_segments)
            self.ctx._check(lib().randt_register_batch_weighted(self.ctx._h, self._h, int(variant), _ptr(poses), lp, _ptr(w), C.byref(opt), _ptr(result)))
        return poses, result

    def register_batch_dev(self, d_poses, d_result, loss, opt, variant=VAR_SE2_INTENSITY):
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_register_batch_dev(self.ctx._h, self._h, int(variant), _ptr(d_poses), lp, C.byref(opt), _ptr(d_result)))

    def sweep_costs(self, seg, poses, loss=None, variant=VAR_SE2_INTENSITY):
        npar = 4 if variant <= 1 else 3
        poses = _f64(poses).reshape(-1, npar)
        cost = np.zeros(len(poses), np.float64)
        lp = C.byref(loss) if loss is not None else None
        self.ctx._check(lib().randt_sweep_costs(self.ctx._h, self._h, int(seg), int(variant), _ptr(poses), len(poses), lp, _ptr(cost)))
        return cost


def unpack_fused(o):
    o = np.asarray(o)
    return dict(H=o[..., :16].reshape(o.shape[:-1] + (4, 4)), g=o[..., 16:20], cost=o[..., 20], max_r=o[..., 21], sum_sq=o[..., 22],
                n=o[..., 23])


class PinnedArray:
    """float64 array in pinned host memory from randt_host_alloc (what the *_async entry points take); .a is the numpy view"""

    def __init__(self, shape):
        n = int(np.prod(shape))
        self._p = lib().randt_host_alloc(C.c_size_t(max(8, 8 * n)))
        if not self._p:
            raise MemoryError("randt_host_alloc failed")
        self.a = np.ctypeslib.as_array((C.c_double * n).from_address(self._p)).reshape(shape)
        self.a[...] = 0.0

    def close(self):
        if self._p:
            self.a = None
            lib().randt_host_free(C.c_void_p(self._p)); self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
