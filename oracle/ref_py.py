"""ORACLE / _ref — TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_ref/libref.so: the reference's OWN ceres_loss_functions.cpp,
grid.cpp, radar_preprocessor.cpp, ndt_cell.cpp and ndt_map.cpp, compiled unmodified from /root/reference against the shim headers
under oracle/shim/ (oracle/Makefile, target `ref`; harness: oracle/ref_harness.cpp).

Only tests/ and tests/golden/gen_ref_golden.py import this module.  /root/reference exists in the build container only: there
build() compiles the library; on the GPU box the prebuilt oracle/_ref/libref.so travels with the snapshot, and available() is
False when neither is there (the tests then fall back to the committed fixtures tests/golden/ref_golden.npz this library produced).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
REF_ROOT = os.environ.get("RANDT_REFERENCE", "/root/reference/ros/ndt_radar_slam")

LOOKUP_MAHALANOBIS, LOOKUP_EUCLID = 0, 1


def build(force=False):
    """Compile the reference sources where they lie (g++, ~10 s).  No-op when /root/reference is absent."""
    if not os.path.isdir(REF_ROOT):
        return _LIB_PATH if os.path.exists(_LIB_PATH) else None
    deps = [os.path.join(_HERE, "ref_harness.cpp"), os.path.join(_HERE, "Makefile")]
    for d, _, fs in os.walk(os.path.join(_HERE, "shim")):
        deps += [os.path.join(d, f) for f in fs]
    if (not force) and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in deps):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s", "ref", "REF=" + REF_ROOT] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def available():
    return os.path.exists(_LIB_PATH) or os.path.isdir(REF_ROOT)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        d, i, u, f = C.c_double, C.c_int, C.c_uint32, C.c_float
        pf, pd, pi, pu, vp = C.POINTER(f), C.POINTER(d), C.POINTER(C.c_int32), C.POINTER(u), C.c_void_p
        L.ref_barron_evaluate.restype = None; L.ref_barron_evaluate.argtypes = [d, d, d, i, pd, i, pd]
        L.ref_welsch_evaluate.restype = None; L.ref_welsch_evaluate.argtypes = [d, d, i, pd, i, pd]
        L.ref_grid_cluster.restype = None; L.ref_grid_cluster.argtypes = [pf, u, C.c_uint64, d, pi]
        L.ref_map_from_scan.restype = vp; L.ref_map_from_scan.argtypes = [pf, u, i, d, i, i, i, d, d, C.POINTER(i)]
        L.ref_map_empty.restype = vp; L.ref_map_empty.argtypes = [i, i, i, d, d]
        L.ref_map_clone.restype = vp; L.ref_map_clone.argtypes = [vp]
        L.ref_map_free.restype = None; L.ref_map_free.argtypes = [vp]
        L.ref_map_n_cells.restype = u; L.ref_map_n_cells.argtypes = [vp]
        L.ref_map_n_slots.restype = u; L.ref_map_n_slots.argtypes = [vp]
        L.ref_map_get.restype = None; L.ref_map_get.argtypes = [vp, pf, pu, pi]
        L.ref_map_transform.restype = None; L.ref_map_transform.argtypes = [vp, pd]
        L.ref_affine_from_se2d.restype = None; L.ref_affine_from_se2d.argtypes = [pd, pf]
        L.ref_rotation_of_affine.restype = None; L.ref_rotation_of_affine.argtypes = [pf, pf]
        L.ref_map_merge.restype = i; L.ref_map_merge.argtypes = [vp, vp]
        L.ref_closest_cells.restype = i; L.ref_closest_cells.argtypes = [vp, vp, u, pd, i, i, C.POINTER(C.c_uint64)]
        L.ref_cs_divergence.restype = d; L.ref_cs_divergence.argtypes = [vp, vp]
        L.ref_filter_scan.restype = i; L.ref_filter_scan.argtypes = [pf, u, u, d, d, d, d, pf, pf, u, pd, pd, pu]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def barron(a, alpha, mu, s):
    """ceres::BarronLoss(a, alpha, mu).Evaluate(s) -> [n, 3] (rho, rho', rho'');  mu=None uses the two-argument constructor"""
    s = np.ascontiguousarray(np.atleast_1d(s), np.float64); out = np.zeros((len(s), 3))
    lib().ref_barron_evaluate(float(a), float(alpha), float(1.0 if mu is None else mu), int(mu is not None), _p(s, C.c_double), len(s), _p(out, C.c_double))
    return out


def welsch(a, mu, s):
    s = np.ascontiguousarray(np.atleast_1d(s), np.float64); out = np.zeros((len(s), 3))
    lib().ref_welsch_evaluate(float(a), float(1.0 if mu is None else mu), int(mu is not None), _p(s, C.c_double), len(s), _p(out, C.c_double))
    return out


def grid_cluster(pts, n_clusters, max_range):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4); lab = np.zeros(len(pts), np.int32)
    lib().ref_grid_cluster(_p(pts, C.c_float), len(pts), int(n_clusters), float(max_range), _p(lab, C.c_int32))
    return lab


class RefMap:
    """rc::navigation::ndt::Map of the reference, held by the harness."""

    def __init__(self, h):
        if not h:
            raise RuntimeError("the reference threw (cell mean outside the map)")
        self._h = h

    @classmethod
    def from_scan(cls, pts, n_clusters, max_range, min_points, size_x, size_y, res, max_linf=0.0):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4); err = C.c_int(0)
        return cls(lib().ref_map_from_scan(_p(pts, C.c_float), len(pts), int(n_clusters), float(max_range), int(min_points), int(size_x), int(size_y),
                                           float(res), float(max_linf), C.byref(err)))

    @classmethod
    def empty(cls, min_points, size_x, size_y, res, max_linf=0.0):
        return cls(lib().ref_map_empty(int(min_points), int(size_x), int(size_y), float(res), float(max_linf)))

    def clone(self):
        return RefMap(lib().ref_map_clone(self._h))

    def close(self):
        if self._h:
            lib().ref_map_free(self._h); self._h = None

    __del__ = close

    def get(self):
        n, ns = lib().ref_map_n_cells(self._h), lib().ref_map_n_slots(self._h)
        cells = np.zeros((n, 12), np.float32); npts = np.zeros(n, np.uint32); slot = np.zeros(ns, np.int32)
        lib().ref_map_get(self._h, _p(cells, C.c_float), _p(npts, C.c_uint32), _p(slot, C.c_int32))
        return dict(cells=cells, npts=npts, slot=slot)

    def transform(self, pose):
        pose = np.ascontiguousarray(pose, np.float64)
        lib().ref_map_transform(self._h, _p(pose, C.c_double))

    def merge(self, moving):
        if lib().ref_map_merge(self._h, moving._h):
            raise RuntimeError("the reference threw in mergeMapCell")

    def closest_cells(self, moving, i, pose, k, metric=LOOKUP_MAHALANOBIS):
        pose = np.ascontiguousarray(pose, np.float64); idx = np.zeros(max(k, 1) + 8, np.uint64)
        n = lib().ref_closest_cells(self._h, moving._h, int(i), _p(pose, C.c_double), int(k), int(metric), _p(idx, C.c_uint64))
        if n < 0:
            raise RuntimeError("the reference threw in getClosestCells")
        return idx[:n].astype(np.uint32)

    def associate(self, moving, pose, k, metric=LOOKUP_MAHALANOBIS):
        """the pair list Matcher::addNDTFactor builds (ndt_matcher.cpp:200-246): -> (im, jf) in residual-block order"""
        im, jf = [], []
        for i in range(lib().ref_map_n_cells(moving._h)):
            nb = self.closest_cells(moving, i, pose, k, metric)
            im += [i] * len(nb); jf += list(nb)
        return np.array(im, np.uint32), np.array(jf, np.uint32)


def affine_from_se2d(pose):
    pose = np.ascontiguousarray(pose, np.float64); out = np.zeros(4, np.float32)
    lib().ref_affine_from_se2d(_p(pose, C.c_double), _p(out, C.c_float))
    return out


def rotation_of_affine(csxy):
    a = np.ascontiguousarray(csxy, np.float32); out = np.zeros(9, np.float32)
    lib().ref_rotation_of_affine(_p(a, C.c_float), _p(out, C.c_float))
    return out.reshape(3, 3)


def filter_scan(raw4, n_az, n_bins, min_range, max_range, min_intensity, beam_thr, tf12=None):
    """RadarPreprocessor::filterScan -> (kept points [n, 4] in the base frame, polar [n, 2], max detections [m, 3])"""
    raw = np.ascontiguousarray(raw4, np.float32).reshape(-1, 4)
    assert len(raw) == n_az * n_bins
    tf = np.ascontiguousarray(np.eye(4)[:3] if tf12 is None else tf12, np.float32).reshape(12)
    cap = len(raw)
    out = np.zeros((cap, 4), np.float32); polar = np.zeros((cap, 2)); md = np.zeros((n_az + 1, 3)); nmd = C.c_uint32(0)
    n = lib().ref_filter_scan(_p(raw, C.c_float), int(n_az), int(n_bins), float(min_range), float(max_range), float(min_intensity), float(beam_thr),
                              _p(tf, C.c_float), _p(out, C.c_float), cap, _p(polar, C.c_double), _p(md, C.c_double), C.byref(nmd))
    if n < 0:
        raise RuntimeError("the reference threw in filterScan")
    return out[:n].copy(), polar[:n].copy(), md[:nmd.value].copy()
