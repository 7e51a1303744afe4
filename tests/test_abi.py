"""The C-ABI shared library loads without a GPU and exports every symbol include/randt_gpu.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from randt_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "randt_gpu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"RANDT_API\s+[\w\s\*]+?\b(randt_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from randt_slam_b200 import build
    path = build.build_all()
    assert os.path.exists(path)
    return path


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ("randt_ctx_create", "randt_voxelize", "randt_associate", "randt_problem_create", "randt_eval_emit", "randt_eval_fused",
                 "randt_sweep_costs", "randt_map_merge", "randt_map_transform", "randt_last_error"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, "declared in include/randt_gpu.h but not exported: %s" % missing


def test_python_binding_table_matches_header(built_lib):
    assert sorted(capi._SIGS) == declared_symbols()
    capi.lib()          # binds every signature; raises on a missing symbol


def test_every_entry_point_cites_the_reference_interface_it_replaces():
    """include/*.h must name the reference file:line each compute entry point stands in for"""
    src = open(HEADER).read()
    for anchor in ("grid.cpp:7-14", "radar_preprocessor.cpp:151-169", "ndt_map.cpp:238-245", "ndt_cell.cpp:25-114", "ndt_matcher.cpp:200-217",
                   "ndt_map.cpp:101-151", "ceres_residuals.h", "ndt_map.cpp:191-207", "ndt_map.cpp:177-182", "ndt_matcher.cpp:560-576", "ndt_map.cpp:42-99",
                   "ndt_matcher.cpp:457-492"):
        assert anchor in src, anchor


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(capi.Loss) == 40           # int32 + pad, 4 doubles
    assert ctypes.sizeof(capi.GridParams) == 40     # float, 4 int32, pad, 2 doubles
    assert capi.GridParams.resolution.offset == 24 and capi.GridParams.max_linf.offset == 32
    assert capi.Loss.scale.offset == 8


def test_version_and_error_paths_without_a_gpu(built_lib):
    L = capi.lib()
    assert L.randt_version() >= 100
    assert L.randt_last_error(None) == b"null context"
    import torch
    if not torch.cuda.is_available():
        # the product path fails loudly without a device: no CPU fallback
        with pytest.raises(capi.RandtError):
            capi.Context(0)


def test_record_layout_constants_match_the_header():
    """RANDT_FUSED_* / RANDT_PACKED_* / RANDT_REG_* enumerators in include/randt_gpu.h == the constants the Python bindings index with"""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(RANDT_(?:FUSED|PACKED|REG)_[A-Z_]+)\s*=\s*(\d+)", src)}
    assert enums["RANDT_FUSED_STRIDE"] == capi.FUSED_STRIDE == 24
    assert (enums["RANDT_FUSED_H"], enums["RANDT_FUSED_G"], enums["RANDT_FUSED_COST"], enums["RANDT_FUSED_MAXR"], enums["RANDT_FUSED_SUMSQ"],
            enums["RANDT_FUSED_N"]) == (capi.FUSED_H, capi.FUSED_G, capi.FUSED_COST, capi.FUSED_MAXR, capi.FUSED_SUMSQ, capi.FUSED_N)
    assert enums["RANDT_PACKED_STRIDE"] == capi.PACKED_STRIDE == 18
    # packed = upper triangle of the 4x4 H row by row, then everything after H
    want = [4 * r + c for r in range(4) for c in range(r, 4)] + list(range(16, 24))
    assert capi._PACKED_SRC == want
    assert (enums["RANDT_PACKED_G"], enums["RANDT_PACKED_COST"], enums["RANDT_PACKED_MAXR"], enums["RANDT_PACKED_SUMSQ"], enums["RANDT_PACKED_N"]) == \
        (want.index(16), want.index(20), want.index(21), want.index(22), want.index(23))
    import numpy as np
    full = np.arange(48, dtype=np.float64).reshape(2, 24)
    assert np.array_equal(capi.pack_fused(full)[1], full[1][want])
    if "RANDT_REG_STRIDE" in enums:
        assert enums["RANDT_REG_STRIDE"] == capi.REG_STRIDE
        assert enums["RANDT_REG_SCORE"] == capi.REG_SCORE and enums["RANDT_REG_STATUS"] == capi.REG_STATUS
