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


def test_header_is_plain_c_and_every_symbol_links_from_c(built_lib, tmp_path):
    """the drop-in boundary is a C ABI: include/randt_gpu.h compiles as C99 (-pedantic), and a C translation unit that takes the address
    of every declared entry point links against librandt_gpu.so"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    syms = declared_symbols()
    src = tmp_path / "abi.c"
    src.write_text('#include "randt_gpu.h"\n#include <stdio.h>\nint main(void) {\n  const void* f[] = {%s};\n'
                   '  printf("%%d %%d\\n", (int)(sizeof(f) / sizeof(f[0])), randt_version());\n  return 0;\n}\n'
                   % ", ".join("(const void*)%s" % s_ for s_ in syms))
    exe = tmp_path / "abi"
    lib_dir = os.path.dirname(built_lib)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", lib_dir, "-lrandt_gpu", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == [str(len(syms)), "100"]


def test_host_layer_compiles_against_a_ceres_cost_function(tmp_path):
    """include/randt_host.hpp derives NdtCostFunction from <ceres/cost_function.h> when that header exists (it does not in this image):
    the branch is compiled here against a header with Ceres 2.1.0's CostFunction interface, so Evaluate's signature, the protected
    setters and the deleted copy are all what a real ceres::Problem expects"""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    fake = os.path.join(ROOT, "tests", "fake_ceres")
    host = os.path.join(ROOT, "randt_slam_b200", "host", "randt_host.cpp")
    r = subprocess.run([gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", fake, host], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    probe = tmp_path / "probe.cpp"
    probe.write_text('#include "randt_host.hpp"\n#include <type_traits>\n'
                     'static_assert(std::is_base_of<ceres::CostFunction, randt::NdtCostFunction>::value, "derives from ceres");\n'
                     'static_assert(!std::is_abstract<randt::NdtCostFunction>::value, "Evaluate is overridden");\n'
                     'static_assert(!std::is_copy_constructible<randt::NdtCostFunction>::value, "ceres forbids copies");\nint main() { return 0; }\n')
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", fake, "-I", os.path.join(ROOT, "include"), str(probe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_example_links_against_the_exported_host_classes(built_lib, tmp_path):
    """examples/local_fuser_flow.cpp (the reference's per-scan flow on include/randt_host.hpp) compiles, links against librandt_host.so /
    librandt_gpu.so, and — without a GPU — fails loudly instead of falling back to anything"""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    lib_dir = os.path.dirname(built_lib)
    exe = tmp_path / "flow"
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "local_fuser_flow.cpp"),
                        "-L", lib_dir, "-lrandt_host", "-lrandt_gpu", "-Wl,-rpath," + lib_dir, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert out.returncode == 0 and "cells exported" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 1 and "randt error" in out.stderr and "CUDA" in out.stderr
