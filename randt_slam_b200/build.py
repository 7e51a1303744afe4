"""Builds the CUDA extension in-tree: randt_slam_b200/librandt_gpu.so (sm_100a only).

nvcc cross-compiles without a GPU.  K1/K2 are compiled with -fmad=false: their float32 arithmetic reproduces the reference's
un-fused multiply/add sequences bit for bit; K3 (fp64) keeps FMA contraction.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "librandt_gpu.so")
HOST_LIB = os.path.join(HERE, "librandt_host.so")       # C++ mirror of the reference's Map / Matcher surface (host/randt_host.cpp)
HOST_SRC = os.path.join(HERE, "host", "randt_host.cpp")
HOST_SRCS = [HOST_SRC, os.path.join(HERE, "host", "window_solver.cpp")]   # + the joint window problem of estimateTransformCeres
HOST_DEPS = HOST_SRCS + [os.path.join(HERE, "host", "window_solver.hpp")]
HOST_HDR = os.path.join(HERE, "..", "include", "randt_host.hpp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("RANDT_NVCC_FLAGS", "").split()
COMMON = EXTRA + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
UNITS = {
    "k3_pair_eval.cu": [],
    "k4_lm_step.cu": [],
    "k7_solve.cu": [],
    "k8_allpairs.cu": [],
    "k2_associate.cu": ["-fmad=false"],
    "k1_voxelize.cu": ["-fmad=false"],
    "k5_cs_divergence.cu": ["-fmad=false"],
    "k6_filter_scan.cu": ["-fmad=false"],
    "capi.cu": [],
}


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_all(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "schedule.hpp"), os.path.join(CSRC, "k3_device.cuh"), os.path.join(CSRC, "k4_device.cuh"), os.path.join(HERE, "..", "include", "randt_gpu.h"),
               os.path.abspath(__file__)]
    jobs = []
    for src, extra in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + headers):
            jobs.append([nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, src.replace(".cu", ".o")) for src in UNITS]
    if force or jobs or _newer(LIB, objs):
        run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    # host layer: plain C++17 over the C-ABI (no CUDA headers), linked against librandt_gpu.so next to it
    if force or _newer(HOST_LIB, HOST_DEPS + [HOST_HDR, LIB, os.path.join(CSRC, "schedule.hpp")]):
        cxx = os.environ.get("CXX", "g++")
        # RANDT_HOST_CXXFLAGS: e.g. "-I/usr/include/eigen3 -I<ceres>/include" so that the host layer is compiled against the real
        # ceres/cost_function.h of the tree it is going to be linked into
        run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall"] + os.environ.get("RANDT_HOST_CXXFLAGS", "").split() +
            ["-o", HOST_LIB] + HOST_SRCS + [
             "-L" + HERE, "-lrandt_gpu", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
