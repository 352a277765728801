"""In-tree build of the two shared libraries (and, for tests, the CPU shim of the device header).

    python -m rendering_b200.build [--force]

librtb_host.so : g++ -std=c++17 -O2 -ffp-contract=off   (no -march=native: FMA contraction changes
                 loader / tree-builder results)
librtb_cuda.so : nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")

HOST_SOURCES = ["host/util.cpp", "host/objects.cpp", "host/scene.cpp", "host/flatten.cpp", "host/host_abi.cpp", "host/render.cpp"]
CUDA_SOURCES = ["cuda/rtb_api.cu"]
CUDA_DEPS = ["cuda/rtb_kernels.cuh", "cuda/rtb_tile.cuh", "cuda/rt_device.cuh", "cuda/scene_pack.h", "cuda/bvh_build.h", "cuda/lbvh_build.cuh"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def _glob_headers(sub):
    d = os.path.join(CSRC, sub)
    return [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".cuh"))] + [os.path.join(ROOT, "include", "rtb.h")]


def build_host(force=False):
    out = os.path.join(HERE, "librtb_host.so")
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    if force or _stale(out, srcs + _glob_headers("host")):
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra", *srcs, "-o", out, "-ldl"])
    exe = os.path.join(HERE, "RayTracing")
    main = os.path.join(CSRC, "host/main.cpp")
    if force or _stale(exe, [main, out]):
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", main, "-o", exe, "-L" + HERE, "-lrtb_host", "-Wl,-rpath,$ORIGIN"])
    return out


def build_cuda(force=False):
    out = os.path.join(HERE, "librtb_cuda.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = [os.path.join(CSRC, s) for s in CUDA_DEPS] + [os.path.join(ROOT, "include", "rtb.h")]
    extra = os.environ.get("RTB_NVCC_EXTRA", "").split()      # experiment knobs, e.g. -DRTB_CHUNK=64
    if force or extra or _stale(out, srcs + deps):
        _run(["nvcc", *NVCC_FLAGS, *extra, *srcs, "-o", out])
    return out


def build_shim(force=False):
    """tests/shim: the device arithmetic header compiled for the CPU (test-only)."""
    src = os.path.join(ROOT, "tests", "shim", "shim_render.cpp")
    out = os.path.join(ROOT, "tests", "shim", "libshim.so")
    deps = [src] + [os.path.join(CSRC, s) for s in ("cuda/rt_device.cuh", "cuda/scene_pack.h")]
    if force or _stale(out, deps):
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", out])
    return out


def build_all(force=False):
    return [build_host(force), build_cuda(force), build_shim(force)]


if __name__ == "__main__":
    build_all("--force" in sys.argv)
