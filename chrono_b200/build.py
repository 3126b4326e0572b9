"""In-tree build of the CUDA engine for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("DEMB200_LIB") or os.path.join(HERE, "libchrono_b200_dem.so")  # DEMB200_LIB: tuning variants
SOURCES = [os.path.join(HERE, "csrc", "dem_engine.cu"), os.path.join(HERE, "csrc", "ChSystemDem.cpp")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", f) for f in ("dem_kernels.cuh", "dem_types.h")] + [
    os.path.join(ROOT, "include", "chrono_b200_dem.h"), os.path.join(ROOT, "include", "chrono_dem", "physics", "ChSystemDem.h"),
    os.path.join(ROOT, "include", "chrono_dem", "ChDemDefines.h"),
    os.path.join(ROOT, "include", "chrono", "geometry", "ChTriangleMeshConnected.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-I", os.path.join(ROOT, "include"), "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("DEMB200_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    subprocess.check_call(cmd)
    return LIB_PATH
