"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every entry point that
include/chrono_b200_dem.h declares, and fails loudly (DEMB200_ECUDA, no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "chrono_b200_dem.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dem_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from chrono_b200.build import build_library
    lib = C.CDLL(build_library())
    names = declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return  # this is the CPU-box check
    from chrono_b200 import dem
    try:
        dem.DemSystem(dem.config())
    except dem.DemError as e:
        assert e.code == -1  # DEMB200_ECUDA
    else:
        raise AssertionError("engine was created without a CUDA device")


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under chrono_b200/ may reference it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "chrono_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|dem_oracle|pyoracle|libdem_oracle", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad
