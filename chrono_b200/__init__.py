"""chrono_b200 -- B200-native smooth-contact (SMC) granular DEM step, drop-in for the hot path of
Chrono::Dem (ChSystemDem / ChSystemGpu).  The product is the C-ABI shared library
``libchrono_b200_dem.so`` (hand-written sm_100a CUDA, see csrc/); this package only holds its
ctypes binding, the build recipe and synthetic scene generators.  There is no CPU fallback."""
from .build import build_library, LIB_PATH  # noqa: F401
