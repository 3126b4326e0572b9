// Stand-in for the CMake-generated chrono/ChConfig.h (src/chrono/ChConfig.h.in; the reference's build system is not run
// here).  The incumbent Chrono::Dem device code only needs the GPU back-end switch.
#pragma once
#define CHRONO_HAS_CUDA
#define CHRONO_CUDA_VERSION "12.9"
