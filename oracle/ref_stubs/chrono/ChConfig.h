// Stand-in for the CMake-generated chrono/ChConfig.h (reference: src/chrono/ChConfig.h.in is
// produced by the reference's build system, which is not run here).  Only the switches that the
// Chrono::Multicore collision/math sources test are defined: fp64 "real" and OpenMP Thrust.
// No SIMD (CHRONO_HAS_SSE/AVX undefined) -> simd_non.h scalar code paths.
#pragma once
#define USE_COLLISION_DOUBLE
#define CHRONO_MULTICORE_USE_DOUBLE
#define CHRONO_OPENMP_ENABLED
#ifndef THRUST_DEVICE_SYSTEM
#define THRUST_DEVICE_SYSTEM THRUST_DEVICE_SYSTEM_OMP
#endif
