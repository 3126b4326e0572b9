// Force-included ahead of the function_CalcContactForces excerpt when oracle/Makefile pipes that
// function (and only that function) from the reference source
//   src/chrono_multicore/solver/ChIterativeSolverMulticoreSMC.cpp
// straight into the compiler.  The rest of that file needs Eigen (ChDataManager sparse matrices),
// so it cannot be compiled here; the excerpt needs only:
//   - the multicore_math types (compiled from the reference as-is),
//   - the ChSystemSMC enumerations, src/chrono/physics/ChSystemSMC.h:34-53 (same order/values),
//   - max_shear, src/chrono_multicore/ChDataManager.h:185.
#pragma once
#include "prelude.h"
#include <limits>
#include <stdexcept>
#include "chrono/multicore_math/types.h"
#include "chrono/multicore_math/utility.h"
// Same headers the original translation unit pulls in (ChIterativeSolverMulticoreSMC.cpp:31-41).  They matter:
// the function calls an UNQUALIFIED abs() on a double (:173); only with these headers in scope does it resolve
// to the floating-point overload (without them it binds to ::abs(int) and truncates |v_n| to 0).
#include <algorithm>
#include "chrono/multicore_math/thrust.h"
#include <thrust/sort.h>
namespace chrono {
class ChSystemSMC {
  public:
    enum ContactForceModel { Hooke, Hertz, PlainCoulomb, Flores };
    enum AdhesionForceModel { Constant, DMT, Perko };
    enum TangentialDisplacementModel { None, OneStep, MultiStep };
};
}  // namespace chrono
#define max_shear 20
using namespace chrono;
