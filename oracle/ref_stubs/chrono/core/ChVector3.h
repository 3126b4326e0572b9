// Stand-in that shadows chrono/core/ChVector3.h (the real one needs Eigen3, absent here).
// Only the members that chrono/multicore_math/utility.h's conversion helpers touch.
#pragma once
#include "chrono/utils/ChConstants.h"
namespace chrono {
template <class Real = double>
class ChVector3 {
  public:
    ChVector3() : m{0, 0, 0} {}
    ChVector3(Real a, Real b, Real c) : m{a, b, c} {}
    Real x() const { return m[0]; }
    Real y() const { return m[1]; }
    Real z() const { return m[2]; }
  private:
    Real m[3];
};
}  // namespace chrono
