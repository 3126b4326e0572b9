// Stand-in that shadows chrono/core/ChVector2.h (see ChVector3.h stand-in).
#pragma once
namespace chrono {
template <class Real = double>
class ChVector2 {
  public:
    ChVector2() : m{0, 0} {}
    ChVector2(Real a, Real b) : m{a, b} {}
    Real x() const { return m[0]; }
    Real y() const { return m[1]; }
  private:
    Real m[2];
};
}  // namespace chrono
