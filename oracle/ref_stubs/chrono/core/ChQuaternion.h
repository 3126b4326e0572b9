// Stand-in that shadows chrono/core/ChQuaternion.h (see ChVector3.h stand-in).
#pragma once
namespace chrono {
template <class Real = double>
class ChQuaternion {
  public:
    ChQuaternion() : m{1, 0, 0, 0} {}
    ChQuaternion(Real a, Real b, Real c, Real d) : m{a, b, c, d} {}
    Real e0() const { return m[0]; }
    Real e1() const { return m[1]; }
    Real e2() const { return m[2]; }
    Real e3() const { return m[3]; }
  private:
    Real m[4];
};
}  // namespace chrono
