// Stand-in that shadows chrono/core/ChMatrix33.h (see ChVector3.h stand-in).
#pragma once
namespace chrono {
template <class Real = double>
class ChMatrix33 {
  public:
    ChMatrix33() : m{} {}
    Real& operator()(int r, int c) { return m[r][c]; }
    const Real& operator()(int r, int c) const { return m[r][c]; }
  private:
    Real m[3][3];
};
}  // namespace chrono
