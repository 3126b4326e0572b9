// Minimal stand-in for chrono::ChMatrix33<T> (reference: src/chrono/core/ChMatrix33.h; Eigen-free).
#ifndef CHRONO_B200_CHMATRIX33_H
#define CHRONO_B200_CHMATRIX33_H
#include "chrono/core/ChQuaternion.h"
#include "chrono/core/ChVector3.h"

namespace chrono {

template <class Real = double>
class ChMatrix33 {
  public:
    ChMatrix33() { SetIdentity(); }
    explicit ChMatrix33(Real diag) { SetIdentity(); m[0][0] = m[1][1] = m[2][2] = diag; }
    template <class R2>
    explicit ChMatrix33(const ChQuaternion<R2>& q) { SetFromQuaternion(q); }
    /// Diagonal matrix (scaling), as ChMatrix33(const ChVector3d&) of the reference.
    template <class R2>
    explicit ChMatrix33(const ChVector3<R2>& diag) { SetIdentity(); m[0][0] = (Real)diag.x(); m[1][1] = (Real)diag.y(); m[2][2] = (Real)diag.z(); }
    void SetIdentity() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = (i == j) ? (Real)1 : (Real)0; }
    Real& operator()(int i, int j) { return m[i][j]; }
    const Real& operator()(int i, int j) const { return m[i][j]; }
    template <class R2>
    void SetFromQuaternion(const ChQuaternion<R2>& q) {
        Real e0 = (Real)q.e0(), e1 = (Real)q.e1(), e2 = (Real)q.e2(), e3 = (Real)q.e3();
        m[0][0] = (e0 * e0 + e1 * e1) * 2 - 1; m[0][1] = (e1 * e2 - e0 * e3) * 2;     m[0][2] = (e1 * e3 + e0 * e2) * 2;
        m[1][0] = (e1 * e2 + e0 * e3) * 2;     m[1][1] = (e0 * e0 + e2 * e2) * 2 - 1; m[1][2] = (e2 * e3 - e0 * e1) * 2;
        m[2][0] = (e1 * e3 - e0 * e2) * 2;     m[2][1] = (e2 * e3 + e0 * e1) * 2;     m[2][2] = (e0 * e0 + e3 * e3) * 2 - 1;
    }
    ChVector3<Real> operator*(const ChVector3<Real>& v) const {
        return ChVector3<Real>(m[0][0] * v.x() + m[0][1] * v.y() + m[0][2] * v.z(), m[1][0] * v.x() + m[1][1] * v.y() + m[1][2] * v.z(),
                               m[2][0] * v.x() + m[2][1] * v.y() + m[2][2] * v.z());
    }

  private:
    Real m[3][3];
};

}  // namespace chrono
#endif
