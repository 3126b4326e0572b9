// Minimal stand-in for chrono::ChQuaternion<T> (reference: src/chrono/core/ChQuaternion.h).
#ifndef CHRONO_B200_CHQUATERNION_H
#define CHRONO_B200_CHQUATERNION_H
#include <cmath>
#include "chrono/core/ChVector3.h"

namespace chrono {

template <class Real = double>
class ChQuaternion {
  public:
    ChQuaternion() : m_data{1, 0, 0, 0} {}
    ChQuaternion(Real e0, Real e1, Real e2, Real e3) : m_data{e0, e1, e2, e3} {}
    Real& e0() { return m_data[0]; }
    Real& e1() { return m_data[1]; }
    Real& e2() { return m_data[2]; }
    Real& e3() { return m_data[3]; }
    const Real& e0() const { return m_data[0]; }
    const Real& e1() const { return m_data[1]; }
    const Real& e2() const { return m_data[2]; }
    const Real& e3() const { return m_data[3]; }
    void SetFromAngleAxis(Real angle, const ChVector3<Real>& axis) {
        Real h = angle / 2, s = std::sin(h);
        ChVector3<Real> a = axis.GetNormalized();
        m_data[0] = std::cos(h); m_data[1] = a.x() * s; m_data[2] = a.y() * s; m_data[3] = a.z() * s;
    }
    void Normalize() {
        Real l = std::sqrt(m_data[0] * m_data[0] + m_data[1] * m_data[1] + m_data[2] * m_data[2] + m_data[3] * m_data[3]);
        if (l > 0) for (auto& v : m_data) v /= l;
    }
    /// Rotate a vector by this (unit) quaternion.
    ChVector3<Real> Rotate(const ChVector3<Real>& v) const {
        ChVector3<Real> u(m_data[1], m_data[2], m_data[3]);
        ChVector3<Real> t = u.Cross(v) * (Real)2;
        return v + t * m_data[0] + u.Cross(t);
    }
    ChQuaternion operator*(const ChQuaternion& o) const {
        return ChQuaternion(e0() * o.e0() - e1() * o.e1() - e2() * o.e2() - e3() * o.e3(),
                            e0() * o.e1() + e1() * o.e0() + e2() * o.e3() - e3() * o.e2(),
                            e0() * o.e2() - e1() * o.e3() + e2() * o.e0() + e3() * o.e1(),
                            e0() * o.e3() + e1() * o.e2() - e2() * o.e1() + e3() * o.e0());
    }

  private:
    Real m_data[4];
};
typedef ChQuaternion<double> ChQuaterniond;
typedef ChQuaternion<float> ChQuaternionf;
const ChQuaterniond QUNIT(1., 0., 0., 0.);

// src/chrono/core/ChRotation.h: rotations from angle-axis, double precision
inline ChQuaterniond QuatFromAngleAxis(double angle, const ChVector3<double>& axis) {
    ChQuaterniond q;
    q.SetFromAngleAxis(angle, axis);
    return q;
}
inline ChQuaterniond QuatFromAngleX(double a) { return QuatFromAngleAxis(a, ChVector3<double>(1, 0, 0)); }
inline ChQuaterniond QuatFromAngleY(double a) { return QuatFromAngleAxis(a, ChVector3<double>(0, 1, 0)); }
inline ChQuaterniond QuatFromAngleZ(double a) { return QuatFromAngleAxis(a, ChVector3<double>(0, 0, 1)); }

}  // namespace chrono
#endif
