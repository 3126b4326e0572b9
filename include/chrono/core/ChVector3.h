// Minimal stand-in for chrono::ChVector3<T> (reference: src/chrono/core/ChVector3.h).  Only the value-type surface that
// leaks through the Chrono::Dem API (ChSystemDem.h:20-23) is provided; real Chrono (which needs Eigen3) is not required.
#ifndef CHRONO_B200_CHVECTOR3_H
#define CHRONO_B200_CHVECTOR3_H
#include <cmath>
#include "chrono/core/ChTypes.h"
#include "chrono/utils/ChConstants.h"

namespace chrono {

template <class Real = double>
class ChVector3 {
  public:
    ChVector3() : m_data{0, 0, 0} {}
    ChVector3(Real x, Real y, Real z) : m_data{x, y, z} {}
    explicit ChVector3(Real a) : m_data{a, a, a} {}
    template <class R2>
    ChVector3(const ChVector3<R2>& o) : m_data{(Real)o.x(), (Real)o.y(), (Real)o.z()} {}

    Real& x() { return m_data[0]; }
    Real& y() { return m_data[1]; }
    Real& z() { return m_data[2]; }
    const Real& x() const { return m_data[0]; }
    const Real& y() const { return m_data[1]; }
    const Real& z() const { return m_data[2]; }
    Real& operator[](unsigned i) { return m_data[i]; }
    const Real& operator[](unsigned i) const { return m_data[i]; }
    const Real* data() const { return m_data; }
    void Set(Real x, Real y, Real z) { m_data[0] = x; m_data[1] = y; m_data[2] = z; }

    ChVector3 operator+(const ChVector3& o) const { return ChVector3(x() + o.x(), y() + o.y(), z() + o.z()); }
    ChVector3 operator-(const ChVector3& o) const { return ChVector3(x() - o.x(), y() - o.y(), z() - o.z()); }
    ChVector3 operator-() const { return ChVector3(-x(), -y(), -z()); }
    ChVector3 operator*(Real s) const { return ChVector3(x() * s, y() * s, z() * s); }
    ChVector3 operator/(Real s) const { return ChVector3(x() / s, y() / s, z() / s); }
    ChVector3& operator+=(const ChVector3& o) { m_data[0] += o.x(); m_data[1] += o.y(); m_data[2] += o.z(); return *this; }
    ChVector3& operator-=(const ChVector3& o) { m_data[0] -= o.x(); m_data[1] -= o.y(); m_data[2] -= o.z(); return *this; }
    ChVector3& operator*=(Real s) { m_data[0] *= s; m_data[1] *= s; m_data[2] *= s; return *this; }
    Real Dot(const ChVector3& o) const { return x() * o.x() + y() * o.y() + z() * o.z(); }
    Real operator^(const ChVector3& o) const { return Dot(o); }
    ChVector3 Cross(const ChVector3& o) const {
        return ChVector3(y() * o.z() - z() * o.y(), z() * o.x() - x() * o.z(), x() * o.y() - y() * o.x());
    }
    ChVector3 operator%(const ChVector3& o) const { return Cross(o); }
    Real Length2() const { return Dot(*this); }
    Real Length() const { return std::sqrt(Length2()); }
    ChVector3 GetNormalized() const { Real l = Length(); return l > 0 ? (*this) / l : *this; }
    bool Normalize() { Real l = Length(); if (l > 0) { *this = (*this) / l; return true; } return false; }

  private:
    Real m_data[3];
};

template <class Real>
ChVector3<Real> operator*(Real s, const ChVector3<Real>& v) { return v * s; }
template <class Real>
Real Vdot(const ChVector3<Real>& a, const ChVector3<Real>& b) { return a.Dot(b); }
template <class Real>
ChVector3<Real> Vcross(const ChVector3<Real>& a, const ChVector3<Real>& b) { return a.Cross(b); }

typedef ChVector3<double> ChVector3d;
typedef ChVector3<float> ChVector3f;
typedef ChVector3<int> ChVector3i;

}  // namespace chrono
#endif
