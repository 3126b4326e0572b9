// =============================================================================
// oracle/omath.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// fp64 vector / quaternion arithmetic restating the *scalar* code path of the reference's
// chrono/multicore_math (the path selected when neither CHRONO_HAS_SSE nor CHRONO_HAS_AVX is
// defined).  Operation order is kept identical so results are bit-identical to the reference
// objects compiled into oracle/_ref (checked by tests/test_oracle_vs_ref.py):
//   - Dot / Cross / componentwise ops ........ src/chrono/multicore_math/simd_non.h:20-75
//   - Length, Normalize ...................... src/chrono/multicore_math/real3.cpp:110-118
//   - Rotate, RotateT, Mult, AbsRotate ....... src/chrono/multicore_math/real4.cpp:139-187
//   - TransformLocalToParent/ParentToLocal ... src/chrono/multicore_math/utility.h:46-55
// Compile with -ffp-contract=off: the reference expressions are plain mul/add sequences.
// =============================================================================
#pragma once
#include <cmath>
#include <algorithm>

namespace orc {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a) : x(a), y(a), z(a) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return (&x)[i]; }
    double& operator[](int i) { return (&x)[i]; }
};

struct Q4 {  // w + xi + yj + zk  (reference "quaternion": members w,x,y,z)
    double w, x, y, z;
    Q4() : w(1), x(0), y(0), z(0) {}
    Q4(double a, double b, double c, double d) : w(a), x(b), y(c), z(d) {}
    V3 vect() const { return V3(x, y, z); }
};

inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(const V3& a, const V3& b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(const V3& a, const V3& b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator+(const V3& a, double b) { return V3(a.x + b, a.y + b, a.z + b); }
inline V3 operator-(const V3& a, double b) { return V3(a.x - b, a.y - b, a.z - b); }
inline V3 operator*(const V3& a, double b) { return V3(a.x * b, a.y * b, a.z * b); }
inline V3 operator/(const V3& a, double b) { return V3(a.x / b, a.y / b, a.z / b); }
inline V3 operator*(double a, const V3& b) { return V3(a * b.x, a * b.y, a * b.z); }
inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
inline V3& operator+=(V3& a, const V3& b) { a = a + b; return a; }
inline V3& operator-=(V3& a, const V3& b) { a = a - b; return a; }
inline V3& operator*=(V3& a, double b) { a = a * b; return a; }

// simd_non.h:56-62
inline double Dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double Dot(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
// simd_non.h:68-74
inline V3 Cross(const V3& a, const V3& b) {
    return V3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
// real3.cpp:115
inline double Length(const V3& v) { return std::sqrt(Dot(v)); }
inline V3 Min(const V3& a, const V3& b) { return V3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
inline V3 Max(const V3& a, const V3& b) { return V3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }
inline V3 Abs(const V3& a) { return V3(std::abs(a.x), std::abs(a.y), std::abs(a.z)); }

// real4.cpp:124 (operator~ = conjugate)
inline Q4 Conj(const Q4& q) { return Q4(q.w, -q.x, -q.y, -q.z); }
// real4.cpp:141-152 (non-AVX2 branch)
inline Q4 Mult(const Q4& a, const Q4& b) {
    Q4 t;
    t.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    t.x = a.w * b.x + a.x * b.w - a.z * b.y + a.y * b.z;
    t.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    t.z = a.w * b.z + a.z * b.w - a.y * b.x + a.x * b.y;
    return t;
}
// real4.cpp:158-161
inline V3 Rotate(const V3& v, const Q4& q) {
    V3 t = 2 * Cross(q.vect(), v);
    return v + q.w * t + Cross(q.vect(), t);
}
// real4.cpp:163-165
inline V3 RotateT(const V3& v, const Q4& q) { return Rotate(v, Conj(q)); }
// real4.cpp:168-187
inline V3 AbsRotate(const Q4& q, const V3& v) {
    double e0e0 = q.w * q.w, e1e1 = q.x * q.x, e2e2 = q.y * q.y, e3e3 = q.z * q.z;
    double e0e1 = q.w * q.x, e0e2 = q.w * q.y, e0e3 = q.w * q.z;
    double e1e2 = q.x * q.y, e1e3 = q.x * q.z, e2e3 = q.y * q.z;
    V3 r;
    r.x = std::abs((e0e0 + e1e1) * 2 - 1) * v.x + std::abs((e1e2 - e0e3) * 2) * v.y + std::abs((e1e3 + e0e2) * 2) * v.z;
    r.y = std::abs((e1e2 + e0e3) * 2) * v.x + std::abs((e0e0 + e2e2) * 2 - 1) * v.y + std::abs((e2e3 - e0e1) * 2) * v.z;
    r.z = std::abs((e1e3 - e0e2) * 2) * v.x + std::abs((e2e3 + e0e1) * 2) * v.y + std::abs((e0e0 + e3e3) * 2 - 1) * v.z;
    return r;
}
// utility.h:46-55
inline V3 TransformLocalToParent(const V3& p, const Q4& q, const V3& rl) { return p + Rotate(rl, q); }
inline V3 TransformParentToLocal(const V3& p, const Q4& q, const V3& rp) { return RotateT(rp - p, q); }

}  // namespace orc
