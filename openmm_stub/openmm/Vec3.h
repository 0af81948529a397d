#ifndef OPENMM_STUB_VEC3_H_
#define OPENMM_STUB_VEC3_H_
// Stand-in for OpenMM's Vec3 (three doubles).  Division multiplies by the reciprocal, as
// OpenMM's does, so that results of code compiled against this stub match a real build.
#include <cassert>
#include <cmath>
#include <iosfwd>

namespace OpenMM {

class Vec3 {
public:
    Vec3() : v{0.0, 0.0, 0.0} {}
    Vec3(double x, double y, double z) : v{x, y, z} {}
    double operator[](int i) const { assert(i >= 0 && i < 3); return v[i]; }
    double& operator[](int i) { assert(i >= 0 && i < 3); return v[i]; }
    bool operator==(const Vec3& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    Vec3 operator+() const { return *this; }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
    Vec3& operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    Vec3 operator/(double s) const { double r = 1.0 / s; return Vec3(v[0] * r, v[1] * r, v[2] * r); }
    Vec3& operator/=(double s) { double r = 1.0 / s; v[0] *= r; v[1] *= r; v[2] *= r; return *this; }
    double dot(const Vec3& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
private:
    double v[3];
};

inline Vec3 operator*(double s, const Vec3& a) { return a * s; }

} // namespace OpenMM
#endif
