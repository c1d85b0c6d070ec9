// Point3 (= double3) and the small vector algebra the reference exposes to its users
// (/root/reference/src/common/cuda_math.cuh).  Host-side convenience only: the kernels have their own
// helpers (csrc/i2_vec.cuh).
#ifndef CUDA_MATH_CUH
#define CUDA_MATH_CUH

#include <cuda_runtime.h>
#include <cmath>
#include "constants.h"

typedef double3 Point3;

#define I2_FN __host__ __device__ inline
I2_FN double sqr(double x) { return x * x; }
I2_FN double sign(double x) { return fabs(x) < CONSTANTS::DOUBLE_MIN ? 0.0 : (x > CONSTANTS::DOUBLE_MIN ? 1.0 : -1.0); }
I2_FN double arg(double x) { return x > CONSTANTS::DOUBLE_MIN ? 0.0 : CONSTANTS::PI; }
I2_FN double4 assign_vector_part(const Point3 &v) { return make_double4(v.x, v.y, v.z, 0); }
I2_FN Point3 extract_vector_part(const double4 &v) { return make_double3(v.x, v.y, v.z); }
I2_FN double4 operator+(const double4 &a, const double4 &b) { return make_double4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
I2_FN double4 operator-(const double4 &a, const double4 &b) { return make_double4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
I2_FN double4 operator*(double s, const double4 &v) { return make_double4(v.x * s, v.y * s, v.z * s, v.w * s); }
I2_FN void operator+=(double4 &v, const double4 &a) { v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
I2_FN Point3 operator+(const Point3 &a, const Point3 &b) { return make_double3(a.x + b.x, a.y + b.y, a.z + b.z); }
I2_FN Point3 operator-(const Point3 &a, const Point3 &b) { return make_double3(a.x - b.x, a.y - b.y, a.z - b.z); }
I2_FN Point3 operator-(const Point3 &a) { return make_double3(-a.x, -a.y, -a.z); }
I2_FN Point3 operator*(double s, const Point3 &v) { return make_double3(v.x * s, v.y * s, v.z * s); }
I2_FN void operator+=(Point3 &v, const Point3 &a) { v.x += a.x; v.y += a.y; v.z += a.z; }
I2_FN void operator*=(Point3 &v, const double &s) { v.x *= s; v.y *= s; v.z *= s; }
I2_FN void operator/=(Point3 &v, const double &s) { v.x /= s; v.y /= s; v.z /= s; }
I2_FN double dot(const Point3 &a, const Point3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
I2_FN Point3 cross(const Point3 &a, const Point3 &b) { return make_double3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
I2_FN double vector_length2(const Point3 &v) { return dot(v, v); }
I2_FN double vector_length(const Point3 &v) { return sqrt(dot(v, v)); }
I2_FN Point3 normalize(const Point3 &v) { const double inv = 1.0 / vector_length(v); return make_double3(v.x * inv, v.y * inv, v.z * inv); }
I2_FN double norm1(const Point3 &v) { return fabs(v.x) + fabs(v.y) + fabs(v.z); }
I2_FN double norm1(const double4 &v) { return fabs(v.x) + fabs(v.y) + fabs(v.z) + fabs(v.w); }
I2_FN double4 divide(const double4 &n, const double4 &d) {
    double4 r;
    r.x = (fabs(n.x) < CONSTANTS::DOUBLE_MIN && fabs(d.x) < CONSTANTS::DOUBLE_MIN) ? 0.0 : n.x / d.x;
    r.y = (fabs(n.y) < CONSTANTS::DOUBLE_MIN && fabs(d.y) < CONSTANTS::DOUBLE_MIN) ? 0.0 : n.y / d.y;
    r.z = (fabs(n.z) < CONSTANTS::DOUBLE_MIN && fabs(d.z) < CONSTANTS::DOUBLE_MIN) ? 0.0 : n.z / d.z;
    r.w = (fabs(n.w) < CONSTANTS::EPS_ZERO2 && fabs(d.w) < CONSTANTS::EPS_ZERO2) ? 0.0 : n.w / d.w;
    return r;
}
I2_FN double angle(const Point3 &a, const Point3 &b) {
    const double den = sqrt(vector_length2(a) * vector_length2(b));
    if (den < CONSTANTS::EPS_ZERO) return 0;
    const double c = dot(a, b) / den;
    if (c >= 1.0) return 0;
    if (c <= -1.0) return CONSTANTS::PI;
    return acos(c);
}
#undef I2_FN

#endif  // CUDA_MATH_CUH
