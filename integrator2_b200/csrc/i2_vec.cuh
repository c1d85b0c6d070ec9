// Small FP64 vector helpers shared by the sm_100a kernels.  Compiles as host code too
// (tests/host_emu builds the per-point functions with g++ to pre-check numerics without a GPU).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define I2_HD __host__ __device__ __forceinline__
#define I2_D __device__ __forceinline__
#else
#define I2_HD inline
#define I2_D inline
#endif

// Warp-uniform control flow: a rare, data-dependent alternative is evaluated by the WHOLE warp when any lane needs it
// and selected per lane afterwards; on the host (tests/host_emu) the vote degenerates to the lane's own predicate.
// Callers on the device must have all 32 lanes of the warp active.
#if defined(__CUDA_ARCH__)
#define I2_WARP_ANY(pred) (__any_sync(0xffffffffu, (pred)) != 0)
#else
#define I2_WARP_ANY(pred) (pred)
#endif

namespace i2 {

// thresholds and constants, bit-identical to the reference (src/common/constants.h:11-65):
// they select branches, so they are part of the numerics.
constexpr double DOUBLE_MIN = 2e-6;
constexpr double EPS_ZERO = 1e-6;
constexpr double EPS_ZERO2 = 1e-10;
constexpr double EPS_PSI_THETA2 = EPS_ZERO * EPS_ZERO;
constexpr double EPS_INTEGRATION = 1e-5;
constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 6.28318530717958647692;
constexpr double RECIPROCAL_FOUR_PI = 0.079577471545947667884;
constexpr int MAX_REFINE_LEVEL = 5;
constexpr int MAX_GAUSS_POINTS = 13;

struct d3 { double x, y, z; };
struct d4 { double x, y, z, w; };

I2_HD d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
I2_HD d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
I2_HD d3 operator-(d3 a) { return {-a.x, -a.y, -a.z}; }
I2_HD d3 operator*(double s, d3 a) { return {a.x * s, a.y * s, a.z * s}; }
I2_HD double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
I2_HD d3 cross(d3 a, d3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
I2_HD double norm2(d3 a) { return dot(a, a); }
I2_HD double norm(d3 a) { return sqrt(dot(a, a)); }
I2_HD d3 over(d3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
I2_HD d3 unit(d3 a) { const double inv = 1.0 / norm(a); return {a.x * inv, a.y * inv, a.z * inv}; }
I2_HD double sq(double x) { return x * x; }

I2_HD d4 operator+(d4 a, d4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
I2_HD d4 operator-(d4 a, d4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
I2_HD d4 operator*(double s, d4 a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
I2_HD d4 vec4(d3 a, double w = 0.0) { return {a.x, a.y, a.z, w}; }
I2_HD double l1(d3 a) { return fabs(a.x) + fabs(a.y) + fabs(a.z); }
I2_HD double l1(d4 a) { return fabs(a.x) + fabs(a.y) + fabs(a.z) + fabs(a.w); }

// sign()/arg() with the reference's 2e-6 dead zone (src/common/cuda_math.cuh:31-49)
I2_HD double sgn_dz(double x) { return fabs(x) < DOUBLE_MIN ? 0.0 : (x > DOUBLE_MIN ? 1.0 : -1.0); }
I2_HD double arg_dz(double x) { return x > DOUBLE_MIN ? 0.0 : PI; }

// angle between two vectors (src/common/cuda_math.cu:14-27): 0 for degenerate input, cosine clamped to [-1, 1]
// (acos(1) = 0 and acos(-1) = pi exactly, which are the reference's early-return values) — branch-free
I2_HD double angle_between(d3 a, d3 b) {
    const double den = sqrt(norm2(a) * norm2(b));
    const double c = fmin(fmax(dot(a, b) / den, -1.0), 1.0);
    const double r = acos(c);
    return den < EPS_ZERO ? 0.0 : r;
}

struct tri3 { int a, b, c; };
I2_HD int tri_at(tri3 t, int k) { return k == 0 ? t.a : (k == 1 ? t.b : t.c); }
// cyclic rotation so that vertex #shift comes first (src/evaluators/evaluatorJ3DK.cu:827-847)
I2_HD tri3 rot_left(tri3 t, int shift) {
    if (shift == 1) return {t.b, t.c, t.a};
    if (shift == 2) return {t.c, t.a, t.b};
    return t;
}

}  // namespace i2
