// Branch-free FP64 primitives for the regular-pair kernel (the FP64 pipe is the roofline, so every DFMA counts):
//   fast_sqrt      MUFU.RSQ64H seed + one cubic (Halley-type) step                              (5 FP64 ops)
//   fast_rcp       MUFU.RCP64H seed + one cubic step                                            (3 FP64 ops)
//   log_ratio      ln(N/D) with the division folded into the atanh argument (N-D)/(N+D)        (~25 FP64 ops)
//   atan2_fast     atan2(y,x) with a 5-entry argument reduction, one division                  (~26 FP64 ops)
// libdevice spends ~17 (sqrt) / ~19 (div) / ~30 (log) / ~46 (atan2) FP64-pipe instructions on the same jobs and
// branches into slow paths; these versions assume finite, non-denormal inputs (triangle geometry) and have no
// data-dependent branches: warp-uniform control flow by construction.
// Accuracy (tests/test_math_primitives.py, host emulation with a 2^-20 seed): <= 2 ulp each.
// Fast-math relaxations, stated: no denormal / inf / NaN handling, results not correctly rounded (<= 2 ulp).
#pragma once
#include "i2_vec.cuh"

#if !defined(__CUDA_ARCH__)
#include <cstdint>
#include <cstring>
#endif

namespace i2 {

// Polynomial coefficients live in constant memory on the device, uploaded at context creation (i2_create): a DFMA
// then takes them as c[bank][offset] operands, which costs no issue slot.  As literals each double needs two UMOVs into
// uniform registers (~250 issue slots per pair); initialised __constant__ arrays get folded back into literals.
#define I2_MATH_TABLE_VALUES                                                                                                   \
    {2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0, 2.0 / 17.0, 2.0 / 19.0, 2.0 / 21.0,       /* atanh series  [0..9]  */ \
     -1.0 / 3.0, 1.0 / 5.0, -1.0 / 7.0, 1.0 / 9.0, -1.0 / 11.0, 1.0 / 13.0, -1.0 / 15.0, 1.0 / 17.0, -1.0 / 19.0,              /* atan series  [10..18] */ \
     6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.57079632679489655800e+00, 3.14159265358979311600e+00,         /* ln2 hi, ln2 lo, pi/2, pi [19..22] */ \
     1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0}       /* atanh series, halved [23..32] */
constexpr int I2_MATH_TABLE_SIZE = 33;
#if defined(__CUDACC__)
__constant__ double c_mathTable[I2_MATH_TABLE_SIZE];
static const double h_mathTable[I2_MATH_TABLE_SIZE] = I2_MATH_TABLE_VALUES;
#if defined(__CUDA_ARCH__)
#define I2_K(idx) c_mathTable[idx]
#else
#define I2_K(idx) h_mathTable[idx]
#endif
#else
static const double h_mathTable[I2_MATH_TABLE_SIZE] = I2_MATH_TABLE_VALUES;
#define I2_K(idx) h_mathTable[idx]
#endif

I2_HD int hi_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
I2_HD int lo_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu);
#endif
}
I2_HD double make_double(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

// hardware seeds: ~2^-22 relative accuracy, use only the high word of the operand
I2_HD double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y;
#else
    return (1.0 / sqrt(x)) * (1.0 + 9.5e-7);   // host emulation: deliberately ~2^-20 off, any magnitude
#endif
}
I2_HD double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y;
#else
    return (1.0 / x) * (1.0 - 9.5e-7);
#endif
}

I2_HD double fast_sqrt(double x) {
    // one cubic step: sqrt x = g (1 - e)^(-1/2) with g = x y, e = 1 - x y^2 (|e| ~ 2^-19 for the 2^-20 seed):
    // 1 + e/2 + 3 e^2/8, next term 5 e^3/16 < 2^-58.  The rounding of g is half compensated through e: <= 1 ulp.
    const double y = rsqrt_seed(x);
    const double g = x * y;
    const double e = fma(-g, y, 1.0);
    const double q = e * fma(0.375, e, 0.5);
    return fma(g, q, g);
}

I2_HD double fast_rcp(double x) {
    const double y = rcp_seed(x);
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);    // y (1 + e + e^2), next term e^3 < 2^-60
}

// ln(N/D), N, D > 0.  N = 2^eN mN, D = 2^eD mD with mN, mD in [1,2); the pair is rescaled by one more factor 2 if
// needed so that mN/mD lies in about [1/sqrt2, sqrt2] (decided on the FP32 pipe, the exact cut does not matter);
// then ln(mN/mD) = 2 atanh(f), f = (mN-mD)/(mN+mD), |f| <= 0.1716: odd series in f up to f^21.
// mN - mD is exact (Sterbenz), so the result keeps full RELATIVE accuracy when N/D -> 1, which log(N/D) does not.
template <bool RESID = true>
I2_HD double log_ratio(double N, double D) {
    const int hN = hi_word(N), hD = hi_word(D);
    int e = (hN >> 20) - (hD >> 20);
    int mhN = (hN & 0x000fffff) | 0x3ff00000, mhD = (hD & 0x000fffff) | 0x3ff00000;
    // FP32 views of the mantissas (top 20 bits are plenty to pick the branch)
#if defined(__CUDA_ARCH__)
    const float fN = __int_as_float(0x3f800000 | ((hN & 0x000fffff) << 3));
    const float fD = __int_as_float(0x3f800000 | ((hD & 0x000fffff) << 3));
#else
    float fN, fD; { int a = 0x3f800000 | ((hN & 0x000fffff) << 3), b = 0x3f800000 | ((hD & 0x000fffff) << 3); memcpy(&fN, &a, 4); memcpy(&fD, &b, 4); }
#endif
    const bool big = fN > 1.41421356f * fD, small = fN * 1.41421356f < fD;
    mhN -= big ? 0x00100000 : 0;     // mN /= 2
    mhD -= small ? 0x00100000 : 0;   // mD /= 2
    e += (big ? 1 : 0) - (small ? 1 : 0);
    const double mN = make_double(mhN, lo_word(N)), mD = make_double(mhD, lo_word(D));
    const double s = mN + mD, d = mN - mD;
    const double r = fast_rcp(s);
    double f = d * r;
    if (RESID) f = fma(fma(-f, s, d), r, f);    // residual correction: f is now (mN-mD)/(mN+mD) to ~0.5 ulp
    const double z = f * f;
    double p = I2_K(9);
    p = fma(p, z, I2_K(8));
    p = fma(p, z, I2_K(7));
    p = fma(p, z, I2_K(6));
    p = fma(p, z, I2_K(5));
    p = fma(p, z, I2_K(4));
    p = fma(p, z, I2_K(3));
    p = fma(p, z, I2_K(2));
    p = fma(p, z, I2_K(1));
    p = fma(p, z, I2_K(0));
    const double lg = fma(f * z, p, f + f);                 // 2 atanh(f)
    const double ed = (double)e;
    return fma(ed, I2_K(19), fma(ed, I2_K(20), lg));   // ln2 split hi/lo
}

// Far-field shortcuts, used when a whole warp qualifies (the caller reduces the "margin" over the warp):
//   atanh_series<NC>: atanh(d/s) for |d/s| small — f = (N-D)/(N+D) is scale invariant, so no exponent / mantissa surgery;
//   atan_series<NC> : atan(y/x) for x > 0 and |y/x| small — no octant logic, no argument reduction.
// Same series as the general versions, truncated to the NC terms the margin allows; they drop the integer-pipe
// bookkeeping (~20 issue slots each) and 4-7 Horner steps for far pairs.
// The margin is measured on the integer pipe: positive doubles order like their bit patterns and
// g(x) = (high word of x) / 2^20 satisfies g(x) <= log2(x) + const <= g(x) + 0.0861, so
//   high(s) - high(|d|) >= k * 2^20   implies   |d/s| <= 2^-(k - 0.0862).
constexpr int kMarginNear1 = 2768241;   // 2.64 * 2^20: |f| <= 0.1716 -> 10 terms (the general version's reduced range)
constexpr int kMarginFar = 4 << 20;     // |f| <= 0.0664 ->  6 terms (next term f^14/15 < 5e-18)
constexpr int kMarginVeryFar = 7 << 20; // |f| <= 0.0084 ->  3 terms (next term f^8/9   < 3e-18)
constexpr int kAngleTiny = 3 << 20;     // |y/x| <= 0.133 -> 9 terms (the general version's reduced range)
constexpr int kAngleFar = 6 << 20;      // |y/x| <= 0.0166 -> 4 terms (next term t^10/11 < 2e-19)
// s = N + D > 0, d = N - D: margin of |d/s| (large = far); d = 0 gives the largest margin
I2_HD int ratio_margin(double s, double d) { return hi_word(s) - (hi_word(d) & 0x7fffffff); }
// margin of |y/x|, or a negative number when x <= 0
I2_HD int angle_margin(double y, double x) {
    const int hx = hi_word(x);
    return hx > 0 ? hx - (hi_word(y) & 0x7fffffff) : -1;
}
// atanh(d/s) = ln(N/D) / 2: the caller folds the factor 2 into the quadrature weight
template <int NC, bool RESID = true>
I2_HD double atanh_series(double s, double d) {
    const double r = fast_rcp(s);
    double f = d * r;
    if (RESID) f = fma(fma(-f, s, d), r, f);
    const double z = f * f;
    double p = I2_K(23 + NC - 1);
#pragma unroll
    for (int k = NC - 2; k >= 0; --k) p = fma(p, z, I2_K(23 + k));
    return fma(f * z, p, f);
}
template <int NC, bool RESID = true>
I2_HD double atan_series(double y, double x) {
    const double r = fast_rcp(x);
    double t = y * r;
    if (RESID) t = fma(fma(-t, x, y), r, t);
    const double z = t * t;
    double p = I2_K(10 + NC - 1);
#pragma unroll
    for (int k = NC - 2; k >= 0; --k) p = fma(p, z, I2_K(10 + k));
    return fma(t * z, p, t);
}

// atan2(y, x) for finite arguments, not both zero.  t = min/max in [0,1]; c = nearest of {0, 1/4, 1/2, 3/4, 1};
// atan t = atan c + atan((t-c)/(1+tc)) = atan c + atan((mn - c mx)/(mx + c mn)), |arg| <= 0.1244: odd series to ^19.
template <bool RESID = true>
I2_HD double atan2_fast(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    // choose c on the integer / FP32 pipes from the leading bits of the operands (any magnitude: only the exponent
    // difference and the top 20 mantissa bits are used, so nothing overflows a float)
    const int hm = hi_word(mn), hx = hi_word(mx);
    const int de = (hx >> 20) - (hm >> 20);                      // >= 0
    const int bm = de < 64 ? (((127 - de) << 23) | ((hm & 0x000fffff) << 3)) : 0;
    const int bx = 0x3f800000 | ((hx & 0x000fffff) << 3);
#if defined(__CUDA_ARCH__)
    const float q = __fdividef(__int_as_float(bm), __int_as_float(bx));
#else
    float fm, fx; memcpy(&fm, &bm, 4); memcpy(&fx, &bx, 4);
    const float q = fm / fx;
#endif
    const int k = (int)(q * 4.0f + 0.5f);                       // 0..4
    const double c = 0.25 * (double)k;
    // atan(k/4), k = 0..4
    const double at = k == 0 ? 0.0 : (k == 1 ? 2.44978663126864154172e-01 : (k == 2 ? 4.63647609000806116214e-01
                      : (k == 3 ? 6.43501108793284386803e-01 : 7.85398163397448309616e-01)));
    const double num = fma(-c, mx, mn), den = fma(c, mn, mx);
    const double r = fast_rcp(den);
    double t = num * r;
    if (RESID) t = fma(fma(-t, den, num), r, t);
    const double z = t * t;
    double p = I2_K(18);
    p = fma(p, z, I2_K(17));
    p = fma(p, z, I2_K(16));
    p = fma(p, z, I2_K(15));
    p = fma(p, z, I2_K(14));
    p = fma(p, z, I2_K(13));
    p = fma(p, z, I2_K(12));
    p = fma(p, z, I2_K(11));
    p = fma(p, z, I2_K(10));
    double a = at + fma(t * z, p, t);                           // atan(mn/mx) in [0, pi/4]
    a = ay > ax ? I2_K(21) - a : a;           // pi/2 - a
    a = x < 0.0 ? I2_K(22) - a : a;           // pi - a
    return copysign(a, y);
}

}  // namespace i2
