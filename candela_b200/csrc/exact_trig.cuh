// sin / cos / acos / pow with DEFINED bits ("cndl exact math v1").
//
// GLSL leaves the results of these built-ins implementation-defined, and CUDA's and a host libm's versions differ in the
// last bit, so the ray generators (kernels_raygen.cu) — whose output must be reproducible bit for bit by a CPU
// checker — use the definitions below: IEEE double arithmetic with separately rounded operations (the intrinsics
// never contract into FMAs), rounded to float once at the end.  Formulas:
//   xsincos : k = floor(x * 2/pi + 0.5); r = (x - k*PIO2_HI) - k*PIO2_LO; Taylor polynomials to r^13 / r^14 (Horner);
//             quadrant from k & 3.
//   xacos   : asin series (22 terms; term ratio (2n-1)^2 / (2n (2n+1))) on |x| <= 0.5, else via asin(sqrt((1-|x|)/2)).
//   xpow    : exp(y * log(x)) for x > 0: log = e*ln2 + 2 atanh((m-1)/(m+1)), m in [sqrt(1/2), sqrt(2));
//             exp = degree-14 Taylor polynomial after reduction by ln 2.
// Within 1 ulp of sinf / cosf / acosf / powf.
#pragma once
#include <cuda_runtime.h>

namespace cndl {
namespace xm {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

#define CNDL_TWO_OVER_PI 0.63661977236758138
#define CNDL_PIO2_HI 1.5707963267948966
#define CNDL_PIO2_LO 6.123233995736766e-17
#define CNDL_PI_D 3.1415926535897931
#define CNDL_LN2_HI 0.693147180369123816490
#define CNDL_LN2_LO 1.90821492927058770002e-10
#define CNDL_INV_LN2 1.4426950408889634

__device__ __forceinline__ void sincos_reduced(double r, double& s, double& c) {
    const double r2 = dmul(r, r);
    double ps = 1.0 / 6227020800.0;
    ps = dadd(dmul(ps, r2), -1.0 / 39916800.0);
    ps = dadd(dmul(ps, r2), 1.0 / 362880.0);
    ps = dadd(dmul(ps, r2), -1.0 / 5040.0);
    ps = dadd(dmul(ps, r2), 1.0 / 120.0);
    ps = dadd(dmul(ps, r2), -1.0 / 6.0);
    s = dadd(r, dmul(dmul(r, r2), ps));
    double pc = -1.0 / 87178291200.0;
    pc = dadd(dmul(pc, r2), 1.0 / 479001600.0);
    pc = dadd(dmul(pc, r2), -1.0 / 3628800.0);
    pc = dadd(dmul(pc, r2), 1.0 / 40320.0);
    pc = dadd(dmul(pc, r2), -1.0 / 720.0);
    pc = dadd(dmul(pc, r2), 1.0 / 24.0);
    pc = dadd(dmul(pc, r2), -0.5);
    c = dadd(1.0, dmul(r2, pc));
}

__device__ __forceinline__ void xsincos(float x, float& s_out, float& c_out) {
    const double xd = (double)x;
    const double kd = floor(dadd(dmul(xd, CNDL_TWO_OVER_PI), 0.5));
    const double r = dsub(dsub(xd, dmul(kd, CNDL_PIO2_HI)), dmul(kd, CNDL_PIO2_LO));
    double s, c;
    sincos_reduced(r, s, c);
    const int q = (int)((long long)kd & 3ll);
    const double so = q == 0 ? s : (q == 1 ? c : (q == 2 ? -s : -c));
    const double co = q == 0 ? c : (q == 1 ? -s : (q == 2 ? -c : s));
    s_out = __double2float_rn(so);
    c_out = __double2float_rn(co);
}

__device__ __forceinline__ double asin_series(double a) {  // 0 <= a <= 0.5
    const double a2 = dmul(a, a);
    double term = a, sum = a;
#pragma unroll 1
    for (int n = 1; n <= 22; ++n) {
        const double num = (double)((2 * n - 1) * (2 * n - 1)), den = (double)((2 * n) * (2 * n + 1));
        term = ddiv(dmul(dmul(term, a2), num), den);
        sum = dadd(sum, term);
    }
    return sum;
}

__device__ __forceinline__ float xacos(float x) {
    const double xd = (double)x;
    const double a = xd < 0.0 ? -xd : xd;
    if (!(a <= 1.0)) return __int_as_float(0x7FC00000);
    double r;
    if (a <= 0.5) {
        const double as = asin_series(a);
        r = xd < 0.0 ? dadd(dadd(CNDL_PIO2_HI, as), CNDL_PIO2_LO) : dadd(dsub(CNDL_PIO2_HI, as), CNDL_PIO2_LO);
    } else {
        const double z = dmul(dsub(1.0, a), 0.5);
        const double as2 = dmul(2.0, asin_series(__dsqrt_rn(z)));
        r = xd < 0.0 ? dsub(CNDL_PI_D, as2) : as2;
    }
    return __double2float_rn(r);
}

__device__ __forceinline__ double xlog_d(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    int e = (int)((b >> 52) & 0x7FFull) - 1023;
    b = (b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m = __longlong_as_double((long long)b);
    if (m > 1.4142135623730951) { m = dmul(m, 0.5); e = e + 1; }
    const double f = ddiv(dsub(m, 1.0), dadd(m, 1.0)), f2 = dmul(f, f);
    double p = 1.0 / 25.0;
#pragma unroll 1
    for (int k = 23; k >= 3; k -= 2) p = dadd(dmul(p, f2), ddiv(1.0, (double)k));
    const double lm = dmul(2.0, dadd(f, dmul(dmul(f, f2), p)));
    const double ed = (double)e;
    return dadd(dmul(ed, CNDL_LN2_HI), dadd(dmul(ed, CNDL_LN2_LO), lm));
}

__device__ __forceinline__ double xexp_d(double t) {
    if (t > 700.0) return __longlong_as_double(0x7FF0000000000000ll);
    if (t < -740.0) return 0.0;
    const double kd = floor(dadd(dmul(t, CNDL_INV_LN2), 0.5));
    const double r = dsub(dsub(t, dmul(kd, CNDL_LN2_HI)), dmul(kd, CNDL_LN2_LO));
    const double inv_fact[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                                 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
    double p = 1.0 / 87178291200.0;
#pragma unroll
    for (int k = 0; k < 14; ++k) p = dadd(dmul(p, r), inv_fact[k]);
    const int k = (int)kd;
    const int k1 = k / 2, k2 = k - k1;
    const double s1 = __longlong_as_double((long long)(k1 + 1023) << 52), s2 = __longlong_as_double((long long)(k2 + 1023) << 52);
    return dmul(dmul(p, s1), s2);
}

__device__ __forceinline__ float xpow(float x, float y) {
    if (x != x || y != y || x < 0.0f) return __int_as_float(0x7FC00000);
    if (x == 0.0f) return y > 0.0f ? 0.0f : (y == 0.0f ? 1.0f : __int_as_float(0x7F800000));
    if (isinf(x)) return y > 0.0f ? __int_as_float(0x7F800000) : (y == 0.0f ? 1.0f : 0.0f);
    return __double2float_rn(xexp_d(dmul((double)y, xlog_d((double)x))));
}

}  // namespace xm
}  // namespace cndl
