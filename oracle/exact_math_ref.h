// TEST INFRASTRUCTURE ONLY — never included by the product library.
//
// GLSL leaves the bits of sin / cos / acos / pow implementation-defined, so a ray generator that must agree bit for
// bit on a CPU and a GPU needs ONE definition of them.  This header is the CPU statement of that definition
// ("cndl exact math v1"); candela_b200/csrc/exact_trig.cuh is the CUDA statement of the same formulas.  Everything is
// evaluated in IEEE double with separately rounded +, -, *, / and sqrt (compile with -ffp-contract=off) and rounded
// to float once at the end, so both sides produce the same bits on every input.  The results are within 1 ulp of
// libm's sinf / cosf / acosf / powf (tests/test_raygen_oracle.py measures it).
//
//   xsin / xcos : k = floor(x * 2/pi + 0.5); r = (x - k*PIO2_HI) - k*PIO2_LO; Taylor polynomials to r^13 / r^14 in
//                 Horner form; quadrant from k & 3.  Meant for |x| < 1e5 (the generators pass angles in [0, 2 pi]).
//   xacos       : asin series (22 terms, ratio (2n-1)^2 / (2n (2n+1))) on |x| <= 0.5, else through
//                 asin(sqrt((1 - |x|) / 2)).
//   xpow(x, y)  : x > 0: exp(y * log(x)); log through 2 atanh((m-1)/(m+1)) on the mantissa folded to
//                 [sqrt(1/2), sqrt(2)), exp through a degree-14 Taylor polynomial after reduction by ln 2.
//                 x == 0: 0 for y > 0.  x < 0 or NaN: NaN (GLSL: undefined).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace xm {

constexpr double TWO_OVER_PI = 0.63661977236758138;      // 0x3FE45F306DC9C883
constexpr double PIO2_HI = 1.5707963267948966;           // 0x3FF921FB54442D18
constexpr double PIO2_LO = 6.123233995736766e-17;        // pi/2 - PIO2_HI
constexpr double PI_D = 3.1415926535897931;
constexpr double LN2_HI = 0.693147180369123816490;       // 0x3FE62E42FEE00000
constexpr double LN2_LO = 1.90821492927058770002e-10;    // 0x3DEA39EF35793C76
constexpr double INV_LN2 = 1.4426950408889634;

inline void sincos_reduced(double r, double& s, double& c) {
    const double r2 = r * r;
    double ps = 1.0 / 6227020800.0;                      //  1/13!
    ps = ps * r2 + -1.0 / 39916800.0;                    // -1/11!
    ps = ps * r2 + 1.0 / 362880.0;                       //  1/9!
    ps = ps * r2 + -1.0 / 5040.0;                        // -1/7!
    ps = ps * r2 + 1.0 / 120.0;                          //  1/5!
    ps = ps * r2 + -1.0 / 6.0;                           // -1/3!
    s = r + (r * r2) * ps;
    double pc = -1.0 / 87178291200.0;                    // -1/14!
    pc = pc * r2 + 1.0 / 479001600.0;                    //  1/12!
    pc = pc * r2 + -1.0 / 3628800.0;                     // -1/10!
    pc = pc * r2 + 1.0 / 40320.0;                        //  1/8!
    pc = pc * r2 + -1.0 / 720.0;                         // -1/6!
    pc = pc * r2 + 1.0 / 24.0;                           //  1/4!
    pc = pc * r2 + -0.5;                                 // -1/2!
    c = 1.0 + r2 * pc;
}

inline void xsincos(float x, float& s_out, float& c_out) {
    const double xd = (double)x;
    const double kd = std::floor(xd * TWO_OVER_PI + 0.5);
    const double r = (xd - kd * PIO2_HI) - kd * PIO2_LO;
    double s, c;
    sincos_reduced(r, s, c);
    const long long k = (long long)kd;
    switch ((int)(k & 3)) {
        case 0: s_out = (float)s; c_out = (float)c; break;
        case 1: s_out = (float)c; c_out = (float)(-s); break;
        case 2: s_out = (float)(-s); c_out = (float)(-c); break;
        default: s_out = (float)(-c); c_out = (float)s; break;
    }
}
inline float xsin(float x) { float s, c; xsincos(x, s, c); return s; }
inline float xcos(float x) { float s, c; xsincos(x, s, c); return c; }

inline double asin_series(double a) {  // 0 <= a <= 0.5
    const double a2 = a * a;
    double term = a, sum = a;
    for (int n = 1; n <= 22; ++n) {
        const double num = (double)((2 * n - 1) * (2 * n - 1)), den = (double)((2 * n) * (2 * n + 1));
        term = ((term * a2) * num) / den;
        sum = sum + term;
    }
    return sum;
}

inline float xacos(float x) {
    const double xd = (double)x;
    const double a = xd < 0.0 ? -xd : xd;
    if (!(a <= 1.0)) return std::nanf("");
    double r;
    if (a <= 0.5) {
        const double as = asin_series(a);
        r = xd < 0.0 ? (PIO2_HI + as) + PIO2_LO : (PIO2_HI - as) + PIO2_LO;
    } else {
        const double z = (1.0 - a) * 0.5;
        const double as2 = 2.0 * asin_series(std::sqrt(z));
        r = xd < 0.0 ? PI_D - as2 : as2;
    }
    return (float)r;
}

inline double xlog_d(double x) {  // x > 0, finite, normal or subnormal double that came from a float (so normal as a double)
    uint64_t b;
    std::memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7FF) - 1023;
    b = (b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m;
    std::memcpy(&m, &b, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e = e + 1; }
    const double f = (m - 1.0) / (m + 1.0), f2 = f * f;
    double p = 1.0 / 25.0;
    for (int k = 23; k >= 3; k -= 2) p = p * f2 + 1.0 / (double)k;
    const double lm = 2.0 * (f + (f * f2) * p);
    const double ed = (double)e;
    return ed * LN2_HI + (ed * LN2_LO + lm);
}

inline double xexp_d(double t) {
    if (t > 700.0) return std::numeric_limits<double>::infinity();
    if (t < -740.0) return 0.0;
    const double kd = std::floor(t * INV_LN2 + 0.5);
    const double r = (t - kd * LN2_HI) - kd * LN2_LO;
    double p = 1.0 / 87178291200.0;  // 1/14!
    const double inv_fact[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                                 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
    for (int k = 0; k < 14; ++k) p = p * r + inv_fact[k];
    // p * 2^k in two steps so that results below the normal range round once, in the final conversion to float
    const int k = (int)kd;
    const int k1 = k / 2, k2 = k - k1;
    uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
    double s1, s2;
    std::memcpy(&s1, &b1, 8);
    std::memcpy(&s2, &b2, 8);
    return (p * s1) * s2;
}

inline float xpow(float x, float y) {
    if (x != x || y != y || x < 0.0f) return std::nanf("");
    if (x == 0.0f) return y > 0.0f ? 0.0f : (y == 0.0f ? 1.0f : std::numeric_limits<float>::infinity());
    if (std::isinf(x)) return y > 0.0f ? std::numeric_limits<float>::infinity() : (y == 0.0f ? 1.0f : 0.0f);
    return (float)xexp_d((double)y * xlog_d((double)x));
}

}  // namespace xm
