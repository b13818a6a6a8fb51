/*
 * tcr_libm.h -- bit-reproducible float64 elementary functions.
 *
 * Why this exists.  The reference's adaptive RK45 integration (scipy solve_ivp, call site
 * intensity/coupled_fast.py:264) on grid-rough fields amplifies a 1-ulp perturbation of the
 * state by roughly 3x per accepted step (measured: DESIGN.md "Chaos floor"), and its exact
 * `f_land.ev(...) == 1` test (coupled_fast.py:38) is decided by the last rounding bit.  Two
 * implementations that call two different libm's (glibc on the host, libdevice on the GPU)
 * therefore drift apart to the integrator's own rtol (1e-3) on long-lived storms.  To make
 * the CUDA path and the CPU oracle comparable BIT FOR BIT, both evaluate every transcendental
 * through the routines below: straight-line IEEE-754 binary64 add/mul/div/sqrt plus explicit
 * fma(), no table look-ups, no data-dependent library calls.  Compiled with contraction
 * disabled (gcc -ffp-contract=off, nvcc -fmad=false) they round identically on x86-64 and
 * sm_100a.
 *
 * The algorithms are the classical Sun fdlibm ones (Cody-Waite reduction + minimax kernels),
 * restated; each is accurate to < 1.5 ulp on the argument ranges the hot path uses, which is
 * checked against glibc in tests/test_libm.py.
 *
 * This header is shared by the product (csrc/) and by the test-only oracle (oracle/) in the
 * same way both would otherwise share a system libm.
 */
#ifndef TCR_LIBM_H
#define TCR_LIBM_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TCR_HD __host__ __device__ __forceinline__
#else
#define TCR_HD static inline
#endif

#define TCR_PI 3.14159265358979323846
#define TCR_DEG2RAD (TCR_PI / 180.0)

TCR_HD int64_t tcr_d2bits(double x)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    union { double d; int64_t i; } u; u.d = x; return u.i;
#endif
}

TCR_HD double tcr_bits2d(int64_t i)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(i);
#else
    union { double d; int64_t i; } u; u.i = i; return u.d;
#endif
}

TCR_HD int tcr_isnan(double x) { return x != x; }

/* 2^k for -1022 <= k <= 1023 */
TCR_HD double tcr_pow2i(int k) { return tcr_bits2d((int64_t)(1023 + k) << 52); }

/* ---- exp ---------------------------------------------------------------------------- */
TCR_HD double tcr_exp(double x)
{
    const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10;
    const double invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    if (tcr_isnan(x)) return x;
    if (x > 709.0) return INFINITY;
    if (x < -708.0) return 0.0;               /* flush: the hot path never needs subnormal results */
    double fk = floor(x * invln2 + 0.5);
    int k = (int)fk;
    double hi = x - fk * ln2hi;
    double lo = fk * ln2lo;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    return y * tcr_pow2i(k);
}

/* ---- log (x > 0 normal; <= 0 and NaN give NaN / -inf like libm) ----------------------- */
TCR_HD double tcr_log(double x)
{
    const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    if (tcr_isnan(x)) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    int k = 0;
    if (x < 2.2250738585072014e-308) { x *= 18014398509481984.0; k -= 54; }   /* subnormal */
    int64_t b = tcr_d2bits(x);
    int32_t hx = (int32_t)(b >> 32);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    b = ((int64_t)(hx | (i ^ 0x3ff00000)) << 32) | (b & 0xffffffffLL);
    k += (i >> 20);
    double f = tcr_bits2d(b) - 1.0;
    double dk = (double)k;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    return dk * ln2hi - ((hfsq - (s * (hfsq + R) + dk * ln2lo)) - f);
}

/* x^p for x >= 0 through exp(p log x); x == 0 gives 0 for p > 0 and +inf for p < 0 */
TCR_HD double tcr_pow(double x, double p)
{
    if (x == 0.0) return p > 0.0 ? 0.0 : (p < 0.0 ? INFINITY : 1.0);
    return tcr_exp(p * tcr_log(x));
}

/* ---- sin / cos kernels on |x| <= pi/4 -------------------------------------------------- */
TCR_HD double tcr_ksin(double x)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x + v * (S1 + z * r);
}

TCR_HD double tcr_kcos(double x)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + z * r);
}

/* sin and cos of x (radians), |x| < ~1e5: two-term Cody-Waite reduction by pi/2 */
TCR_HD void tcr_sincos(double x, double* s, double* c)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;    /* first 33 bits of pi/2 */
    const double pio2_1t = 6.07710050650619224932e-11;   /* pi/2 - pio2_1 */
    double fn = floor(x * invpio2 + 0.5);
    double r = (x - fn * pio2_1) - fn * pio2_1t;
    int n = (int)fn & 3;
    double ks = tcr_ksin(r), kc = tcr_kcos(r);
    double ss = (n & 1) ? kc : ks;
    double cc = (n & 1) ? ks : kc;
    if (n == 1 || n == 2) cc = -cc;
    if (n >= 2) ss = -ss;
    *s = ss; *c = cc;
}

TCR_HD double tcr_sin(double x) { double s, c; tcr_sincos(x, &s, &c); return s; }
TCR_HD double tcr_cos(double x) { double s, c; tcr_sincos(x, &s, &c); return c; }

/* sin and cos of 2*pi*u for any finite u: exact reduction in turns */
TCR_HD void tcr_sincos2pi(double u, double* s, double* c)
{
    double fq = floor(4.0 * u + 0.5);
    double r = (u - 0.25 * fq) * (2.0 * TCR_PI);          /* |r| <= pi/4, u - fq/4 is exact */
    int n = (int)(fq - 4.0 * floor(0.25 * fq)) & 3;
    double ks = tcr_ksin(r), kc = tcr_kcos(r);
    double ss = (n & 1) ? kc : ks;
    double cc = (n & 1) ? ks : kc;
    if (n == 1 || n == 2) cc = -cc;
    if (n >= 2) ss = -ss;
    *s = ss; *c = cc;
}

/* ---- asin on [-1, 1] -------------------------------------------------------------------- */
TCR_HD double tcr_asin(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double ax = fabs(x);
    if (!(ax <= 1.0)) return NAN;
    if (ax == 1.0) return x * pio2_hi + x * pio2_lo;      /* asin(+-1) = +-pi/2 (fdlibm special case) */
    if (ax < 0.5) {
        double t = x * x;
        double p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
        double q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
        return x + x * (p / q);
    }
    double w = 1.0 - ax;
    double t = w * 0.5;
    double p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
    double q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
    double s = sqrt(t);
    double r = p / q;
    /* asin(|x|) = pi/2 - 2 asin(sqrt((1-|x|)/2)); sf = s with the low word cleared so that
       c = (t - sf*sf)/(s + sf) recovers the bits lost in the square root */
    double sf = tcr_bits2d(tcr_d2bits(s) & (int64_t)0xffffffff00000000LL);
    double c = (t - sf * sf) / (s + sf);
    double pp = 2.0 * s * r - (pio2_lo - 2.0 * c);
    double qq = 0.5 * pio2_hi - 2.0 * sf;
    double res = 0.5 * pio2_hi - (pp - qq);
    return x > 0.0 ? res : -res;
}

/* ---- tanh (absolute accuracy ~2e-16; used only in an additive blend, tc_wind.py:8) ------ */
TCR_HD double tcr_tanh(double x)
{
    if (tcr_isnan(x)) return x;
    double ax = fabs(x);
    double t = (ax > 20.0) ? 1.0 : 1.0 - 2.0 / (tcr_exp(2.0 * ax) + 1.0);
    return x < 0.0 ? -t : t;
}

/* ---- great-circle distance, km (haversine; reference util/sphere.py:15-30 and the notebook's
 * own copy, notebooks/sample_analysis.ipynb cell 13), radius of the sphere in km ---------------- */
TCR_HD double tcr_haversine_r(double r_km, double lon1, double lat1, double lon2, double lat2)
{
    lon1 = lon1 * TCR_DEG2RAD; lat1 = lat1 * TCR_DEG2RAD;
    lon2 = lon2 * TCR_DEG2RAD; lat2 = lat2 * TCR_DEG2RAD;
    double dlon = lon2 - lon1, dlat = lat2 - lat1;
    double sa = tcr_sin(dlat / 2), sb = tcr_sin(dlon / 2);
    double a = sa * sa + tcr_cos(lat1) * tcr_cos(lat2) * (sb * sb);
    double c = 2.0 * tcr_asin(sqrt(a));
    return r_km * c;
}

#endif /* TCR_LIBM_H */
