// fp64 building blocks of the CMEM kernel (sm_100a): exp, log2, pow, acos and the sin / cos of an
// angle that is itself a rounded acos.  ~1e-15 accurate, i.e. rounding one of them to fp32 gives the
// correctly rounded fp32 result except within ~1e-8 of a rounding boundary (checked on the device
// against the CUDA fp64 library, tests/test_gpu_parity.py::test_cmem_math_accuracy).
//
// Why not the CUDA fp64 library: it inlines every polynomial coefficient as a 64-bit immediate, and
// sm_100 materialises each one with two UMOV / IMAD.MOV instructions in front of the DFMA that uses
// it -- in the library-based kernel those were 600 of 2800 static instructions (profiles/r1l), and
// issue slots are what that kernel runs out of.  Here every coefficient sits in __constant__ memory
// (NOT `const`: a const table is folded back into immediates) and is read by the DFMA itself as a
// constant-bank operand.  And the geometry of CMEM (theta = acosf(x/r), then sinf / cosf / cosf(theta/2)
// of that very theta) lets the sines and cosines come from two square roots instead of two sincos
// evaluations (angle_from_cosine).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace birp {
namespace cm {

// Taylor coefficients 1/11! ... 1/2! of e^r (exact rationals rounded), highest first
static __constant__ double kExpC[10] = {1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                                        1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,     1.0 / 6.0,     0.5};
// log1p(z) = z - z^2/2 + ... + z^7/7, highest first
static __constant__ double kLog1pC[6] = {1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5};
// asin(t) = t + t^3 R(t^2), |t| <= 1/2, R of degree 10 in z = t^2, highest first; max abs error 5.0e-16
// (tools/gen_cmem_coeffs.py 10: Chebyshev-node interpolant computed with mpmath, error measured in fp64 Horner form)
static __constant__ double kAsinC[11] = {0x1.c8a4a8d5d7026p-6,  -0x1.bf16e7c9f283cp-8, 0x1.fa1b2b4831188p-7, 0x1.512bc40e88a9ep-7,
                                         0x1.cf5ed14c7cb7ep-7,  0x1.1c0d74beb3610p-6,  0x1.6e8f34a32a3ecp-6, 0x1.f1c6ff7f5507fp-6,
                                         0x1.6db6dba99e56dp-5,  0x1.33333333030cfp-4,  0x1.55555555555bbp-3};
// {ln2 high part (32 trailing zero bits), ln2 low part, log2(e), 1/sqrt(2), pi, pi/2, ln2}
static __constant__ double kMisc[7] = {0x1.62e42fefa3800p-1, 0x1.ef35793c76730p-45, 1.44269504088896338700e+00,
                                       0.70710678118654752440, 3.14159265358979323846, 1.57079632679489661923,
                                       6.93147180559945286227e-01};

// ---- exp / exp2 --------------------------------------------------------------------------------
// e^r for |r| <= ln2/2, Taylor to r^11: truncation < 2^-47 relative... (r^12/12! <= 6.3e-15)
__device__ __forceinline__ double expm_core(double r) {
    double p = kExpC[0];
#pragma unroll
    for (int k = 1; k < 10; ++k) p = fma(p, r, kExpC[k]);
    p = fma(p, r, 1.0);
    return fma(p, r, 1.0);
}

// v * 2^k for v in [0.5, 2) and |k| <= 1000: one integer add on the exponent field
__device__ __forceinline__ double scale2(double v, int k) {
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

static __device__ __noinline__ double exp2_slow(double t) { return exp2(t); }
static __device__ __noinline__ double exp_slow(double x) { return exp(x); }

__device__ __forceinline__ double exp2_d(double t) {
    if (!(fabs(t) < 1000.0)) return exp2_slow(t);
    const double kd = rint(t);
    const double r = (t - kd) * kMisc[6];  // exact difference, |r| <= ln2/2
    return scale2(expm_core(r), (int)kd);
}

// e^x: Cody-Waite reduction x = k ln2 + r with ln2 split in two, so r carries no argument-size error
__device__ __forceinline__ double exp_d(double x) {
    if (!(fabs(x) < 690.0)) return exp_slow(x);
    const double kd = rint(x * kMisc[2]);
    double r = fma(-kd, kMisc[0], x);
    r = fma(-kd, kMisc[1], r);
    return scale2(expm_core(r), (int)kd);
}

// ---- log2 --------------------------------------------------------------------------------------
// 32 intervals of the mantissa m in [1,2): {1/c_i rounded, -ln(that)} with c_i the interval centre,
// so z = m * invc - 1 has |z| <= 1/64 and ln m = log1p(z) + logc.
__device__ const double2 kLogTab[32] = {
    {0x1.f81f81f81f820p-1, 0x1.fc0a8b0fc03c4p-7},
    {0x1.e9131abf0b767p-1, 0x1.77458f632dcffp-5},
    {0x1.dae6076b981dbp-1, 0x1.341d7961bd1d0p-4},
    {0x1.cd85689039b0bp-1, 0x1.a926d3a4ad562p-4},
    {0x1.c0e070381c0e0p-1, 0x1.0d77e7cd08e5bp-3},
    {0x1.b4e81b4e81b4fp-1, 0x1.44d2b6ccb7d1cp-3},
    {0x1.a98ef606a63bep-1, 0x1.7ab890210d907p-3},
    {0x1.9ec8e951033d9p-1, 0x1.af3c94e80bff3p-3},
    {0x1.948b0fcd6e9e0p-1, 0x1.e27076e2af2e8p-3},
    {0x1.8acb90f6bf3aap-1, 0x1.0a324e27390e2p-2},
    {0x1.8181818181818p-1, 0x1.22941fbcf7966p-2},
    {0x1.78a4c8178a4c8p-1, 0x1.3a64c556945eap-2},
    {0x1.702e05c0b8170p-1, 0x1.51aad872df82ep-2},
    {0x1.6816816816817p-1, 0x1.686c81e9b14adp-2},
    {0x1.6058160581606p-1, 0x1.7eaf83b82afc2p-2},
    {0x1.58ed2308158edp-1, 0x1.947941c2116fbp-2},
    {0x1.51d07eae2f815p-1, 0x1.a9cec9a9a084ap-2},
    {0x1.4afd6a052bf5bp-1, 0x1.beb4d9da71b7ap-2},
    {0x1.446f86562d9fbp-1, 0x1.d32fe7e00ebd5p-2},
    {0x1.3e22cbce4a902p-1, 0x1.e744261d68789p-2},
    {0x1.3813813813814p-1, 0x1.faf588f78f31dp-2},
    {0x1.323e34a2b10bfp-1, 0x1.0723e5c1cdf41p-1},
    {0x1.2c9fb4d812ca0p-1, 0x1.109f39e2d4c96p-1},
    {0x1.27350b8812735p-1, 0x1.19ee6b467c96fp-1},
    {0x1.21fb78121fb78p-1, 0x1.23130d7bebf43p-1},
    {0x1.1cf06ada2811dp-1, 0x1.2c0e9ed448e8cp-1},
    {0x1.1811811811812p-1, 0x1.34e289d9ce1d2p-1},
    {0x1.135c81135c811p-1, 0x1.3d9026a7156fbp-1},
    {0x1.0ecf56be69c90p-1, 0x1.4618bc21c5ec2p-1},
    {0x1.0a6810a6810a7p-1, 0x1.4e7d811b75bb0p-1},
    {0x1.0624dd2f1a9fcp-1, 0x1.56bf9d5b3f399p-1},
    {0x1.0204081020408p-1, 0x1.5ee02a9241676p-1},
};

// x finite, positive and normal (every fp32 value widened to fp64 is normal)
__device__ __forceinline__ double log2_d(double x) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 15) & 31;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = kLogTab[idx];
    const double z = fma(m, t.x, -1.0);
    // log1p(z), alternating Taylor series to z^7: truncation < 2^-51
    double p = kLog1pC[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) p = fma(p, z, kLog1pC[k]);
    const double ln_m = fma(p * z, z, z) + t.y;
    return fma(ln_m, kMisc[2], (double)e);
}

// x^y for finite x > 0 (other bases: the caller keeps powf's special cases through pow())
__device__ __forceinline__ double pow_pos_d(double x, double y) { return exp2_d(y * log2_d(x)); }
static __device__ __noinline__ double pow_slow(double x, double y) { return pow(x, y); }

// ---- acos and the angle bundle ------------------------------------------------------------------
// sqrt(x) for 0 <= x <= 4 (1 +- cosine, (1 - |x|) / 2), to within an ulp or two: the fp32 reciprocal-square-root unit
// seeds it (relative error e <= 2^-22) and two coupled Newton steps on s = x y take the error to 1.5 e^2, then to
// 1.5 e^3 -- 10 instructions where the correctly rounded fp64 sqrt takes 20 (64-bit seed, iterations, range fix-up).
// Zero gives zero, a negative or NaN argument NaN, like the library's.
__device__ __forceinline__ double sqrt_small(double x) {
#ifdef __CUDACC__
    float y0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"((float)x));
    const double y0 = (double)y0f, hy = 0.5 * y0;
    double s = x * y0;
    s = fma(fma(-s, s, x), hy, s);
    s = fma(fma(-s, s, x), hy, s);
    return x == 0.0 ? 0.0 : s;
#else
    return sqrt(x);
#endif
}

// asin(t) for |t| <= 1/2
__device__ __forceinline__ double asin_core(double t) {
    const double z = t * t;
    double p = kAsinC[0];
#pragma unroll
    for (int k = 1; k < 11; ++k) p = fma(p, z, kAsinC[k]);
    return fma(t * z, p, t);
}

// acos(x) for any double x (NaN outside [-1,1], like the library):
//   |x| < 1/2:  pi/2 - asin(x);   x >= 1/2:  2 asin(sqrt((1-x)/2));   x <= -1/2:  pi - 2 asin(sqrt((1+x)/2))
__device__ __forceinline__ double acos_d(double x) {
    const double ax = fabs(x);
    if (ax < 0.5) return kMisc[5] - asin_core(x);
    const double t = 2.0 * asin_core(sqrt_small((1.0 - ax) * 0.5));
    return x > 0.0 ? t : kMisc[4] - t;
}

// Everything CMEM needs of an angle a = RN32(acos(q)), q an fp32 cosine (src/Cmem.cpp:224-246 then
// :196-216): the fp32 angle itself and, in fp64, sin(a), cos(a), cos(a/2) of THAT rounded angle.
// With A = sqrt(1+q), B = sqrt(1-q) (1 +- q are exact in fp64):  sin(acos q) = A B,
// cos(acos(q)/2) = A / sqrt2, sin(acos(q)/2) = B / sqrt2, and acos(q) itself is 2 asin(B / sqrt2)
// (or pi - 2 asin(A / sqrt2), or pi/2 - asin(q)).  The rounded angle differs from acos(q) by
// d = a - acos(q), |d| <= 1.2e-7, so   cos a = q cos d - AB sin d,   sin a = AB cos d + q sin d,
// cos(a/2) = (A cos(d/2) - B sin(d/2)) / sqrt2   with cos d = 1 - d^2/2, sin d = d (next terms < 1e-21).
// Two square roots and one asin polynomial replace acos + sincos + cos.
struct Angle {
    float a;        // RN32(acos(q))
    double s, c;    // sin(a), cos(a)
    double c_half;  // cos(a / 2)
};

__device__ __forceinline__ Angle angle_from_cosine(float qf) {
    const double q = (double)qf;
    const double A = sqrt_small(1.0 + q), B = sqrt_small(1.0 - q);
    const double aq = fabs(q);
    // one asin_core for all three ranges
    const double t = (aq < 0.5) ? q : ((q > 0.0 ? B : A) * kMisc[3]);
    const double as = asin_core(t);
    const double acq = (aq < 0.5) ? (kMisc[5] - as) : (q > 0.0 ? 2.0 * as : fma(-2.0, as, kMisc[4]));
    Angle r;
    r.a = (float)acq;
    const double d = (double)r.a - acq;
    const double h = 0.5 * d * d;
    const double sq = A * B;
    r.c = fma(-q, h, fma(-sq, d, q));
    r.s = fma(-sq, h, fma(q, d, sq));
    r.c_half = fma(-B, 0.5 * d, fma(-A, 0.25 * h, A)) * kMisc[3];
    return r;
}

}  // namespace cm
}  // namespace birp
