// powf for CMEM, evaluated in fp64: x^y = 2^(y * log2 x) with a table-driven log2 and a short
// exp2, ~2^-47 relative accuracy (checked on the device against the CUDA fp64 library,
// tests/test_gpu_parity.py::test_cmem_math_accuracy).  CMEM makes 5-6 powf calls per sample; the
// general-purpose fp64 pow() costs ~250 instructions and exp2(y*log2(x)) through the library ~110,
// this version ~45 (Taylor coefficients are exact rationals; the only table is 32 x {1/c, -ln(1/c)}).
// The other transcendental calls of CMEM go through the CUDA fp64 library: hand-written Taylor
// versions of sincos / exp / acos were measured and were slower (more fp64 operations than the
// library's minimax polynomials) -- see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace birp {
namespace cm {

// ---- exp / exp2 --------------------------------------------------------------------------------
// e^r for |r| <= ln2/2, Taylor to r^11: truncation < 2^-47
__device__ __forceinline__ double expm_core(double r) {
    double p = 1.0 / 39916800.0;           // 1/11!
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    return fma(p, r, 1.0);
}

// multiply by 2^k for |k| <= 1000 (two steps keep every factor a normal number)
__device__ __forceinline__ double scale2(double v, int k) {
    const int k1 = k / 2, k2 = k - k1;
    const double f1 = __hiloint2double((1023 + k1) << 20, 0);
    const double f2 = __hiloint2double((1023 + k2) << 20, 0);
    return v * f1 * f2;
}

__device__ __forceinline__ double exp2_d(double t) {
    if (!(fabs(t) < 1000.0)) return exp2(t);
    const double kd = rint(t);
    const double r = (t - kd) * 6.93147180559945286227e-01;  // exact difference, |r| <= ln2/2
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
    double p = 1.0 / 7.0;
    p = fma(p, z, -1.0 / 6.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, -1.0 / 4.0);
    p = fma(p, z, 1.0 / 3.0);
    p = fma(p, z, -0.5);
    const double ln_m = fma(p * z, z, z) + t.y;
    return fma(ln_m, 1.44269504088896338700e+00, (double)e);
}

// x^y for finite x > 0 (other bases: the caller keeps powf's special cases through pow())
__device__ __forceinline__ double pow_pos_d(double x, double y) { return exp2_d(y * log2_d(x)); }

}  // namespace cm
}  // namespace birp
