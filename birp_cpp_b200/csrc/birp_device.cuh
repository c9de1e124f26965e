// Device-side building blocks of the line-of-sight integrator (sm_100a).
//
// Everything that decides WHICH voxel a sample reads, WHEN a ray terminates and in what
// order a pixel accumulates is written with explicitly rounded IEEE intrinsics
// (__fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn: never contracted into FMA), in the
// operation order of the reference, so the DataCube path is bit-identical to the
// reference CPU render (which contains no FMA instructions: x86-64 baseline build,
// reference CMakeLists.txt:18-22).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/birp_b200.h"
#include "birp_cmem_math.cuh"

namespace birp {

// ---------------------------------------------------------------------------------
// Device view of a resident cube (x fastest, like slices[z][y][x]).
// ---------------------------------------------------------------------------------
struct CubeView {
    const float* __restrict__ v;
    int nx, ny, nz;
    int slab;             // nx * ny
    long long slab16;     // 16 * nx * ny: byte distance between the z layers of the quad copy
    float fnx, fny, fnz;  // (float)size.{x,y,z}: right-hand side of the reference's bounds test
    float ox, oy, oz;     // coords_{x,y,z}[0]
    float sx, sy, sz;     // spacing[0..2]
    // ---- derived once at upload (birp_capi.cu: finish_cube) ----
    float rx, ry, rz;     // RN(1 / spacing), for div_spacing
    float lx, ly, lz;     // smallest ingress >= 0 that the reference's bounds test rejects
    float lox, loy, loz;  // smallest ingress whose quotient by the spacing is >= 1 (and >= 2^-100)
                          // (hot window: lo? <= ingress < l?, i.e. index in [1, n-1); everything else is cold)
    int divmode;          // cheapest div_spacing variant proven exact for all three axes (0,1) or 2
    int positive;         // 1: every voxel is finite and > 0 (std::lerp never takes its sign branch);
                          // 2: finite and > 0 or +0 (lerp_nonneg); 0: anything else (full std::lerp)
    // layered 2D texture over the same voxels (layer = z), point-sampled; 0 when the cube exceeds
    // the layered-texture limits or the upload-time gather self-check failed
    cudaTextureObject_t tex;
    // "quad" copy of the grid for the trilinear corner fetch, or nullptr:
    //   quad[(z*ny + y)*nx + x] = { v[z][y][x], v[z][y][x+1], v[z][y+1][x], v[z][y+1][x+1] }
    // (neighbours clamped at the last row / column, which no in-bounds sample reads): the four
    // corners of one z layer are ONE aligned 16-byte load.
    const float4* __restrict__ quad;
    // quad + nx * ny, the z + 1 layer of the same cell: set by the march kernel itself (both bases are then
    // plain registers and each corner address is one IMAD.WIDE)
    const float4* __restrict__ quad_up;
    int quad_bytes;  // 16, opaque to the compiler (likewise set by the kernel)
    // "oct" copy of the grid, or nullptr: 32-byte records, oct[(z*ny + y)*nx + x] = the quad of layer z followed by
    // the quad of layer z + 1 (clamped at the last layer like the quads at the last row / column): all eight
    // corners of a cell are ONE aligned 32-byte load (LDG.256), i.e. one L1 request per sample instead of two.
    // 8x the grid in device memory.
    const float* __restrict__ oct;
    // 1: the oct copy MARKS every record that is neither "all eight corners finite and > 0" nor "all eight corners +0"
    // by storing a NaN as its first corner (cubes outside the positive class): a sample in a marked record comes out
    // NaN, fails the scale's range test and takes the literal path; every other record interpolates exactly with the
    // three-operation lerp (LERP 3)
    int oct_marked;
    // smallest / largest voxel of a positive cube, found by scan_positive_kernel at upload (0, 0 otherwise); every
    // sample interpolated from a positive cube lies between them (lean march; host-side use only)
    float vmin, vmax;
    // "poly" copy of the grid for the tolerance-mode march (BIRP_FAST_MARCH), or nullptr: 32-byte records, the
    // trilinear interpolant of cell (x, y, z) as a polynomial in its three weights (see fast_sample); built at the
    // first BIRP_FAST_MARCH render of the cube.  8x the grid.
    const float* __restrict__ poly;
    int poly_bytes;  // 32, opaque to the compiler (set by the kernel: the multiplier of the record address stays in a register)
    int poly_bias;   // set by the kernel: (B * ny + B) * nx + B (mod 2^32) with B the bit pattern of 1.5 * 2^23 (fast_sample<0>)
    // 1 / spacing in double: the tolerance-mode march makes its per-ray index-space origin and direction with them
    double rx64, ry64, rz64;
};

// one 32-byte load through the read-only path (sm_100: LDG.E.NA.256.CONSTANT).
// L1::no_allocate: the records of a cell are read by a handful of neighbouring rays at about the same time and then
// never again, so the L1 hit rate is low (41 % on the orbit sweep) and every miss that ALLOCATES pays for its fill
// in L1 data-pipe wavefronts -- the unit that bounds the tolerance-mode march.  Without allocation the data goes
// from L2 to the registers: measured on the orbit sweep (profiles/r2c_ab_load_*.json) tolerance mode 23.6 -> 18.1 ms,
// bit-exact march 28.4 -> 27.6 ms; plain ld.global, evict_first and evict_last are all within 0.2 % of the default.
__device__ __forceinline__ void load_oct(const void* p, float& a0, float& a1, float& a2, float& a3, float& b0, float& b1,
                                         float& b2, float& b3) {
#ifdef __CUDACC__
#ifndef BIRP_OCT_LOAD
#define BIRP_OCT_LOAD "ld.global.nc.L1::no_allocate.v8.f32"  // (A/B builds override it)
#endif
    asm(BIRP_OCT_LOAD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
        : "l"(p));
#else
    const float* f = static_cast<const float*>(p);
    a0 = f[0], a1 = f[1], a2 = f[2], a3 = f[3], b0 = f[4], b1 = f[5], b2 = f[6], b3 = f[7];
#endif
}

// How does the oct record of these eight corners interpolate (cubes outside the positive class, CubeView::oct_marked)?
//   0  exactly with the three-operation lerp: all corners finite and > 0 (proof at cube_sample_hot), or all +0;
//   1  all corners finite and >= +0 but zeros next to positive values: std::lerp's sign branch gives RN((1-t) a) where
//      the three-operation form gives RN(a - RN(t a)) -- the record is SIGN-marked (its first corner, which is >= +0,
//      is stored negated) and the march interpolates it with lerp_nonneg, inside the hot loop;
//   2  anything else (a negative, -0, NaN or infinite corner): NaN-marked, the sample takes the literal path (where
//      a negative sample ends the ray).
__device__ __forceinline__ int oct_record_class(float a0, float a1, float a2, float a3, float b0, float b1, float b2,
                                                float b3) {
    const float lo = fminf(fminf(fminf(a0, a1), fminf(a2, a3)), fminf(fminf(b0, b1), fminf(b2, b3)));
    const float hi = fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)), fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)));
    const float sum = a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;  // NaN iff a corner is NaN (or +inf and -inf meet)
    const unsigned any_bits = __float_as_uint(a0) | __float_as_uint(a1) | __float_as_uint(a2) | __float_as_uint(a3) |
                              __float_as_uint(b0) | __float_as_uint(b1) | __float_as_uint(b2) | __float_as_uint(b3);
    const bool finite = (sum == sum) && hi < CUDART_INF_F && lo > -CUDART_INF_F;
    if (finite && lo > 0.f) return 0;
    if (any_bits == 0u) return 0;
    // no sign bit anywhere: every corner is +0 or positive
    if (finite && (any_bits & 0x80000000u) == 0u) return 1;
    return 2;
}

// The four voxels (x0..x0+1, y0..y0+1) of layer z in ONE texture instruction: tld4 returns the
// raw fp32 texels of the bilinear footprint, unfiltered (no 9-bit weights are involved -- the
// interpolation itself stays in fp32 registers, below).  Passing the footprint CENTRE
// (x0 + 1, y0 + 1 in unnormalised texel coordinates) keeps the footprint selection away from
// any rounding edge of the sampler's fixed-point coordinate conversion.
// Order: {T[y0+1][x0], T[y0+1][x0+1], T[y0][x0+1], T[y0][x0]}.
__device__ __forceinline__ void gather4(cudaTextureObject_t tex, int layer, float u, float v, float& t01,
                                        float& t11, float& t10, float& t00) {
#ifdef __CUDACC__
    asm("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %8}];"
        : "=f"(t01), "=f"(t11), "=f"(t10), "=f"(t00)
        : "l"(tex), "r"(layer), "f"(u), "f"(v), "f"(0.f));
#else  // host build of this header (tests/cpp/device_on_host.cpp): the provider of that build
    birp_host_gather4(tex, layer, u, v, t01, t11, t10, t00);
#endif
}


// Ingresses below this (other than the cold-path cases) are so close to the first grid plane
// that the residual of div_spacing would underflow; they take the generic sampler instead.
#define BIRP_TINY_INGRESS 7.888609052210118e-31f /* 2^-100 */

// ingress / spacing, correctly rounded.
//   MODE 2: IEEE division.
//   MODE 0/1: multiply by the correctly rounded reciprocal, then one / two residual corrections.
//   MODE 3: the multiply alone -- exact when the spacing is a power of two (0.5 R_E in the reference's own cubes):
//           the reciprocal is exact and scaling by a power of two rounds nothing in the normal range.
// MODE 0, 1 and 3 are only selected for a cube after an EXHAUSTIVE device-side comparison with
// __fdiv_rn over every fp32 ingress of the hot window (2^-100 <= a < limit), for each
// axis spacing (verify_div_kernel), so the quotient -- and with it the voxel index and the
// interpolation weight -- is bit-identical to the reference's `x_ingress / spacing[0]`.
template <int MODE>
__device__ __forceinline__ float div_spacing(float a, float s, float r) {
    if (MODE == 2) return __fdiv_rn(a, s);
    float q = __fmul_rn(a, r);
    if (MODE == 3) return q;
    float e = __fmaf_rn(-q, s, a);
    q = __fmaf_rn(e, r, q);
    if (MODE == 1) {
        e = __fmaf_rn(-q, s, a);
        q = __fmaf_rn(e, r, q);
    }
    return q;
}

// The fp32 -> fp64 -> fp32 scale of reference src/Camera.cpp:201,
//   sample = (sample * ray_width * (1000 * 100 * 6378.1) / (4 / M_PI)) / 10000;
// i.e. fl32( ((double)fl32(sample*ray_width) * 637810000.0 / (4/pi)) / 10000 ).
#define BIRP_SCALE_C1 (1000 * 100 * 6378.1)
#define BIRP_SCALE_C2 (4 / 3.14159265358979323846)
#define BIRP_SCALE_C3 10000.0

__device__ __forceinline__ float scale_sample_chain(float p) {
    double d = (double)p;
    d = __dmul_rn(d, BIRP_SCALE_C1);
    d = __ddiv_rn(d, BIRP_SCALE_C2);
    d = __ddiv_rn(d, BIRP_SCALE_C3);
    return __double2float_rn(d);
}

// Same result as scale_sample_chain for every input, cheaper on average.
// One fp64 multiply by K = C1/C2/C3 lands within a few fp64 ulps of the three-step
// chain (each step of either route is correctly rounded, so both stay within 4 ulp64
// of the real product).  The two can only round to different fp32 values when that
// fp64 result sits within a few ulp64 of an fp32 rounding boundary -- either a midpoint
// between consecutive fp32 values (low 29 mantissa bits ~ 0x10000000) or, because the
// chain's fp32 result may be subnormal/overflow, outside the normal fp32 range.  Those
// cases (about 1 in 2^24) take the exact chain.
__device__ __forceinline__ float scale_sample(float s, float ray_width) {
    const float p = __fmul_rn(s, ray_width);
    constexpr double K = BIRP_SCALE_C1 / BIRP_SCALE_C2 / BIRP_SCALE_C3;
    const double d = __dmul_rn((double)p, K);
    const float y = __double2float_rn(d);
    // distance of the 29 dropped mantissa bits from the midpoint pattern, window +-32 ulp64
    const unsigned frac = (unsigned)__double2loint(d) & 0x1fffffffu;
    const bool near_mid = (frac - (0x10000000u - 32u)) < 64u;
    // result comfortably inside the normal fp32 range (2^-120 <= |y| < 2^120); zero, subnormal,
    // overflowing and NaN products fail this test and take the literal chain
    const float ay = fabsf(y);
    const bool normal = (ay >= 7.52316384526264e-37f) && (ay < 1.329227995784916e+36f);
    if (__builtin_expect(normal && !near_mid, 1)) return y;
    if (p == 0.f) return p;  // (+-0) * C1 / C2 / C3 = +-0
    return scale_sample_chain(p);
}

// The hot-loop form of scale_sample: `y` is the single-multiply result and the return value says
// whether it is proven equal to the reference chain (same argument as above).  Negative, NaN,
// subnormal and overflowing results return false (no fabsf: a negative sample ends the ray, which
// the cold path handles); a zero product is its own exact result when MAYBE_ZERO (cubes that are
// not strictly positive -- for a positive cube a zero product can only be an underflow).
#define BIRP_SCALE_K (BIRP_SCALE_C1 / BIRP_SCALE_C2 / BIRP_SCALE_C3)
#define BIRP_SCALE_SAFE_LO 7.52316384526264e-37f  /* 2^-120 */
#define BIRP_SCALE_SAFE_HI 1.329227995784916e+36f /* 2^120 */
// RANGE_SAFE: the caller has shown on the host that every sample the hot path can produce scales into
// [2^-120, 2^120) (launch_render_cube: a positive cube's samples lie in [vmin, vmax], the scale is monotone), so
// the range test is dropped -- and with it the midpoint test: an enumeration of ALL 2^32 fp32 products
// (tests/test_device_on_host.py::test_scale_single_multiply_is_the_chain_for_every_product, 4 s on the CPU, same
// source) shows that for this K no product inside that range ever lands within 32 ulp64 of an fp32 midpoint, i.e.
// inside the range the single multiply IS the reference's chain for every input.  (The general hot loop keeps
// both tests: they cost two instructions and make the argument independent of the value of K.)
template <bool MAYBE_ZERO, bool RANGE_SAFE = false>
__device__ __forceinline__ bool scale_fast(float s, float ray_width, double K, float& y) {
    const float p = __fmul_rn(s, ray_width);
    const double d = __dmul_rn((double)p, K);
    y = __double2float_rn(d);
    // (frac - (2^28 - 32)) mod 2^29 < 64 with frac = the 29 dropped bits: times 8, the modulus is the word size and
    // the mask disappears (one multiply-add and one compare)
    const bool near_mid = ((unsigned)__double2loint(d) * 8u - ((0x10000000u - 32u) << 3)) < 512u;
    if (RANGE_SAFE) return true;
    bool exact = (y >= BIRP_SCALE_SAFE_LO) && (y < BIRP_SCALE_SAFE_HI);
    if (MAYBE_ZERO) exact = exact || (p == 0.f);  // (+-0) * C1 / C2 / C3 = +-0 = y
    return exact && !near_mid;
}

// libstdc++ 13 std::lerp (<cmath> __lerp), the interpolation the reference calls
// (src/DataCube.cpp:260-270, src/Cmem.cpp:149,170).
__device__ __forceinline__ float lerp_ref(float a, float b, float t) {
    if ((a <= 0.f && b >= 0.f) || (a >= 0.f && b <= 0.f))
        return __fadd_rn(__fmul_rn(t, b), __fmul_rn(__fsub_rn(1.f, t), a));
    if (t == 1.f) return b;
    const float x = __fadd_rn(a, __fmul_rn(t, __fsub_rn(b, a)));
    return ((t > 1.f) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
}

// Helper::VectorDistance, src/Helper.cpp:591-594
__device__ __forceinline__ float vec_len3(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// World-space unit direction of pixel (i, j): src/Camera.cpp:148-168 with
// Helper::ApplyRotation (src/Helper.cpp:136-148).
__device__ __forceinline__ void pixel_direction(const birp_frame& fr, int i, int j, float fW, float fH,
                                                float& dx, float& dy, float& dz) {
    const float device_x = __fdiv_rn(__fadd_rn((float)i, 0.5f), fW);
    const float device_y = __fdiv_rn(__fadd_rn((float)j, 0.5f), fH);
    const float screen_x =
        __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(2.f, device_x), 1.f), fr.tan_half_x), fr.aspect);
    const float screen_y = __fmul_rn(__fsub_rn(1.f, __fmul_rn(2.f, device_y)), fr.tan_half_y);
    const float nd = vec_len3(screen_x, screen_y, -1.f);
    const float ux = __fdiv_rn(screen_x, nd), uy = __fdiv_rn(screen_y, nd), uz = __fdiv_rn(-1.f, nd);
    const float* r = fr.rinv;
    const float wx = __fadd_rn(__fadd_rn(__fmul_rn(r[0], ux), __fmul_rn(r[1], uy)), __fmul_rn(r[2], uz));
    const float wy = __fadd_rn(__fadd_rn(__fmul_rn(r[3], ux), __fmul_rn(r[4], uy)), __fmul_rn(r[5], uz));
    const float wz = __fadd_rn(__fadd_rn(__fmul_rn(r[6], ux), __fmul_rn(r[7], uy)), __fmul_rn(r[8], uz));
    const float wd = vec_len3(wx, wy, wz);
    dx = __fdiv_rn(wx, wd);
    dy = __fdiv_rn(wy, wd);
    dz = __fdiv_rn(wz, wd);
}

// ---------------------------------------------------------------------------------
// DataCube::GetSample, src/DataCube.cpp:203-277.
// Returns -1 for the reference's out-of-bounds result.  *nan_index is set when an index
// is NaN after passing the bounds test (where the reference's .at() throws).
// ---------------------------------------------------------------------------------
template <bool TRILINEAR>
__device__ __forceinline__ float cube_sample(const CubeView& c, float x, float y, float z, bool& nan_index) {
    const float xi = __fsub_rn(x, c.ox);
    const float fx = __fdiv_rn(xi, c.sx);
    const float yi = __fsub_rn(y, c.oy);
    const float fy = __fdiv_rn(yi, c.sy);
    const float zi = __fsub_rn(z, c.oz);
    const float fz = __fdiv_rn(zi, c.sz);
    if ((0.f > xi || __fadd_rn(fx, 1.f) >= c.fnx) || (0.f > yi || __fadd_rn(fy, 1.f) >= c.fny) ||
        (0.f > zi || __fadd_rn(fz, 1.f) >= c.fnz))
        return -1.f;
    if (!(fx == fx) || !(fy == fy) || !(fz == fz)) {
        nan_index = true;
        return 0.f;
    }
    // indices are in [0, n-1): truncation == floor
    // (the cube has < 2^31 voxels, checked at upload: 32-bit index arithmetic)
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const int row = c.nx;
    const int base = (z0 * c.ny + y0) * row + x0;
    if (!TRILINEAR) return __ldg(c.v + base);

    const float tx = __fsub_rn(fx, (float)x0);
    const float ty = __fsub_rn(fy, (float)y0);
    const float tz = __fsub_rn(fz, (float)z0);
    // ceil == floor on an exact integer index: the "max" corner is then the same voxel
    const int ddx = (tx > 0.f) ? 1 : 0;
    const int ddy = (ty > 0.f) ? row : 0;
    const int ddz = (tz > 0.f) ? row * c.ny : 0;
    const float* p = c.v + base;
    const float s000 = __ldg(p);
    const float s001 = __ldg(p + ddx);
    const float s010 = __ldg(p + ddy);
    const float s011 = __ldg(p + ddy + ddx);
    const float s100 = __ldg(p + ddz);
    const float s101 = __ldg(p + ddz + ddx);
    const float s110 = __ldg(p + ddz + ddy);
    const float s111 = __ldg(p + ddz + ddy + ddx);
    const float low_x1 = lerp_ref(s000, s001, tx);
    const float low_x2 = lerp_ref(s010, s011, tx);
    const float high_x1 = lerp_ref(s100, s101, tx);
    const float high_x2 = lerp_ref(s110, s111, tx);
    const float low_y = lerp_ref(low_x1, low_x2, ty);
    const float high_y = lerp_ref(high_x1, high_x2, ty);
    return lerp_ref(low_y, high_y, tz);
}

// std::lerp for a NON-NEGATIVE finite cube (every voxel > 0 or +0: emissivity grids of real runs are zero
// inside the magnetosphere) and a weight t <= 1 - 2^-23, omt = RN(1 - t):
//   a, b > 0:  a + t*(b - a), the clamp never fires (proof at cube_sample_hot);
//   a == 0:    the sign branch of __lerp gives t*b + (1-t)*0 = RN(t*b), and so does 0 + t*(b - 0);
//   b == 0:    the sign branch gives t*0 + (1-t)*a = RN(omt * a)   -- the one case that needs its own formula.
// Results are again finite and >= +0, so the seven lerps of a trilinear sample compose.
__device__ __forceinline__ float lerp_nonneg(float a, float b, float t, float omt) {
    const float x = __fadd_rn(a, __fmul_rn(t, __fsub_rn(b, a)));
    const float y = __fmul_rn(omt, a);
    return b == 0.f ? y : x;
}

// Hot-path form of DataCube::GetSample for the march kernel.  Same results as cube_sample();
// what changes is the cost:
//  * the bounds test `0 > ingress || index + 1 >= size` is monotone in the ingress, so it is a
//    comparison against a precomputed limit (c.l?) -- no division for dead samples -- and the hot
//    window also leaves the FIRST cell of every axis (index < 1, c.lo?) to the cold path;
//  * the division by the spacing is div_spacing<DIVMODE>;
//  * the eight corners are read at fixed offsets (x0+1, y0+1, z0+1 are always inside the grid for
//    an in-bounds sample); on an exact integer index the reference reads the same voxel twice
//    (ceil == floor) -- with weight 0 that gives the same result for finite data (LERP 1, 2),
//    otherwise the duplicate is selected explicitly;
//  * for a positive cube (LERP 1) std::lerp is `a + t*(b-a)`, three instructions (proof below); for a
//    non-negative one (LERP 2) lerp_nonneg; anything else (LERP 0) runs the full __lerp.
// cube_hot_window() says whether a sample is inside the hot window (false: dead ray, NaN coordinate,
// first cell of an axis -- the caller falls back to cube_sample()), cube_sample_hot() evaluates it.
//
// Corner fetch (FETCH):
//   0  eight global loads from the linear grid
//   1  two texture gathers (tld4) on the layered copy
//   2  lower z layer: one tld4; upper z layer: four global loads
//   4  two 16-byte loads from the quad copy (one address computation each, fewer cache lines per warp request
//      than 4-byte loads)
//   5  one 32-byte load (LDG.256) from the oct copy (default: one L1 request per sample instead of two, three
//      instructions fewer; measured fastest, profiles/README.md)
// (A form 3, tld4 for the lower layer + one quad load for the upper, was measured 9 % slower than 4.)
// With the 4-byte forms the texture unit and the load/store unit each saturate their return path
// into the register file first (profiles/r1d, r1e, r1g), hence the mixed form 2.
//
// Why the three-instruction lerp is exact here.  std::lerp(a, b, t) for finite a, b > 0 and t < 1 is
//   x = a + t*(b - a);  return b > a ? min(x, b) : max(x, b)      (libstdc++ <cmath> __lerp)
// and the clamp can only matter if x overshoots b.  In the hot window the index f is >= 1, so the
// weight t = f - floor(f) is a multiple of 2^-23 below 1: t <= 1 - 2u with u = 2^-24.  Take b > a,
// D = b - a (real), d = RN(b - a) <= D(1 + u), and p = RN(t*d):
//   normal range:    p <= t*d*(1 + u) <= (1 - 2u)(1 + u)^2 D = (1 - 3u^2 - 2u^3) D < D;
//   subnormal t*d:   p <= t*d + 2^-150 < D(1 - u - 2u^2) + 2^-150 <= D when D >= 2^-126, and when
//                    D < 2^-126 the difference was exact (d = D) so p = RN(t*d) <= d = D.
// Hence a + p <= b in real arithmetic and, rounding being monotone and b representable,
// RN(a + p) <= b: the clamp never fires.  b < a is the mirror image, b == a gives x = a = b.
// (1) the window test on the ingresses xi = x - coords_x[0], ...
__device__ __forceinline__ bool cube_hot_window(const CubeView& c, float xi, float yi, float zi) {
    return (xi >= c.lox) && (xi < c.lx) && (yi >= c.loy) && (yi < c.ly) && (zi >= c.loz) && (zi < c.lz);
}

// (2) the sample, for ingresses inside the window
// FIDX (grids of at most 2^24 voxels, trilinear): the cell index is computed in fp32 -- floor of the three
// indices (FRND), two FFMA, ONE float->int conversion -- instead of three conversions each way and two IMAD.
// Every intermediate is an integer below 2^24, so the fp32 arithmetic is exact and the cell is the same.
template <bool TRILINEAR, int DIVMODE, int LERP, int FETCH, bool FIDX = false>
__device__ __forceinline__ float cube_sample_hot(const CubeView& c, float xi, float yi, float zi) {
    const float fx = div_spacing<DIVMODE>(xi, c.sx, c.rx);
    const float fy = div_spacing<DIVMODE>(yi, c.sy, c.ry);
    const float fz = div_spacing<DIVMODE>(zi, c.sz, c.rz);
    const int row = c.nx;
    const int slab = c.slab;
    // indices are in [1, n-1): truncation == floor == the reference's (int) cast
    int x0, y0, z0, cell;
    float gx, gy, gz;
    if (FIDX && TRILINEAR) {
        gx = truncf(fx), gy = truncf(fy), gz = truncf(fz);
        cell = (int)__fmaf_rn(__fmaf_rn(gz, c.fny, gy), c.fnx, gx);
        x0 = y0 = z0 = 0;  // (only the texture forms use them)
    } else {
        x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
        cell = (z0 * c.ny + y0) * row + x0;
        if (!TRILINEAR) return __ldg(c.v + cell);
        // One float->int conversion per axis (XU pipe) and the int->float back on the ALU pipe;
        // the fractions are exact.
        gx = (float)x0, gy = (float)y0, gz = (float)z0;
    }
    const float tx = __fsub_rn(fx, gx);
    const float ty = __fsub_rn(fy, gy);
    const float tz = __fsub_rn(fz, gz);
    float s000, l001, l010, l011, l100, l101, l110, l111;
    if (FETCH == 1) {
        const float u = gx + 1.0f, v = gy + 1.0f;  // exact: small integers
        gather4(c.tex, z0, u, v, l010, l011, l001, s000);
        gather4(c.tex, z0 + 1, u, v, l110, l111, l101, l100);
    } else if (FETCH == 2) {
        gather4(c.tex, z0, gx + 1.0f, gy + 1.0f, l010, l011, l001, s000);
        const float* __restrict__ p = c.v + slab + cell;
        l100 = __ldg(p);
        l101 = __ldg(p + 1);
        l110 = __ldg(p + row);
        l111 = __ldg(p + row + 1);
    } else if (FETCH == 5) {
        load_oct(reinterpret_cast<const char*>(c.oct) + (long long)cell * 32, s000, l001, l010, l011, l100, l101, l110,
                 l111);
    } else if (FETCH == 4) {
        // address = uniform base + cell * 16, the 16 held in a register: one IMAD.WIDE per layer
        const float4 lo = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(c.quad) + (long long)cell * c.quad_bytes));
        const float4 hi = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(c.quad_up) + (long long)cell * c.quad_bytes));
        s000 = lo.x, l001 = lo.y, l010 = lo.z, l011 = lo.w;
        l100 = hi.x, l101 = hi.y, l110 = hi.z, l111 = hi.w;
    } else {
        const float* __restrict__ p = c.v + cell;
        s000 = __ldg(p);
        l001 = __ldg(p + 1);
        l010 = __ldg(p + row);
        l011 = __ldg(p + row + 1);
        l100 = __ldg(p + slab);
        l101 = __ldg(p + slab + 1);
        l110 = __ldg(p + slab + row);
        l111 = __ldg(p + slab + row + 1);
    }
    if (LERP == 2) {
        const float ox = __fsub_rn(1.f, tx), oy = __fsub_rn(1.f, ty), oz = __fsub_rn(1.f, tz);
        const float low_x1 = lerp_nonneg(s000, l001, tx, ox);
        const float low_x2 = lerp_nonneg(l010, l011, tx, ox);
        const float high_x1 = lerp_nonneg(l100, l101, tx, ox);
        const float high_x2 = lerp_nonneg(l110, l111, tx, ox);
        const float low_y = lerp_nonneg(low_x1, low_x2, ty, oy);
        const float high_y = lerp_nonneg(high_x1, high_x2, ty, oy);
        return lerp_nonneg(low_y, high_y, tz, oz);
    }
    if (LERP == 3 && __float_as_int(s000) < 0) {
        // sign-marked record (zeros next to positive corners, oct_record_class 1): the non-negative form of std::lerp,
        // here in the hot loop -- a ray crosses the surface of a zero region at a step of its own, and a lane that left
        // the loop for it would fall out of step with its warp
        s000 = fabsf(s000);
        const float ox = __fsub_rn(1.f, tx), oy = __fsub_rn(1.f, ty), oz = __fsub_rn(1.f, tz);
        const float low_x1 = lerp_nonneg(s000, l001, tx, ox);
        const float low_x2 = lerp_nonneg(l010, l011, tx, ox);
        const float high_x1 = lerp_nonneg(l100, l101, tx, ox);
        const float high_x2 = lerp_nonneg(l110, l111, tx, ox);
        const float low_y = lerp_nonneg(low_x1, low_x2, ty, oy);
        const float high_y = lerp_nonneg(high_x1, high_x2, ty, oy);
        return lerp_nonneg(low_y, high_y, tz, oz);
    }
    if (LERP == 1 || LERP == 3) {
        // LERP 3 (cubes outside the positive class, oct records with marks): a record whose corners are all positive
        // interpolates exactly as in a positive cube; one whose corners are all +0 gives 0 + t * (0 - 0) = +0, which is
        // what std::lerp's sign branch t * 0 + (1 - t) * 0 gives; a NaN-marked record (oct_record_class 2) yields a NaN
        // sample, and the caller's range test sends it to the literal sampler
        const float low_x1 = __fadd_rn(s000, __fmul_rn(tx, __fsub_rn(l001, s000)));
        const float low_x2 = __fadd_rn(l010, __fmul_rn(tx, __fsub_rn(l011, l010)));
        const float high_x1 = __fadd_rn(l100, __fmul_rn(tx, __fsub_rn(l101, l100)));
        const float high_x2 = __fadd_rn(l110, __fmul_rn(tx, __fsub_rn(l111, l110)));
        const float low_y = __fadd_rn(low_x1, __fmul_rn(ty, __fsub_rn(low_x2, low_x1)));
        const float high_y = __fadd_rn(high_x1, __fmul_rn(ty, __fsub_rn(high_x2, high_x1)));
        return __fadd_rn(low_y, __fmul_rn(tz, __fsub_rn(high_y, low_y)));
    }
    const bool ux = tx > 0.f, uy = ty > 0.f, uz = tz > 0.f;
    const float s010 = uy ? l010 : s000;
    const float s100 = uz ? l100 : s000;
    const float s110 = uz ? (uy ? l110 : l100) : s010;
    const float s001 = ux ? l001 : s000;
    const float s011 = ux ? (uy ? l011 : l001) : s010;
    const float s101 = ux ? (uz ? l101 : l001) : s100;
    const float s111 = ux ? (uz ? (uy ? l111 : l101) : s011) : s110;
    const float low_x1 = lerp_ref(s000, s001, tx);
    const float low_x2 = lerp_ref(s010, s011, tx);
    const float high_x1 = lerp_ref(s100, s101, tx);
    const float high_x2 = lerp_ref(s110, s111, tx);
    const float low_y = lerp_ref(low_x1, low_x2, ty);
    const float high_y = lerp_ref(high_x1, high_x2, ty);
    return lerp_ref(low_y, high_y, tz);
}

// ---------------------------------------------------------------------------------
// The march of one ray through the cube (called by march_cube_kernel, birp_kernels.cu; compiled for the host
// as well by tests/cpp/device_on_host.cpp, where the whole control flow below runs against the oracle).
// ---------------------------------------------------------------------------------
constexpr int kSmemDist = 2048;  // ray-table entries kept in shared memory by the cube kernel (+ 1 of padding)

// bits in the device status word
constexpr int kStatusNanIndex = 1;  // BIRP_ERR_RANGE
constexpr int kStatusSingular = 2;  // BIRP_ERR_SINGULAR (device frame setup)
constexpr int kStatusPeerTimeout = 4;  // birp_peer_wait gave up (BIRP_ERR_CUDA)

// entry of the staged ray table at 32-bit shared-window address `sdist` (one LDS; the host build reads its own
// copy of the table at that byte offset)
__device__ __forceinline__ float table_entry(unsigned sdist) {
#ifdef __CUDACC__
    float tk;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(tk) : "r"(sdist));
    return tk;
#else
    return birp_host_table_entry(sdist);
#endif
}

struct RayState {
    float dx, dy, dz;  // unit direction
    float px, py, pz;  // camera position
    float acc;         // pixel value so far
};

// ingresses of the sample whose distance sits at shared address `sdist`, and whether it is hot
__device__ __forceinline__ bool step_ingress(const CubeView& c, const RayState& r, unsigned sdist, float& xi, float& yi,
                                             float& zi) {
    const float tk = table_entry(sdist);
    // Helper::MatrixScalarMultiply then "+ position" (src/Camera.cpp:183-189): two roundings,
    // then DataCube::GetSample's `x - coords_x[0]` (src/DataCube.cpp:205)
    xi = __fsub_rn(__fadd_rn(__fmul_rn(r.dx, tk), r.px), c.ox);
    yi = __fsub_rn(__fadd_rn(__fmul_rn(r.dy, tk), r.py), c.oy);
    zi = __fsub_rn(__fadd_rn(__fmul_rn(r.dz, tk), r.pz), c.oz);
    return cube_hot_window(c, xi, yi, zi);
}

// the ingresses alone (SPAN kernels: the window was settled for the whole run)
__device__ __forceinline__ void step_ingress_only(const CubeView& c, const RayState& r, unsigned sdist, float& xi,
                                                  float& yi, float& zi) {
    const float tk = table_entry(sdist);
    xi = __fsub_rn(__fadd_rn(__fmul_rn(r.dx, tk), r.px), c.ox);
    yi = __fsub_rn(__fadd_rn(__fmul_rn(r.dy, tk), r.py), c.oy);
    zi = __fsub_rn(__fadd_rn(__fmul_rn(r.dz, tk), r.pz), c.oz);
}

// March the steps whose distances sit in shared memory at byte addresses [sdist, send) (32-bit
// shared-window addresses, advanced by 4 per step: one LDS + one IADD, no per-step index
// arithmetic and no separate step counter).  Returns the address of the first step NOT
// accumulated (send when the ray is still alive); `alive` is cleared when the ray ended.
//
// Control flow: the inner loop is the hot path only -- a sample inside the hot window whose
// single-multiply scale is proven exact.  Anything else (ray leaving the cube, first cell of an
// axis, NaN, a scale product next to a rounding midpoint, ...) BREAKS out of it; the lane then
// waits at the loop exit for the rest of its warp, evaluates that one sample with the literal
// transcription (cube_sample + scale_sample) and, if the ray lives on, re-enters the hot loop.
// So the hot loop carries no reconvergence bookkeeping and no cold code; almost every break is a
// ray ending, and a lane that re-enters alone (about 1 warp in 1000) costs its warp at most one
// extra pass.
template <bool TRILINEAR, int DIVMODE, int LERP, int FETCH>
__device__ __forceinline__ unsigned march_steps(const CubeView& c, RayState& r, unsigned sdist, const unsigned send,
                                                float ray_width, double K, int* status, bool& alive) {
    for (;;) {
        // (an opaque statement of the OUTER loop only: keeps the compiler from merging the two loops
        // into one, which would put the cold code and its reconvergence barrier back into every step)
        asm volatile("" : "+r"(sdist));
        if (sdist == send) return sdist;
        float xi, yi, zi;
        if (step_ingress(c, r, sdist, xi, yi, zi)) {
            // Hot loop, rotated by hand: the body evaluates the sample at `sdist` (known to be in the
            // window) AND the window test of the next one, so a step has a single branch.  (The table
            // has one padding entry: the look-ahead of the last step reads it and discards the result.)
            bool more;
            do {
                const float s = cube_sample_hot<TRILINEAR, DIVMODE, LERP, FETCH>(c, xi, yi, zi);
                float v;
                const bool exact = scale_fast<LERP != 1>(s, ray_width, K, v);
                const bool hot_next = step_ingress(c, r, sdist + 4, xi, yi, zi);
                if (exact) {
                    r.acc = __fadd_rn(r.acc, v);
                    sdist += 4;
                }
                more = exact && hot_next && (sdist != send);
            } while (more);
            if (sdist == send) return sdist;
        }
        const float tk = table_entry(sdist);
        const float x = __fadd_rn(__fmul_rn(r.dx, tk), r.px);
        const float y = __fadd_rn(__fmul_rn(r.dy, tk), r.py);
        const float z = __fadd_rn(__fmul_rn(r.dz, tk), r.pz);
        bool nan_index = false;
        float s = cube_sample<TRILINEAR>(c, x, y, z, nan_index);
        if (nan_index) atomicOr(status, kStatusNanIndex);
        s = scale_sample(s, ray_width);
        if (s < 0.f) {  // src/Camera.cpp:207-210: the rest of the ray is ignored
            alive = false;
            return sdist;
        }
        r.acc = __fadd_rn(r.acc, s);
        sdist += 4;
    }
}

// The same march with the window test taken OUT of the step (SPAN kernels, the default).
//
// The ray table is monotone (checked on the host when the table is made: RayTable::monotone), and every
// operation between a table entry and an ingress -- RN(d * t), RN(. + p), RN(. - o) -- is a correctly
// rounded, hence monotone, function of t.  So each ingress is a monotone function of the step index, each
// axis is inside its window [lo, l) over a contiguous range of steps, and so is the conjunction of the three:
// the hot samples of a ray form ONE run of consecutive steps.  Once a step is found hot, the end of its run
// is located by bisection with the very same window test (step_ingress; about log2(nsamples) evaluations per
// ray instead of one per step), and the steps up to there are marched without any test: six compares and
// their predicate chain leave the loop body.  Results are those of march_steps bit for bit -- the window test
// only ever decided WHICH code evaluates a sample, and both codes return the reference's value.
// LEAN (BIRP_LEAN_MARCH, positive cubes of at most 2^24 voxels whose value range was checked on the host): the
// hot loop uses the fp32 cell index and the scale without its range test (cube_sample_hot<FIDX>,
// scale_fast<RANGE_SAFE>): four instructions fewer per sample.
template <bool TRILINEAR, int DIVMODE, int LERP, int FETCH, bool LEAN = false>
__device__ __forceinline__ unsigned march_steps_span(const CubeView& c, RayState& r, unsigned sdist, const unsigned send,
                                                     float ray_width, double K, int* status, bool& alive) {
    for (;;) {
        asm volatile("" : "+r"(sdist));  // (keeps the outer loop separate from the hot one, see march_steps)
        if (sdist == send) return sdist;
        float xi, yi, zi;
        if (step_ingress(c, r, sdist, xi, yi, zi)) {
            // end of the hot run that starts here: every step in [sdist, lo) is hot, `hi` is cold or the end
            unsigned lo = sdist + 4, hi = send;
            while (lo < hi) {
                const unsigned mid = lo + (((hi - lo) >> 3) << 2);
                float ux, uy, uz;
                if (step_ingress(c, r, mid, ux, uy, uz)) lo = mid + 4;
                else hi = mid;
            }
            const unsigned hend = lo;
            // (rotated like march_steps: the body evaluates the sample at `sdist` and the ingresses of the next)
            bool more;
            do {
                const float s = cube_sample_hot<TRILINEAR, DIVMODE, LERP, FETCH, LEAN>(c, xi, yi, zi);
                float v;
                const bool exact = scale_fast<LERP != 1, LEAN>(s, ray_width, K, v);  // (LERP 3: a NaN sample is not exact)
                step_ingress_only(c, r, sdist + 4, xi, yi, zi);
                if (exact) {
                    r.acc = __fadd_rn(r.acc, v);
                    sdist += 4;
                }
                more = exact && (sdist != hend);
            } while (more);
            if (sdist == send) return sdist;
        }
        const float tk = table_entry(sdist);
        const float x = __fadd_rn(__fmul_rn(r.dx, tk), r.px);
        const float y = __fadd_rn(__fmul_rn(r.dy, tk), r.py);
        const float z = __fadd_rn(__fmul_rn(r.dz, tk), r.pz);
        bool nan_index = false;
        float s = cube_sample<TRILINEAR>(c, x, y, z, nan_index);
        if (nan_index) atomicOr(status, kStatusNanIndex);
        s = scale_sample(s, ray_width);
        if (s < 0.f) {  // src/Camera.cpp:207-210: the rest of the ray is ignored
            alive = false;
            return sdist;
        }
        r.acc = __fadd_rn(r.acc, s);
        sdist += 4;
    }
}

// ---------------------------------------------------------------------------------
// Tolerance-mode march (BIRP_FAST_MARCH; trilinear only; opt-in -- the bit-exact march stays the default).
//
// north_star asks for 1e-5 per pixel and 1e-4 in the integrated flux, not for bit parity.  What bit parity costs
// is the arithmetic of the reference's un-fused x86 code: three rounded operations per coordinate, three
// corrected divisions, seven three-operation lerps and an fp64 scale with exactness tests per SAMPLE.  This march
// keeps everything that DECIDES something exact and relaxes only the value of a hot sample:
//
//  * which steps of a ray are inside the cube is decided with the reference's own arithmetic -- the same window
//    test and the same bisection as march_steps_span -- and every sample outside the hot run (cube exit, first
//    cell of an axis, NaN coordinates) goes through the literal transcription (cube_sample + scale_sample), which
//    also decides where the ray ends.  So a ray ends at the same step as in the reference;
//  * a hot sample whose cell has a negative or non-finite corner (only cubes outside the non-negative class have
//    such cells) is evaluated by the literal transcription too: the sign of its value ends the ray.  Those cells
//    are marked in their record (first coefficient NaN), the test is one compare on the interpolated value.  In
//    every other cell the reference's sample is >= 0 whatever the rounding (a lerp of non-negative numbers with
//    weights in [0, 1]), so the ray lives on and only the VALUE matters;
//  * the value: per axis, index - 1/2 = g + t_k * e with the per-ray index-space origin g = (p - o) / s - 1/2 and
//    direction e = d / s made once per ray in fp64 (one FFMA per axis and step instead of three roundings and a
//    division); i = rint(index - 1/2) is the cell, u = (index - 1/2) - i in [-1/2, 1/2] its CENTRED weight -- either
//    one FRND (XU pipe, a quarter of the FMA pipe's rate) or two FADD through the constant 1.5 * 2^23; on an
//    integer index the tie may pick the cell below with u = +1/2, which is the same point of the same continuous
//    interpolant; the cell's interpolant as a polynomial in the centred weights
//        s = C0 + C1 u + C2 v + C3 uv + w (C4 + C5 u + C6 v + C7 uv)
//    whose eight coefficients are ONE 32-byte record of the poly copy (7 FFMA instead of 21 operations), the raw
//    samples summed in fp32 and the scale (src/Camera.cpp:201) applied once per ray.
//
// Error against the reference: both sides place a sample within a few ulp of the coordinate magnitude (~1e-5 of a
// cell); pixels differ by ~1e-6 relative (tests/test_gpu_parity.py::test_fast_march_*: >= 99.9 % of pixels within
// 1e-5, flux within 1e-4; the same source on the CPU: tests/test_device_on_host.py::test_fast_march_*).
// ---------------------------------------------------------------------------------
#define BIRP_RINT_MAGIC 12582912.0f /* 1.5 * 2^23: x + M - M = rint(x) for |x| < 2^22 */

struct FastRay {
    float gx, gy, gz;  // index-space position of the camera minus 1/2: (p - o) / s - 1/2
    float ex, ey, ez;  // index-space direction: d / s (KSTEP: times the step length, d / s * maxdist / nsamples)
};

// `step`: 1 (the march reads the distance of every step from the ray table) or, KSTEP, the exact step length
// maxdist / nsamples of the uniform table of Camera::GenerateRayDistWidth (src/Camera.cpp:73:
// dist[k] = max * ((k + 0.5f) / ray_samples)), so that (k + 1/2) * e is the index-space offset of step k and the
// march needs no table read
__device__ __forceinline__ FastRay fast_ray(const CubeView& c, const RayState& r, double step) {
    FastRay q;
    q.gx = (float)(((double)r.px - (double)c.ox) * c.rx64 - 0.5);
    q.gy = (float)(((double)r.py - (double)c.oy) * c.ry64 - 0.5);
    q.gz = (float)(((double)r.pz - (double)c.oz) * c.rz64 - 0.5);
    q.ex = (float)((double)r.dx * c.rx64 * step);
    q.ey = (float)((double)r.dy * c.ry64 * step);
    q.ez = (float)((double)r.dz * c.rz64 * step);
    return q;
}

// Coefficients of one cell from its eight corners (pack_poly_kernel and the host harness), in fp64, rounded once:
// index bit 0 / 1 / 2 of a coefficient = it multiplies u / v / w.  A cell with a negative or non-finite corner gets
// C0 = NaN: every sample in it comes out NaN and takes the literal path.
__device__ __forceinline__ void poly_coefficients(float v000, float v001, float v010, float v011, float v100, float v101,
                                                  float v110, float v111, float c[8]) {
    const double a = v000, b = v001, d = v010, e = v011, f = v100, g = v101, h = v110, k = v111;
    c[0] = (float)((((a + b) + (d + e)) + ((f + g) + (h + k))) * 0.125);
    c[1] = (float)((((b - a) + (e - d)) + ((g - f) + (k - h))) * 0.25);
    c[2] = (float)((((d - a) + (e - b)) + ((h - f) + (k - g))) * 0.25);
    c[4] = (float)((((f - a) + (g - b)) + ((h - d) + (k - e))) * 0.25);
    c[3] = (float)((((e - d) - (b - a)) + ((k - h) - (g - f))) * 0.5);
    c[5] = (float)((((g - f) - (b - a)) + ((k - h) - (e - d))) * 0.5);
    c[6] = (float)((((h - f) - (d - a)) + ((k - g) - (e - b))) * 0.5);
    c[7] = (float)(((k - h) - (g - f)) - ((e - d) - (b - a)));
    const float lo = fminf(fminf(fminf(v000, v001), fminf(v010, v011)), fminf(fminf(v100, v101), fminf(v110, v111)));
    const float sum = fabsf(v000) + fabsf(v001) + fabsf(v010) + fabsf(v011) + fabsf(v100) + fabsf(v101) + fabsf(v110) + fabsf(v111);
    // (fminf drops NaNs, the sum of magnitudes keeps them: inf or NaN iff a corner is non-finite -- or the eight
    // magnitudes overflow together, which only marks a cell that was going to overflow anyway)
    if (!(lo >= 0.f) || !(sum <= 3.402823466e+38f)) c[0] = __int_as_float(0x7fc00000);
}

// cell index i = rint(f) and centred weight u = f - i of one axis.  FRND: one XU instruction and one FADD; else three
// FADD (no XU); *bits = the bit pattern of f + M, whose low mantissa bits are i (integer cell index without F2I)
template <bool FRND>
__device__ __forceinline__ void fast_axis(float f, float& i, float& u, unsigned& bits) {
    if (FRND) {
        i = rintf(f);
        bits = 0u;
    } else {
        const float m = __fadd_rn(f, BIRP_RINT_MAGIC);
        bits = __float_as_uint(m);
        i = __fsub_rn(m, BIRP_RINT_MAGIC);
    }
    u = __fsub_rn(f, i);
}

// Value of the hot sample whose three (index - 1/2) are fx, fy, fz.
//   VARIANT 0: rint by FADD pairs on all axes, INTEGER cell index from the bit patterns (no XU instruction at all;
//              the only form for grids of more than 2^24 voxels)
//   VARIANT 1: rint by FADD pairs, cell index in fp32 + one F2I            (1 XU instruction per sample)
//   VARIANT 2: z by FRND, x and y by FADD pairs, fp32 cell index + F2I     (2)
//   VARIANT 3: y and z by FRND                                              (3)
//   VARIANT 4: all three by FRND                                            (4: 26 instructions, but XU-bound)
// (the XU pipe issues a warp instruction every 8 cycles per scheduler; which mix is fastest is a measurement)
template <int VARIANT>
__device__ __forceinline__ float fast_sample(const CubeView& c, float fx, float fy, float fz) {
    float ix, iy, iz, u, v, w;
    unsigned bx, by, bz;
    fast_axis<(VARIANT >= 4)>(fx, ix, u, bx);
    fast_axis<(VARIANT >= 3)>(fy, iy, v, by);
    fast_axis<(VARIANT >= 2)>(fz, iz, w, bz);
    int cell;
    if (VARIANT == 0) {
        // (bits - B) is the integer index, B = bits(M); the B terms are folded into c.poly_bias by the kernel
        cell = (int)((bz * (unsigned)c.ny + by) * (unsigned)c.nx + bx) - c.poly_bias;
    } else {
        cell = (int)__fmaf_rn(__fmaf_rn(iz, c.fny, iy), c.fnx, ix);
    }
    float c0, c1, c2, c3, c4, c5, c6, c7;
    load_oct(reinterpret_cast<const char*>(c.poly) + (long long)cell * c.poly_bytes, c0, c1, c2, c3, c4, c5, c6, c7);
    const float hi = __fmaf_rn(v, __fmaf_rn(u, c7, c6), __fmaf_rn(u, c5, c4));
    const float lo = __fmaf_rn(v, __fmaf_rn(u, c3, c2), __fmaf_rn(u, c1, c0));
    return __fmaf_rn(w, hi, lo);
}

// the literal sample of one step (cold path of the tolerance-mode march): returns false when the ray ends here
__device__ __forceinline__ bool fast_cold_step(const CubeView& c, RayState& r, unsigned sdist, float ray_width,
                                               int* status) {
    const float tk = table_entry(sdist);
    const float x = __fadd_rn(__fmul_rn(r.dx, tk), r.px);
    const float y = __fadd_rn(__fmul_rn(r.dy, tk), r.py);
    const float z = __fadd_rn(__fmul_rn(r.dz, tk), r.pz);
    bool nan_index = false;
    const float s = cube_sample<true>(c, x, y, z, nan_index);
    if (nan_index) atomicOr(status, kStatusNanIndex);
    if (scale_sample(s, ray_width) < 0.f) return false;  // src/Camera.cpp:207-210, decided on the reference's value
    r.acc = __fadd_rn(r.acc, s);                         // (r.acc is the sum of RAW samples in this march)
    return true;
}

// SIGNS: the cube has cells with negative / non-finite corners (class 0): test every hot sample for the NaN mark.
// KSTEP: the distance of step k is (k + 1/2) * maxdist / nsamples (folded into q.e by fast_ray) instead of a read of
// the staged table -- one LDS less per sample on the L1 data path, which is what bounds this march; `kfirst` is the
// step index of shared address `sbase` (the chunk of the table that is staged).
template <bool SIGNS, int VARIANT, bool KSTEP>
__device__ __forceinline__ unsigned march_steps_fast(const CubeView& c, RayState& r, const FastRay& q, unsigned sdist,
                                                     const unsigned send, float ray_width, int* status, bool& alive,
                                                     unsigned sbase = 0u, int kfirst = 0) {
    for (;;) {
        asm volatile("" : "+r"(sdist));  // (keeps the outer loop separate from the hot one, see march_steps)
        if (sdist == send) return sdist;
        float xi, yi, zi;
        if (step_ingress(c, r, sdist, xi, yi, zi)) {
            // end of the hot run that starts here, by the reference's own window test (see march_steps_span)
            unsigned lo = sdist + 4, hi = send;
            while (lo < hi) {
                const unsigned mid = lo + (((hi - lo) >> 3) << 2);
                float ux, uy, uz;
                if (step_ingress(c, r, mid, ux, uy, uz)) lo = mid + 4;
                else hi = mid;
            }
            const unsigned hend = lo;
            for (;;) {
                asm volatile("" : "+r"(sdist));
                bool marked = false;
                if (KSTEP) {
                    // k + 1/2 as a float (exact below 2^23 steps), advanced by 1 per step
                    const float k0 = __fadd_rn((float)(kfirst + (int)((sdist - sbase) >> 2)), 0.5f);
                    const float kend = __fadd_rn(k0, (float)((hend - sdist) >> 2));
                    float kf = k0;
                    do {
                        const float s = fast_sample<VARIANT>(c, __fmaf_rn(kf, q.ex, q.gx), __fmaf_rn(kf, q.ey, q.gy),
                                                             __fmaf_rn(kf, q.ez, q.gz));
                        if (SIGNS && !(s >= 0.f)) {
                            marked = true;
                            break;
                        }
                        r.acc = __fadd_rn(r.acc, s);
                        kf = __fadd_rn(kf, 1.0f);
                    } while (kf != kend);
                    sdist += 4u * (unsigned)(int)__fsub_rn(kf, k0);
                } else {
                    do {
                        const float tk = table_entry(sdist);
                        const float s = fast_sample<VARIANT>(c, __fmaf_rn(tk, q.ex, q.gx), __fmaf_rn(tk, q.ey, q.gy),
                                                             __fmaf_rn(tk, q.ez, q.gz));
                        if (SIGNS && !(s >= 0.f)) {
                            marked = true;
                            break;
                        }
                        r.acc = __fadd_rn(r.acc, s);
                        sdist += 4;
                    } while (sdist != hend);
                }
                if (!SIGNS || !marked) break;
                // a cell with a negative / non-finite corner (or a rounding-level negative next to a zero corner):
                // the literal sample decides
                if (!fast_cold_step(c, r, sdist, ray_width, status)) {
                    alive = false;
                    return sdist;
                }
                sdist += 4;
                if (sdist == hend) break;
            }
            if (sdist == send) return sdist;
        }
        if (!fast_cold_step(c, r, sdist, ray_width, status)) {
            alive = false;
            return sdist;
        }
        sdist += 4;
    }
}

// ---------------------------------------------------------------------------------
// CMEM::GetSample, src/Cmem.cpp:117-186, with ShueCoords :224-246, LinScaled :188-222
// and ShueModel :280-288.
//
// The reference evaluates fp32 glibc functions (correctly rounded in the vast majority
// of cases).  PRECISE = true evaluates every transcendental in fp64 and rounds once to
// fp32, which reproduces a correctly rounded fp32 result except within ~1e-8 of a
// rounding boundary; PRECISE = false calls CUDA's fp32 libdevice functions (1-4 ulp).
// The algebra between the calls keeps the reference's types and operation order.
// ---------------------------------------------------------------------------------
template <bool PRECISE>
struct M {
    static __device__ __forceinline__ float sin_(float x) { return PRECISE ? (float)sin((double)x) : sinf(x); }
    static __device__ __forceinline__ float cos_(float x) { return PRECISE ? (float)cos((double)x) : cosf(x); }
    // sinf(x), cosf(x) and sinf(2x) together: 2x is exact in fp32, and sin(2x) = 2 sin(x) cos(x) holds for
    // the real numbers, so the fp64 product rounds to the same fp32 value as a separate evaluation would
    static __device__ __forceinline__ void sincos_(float x, float& s, float& c, float& s2) {
        if (PRECISE) {
            double ds, dc;
            sincos((double)x, &ds, &dc);
            s = (float)ds;
            c = (float)dc;
            s2 = (float)(2.0 * ds * dc);
        } else {
            sincosf(x, &s, &c);
            s2 = sinf(__fmul_rn(2.f, x));
        }
    }
    static __device__ __forceinline__ float acos_(float x) {
        return PRECISE ? (float)acos((double)x) : acosf(x);
    }
    static __device__ __forceinline__ float exp_(float x) { return PRECISE ? (float)exp((double)x) : expf(x); }
    // powf.  For a finite positive base, x^y = 2^(y * log2 x) in fp64, within ~2^-44 relative of the true
    // power for |y * log2 x| < 2^8 -- the same "round an almost exact value once" recipe as the other
    // calls.  Zero, negative, infinite and NaN bases keep powf's special-case semantics through pow().
    static __device__ __forceinline__ float pow_(float x, float y) {
        if (!PRECISE) return powf(x, y);
        if (x > 0.f && x < CUDART_INF_F) return (float)cm::pow_pos_d((double)x, (double)y);
        return (float)pow((double)x, (double)y);
    }
    // the two calls the reference itself makes in fp64 (src/Cmem.cpp:203-210)
    static __device__ __forceinline__ double cos64(double x) { return cos(x); }
    static __device__ __forceinline__ double acos64(double x) { return acos(x); }
};

// CMEM parameters plus everything that does not depend on the sample, evaluated once on the
// host with glibc (make_cmem_dev in birp_kernels.cu) -- the reference recomputes these per sample
// with the same glibc calls, so hoisting them is exact.
struct CmemDev {
    birp_cmem_params P;
    float cos_tn, sin_tn, cos_ts, sin_ts;  // cosf/sinf of theta_n, theta_s (src/Cmem.cpp:204-209)
    float p0r0;                            // p0 * r0_lin            (:219)
    float p2c;                             // p2 * charlie           (:217)
    float p3dn, p3ds;                      // p3 * dn, p3 * ds       (:217)
    int same_flare;                        // ay_bs == az_bs: ShueModel's two powf calls coincide
    int q_nonpositive;                     // p2 * charlie <= 0: the cusp term of LinScaled can only lower the radius
    int e_one;                             // expf(-B (theta/2)^4) rounds to 1.0f for EVERY theta in [0, pi] (see cmem_b1)
    // Region table of this parameter set (cmem_table_cell; nullptr: none): kCmemTableT bins over cos(theta) x
    // kCmemTableP bins over cos(phi), each {mp_lo, mp_hi, bs_lo, bs_hi}: a sample whose direction falls into the cell
    // and whose r is < mp_lo lies inside the magnetopause layer, with r > mp_hi outside it, with r < bs_lo inside the
    // bow-shock layer, with r > bs_hi beyond it -- bounds over the cell and its eight neighbours, layer half-thickness
    // and estimate margins included, so that a comparison of the ESTIMATED r with them is safe.  A cell near which
    // the fp32 estimates are not trusted (next to a cusp axis, where the Lin radius has a needle; next to the -x
    // axis) bounds nothing.
    const float4* table;
};

constexpr int kCmemTableT = 256, kCmemTableP = 64;

__device__ __forceinline__ int cmem_table_index(float q_t, float q_p) {
    int bt = (int)((q_t + 1.f) * (0.5f * kCmemTableT));
    int bp = (int)((q_p + 1.f) * (0.5f * kCmemTableP));
    bt = bt < 0 ? 0 : (bt > kCmemTableT - 1 ? kCmemTableT - 1 : bt);
    bp = bp < 0 ? 0 : (bp > kCmemTableP - 1 ? kCmemTableP - 1 : bp);
    return bt * kCmemTableP + bp;
}

// Sample-independent CMEM terms, with the host libm the reference itself calls.
inline CmemDev make_cmem_dev(const birp_cmem_params& P) {
    CmemDev D;
    D.P = P;
    D.cos_tn = cosf(P.theta_n);
    D.sin_tn = sinf(P.theta_n);
    D.cos_ts = cosf(P.theta_s);
    D.sin_ts = sinf(P.theta_s);
    const volatile float p0r0 = P.p0 * P.r0_lin, p2c = P.p2 * P.charlie, p3dn = P.p3 * P.dn, p3ds = P.p3 * P.ds;
    D.p0r0 = p0r0;
    D.p2c = p2c;
    D.p3dn = p3dn;
    D.p3ds = p3ds;
    D.same_flare = (P.ay_bs == P.az_bs) ? 1 : 0;
    D.q_nonpositive = (p2c <= 0.f) ? 1 : 0;
    // theta = RN32(acos q) <= RN32(pi), so (theta/2)^4 with its three roundings stays below 6.0881: when |B| * 6.0882 is
    // below 2^-26 the argument of the expf of src/Cmem.cpp:157 is too, for every sample, and the exponential is 1.0f
    D.e_one = ((double)fabsf(P.B) * 6.0882 < 1.4901161e-08) ? 1 : 0;
    D.table = nullptr;
    return D;
}

template <bool PRECISE>
__device__ __forceinline__ float cmem_sample(const CmemDev& D, float x, float y, float z) {
    typedef M<PRECISE> m;
    const birp_cmem_params& P = D.P;
    // ShueCoords
    const float r = vec_len3(x, y, z);
    const float theta = (x != 0.f) ? m::acos_(__fdiv_rn(x, r)) : 0.f;
    const float phi =
        (y != 0.f) ? m::acos_(__fdiv_rn(y, __fsqrt_rn(__fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))))) : 0.f;

    float sin_t, cos_t, sin_2t, sin_p, cos_p, unused;
    m::sincos_(theta, sin_t, cos_t, sin_2t);
    m::sincos_(phi, sin_p, cos_p, unused);

    // LinScaled: the cos(phi - pi/2) factors and the outer acos are fp64 in the reference
    // cos(phi - 3pi/2) = -cos(phi - pi/2); the reference evaluates both in fp64 (they agree to ~1e-16)
    const double c_n = m::cos64((double)phi - CUDART_PI / 2);
    const double c_s = PRECISE ? -c_n : m::cos64((double)phi - (3 * CUDART_PI / 2));
    const float cos_tn = D.cos_tn, sin_tn = D.sin_tn, cos_ts = D.cos_ts, sin_ts = D.sin_ts;
    const float phi_n = (float)m::acos64(
        __dadd_rn((double)__fmul_rn(cos_t, cos_tn), __dmul_rn((double)__fmul_rn(sin_t, sin_tn), c_n)));
    const float phi_s = (float)m::acos64(
        __dadd_rn((double)__fmul_rn(cos_t, cos_ts), __dmul_rn((double)__fmul_rn(sin_t, sin_ts), c_s)));

    const float fbase = __fadd_rn(
        m::cos_(__fdiv_rn(theta, 2.f)),
        __fmul_rn(__fmul_rn(P.a5, sin_2t), __fsub_rn(1.f, m::exp_(-theta))));
    const float fexp = __fmul_rn(
        P.p1, __fadd_rn(__fadd_rn(P.beta_c[0], __fmul_rn(P.beta_c[1], cos_p)),
                        __fmul_rn(P.beta_c[3], __fmul_rn(sin_p, sin_p))));
    const float f = m::pow_(fbase, fexp);
    const float p2c = D.p2c;
    const float q = __fadd_rn(__fmul_rn(p2c, m::exp_(__fmul_rn(D.p3dn, m::pow_(phi_n, P.a21)))),
                              __fmul_rn(p2c, m::exp_(__fmul_rn(D.p3ds, m::pow_(phi_s, P.a21)))));
    const float radius_mp = __fadd_rn(__fmul_rn(D.p0r0, f), q);

    // ShueModel(theta, phi, bs, ay_bs, az_bs)
    const float sbase = __fdiv_rn(2.f, __fadd_rn(1.f, cos_t));
    const float ry = __fmul_rn(P.bs, m::pow_(sbase, P.ay_bs));
    const float rz = D.same_flare ? ry : __fmul_rn(P.bs, m::pow_(sbase, P.az_bs));
    const float rzc = __fmul_rn(rz, cos_p);
    const float rys = __fmul_rn(ry, sin_p);
    const float radius_bs =
        __fdiv_rn(__fmul_rn(ry, rz), __fsqrt_rn(__fadd_rn(__fmul_rn(rzc, rzc), __fmul_rn(rys, rys))));

    const float half_thickness = 0.025f;  // 0.05f / 2
    const float mp_lo = __fsub_rn(radius_mp, half_thickness), mp_hi = __fadd_rn(radius_mp, half_thickness);
    const float bs_lo = __fsub_rn(radius_bs, half_thickness), bs_hi = __fadd_rn(radius_bs, half_thickness);

    if (r < mp_lo) return 0.f;

    const bool in_mp_layer = (r >= mp_lo && r <= mp_hi);
    const bool in_sheath = (r > mp_hi && r < bs_lo);
    const bool in_bs_layer = (r >= bs_lo && r <= bs_hi);

    float b1 = 0.f;
    if (in_mp_layer || in_sheath || in_bs_layer) {
        const float h = __fdiv_rn(theta, 2.f);
        const float h4 = __fmul_rn(__fmul_rn(__fmul_rn(h, h), h), h);
        const float e = m::exp_(__fmul_rn(-P.B, h4));
        const float pw = m::pow_(__fdiv_rn(r, 10.f),
                                 __fsub_rn(-P.alpha, __fmul_rn(P.beta, __fmul_rn(sin_t, sin_t))));
        b1 = __fmul_rn(__fmul_rn(P.A1, e), pw);
    }
    if (in_mp_layer) {
        const float shue_norm = __fdiv_rn(__fsub_rn(r, mp_lo), mp_hi);
        return lerp_ref(0.f, b1, shue_norm);
    }
    if (in_sheath) return b1;

    // beyond the magnetosheath: b2 is computed in double from the fp32 quotient r/10
    const double sd = (double)__fdiv_rn(r, 10.f);
    const double cube = __dmul_rn(__dmul_rn(sd, sd), sd);
    const float b2 = __double2float_rn(__dmul_rn((double)P.A2, __ddiv_rn(1.0, cube)));
    if (in_bs_layer) {
        const float shue_norm = __fdiv_rn(__fsub_rn(r, bs_lo), bs_hi);
        return lerp_ref(b1, b2, shue_norm);
    }
    return b2;
}

// ---------------------------------------------------------------------------------
// CMEM::GetSample once more, the default: the same "fp64, rounded once to fp32" evaluation of every
// libm call as cmem_sample<true>, with the building blocks of birp_cmem_math.cuh instead of the CUDA
// fp64 library -- coefficients read from the constant bank, and the sines / cosines of theta and phi
// taken from the square roots the two acos evaluations need anyway (cm::angle_from_cosine).  Types
// and operation order between the calls are the reference's, as above.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float pow32(float x, float y) {
    if (x > 0.f && x < CUDART_INF_F) return (float)cm::pow_pos_d((double)x, (double)y);
    return (float)cm::pow_slow((double)x, (double)y);  // zero, negative, infinite, NaN bases: powf's special cases
}

// the cosines ShueCoords takes its angles from: theta = acosf(q_t) (or 0 when x == 0: acos(1) = 0 gives the
// same angle), phi = acosf(q_p) (or 0 when y == 0); src/Cmem.cpp:224-246
__device__ __forceinline__ void cmem_cosines(float x, float y, float z, float& r, float& q_t, float& q_p) {
    r = vec_len3(x, y, z);
    q_t = (x != 0.f) ? __fdiv_rn(x, r) : 1.f;
    q_p = (y != 0.f) ? __fdiv_rn(y, __fsqrt_rn(__fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z)))) : 1.f;
}

// r / 10.f (src/Cmem.cpp:158,165) without the division: q0 = r * RN(1/10), the exact remainder r - 10 q0 (one FMA), one
// correction step.  Compared with r / 10.f for EVERY fp32 r (birp_selftest_div10 on the device, and on the CPU):
// identical for all r >= 2^-125 -- the 419 431 failures all lie in [2^-126, 2^-125), where the quotient is subnormal;
// cmem_r_is_ordinary keeps it to [2^-100, 2^100).
__device__ __forceinline__ float div10(float r) {
    const float q0 = __fmul_rn(r, 0.1f);
    return __fmaf_rn(__fmaf_rn(-q0, 10.f, r), 0.1f, q0);
}
// 2^-100 <= r < 2^100 (false for zero, negative, infinite, NaN): one integer compare
__device__ __forceinline__ bool cmem_r_is_ordinary(float r) {
    return (__float_as_uint(r) - 0x0D800000u) < 0x64000000u;
}

// 1 / c in fp64 to within an ulp or so: hardware seed (MUFU.RCP64H, ~20 bits) and three Newton steps, 7 instructions
// where the IEEE division takes ~30 (seed, iterations, range fix-up).  Used where the quotient is rounded to fp32
// right away: the fp32 result differs from the correctly rounded division's only when the fp64 value lies within an
// ulp64 of an fp32 rounding boundary (about one sample in 2^27).
__device__ __forceinline__ double rcp_newton(double c) {
#ifdef __CUDACC__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c));
    y = fma(y, fma(-c, y, 1.0), y);
    y = fma(y, fma(-c, y, 1.0), y);
    return fma(y, fma(-c, y, 1.0), y);
#else
    return 1.0 / c;
#endif
}

// expf(-B h^4) of src/Cmem.cpp:157.  CMEM::Init overwrites B with |b| * 1e-9 (:310), so the argument is ~ -1e-8 and the
// correctly rounded result is 1.0f whenever |argument| < 2^-26 (e^x is then within a quarter ulp of 1).
__device__ __forceinline__ float cmem_b1_exp(float B, float theta) {
    const float h = __fdiv_rn(theta, 2.f);
    const float h4 = __fmul_rn(__fmul_rn(__fmul_rn(h, h), h), h);
    const float ea = __fmul_rn(-B, h4);
    return (fabsf(ea) < 1.4901161e-08f) ? 1.0f : (float)cm::exp_d((double)ea);
}

// b1 and b2 (src/Cmem.cpp:155-173) for an r that is zero, tiny, huge, infinite or NaN: the literal forms, out of line
// (no sample of a real render comes here; scalars by value, so that the parameter block need not be copied to the stack)
#ifdef __CUDACC__
#define BIRP_NOINLINE __noinline__
#else
#define BIRP_NOINLINE
#endif
static __device__ BIRP_NOINLINE float cmem_b1_odd_r(float A1, float B, float alpha, float beta, float r, float theta,
                                                    float sin_t) {
    const float pw = pow32(__fdiv_rn(r, 10.f), __fsub_rn(-alpha, __fmul_rn(beta, __fmul_rn(sin_t, sin_t))));
    return __fmul_rn(__fmul_rn(A1, cmem_b1_exp(B, theta)), pw);
}
static __device__ BIRP_NOINLINE float cmem_b2_odd_r(float A2, float r) {
    const double sd = (double)__fdiv_rn(r, 10.f);
    return __double2float_rn(__dmul_rn((double)A2, __ddiv_rn(1.0, __dmul_rn(__dmul_rn(sd, sd), sd))));
}

// b1 of src/Cmem.cpp:155-160: the magnetosheath emissivity, from the rounded theta and its fp32 sine
__device__ __forceinline__ float cmem_b1(const CmemDev& D, float r, float theta, float sin_t) {
    const birp_cmem_params& P = D.P;
    if (!cmem_r_is_ordinary(r)) return cmem_b1_odd_r(P.A1, P.B, P.alpha, P.beta, r, theta, sin_t);
    // D.e_one: the exponential is 1.0f for every theta (known once per parameter set); a NaN theta reaches the result
    // through sin_t all the same
    const float e = D.e_one ? 1.0f : cmem_b1_exp(P.B, theta);
    const float ex = __fsub_rn(-P.alpha, __fmul_rn(P.beta, __fmul_rn(sin_t, sin_t)));
    // (r / 10 of an ordinary r is finite and positive: none of powf's special cases)
    const float pw = (float)cm::pow_pos_d((double)div10(r), (double)ex);
    return __fmul_rn(__fmul_rn(P.A1, e), pw);
}

// b2 of src/Cmem.cpp:162-173: beyond the magnetosheath, computed in double from the fp32 quotient r/10
__device__ __forceinline__ float cmem_b2(const birp_cmem_params& P, float r) {
    if (!cmem_r_is_ordinary(r)) return cmem_b2_odd_r(P.A2, r);
    const double sd = (double)div10(r);  // (its cube neither overflows nor underflows in fp64)
    return __double2float_rn(__dmul_rn((double)P.A2, rcp_newton(__dmul_rn(__dmul_rn(sd, sd), sd))));
}

// ShueModel(theta, phi, bs, ay_bs, az_bs) of src/Cmem.cpp:280-288: the bow-shock radius in the sample's direction
__device__ __forceinline__ float cmem_radius_bs(const CmemDev& D, float cos_t, float sin_p, float cos_p) {
    const birp_cmem_params& P = D.P;
    const float sbase = __fdiv_rn(2.f, __fadd_rn(1.f, cos_t));
    const float ry = __fmul_rn(P.bs, pow32(sbase, P.ay_bs));
    const float rz = D.same_flare ? ry : __fmul_rn(P.bs, pow32(sbase, P.az_bs));
    const float rzc = __fmul_rn(rz, cos_p);
    const float rys = __fmul_rn(ry, sin_p);
    return __fdiv_rn(__fmul_rn(ry, rz), __fsqrt_rn(__fadd_rn(__fmul_rn(rzc, rzc), __fmul_rn(rys, rys))));
}

// magnetopause (LinScaled, :188-222) and bow-shock (ShueModel, :280-288) radii in the sample's direction
// above_mp: the sample lies ABOVE the magnetopause layer for certain (r > mp_hi: the region filter knows, see
// cmem_try_filter) -- the Lin radius, two thirds of this function, decides nothing there and is not evaluated: a radius
// of -inf makes every comparison of cmem_regions come out as it would with the true one (r < mp_lo false, r <= mp_hi
// false, r > mp_hi true).
__device__ __forceinline__ void cmem_radii_v4(const CmemDev& D, float q_t, float q_p, float& theta, float& sin_t,
                                              float& radius_mp, float& radius_bs, bool above_mp = false) {
    const birp_cmem_params& P = D.P;
    const cm::Angle T = cm::angle_from_cosine(q_t);
    const cm::Angle F = cm::angle_from_cosine(q_p);
    theta = T.a;
    sin_t = (float)T.s;
    const float cos_t = (float)T.c;
    // sinf(2 theta): 2 theta is exact in fp32 and sin 2a = 2 sin a cos a
    const float sin_2t = (float)(2.0 * T.s * T.c);
    const float sin_p = (float)F.s, cos_p = (float)F.c;
    radius_bs = cmem_radius_bs(D, cos_t, sin_p, cos_p);
    if (above_mp) {
        radius_mp = -CUDART_INF_F;
        return;
    }

    // cos(phi - pi/2) = sin(phi) and cos(phi - 3pi/2) = -sin(phi): the reference evaluates both in fp64
    // (they agree with these to ~1e-16), as well as the outer acos
    const double c_n = F.s, c_s = -F.s;
    const float phi_n = (float)cm::acos_d(
        __dadd_rn((double)__fmul_rn(cos_t, D.cos_tn), __dmul_rn((double)__fmul_rn(sin_t, D.sin_tn), c_n)));
    const float phi_s = (float)cm::acos_d(
        __dadd_rn((double)__fmul_rn(cos_t, D.cos_ts), __dmul_rn((double)__fmul_rn(sin_t, D.sin_ts), c_s)));

    const float fbase = __fadd_rn((float)T.c_half,  // cosf(theta / 2): theta / 2 is exact in fp32
                                  __fmul_rn(__fmul_rn(P.a5, sin_2t), __fsub_rn(1.f, (float)cm::exp_d(-(double)theta))));
    const float fexp = __fmul_rn(
        P.p1, __fadd_rn(__fadd_rn(P.beta_c[0], __fmul_rn(P.beta_c[1], cos_p)),
                        __fmul_rn(P.beta_c[3], __fmul_rn(sin_p, sin_p))));
    const float f = pow32(fbase, fexp);
    const float p2c = D.p2c;
    const float q = __fadd_rn(__fmul_rn(p2c, (float)cm::exp_d((double)__fmul_rn(D.p3dn, pow32(phi_n, P.a21)))),
                              __fmul_rn(p2c, (float)cm::exp_d((double)__fmul_rn(D.p3ds, pow32(phi_s, P.a21)))));
    radius_mp = __fadd_rn(__fmul_rn(D.p0r0, f), q);
}

#define BIRP_CMEM_HALF_THICKNESS 0.025f /* 0.05f / 2, src/Cmem.cpp:133 */

// the five regions of CMEM::GetSample (:135-176) for known radii
__device__ __forceinline__ float cmem_regions(const CmemDev& D, float r, float theta, float sin_t,
                                              float radius_mp, float radius_bs) {
    const birp_cmem_params& P = D.P;
    const float half_thickness = BIRP_CMEM_HALF_THICKNESS;
    const float mp_lo = __fsub_rn(radius_mp, half_thickness), mp_hi = __fadd_rn(radius_mp, half_thickness);
    const float bs_lo = __fsub_rn(radius_bs, half_thickness), bs_hi = __fadd_rn(radius_bs, half_thickness);

    if (r < mp_lo) return 0.f;

    const bool in_mp_layer = (r >= mp_lo && r <= mp_hi);
    const bool in_sheath = (r > mp_hi && r < bs_lo);
    const bool in_bs_layer = (r >= bs_lo && r <= bs_hi);

    float b1 = 0.f;
    if (in_mp_layer || in_sheath || in_bs_layer) b1 = cmem_b1(D, r, theta, sin_t);
    if (in_mp_layer) {
        const float shue_norm = __fdiv_rn(__fsub_rn(r, mp_lo), mp_hi);
        return lerp_ref(0.f, b1, shue_norm);
    }
    if (in_sheath) return b1;
    const float b2 = cmem_b2(P, r);
    if (in_bs_layer) {
        const float shue_norm = __fdiv_rn(__fsub_rn(r, bs_lo), bs_hi);
        return lerp_ref(b1, b2, shue_norm);
    }
    return b2;
}

__device__ __forceinline__ float cmem_sample_v4(const CmemDev& D, float x, float y, float z, bool above_mp = false) {
    float r, q_t, q_p, theta, sin_t, radius_mp, radius_bs;
    cmem_cosines(x, y, z, r, q_t, q_p);
    cmem_radii_v4(D, q_t, q_p, theta, sin_t, radius_mp, radius_bs, above_mp);
    return cmem_regions(D, r, theta, sin_t, radius_mp, radius_bs);
}

// ---------------------------------------------------------------------------------
// Region filter.  Three samples in four lie well inside the magnetopause (emissivity 0) or well beyond
// the bow shock (emissivity A2 / (r/10)^3, a function of r alone), and the rest mostly in the open
// magnetosheath, whose emissivity needs theta but neither radius.  WHICH region a sample is in follows
// from comparing r with the two radii -- and away from the boundaries a cheap fp32 estimate of the radii
// (hardware exp2 / log2 / rsqrt approximations, ~1e-6 relative) decides that with certainty: if r clears
// a boundary by more than the estimate can possibly be off (BIRP_CMEM_MARGIN_*, more than 20x the largest deviation found
// by the device self-test over 10^8 directions, tests/test_gpu_parity.py::test_cmem_filter_margin), the
// exact evaluation would take the same branch, so its result is reproduced bit for bit at a fraction of
// the cost.  Everything else -- within the margin of a boundary, directions where the estimate is
// ill-conditioned (near the -x axis, near a cusp), NaNs -- takes the full evaluation.
// ---------------------------------------------------------------------------------
#define BIRP_CMEM_MARGIN_REL 2e-3f
#define BIRP_CMEM_MARGIN_ABS 2e-3f

// hardware approximations for the estimate (MUFU.SQRT / MUFU.RCP, ~2^-22 relative)
__device__ __forceinline__ float est_sqrt(float x) {
#ifdef __CUDACC__
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}
// 1 / sqrt(x) from the hardware unit alone (MUFU.RSQ; rsqrtf() wraps it in a rescaling of subnormal arguments, four
// instructions per call that the filter would pay at every sample)
__device__ __forceinline__ float est_rsqrt(float x) {
#ifdef __CUDACC__
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}
// acos(x), |error| < 1e-7: Abramowitz & Stegun 4.4.46, acos(x) = sqrt(1 - x) (a0 + a1 x + ... + a7 x^7) on [0, 1]
// (|eps| <= 2e-8), mirrored for x < 0
__device__ __forceinline__ float est_acos(float x) {
    const float ax = fabsf(x);
    float p = -0.0012624911f;
    p = fmaf(p, ax, 0.0066700901f);
    p = fmaf(p, ax, -0.0170881256f);
    p = fmaf(p, ax, 0.0308918810f);
    p = fmaf(p, ax, -0.0501743046f);
    p = fmaf(p, ax, 0.0889789874f);
    p = fmaf(p, ax, -0.2145988016f);
    p = fmaf(p, ax, 1.5707963050f);
    const float a = est_sqrt(1.f - ax) * p;
    return x < 0.f ? 3.14159265f - a : a;
}

// hardware-approximation versions of cmem_cosines' outputs (one MUFU.RSQ each instead of IEEE sqrt + division)
__device__ __forceinline__ void cmem_cosines_estimate(float x, float y, float z, float& r, float& q_t, float& q_p) {
    const float yz2 = y * y + z * z;
    const float s2 = x * x + yz2;
    const float rinv = est_rsqrt(s2);
    r = s2 * rinv;
    q_t = (x != 0.f) ? x * rinv : 1.f;
    q_p = (y != 0.f) ? y * est_rsqrt(yz2) : 1.f;
    // The hardware unit flushes a subnormal argument to zero: a sample within 1e-15 RE of the origin (or of the x axis,
    // where cos(phi) comes from y^2 + z^2) is left to the exact evaluation -- a NaN radius fails every comparison of the
    // filter, the region table's included.
    if (!((y != 0.f ? yz2 : s2) >= 1e-30f)) r = __int_as_float(0x7fc00000);
}

// fp32 estimates for the region filter, in two stages so that a sample beyond the bow shock never pays for the
// cusp terms.  Stage 1: the bow-shock radius and the first term of the magnetopause radius, p0 r0 f.  Stage 2:
// the cusp indentation q (two acos, two pow, two exp).  With p2 * charlie <= 0 (D.q_nonpositive, every parameter
// set the reference can produce) q <= 0, i.e. p0 r0 f alone is an UPPER bound of the magnetopause radius.
struct CmemEstimate {
    float cos_t, sin_t, cos_p;
    float sin2_p;    // sin(phi)^2 = 1 - q_p^2 (stage 1 needs no sin(phi) itself)
    float mp_upper;  // p0 r0 f
    float bs;
    bool trusted;    // false in the ill-conditioned directions: 1 + cos(theta) -> 0 (both radii blow up)
};

__device__ __forceinline__ CmemEstimate cmem_estimate_stage1(const CmemDev& D, float q_t, float q_p) {
    const birp_cmem_params& P = D.P;
    CmemEstimate e;
    e.cos_t = q_t, e.cos_p = q_p;  // cos(acos q) = q
    e.sin_t = est_sqrt((1.f - q_t) * (1.f + q_t));
    e.sin2_p = (1.f - q_p) * (1.f + q_p);
    const float theta = est_acos(q_t);
    const float fbase = fmaf(P.a5 * (2.f * e.sin_t * e.cos_t), 1.f - __expf(-theta), est_sqrt(0.5f * (1.f + q_t)));
    const float fexp = P.p1 * (P.beta_c[0] + P.beta_c[1] * e.cos_p + P.beta_c[3] * e.sin2_p);
    e.mp_upper = D.p0r0 * __powf(fbase, fexp);
    // (2 / (1 + cos theta))^a = 2^(a (1 - log2(1 + cos theta)))
    const float lb = 1.f - __log2f(1.f + e.cos_t);
    const float ry = P.bs * exp2f(P.ay_bs * lb);
    if (D.same_flare) {
        e.bs = ry;  // ry = rz: ry rz / sqrt(ry^2 (cos^2 + sin^2)) = ry
    } else {
        const float rz = P.bs * exp2f(P.az_bs * lb);
        const float rzc = rz * e.cos_p;
        e.bs = ry * rz * est_rsqrt(fmaf(rzc, rzc, ry * ry * e.sin2_p));
    }
    e.trusted = (1.f + q_t > 0.02f) && (fbase > 0.05f);
    return e;
}

// the full magnetopause estimate; `trusted` is cleared next to a cusp (acos near +-1 is ill-conditioned)
__device__ __forceinline__ float cmem_estimate_stage2(const CmemDev& D, CmemEstimate& e) {
    const birp_cmem_params& P = D.P;
    const float sn = e.sin_t * est_sqrt(e.sin2_p);
    const float arg_n = fmaf(e.cos_t, D.cos_tn, sn * D.sin_tn);
    const float arg_s = fmaf(e.cos_t, D.cos_ts, -sn * D.sin_ts);
    e.trusted = e.trusted && (fabsf(arg_n) < 0.9995f) && (fabsf(arg_s) < 0.9995f);
    const float phi_n = est_acos(arg_n), phi_s = est_acos(arg_s);
    const float q = D.p2c * (__expf(D.p3dn * __powf(phi_n, P.a21)) + __expf(D.p3ds * __powf(phi_s, P.a21)));
    return e.mp_upper + q;
}

#define BIRP_CMEM_MARGIN(v) (BIRP_CMEM_MARGIN_REL * fabsf(v) + BIRP_CMEM_MARGIN_ABS)

// The emissivity of a sample whose REGION is known.  FASTVAL = false: the fp64-then-round evaluation of the default
// path (bit for bit what the unfiltered evaluation returns).  FASTVAL = true (BIRP_CMEM_FAST, tolerance mode): fp32
// with the hardware approximations, ~1e-6 relative -- the estimated r and cos(theta) the filter already holds are
// all it needs: sin^2(theta) = 1 - cos^2(theta), x^y = 2^(y log2 x), and exp(-B (theta/2)^4) is 1 to within 1e-7
// unless B is far from what CMEM::Init leaves there (src/Cmem.cpp:310: |b| * 1e-9).
template <bool FASTVAL>
__device__ __forceinline__ float cmem_outside_value(const CmemDev& D, float x, float y, float z, float r_est) {
    if (!FASTVAL) return cmem_b2(D.P, vec_len3(x, y, z));
    const float sd = r_est * 0.1f;
    return __fdividef(D.P.A2, sd * sd * sd);
}

template <bool FASTVAL>
__device__ __forceinline__ float cmem_sheath_value(const CmemDev& D, float x, float y, float z, float r_est, float q_t) {
    if (!FASTVAL) {
        const float r_exact = vec_len3(x, y, z);
        const cm::Angle T = cm::angle_from_cosine((x != 0.f) ? __fdiv_rn(x, r_exact) : 1.f);
        return cmem_b1(D, r_exact, T.a, (float)T.s);
    }
    const birp_cmem_params& P = D.P;
    const float sin2 = (1.f - q_t) * (1.f + q_t);
    const float pw = exp2f((-P.alpha - P.beta * sin2) * __log2f(r_est * 0.1f));
    float e = 1.f;
    if (fabsf(P.B) > 1e-8f) {  // (theta <= pi: B (theta/2)^4 <= 6.1 B)
        const float h = 0.5f * est_acos(q_t);
        e = __expf(-P.B * (h * h) * (h * h));
    }
    return P.A1 * e * pw;
}

// The filter: kCmemDecided with the sample in `value` when its region is certain; kCmemUndecided when the full
// evaluation has to decide; kCmemAboveMp when the sample lies above the magnetopause layer for certain and only the bow
// shock is left to decide (cmem_sample_v4 with above_mp).  "Above for certain" needs the TRUE radius to be a number: next
// to a cusp axis rounding can push the argument of the reference's acos beyond 1, the radius is NaN there and the
// reference then takes NEITHER the sheath nor a layer (every comparison with it is false) -- so only the region table
// (whose cells next to a cusp axis bound nothing) and stage 2 (whose trust flag excludes those directions) may say
// so, not the upper bound p0 r0 f of stage 1 on its own.
enum { kCmemUndecided = 0, kCmemDecided = 1, kCmemAboveMp = 2 };
template <bool FASTVAL>
__device__ __forceinline__ int cmem_try_filter(const CmemDev& D, float x, float y, float z, float& value) {
    float r, q_t, q_p;
    cmem_cosines_estimate(x, y, z, r, q_t, q_p);
    // Stage 0: the region table.  Most samples lie far from both boundaries of their direction's cell; for those one
    // 16-byte load decides the region, and neither radius is estimated at all.
    bool table_above_mp = false;
    if (D.table) {
        const float4 tb = __ldg(D.table + cmem_table_index(q_t, q_p));  // {mp_lo, mp_hi, bs_lo, bs_hi}; never NaN
        if (r < tb.x) {
            value = 0.f;
            return kCmemDecided;
        }
        if (r > tb.y) {
            if (r > tb.w) {
                value = cmem_outside_value<FASTVAL>(D, x, y, z, r);
                return kCmemDecided;
            }
            if (r < tb.z) {
                value = cmem_sheath_value<FASTVAL>(D, x, y, z, r, q_t);
                return kCmemDecided;
            }
            table_above_mp = true;  // r > mp_hi for certain, between the table's bounds of the bow shock
        }
    }
    CmemEstimate e = cmem_estimate_stage1(D, q_t, q_p);
    const float ht = BIRP_CMEM_HALF_THICKNESS;
    const float m_bs = BIRP_CMEM_MARGIN(e.bs);
    // every comparison is written so that a NaN anywhere makes it false (-> full evaluation)
    if (D.q_nonpositive && e.trusted && r > e.bs + ht + m_bs && r > e.mp_upper + ht + BIRP_CMEM_MARGIN(e.mp_upper)) {
        // certainly r > bs_hi and r > mp_hi: neither sheath nor a layer
        value = cmem_outside_value<FASTVAL>(D, x, y, z, r);
        return kCmemDecided;
    }
    const float mp = cmem_estimate_stage2(D, e);
    const float m_mp = BIRP_CMEM_MARGIN(mp);
    if (e.trusted && r < mp - ht - m_mp) {  // certainly r < mp_lo
        value = 0.f;
        return kCmemDecided;
    }
    const bool above_mp = e.trusted && r > mp + ht + m_mp;  // certainly r > mp_hi
    if (above_mp && r > e.bs + ht + m_bs) {                 // (parameter sets with q > 0)
        value = cmem_outside_value<FASTVAL>(D, x, y, z, r);
        return kCmemDecided;
    }
    if (above_mp && r < e.bs - ht - m_bs) {  // certainly mp_hi < r < bs_lo: open magnetosheath
        value = cmem_sheath_value<FASTVAL>(D, x, y, z, r, q_t);
        return kCmemDecided;
    }
    return (above_mp || table_above_mp) ? kCmemAboveMp : kCmemUndecided;
}

__device__ __forceinline__ float cmem_sample_v5(const CmemDev& D, float x, float y, float z) {
    float value;
    const int decided = cmem_try_filter<false>(D, x, y, z, value);
    if (decided == kCmemDecided) return value;
    return cmem_sample_v4(D, x, y, z, decided == kCmemAboveMp);
}

// BIRP_CMEM_FAST (opt-in, tolerance mode): the same region decisions from the table and the fp32 estimates, values in
// fp32; a sample next to a boundary is evaluated with the fp32 math library (region decided by fp32 radii, which
// agree with the reference's on all but about one sample in 10^4 crossings)
__device__ __forceinline__ float cmem_sample_fast(const CmemDev& D, float x, float y, float z) {
    float value;
    if (cmem_try_filter<true>(D, x, y, z, value) == kCmemDecided) return value;
    return cmem_sample<false>(D, x, y, z);
}

// One cell of the region table (cmem_table_kernel; tests/cpp/device_on_host.cpp runs the same function on the host).
// The EXACT radii (cmem_radii_v4, what the march itself computes) on a 13 x 13 grid of directions across the cell and
// its eight neighbours (a sample's estimated cosines may land it next door), the extremes widened by twice the
// largest difference between neighbouring grid points (what the grid can miss between its points) and a quarter of
// their range, by the layers' half thickness and by the margin the filter allows its fp32 estimates (the table is
// compared with the ESTIMATED r).  Where the fp32 estimates are not trusted the exact radius is not smooth either --
// within 2 degrees of a cusp axis the Lin radius has a needle no grid resolves, towards the -x axis both radii blow
// up -- so a cell with an untrusted grid point bounds nothing; nor does a cell with a non-finite radius.  The first
// and the last row (directions along the x axis, where cos(phi) is ill-defined) hold the bounds of the whole row.
__device__ __forceinline__ void cmem_table_cell(const CmemDev& D, int bt, int bp, float out[4]) {
    constexpr int NS = 12;
    const float wt = 2.f / kCmemTableT, wp = 2.f / kCmemTableP;
    const bool polar = (bt == 0 || bt == kCmemTableT - 1);
    const float t_lo = fmaxf(-1.f, -1.f + (float)(bt - 1) * wt), t_hi = fminf(1.f, -1.f + (float)(bt + 2) * wt);
    const float p_lo = polar ? -1.f : fmaxf(-1.f, -1.f + (float)(bp - 1) * wp);
    const float p_hi = polar ? 1.f : fminf(1.f, -1.f + (float)(bp + 2) * wp);
    const int np = polar ? 16 * NS : NS;
    float mn[2] = {CUDART_INF_F, CUDART_INF_F}, mx[2] = {-CUDART_INF_F, -CUDART_INF_F}, dif[2] = {0.f, 0.f};
    bool usable[2] = {true, true};
    float prev_row[2][16 * NS + 1];
    for (int it = 0; it <= NS; ++it) {
        const float q_t = t_lo + (t_hi - t_lo) * ((float)it / (float)NS);
        float prev[2] = {0.f, 0.f};
        for (int ip = 0; ip <= np; ++ip) {
            const float q_p = p_lo + (p_hi - p_lo) * ((float)ip / (float)np);
            float theta, sin_t, rad[2];
            cmem_radii_v4(D, q_t, q_p, theta, sin_t, rad[0], rad[1]);
            CmemEstimate e = cmem_estimate_stage1(D, q_t, q_p);
            if (!e.trusted) usable[0] = usable[1] = false;
            (void)cmem_estimate_stage2(D, e);
            if (!e.trusted) usable[0] = false;  // (next to a cusp axis: only the magnetopause has the needle)
            for (int a = 0; a < 2; ++a) {
                const float v = rad[a];
                if (!(fabsf(v) < CUDART_INF_F)) usable[a] = false;
                mn[a] = fminf(mn[a], v);
                mx[a] = fmaxf(mx[a], v);
                if (ip > 0) dif[a] = fmaxf(dif[a], fabsf(v - prev[a]));
                if (it > 0) dif[a] = fmaxf(dif[a], fabsf(v - prev_row[a][ip]));
                prev[a] = v;
                prev_row[a][ip] = v;
            }
        }
    }
    for (int a = 0; a < 2; ++a) {
        const float slack = 2.f * dif[a] + 0.25f * (mx[a] - mn[a]) + 1e-4f * fabsf(mx[a]);
        float lo = mn[a] - slack - BIRP_CMEM_HALF_THICKNESS - BIRP_CMEM_MARGIN(mn[a]);
        float hi = mx[a] + slack + BIRP_CMEM_HALF_THICKNESS + BIRP_CMEM_MARGIN(mx[a]);
        if (!usable[a] || !(lo == lo) || !(hi == hi)) lo = -CUDART_INF_F, hi = CUDART_INF_F;
        out[2 * a] = lo;
        out[2 * a + 1] = hi;
    }
}

// CMEM evaluation mode: 0 fp32 libdevice calls for every sample (diagnostic), 1 CUDA fp64 library rounded once, 2 own fp64
// blocks, 3 (default) own fp64 blocks behind the region table and filter, 4 (BIRP_CMEM_FAST) fp32 values behind them
template <int MODE>
__device__ __forceinline__ float cmem_sample_mode(const CmemDev& D, float x, float y, float z) {
    if (MODE == 4) return cmem_sample_fast(D, x, y, z);
    if (MODE == 3) return cmem_sample_v5(D, x, y, z);
    if (MODE == 2) return cmem_sample_v4(D, x, y, z);
    return cmem_sample<MODE == 1>(D, x, y, z);
}

}  // namespace birp
