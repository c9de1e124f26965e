// TEST INFRASTRUCTURE.  Lets g++ compile birp_cpp_b200/csrc/birp_device.cuh (the device code of the march kernels)
// for the host, so that the very source the GPU runs -- hot-path sampler, clamp-free lerps, reciprocal division,
// single-multiply scale, the span march with its bisection, the CMEM fp64 blocks -- can be checked against the
// oracle on the CPU (tests/test_device_on_host.py).  Under g++ the CUDA qualifiers (__device__, __constant__, ...)
// expand to nothing (crt/host_defines.h); what is left to provide are the arithmetic intrinsics.
//
// Everything the DataCube path and the fp64 CMEM blocks compute is IEEE arithmetic with explicit roundings
// (x86-64 SSE2, -ffp-contract=off: a * b is __fmul_rn(a, b)), so the host build is bit-identical to the device
// there.  Only the hardware approximations of the CMEM region filter's ESTIMATES (MUFU ex2 / lg2 / rsq / sqrt)
// have libm stand-ins; the filter is built so that their error cannot change a result (margins).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>  // vector types, cudaTextureObject_t, qualifiers
#include <math_constants.h>

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

static inline float __int_as_float(int i) {  // (math_constants.h spells CUDART_INF_F with it)
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
static inline int __float_as_int(float f) {
    int i;
    std::memcpy(&i, &f, 4);
    return i;
}
static inline unsigned __float_as_uint(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(unsigned u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __double2loint(double a) {
    uint64_t u;
    std::memcpy(&u, &a, 8);
    return (int)(uint32_t)u;
}
static inline int __double2hiint(double a) {
    uint64_t u;
    std::memcpy(&u, &a, 8);
    return (int)(uint32_t)(u >> 32);
}
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double a;
    std::memcpy(&a, &u, 8);
    return a;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int atomicOr(int* p, int v) {
    const int old = *p;
    *p = old | v;
    return old;
}
// stand-ins for the hardware approximations (region-filter estimates only)
// (glibc declares __powf / __expf / __log2f as internal aliases of its own: map the names, do not redefine them)
#define __powf powf
#define __expf expf
#define __log2f log2f
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }

// provided by the harness (device_on_host.cpp): the staged ray table at a byte offset, and the texture gather
float birp_host_table_entry(unsigned sdist);
void birp_host_gather4(cudaTextureObject_t tex, int layer, float u, float v, float& t01, float& t11, float& t10,
                       float& t00);
