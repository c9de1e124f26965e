// sm_100a kernels of the line-of-sight integrator.
//
//   march_cube_kernel<..> K1/K3: ray generation + fixed-step march + DataCube sampling (nearest /
//                         trilinear) + scale + terminate + accumulate, for a batch of frames
//                         (reference src/Camera.cpp:140-232 x src/DataCube.cpp:203-277).
//   march_kernel<Field>   K2: the same march around a field functor (CMEM, src/Cmem.cpp:117-288).
//   accumulate_kernel     K4: scale + terminate + accumulate of host-sampled values
//                         (reference src/Camera.cpp:201-223), for user-defined Space classes.
//   sample_*_kernel, ray_directions_kernel: Space::GetSample on points; the ray generator alone.
//   verify_div_kernel, verify_gather_kernel, scan_positive_kernel: upload-time proofs / analysis;
//   pack_octs_kernel / pack_quads_kernel (+ verify_*): the copies of the grid for the corner fetch.
//
// What a single ray does -- samplers, lerps, scale, the march loops (march_steps, march_steps_span) -- lives in
// birp_device.cuh, which also compiles for the host (tests/cpp/device_on_host.cpp checks it against the oracle
// on the CPU); this file holds the __global__ kernels and their launches.
//
// Work decomposition of the march kernels: one warp owns a 4x8 / 8x4 pixel tile of one frame, one
// lane per ray.  The 32 rays of a tile stay within a few voxels of one another along the whole
// march, so the corner fetches of a step hit the same handful of cache lines / texture tiles, and
// the serial per-pixel accumulation order of the reference is kept.
#include <atomic>
#include <cmath>
#include <cstdlib>

#include "birp_frame_math.h"
#include "birp_internal.h"

namespace birp {

namespace {

std::atomic<unsigned long long> g_launches{0};

// The LARGEST CTA (the launch bound).  A launch uses 8 warps per CTA, 4 for very large images (pick_cta_warps).
constexpr int kWarpsPerBlock = 8;
constexpr int kFastVariantDefault = 2;  // fast_sample<VARIANT> of the tolerance-mode march (see birp_device.cuh)
constexpr int kBlockThreads = kWarpsPerBlock * 32;

struct MarchArgs {
    const birp_frame* __restrict__ frames;
    const float* __restrict__ dist;
    float* __restrict__ images;
    unsigned long long* valid;
    int* status;
    int W, H;
    int row_begin, rows;
    // interleaved row bands (one large image over several GPUs): this launch renders, of the bands of `row_band`
    // rows counted from row_begin, those with index = row_phase mod row_nphases; row_band = 0: every row of
    // [row_begin, row_end).  `rows` counts the rows of this launch's own bands, row_end bounds the last one.
    int row_band, row_phase, row_nphases, row_end;
    int nsamples;
    int cta_warps;    // warps per CTA of this launch (blockDim.x / 32)
    int tile_w_log2;  // a warp's 32 rays form a (1 << tile_w_log2) x (32 >> tile_w_log2) pixel tile
    int swizzle;      // lanes 4n .. 4n+3 own a 2x2 pixel block of the tile instead of four pixels of one row
    int tiles_x, tiles_per_frame;
    long long total_tiles;
    float ray_width;
    int accumulate;
    double scale_k;  // BIRP_SCALE_K: a kernel parameter so the fp64 multiply reads it from the constant bank
    // frame f of this launch is written to image slot out_offset + f * out_stride of `images`
    // (1-GPU: 0 / 1; multi-GPU sweeps scatter straight into the gathering rank's buffer)
    long long out_offset, out_stride;
    // tolerance-mode march without table reads (KSTEP): maxdist / nsamples of the uniform ray table; 0 = read the table
    double fast_step;
};

template <int MODE>  // 0 fp32 libdevice, 1 CUDA fp64 library, 2 own fp64 blocks, 3 + table and filter, 4 fast (cmem_sample_mode)
struct CmemField {
    static constexpr bool kScaleOnce = true;
    CmemDev D;
    __device__ __forceinline__ float operator()(float x, float y, float z, bool&) const {
        return cmem_sample_mode<MODE>(D, x, y, z);
    }
};

// Which ray this lane owns; false when the lane falls outside the image / row band.
__device__ __forceinline__ bool lane_ray(const MarchArgs& a, int& f, int& i, int& j) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * a.cta_warps + (threadIdx.x >> 5);
    if (tile >= a.total_tiles) return false;
    f = (int)(tile / a.tiles_per_frame);
    const int t = (int)(tile - (long long)f * a.tiles_per_frame);
    const int ty = t / a.tiles_x;
    const int tx = t - ty * a.tiles_x;
    int lx = lane & ((1 << a.tile_w_log2) - 1), ly = lane >> a.tile_w_log2;
    if (a.swizzle) {
        // A 32-byte-per-lane load is served four lanes at a time, one L1 wavefront per 128-byte line those four lanes
        // touch: four rays of a 2x2 pixel block stay closer together than four rays of a row.
        lx = (lane & 1) | (((lane >> 2) & ((1 << (a.tile_w_log2 - 1)) - 1)) << 1);
        ly = ((lane >> 1) & 1) | ((lane >> (a.tile_w_log2 + 1)) << 1);
    }
    i = (tx << a.tile_w_log2) + lx;
    const int jr = ty * (32 >> a.tile_w_log2) + ly;
    if (a.row_band > 0) {
        const int lb = jr / a.row_band;
        j = a.row_begin + (lb * a.row_nphases + a.row_phase) * a.row_band + (jr - lb * a.row_band);
    } else {
        j = a.row_begin + jr;
    }
    return i < a.W && jr < a.rows && j < a.row_end;
}

// Generic march: any Field functor (used for CMEM; the DataCube path has its own kernel below).
// Field::kScaleOnce: the raw samples are summed in fp32 and the scale of src/Camera.cpp:201 is applied to the sum (the
// analytic field is a tolerance path -- its transcendentals are not glibc's -- so the order of the roundings is free;
// a ray still ends at its first negative sample, whose sign is that of its scaled value for the positive ray width).
// (no minimum of resident CTAs: ptxas then settles on 64 registers = 4 CTAs of 8 warps; A/B builds set BIRP_CMEM_MIN_BLOCKS
// to 5 or 6, which caps the registers at 48 / 40 -- see profiles/README.md for what that measured)
#ifdef BIRP_CMEM_MIN_BLOCKS
#define BIRP_CMEM_LAUNCH_BOUNDS __launch_bounds__(kBlockThreads, BIRP_CMEM_MIN_BLOCKS)
#else
#define BIRP_CMEM_LAUNCH_BOUNDS __launch_bounds__(kBlockThreads)
#endif
template <class Field>
__global__ void BIRP_CMEM_LAUNCH_BOUNDS march_kernel(const MarchArgs a, const Field field) {
    int f, i, j;
    if (!lane_ray(a, f, i, j)) return;
    const birp_frame fr = a.frames[f];
    float dx, dy, dz;
    pixel_direction(fr, i, j, (float)a.W, (float)a.H, dx, dy, dz);

    float* pix = a.images + ((size_t)(a.out_offset + f * a.out_stride) * a.H + j) * a.W + i;
    float acc = (a.accumulate && !Field::kScaleOnce) ? *pix : 0.f;
    bool nan_index = false;
    unsigned nvalid = 0;
    const float ray_width = a.ray_width;
    for (int k = 0; k < a.nsamples; ++k) {
        const float tk = __ldg(a.dist + k);
        // Helper::MatrixScalarMultiply then "+ position" (src/Camera.cpp:183-189): two roundings
        const float x = __fadd_rn(__fmul_rn(dx, tk), fr.pos[0]);
        const float y = __fadd_rn(__fmul_rn(dy, tk), fr.pos[1]);
        const float z = __fadd_rn(__fmul_rn(dz, tk), fr.pos[2]);
        float s = field(x, y, z, nan_index);
        if (!Field::kScaleOnce) s = scale_sample(s, ray_width);
        if (s < 0.f) break;  // src/Camera.cpp:207-210: the rest of the ray is ignored
        acc = __fadd_rn(acc, s);
        ++nvalid;
    }
    if (Field::kScaleOnce) {
        const float total = __double2float_rn(__dmul_rn((double)__fmul_rn(acc, ray_width), a.scale_k));
        acc = a.accumulate ? __fadd_rn(*pix, total) : total;
    }
    *pix = acc;
    if (nan_index && a.status) atomicOr(a.status, kStatusNanIndex);
    if (a.valid) atomicAdd(a.valid, (unsigned long long)nvalid);
}

// DataCube march (K1/K3).  One lane per ray; per step: position (6 flops), bounds window
// (6 compares), 3 exact divisions by the spacing, 8 corner loads at fixed offsets from one
// base address, 7 lerps, the fp64 scale, the fp32 accumulate.  Everything a sample outside
// the window needs (dead ray, NaN coordinate) is on the cold side of one branch.
#ifndef BIRP_CUBE_MIN_BLOCKS
// resident CTAs per SM the register allocation must allow: 8 = all 64 warps at 32 registers per thread.  The hot
// loops of every variant fit into 32 registers without local-memory traffic (one instruction more than at 40
// registers / 6 CTAs), and since v9 the kernel waits for L1 data more than for issue slots, so the extra warps pay:
// orbit sweep 31.5 instead of 32.3 ms with quads, 29.2 instead of 30.3 ms with oct records.
#define BIRP_CUBE_MIN_BLOCKS 8
#endif

// Make a warp-uniform kernel parameter opaque to ptxas AND keep it in a uniform register: the OR-reduction
// of 32 identical words is the word itself, the result of redux.sync lives in the uniform register
// file, and ptxas cannot rematerialise it from the constant bank inside the march loop (left to
// itself it re-loads a dozen constants in every step; issue slots are the scarce resource here).
// Must be called by all 32 lanes.
__device__ __forceinline__ void pin_uniform(float& v) {
    v = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(v)));
}

__device__ __forceinline__ void pin_uniform(int& v) { v = (int)__reduce_or_sync(0xffffffffu, (unsigned)v); }
__device__ __forceinline__ void pin_uniform(long long& v) {
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)((unsigned long long)v >> 32));
    v = (long long)(((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ void pin_uniform(double& v) {
    long long b = __double_as_longlong(v);
    pin_uniform(b);
    v = __longlong_as_double(b);
}

// SPAN: 0 the window test in every step (march_steps), 1 run end by bisection (march_steps_span), 2 the same, lean
template <bool TRILINEAR, int DIVMODE, int LERP, int FETCH, int SPAN>
__global__ void __launch_bounds__(kBlockThreads, BIRP_CUBE_MIN_BLOCKS) march_cube_kernel(const MarchArgs a, const CubeView c) {
    // The ray table is staged through shared memory in chunks of kSmemDist distances (one chunk
    // for any march up to 2048 samples), so a step costs one LDS instead of a global load.
    // (Sizing the buffer to the march -- 0.8 KB of dynamic shared memory for 200 samples, i.e. a 16 KB instead of
    // a 100 KB shared-memory carve-out at 8 CTAs per SM and that much more L1 -- was measured: the trilinear
    // kernel got 1.7 % SLOWER (29.68 vs 29.19 ms), nearest neighbour 1.4 % faster; the static table stays.)
    __shared__ float s_dist[kSmemDist + 1];  // + the look-ahead padding entry
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_dist);
    const int first_chunk = min(kSmemDist, a.nsamples);
    for (int k = threadIdx.x; k < first_chunk; k += blockDim.x) s_dist[k] = __ldg(a.dist + k);
    if (threadIdx.x == 0) s_dist[kSmemDist] = 0.f;
    __syncthreads();
    CubeView cc = c;
    double K = a.scale_k;
    float ray_width = a.ray_width;
    pin_uniform(cc.ox), pin_uniform(cc.oy), pin_uniform(cc.oz);
    pin_uniform(cc.lox), pin_uniform(cc.loy), pin_uniform(cc.loz);
    pin_uniform(cc.lx), pin_uniform(cc.ly), pin_uniform(cc.lz);
    pin_uniform(cc.rx), pin_uniform(cc.ry), pin_uniform(cc.rz);
    pin_uniform(cc.sx), pin_uniform(cc.sy), pin_uniform(cc.sz);
    pin_uniform(ray_width);
    pin_uniform(K);
    pin_uniform(cc.nx), pin_uniform(cc.ny);
    if (FETCH == 4) {
        long long q0 = (long long)cc.quad, q1 = q0 + cc.slab16;
        pin_uniform(q0), pin_uniform(q1);
        cc.quad = reinterpret_cast<const float4*>(q0);
        cc.quad_up = reinterpret_cast<const float4*>(q1);
        cc.quad_bytes = 16;
        asm volatile("" : "+r"(cc.quad_bytes));
    }
    if (SPAN == 2) pin_uniform(cc.fnx), pin_uniform(cc.fny);
    if (FETCH == 5) {
        long long q0 = (long long)cc.oct;
        pin_uniform(q0);
        cc.oct = reinterpret_cast<const float*>(q0);
    }

    int f = 0, i = 0, j = 0;
    const bool owns_pixel = lane_ray(a, f, i, j);
    if (!owns_pixel && a.nsamples <= kSmemDist) return;  // (longer marches: keep helping to stage the table)
    RayState r = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* pix = nullptr;
    if (owns_pixel) {
        const birp_frame* __restrict__ frp = a.frames + f;
        pixel_direction(*frp, i, j, (float)a.W, (float)a.H, r.dx, r.dy, r.dz);
        r.px = frp->pos[0];
        r.py = frp->pos[1];
        r.pz = frp->pos[2];
        pix = a.images + ((size_t)(a.out_offset + f * a.out_stride) * a.H + j) * a.W + i;
        r.acc = a.accumulate ? *pix : 0.f;
    }
    bool alive = owns_pixel;
    unsigned nlive4 = 0;  // 4 x the number of samples accumulated
    if (alive)
        nlive4 = (SPAN ? march_steps_span<TRILINEAR, DIVMODE, LERP, FETCH, SPAN == 2>(cc, r, s_base, s_base + 4u * first_chunk,
                                                                                         ray_width, K, a.status, alive)
                       : march_steps<TRILINEAR, DIVMODE, LERP, FETCH>(cc, r, s_base, s_base + 4u * first_chunk,
                                                                         ray_width, K, a.status, alive)) - s_base;
    // marches longer than the staged table: refill it chunk by chunk (every thread of the CTA takes part)
    for (int base = kSmemDist; base < a.nsamples; base += kSmemDist) {
        const int chunk = min(kSmemDist, a.nsamples - base);
        __syncthreads();
        for (int k = threadIdx.x; k < chunk; k += blockDim.x) s_dist[k] = __ldg(a.dist + base + k);
        __syncthreads();
        if (alive)
            nlive4 += (SPAN ? march_steps_span<TRILINEAR, DIVMODE, LERP, FETCH, SPAN == 2>(cc, r, s_base, s_base + 4u * chunk,
                                                                                              ray_width, K, a.status, alive)
                            : march_steps<TRILINEAR, DIVMODE, LERP, FETCH>(cc, r, s_base, s_base + 4u * chunk,
                                                                              ray_width, K, a.status, alive)) - s_base;
    }
    if (owns_pixel) {
        *pix = r.acc;
        if (a.valid) atomicAdd(a.valid, (unsigned long long)(nlive4 >> 2));
    }
}


// Tolerance-mode DataCube march (BIRP_FAST_MARCH): same work decomposition and prologue as march_cube_kernel; the
// per-ray loop is march_steps_fast (birp_device.cuh), the pixel is the fp32 sum of the raw samples scaled once
// (src/Camera.cpp:201 applied to the sum).
#ifndef BIRP_FAST_MIN_BLOCKS
#define BIRP_FAST_MIN_BLOCKS 8
#endif
template <bool SIGNS, int VARIANT, bool KSTEP>
__global__ void __launch_bounds__(kBlockThreads, BIRP_FAST_MIN_BLOCKS) march_cube_fast_kernel(const MarchArgs a, const CubeView c) {
    __shared__ float s_dist[kSmemDist + 1];
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_dist);
    const int first_chunk = min(kSmemDist, a.nsamples);
    for (int k = threadIdx.x; k < first_chunk; k += blockDim.x) s_dist[k] = __ldg(a.dist + k);
    if (threadIdx.x == 0) s_dist[kSmemDist] = 0.f;
    __syncthreads();
    CubeView cc = c;
    float ray_width = a.ray_width;
    if (VARIANT == 0) {
        pin_uniform(cc.nx), pin_uniform(cc.ny);
        const unsigned B = __float_as_uint(BIRP_RINT_MAGIC);
        cc.poly_bias = (int)((B * (unsigned)cc.ny + B) * (unsigned)cc.nx + B);
        pin_uniform(cc.poly_bias);
    } else {
        pin_uniform(cc.fnx), pin_uniform(cc.fny);
    }
    cc.poly_bytes = 32;

    int f = 0, i = 0, j = 0;
    const bool owns_pixel = lane_ray(a, f, i, j);
    if (!owns_pixel && a.nsamples <= kSmemDist) return;
    RayState r = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    FastRay q = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* pix = nullptr;
    if (owns_pixel) {
        const birp_frame* __restrict__ frp = a.frames + f;
        pixel_direction(*frp, i, j, (float)a.W, (float)a.H, r.dx, r.dy, r.dz);
        r.px = frp->pos[0];
        r.py = frp->pos[1];
        r.pz = frp->pos[2];
        q = fast_ray(cc, r, KSTEP ? a.fast_step : 1.0);
        pix = a.images + ((size_t)(a.out_offset + f * a.out_stride) * a.H + j) * a.W + i;
    }
    bool alive = owns_pixel;
    unsigned nlive4 = 0;
    if (alive)
        nlive4 = march_steps_fast<SIGNS, VARIANT, KSTEP>(cc, r, q, s_base, s_base + 4u * first_chunk, ray_width, a.status, alive,
                                                         s_base, 0) - s_base;
    for (int base = kSmemDist; base < a.nsamples; base += kSmemDist) {
        const int chunk = min(kSmemDist, a.nsamples - base);
        __syncthreads();
        for (int k = threadIdx.x; k < chunk; k += blockDim.x) s_dist[k] = __ldg(a.dist + base + k);
        __syncthreads();
        if (alive)
            nlive4 += march_steps_fast<SIGNS, VARIANT, KSTEP>(cc, r, q, s_base, s_base + 4u * chunk, ray_width, a.status, alive,
                                                              s_base, base) - s_base;
    }
    if (owns_pixel) {
        const float total = __double2float_rn(__dmul_rn((double)__fmul_rn(r.acc, ray_width), a.scale_k));
        *pix = a.accumulate ? __fadd_rn(*pix, total) : total;
        if (a.valid) atomicAdd(a.valid, (unsigned long long)(nlive4 >> 2));
    }
}


// EXPERIMENT (north_star: "warp-shuffle reductions of partial segment sums"): the tolerance-mode march with the 32
// lanes of a warp cooperating on ONE ray at a time -- lane l takes steps l, l + 32, ... of the ray's first hot run,
// the partial sums are reduced with shuffles -- instead of one lane per ray.  Only the accumulation order changes
// (allowed in tolerance mode).  Measured on the orbit sweep (profiles/r2c_*): the 32 samples of a request are spread
// over ~22 voxels along the ray instead of 2-3 across a pixel tile, i.e. many more 128-byte lines per load on an
// L1-wavefront-bound kernel.  Selected by birp_set_tuning("fast_coop", 1); positive / non-negative cubes of at most
// 2^24 voxels, uniform ray table.
template <int VARIANT>
__global__ void __launch_bounds__(kBlockThreads, BIRP_FAST_MIN_BLOCKS) march_cube_fast_coop_kernel(const MarchArgs a, const CubeView c) {
    __shared__ float s_dist[kSmemDist + 1];
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_dist);
    const int first_chunk = min(kSmemDist, a.nsamples);
    for (int k = threadIdx.x; k < first_chunk; k += blockDim.x) s_dist[k] = __ldg(a.dist + k);
    if (threadIdx.x == 0) s_dist[kSmemDist] = 0.f;
    __syncthreads();
    CubeView cc = c;
    float ray_width = a.ray_width;
    pin_uniform(cc.fnx), pin_uniform(cc.fny);
    cc.poly_bytes = 32;
    int f = 0, i = 0, j = 0;
    const bool owns_pixel = lane_ray(a, f, i, j);  // (a whole warp shares `tile < total_tiles`: no early exit needed)
    RayState r = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    FastRay q = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* pix = nullptr;
    unsigned sdist = s_base;
    const unsigned send = s_base + 4u * first_chunk;
    int nrun = 0;
    if (owns_pixel) {
        const birp_frame* __restrict__ frp = a.frames + f;
        pixel_direction(*frp, i, j, (float)a.W, (float)a.H, r.dx, r.dy, r.dz);
        r.px = frp->pos[0], r.py = frp->pos[1], r.pz = frp->pos[2];
        q = fast_ray(cc, r, a.fast_step);
        pix = a.images + ((size_t)(a.out_offset + f * a.out_stride) * a.H + j) * a.W + i;
        float xi, yi, zi;
        if (step_ingress(cc, r, sdist, xi, yi, zi)) {  // the run of hot steps that starts at step 0
            unsigned lo = sdist + 4, hi = send;
            while (lo < hi) {
                const unsigned mid = lo + (((hi - lo) >> 3) << 2);
                float ux, uy, uz;
                if (step_ingress(cc, r, mid, ux, uy, uz)) lo = mid + 4;
                else hi = mid;
            }
            nrun = (int)((lo - sdist) >> 2);
        }
    }
    const int lane = threadIdx.x & 31;
    for (int src = 0; src < 32; ++src) {
        const int n = __shfl_sync(0xffffffffu, nrun, src);
        if (n == 0) continue;
        const float gx = __shfl_sync(0xffffffffu, q.gx, src), gy = __shfl_sync(0xffffffffu, q.gy, src),
                    gz = __shfl_sync(0xffffffffu, q.gz, src);
        const float ex = __shfl_sync(0xffffffffu, q.ex, src), ey = __shfl_sync(0xffffffffu, q.ey, src),
                    ez = __shfl_sync(0xffffffffu, q.ez, src);
        float part = 0.f;
        for (int k = lane; k < n; k += 32) {
            const float kf = (float)k + 0.5f;
            part += fast_sample<VARIANT>(cc, __fmaf_rn(kf, ex, gx), __fmaf_rn(kf, ey, gy), __fmaf_rn(kf, ez, gz));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == src) r.acc = part;
    }
    bool alive = owns_pixel;
    unsigned nlive4 = 0;
    if (alive) {
        sdist += 4u * (unsigned)nrun;
        nlive4 = march_steps_fast<false, VARIANT, true>(cc, r, q, sdist, send, ray_width, a.status, alive, s_base, 0) - s_base;
    }
    if (owns_pixel) {
        const float total = __double2float_rn(__dmul_rn((double)__fmul_rn(r.acc, ray_width), a.scale_k));
        *pix = a.accumulate ? __fadd_rn(*pix, total) : total;
        if (a.valid) atomicAdd(a.valid, (unsigned long long)(nlive4 >> 2));
    }
}

// poly copy: the trilinear interpolant of every cell as eight polynomial coefficients (poly_coefficients), neighbours
// clamped at the last row / column / layer like the oct copy
__global__ void __launch_bounds__(256) pack_poly_kernel(const float* __restrict__ v, int nx, int ny, int nz,
                                                        float4* __restrict__ poly) {
    const size_t n = (size_t)nx * ny * nz;
    const size_t slab = (size_t)nx * ny;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int y = (int)((t / nx) % ny);
        const int z = (int)(t / slab);
        const int dx = (x + 1 < nx) ? 1 : 0;
        const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
        const size_t u = t + ((z + 1 < nz) ? slab : 0);
        float c[8];
        poly_coefficients(__ldg(v + t), __ldg(v + t + dx), __ldg(v + t + dy), __ldg(v + t + dy + dx), __ldg(v + u),
                          __ldg(v + u + dx), __ldg(v + u + dy), __ldg(v + u + dy + dx), c);
        poly[2 * t] = make_float4(c[0], c[1], c[2], c[3]);
        poly[2 * t + 1] = make_float4(c[4], c[5], c[6], c[7]);
    }
}

// Tile shape: split the warp's 32 rays so that the tile is as square as possible in ANGLE
// (the SXI pixel is 1.7x wider than tall once the reference's extra aspect factor is applied):
// compact ray bundles touch fewer cache lines per corner load.
// Warps per CTA.  Measured with tools/ab_rows.py (profiles/r2m_ab_rows_*.json: separate builds; profiles/r2n_ab_rows_*.json:
// this launch-time choice), 8 / 4 / 2 warps: one 4096^2 image of the 512^3 cube 7.09 / 7.09 / 7.21 ms (separate builds:
// 7.13 / 6.99 / 7.09), one rank's eighth of it 0.924 / 0.918 / 0.940 ms (tolerance mode 0.510 / 0.502 / 0.574); a 512^2
// image 0.158 / 0.164 / 0.173 ms; the orbit sweep of 64 x 108 frames 27.57 / 27.67 / 27.79 ms.  So: 4 warps for frames of
// half a million rays and more (a CTA's slot is refilled only when its slowest warp is done), 8 otherwise -- a 1-2 % matter.
// The CMEM march (`analytic`): its rays differ far more in cost (which regions they cross), and a CTA's slot waits for
// its slowest warp: 107.1 / 104.5 / 104.0 ms for the 10 000-frame sweep with 8 / 4 / 2 warps (profiles/r2p_ab_cmem_*.json;
// capping the registers at 48 or 40 for more resident warps instead: 107.1 / 104.0, but the tolerance mode loses 5 %).
int pick_cta_warps(const RenderGeom& g, int tiles_per_frame, bool analytic) {
    if (g.cta_warps == 2 || g.cta_warps == 4 || g.cta_warps == 8) return g.cta_warps;
    if (analytic) return 4;
    return tiles_per_frame >= 16384 ? 4 : kWarpsPerBlock;  // >= 2^19 rays of one frame in this launch
}

int pick_tile_w_log2(const RenderGeom& g) {
    if (g.tile_w_log2 >= 0 && g.tile_w_log2 <= 5) return g.tile_w_log2;
    // default: the reference ties the image size to the field of view (W ~ fov_w / px,
    // H ~ fov_h / px) and then stretches x by W/H once more, so a pixel step in x spans about
    // W/H times the angle of a step in y
    const double ax = g.pixel_angle_x > 0 ? g.pixel_angle_x : (double)g.W;
    const double ay = g.pixel_angle_y > 0 ? g.pixel_angle_y : (double)g.H;
    int best = 3;
    double best_cost = 1e300;
    for (int l = 0; l <= 5; ++l) {
        const int tw = 1 << l, th = 32 >> l;
        if (tw > 2 * g.W && l > 0) continue;
        const double cost = tw * ax + th * ay;  // half-perimeter of the tile's angular footprint
        if (cost < best_cost) {
            best_cost = cost;
            best = l;
        }
    }
    return best;
}

bool fill_args(MarchArgs& a, const birp_frame* d_frames, const RenderGeom& g, const float* d_dist, int flags,
               float* d_images, unsigned long long* d_valid, int* d_status, unsigned* blocks, bool analytic = false) {
    a.frames = d_frames;
    a.dist = d_dist;
    a.images = d_images;
    a.valid = d_valid;
    a.status = d_status;
    a.W = g.W;
    a.H = g.H;
    a.row_begin = g.row_begin;
    a.rows = g.row_end - g.row_begin;
    a.row_end = g.row_end;
    a.row_band = 0, a.row_phase = 0, a.row_nphases = 1;
    if (g.row_band > 0 && g.row_nphases > 1) {
        const int total_bands = (a.rows + g.row_band - 1) / g.row_band;
        const int own = (total_bands > g.row_phase) ? (total_bands - g.row_phase + g.row_nphases - 1) / g.row_nphases : 0;
        a.row_band = g.row_band, a.row_phase = g.row_phase, a.row_nphases = g.row_nphases;
        a.rows = own * g.row_band;
    }
    a.nsamples = g.nsamples;
    a.tile_w_log2 = pick_tile_w_log2(g);
    a.swizzle = (g.tile_swizzle && a.tile_w_log2 >= 1 && a.tile_w_log2 <= 4) ? 1 : 0;
    const int tw = 1 << a.tile_w_log2, th = 32 >> a.tile_w_log2;
    a.tiles_x = (g.W + tw - 1) / tw;
    a.tiles_per_frame = a.tiles_x * ((a.rows + th - 1) / th);
    a.total_tiles = (long long)a.tiles_per_frame * g.nframes;
    a.ray_width = g.ray_width;
    a.accumulate = (flags & BIRP_ACCUMULATE) ? 1 : 0;
    a.scale_k = BIRP_SCALE_K;
    a.fast_step = 0.0;
    a.out_offset = g.out_frame_offset;
    a.out_stride = g.out_frame_stride;
    a.cta_warps = pick_cta_warps(g, a.tiles_per_frame, analytic);
    const long long nb = (a.total_tiles + a.cta_warps - 1) / a.cta_warps;
    *blocks = (unsigned)nb;
    return nb <= 0x7fffffffLL;
}

template <class Field>
cudaError_t launch_march(const Field& field, const birp_frame* d_frames, const RenderGeom& g,
                         const float* d_dist, int flags, float* d_images, unsigned long long* d_valid,
                         int* d_status, cudaStream_t stream) {
    MarchArgs a;
    unsigned blocks;
    if (!fill_args(a, d_frames, g, d_dist, flags, d_images, d_valid, d_status, &blocks, /*analytic=*/true))
        return cudaErrorInvalidConfiguration;
    if (a.total_tiles <= 0) return cudaSuccess;
    march_kernel<Field><<<blocks, a.cta_warps * 32, 0, stream>>>(a, field);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <bool TRI, int DIV, int LERP, int FETCH, int SPAN>
cudaError_t launch_cube_variant(const MarchArgs& a, const CubeView& c, unsigned blocks, cudaStream_t stream) {
    march_cube_kernel<TRI, DIV, LERP, FETCH, SPAN><<<blocks, a.cta_warps * 32, 0, stream>>>(a, c);
    return cudaGetLastError();
}

// `span`: the window test is taken out of the step (march_steps_span).  Built for the default corner fetch
// (quads) and for nearest neighbour; the other fetch forms -- diagnostics, or a cube whose quad copy did not
// fit -- keep the per-step test.
template <int DIV, int LERP>
cudaError_t launch_cube_fetch(const MarchArgs& a, const CubeView& c, unsigned blocks, int fetch, bool span, bool lean,
                              cudaStream_t stream) {
    switch (fetch) {
        case 1: return launch_cube_variant<true, DIV, LERP, 1, 0>(a, c, blocks, stream);
        case 2: return launch_cube_variant<true, DIV, LERP, 2, 0>(a, c, blocks, stream);
        case 5:
            if (lean && LERP == 1) return launch_cube_variant<true, DIV, 1, 5, 2>(a, c, blocks, stream);
            // a marked oct copy (cubes outside the positive class) is only ever read by the mark-aware form
            if (LERP != 1) return launch_cube_variant<true, DIV, 3, 5, 1>(a, c, blocks, stream);
            return launch_cube_variant<true, DIV, 1, 5, 1>(a, c, blocks, stream);
        case 4:
            if (span) return launch_cube_variant<true, DIV, LERP, 4, 1>(a, c, blocks, stream);
            return launch_cube_variant<true, DIV, LERP, 4, 0>(a, c, blocks, stream);
        default: return launch_cube_variant<true, DIV, LERP, 0, 0>(a, c, blocks, stream);
    }
}

// lerp: 0 full std::lerp, 1 positive cube, 2 non-negative cube
template <int DIV>
cudaError_t launch_cube_div(const MarchArgs& a, const CubeView& c, unsigned blocks, bool tri, int lerp, int fetch,
                            bool span, bool lean, cudaStream_t stream) {
    if (!tri) {
        // nearest neighbour interpolates nothing; the class only decides whether a zero product can be a sample
        if (lerp == 1) {
            if (span) return launch_cube_variant<false, DIV, 1, 0, 1>(a, c, blocks, stream);
            return launch_cube_variant<false, DIV, 1, 0, 0>(a, c, blocks, stream);
        }
        if (span) return launch_cube_variant<false, DIV, 0, 0, 1>(a, c, blocks, stream);
        return launch_cube_variant<false, DIV, 0, 0, 0>(a, c, blocks, stream);
    }
    if (lerp == 1) return launch_cube_fetch<DIV, 1>(a, c, blocks, fetch, span, lean, stream);
    if (lerp == 2) return launch_cube_fetch<DIV, 2>(a, c, blocks, fetch, span, false, stream);
    return launch_cube_fetch<DIV, 0>(a, c, blocks, fetch, span, false, stream);
}

// upload-time self-check of the texture path: gathered texels == linear voxels at `n` cells
__global__ void __launch_bounds__(256) verify_gather_kernel(const CubeView c, int n, int* bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned h = (unsigned)t * 2654435761u + 17u;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    // corners first (t < 8), then pseudo-random cells
    int x0, y0, z0;
    if (t < 8) {
        x0 = (t & 1) ? c.nx - 2 : 0;
        y0 = (t & 2) ? c.ny - 2 : 0;
        z0 = (t & 4) ? c.nz - 2 : 0;
    } else {
        x0 = (int)(h % (unsigned)(c.nx - 1));
        y0 = (int)((h >> 8) % (unsigned)(c.ny - 1));
        z0 = (int)((h >> 16) % (unsigned)(c.nz - 1));
    }
    bool ok = true;
    for (int dz = 0; dz < 2; ++dz) {
        float t01, t11, t10, t00;
        gather4(c.tex, z0 + dz, (float)x0 + 1.0f, (float)y0 + 1.0f, t01, t11, t10, t00);
        const float* p = c.v + (((size_t)(z0 + dz) * c.ny + y0) * c.nx + x0);
        ok &= __float_as_uint(t00) == __float_as_uint(p[0]) && __float_as_uint(t10) == __float_as_uint(p[1]) &&
              __float_as_uint(t01) == __float_as_uint(p[c.nx]) && __float_as_uint(t11) == __float_as_uint(p[c.nx + 1]);
    }
    if (!ok) atomicOr(bad, 1);
}

// quad copy: quad[(z*ny + y)*nx + x] = {v[z][y][x], v[z][y][x+1], v[z][y+1][x], v[z][y+1][x+1]}, neighbours
// clamped at the last row / column (entries no in-bounds sample reads)
__global__ void __launch_bounds__(256) pack_quads_kernel(const float* __restrict__ v, int nx, int ny, int nz,
                                                         float4* __restrict__ quad) {
    const size_t n = (size_t)nx * ny * nz;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int y = (int)((t / nx) % ny);
        const int dx = (x + 1 < nx) ? 1 : 0;
        const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
        quad[t] = make_float4(__ldg(v + t), __ldg(v + t + dx), __ldg(v + t + dy), __ldg(v + t + dy + dx));
    }
}

// oct copy: oct[t] = {quad of layer z, quad of layer z + 1 (clamped at the last layer)}, 32 bytes per cell
__global__ void __launch_bounds__(256) pack_octs_kernel(const float* __restrict__ v, int nx, int ny, int nz,
                                                        float4* __restrict__ oct, int mark) {
    const size_t n = (size_t)nx * ny * nz;
    const size_t slab = (size_t)nx * ny;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int y = (int)((t / nx) % ny);
        const int z = (int)(t / slab);
        const int dx = (x + 1 < nx) ? 1 : 0;
        const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
        const size_t u = t + ((z + 1 < nz) ? slab : 0);
        float4 lo = make_float4(__ldg(v + t), __ldg(v + t + dx), __ldg(v + t + dy), __ldg(v + t + dy + dx));
        const float4 hi = make_float4(__ldg(v + u), __ldg(v + u + dx), __ldg(v + u + dy), __ldg(v + u + dy + dx));
        // (CubeView::oct_marked) a record the three-operation lerp cannot interpolate exactly gets a NaN first corner
        if (mark) {
            const int cls = oct_record_class(lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w);
            if (cls == 1) lo.x = -lo.x;
            if (cls == 2) lo.x = __int_as_float(0x7fc00000);
        }
        oct[2 * t] = lo;
        oct[2 * t + 1] = hi;
    }
}

// upload-time self-check of the oct copy, every entry, bit for bit, through the load the march uses
__global__ void __launch_bounds__(256) verify_octs_kernel(const float* __restrict__ v, int nx, int ny, int nz,
                                                          const float* __restrict__ oct, int* bad, int mark) {
    const size_t n = (size_t)nx * ny * nz;
    const size_t slab = (size_t)nx * ny;
    bool ok = true;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int y = (int)((t / nx) % ny);
        const int z = (int)(t / slab);
        const int dx = (x + 1 < nx) ? 1 : 0;
        const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
        const size_t u = t + ((z + 1 < nz) ? slab : 0);
        float a0, a1, a2, a3, b0, b1, b2, b3;
        load_oct(oct + 8 * t, a0, a1, a2, a3, b0, b1, b2, b3);
        // the first corner of a record that must carry the mark is the quiet NaN, nothing else
        const int cls = mark ? oct_record_class(v[t], v[t + dx], v[t + dy], v[t + dy + dx], v[u], v[u + dx], v[u + dy],
                                                v[u + dy + dx]) : 0;
        const unsigned want0 = cls == 2 ? 0x7fc00000u : cls == 1 ? (__float_as_uint(v[t]) | 0x80000000u) : __float_as_uint(v[t]);
        ok &= __float_as_uint(a0) == want0 && __float_as_uint(a1) == __float_as_uint(v[t + dx]) &&
              __float_as_uint(a2) == __float_as_uint(v[t + dy]) && __float_as_uint(a3) == __float_as_uint(v[t + dy + dx]) &&
              __float_as_uint(b0) == __float_as_uint(v[u]) && __float_as_uint(b1) == __float_as_uint(v[u + dx]) &&
              __float_as_uint(b2) == __float_as_uint(v[u + dy]) && __float_as_uint(b3) == __float_as_uint(v[u + dy + dx]);
    }
    if (!ok) atomicOr(bad, 1);
}

// upload-time self-check of the quad copy, every entry, bit for bit, through the load the march uses
__global__ void __launch_bounds__(256) verify_quads_kernel(const float* __restrict__ v, int nx, int ny, int nz,
                                                           const float4* __restrict__ quad, int* bad) {
    const size_t n = (size_t)nx * ny * nz;
    bool ok = true;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int y = (int)((t / nx) % ny);
        const int dx = (x + 1 < nx) ? 1 : 0;
        const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
        const float4 q = __ldg(quad + t);
        ok &= __float_as_uint(q.x) == __float_as_uint(v[t]) && __float_as_uint(q.y) == __float_as_uint(v[t + dx]) &&
              __float_as_uint(q.z) == __float_as_uint(v[t + dy]) && __float_as_uint(q.w) == __float_as_uint(v[t + dy + dx]);
    }
    if (!ok) atomicOr(bad, 1);
}

// ---- upload-time analysis kernels -------------------------------------------------------
// flag[0] |= 1 unless every voxel is finite and > 0; flag[0] |= 2 unless every voxel is finite and > 0 or +0;
// flag[1] / flag[2] = bit patterns of the smallest / largest finite positive voxel (positive floats order like
// their bit patterns; flag[1] starts at 0xffffffff, flag[2] at 0) -- the value range the lean march relies on
__global__ void __launch_bounds__(256) scan_positive_kernel(const float* __restrict__ v, size_t n, int* flag) {
    int bad = 0;
    unsigned lo = 0xffffffffu, hi = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = __ldg(v + i);
        const bool pos = x > 0.f && x < CUDART_INF_F;
        if (!pos) bad |= 1;
        if (!pos && __float_as_uint(x) != 0u) bad |= 2;
        if (pos) {
            lo = min(lo, __float_as_uint(x));
            hi = max(hi, __float_as_uint(x));
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicOr(flag, bad);
        atomicMin(reinterpret_cast<unsigned*>(flag) + 1, lo);
        atomicMax(reinterpret_cast<unsigned*>(flag) + 2, hi);
    }
}

// Exhaustive proof for one spacing: for EVERY fp32 ingress a of the hot window,
// 2^-100 <= a < limit (all bit patterns in [bits(2^-100), limit_bits)), do the cheap quotients
// equal the IEEE one?  bad[0] / bad[1] / bad[2] are set when div_spacing<0> / <1> / <3> (the multiply alone) ever
// differs from __fdiv_rn.
__global__ void __launch_bounds__(256) verify_div_kernel(float s, float r, unsigned limit_bits, int* bad) {
    bool b0 = false, b1 = false, b3 = false;
    const unsigned long long first = __float_as_uint(BIRP_TINY_INGRESS);
    for (unsigned long long u = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < limit_bits;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((unsigned)u);
        const unsigned want = __float_as_uint(__fdiv_rn(a, s));
        b0 |= __float_as_uint(div_spacing<0>(a, s, r)) != want;
        b1 |= __float_as_uint(div_spacing<1>(a, s, r)) != want;
        b3 |= __float_as_uint(div_spacing<3>(a, s, r)) != want;
    }
    if (__any_sync(0xffffffffu, b0) && (threadIdx.x & 31) == 0) atomicOr(bad + 0, 1);
    if (__any_sync(0xffffffffu, b1) && (threadIdx.x & 31) == 0) atomicOr(bad + 1, 1);
    if (__any_sync(0xffffffffu, b3) && (threadIdx.x & 31) == 0) atomicOr(bad + 2, 1);
}

// ---- per-frame setup on the device ----------------------------------------------
// The math provider of birp_frame_math.h for the device: every transcendental in fp64, rounded
// once to fp32 (glibc's fp32 functions, which the host instantiation and the reference call, are
// correctly rounded in the vast majority of cases; the two agree except within ~1e-8 of an fp32
// rounding boundary).  sqrt and the arithmetic between the calls are IEEE fp32 in both.
struct DeviceFrameMath {
    static __host__ __device__ float sqrt(float x) { return sqrtf(x); }
    static __host__ __device__ float sin(float x) { return (float)::sin((double)x); }
    static __host__ __device__ float cos(float x) { return (float)::cos((double)x); }
    static __host__ __device__ float acos(float x) { return (float)::acos((double)x); }
    static __host__ __device__ float asin(float x) { return (float)::asin((double)x); }
    static __host__ __device__ float atan2(float y, float x) { return (float)::atan2((double)y, (double)x); }
    static __host__ __device__ float abs(float x) { return fabsf(x); }
};

// one thread per pose: Camera::Integrate's prologue + Camera::Orient (src/Camera.cpp:85-137,309-508)
__global__ void __launch_bounds__(128) frame_setup_kernel(const float* __restrict__ poses, int n, float aim_x,
                                                          float tan_half_x, float tan_half_y, float aspect,
                                                          birp_frame* __restrict__ out, int* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pos[3] = {poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]};
    birp_frame fr;
    const float aim[3] = {aim_x, 0.f, 0.f};
    const int rc = fm::frame_setup_T<DeviceFrameMath>(pos, aim, tan_half_x, tan_half_y, aspect, &fr);
    if (rc != BIRP_OK) {
        atomicOr(status, kStatusSingular);
        return;
    }
    out[i] = fr;
}

// n camera positions, uniform in time, on a Kepler ellipse (fp64, Newton on Kepler's equation, in
// the operation order of birp_cpp_b200/synthetic.py: orbit_poses): pos = -xp * u + yp * w
__global__ void __launch_bounds__(128) orbit_poses_kernel(int n, double a, double e, double b_over_a, double ux,
                                                          double uy, double uz, double wx, double wy, double wz,
                                                          float* __restrict__ poses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double M = 2.0 * BIRP_PI * ((double)i + 0.5) / (double)n;
    double E = M;
    for (int it = 0; it < 60; ++it) E -= (E - e * sin(E) - M) / (1.0 - e * cos(E));
    const double xp = a * (cos(E) - e);
    const double yp = a * b_over_a * sin(E);
    poses[3 * i + 0] = (float)((-xp) * ux + yp * wx);
    poses[3 * i + 1] = (float)((-xp) * uy + yp * wy);
    poses[3 * i + 2] = (float)((-xp) * uz + yp * wz);
}

// K4: one thread per ray, samples[ray][k] (row-major, as the host sampled them)
__global__ void __launch_bounds__(256) accumulate_kernel(const float* __restrict__ samples, long nrays,
                                                         int nsamples, float ray_width, int accumulate,
                                                         float* __restrict__ image) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrays) return;
    float acc = accumulate ? image[r] : 0.f;
    const float* s = samples + r * (long)nsamples;
    for (int k = 0; k < nsamples; ++k) {
        const float v = scale_sample(__ldg(s + k), ray_width);
        if (v < 0.f) break;
        acc = __fadd_rn(acc, v);
    }
    image[r] = acc;
}

template <bool TRILINEAR>
__global__ void __launch_bounds__(256) sample_cube_kernel(const CubeView c, const float* __restrict__ xyz,
                                                          long n, float* __restrict__ out, int* status) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    bool nan_index = false;
    out[p] = cube_sample<TRILINEAR>(c, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], nan_index);
    if (nan_index) atomicOr(status, kStatusNanIndex);
}

template <int MODE>
__global__ void __launch_bounds__(256) sample_cmem_kernel(const CmemDev D, const float* __restrict__ xyz, long n,
                                                          float* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    out[p] = cmem_sample_mode<MODE>(D, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]);
}

__global__ void __launch_bounds__(256) ray_directions_kernel(const birp_frame* __restrict__ frame, int W, int H,
                                                             float* __restrict__ dirs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    float dx, dy, dz;
    pixel_direction(*frame, p % W, p / W, (float)W, (float)H, dx, dy, dz);
    dirs[3 * p] = dx;
    dirs[3 * p + 1] = dy;
    dirs[3 * p + 2] = dz;
}

// max error of the birp::cm fp64 routines against the CUDA fp64 library on pseudo-random arguments from
// CMEM's ranges; err[6] = {exp2, log2, pow, exp (relative), acos (absolute), angle bundle (absolute: sin, cos,
// cos(a/2) of the rounded angle, and how far the fp32 angle is from a correct rounding of acos)}
// (double bit patterns reduced with atomicMax)
__global__ void __launch_bounds__(256) cmem_math_selftest_kernel(long n, unsigned long long* err) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned h = (unsigned)t * 2654435761u + 99u;
    auto next01 = [&]() {
        h ^= h >> 16;
        h *= 0x85ebca6bu;
        h ^= h >> 13;
        h *= 0xc2b2ae35u;
        h ^= h >> 16;
        return (double)h * (1.0 / 4294967296.0);
    };
    auto rel = [](double got, double want) {
        const double d = fabs(got - want);
        return want != 0.0 ? d / fabs(want) : d;
    };
    double e[6];
    const double x = (next01() - 0.5) * 250.0;  // [-125, 125]
    e[0] = rel(cm::exp2_d(x), exp2(x));
    const double p = exp2((next01() - 0.5) * 250.0);
    e[1] = fabs(cm::log2_d(p) - log2(p)) / fmax(fabs(log2(p)), 1.0);
    const double y = (next01() - 0.5) * 8.0;
    const double base = exp2((next01() - 0.5) * 24.0);
    e[2] = rel(cm::pow_pos_d(base, y), pow(base, y));
    const double xe = (next01() - 0.5) * 80.0;  // [-40, 40]
    e[3] = rel(cm::exp_d(xe), exp(xe));
    // acos on [-1, 1], with extra weight on the ends and on the +-1/2 seams
    double u = 2.0 * next01() - 1.0;
    const unsigned pick = h & 7u;
    if (pick == 0) u = 1.0 - next01() * 1e-6;
    if (pick == 1) u = -1.0 + next01() * 1e-6;
    if (pick == 2) u = 0.5 + (next01() - 0.5) * 1e-6;
    if (pick == 3) u = -0.5 + (next01() - 0.5) * 1e-6;
    e[4] = fabs(cm::acos_d(u) - acos(u));
    const float qf = (float)u;
    const cm::Angle A = cm::angle_from_cosine(qf);
    const double a = (double)A.a, want_a = acos((double)qf);
    const double ulp = (double)__uint_as_float(__float_as_uint(fmaxf(A.a, 1e-30f)) + 1u) - (double)fmaxf(A.a, 1e-30f);
    double ea = fmax(fabs(A.s - sin(a)), fabs(A.c - cos(a)));
    ea = fmax(ea, fabs(A.c_half - cos(0.5 * a)));
    ea = fmax(ea, fmax(0.0, fabs(a - want_a) - 0.5 * ulp));
    e[5] = ea;
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicMax(err + k, (unsigned long long)__double_as_longlong(e[k]));
}

// How much of the region filter's margin do the fp32 radius estimates actually use?  n pseudo-random sample
// positions (a box of +-35 RE, plus positions squeezed towards the planes / axes where an angle degenerates):
// out[0] / out[1] = max over the trusted directions of |estimate - exact| / margin for the magnetopause / the
// bow-shock radius, out[2] = fraction of positions whose direction is trusted (double bit patterns, atomicMax;
// out[2] accumulates a count)
__global__ void __launch_bounds__(256) cmem_filter_selftest_kernel(const CmemDev D, long n, unsigned long long* out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned h = (unsigned)t * 2654435761u + 12345u;
    auto next01 = [&]() {
        h ^= h >> 16;
        h *= 0x85ebca6bu;
        h ^= h >> 13;
        h *= 0xc2b2ae35u;
        h ^= h >> 16;
        return (float)h * (1.0f / 4294967296.0f);
    };
    float x = (next01() - 0.5f) * 70.f, y = (next01() - 0.5f) * 70.f, z = (next01() - 0.5f) * 70.f;
    const unsigned pick = h & 15u;  // squeeze some positions towards degenerate directions
    const float tiny = exp2f(-20.f * next01());
    if (pick == 0) y *= tiny;
    if (pick == 1) z *= tiny;
    if (pick == 2) x *= tiny;
    if (pick == 3) y *= tiny, z *= tiny;
    if (pick == 4) x = -fabsf(x), y *= 0.2f * tiny, z *= 0.2f * tiny;  // around the -x axis guard
    float r, q_t, q_p, re, qe_t, qe_p;
    cmem_cosines(x, y, z, r, q_t, q_p);
    cmem_cosines_estimate(x, y, z, re, qe_t, qe_p);
    CmemEstimate est = cmem_estimate_stage1(D, qe_t, qe_p);
    const bool trusted1 = est.trusted;  // what the stage-1 decision (beyond the bow shock) relies on
    const float mp_e = cmem_estimate_stage2(D, est);
    const float bs_e = est.bs, up_e = est.mp_upper;
    if (!trusted1) return;
    float theta, sin_t, mp, bs;
    cmem_radii_v4(D, q_t, q_p, theta, sin_t, mp, bs);
    // the filter compares the ESTIMATED r with the estimated radii: the error of r counts against the margin too
    const double dr = fabs((double)re - (double)r);
    double u_mp = est.trusted ? (fabs((double)mp_e - (double)mp) + dr) / (double)BIRP_CMEM_MARGIN(mp_e) : 0.0;
    const double u_bs = (fabs((double)bs_e - (double)bs) + dr) / (double)BIRP_CMEM_MARGIN(bs_e);
    // stage 1 uses p0 r0 f as an upper bound of the magnetopause radius: how far below the exact radius can it fall?
    if (D.q_nonpositive) u_mp = fmax(u_mp, ((double)mp - (double)up_e + dr) / (double)BIRP_CMEM_MARGIN(up_e));
    // a NaN on either side would make the filter's comparisons false (full evaluation); count it as unused margin
    if (u_mp == u_mp) atomicMax(out + 0, (unsigned long long)__double_as_longlong(u_mp));
    if (u_bs == u_bs) atomicMax(out + 1, (unsigned long long)__double_as_longlong(u_bs));
    atomicAdd(out + 2, 1ull);
}


// ---- CMEM region table (CmemDev::table): one thread per cell (cmem_table_cell, birp_device.cuh) -----------------
__global__ void __launch_bounds__(128) cmem_table_kernel(const CmemDev D, float4* __restrict__ table) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= kCmemTableT * kCmemTableP) return;
    float out[4];
    cmem_table_cell(D, cell / kCmemTableP, cell % kCmemTableP, out);
    table[cell] = make_float4(out[0], out[1], out[2], out[3]);
}

// the march's sampler with the table (mode 3 with D.table) against the unfiltered evaluation (mode 2) on n pseudo-random
// positions (the generator of cmem_filter_selftest_kernel): out[0] = samples that differ bitwise (must be 0),
// out[1] = samples the table decided
__global__ void __launch_bounds__(256) cmem_table_selftest_kernel(const CmemDev D, long n, unsigned long long* out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned h = (unsigned)t * 2654435761u + 777u;
    auto next01 = [&]() {
        h ^= h >> 16;
        h *= 0x85ebca6bu;
        h ^= h >> 13;
        h *= 0xc2b2ae35u;
        h ^= h >> 16;
        return (float)h * (1.0f / 4294967296.0f);
    };
    float x = (next01() - 0.5f) * 70.f, y = (next01() - 0.5f) * 70.f, z = (next01() - 0.5f) * 70.f;
    const unsigned pick = h & 15u;
    const float tiny = exp2f(-20.f * next01());
    if (pick == 0) y *= tiny;
    if (pick == 1) z *= tiny;
    if (pick == 2) x *= tiny;
    if (pick == 3) y *= tiny, z *= tiny;
    if (pick == 4) x = -fabsf(x), y *= 0.2f * tiny, z *= 0.2f * tiny;
    if (pick == 5) {  // on a ray through the origin region, at a radius near the table's bounds: scale towards 8 .. 16 RE
        const float s = (8.f + 8.f * next01()) * rsqrtf(x * x + y * y + z * z);
        x *= s, y *= s, z *= s;
    }
    const float a = cmem_sample_v5(D, x, y, z);
    const float b = cmem_sample_v4(D, x, y, z);
    if (__float_as_uint(a) != __float_as_uint(b)) atomicAdd(out + 0, 1ull);
    float r, q_t, q_p;
    cmem_cosines_estimate(x, y, z, r, q_t, q_p);
    const float4 tb = D.table[cmem_table_index(q_t, q_p)];
    if (r < tb.x || (r > tb.y && (r > tb.w || r < tb.z))) atomicAdd(out + 1, 1ull);
}

// ---- completion flags between the GPUs of a box (one process per GPU, results in one GPU's memory) -----------------
// A rank's render kernel stores its pixels into the gathering rank's buffer over NVLink; "all ranks are done" used to be
// a 4-byte NCCL all-reduce (30-40 us: 4 % of a 0.9 ms step).  Now each rank's stream ends a step with signal_kernel --
// one store of the step number into its slot of a flag array in the gathering rank's memory, ordered after the render
// kernel's stores by the stream and a system-scope fence -- and the gathering rank's stream waits in wait_kernel until
// every slot holds the step number.  The wait gives up after ~2 s of clock (a rank died): the status word says so.
__global__ void signal_kernel(volatile unsigned* flag, unsigned value) {
    __threadfence_system();
    *flag = value;
    __threadfence_system();
}

__global__ void wait_kernel(const volatile unsigned* flags, int n, unsigned value, int* status) {
    const int t = threadIdx.x;
    if (t < n) {
        const long long t0 = clock64();
        while ((int)(flags[t] - value) < 0) {  // (wrap-safe: a slot may already hold a LATER step)
            __nanosleep(200);
            if (clock64() - t0 > 4000000000LL) {
                if (status) atomicOr(status, kStatusPeerTimeout);
                break;
            }
        }
    }
    __threadfence_system();
}

inline unsigned blocks_for(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

cudaError_t launch_render_cube(const CubeView& cube, const birp_frame* d_frames, const RenderGeom& g,
                               const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, int* d_status, cudaStream_t stream) {
    MarchArgs a;
    unsigned blocks;
    if (!fill_args(a, d_frames, g, d_dist, flags, d_images, d_valid, d_status, &blocks))
        return cudaErrorInvalidConfiguration;
    if (a.total_tiles <= 0) return cudaSuccess;
    const bool tri = (flags & BIRP_TRILINEAR) != 0;
    const int lerp = (flags & BIRP_GENERAL_LERP) ? 0 : cube.positive;
    int divmode = cube.divmode;  // 3 multiply only (power-of-two spacings), 0 one correction step, 1 two, 2 IEEE
    if ((flags & BIRP_DIV_TWO_STEP) && (divmode == 3 || divmode == 0)) divmode = 1;
    if (flags & BIRP_DIV_ONE_STEP && divmode == 3) divmode = 0;
    if (flags & BIRP_DIV_IEEE) divmode = 2;
    // corner fetch of the trilinear kernel (cube_sample_hot): what the cube has, minus what the flags forbid
    const bool has_tex = cube.tex != 0 && !(flags & BIRP_NO_TEXTURE);
    const bool has_quad = cube.quad != nullptr && !(flags & BIRP_NO_QUADS);
    // the run-end search needs a monotone ray table (RenderGeom::monotone_table, checked where the table is made)
    const bool span = g.monotone_table && !(flags & BIRP_WINDOW_PER_STEP);
    // (the oct kernel exists in the span form only; the oct copy of a cube outside the positive class carries marks
    // that only the default, mark-aware lerp form understands: the cross-check forms read the quad copy)
    const bool has_oct = cube.oct != nullptr && span &&
                         !(flags & (BIRP_NO_OCTS | BIRP_NO_QUADS | BIRP_QUADS_ONLY | BIRP_GENERAL_LERP)) &&
                         (cube.positive == 1) == (cube.oct_marked == 0);
    int fetch = 0;
    if (tri) {
        if (has_tex && (flags & BIRP_TEXTURE_ONLY)) fetch = 1;
        else if (has_oct) fetch = 5;
        else if (has_quad) fetch = 4;
        else if (has_tex) fetch = 2;
    }
    // BIRP_FAST_MARCH (trilinear, monotone ray table, poly copy present): the tolerance-mode march
    if (tri && (flags & BIRP_FAST_MARCH) && cube.poly != nullptr && g.monotone_table) {
        const bool signs = cube.positive == 0;
        // grids of more than 2^24 voxels need the integer cell index (variant 0); else g.fast_variant (BIRP_FAST_VARIANT
        // in the environment, experiments) or the measured default
        int variant = ((long long)cube.nx * cube.ny * cube.nz <= (1LL << 24)) ? g.fast_variant : 0;
        if (variant < 0 || variant > 4) variant = kFastVariantDefault;
        // the uniform table of Camera::GenerateRayDistWidth lets the march compute its distances (KSTEP); any other
        // table (none exists behind this API today) is read step by step
        const bool kstep = g.fast_step > 0.0 && g.fast_kstep != 0;
        a.fast_step = kstep ? g.fast_step : 0.0;
        if (g.fast_coop && kstep && !signs && variant != 0 && a.nsamples <= kSmemDist) {  // (experiment, see the kernel)
            march_cube_fast_coop_kernel<2><<<blocks, a.cta_warps * 32, 0, stream>>>(a, cube);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            return cudaGetLastError();
        }
#define BIRP_FAST_LAUNCH(S, V)                                                                       \
    if (kstep) march_cube_fast_kernel<S, V, true><<<blocks, a.cta_warps * 32, 0, stream>>>(a, cube);   \
    else march_cube_fast_kernel<S, V, false><<<blocks, a.cta_warps * 32, 0, stream>>>(a, cube)
#define BIRP_FAST_VARIANTS(S)                       \
    switch (variant) {                              \
        case 0: BIRP_FAST_LAUNCH(S, 0); break;      \
        case 1: BIRP_FAST_LAUNCH(S, 1); break;      \
        case 2: BIRP_FAST_LAUNCH(S, 2); break;      \
        case 3: BIRP_FAST_LAUNCH(S, 3); break;      \
        default: BIRP_FAST_LAUNCH(S, 4); break;     \
    }
        if (signs) {
            BIRP_FAST_VARIANTS(true)
        } else {
            BIRP_FAST_VARIANTS(false)
        }
#undef BIRP_FAST_VARIANTS
#undef BIRP_FAST_LAUNCH
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    CubeView view = cube;
    if (!tri) {  // nearest neighbour has no interpolation weight to protect: the first cells stay hot
        view.lox = BIRP_TINY_INGRESS;
        view.loy = view.lox;
        view.loz = view.lox;
    }
    // Lean march (the default where its preconditions hold; BIRP_NO_LEAN_MARCH keeps the general loop as the
    // cross-check): fp32 cell index (grids of at most 2^24 voxels) and no range test in the scale -- allowed when
    // every sample of the positive cube, which lies in [vmin, vmax] (scan_positive_kernel at upload), scales into
    // [2^-120, 2^120) (the scale is a monotone function of the sample for a positive ray width)
    bool lean = false;
    if (!(flags & BIRP_NO_LEAN_MARCH) && fetch == 5 && lerp == 1 && (long long)cube.nx * cube.ny * cube.nz <= (1LL << 24) &&
        cube.vmin > 0.f && cube.vmax >= cube.vmin && a.ray_width > 0.f) {
        const volatile float plo = cube.vmin * a.ray_width, phi = cube.vmax * a.ray_width;
        const float ylo = (float)((double)plo * BIRP_SCALE_K), yhi = (float)((double)phi * BIRP_SCALE_K);
        lean = ylo >= BIRP_SCALE_SAFE_LO && yhi < BIRP_SCALE_SAFE_HI;
    }
    cudaError_t e;
    switch (divmode) {
        case 0: e = launch_cube_div<0>(a, view, blocks, tri, lerp, fetch, span, lean, stream); break;
        case 1: e = launch_cube_div<1>(a, view, blocks, tri, lerp, fetch, span, lean, stream); break;
        case 3: e = launch_cube_div<3>(a, view, blocks, tri, lerp, fetch, span, lean, stream); break;
        default: e = launch_cube_div<2>(a, view, blocks, tri, lerp, fetch, span, lean, stream); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return e;
}

cudaError_t launch_scan_positive(const float* d_v, size_t n, int* d_flag, cudaStream_t stream) {
    scan_positive_kernel<<<148 * 8, 256, 0, stream>>>(d_v, n, d_flag);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_pack_quads(const CubeView& cube, float4* d_quad, int* d_bad, cudaStream_t stream) {
    pack_quads_kernel<<<148 * 8, 256, 0, stream>>>(cube.v, cube.nx, cube.ny, cube.nz, d_quad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    verify_quads_kernel<<<148 * 8, 256, 0, stream>>>(cube.v, cube.nx, cube.ny, cube.nz, d_quad, d_bad);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_pack_octs(const CubeView& cube, float* d_oct, int* d_bad, int mark, cudaStream_t stream) {
    pack_octs_kernel<<<148 * 8, 256, 0, stream>>>(cube.v, cube.nx, cube.ny, cube.nz, reinterpret_cast<float4*>(d_oct), mark);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    verify_octs_kernel<<<148 * 8, 256, 0, stream>>>(cube.v, cube.nx, cube.ny, cube.nz, d_oct, d_bad, mark);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_pack_poly(const CubeView& cube, float* d_poly, cudaStream_t stream) {
    pack_poly_kernel<<<148 * 8, 256, 0, stream>>>(cube.v, cube.nx, cube.ny, cube.nz, reinterpret_cast<float4*>(d_poly));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_verify_gather(const CubeView& cube, int ncells, int* d_bad, cudaStream_t stream) {
    verify_gather_kernel<<<(ncells + 255) / 256, 256, 0, stream>>>(cube, ncells, d_bad);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_verify_div(float s, float r, unsigned limit_bits, int* d_bad2, cudaStream_t stream) {
    verify_div_kernel<<<148 * 16, 256, 0, stream>>>(s, r, limit_bits, d_bad2);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_cmem_table(const birp_cmem_params& params, float4* d_table, cudaStream_t stream) {
    cmem_table_kernel<<<(kCmemTableT * kCmemTableP + 127) / 128, 128, 0, stream>>>(make_cmem_dev(params), d_table);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_cmem_table_selftest(const birp_cmem_params& params, const float4* d_table, long n,
                                       unsigned long long* d_out2, cudaStream_t stream) {
    CmemDev D = make_cmem_dev(params);
    D.table = d_table;
    cmem_table_selftest_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(D, n, d_out2);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_render_cmem(const birp_cmem_params& params, const birp_frame* d_frames,
                               const RenderGeom& g, const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, cudaStream_t stream, const float4* d_table) {
    if (flags & BIRP_CMEM_FAST) {  // tolerance mode: fp32 values behind the table and the filter
        CmemField<4> f{make_cmem_dev(params)};
        f.D.table = (flags & BIRP_CMEM_NO_TABLE) ? nullptr : d_table;
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
    }
    if (flags & BIRP_CMEM_LIBM32) {  // diagnostic: the fp32 math library for every sample, no filter
        CmemField<0> f{make_cmem_dev(params)};
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
    }
    if (flags & BIRP_CMEM_LIBM64) {
        CmemField<1> f{make_cmem_dev(params)};
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
    }
    if (flags & BIRP_CMEM_NO_FILTER) {
        CmemField<2> f{make_cmem_dev(params)};
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
    }
    CmemField<3> f{make_cmem_dev(params)};
    f.D.table = (flags & BIRP_CMEM_NO_TABLE) ? nullptr : d_table;
    return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
}

cudaError_t launch_peer_signal(unsigned* d_flag, unsigned value, cudaStream_t stream) {
    signal_kernel<<<1, 1, 0, stream>>>(d_flag, value);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_peer_wait(const unsigned* d_flags, int n, unsigned value, int* d_status, cudaStream_t stream) {
    wait_kernel<<<1, 32, 0, stream>>>(d_flags, n, value, d_status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_accumulate(const float* d_samples, long nrays, int nsamples, float ray_width, int flags,
                              float* d_image, cudaStream_t stream) {
    if (nrays <= 0) return cudaSuccess;
    accumulate_kernel<<<blocks_for(nrays, 256), 256, 0, stream>>>(d_samples, nrays, nsamples, ray_width,
                                                                  (flags & BIRP_ACCUMULATE) ? 1 : 0, d_image);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_frame_setup(const float* d_poses, int n, float aim_x, float tan_half_x, float tan_half_y,
                               float aspect, birp_frame* d_frames, int* d_status, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    frame_setup_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(d_poses, n, aim_x, tan_half_x, tan_half_y, aspect,
                                                              d_frames, d_status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_orbit_poses(int n, const double k[9], float* d_poses, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    orbit_poses_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(n, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7],
                                                              k[8], d_poses);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// div10 (birp_device.cuh) against the IEEE division for EVERY fp32 bit pattern: bad[0] counts the patterns where the form
// the CMEM kernel uses -- div10 for an ordinary r, the division otherwise -- differs from r / 10.f, bad[1] the positive
// normal r where div10 alone would (the subnormal quotients the range test keeps away from it)
__global__ void __launch_bounds__(256) div10_selftest_kernel(unsigned long long* bad) {
    unsigned long long n0 = 0, n1 = 0;
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < (1ull << 32);
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float r = __uint_as_float((unsigned)u);
        const unsigned want = __float_as_uint(__fdiv_rn(r, 10.f));
        const unsigned fast = __float_as_uint(div10(r));
        const bool nan_both = (r != r);  // any NaN payload is a NaN
        if (!nan_both && (cmem_r_is_ordinary(r) ? fast : want) != want) ++n0;
        if (u >= 0x00800000ull && u < 0x7f800000ull && fast != want) ++n1;  // positive normal r
    }
    if (n0) atomicAdd(bad + 0, n0);
    if (n1) atomicAdd(bad + 1, n1);
}

cudaError_t launch_div10_selftest(unsigned long long* d_bad2, cudaStream_t stream) {
    div10_selftest_kernel<<<148 * 16, 256, 0, stream>>>(d_bad2);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_cmem_math_selftest(long n, unsigned long long* d_err6, cudaStream_t stream) {
    cmem_math_selftest_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(n, d_err6);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_cmem_filter_selftest(const birp_cmem_params& params, long n, unsigned long long* d_out3,
                                        cudaStream_t stream) {
    cmem_filter_selftest_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(make_cmem_dev(params), n, d_out3);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_ray_directions(const birp_frame* d_frame, int W, int H, float* d_dirs, cudaStream_t stream) {
    ray_directions_kernel<<<blocks_for((long)W * H, 256), 256, 0, stream>>>(d_frame, W, H, d_dirs);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_sample_cube(const CubeView& cube, const float* d_xyz, long n, int flags, float* d_out,
                               int* d_status, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (flags & BIRP_TRILINEAR)
        sample_cube_kernel<true><<<blocks_for(n, 256), 256, 0, stream>>>(cube, d_xyz, n, d_out, d_status);
    else
        sample_cube_kernel<false><<<blocks_for(n, 256), 256, 0, stream>>>(cube, d_xyz, n, d_out, d_status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_sample_cmem(const birp_cmem_params& params, const float* d_xyz, long n, int flags,
                               float* d_out, cudaStream_t stream, const float4* d_table) {
    if (n <= 0) return cudaSuccess;
    CmemDev D = make_cmem_dev(params);
    D.table = (flags & BIRP_CMEM_NO_TABLE) ? nullptr : d_table;
    if (flags & BIRP_CMEM_FAST)
        sample_cmem_kernel<4><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    else if (flags & BIRP_CMEM_LIBM32)
        sample_cmem_kernel<0><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    else if (flags & BIRP_CMEM_LIBM64)
        sample_cmem_kernel<1><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    else if (flags & BIRP_CMEM_NO_FILTER)
        sample_cmem_kernel<2><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    else
        sample_cmem_kernel<3><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace birp
