// sm_100a kernels of the line-of-sight integrator.
//
//   march_cube_kernel<..> K1/K3: ray generation + fixed-step march + DataCube sampling (nearest /
//                         trilinear) + scale + terminate + accumulate, for a batch of frames
//                         (reference src/Camera.cpp:140-232 x src/DataCube.cpp:203-277).
//   march_kernel<Field>   K2: the same march around a field functor (CMEM, src/Cmem.cpp:117-288).
//   accumulate_kernel     K4: scale + terminate + accumulate of host-sampled values
//                         (reference src/Camera.cpp:201-223), for user-defined Space classes.
//   sample_*_kernel, ray_directions_kernel: Space::GetSample on points; the ray generator alone.
//   verify_div_kernel, verify_gather_kernel, scan_positive_kernel: upload-time proofs / analysis.
//
// Work decomposition of the march kernels: one warp owns a 4x8 / 8x4 pixel tile of one frame, one
// lane per ray.  The 32 rays of a tile stay within a few voxels of one another along the whole
// march, so the corner fetches of a step hit the same handful of cache lines / texture tiles, and
// the serial per-pixel accumulation order of the reference is kept.
#include <atomic>
#include <cmath>

#include "birp_internal.h"

namespace birp {

namespace {

std::atomic<unsigned long long> g_launches{0};

constexpr int kWarpsPerBlock = 8;
constexpr int kSmemDist = 2048;  // ray-table entries kept in shared memory by the cube kernel
constexpr int kBlockThreads = kWarpsPerBlock * 32;

struct MarchArgs {
    const birp_frame* __restrict__ frames;
    const float* __restrict__ dist;
    float* __restrict__ images;
    unsigned long long* valid;
    int* status;
    int W, H;
    int row_begin, rows;
    int nsamples;
    int tile_w_log2;  // a warp's 32 rays form a (1 << tile_w_log2) x (32 >> tile_w_log2) pixel tile
    int tiles_x, tiles_per_frame;
    long long total_tiles;
    float ray_width;
    int accumulate;
    // frame f of this launch is written to image slot out_offset + f * out_stride of `images`
    // (1-GPU: 0 / 1; multi-GPU sweeps scatter straight into the gathering rank's buffer)
    long long out_offset, out_stride;
};

template <bool PRECISE>
struct CmemField {
    CmemDev D;
    __device__ __forceinline__ float operator()(float x, float y, float z, bool&) const {
        return cmem_sample<PRECISE>(D, x, y, z);
    }
};

// Sample-independent CMEM terms, with the host libm the reference itself calls.
CmemDev make_cmem_dev(const birp_cmem_params& P) {
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
    return D;
}

// Which ray this lane owns; false when the lane falls outside the image / row band.
__device__ __forceinline__ bool lane_ray(const MarchArgs& a, int& f, int& i, int& j) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (tile >= a.total_tiles) return false;
    f = (int)(tile / a.tiles_per_frame);
    const int t = (int)(tile - (long long)f * a.tiles_per_frame);
    const int ty = t / a.tiles_x;
    const int tx = t - ty * a.tiles_x;
    i = (tx << a.tile_w_log2) + (lane & ((1 << a.tile_w_log2) - 1));
    const int jr = ty * (32 >> a.tile_w_log2) + (lane >> a.tile_w_log2);
    j = a.row_begin + jr;
    return i < a.W && jr < a.rows;
}

// Generic march: any Field functor (used for CMEM; the DataCube path has its own kernel below).
template <class Field>
__global__ void __launch_bounds__(kBlockThreads) march_kernel(const MarchArgs a, const Field field) {
    int f, i, j;
    if (!lane_ray(a, f, i, j)) return;
    const birp_frame fr = a.frames[f];
    float dx, dy, dz;
    pixel_direction(fr, i, j, (float)a.W, (float)a.H, dx, dy, dz);

    float* pix = a.images + ((size_t)(a.out_offset + f * a.out_stride) * a.H + j) * a.W + i;
    float acc = a.accumulate ? *pix : 0.f;
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
        s = scale_sample(s, ray_width);
        if (s < 0.f) break;  // src/Camera.cpp:207-210: the rest of the ray is ignored
        acc = __fadd_rn(acc, s);
        ++nvalid;
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
#define BIRP_CUBE_MIN_BLOCKS 6  // resident CTAs per SM the register allocation must allow (48 of 64 warps)
#endif

struct RayState {
    float dx, dy, dz;  // unit direction
    float px, py, pz;  // camera position
    float acc;         // pixel value so far
};

// March `count` steps whose distances sit in shared memory at byte address `sdist` (a 32-bit shared
// window address, advanced by 4 per step: one LDS + one IADD, no per-step address arithmetic).
// Returns the number of samples accumulated; `alive` is cleared when the ray ended.
template <bool TRILINEAR, int DIVMODE, bool POSITIVE, int TEXGATHER>
__device__ __forceinline__ int march_steps(const CubeView& c, RayState& r, unsigned sdist, int count, float ray_width,
                                           int* status, bool& alive) {
    int k = 0;
    for (; k < count; ++k, sdist += 4) {
        float tk;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(tk) : "r"(sdist));
        const float x = __fadd_rn(__fmul_rn(r.dx, tk), r.px);
        const float y = __fadd_rn(__fmul_rn(r.dy, tk), r.py);
        const float z = __fadd_rn(__fmul_rn(r.dz, tk), r.pz);
        float s;
        if (!cube_sample_inb<TRILINEAR, DIVMODE, POSITIVE, TEXGATHER>(c, x, y, z, s)) {
            // outside the hot window: almost always a ray leaving the cube (the reference returns
            // -1, which ends the ray); also NaN coordinates and ingresses below 2^-100.  The
            // generic transcription of DataCube::GetSample decides.
            bool nan_index = false;
            s = cube_sample<TRILINEAR>(c, x, y, z, nan_index);
            if (nan_index) atomicOr(status, kStatusNanIndex);
        }
        s = scale_sample(s, ray_width);
        if (s < 0.f) {  // src/Camera.cpp:207-210: the rest of the ray is ignored
            alive = false;
            break;
        }
        r.acc = __fadd_rn(r.acc, s);
    }
    return k;
}

template <bool TRILINEAR, int DIVMODE, bool POSITIVE, int TEXGATHER>
__global__ void __launch_bounds__(kBlockThreads, BIRP_CUBE_MIN_BLOCKS) march_cube_kernel(const MarchArgs a, const CubeView c) {
    // The ray table is staged through shared memory in chunks of kSmemDist distances (one chunk
    // for any march up to 2048 samples), so a step costs one LDS instead of a global load.
    __shared__ float s_dist[kSmemDist];
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_dist);
    const int first_chunk = min(kSmemDist, a.nsamples);
    for (int k = threadIdx.x; k < first_chunk; k += kBlockThreads) s_dist[k] = __ldg(a.dist + k);
    __syncthreads();

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
    int nlive = 0;
    // Keep the per-axis cube constants in ordinary registers for the whole march: left to itself ptxas
    // re-loads ~10 of them from the constant bank in every step, and issue slots are the scarce resource
    // here (+7 % measured).  The empty asm makes each value opaque, so it cannot be rematerialised.
    CubeView cc = c;
    asm volatile("" : "+f"(cc.ox), "+f"(cc.oy), "+f"(cc.oz), "+f"(cc.lx), "+f"(cc.ly), "+f"(cc.lz));
    asm volatile("" : "+f"(cc.rx), "+f"(cc.ry), "+f"(cc.rz), "+f"(cc.sx), "+f"(cc.sy), "+f"(cc.sz));
    asm volatile("" : "+r"(cc.nx), "+r"(cc.ny), "+l"(cc.v));
    if (alive) nlive = march_steps<TRILINEAR, DIVMODE, POSITIVE, TEXGATHER>(cc, r, s_base, first_chunk, a.ray_width, a.status, alive);
    // marches longer than the staged table: refill it chunk by chunk (every thread of the CTA takes part)
    for (int base = kSmemDist; base < a.nsamples; base += kSmemDist) {
        const int chunk = min(kSmemDist, a.nsamples - base);
        __syncthreads();
        for (int k = threadIdx.x; k < chunk; k += kBlockThreads) s_dist[k] = __ldg(a.dist + base + k);
        __syncthreads();
        if (alive) nlive += march_steps<TRILINEAR, DIVMODE, POSITIVE, TEXGATHER>(cc, r, s_base, chunk, a.ray_width, a.status, alive);
    }
    if (owns_pixel) {
        *pix = r.acc;
        if (a.valid) atomicAdd(a.valid, (unsigned long long)nlive);
    }
}

// Tile shape: split the warp's 32 rays so that the tile is as square as possible in ANGLE
// (the SXI pixel is 1.7x wider than tall once the reference's extra aspect factor is applied):
// compact ray bundles touch fewer cache lines per corner load.
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
               float* d_images, unsigned long long* d_valid, int* d_status, unsigned* blocks) {
    a.frames = d_frames;
    a.dist = d_dist;
    a.images = d_images;
    a.valid = d_valid;
    a.status = d_status;
    a.W = g.W;
    a.H = g.H;
    a.row_begin = g.row_begin;
    a.rows = g.row_end - g.row_begin;
    a.nsamples = g.nsamples;
    a.tile_w_log2 = pick_tile_w_log2(g);
    const int tw = 1 << a.tile_w_log2, th = 32 >> a.tile_w_log2;
    a.tiles_x = (g.W + tw - 1) / tw;
    a.tiles_per_frame = a.tiles_x * ((a.rows + th - 1) / th);
    a.total_tiles = (long long)a.tiles_per_frame * g.nframes;
    a.ray_width = g.ray_width;
    a.accumulate = (flags & BIRP_ACCUMULATE) ? 1 : 0;
    a.out_offset = g.out_frame_offset;
    a.out_stride = g.out_frame_stride;
    const long long nb = (a.total_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    *blocks = (unsigned)nb;
    return nb <= 0x7fffffffLL;
}

template <class Field>
cudaError_t launch_march(const Field& field, const birp_frame* d_frames, const RenderGeom& g,
                         const float* d_dist, int flags, float* d_images, unsigned long long* d_valid,
                         int* d_status, cudaStream_t stream) {
    MarchArgs a;
    unsigned blocks;
    if (!fill_args(a, d_frames, g, d_dist, flags, d_images, d_valid, d_status, &blocks))
        return cudaErrorInvalidConfiguration;
    if (a.total_tiles <= 0) return cudaSuccess;
    march_kernel<Field><<<blocks, kBlockThreads, 0, stream>>>(a, field);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <bool TRI, int DIV, bool POS, int TEX>
cudaError_t launch_cube_variant(const MarchArgs& a, const CubeView& c, unsigned blocks, cudaStream_t stream) {
    march_cube_kernel<TRI, DIV, POS, TEX><<<blocks, kBlockThreads, 0, stream>>>(a, c);
    return cudaGetLastError();
}

template <int DIV>
cudaError_t launch_cube_div(const MarchArgs& a, const CubeView& c, unsigned blocks, bool tri, bool pos, int tex,
                            cudaStream_t stream) {
    if (!tri) return launch_cube_variant<false, DIV, false, 0>(a, c, blocks, stream);
    if (pos) {
        if (tex == 2) return launch_cube_variant<true, DIV, true, 2>(a, c, blocks, stream);
        if (tex == 1) return launch_cube_variant<true, DIV, true, 1>(a, c, blocks, stream);
        return launch_cube_variant<true, DIV, true, 0>(a, c, blocks, stream);
    }
    if (tex == 2) return launch_cube_variant<true, DIV, false, 2>(a, c, blocks, stream);
    if (tex == 1) return launch_cube_variant<true, DIV, false, 1>(a, c, blocks, stream);
    return launch_cube_variant<true, DIV, false, 0>(a, c, blocks, stream);
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

// ---- upload-time analysis kernels -------------------------------------------------------
// flag |= 1 unless every voxel is finite and > 0
__global__ void __launch_bounds__(256) scan_positive_kernel(const float* __restrict__ v, size_t n, int* flag) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = __ldg(v + i);
        bad |= !(x > 0.f && x < CUDART_INF_F);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Exhaustive proof for one spacing: for EVERY fp32 ingress a of the hot window,
// 2^-100 <= a < limit (all bit patterns in [bits(2^-100), limit_bits)), do the cheap quotients
// equal the IEEE one?  bad[0] / bad[1] are set when div_spacing<0> / div_spacing<1> ever differs
// from __fdiv_rn.
__global__ void __launch_bounds__(256) verify_div_kernel(float s, float r, unsigned limit_bits, int* bad) {
    bool b0 = false, b1 = false;
    const unsigned long long first = __float_as_uint(BIRP_TINY_INGRESS);
    for (unsigned long long u = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < limit_bits;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((unsigned)u);
        const unsigned want = __float_as_uint(__fdiv_rn(a, s));
        b0 |= __float_as_uint(div_spacing<0>(a, s, r)) != want;
        b1 |= __float_as_uint(div_spacing<1>(a, s, r)) != want;
    }
    if (__any_sync(0xffffffffu, b0) && (threadIdx.x & 31) == 0) atomicOr(bad + 0, 1);
    if (__any_sync(0xffffffffu, b1) && (threadIdx.x & 31) == 0) atomicOr(bad + 1, 1);
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

template <bool PRECISE>
__global__ void __launch_bounds__(256) sample_cmem_kernel(const CmemDev D, const float* __restrict__ xyz, long n,
                                                          float* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    out[p] = cmem_sample<PRECISE>(D, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]);
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

// max relative error of the birp::cm fp64 routines against the CUDA fp64 library on pseudo-random
// arguments from CMEM's ranges; err[3] = {exp2, log2, pow} (double bit patterns reduced with atomicMax)
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
    double e[3];
    const double x = (next01() - 0.5) * 250.0;  // [-125, 125]
    e[0] = rel(cm::exp2_d(x), exp2(x));
    const double p = exp2((next01() - 0.5) * 250.0);
    e[1] = fabs(cm::log2_d(p) - log2(p)) / fmax(fabs(log2(p)), 1.0);
    const double y = (next01() - 0.5) * 8.0;
    const double base = exp2((next01() - 0.5) * 24.0);
    e[2] = rel(cm::pow_pos_d(base, y), pow(base, y));
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicMax(err + k, (unsigned long long)__double_as_longlong(e[k]));
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
    const bool pos = cube.positive != 0 && !(flags & BIRP_GENERAL_LERP);
    int divmode = cube.divmode;
    if ((flags & BIRP_DIV_TWO_STEP) && divmode < 1) divmode = 1;
    if (flags & BIRP_DIV_IEEE) divmode = 2;
    int tex = 0;  // corner fetch: 0 eight loads, 1 two texture gathers, 2 one gather + four loads
    if (tri && cube.tex != 0 && !(flags & BIRP_NO_TEXTURE)) tex = (flags & BIRP_TEXTURE_ONLY) ? 1 : 2;
    cudaError_t e;
    switch (divmode) {
        case 0: e = launch_cube_div<0>(a, cube, blocks, tri, pos, tex, stream); break;
        case 1: e = launch_cube_div<1>(a, cube, blocks, tri, pos, tex, stream); break;
        default: e = launch_cube_div<2>(a, cube, blocks, tri, pos, tex, stream); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return e;
}

cudaError_t launch_scan_positive(const float* d_v, size_t n, int* d_flag, cudaStream_t stream) {
    scan_positive_kernel<<<148 * 8, 256, 0, stream>>>(d_v, n, d_flag);
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

cudaError_t launch_render_cmem(const birp_cmem_params& params, const birp_frame* d_frames,
                               const RenderGeom& g, const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, cudaStream_t stream) {
    if (flags & BIRP_CMEM_FAST) {
        CmemField<false> f{make_cmem_dev(params)};
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
    }
    CmemField<true> f{make_cmem_dev(params)};
    return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
}

cudaError_t launch_accumulate(const float* d_samples, long nrays, int nsamples, float ray_width, int flags,
                              float* d_image, cudaStream_t stream) {
    if (nrays <= 0) return cudaSuccess;
    accumulate_kernel<<<blocks_for(nrays, 256), 256, 0, stream>>>(d_samples, nrays, nsamples, ray_width,
                                                                  (flags & BIRP_ACCUMULATE) ? 1 : 0, d_image);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_cmem_math_selftest(long n, unsigned long long* d_err6, cudaStream_t stream) {
    cmem_math_selftest_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(n, d_err6);
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
                               float* d_out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const CmemDev D = make_cmem_dev(params);
    if (flags & BIRP_CMEM_FAST)
        sample_cmem_kernel<false><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    else
        sample_cmem_kernel<true><<<blocks_for(n, 256), 256, 0, stream>>>(D, d_xyz, n, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace birp
