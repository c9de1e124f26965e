// sm_100a kernels of the line-of-sight integrator.
//
//   march_kernel<Field>   K1/K2/K3: ray generation + fixed-step march + field evaluation
//                         (DataCube gather or CMEM) + scale + terminate + accumulate, for a
//                         batch of frames (reference src/Camera.cpp:140-232).
//   accumulate_kernel     K4: scale + terminate + accumulate of host-sampled values
//                         (reference src/Camera.cpp:201-223), for user-defined Space classes.
//   sample_*_kernel       Space::GetSample on a list of points.
//
// Work decomposition of march_kernel: one warp owns an 8x4 pixel tile of one frame, one
// lane per ray.  The 32 rays of a tile stay within a few voxels of one another along the
// whole march, so the eight corner loads of a step hit the same handful of 128 B lines
// (L1 broadcast) and the serial per-pixel accumulation order of the reference is kept.
#include <atomic>

#include "birp_internal.h"

namespace birp {

namespace {

std::atomic<unsigned long long> g_launches{0};

constexpr int kTileW = 8;
constexpr int kTileH = 4;
constexpr int kWarpsPerBlock = 8;
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
    int tiles_x, tiles_per_frame;
    long long total_tiles;
    float ray_width;
    int accumulate;
};

template <bool TRILINEAR>
struct CubeField {
    CubeView c;
    __device__ __forceinline__ float operator()(float x, float y, float z, bool& nan_index) const {
        return cube_sample<TRILINEAR>(c, x, y, z, nan_index);
    }
};

template <bool PRECISE>
struct CmemField {
    birp_cmem_params P;
    __device__ __forceinline__ float operator()(float x, float y, float z, bool&) const {
        return cmem_sample<PRECISE>(P, x, y, z);
    }
};

template <class Field>
__global__ void __launch_bounds__(kBlockThreads) march_kernel(const MarchArgs a, const Field field) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (tile >= a.total_tiles) return;
    const int f = (int)(tile / a.tiles_per_frame);
    const int t = (int)(tile - (long long)f * a.tiles_per_frame);
    const int ty = t / a.tiles_x;
    const int tx = t - ty * a.tiles_x;
    const int i = tx * kTileW + (lane & (kTileW - 1));
    const int jr = ty * kTileH + (lane / kTileW);
    if (i >= a.W || jr >= a.rows) return;
    const int j = a.row_begin + jr;

    const birp_frame fr = a.frames[f];
    float dx, dy, dz;
    pixel_direction(fr, i, j, (float)a.W, (float)a.H, dx, dy, dz);

    float* pix = a.images + ((size_t)f * a.H + j) * a.W + i;
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
    if (nan_index) atomicOr(a.status, kStatusNanIndex);
    if (a.valid) atomicAdd(a.valid, (unsigned long long)nvalid);
}

template <class Field>
cudaError_t launch_march(const Field& field, const birp_frame* d_frames, const RenderGeom& g,
                         const float* d_dist, int flags, float* d_images, unsigned long long* d_valid,
                         int* d_status, cudaStream_t stream) {
    MarchArgs a;
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
    a.tiles_x = (g.W + kTileW - 1) / kTileW;
    a.tiles_per_frame = a.tiles_x * ((a.rows + kTileH - 1) / kTileH);
    a.total_tiles = (long long)a.tiles_per_frame * g.nframes;
    a.ray_width = g.ray_width;
    a.accumulate = (flags & BIRP_ACCUMULATE) ? 1 : 0;
    if (a.total_tiles <= 0) return cudaSuccess;
    const long long blocks = (a.total_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    march_kernel<Field><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(a, field);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
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
__global__ void __launch_bounds__(256) sample_cmem_kernel(const birp_cmem_params P,
                                                          const float* __restrict__ xyz, long n,
                                                          float* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    out[p] = cmem_sample<PRECISE>(P, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]);
}

inline unsigned blocks_for(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

cudaError_t launch_render_cube(const CubeView& cube, const birp_frame* d_frames, const RenderGeom& g,
                               const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, int* d_status, cudaStream_t stream) {
    if (flags & BIRP_TRILINEAR) {
        CubeField<true> f{cube};
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, d_status, stream);
    }
    CubeField<false> f{cube};
    return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, d_status, stream);
}

cudaError_t launch_render_cmem(const birp_cmem_params& params, const birp_frame* d_frames,
                               const RenderGeom& g, const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, cudaStream_t stream) {
    if (flags & BIRP_CMEM_FAST) {
        CmemField<false> f{params};
        return launch_march(f, d_frames, g, d_dist, flags, d_images, d_valid, nullptr, stream);
    }
    CmemField<true> f{params};
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
    if (flags & BIRP_CMEM_FAST)
        sample_cmem_kernel<false><<<blocks_for(n, 256), 256, 0, stream>>>(params, d_xyz, n, d_out);
    else
        sample_cmem_kernel<true><<<blocks_for(n, 256), 256, 0, stream>>>(params, d_xyz, n, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace birp
