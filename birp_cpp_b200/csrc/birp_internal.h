// Internal C++ interface between the C ABI (birp_capi.cu) and the kernels (birp_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "birp_device.cuh"

namespace birp {

struct RenderGeom {
    int nframes;
    int W, H;
    int row_begin, row_end;  // rows [row_begin,row_end) of every frame are rendered
    int nsamples;
    float ray_width;
};

// bit in the device status word
constexpr int kStatusNanIndex = 1;

cudaError_t launch_render_cube(const CubeView& cube, const birp_frame* d_frames, const RenderGeom& g,
                               const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, int* d_status, cudaStream_t stream);
cudaError_t launch_render_cmem(const birp_cmem_params& params, const birp_frame* d_frames,
                               const RenderGeom& g, const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, cudaStream_t stream);
cudaError_t launch_accumulate(const float* d_samples, long nrays, int nsamples, float ray_width, int flags,
                              float* d_image, cudaStream_t stream);
cudaError_t launch_sample_cube(const CubeView& cube, const float* d_xyz, long n, int flags, float* d_out,
                               int* d_status, cudaStream_t stream);
cudaError_t launch_sample_cmem(const birp_cmem_params& params, const float* d_xyz, long n, int flags,
                               float* d_out, cudaStream_t stream);

unsigned long long launch_count();

}  // namespace birp
