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
    // the ray table never decreases (or never increases) from one step to the next: the cube kernel may then
    // locate the end of a ray's hot run by bisection instead of testing every step (march_steps_span)
    bool monotone_table = false;
    // angular size of one pixel step in x / y (any common unit) and an optional forced tile
    // width (log2, -1 = choose from the angles); only used to shape the per-warp pixel tile
    float pixel_angle_x = 0.f, pixel_angle_y = 0.f;
    int tile_w_log2 = -1;
    // interleaved row bands: of the bands of row_band rows counted from row_begin, only those with
    // index = row_phase (mod row_nphases) are rendered (row_band = 0 or row_nphases = 1: all rows)
    int row_band = 0, row_phase = 0, row_nphases = 1;
    int tile_swizzle = 0;  // lanes in 2x2 pixel blocks (MarchArgs::swizzle)
    int cta_warps = 0;     // warps per CTA: 2, 4 or 8; 0 = chosen from the frame size (pick_cta_warps)
    // maxdist / nsamples when the ray table is the uniform one of Camera::GenerateRayDistWidth (else 0): the
    // tolerance-mode march may then compute the distance of a step instead of reading it; fast_kstep = 0 forbids that
    double fast_step = 0.0;
    int fast_kstep = 1;
    int fast_coop = 0;  // experiment: the warp-per-ray form of the tolerance-mode march (march_cube_fast_coop_kernel)
    int fast_variant = -1;  // fast_sample<VARIANT> of the tolerance-mode march; -1 = the default (experiments: BIRP_FAST_VARIANT)
    // where frame f lands in the image buffer: slot out_frame_offset + f * out_frame_stride
    long long out_frame_offset = 0, out_frame_stride = 1;
};

cudaError_t launch_render_cube(const CubeView& cube, const birp_frame* d_frames, const RenderGeom& g,
                               const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, int* d_status, cudaStream_t stream);
// d_table: the region table of this parameter set on this device (launch_cmem_table), or nullptr
cudaError_t launch_render_cmem(const birp_cmem_params& params, const birp_frame* d_frames,
                               const RenderGeom& g, const float* d_dist, int flags, float* d_images,
                               unsigned long long* d_valid, cudaStream_t stream, const float4* d_table = nullptr);
cudaError_t launch_cmem_table(const birp_cmem_params& params, float4* d_table, cudaStream_t stream);
cudaError_t launch_cmem_table_selftest(const birp_cmem_params& params, const float4* d_table, long n,
                                       unsigned long long* d_out2, cudaStream_t stream);
cudaError_t launch_accumulate(const float* d_samples, long nrays, int nsamples, float ray_width, int flags,
                              float* d_image, cudaStream_t stream);
cudaError_t launch_sample_cube(const CubeView& cube, const float* d_xyz, long n, int flags, float* d_out,
                               int* d_status, cudaStream_t stream);
cudaError_t launch_sample_cmem(const birp_cmem_params& params, const float* d_xyz, long n, int flags,
                               float* d_out, cudaStream_t stream, const float4* d_table = nullptr);

// per-frame setup / orbit poses on the device (k = {a, e, sqrt(1-e^2), u[3], w[3]})
cudaError_t launch_frame_setup(const float* d_poses, int n, float aim_x, float tan_half_x, float tan_half_y,
                               float aspect, birp_frame* d_frames, int* d_status, cudaStream_t stream);
cudaError_t launch_orbit_poses(int n, const double k[9], float* d_poses, cudaStream_t stream);
cudaError_t launch_cmem_math_selftest(long n, unsigned long long* d_err6, cudaStream_t stream);
cudaError_t launch_div10_selftest(unsigned long long* d_bad2, cudaStream_t stream);
cudaError_t launch_cmem_filter_selftest(const birp_cmem_params& params, long n, unsigned long long* d_out3,
                                        cudaStream_t stream);
cudaError_t launch_ray_directions(const birp_frame* d_frame, int W, int H, float* d_dirs, cudaStream_t stream);

// upload-time analysis of a resident cube (see CubeView)
cudaError_t launch_scan_positive(const float* d_v, size_t n, int* d_flag, cudaStream_t stream);
// build the quad copy of the grid and compare every entry with the linear copy (*d_bad |= 1 on a mismatch)
cudaError_t launch_pack_quads(const CubeView& cube, float4* d_quad, int* d_bad, cudaStream_t stream);
// the same for the oct copy (32-byte records: both z layers of a cell)
// (mark: CubeView::oct_marked, for cubes outside the positive class)
cudaError_t launch_pack_octs(const CubeView& cube, float* d_oct, int* d_bad, int mark, cudaStream_t stream);
// the poly copy for the tolerance-mode march (32-byte records: the cell's interpolant as polynomial coefficients)
cudaError_t launch_pack_poly(const CubeView& cube, float* d_poly, cudaStream_t stream);
cudaError_t launch_verify_gather(const CubeView& cube, int ncells, int* d_bad, cudaStream_t stream);
cudaError_t launch_verify_div(float s, float r, unsigned limit_bits, int* d_bad2, cudaStream_t stream);

// completion flags between the GPUs of a box (see signal_kernel / wait_kernel)
cudaError_t launch_peer_signal(unsigned* d_flag, unsigned value, cudaStream_t stream);
cudaError_t launch_peer_wait(const unsigned* d_flags, int n, unsigned value, int* d_status, cudaStream_t stream);

unsigned long long launch_count();

}  // namespace birp
