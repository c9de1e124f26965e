// Host-side setup routines (birp_hostmath.cpp), wrapped by the C ABI in birp_capi.cu.
#pragma once
#include "../../include/birp_b200.h"

namespace birp {
// pose-independent frame terms: tanf of the half fields of view, aspect ratio (src/Camera.cpp:137,150-151)
void host_fov_terms(float fov_x_deg, float fov_y_deg, int W, int H, float* tan_half_x, float* tan_half_y,
                    float* aspect);
int host_frame_setup(const float pos[3], float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                     birp_frame* out);
// the same with the whole aim vector (Camera::SetAim(x, y, z): Orient() reads all three components)
int host_frame_setup_aim(const float pos[3], const float aim[3], float fov_x_deg, float fov_y_deg, int W, int H,
                         birp_frame* out);
int host_frame_setup_batch(const float* poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                           birp_frame* out, int nthreads);
void host_ray_table(int nsamples, int maxdist, float* dist_out, float* ray_width_out);
// hot window of the cube kernel for one axis (see CubeView in birp_device.cuh): the smallest ingress the
// reference's bounds test rejects, and the smallest ingress whose index is >= 1 (never below 2^-100)
constexpr float kTinyIngress = 7.888609052210118e-31f;  // 2^-100 (= BIRP_TINY_INGRESS)
float host_bounds_limit(float spacing, int n);
float host_first_cell_limit(float spacing);
void host_cmem_init(int has_v, const int v_in[3], int has_b, const int b_in[3], float dipole, float p0_passed,
                    int p1_passed, int p2_passed, int p3_passed, float B_passed, float alpha, float beta,
                    float bs_passed, float A1_passed, float A2_passed, float ay_bs_passed, float az_bs_passed,
                    float density_passed, birp_cmem_params* out);
}  // namespace birp
