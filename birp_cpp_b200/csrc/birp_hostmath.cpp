// Host-side, once-per-frame / once-per-run setup of the integrator.
//
// The reference does this work outside its pixel loop, with glibc's fp32 libm; it stays on
// the host here, in the reference's operation order, so that the 13 floats a frame hands
// to the kernel (and the CMEM constants) are bit-identical to the reference's.
// Build this file WITHOUT FMA contraction (-ffp-contract=off, no -march): the reference
// binary contains no fused multiply-adds.
//
//   birp_frame_setup  <- Camera::Integrate prologue   reference src/Camera.cpp:85-137
//                        Camera::Orient               reference src/Camera.cpp:309-417
//                        Camera::PointToPlane         reference src/Camera.cpp:419-508
//                        Helper::MatrixMultiply/GetInverse/GetCofactor/GetDeterminant
//                                                     reference src/Helper.cpp:70-98,150-264
//   birp_ray_table    <- Camera::GenerateRayDistWidth reference src/Camera.cpp:69-81
//   birp_cmem_init    <- CMEM::Init and helpers       reference src/Cmem.cpp:38-115,250-325
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/birp_b200.h"
#include "birp_frame_math.h"
#include "birp_host.h"

namespace {

// the functions the reference itself calls (fp32 overloads of <cmath> -> glibc sinf, cosf, ...)
struct HostMath {
    static float sqrt(float x) { return std::sqrt(x); }
    static float sin(float x) { return std::sin(x); }
    static float cos(float x) { return std::cos(x); }
    static float acos(float x) { return std::acos(x); }
    static float asin(float x) { return std::asin(x); }
    static float atan2(float y, float x) { return std::atan2(y, x); }
    static float abs(float x) { return std::abs(x); }
};

}  // namespace

namespace birp {

void host_fov_terms(float fov_x_deg, float fov_y_deg, int W, int H, float* tan_half_x, float* tan_half_y,
                    float* aspect) {
    const float to_rad = M_PI / 180.f;
    *tan_half_x = std::tan(to_rad * (fov_x_deg / 2));  // src/Camera.cpp:150
    *tan_half_y = std::tan(to_rad * (fov_y_deg / 2));  // src/Camera.cpp:151
    *aspect = (float)W / (float)H;                     // src/Camera.cpp:137
}

int host_frame_setup_aim(const float pos[3], const float aim[3], float fov_x_deg, float fov_y_deg, int W, int H,
                         birp_frame* out) {
    float tx, ty, aspect;
    host_fov_terms(fov_x_deg, fov_y_deg, W, H, &tx, &ty, &aspect);
    return fm::frame_setup_T<HostMath>(pos, aim, tx, ty, aspect, out);
}

int host_frame_setup(const float pos[3], float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                     birp_frame* out) {
    const float aim[3] = {aim_x, 0.f, 0.f};
    return host_frame_setup_aim(pos, aim, fov_x_deg, fov_y_deg, W, H, out);
}

// frame_setup for a whole sweep: poses are independent, so they are split over host threads
// (the per-pose math is ~1 us; a 10 000-pose sweep would otherwise cost a fifth of its render time)
int host_frame_setup_batch(const float* poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                           birp_frame* out, int nthreads) {
    if (n <= 0) return BIRP_OK;
    unsigned hw = std::thread::hardware_concurrency();
    if (nthreads <= 0) {
        // BIRP_HOST_THREADS: processes that share a box (one per GPU) divide the cores between them
        const char* env = std::getenv("BIRP_HOST_THREADS");
        const int want = env ? std::atoi(env) : 0;
        nthreads = want > 0 ? want : (int)(hw ? (hw < 16 ? hw : 16) : 4);
    }
    if (nthreads > (n + 255) / 256) nthreads = (n + 255) / 256;
    std::atomic<int> status{BIRP_OK};
    auto work = [&](int begin, int end) {
        for (int i = begin; i < end; ++i) {
            int rc = host_frame_setup(poses + 3 * (size_t)i, aim_x, fov_x_deg, fov_y_deg, W, H, out + i);
            if (rc != BIRP_OK) status.store(rc);
        }
    };
    if (nthreads <= 1) {
        work(0, n);
    } else {
        std::vector<std::thread> pool;
        const int per = (n + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            const int b = t * per, e = b + per < n ? b + per : n;
            if (b < e) pool.emplace_back(work, b, e);
        }
        for (auto& th : pool) th.join();
    }
    return status.load();
}

// smallest ingress >= 0 that the reference's bounds test `index + 1 >= size` rejects, where
// index = ingress / spacing in fp32 (src/DataCube.cpp:205-218).  The test is monotone in the
// ingress, so a bisection over the fp32 bit patterns finds the exact threshold.
float host_bounds_limit(float spacing, int n) {
    auto rejected = [&](unsigned bits) {
        float xi;
        std::memcpy(&xi, &bits, sizeof(xi));
        const volatile float index = xi / spacing;
        const volatile float plus1 = index + 1;
        return plus1 >= (float)n;
    };
    unsigned lo = 0, hi = 0x7f800000u;  // +0 .. +inf (always rejected)
    if (rejected(lo)) return 0.f;
    while (hi - lo > 1) {
        const unsigned mid = lo + (hi - lo) / 2;
        if (rejected(mid)) hi = mid;
        else lo = mid;
    }
    float out;
    std::memcpy(&out, &hi, sizeof(out));
    return out;
}

// smallest ingress whose quotient by the spacing is >= 1 (monotone, same bisection): below it a
// sample sits in the first cell of the axis, where the interpolation weight can exceed 1 - 2^-23
// and the hot path's clamp-free lerp is not proven (birp_device.cuh).  Never below 2^-100, the
// lower end of the range div_spacing is verified on.
float host_first_cell_limit(float spacing) {
    auto reached = [&](unsigned bits) {
        float xi;
        std::memcpy(&xi, &bits, sizeof(xi));
        const volatile float index = xi / spacing;
        return index >= 1.0f;
    };
    unsigned lo = 0, hi = 0x7f800000u;  // +0 (never: 0 / s = 0) .. +inf (always)
    while (hi - lo > 1) {
        const unsigned mid = lo + (hi - lo) / 2;
        if (reached(mid)) hi = mid;
        else lo = mid;
    }
    float out;
    std::memcpy(&out, &hi, sizeof(out));
    return out < kTinyIngress ? kTinyIngress : out;
}

void host_ray_table(int nsamples, int maxdist, float* dist_out, float* ray_width_out) {
    float d0 = 0.f, d1 = 0.f;
    for (int i = 0; i < nsamples; ++i) {
        const float dist = maxdist * ((i + 0.5f) / nsamples);
        if (dist_out) dist_out[i] = dist;
        if (i == 0) d0 = dist;
        if (i == 1) d1 = dist;
    }
    // the reference reads ray_dist[1] unconditionally; a 1-sample march has no width
    if (ray_width_out) *ray_width_out = (nsamples > 1) ? d1 - d0 : 0.f;
}

void host_cmem_init(int has_v, const int v_in[3], int has_b, const int b_in[3], float dipole, float p0_passed,
                    int p1_passed, int p2_passed, int p3_passed, float B_passed, float alpha, float beta,
                    float bs_passed, float A1_passed, float A2_passed, float ay_bs_passed, float az_bs_passed,
                    float density_passed, birp_cmem_params* out) {
    // solar-wind velocity (km/s) and IMF (nT); defaults of :60-80
    float v[3] = {400, 0, 0};
    float b[3] = {0, 0, 5};
    if (has_v && v_in) { v[0] = v_in[0]; v[1] = v_in[1]; v[2] = v_in[2]; }
    if (has_b && b_in) { b[0] = b_in[0]; b[1] = b_in[1]; b[2] = b_in[2]; }

    birp_cmem_params P;
    std::memset(&P, 0, sizeof(P));
    P.alpha = alpha;
    P.beta = beta;
    // "0 means default" rules of :88-95; p2 deliberately takes p3's value, as the reference does
    P.p0 = p0_passed == 0 ? 0.786300004f : p0_passed;
    P.p1 = p1_passed == 0 ? 1.f : p1_passed;
    P.p2 = p2_passed == 0 ? 3.f : p3_passed;
    P.p3 = p3_passed == 0 ? 4.f : p3_passed;
    P.bs = bs_passed == 0 ? 12.64f : bs_passed;
    P.A1 = A1_passed == 0 ? 7.2e-06f : A1_passed;
    P.A2 = A2_passed == 0 ? 3.5e-06f : A2_passed;
    P.density = density_passed == 0 ? 5 : density_passed;
    (void)B_passed;  // overwritten below, exactly as CalcMagneticPressure overwrites CMEM::B

    // dynamic pressure, nPa (:292-304)
    const float proton_mass = 0.00000000000000000000000000167f;
    const float vms = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * 1000;
    const float n = P.density * 1000000;
    P.pressure_dynamic = 0.5f * proton_mass * n * (vms * vms) * 1000000000;
    // magnetic pressure, nPa (:308-318); B becomes |b| in tesla
    P.B = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) * 0.000000001;
    const float mu0 = 4 * M_PI * 0.0000001;
    P.pressure_magnetic = (P.B * P.B) / (2 * mu0) * 1000000000;
    // bow-shock flaring (:101-110, :320-321)
    const float initial_alpha = (0.58f - 0.010 * b[2]) * (1 + 0.010f * P.pressure_dynamic) + 0.2f;
    P.ay_bs = (ay_bs_passed == 0) ? initial_alpha : ay_bs_passed;
    P.az_bs = (az_bs_passed == 0) ? initial_alpha : az_bs_passed;

    // Lin et al. coefficients (:253-275)
    static const float a[22] = {12.544f, -0.194f, 0.305f, 0.0573f, 2.178f, 0.0571f, -0.999f, 16.473f,
                                0.00152f, 0.382f, 0.0431f, -0.00763f, -0.210f, 0.0405f, -4.430f, -0.636f,
                                -2.600f, 0.832f, -5.328f, 1.103f, -0.907f, 1.450f};
    const float ptot = P.pressure_dynamic + P.pressure_magnetic;
    P.beta_c[0] = a[6] + a[7] * ((std::exp(a[8] * b[2]) - 1) / (std::exp(a[9] * b[2]) + 1));
    P.beta_c[1] = a[10];
    P.beta_c[2] = a[11] + a[12] * dipole;
    P.beta_c[3] = a[13];
    P.charlie = a[14] * std::pow(ptot, a[15]);
    P.dn = a[16] + a[17] * dipole + a[18] * (dipole * dipole);
    P.ds = a[16] - a[17] * dipole + a[18] * (dipole * dipole);
    P.theta_n = a[19] + a[20] * dipole;
    P.theta_s = a[19] - a[20] * dipole;
    P.r0_lin = a[0] * std::pow(ptot, a[1]) * (1 + a[2] * ((std::exp(a[3] * b[2]) - 1) / (std::exp(a[4] * b[2]) + 1)));
    P.a5 = a[5];
    P.a21 = a[21];
    *out = P;
}

}  // namespace birp
