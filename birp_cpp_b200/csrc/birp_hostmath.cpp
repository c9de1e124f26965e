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
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/birp_b200.h"
#include "birp_host.h"

namespace {

struct Vec3 {
    float x, y, z;
};
struct Mat3 {
    float m[3][3];
};

inline float dot3(const Vec3& a, const Vec3& b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline float norm3(const Vec3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// n x n product with a running fp32 sum started at 0 (src/Helper.cpp:80-91)
Mat3 mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k) acc = acc + (a.m[i][k] * b.m[k][j]);
            r.m[i][j] = acc;
        }
    return r;
}

// 2x2 minor obtained by deleting row `skip_r` and column `skip_c`
inline float minor2(const Mat3& a, int skip_r, int skip_c) {
    float s[4];
    int n = 0;
    for (int r = 0; r < 3; ++r) {
        if (r == skip_r) continue;
        for (int c = 0; c < 3; ++c) {
            if (c == skip_c) continue;
            s[n++] = a.m[r][c];
        }
    }
    return s[0] * s[3] - s[1] * s[2];
}

// Laplace expansion along row 0 with an alternating integer sign (src/Helper.cpp:170-193)
float det(const Mat3& a) {
    float result = 0;
    int sign = 1;
    for (int i = 0; i < 3; ++i) {
        result = result + sign * a.m[0][i] * minor2(a, 0, i);
        sign = -sign;
    }
    return result;
}

// transpose(cofactor) scaled by 1/det (src/Helper.cpp:241-263); false when det == 0
bool inverse(const Mat3& a, Mat3* out) {
    if (det(a) == 0) return false;
    const float d = 1.f / det(a);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            // cofactor (j,i): pow(-1, j+i) is evaluated in double in the reference (:235)
            float c = (float)(std::pow(-1, j + i) * minor2(a, j, i));
            c *= d;
            out->m[i][j] = c;
        }
    return true;
}

struct SunProjection {
    float xo, yo;
};

// Camera::PointToPlane for the quantities Orient consumes (xo, yo)
SunProjection project_to_image_plane(const Vec3& obj, float plac, const Vec3& cam, const Vec3& look,
                                     const Vec3& look_unit, const Vec3& north, const Vec3& right) {
    const float to_deg = 180 / M_PI;
    const float snumer = (look.x * obj.x) + (look.y * obj.y) + (look.z * obj.z) + plac;
    const float sdenom = dot3(look, look_unit);
    const float scalar = snumer / sdenom;
    const Vec3 q = {obj.x - (scalar * look_unit.x), obj.y - (scalar * look_unit.y),
                    obj.z - (scalar * look_unit.z)};
    const Vec3 rel = {q.x - cam.x, q.y - cam.y, q.z - cam.z};
    float dp = std::sqrt((rel.x * rel.x) + (rel.y * rel.y) + (rel.z * rel.z));
    const Vec3 rel_unit = {rel.x / dp, rel.y / dp, rel.z / dp};
    dp = std::sqrt((rel_unit.x * rel_unit.x) + (rel_unit.y * rel_unit.y) + (rel_unit.z * rel_unit.z));

    auto clamped_angle_deg = [&](const Vec3& axis, bool clamp_low) {
        const float num = dot3(rel_unit, axis);
        const float den = dp * std::sqrt((axis.x * axis.x) + (axis.y * axis.y) + (axis.z * axis.z));
        float c = num / den;
        if (c > 1) c = 1.f;
        else if (clamp_low && c < -1) c = -1.f;
        return to_deg * std::acos(c);
    };
    const float angle_from_north = clamped_angle_deg(north, true);
    const float angle_from_right = clamped_angle_deg(right, true);
    const float sign_x = (angle_from_right > 90) ? -1.f : 1.f;
    const float sign_y = 1.f;

    Vec3 to_obj = {obj.x - cam.x, obj.y - cam.y, obj.z - cam.z};
    const float to_obj_len = std::sqrt((to_obj.x * to_obj.x) + (to_obj.y * to_obj.y) + (to_obj.z * to_obj.z));
    to_obj = {to_obj.x / to_obj_len, to_obj.y / to_obj_len, to_obj.z / to_obj_len};
    const float t7 = dot3(to_obj, look_unit);
    const float t8 = std::sqrt((to_obj.x * to_obj.x) + (to_obj.y * to_obj.y) + (to_obj.z * to_obj.z));
    float t9 = t7 / t8;
    if (t9 > 1) t9 = 1.f;
    const float off_axis_deg = to_deg * std::acos(t9);

    SunProjection out = {0.f, 0.f};
    if (off_axis_deg > 0.000001f) {
        out.xo = sign_x * off_axis_deg * std::sin(angle_from_north / to_deg);
        out.yo = sign_y * off_axis_deg * std::cos(angle_from_north / to_deg);
    }
    return out;
}

// Camera::Orient: roll angle (degrees, [0,360]) that lays the Sun-Earth line horizontally
float sun_roll_angle(const Vec3& cam, const Vec3& aim) {
    const float to_deg = 180.f / M_PI;
    const Vec3 look = {aim.x - cam.x, aim.y - cam.y, aim.z - cam.z};
    const float len = norm3(look);
    const Vec3 u = {look.x / len, look.y / len, look.z / len};

    const float phi = std::acos(u.z);
    const float s = std::sin(phi);
    // (the reference also derives a theta1 from u.x the same way; nothing reads it)
    float theta2;
    if (std::abs(s) < std::abs(u.y)) {
        // degrees stored where radians are used below: reference behaviour (:352-357)
        theta2 = (s * u.y < 0) ? (float)(std::acos(-1) * to_deg) : (float)(std::acos(1) * to_deg);
    } else {
        theta2 = std::asin(u.y / s);
    }
    Vec3 north = {-1.f * std::cos(theta2) * std::cos(phi), -1.f * std::sin(theta2) * std::cos(phi),
                  std::sin(phi)};
    // octant sign table of :368-392: sign of north.{x,y} as a function of the signs of u
    const bool xn = u.x < 0, yn = u.y < 0, zn = u.z < 0;
    if ((u.x >= 0 || xn) && (u.y >= 0 || yn) && (u.z >= 0 || zn)) {  // false only for NaN components
        const float sx = (xn != zn) ? 1.f : -1.f;
        const float sy = (yn != zn) ? 1.f : -1.f;
        north.x = sx * std::abs(north.x);
        north.y = sy * std::abs(north.y);
    }
    north.z = std::abs(north.z);

    const Vec3 right = {(u.y * north.z) - (u.z * north.y), (u.z * north.x) - (u.x * north.z),
                        (u.x * north.y) - (u.y * north.x)};
    const float plac = (-1.f * look.x * cam.x) + (-1.f * look.y * cam.y) + (-1.f * look.z * cam.z);
    const Vec3 sun = {23455.f, 0.f, 0.f};
    const SunProjection p = project_to_image_plane(sun, plac, cam, look, u, north, right);
    float angle = 270 - (to_deg * std::atan2(0 - p.xo, 0 - p.yo));
    if (angle > 360) angle = angle - 360;
    else if (angle < 0) angle = angle + 360;
    return angle;
}

Mat3 rot_z(float a) {
    return Mat3{{{std::cos(a), -std::sin(a), 0}, {std::sin(a), std::cos(a), 0}, {0, 0, 1}}};
}
Mat3 rot_x(float a) {
    return Mat3{{{1, 0, 0}, {0, std::cos(a), -std::sin(a)}, {0, std::sin(a), std::cos(a)}}};
}

}  // namespace

namespace birp {

int host_frame_setup(const float pos[3], float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                     birp_frame* out) {
    const float to_rad = M_PI / 180.f;
    const Vec3 cam = {pos[0], pos[1], pos[2]};
    const Vec3 aim = {aim_x, 0.f, 0.f};
    const float angle_sun = sun_roll_angle(cam, aim);
    const float angle_sun_corrected = 180.f - angle_sun;

    const float x_diff = aim.x - cam.x;
    const float theta = (2 * M_PI) - std::atan2(x_diff, cam.y);
    const float hypotenuse = std::sqrt(x_diff * x_diff + cam.y * cam.y);
    const float phi = std::atan2(hypotenuse, cam.z);

    const Mat3 rz = rot_z(theta);
    const Mat3 rx = rot_x(phi);
    const Mat3 rz2 = rot_z(angle_sun_corrected * to_rad);
    const Mat3 rotation = mul(mul(rz2, rx), rz);
    Mat3 inv;
    if (!inverse(rotation, &inv)) return BIRP_ERR_SINGULAR;

    out->pos[0] = cam.x;
    out->pos[1] = cam.y;
    out->pos[2] = cam.z;
    std::memcpy(out->rinv, inv.m, sizeof(out->rinv));
    out->tan_half_x = std::tan(to_rad * (fov_x_deg / 2));
    out->tan_half_y = std::tan(to_rad * (fov_y_deg / 2));
    out->aspect = (float)W / (float)H;
    out->sun_angle = angle_sun;
    return BIRP_OK;
}

// frame_setup for a whole sweep: poses are independent, so they are split over host threads
// (the per-pose math is ~1 us; a 10 000-pose sweep would otherwise cost a fifth of its render time)
int host_frame_setup_batch(const float* poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                           birp_frame* out, int nthreads) {
    if (n <= 0) return BIRP_OK;
    unsigned hw = std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = (int)(hw ? (hw < 16 ? hw : 16) : 4);
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
