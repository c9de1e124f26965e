// Per-frame camera setup, written once and instantiated twice:
//
//   * on the host with glibc's fp32 libm (birp_hostmath.cpp) -- the functions the reference itself
//     calls, so the 13 floats of a frame are bit-identical to the reference's;
//   * on the device (frame_setup_kernel, birp_kernels.cu) with every transcendental evaluated in
//     fp64 and rounded once to fp32 -- the same value as glibc's (correctly rounded in the vast
//     majority of cases) result except within ~1e-8 of an fp32 rounding boundary.
//
// Everything between the library calls is plain fp32 / fp64 arithmetic in the reference's
// operation order; both translation units are built without FMA contraction
// (-ffp-contract=off / --fmad=false), like the reference binary.
//
//   frame_setup_T   <- Camera::Integrate prologue   reference src/Camera.cpp:85-137
//   sun_roll_angle  <- Camera::Orient               reference src/Camera.cpp:309-417
//   project_to_image_plane <- Camera::PointToPlane  reference src/Camera.cpp:419-508
//   mul / det / inverse    <- Helper::MatrixMultiply / GetDeterminant / GetCofactor / GetInverse
//                                                   reference src/Helper.cpp:70-98,150-264
//
// The math provider M supplies: sqrt, sin, cos, tan, acos, asin, atan2 (float -> float) and
// abs.
#pragma once
#include "../../include/birp_b200.h"

#ifdef __CUDACC__
#define BIRP_HD __host__ __device__ __forceinline__
#else
#define BIRP_HD inline
#endif

#define BIRP_PI 3.14159265358979323846 /* M_PI */

namespace birp {
namespace fm {

struct Vec3 {
    float x, y, z;
};
struct Mat3 {
    float m[3][3];
};

BIRP_HD float dot3(const Vec3& a, const Vec3& b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }

// n x n product with a running fp32 sum started at 0 (src/Helper.cpp:80-91)
BIRP_HD Mat3 mul(const Mat3& a, const Mat3& b) {
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
BIRP_HD float minor2(const Mat3& a, int skip_r, int skip_c) {
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
BIRP_HD float det(const Mat3& a) {
    float result = 0;
    int sign = 1;
    for (int i = 0; i < 3; ++i) {
        result = result + sign * a.m[0][i] * minor2(a, 0, i);
        sign = -sign;
    }
    return result;
}

// transpose(cofactor) scaled by 1/det (src/Helper.cpp:241-263); false when det == 0
BIRP_HD bool inverse(const Mat3& a, Mat3* out) {
    if (det(a) == 0) return false;
    const float d = 1.f / det(a);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            // cofactor (j,i): the reference multiplies by pow(-1, j+i) in double (:235), i.e. by +-1.0
            float c = (float)((((j + i) & 1) ? -1.0 : 1.0) * minor2(a, j, i));
            c *= d;
            out->m[i][j] = c;
        }
    return true;
}

struct SunProjection {
    float xo, yo;
};

template <class M>
BIRP_HD float clamped_angle_deg(const Vec3& rel_unit, float dp, const Vec3& axis, float to_deg) {
    const float num = dot3(rel_unit, axis);
    const float den = dp * M::sqrt((axis.x * axis.x) + (axis.y * axis.y) + (axis.z * axis.z));
    float c = num / den;
    if (c > 1) c = 1.f;
    else if (c < -1) c = -1.f;
    return to_deg * M::acos(c);
}

// Camera::PointToPlane for the quantities Orient consumes (xo, yo)
template <class M>
BIRP_HD SunProjection project_to_image_plane(const Vec3& obj, float plac, const Vec3& cam, const Vec3& look,
                                             const Vec3& look_unit, const Vec3& north, const Vec3& right) {
    const float to_deg = 180 / BIRP_PI;
    const float snumer = (look.x * obj.x) + (look.y * obj.y) + (look.z * obj.z) + plac;
    const float sdenom = dot3(look, look_unit);
    const float scalar = snumer / sdenom;
    const Vec3 q = {obj.x - (scalar * look_unit.x), obj.y - (scalar * look_unit.y), obj.z - (scalar * look_unit.z)};
    const Vec3 rel = {q.x - cam.x, q.y - cam.y, q.z - cam.z};
    float dp = M::sqrt((rel.x * rel.x) + (rel.y * rel.y) + (rel.z * rel.z));
    const Vec3 rel_unit = {rel.x / dp, rel.y / dp, rel.z / dp};
    dp = M::sqrt((rel_unit.x * rel_unit.x) + (rel_unit.y * rel_unit.y) + (rel_unit.z * rel_unit.z));

    const float angle_from_north = clamped_angle_deg<M>(rel_unit, dp, north, to_deg);
    const float angle_from_right = clamped_angle_deg<M>(rel_unit, dp, right, to_deg);
    const float sign_x = (angle_from_right > 90) ? -1.f : 1.f;
    const float sign_y = 1.f;

    Vec3 to_obj = {obj.x - cam.x, obj.y - cam.y, obj.z - cam.z};
    const float to_obj_len = M::sqrt((to_obj.x * to_obj.x) + (to_obj.y * to_obj.y) + (to_obj.z * to_obj.z));
    to_obj = Vec3{to_obj.x / to_obj_len, to_obj.y / to_obj_len, to_obj.z / to_obj_len};
    const float t7 = dot3(to_obj, look_unit);
    const float t8 = M::sqrt((to_obj.x * to_obj.x) + (to_obj.y * to_obj.y) + (to_obj.z * to_obj.z));
    float t9 = t7 / t8;
    if (t9 > 1) t9 = 1.f;
    const float off_axis_deg = to_deg * M::acos(t9);

    SunProjection out = {0.f, 0.f};
    if (off_axis_deg > 0.000001f) {
        out.xo = sign_x * off_axis_deg * M::sin(angle_from_north / to_deg);
        out.yo = sign_y * off_axis_deg * M::cos(angle_from_north / to_deg);
    }
    return out;
}

// Camera::Orient: roll angle (degrees, [0,360]) that lays the Sun-Earth line horizontally
template <class M>
BIRP_HD float sun_roll_angle(const Vec3& cam, const Vec3& aim) {
    const float to_deg = 180.f / BIRP_PI;
    const Vec3 look = {aim.x - cam.x, aim.y - cam.y, aim.z - cam.z};
    const float len = M::sqrt(look.x * look.x + look.y * look.y + look.z * look.z);
    const Vec3 u = {look.x / len, look.y / len, look.z / len};

    const float phi = M::acos(u.z);
    const float s = M::sin(phi);
    // (the reference also derives a theta1 from u.x the same way; nothing reads it)
    float theta2;
    if (M::abs(s) < M::abs(u.y)) {
        // degrees stored where radians are used below: reference behaviour (:352-357);
        // acos(-1) = pi and acos(1) = 0 in double, times the fp32 to_deg
        theta2 = (s * u.y < 0) ? (float)(BIRP_PI * to_deg) : (float)(0.0 * to_deg);
    } else {
        theta2 = M::asin(u.y / s);
    }
    Vec3 north = {-1.f * M::cos(theta2) * M::cos(phi), -1.f * M::sin(theta2) * M::cos(phi), M::sin(phi)};
    // octant sign table of :368-392: sign of north.{x,y} as a function of the signs of u
    const bool xn = u.x < 0, yn = u.y < 0, zn = u.z < 0;
    if ((u.x >= 0 || xn) && (u.y >= 0 || yn) && (u.z >= 0 || zn)) {  // false only for NaN components
        const float sx = (xn != zn) ? 1.f : -1.f;
        const float sy = (yn != zn) ? 1.f : -1.f;
        north.x = sx * M::abs(north.x);
        north.y = sy * M::abs(north.y);
    }
    north.z = M::abs(north.z);

    const Vec3 right = {(u.y * north.z) - (u.z * north.y), (u.z * north.x) - (u.x * north.z),
                        (u.x * north.y) - (u.y * north.x)};
    const float plac = (-1.f * look.x * cam.x) + (-1.f * look.y * cam.y) + (-1.f * look.z * cam.z);
    const Vec3 sun = {23455.f, 0.f, 0.f};
    const SunProjection p = project_to_image_plane<M>(sun, plac, cam, look, u, north, right);
    float angle = 270 - (to_deg * M::atan2(0 - p.xo, 0 - p.yo));
    if (angle > 360) angle = angle - 360;
    else if (angle < 0) angle = angle + 360;
    return angle;
}

template <class M>
BIRP_HD Mat3 rot_z(float a) {
    return Mat3{{{M::cos(a), -M::sin(a), 0}, {M::sin(a), M::cos(a), 0}, {0, 0, 1}}};
}
template <class M>
BIRP_HD Mat3 rot_x(float a) {
    return Mat3{{{1, 0, 0}, {0, M::cos(a), -M::sin(a)}, {0, M::sin(a), M::cos(a)}}};
}

// tan_half_x / tan_half_y / aspect do not depend on the pose: the caller supplies them
// (host: tanf of the reference, src/Camera.cpp:150-151).  Returns BIRP_OK or BIRP_ERR_SINGULAR.
template <class M>
BIRP_HD int frame_setup_T(const float pos[3], const float aim3[3], float tan_half_x, float tan_half_y, float aspect,
                          birp_frame* out) {
    const float to_rad = BIRP_PI / 180.f;
    const Vec3 cam = {pos[0], pos[1], pos[2]};
    // Orient() reads all three aim components (src/Camera.cpp:323-327); theta and phi below only aim.x (:99-104)
    const Vec3 aim = {aim3[0], aim3[1], aim3[2]};
    const float angle_sun = sun_roll_angle<M>(cam, aim);
    const float angle_sun_corrected = 180.f - angle_sun;

    const float x_diff = aim.x - cam.x;
    const float theta = (2 * BIRP_PI) - M::atan2(x_diff, cam.y);
    const float hypotenuse = M::sqrt(x_diff * x_diff + cam.y * cam.y);
    const float phi = M::atan2(hypotenuse, cam.z);

    const Mat3 rz = rot_z<M>(theta);
    const Mat3 rx = rot_x<M>(phi);
    const Mat3 rz2 = rot_z<M>(angle_sun_corrected * to_rad);
    const Mat3 rotation = mul(mul(rz2, rx), rz);
    Mat3 inv;
    if (!inverse(rotation, &inv)) return BIRP_ERR_SINGULAR;

    out->pos[0] = cam.x;
    out->pos[1] = cam.y;
    out->pos[2] = cam.z;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out->rinv[3 * i + j] = inv.m[i][j];
    out->tan_half_x = tan_half_x;
    out->tan_half_y = tan_half_y;
    out->aspect = aspect;
    out->sun_angle = angle_sun;
    return BIRP_OK;
}

}  // namespace fm
}  // namespace birp
