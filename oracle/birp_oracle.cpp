// TEST INFRASTRUCTURE -- not product code, never linked into libbirp_b200.so.
//
// CPU restatement of the reference's line-of-sight integration on flat arrays,
// written op-for-op so that, built without FMA contraction against the same glibc,
// it is BIT-IDENTICAL to the unmodified reference (oracle/_ref, checked by
// tests/test_oracle_vs_reference.py on every config small enough for the reference).
// It exists because the reference needs minutes for the 1024x1024 / 512^3 /
// orbit-sweep configs; this version is row-parallel and allocation-free.
//
// Parity status: PINNED.  (1) the 12 CMEM known answers of the reference's own
// tests (tests/CMEM_test.cpp:21-44,65-89) reproduce bit-exactly; (2) images, sample
// points and derived constants are compared bit-for-bit with oracle/_ref, i.e. with
// the reference's own objects; (3) committed fixtures under tests/golden/ were made
// from oracle/_ref by tests/golden/make_golden.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
//
// Every function cites the reference lines it follows (paths relative to
// /root/reference).  All arithmetic is fp32 unless a `double` appears here because
// it appears at that point in the reference.
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------
// frame record shared with the tests: 16 floats
//   [0..2] camera position, [3..11] inverse rotation (row-major), [12] tan(fov_x/2),
//   [13] tan(fov_y/2), [14] aspect ratio W/H, [15] sun roll angle (diagnostic)
// ---------------------------------------------------------------------------------
constexpr int kFrameFloats = 16;

// src/Helper.cpp:591-594
inline float vec_len3(float x, float y, float z) { return std::sqrt(x * x + y * y + z * z); }
// src/Helper.cpp:597-600
inline float vec_len2(float x, float y) { return std::sqrt(x * x + y * y); }

// src/Helper.cpp:70-98 (3x3 case)
void mat3_mul(const float a[3][3], const float b[3][3], float out[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float value = 0.0f;
            for (int k = 0; k < 3; ++k) value = value + (a[i][k] * b[k][j]);
            out[i][j] = value;
        }
}

// src/Helper.cpp:150-196 (recursion unrolled for n = 3: 2x2 minors, signed expansion on row 0)
float mat3_det(const float m[3][3]) {
    float result = 0;
    int sign = 1;
    for (int i = 0; i < 3; ++i) {
        float sub[2][2];
        for (int r = 1; r < 3; ++r) {
            int z = 0;
            for (int n = 0; n < 3; ++n)
                if (n != i) sub[r - 1][z++] = m[r][n];
        }
        float det2 = sub[0][0] * sub[1][1] - sub[0][1] * sub[1][0];
        result = result + sign * m[0][i] * det2;
        sign = -sign;
    }
    return result;
}

// src/Helper.cpp:211-264: inverse = transpose(cofactor) * (1/det); the cofactor sign is
// pow(-1, i+j) in double (:235), an exact +-1.  Returns false where the reference throws.
bool mat3_inverse(const float m[3][3], float out[3][3]) {
    if (mat3_det(m) == 0) return false;
    float d = 1.f / mat3_det(m);
    float cof[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float sub[2][2];
            int p = 0;
            for (int x = 0; x < 3; ++x) {
                if (x == i) continue;
                int q = 0;
                for (int y = 0; y < 3; ++y) {
                    if (y == j) continue;
                    sub[p][q++] = m[x][y];
                }
                ++p;
            }
            float det2 = sub[0][0] * sub[1][1] - sub[0][1] * sub[1][0];
            cof[i][j] = (float)(std::pow(-1, i + j) * det2);
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float t = cof[j][i];
            t *= d;
            out[i][j] = t;
        }
    return true;
}

// src/Camera.cpp:419-508.  Only xo, yo are consumed by Orient.
void point_to_plane(const float obj[3], float plac, float prad, const float pos[3],
                    const float dv[3], const float uv[3], const float north[3],
                    const float right[3], float* xo_out, float* yo_out) {
    float to_deg = 180 / M_PI;
    float snumer = (dv[0] * obj[0]) + (dv[1] * obj[1]) + (dv[2] * obj[2]) + plac;
    float sdenom = (dv[0] * uv[0]) + (dv[1] * uv[1]) + (dv[2] * uv[2]);
    float scalar = snumer / sdenom;
    float q[3] = {obj[0] - (scalar * uv[0]), obj[1] - (scalar * uv[1]), obj[2] - (scalar * uv[2])};
    float ov[3] = {q[0] - pos[0], q[1] - pos[1], q[2] - pos[2]};
    float dp = std::sqrt((ov[0] * ov[0]) + (ov[1] * ov[1]) + (ov[2] * ov[2]));
    float ouv[3] = {ov[0] / dp, ov[1] / dp, ov[2] / dp};
    dp = std::sqrt((ouv[0] * ouv[0]) + (ouv[1] * ouv[1]) + (ouv[2] * ouv[2]));

    float t1 = (ouv[0] * north[0]) + (ouv[1] * north[1]) + (ouv[2] * north[2]);
    float t2 = dp * std::sqrt((north[0] * north[0]) + (north[1] * north[1]) + (north[2] * north[2]));
    float t3 = t1 / t2;
    if (t3 > 1) t3 = 1.f;
    else if (t3 < -1) t3 = -1.f;
    float apon = to_deg * std::acos(t3);

    float t4 = (ouv[0] * right[0]) + (ouv[1] * right[1]) + (ouv[2] * right[2]);
    float t5 = dp * std::sqrt((right[0] * right[0]) + (right[1] * right[1]) + (right[2] * right[2]));
    float t6 = t4 / t5;
    if (t6 > 1) t6 = 1.f;
    else if (t6 < -1) t6 = -1.f;
    float apor = to_deg * std::acos(t6);
    float soy = 1.f, sox = 1.f;
    if (apor > 90) sox = -1.f;

    float oov[3] = {obj[0] - pos[0], obj[1] - pos[1], obj[2] - pos[2]};
    float ood = std::sqrt((oov[0] * oov[0]) + (oov[1] * oov[1]) + (oov[2] * oov[2]));
    oov[0] = oov[0] / ood;
    oov[1] = oov[1] / ood;
    oov[2] = oov[2] / ood;
    float aro = to_deg * std::atan(prad / ood);
    (void)aro;
    float t7 = (oov[0] * uv[0]) + (oov[1] * uv[1]) + (oov[2] * uv[2]);
    float t8 = std::sqrt((oov[0] * oov[0]) + (oov[1] * oov[1]) + (oov[2] * oov[2]));
    float t9 = t7 / t8;
    if (t9 > 1) t9 = 1.f;
    float ooa = to_deg * std::acos(t9);
    float xo = 0.f, yo = 0.f;
    if (ooa > 0.000001f) {
        xo = sox * ooa * std::sin(apon / to_deg);
        yo = soy * ooa * std::cos(apon / to_deg);
    }
    *xo_out = xo;
    *yo_out = yo;
}

// src/Camera.cpp:309-417
float orient(const float pos[3], const float aim[3]) {
    float to_deg = 180.f / M_PI;
    float dv[3] = {aim[0] - pos[0], aim[1] - pos[1], aim[2] - pos[2]};
    float distance = vec_len3(dv[0], dv[1], dv[2]);
    float uv[3] = {dv[0] / distance, dv[1] / distance, dv[2] / distance};

    float phi = std::acos(uv[2]);
    float denominator = std::sin(phi);
    float theta1;
    if (std::abs(denominator) < std::abs(uv[0])) {
        if (denominator * uv[0] < 0) theta1 = std::acos(-1) * to_deg;
        else theta1 = std::acos(1) * to_deg;
    } else {
        theta1 = std::asin(uv[0] / denominator);
    }
    (void)theta1;
    float theta2;
    if (std::abs(denominator) < std::abs(uv[1])) {
        if (denominator * uv[1] < 0) theta2 = std::acos(-1) * to_deg;
        else theta2 = std::acos(1) * to_deg;
    } else {
        theta2 = std::asin(uv[1] / denominator);
    }
    float north[3] = {-1.f * std::cos(theta2) * std::cos(phi), -1.f * std::sin(theta2) * std::cos(phi),
                      std::sin(phi)};
    const bool x0 = uv[0] >= 0, y0 = uv[1] >= 0, z0 = uv[2] >= 0;
    const bool x1 = uv[0] < 0, y1 = uv[1] < 0, z1 = uv[2] < 0;
    float ax = std::abs(north[0]), ay = std::abs(north[1]);
    if (x0 && y0 && z0) { north[0] = -1.f * ax; north[1] = -1.f * ay; }
    else if (x0 && y0 && z1) { north[0] = 1.f * ax; north[1] = 1.f * ay; }
    else if (x0 && y1 && z0) { north[0] = -1.f * ax; north[1] = 1.f * ay; }
    else if (x0 && y1 && z1) { north[0] = 1.f * ax; north[1] = -1.f * ay; }
    else if (x1 && y0 && z0) { north[0] = 1.f * ax; north[1] = -1.f * ay; }
    else if (x1 && y0 && z1) { north[0] = -1.f * ax; north[1] = 1.f * ay; }
    else if (x1 && y1 && z0) { north[0] = 1.f * ax; north[1] = 1.f * ay; }
    else if (x1 && y1 && z1) { north[0] = -1.f * ax; north[1] = -1.f * ay; }
    north[2] = std::abs(north[2]);

    float right[3] = {(uv[1] * north[2]) - (uv[2] * north[1]), (uv[2] * north[0]) - (uv[0] * north[2]),
                      (uv[0] * north[1]) - (uv[1] * north[0])};
    float plac = (-1.f * dv[0] * pos[0]) + (-1.f * dv[1] * pos[1]) + (-1.f * dv[2] * pos[2]);
    float sun[3] = {23455.f, 0.f, 0.f};
    float xo, yo;
    point_to_plane(sun, plac, 1, pos, dv, uv, north, right, &xo, &yo);
    float angle_sun = 270 - (to_deg * std::atan2(0 - xo, 0 - yo));
    if (angle_sun > 360) angle_sun = angle_sun - 360;
    else if (angle_sun < 0) angle_sun = angle_sun + 360;
    return angle_sun;
}

// src/Camera.cpp:85-137 plus the per-frame factors of :150-151
int frame_setup3(const float pos[3], const float aim[3], float fov_x, float fov_y, int W, int H, float* fr);
int frame_setup(const float pos[3], float aim_x, float fov_x, float fov_y, int W, int H, float* fr) {
    const float aim[3] = {aim_x, 0.f, 0.f};
    return frame_setup3(pos, aim, fov_x, fov_y, W, H, fr);
}
int frame_setup3(const float pos[3], const float aim[3], float fov_x, float fov_y, int W, int H, float* fr) {
    float to_rad = M_PI / 180.f;
    float angle_sun = orient(pos, aim);
    float angle_sun_corrected = 180.f - angle_sun;
    float x_diff = aim[0] - pos[0];
    float theta = (2 * M_PI) - std::atan2(x_diff, pos[1]);
    float hypotenuse = std::sqrt(x_diff * x_diff + pos[1] * pos[1]);
    float phi = std::atan2(hypotenuse, pos[2]);
    float rz[3][3] = {{std::cos(theta), -std::sin(theta), 0}, {std::sin(theta), std::cos(theta), 0}, {0, 0, 1}};
    float rx[3][3] = {{1, 0, 0}, {0, std::cos(phi), -std::sin(phi)}, {0, std::sin(phi), std::cos(phi)}};
    float a = angle_sun_corrected * to_rad;
    float rz2[3][3] = {{std::cos(a), -std::sin(a), 0}, {std::sin(a), std::cos(a), 0}, {0, 0, 1}};
    float rot[3][3], rot2[3][3], inv[3][3];
    mat3_mul(rz2, rx, rot);
    mat3_mul(rot, rz, rot2);
    if (!mat3_inverse(rot2, inv)) return 2;  // reference throws runtime_error("Determinant is 0")
    fr[0] = pos[0];
    fr[1] = pos[1];
    fr[2] = pos[2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) fr[3 + 3 * i + j] = inv[i][j];
    fr[12] = std::tan(to_rad * (fov_x / 2));
    fr[13] = std::tan(to_rad * (fov_y / 2));
    fr[14] = (float)W / (float)H;
    fr[15] = angle_sun;
    return 0;
}

// src/Camera.cpp:148-168: unit direction in world space of pixel (i, j)
inline void pixel_direction(const float* fr, int i, int j, int W, int H, float d[3]) {
    float device_x = (i + 0.5f) / (W);
    float device_y = (j + 0.5f) / (H);
    float screen_x = ((2 * device_x) - 1) * fr[12] * fr[14];
    float screen_y = (1 - (2 * device_y)) * fr[13];
    float nd = vec_len3(screen_x, screen_y, -1);
    float su[3] = {screen_x / nd, screen_y / nd, -1 / nd};
    const float* r = fr + 3;
    float w[3];
    for (int row = 0; row < 3; ++row)  // src/Helper.cpp:136-148
        w[row] = (r[3 * row] * su[0]) + (r[3 * row + 1] * su[1]) + (r[3 * row + 2] * su[2]);
    float wd = vec_len3(w[0], w[1], w[2]);
    d[0] = w[0] / wd;
    d[1] = w[1] / wd;
    d[2] = w[2] / wd;
}

// src/Camera.cpp:201: fp32 product, then double scale chain, rounded to fp32 on assignment
inline float scale_sample(float sample, float ray_width) {
    sample = (sample * ray_width * (1000 * 100 * 6378.1) / (4 / M_PI)) / 10000;
    return sample;
}

// ---------------------------------------------------------------------------------
// DataCube sampling, src/DataCube.cpp:203-277.  status: 0 ok, 1 = reference would throw
// std::out_of_range from .at() (NaN coordinates make the int conversions garbage).
// ---------------------------------------------------------------------------------
struct Cube {
    const float* v;
    int nx, ny, nz;
    float ox, oy, oz;
    float sx, sy, sz;
    bool trilinear;
};

inline float cube_at(const Cube& c, int z, int y, int x, int* status) {
    if ((unsigned)z >= (unsigned)c.nz || (unsigned)y >= (unsigned)c.ny || (unsigned)x >= (unsigned)c.nx) {
        *status = 1;
        return 0.f;
    }
    return c.v[((size_t)z * c.ny + y) * c.nx + x];
}

inline float cube_sample(const Cube& c, float x, float y, float z, int* status) {
    float x_ingress = x - c.ox;
    float x_index = (x_ingress / c.sx);
    float y_ingress = y - c.oy;
    float y_index = (y_ingress / c.sy);
    float z_ingress = z - c.oz;
    float z_index = (z_ingress / c.sz);
    if ((0 > x_ingress || x_index + 1 >= c.nx) || (0 > y_ingress || y_index + 1 >= c.ny) ||
        (0 > z_ingress || z_index + 1 >= c.nz)) {
        return -1.f;
    }
    if (!(x_index == x_index) || !(y_index == y_index) || !(z_index == z_index)) {
        *status = 1;  // NaN index: the reference's .at() throws
        return 0.f;
    }
    float sample;
    if (c.trilinear) {
        int z_min = std::floor(z_index), z_max = std::ceil(z_index);
        int y_min = std::floor(y_index), y_max = std::ceil(y_index);
        int x_min = std::floor(x_index), x_max = std::ceil(x_index);
        float s000 = cube_at(c, z_min, y_min, x_min, status);
        float s001 = cube_at(c, z_min, y_min, x_max, status);
        float s010 = cube_at(c, z_min, y_max, x_min, status);
        float s011 = cube_at(c, z_min, y_max, x_max, status);
        float s100 = cube_at(c, z_max, y_min, x_min, status);
        float s101 = cube_at(c, z_max, y_min, x_max, status);
        float s110 = cube_at(c, z_max, y_max, x_min, status);
        float s111 = cube_at(c, z_max, y_max, x_max, status);
        float x_frac = x_index - (int)x_index;
        float y_frac = y_index - (int)y_index;
        float z_frac = z_index - (int)z_index;
        float low_x1 = std::lerp(s000, s001, x_frac);
        float low_x2 = std::lerp(s010, s011, x_frac);
        float high_x1 = std::lerp(s100, s101, x_frac);
        float high_x2 = std::lerp(s110, s111, x_frac);
        float low_y = std::lerp(low_x1, low_x2, y_frac);
        float high_y = std::lerp(high_x1, high_x2, y_frac);
        sample = std::lerp(low_y, high_y, z_frac);
    } else {
        sample = cube_at(c, (int)z_index, (int)y_index, (int)x_index, status);
    }
    return sample;
}

// ---------------------------------------------------------------------------------
// CMEM.  Parameter record: 27 floats in the order of birp_cmem_params (include/birp_b200.h)
// ---------------------------------------------------------------------------------
enum {
    P_P0, P_P1, P_P2, P_P3, P_BS, P_A1, P_A2, P_B, P_ALPHA, P_BETA, P_AY_BS, P_AZ_BS, P_R0_LIN,
    P_CHARLIE, P_DN, P_DS, P_THETA_N, P_THETA_S, P_A5, P_A21, P_BC0, P_BC1, P_BC2, P_BC3,
    P_PDYN, P_PMAG, P_DENSITY, P_COUNT
};

// src/Cmem.cpp:38-115, 250-276, 292-325
void cmem_init(int has_v, const int* v_in, int has_b, const int* b_in, float dipole_passed,
               float p0_passed, int p1_passed, int p2_passed, int p3_passed, float B_passed,
               float alpha_passed, float beta_passed, float bs_passed, float A1_passed, float A2_passed,
               float ay_bs_passed, float az_bs_passed, float density_passed, float* P) {
    float v[3], b[3];
    if (!has_v) { v[0] = 400; v[1] = 0; v[2] = 0; }
    else { v[0] = v_in[0]; v[1] = v_in[1]; v[2] = v_in[2]; }
    if (!has_b) { b[0] = 0; b[1] = 0; b[2] = 5; }
    else { b[0] = b_in[0]; b[1] = b_in[1]; b[2] = b_in[2]; }
    float dipole = dipole_passed;
    float B = B_passed;
    float alpha = alpha_passed, beta = beta_passed;
    float p0 = p0_passed == 0 ? 0.786300004f : p0_passed;
    float p1 = p1_passed == 0 ? 1.f : p1_passed;
    float p2 = p2_passed == 0 ? 3.f : p3_passed;  // sic: p2 takes p3's value (:90)
    float p3 = p3_passed == 0 ? 4.f : p3_passed;
    float bs = bs_passed == 0 ? 12.64f : bs_passed;
    float A1 = A1_passed == 0 ? 7.2e-06f : A1_passed;
    float A2 = A2_passed == 0 ? 3.5e-06f : A2_passed;
    float density = density_passed == 0 ? 5 : density_passed;

    // CalcDynamicPressure :292-304
    float proton_mass = 0.00000000000000000000000000167f;
    float vms = vec_len3(v[0], v[1], v[2]) * 1000;
    float n = density * 1000000;
    float pressure_dynamic = 0.5f * proton_mass * n * (vms * vms) * 1000000000;
    // CalcMagneticPressure :308-318 (overwrites B)
    B = vec_len3(b[0], b[1], b[2]) * 0.000000001;
    float mu0 = 4 * M_PI * 0.0000001;
    float pressure_magnetic = (B * B) / (2 * mu0) * 1000000000;
    // CalcInitialAlpha :320-321
    auto initial_alpha = [&]() -> float {
        return (0.58f - 0.010 * b[2]) * (1 + 0.010f * pressure_dynamic) + 0.2f;
    };
    float ay_bs = (ay_bs_passed == 0) ? initial_alpha() : ay_bs_passed;
    float az_bs = (az_bs_passed == 0) ? initial_alpha() : az_bs_passed;

    // DefineLinearCoeffs :250-276
    const float a[22] = {12.544f, -0.194f, 0.305f, 0.0573f, 2.178f, 0.0571f, -0.999f, 16.473f,
                         0.00152f, 0.382f, 0.0431f, -0.00763f, -0.210f, 0.0405f, -4.430f, -0.636f,
                         -2.600f, 0.832f, -5.328f, 1.103f, -0.907f, 1.450f};
    float bc0 = a[6] + a[7] * ((std::exp(a[8] * b[2]) - 1) / (std::exp(a[9] * b[2]) + 1));
    float bc1 = a[10];
    float bc2 = a[11] + a[12] * dipole;
    float bc3 = a[13];
    float charlie = a[14] * std::pow((pressure_dynamic + pressure_magnetic), a[15]);
    float dn = a[16] + a[17] * dipole + a[18] * (dipole * dipole);
    float ds = a[16] - a[17] * dipole + a[18] * (dipole * dipole);
    float theta_n = a[19] + a[20] * dipole;
    float theta_s = a[19] - a[20] * dipole;
    float r0_lin = a[0] * std::pow(pressure_dynamic + pressure_magnetic, a[1]) *
                   (1 + a[2] * ((std::exp(a[3] * b[2]) - 1) / (std::exp(a[4] * b[2]) + 1)));

    P[P_P0] = p0; P[P_P1] = p1; P[P_P2] = p2; P[P_P3] = p3; P[P_BS] = bs; P[P_A1] = A1; P[P_A2] = A2;
    P[P_B] = B; P[P_ALPHA] = alpha; P[P_BETA] = beta; P[P_AY_BS] = ay_bs; P[P_AZ_BS] = az_bs;
    P[P_R0_LIN] = r0_lin; P[P_CHARLIE] = charlie; P[P_DN] = dn; P[P_DS] = ds; P[P_THETA_N] = theta_n;
    P[P_THETA_S] = theta_s; P[P_A5] = a[5]; P[P_A21] = a[21]; P[P_BC0] = bc0; P[P_BC1] = bc1;
    P[P_BC2] = bc2; P[P_BC3] = bc3; P[P_PDYN] = pressure_dynamic; P[P_PMAG] = pressure_magnetic;
    P[P_DENSITY] = density;
}

// src/Cmem.cpp:188-222
inline float lin_scaled(const float* P, float theta, float phi) {
    float theta_n = P[P_THETA_N], theta_s = P[P_THETA_S];
    float phi_n = std::acos((std::cos(theta) * std::cos(theta_n)) +
                            (std::sin(theta) * std::sin(theta_n) * std::cos(phi - M_PI / 2)));
    float phi_s = std::acos((std::cos(theta) * std::cos(theta_s)) +
                            (std::sin(theta) * std::sin(theta_s) * std::cos(phi - (3 * M_PI / 2))));
    float sin_phi = std::sin(phi);
    float f = std::pow(std::cos(theta / 2) + P[P_A5] * std::sin(2 * theta) * (1 - std::exp(-theta)),
                       (P[P_P1] * (P[P_BC0] + P[P_BC1] * std::cos(phi) + P[P_BC3] * (sin_phi * sin_phi))));
    float q = P[P_P2] * P[P_CHARLIE] * std::exp(P[P_P3] * P[P_DN] * (std::pow(phi_n, P[P_A21]))) +
              P[P_P2] * P[P_CHARLIE] * std::exp(P[P_P3] * P[P_DS] * (std::pow(phi_s, P[P_A21])));
    float r = P[P_P0] * P[P_R0_LIN] * f + q;
    return r;
}

// src/Cmem.cpp:280-288
inline float shue_model(float theta, float phi, float r0, float ay, float az) {
    float ry = r0 * std::pow((2 / (1 + std::cos(theta))), ay);
    float rz = r0 * std::pow((2 / (1 + std::cos(theta))), az);
    float rz_x_cos_phi = rz * std::cos(phi);
    float ry_x_sin_phi = ry * std::sin(phi);
    float r = (ry * rz) / std::sqrt((rz_x_cos_phi * rz_x_cos_phi) + (ry_x_sin_phi * ry_x_sin_phi));
    return r;
}

// src/Cmem.cpp:117-186 with ShueCoords :224-246 inlined
inline float cmem_sample(const float* P, float x, float y, float z) {
    float r = vec_len3(x, y, z);
    float theta = (x != 0) ? std::acos(x / r) : 0.f;
    float phi = (y != 0) ? std::acos(y / vec_len2(y, z)) : 0.f;

    float radius_mp = lin_scaled(P, theta, phi);
    float radius_bs = shue_model(theta, phi, P[P_BS], P[P_AY_BS], P[P_AZ_BS]);
    float thickness = 0.05f;
    float half_thickness = thickness / 2;
    const float A1 = P[P_A1], A2 = P[P_A2], B = P[P_B], alpha = P[P_ALPHA], beta = P[P_BETA];

    if (r < radius_mp - half_thickness) {
        return 0;
    } else if (r >= radius_mp - half_thickness && r <= radius_mp + half_thickness) {
        float shue_norm = (r - (radius_mp - half_thickness)) / (radius_mp + half_thickness);
        float b0 = 0;
        float h = theta / 2;
        float s = std::sin(theta);
        float b1 = A1 * (std::exp(-B * (h * h * h * h))) * std::pow(r / 10, (-alpha - (beta * (s * s))));
        return std::lerp(b0, b1, shue_norm);
    } else if (r > radius_mp + half_thickness && r < radius_bs - half_thickness) {
        float h = theta / 2;
        float s = std::sin(theta);
        return A1 * (std::exp(-B * (h * h * h * h))) * std::pow(r / 10, (-alpha - (beta * (s * s))));
    } else if (r >= radius_bs - half_thickness && r <= radius_bs + half_thickness) {
        float h = theta / 2;
        float s = std::sin(theta);
        double shue_div_10 = r / 10;
        float shue_norm = (r - (radius_bs - half_thickness)) / (radius_bs + half_thickness);
        float b1 = A1 * (std::exp(-B * (h * h * h * h))) * std::pow(r / 10, (-alpha - (beta * (s * s))));
        float b2 = A2 * (1 / (shue_div_10 * shue_div_10 * shue_div_10));
        return std::lerp(b1, b2, shue_norm);
    } else {
        double shue_div_10 = r / 10;
        return A2 * (1 / (shue_div_10 * shue_div_10 * shue_div_10));
    }
}

// src/Camera.cpp:69-81
void ray_table(int nsamples, int maxdist, std::vector<float>& dist, float* ray_width) {
    dist.resize(nsamples);
    for (int i = 0; i < nsamples; ++i) dist[i] = maxdist * ((i + 0.5f) / nsamples);
    *ray_width = nsamples > 1 ? dist[1] - dist[0] : 0.f;
}

// src/Camera.cpp:140-232 for rows [row0,row1) of one frame.  SampleFn(x,y,z,&status).
template <class SampleFn>
void integrate_rows(const float* fr, int W, int H, int row0, int row1, const std::vector<float>& dist,
                    float ray_width, float* image, SampleFn&& sample_fn, std::atomic<int>& status,
                    unsigned long long* valid) {
    const int nsamples = (int)dist.size();
    unsigned long long nvalid = 0;
    for (int j = row0; j < row1; ++j) {
        for (int i = 0; i < W; ++i) {
            float d[3];
            pixel_direction(fr, i, j, W, H, d);
            float acc = image[(size_t)j * W + i];
            for (int k = 0; k < nsamples; ++k) {
                float sv[3] = {d[0] * dist[k], d[1] * dist[k], d[2] * dist[k]};  // src/Helper.cpp:127-134
                float xc = sv[0] + fr[0];
                float yc = sv[1] + fr[1];
                float zc = sv[2] + fr[2];
                int st = 0;
                float sample = sample_fn(xc, yc, zc, &st);
                if (st) {
                    status.store(st);
                    return;
                }
                sample = scale_sample(sample, ray_width);
                if (sample < 0) break;  // :207-210
                acc = acc + sample;     // :223
                ++nvalid;
            }
            image[(size_t)j * W + i] = acc;
        }
    }
    if (valid) __atomic_fetch_add(valid, nvalid, __ATOMIC_RELAXED);
}

template <class SampleFn>
int render_frames(const float* frames, int nframes, int W, int H, int row0, int row1, int nsamples,
                  int maxdist, float* images, int nthreads, unsigned long long* valid, SampleFn sample_fn) {
    if (row0 < 0) row0 = 0;
    if (row1 < 0 || row1 > H) row1 = H;
    std::vector<float> dist;
    float ray_width;
    ray_table(nsamples, maxdist, dist, &ray_width);
    if (valid) *valid = 0;
    std::atomic<int> status{0};
    const long total_rows = (long)nframes * (row1 - row0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > total_rows) nthreads = (int)(total_rows > 0 ? total_rows : 1);
    std::atomic<long> next{0};
    auto worker = [&]() {
        for (;;) {
            long t = next.fetch_add(1);
            if (t >= total_rows || status.load()) break;
            int f = (int)(t / (row1 - row0));
            int j = row0 + (int)(t % (row1 - row0));
            integrate_rows(frames + (size_t)f * kFrameFloats, W, H, j, j + 1, dist, ray_width,
                           images + (size_t)f * W * H, sample_fn, status, valid);
        }
    };
    if (nthreads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return status.load();
}

}  // namespace

extern "C" {

int orc_frame_floats() { return kFrameFloats; }
int orc_cmem_param_floats() { return P_COUNT; }

// returns 0 ok, 2 singular rotation (reference: runtime_error)
int orc_frame_setup(const float* pos, float aim_x, float fov_x, float fov_y, int W, int H, float* frame16) {
    return frame_setup(pos, aim_x, fov_x, fov_y, W, H, frame16);
}

int orc_frame_setup3(const float* pos, const float* aim, float fov_x, float fov_y, int W, int H, float* frame16) {
    return frame_setup3(pos, aim, fov_x, fov_y, W, H, frame16);
}

float orc_orient(const float* pos, float aim_x) {
    float aim[3] = {aim_x, 0.f, 0.f};
    return orient(pos, aim);
}

void orc_ray_table(int nsamples, int maxdist, float* dist_out, float* ray_width_out) {
    std::vector<float> d;
    ray_table(nsamples, maxdist, d, ray_width_out);
    std::memcpy(dist_out, d.data(), sizeof(float) * d.size());
}

// world-space sample points of one frame, [H][W][nsamples][3] (pins ray generation)
void orc_trace_points(const float* frame16, int W, int H, int nsamples, int maxdist, float* xyz) {
    std::vector<float> dist;
    float rw;
    ray_table(nsamples, maxdist, dist, &rw);
    size_t n = 0;
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            float d[3];
            pixel_direction(frame16, i, j, W, H, d);
            for (int k = 0; k < nsamples; ++k) {
                xyz[n++] = d[0] * dist[k] + frame16[0];
                xyz[n++] = d[1] * dist[k] + frame16[1];
                xyz[n++] = d[2] * dist[k] + frame16[2];
            }
        }
}

// images[nframes][H][W] are ACCUMULATED INTO (src/Camera.cpp:223); rows outside [row0,row1) untouched.
// returns 0 ok, 1 = the reference would have thrown std::out_of_range (NaN sample coordinates).
int orc_render_cube(const float* zyx, int nx, int ny, int nz, const float* origin, const float* spacing,
                    int trilinear, const float* frames, int nframes, int W, int H, int row0, int row1,
                    int nsamples, int maxdist, float* images, int nthreads, unsigned long long* valid) {
    Cube c{zyx, nx, ny, nz, origin[0], origin[1], origin[2], spacing[0], spacing[1], spacing[2], trilinear != 0};
    return render_frames(frames, nframes, W, H, row0, row1, nsamples, maxdist, images, nthreads, valid,
                         [c](float x, float y, float z, int* st) { return cube_sample(c, x, y, z, st); });
}

int orc_render_cmem(const float* params, const float* frames, int nframes, int W, int H, int row0, int row1,
                    int nsamples, int maxdist, float* images, int nthreads, unsigned long long* valid) {
    return render_frames(frames, nframes, W, H, row0, row1, nsamples, maxdist, images, nthreads, valid,
                         [params](float x, float y, float z, int*) { return cmem_sample(params, x, y, z); });
}

int orc_cube_sample(const float* zyx, int nx, int ny, int nz, const float* origin, const float* spacing,
                    int trilinear, const float* xyz, long n, float* out) {
    Cube c{zyx, nx, ny, nz, origin[0], origin[1], origin[2], spacing[0], spacing[1], spacing[2], trilinear != 0};
    int status = 0;
    for (long i = 0; i < n; ++i) out[i] = cube_sample(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &status);
    return status;
}

void orc_cmem_init(int has_v, const int* v, int has_b, const int* b, float dipole, float p0, int p1, int p2,
                   int p3, float B, float alpha, float beta, float bs, float A1, float A2, float ay_bs,
                   float az_bs, float density, float* params_out) {
    cmem_init(has_v, v, has_b, b, dipole, p0, p1, p2, p3, B, alpha, beta, bs, A1, A2, ay_bs, az_bs, density,
              params_out);
}

void orc_cmem_sample(const float* params, const float* xyz, long n, float* out) {
    for (long i = 0; i < n; ++i) out[i] = cmem_sample(params, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

// src/Camera.cpp:201-223 on a host-sampled [nrays][nsamples] buffer (user-defined Space subclasses)
void orc_accumulate(const float* samples, long nrays, int nsamples, float ray_width, float* image) {
    for (long r = 0; r < nrays; ++r) {
        float acc = image[r];
        for (int k = 0; k < nsamples; ++k) {
            float s = scale_sample(samples[r * nsamples + k], ray_width);
            if (s < 0) break;
            acc = acc + s;
        }
        image[r] = acc;
    }
}

}  // extern "C"
