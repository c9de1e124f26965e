// Camera on top of the C ABI (include/birp_b200.h).
//
// What stays on the host: the once-per-frame orientation (birp_frame_setup, glibc, bit-identical
// to reference src/Camera.cpp:85-137,309-508) and file export.  What moves to the GPU: the whole
// pixel x sample loop of reference src/Camera.cpp:140-232 -- ray generation, march, sampling,
// scaling, termination, accumulation -- for DataCube and CMEM spaces.  A user-defined Space can
// only be sampled through its virtual GetSample on the host; its rays still come from the GPU ray
// generator and its samples are scaled / terminated / accumulated by the GPU (birp_accumulate).
#include "Camera.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "../../include/birp_b200.h"
#include "Cmem.h"

namespace {

[[noreturn]] void raise(int rc, const char* where) {
    const std::string msg = std::string(where) + ": " + birp_last_error();
    switch (rc) {
        case BIRP_ERR_SINGULAR: throw std::runtime_error("Determinant is 0");
        case BIRP_ERR_RANGE: throw std::out_of_range("vector::_M_range_check (NaN sample coordinate)");
        case BIRP_ERR_ARG: throw std::invalid_argument(msg);
        default: throw std::runtime_error(msg);
    }
}

// one 80-column FITS header card
void card(std::string& hdr, const std::string& key, const std::string& value, const std::string& comment) {
    char buf[96];
    std::string v = value;
    if (!comment.empty()) v += " / " + comment;
    std::snprintf(buf, sizeof(buf), "%-8.8s= %-70.70s", key.c_str(), v.c_str());
    hdr.append(buf, 80);
}

// A float keyword value the way cfitsio writes a TFLOAT key (putkey.c: ffpkye with decim = -7 -> ffr2e): "%.7G" of
// the value; an exponent form without a decimal point is redone as "%.1E"; a plain integer-looking result gets a
// trailing '.' so that it reads back as a real; right-justified in columns 11-30.
std::string fits_number(float x) {
    char v[40];
    std::snprintf(v, sizeof(v), "%.7G", (double)x);
    if (!std::strchr(v, '.') && std::strchr(v, 'E')) std::snprintf(v, sizeof(v), "%.1E", (double)x);
    if (!std::strchr(v, '.') && !std::strchr(v, 'E') && !std::strchr(v, 'N') && !std::strchr(v, 'I'))
        std::strncat(v, ".", sizeof(v) - std::strlen(v) - 1);
    char buf[48];
    std::snprintf(buf, sizeof(buf), "%20s", v);
    return buf;
}

std::string fits_int(long x) {
    char buf[40];
    std::snprintf(buf, sizeof(buf), "%20ld", x);
    return buf;
}

// COMMENT card: keyword, two blanks, text from column 11 (cfitsio ffpcom)
void comment_card(std::string& hdr, const std::string& text) {
    char buf[96];
    std::snprintf(buf, sizeof(buf), "COMMENT   %-70.70s", text.c_str());
    hdr.append(buf, 80);
}

void pad_to_block(std::string& s, char fill) { s.append((2880 - s.size() % 2880) % 2880, fill); }

}  // namespace

Camera::Camera(Space& cube, float pixel_size_degrees, float plot_fov_h, float plot_fov_w)
    : dataCube(&cube),
      fov_x(plot_fov_w),
      fov_y(plot_fov_h),
      pixel_size_deg(pixel_size_degrees),
      ray_samples(200),
      ray_max_distance(44) {
    image_dimension_y = (int)std::round(plot_fov_h / pixel_size_deg);
    image_dimension_x = (int)std::round(plot_fov_w / pixel_size_deg);
    if (image_dimension_x < 0) image_dimension_x = 0;
    if (image_dimension_y < 0) image_dimension_y = 0;
    image.assign(image_dimension_y, std::vector<float>(image_dimension_x, 0.f));
}

Camera::~Camera() {}

void Camera::SetPosition(float& x, float& y, float& z) {
    position.x = x;
    position.y = y;
    position.z = z;
}

void Camera::SetAim(float x, float y, float z) {
    aim.x = x;
    aim.y = y;
    aim.z = z;
}

void Camera::SetMarch(int samples, int max_distance) {
    if (samples <= 0) throw std::invalid_argument("Camera::SetMarch: samples must be positive");
    ray_samples = samples;
    ray_max_distance = max_distance;
}

void Camera::Render() {
    const int W = image_dimension_x, H = image_dimension_y;
    if (W <= 0 || H <= 0) return;

    birp_frame frame;
    const float pos[3] = {position.x, position.y, position.z};
    const float aim3[3] = {aim.x, aim.y, aim.z};
    int rc = birp_frame_setup_aim(pos, aim3, fov_x, fov_y, W, H, &frame);
    if (rc != BIRP_OK) raise(rc, "Camera::Render");

    // the reference adds into `image` without clearing it: a second Render() accumulates
    std::vector<float> flat((size_t)W * H);
    bool nonzero = false;
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            const float v = image.at(j).at(i);
            flat[(size_t)j * W + i] = v;
            nonzero |= (v != 0.f) || std::signbit(v);
        }
    const int acc = nonzero ? BIRP_ACCUMULATE : 0;

    if (DataCube* cube = dynamic_cast<DataCube*>(dataCube)) {
        const int flags = (cube->IsTrilinear() ? BIRP_TRILINEAR : BIRP_NEAREST) | (cube->FastMarch() ? BIRP_FAST_MARCH : 0) | acc;
        rc = birp_render_cube(cube->Resident(), &frame, 1, W, H, 0, H, ray_samples, ray_max_distance, flags,
                              flat.data());
        if (rc != BIRP_OK) raise(rc, "Camera::Render(DataCube)");
    } else if (CMEM* cmem = dynamic_cast<CMEM*>(dataCube)) {
        const int flags = (cmem->FastMath() ? BIRP_CMEM_FAST : 0) | acc;
        rc = birp_render_cmem(&cmem->Params(), &frame, 1, W, H, 0, H, ray_samples, ray_max_distance, flags,
                              flat.data());
        if (rc != BIRP_OK) raise(rc, "Camera::Render(CMEM)");
    } else {
        // Any other Space: rays from the GPU ray generator, samples through the virtual call in the
        // reference's order (row by row, near to far, stopping a ray at its first negative scaled
        // sample so that GetSample sees exactly the calls the reference would make).
        std::vector<float> dirs((size_t)W * H * 3);
        rc = birp_ray_directions(&frame, W, H, dirs.data());
        if (rc != BIRP_OK) raise(rc, "Camera::Render(ray generation)");
        std::vector<float> dist(ray_samples);
        float ray_width = 0.f;
        rc = birp_ray_table(ray_samples, ray_max_distance, dist.data(), &ray_width);
        if (rc != BIRP_OK) raise(rc, "Camera::Render(ray table)");
        std::vector<float> samples((size_t)W * H * ray_samples, 0.f);
        for (size_t r = 0; r < (size_t)W * H; ++r) {
            const float* d = &dirs[3 * r];
            float* row = &samples[r * ray_samples];
            for (int k = 0; k < ray_samples; ++k) {
                const float sx = d[0] * dist[k], sy = d[1] * dist[k], sz = d[2] * dist[k];
                const float x = sx + position.x, y = sy + position.y, z = sz + position.z;
                const float s = dataCube->GetSample(x, y, z);
                row[k] = s;
                const float scaled = (s * ray_width * (1000 * 100 * 6378.1) / (4 / M_PI)) / 10000;
                if (scaled < 0) break;
            }
        }
        rc = birp_accumulate(samples.data(), (long)W * H, ray_samples, ray_width, acc, flat.data());
        if (rc != BIRP_OK) raise(rc, "Camera::Render(accumulate)");
    }

    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) image[j][i] = flat[(size_t)j * W + i];
}

int Camera::ToDat(std::string filename) {
    std::ofstream out(filename + ".dat");
    for (int j = 0; j < image_dimension_y; ++j) {
        for (int i = 0; i < image_dimension_x; ++i) out << std::to_string(image[j][i]) << ",";
        out << "\n";
    }
    return 1;
}

// Single-image FITS (the reference goes through EleFits/cfitsio, src/Camera.cpp:252-307): primary
// HDU, BITPIX -32, NAXIS1 = width (x fastest), NAXIS2 = height, the same keywords in the same
// order, big-endian IEEE floats, 2880-byte blocks.  An existing file is overwritten.
int Camera::ToFITS(std::string filename) {
    const int W = image_dimension_x, H = image_dimension_y;
    std::string hdr;
    // the cards, in the order and with the comments cfitsio / EleFits give a single-image file written by
    // src/Camera.cpp:252-307: the mandatory keywords of fits_create_img and its two COMMENT cards, the ten float
    // records of header().writeSeq (unit in brackets: EleFits' raw comment), then the string record "rho"
    card(hdr, "SIMPLE", "                   T", "file does conform to FITS standard");
    card(hdr, "BITPIX", fits_int(-32), "number of bits per data pixel");
    card(hdr, "NAXIS", fits_int(2), "number of data axes");
    card(hdr, "NAXIS1", fits_int(W), "length of data axis 1");
    card(hdr, "NAXIS2", fits_int(H), "length of data axis 2");
    card(hdr, "EXTEND", "                   T", "FITS dataset may contain extensions");
    comment_card(hdr, "FITS (Flexible Image Transport System) format is defined in 'Astronomy");
    comment_card(hdr, "and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H");
    card(hdr, "CRVAL1", fits_number(-(fov_x / 2)), "[deg]");
    card(hdr, "CDELT1", fits_number(pixel_size_deg), "[deg]");
    card(hdr, "CRVAL2", fits_number(-(fov_y / 2)), "[deg]");
    card(hdr, "CDELT2", fits_number(pixel_size_deg), "[deg]");
    card(hdr, "POS_X", fits_number(position.x), "[float]");
    card(hdr, "POS_Y", fits_number(position.y), "[float]");
    card(hdr, "POS_Z", fits_number(position.z), "[float]");
    card(hdr, "AIM_X", fits_number(aim.x), "[float]");
    card(hdr, "AIM_Y", fits_number(aim.y), "[float]");
    card(hdr, "AIM_Z", fits_number(aim.z), "[float]");
    card(hdr, "RHO", "'0       '          ", "[t] comment");  // (a short string is padded to column 30, like a number)
    hdr.append("END");
    pad_to_block(hdr, ' ');

    std::string data((size_t)W * H * 4, '\0');
    size_t n = 0;
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            uint32_t bits;
            std::memcpy(&bits, &image[j][i], 4);
            data[n++] = (char)(bits >> 24);
            data[n++] = (char)(bits >> 16);
            data[n++] = (char)(bits >> 8);
            data[n++] = (char)bits;
        }
    pad_to_block(data, '\0');

    std::ofstream out(filename + ".fits", std::ios::binary | std::ios::trunc);
    if (!out) return 1;  // the reference's callers ignore the return value
    out.write(hdr.data(), (std::streamsize)hdr.size());
    out.write(data.data(), (std::streamsize)data.size());
    return 1;
}
