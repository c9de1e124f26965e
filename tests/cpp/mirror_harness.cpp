// Test harness: drives birp_cpp_b200's C++ classes (the reference's public API) through a C
// interface so that pytest can run the same scenario through the unmodified reference
// (oracle/ref_harness.cpp) and through the product and compare.  Public members only.
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "Batch.h"
#include "Camera.h"
#include "Cmem.h"
#include "DataCube.h"
#include "Helper.h"

namespace {
thread_local std::string g_err;
thread_local int g_err_kind = 0;  // 1 invalid_argument, 2 out_of_range, 3 runtime_error, 4 other
int fail(const std::exception& e) {
    g_err = e.what();
    if (dynamic_cast<const std::invalid_argument*>(&e)) g_err_kind = 1;
    else if (dynamic_cast<const std::out_of_range*>(&e)) g_err_kind = 2;
    else if (dynamic_cast<const std::runtime_error*>(&e)) g_err_kind = 3;
    else g_err_kind = 4;
    return g_err_kind;
}

// the reference tests' fake (tests/Camera_test.cpp:25-29) plus a call counter and a variant
// that turns negative beyond a radius, to exercise the termination rule through a user Space
class CubeFake : public Space {
public:
    float value = 1.f;
    float negative_beyond = -1.f;  // < 0: never negative
    long calls = 0;
    void Init() override {}
    float GetSample(float x, float y, float z) override {
        ++calls;
        if (negative_beyond >= 0 && x * x + y * y + z * z > negative_beyond * negative_beyond) return -value;
        return value + 0.001f * x;
    }
};
}  // namespace

extern "C" {

const char* mir_last_error() { return g_err.c_str(); }
int mir_last_error_kind() { return g_err_kind; }

void* mir_cube_create(const float* zyx, int nx, int ny, int nz, const float* origin, const float* spacing,
                      int trilinear) {
    DataCube* c = new DataCube();
    c->coords_x.resize(nx);
    c->coords_y.resize(ny);
    c->coords_z.resize(nz);
    for (int i = 0; i < nx; ++i) c->coords_x[i] = origin[0] + spacing[0] * i;
    for (int i = 0; i < ny; ++i) c->coords_y[i] = origin[1] + spacing[1] * i;
    for (int i = 0; i < nz; ++i) c->coords_z[i] = origin[2] + spacing[2] * i;
    c->coords_x[0] = origin[0];
    c->coords_y[0] = origin[1];
    c->coords_z[0] = origin[2];
    c->size.x = nx;
    c->size.y = ny;
    c->size.z = nz;
    for (int a = 0; a < 3; ++a) {
        c->spacing[a] = spacing[a];
        c->origin_coords[a] = origin[a];
    }
    c->slices.assign(nz, std::vector<std::vector<float>>(ny));
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j) {
            const float* row = zyx + ((size_t)k * ny + j) * nx;
            c->slices[k][j].assign(row, row + nx);
        }
    c->SetTrilinear(trilinear != 0);
    return c;
}

void* mir_cube_load(const char* path, int trilinear) {
    try {
        DataCube* c = new DataCube();
        std::string p(path);
        c->Init(p, false);
        c->SetTrilinear(trilinear != 0);
        return c;
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

void mir_cube_set_trilinear(void* cube, int t) { static_cast<DataCube*>(cube)->SetTrilinear(t != 0); }

void mir_cube_info(void* cube, int* dims, float* geom) {
    DataCube* c = static_cast<DataCube*>(cube);
    dims[0] = c->size.x;
    dims[1] = c->size.y;
    dims[2] = c->size.z;
    dims[3] = (int)c->slices.size();
    geom[0] = c->coords_x[0];
    geom[1] = c->coords_y[0];
    geom[2] = c->coords_z[0];
    geom[3] = c->spacing[0];
    geom[4] = c->spacing[1];
    geom[5] = c->spacing[2];
}

int mir_cube_export(void* cube, float* zyx) {
    DataCube* c = static_cast<DataCube*>(cube);
    try {
        for (int k = 0; k < c->size.z; ++k)
            for (int j = 0; j < c->size.y; ++j)
                for (int i = 0; i < c->size.x; ++i)
                    zyx[((size_t)k * c->size.y + j) * c->size.x + i] = c->slices.at(k).at(j).at(i);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

void mir_cube_free(void* cube) { delete static_cast<DataCube*>(cube); }

// state[0] = the nested slices have not been built since Load, state[1] = Load was served by a side-car
void mir_cube_state(void* cube, int* state) {
    DataCube* c = static_cast<DataCube*>(cube);
    state[0] = c->SlicesAreLazy() ? 1 : 0;
    state[1] = c->LoadedFromCache() ? 1 : 0;
}

// slices[k][j][i] = v through the public member, like reference-side code that edits the grid by hand
int mir_cube_poke(void* cube, int k, int j, int i, float v) {
    try {
        DataCube* c = static_cast<DataCube*>(cube);
        c->slices.at(k).at(j).at(i) = v;
        c->Invalidate();
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

void mir_cache_policy(int writes, const char* dir) {
    DataCube::SetCacheWrites(writes != 0);
    DataCube::SetCacheDirectory(dir ? dir : "");
}

// DataCube::GetSamples (batched GetSample)
int mir_cube_samples(void* cube, const float* xyz, long n, float* out) {
    try {
        std::vector<float> pts(xyz, xyz + 3 * n);
        const std::vector<float> got = static_cast<DataCube*>(cube)->GetSamples(pts);
        std::copy(got.begin(), got.end(), out);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

void* mir_cmem_create_default() {
    CMEM* m = new CMEM();
    m->Init();
    return m;
}

void* mir_cmem_create(int has_v, const int* v, int has_b, const int* b, float dipole, float p0, int p1, int p2,
                      int p3, float B, float alpha, float beta, float bs, float A1, float A2, float ay_bs,
                      float az_bs, float density) {
    CMEM* m = new CMEM();
    std::vector<int> vv, bb;
    if (has_v) vv.assign(v, v + 3);
    if (has_b) bb.assign(b, b + 3);
    m->Init(true, vv, bb, dipole, p0, p1, p2, p3, B, alpha, beta, bs, A1, A2, ay_bs, az_bs, density);
    return m;
}

void mir_cmem_free(void* m) { delete static_cast<CMEM*>(m); }

void* mir_fake_create(float value, float negative_beyond) {
    CubeFake* f = new CubeFake();
    f->value = value;
    f->negative_beyond = negative_beyond;
    return f;
}
long mir_fake_calls(void* f) { return static_cast<CubeFake*>(f)->calls; }
void mir_fake_free(void* f) { delete static_cast<CubeFake*>(f); }

// kind: 0 DataCube, 1 CMEM, 2 CubeFake
static Space* as_space(void* p, int kind) {
    if (kind == 0) return static_cast<DataCube*>(p);
    if (kind == 1) return static_cast<CMEM*>(p);
    return static_cast<CubeFake*>(p);
}

int mir_space_sample(void* p, int kind, const float* xyz, long n, float* out) {
    try {
        Space* s = as_space(p, kind);
        for (long i = 0; i < n; ++i) out[i] = s->GetSample(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// renders `times` times into one Camera (the reference accumulates); nsamples/maxdist <= 0 keep
// the reference's 200 / 44.  Optionally exports with ToDat / ToFITS.
int mir_render3(void* p, int kind, float px_deg, float fov_h, float fov_w, const float* pos, const float* aim,
                int nsamples, int maxdist, int times, float* image, int* out_wh, const char* dat_path,
                const char* fits_path);
int mir_render(void* p, int kind, float px_deg, float fov_h, float fov_w, const float* pos, float aim_x,
               int nsamples, int maxdist, int times, float* image, int* out_wh, const char* dat_path,
               const char* fits_path) {
    const float aim[3] = {aim_x, 0.f, 0.f};
    return mir_render3(p, kind, px_deg, fov_h, fov_w, pos, aim, nsamples, maxdist, times, image, out_wh, dat_path, fits_path);
}

// the same with the whole aim vector (Camera::SetAim(x, y, z))
int mir_render3(void* p, int kind, float px_deg, float fov_h, float fov_w, const float* pos, const float* aim,
                int nsamples, int maxdist, int times, float* image, int* out_wh, const char* dat_path,
                const char* fits_path) {
    try {
        Space* s = as_space(p, kind);
        Camera cam(*s, px_deg, fov_h, fov_w);
        const int w = cam.Width(), h = cam.Height();
        if (out_wh) {
            out_wh[0] = w;
            out_wh[1] = h;
        }
        if (!image) return 0;
        float x = pos[0], y = pos[1], z = pos[2];
        cam.SetPosition(x, y, z);
        cam.SetAim(aim[0], aim[1], aim[2]);
        if (nsamples > 0) cam.SetMarch(nsamples, maxdist);
        for (int t = 0; t < times; ++t) cam.Render();
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i) image[(size_t)j * w + i] = cam.image[j][i];
        if (dat_path && *dat_path) cam.ToDat(dat_path);
        if (fits_path && *fits_path) cam.ToFITS(fits_path);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// Camera::ToFITS of a given image, no render (and no GPU): a Camera on a dummy Space whose image is filled by hand
int mir_fits_of_image(float px_deg, float fov_h, float fov_w, const float* pos, const float* aim, const float* image,
                      int* out_wh, const char* path) {
    try {
        struct Dummy : Space {
            void Init() override {}
            float GetSample(float, float, float) override { return 0.f; }
        } dummy;
        Camera cam(dummy, px_deg, fov_h, fov_w);
        out_wh[0] = cam.Width();
        out_wh[1] = cam.Height();
        if (!image) return 0;
        float x = pos[0], y = pos[1], z = pos[2];
        cam.SetPosition(x, y, z);
        cam.SetAim(aim[0], aim[1], aim[2]);
        for (int j = 0; j < cam.Height(); ++j)
            for (int i = 0; i < cam.Width(); ++i) cam.image[j][i] = image[(size_t)j * cam.Width() + i];
        cam.ToFITS(path);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// RenderBatchDirectory; returns the number of cubes rendered, -1 on exception
int mir_render_batch(const char* in_dir, const char* out_dir, float px_deg, float fov_h, float fov_w,
                     const float* pos, float aim_x, int trilinear, int threads, int* failed) {
    try {
        auto res = RenderBatchDirectory(in_dir, out_dir, px_deg, fov_h, fov_w, pos[0], pos[1], pos[2], aim_x,
                                        trilinear != 0, threads);
        int ok = 0, bad = 0;
        for (auto& r : res) (r.error.empty() ? ok : bad)++;
        if (failed) *failed = bad;
        return ok;
    } catch (const std::exception& e) {
        fail(e);
        return -1;
    }
}

// RenderSweep: n positions -> n images (kind: 0 DataCube, 1 CMEM); returns 0 / error kind
int mir_render_sweep(void* p, int kind, const float* positions, int n, float aim_x, float px_deg, float fov_h,
                     float fov_w, float* images, int* out_wh) {
    try {
        Space* s = as_space(p, kind);
        std::vector<float> pos(positions, positions + 3 * (size_t)n);
        int w = 0, h = 0;
        std::vector<float> img = RenderSweep(*s, pos, aim_x, px_deg, fov_h, fov_w, &w, &h);
        if (out_wh) {
            out_wh[0] = w;
            out_wh[1] = h;
        }
        if (images) std::copy(img.begin(), img.end(), images);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// Helper's public matrix routines (host only)
int mir_helper_inverse3(const float* m9, float* out9) {
    try {
        std::vector<std::vector<float>> m(3, std::vector<float>(3));
        for (int i = 0; i < 9; ++i) m[i / 3][i % 3] = m9[i];
        auto inv = Helper::GetInverse(m);
        for (int i = 0; i < 9; ++i) out9[i] = inv[i / 3][i % 3];
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

}  // extern "C"
