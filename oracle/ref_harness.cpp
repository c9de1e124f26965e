// TEST INFRASTRUCTURE -- not product code.
//
// C-callable harness around the UNMODIFIED reference classes.  It is compiled
// together with /root/reference/src/{Camera,DataCube,Cmem,Helper}.cpp (where they
// lie, never copied) into oracle/_ref/libbirp_ref.so by oracle/Makefile.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load that library.
//
// What it drives (reference file:line):
//   Camera::Camera               src/Camera.cpp:29-47
//   Camera::SetPosition/SetAim   src/Camera.cpp:49-61
//   Camera::GenerateRayDistWidth src/Camera.cpp:69-81   (private; reached below)
//   Camera::Integrate            src/Camera.cpp:83-233  (private; reached below)
//   DataCube::GetSample          src/DataCube.cpp:203-277
//   CMEM::Init / GetSample       src/Cmem.cpp:18-186
//
// The reference hard-codes 200 samples / 44 RE (src/Camera.cpp:31,65) and derives
// the image size from fov/pixel size.  To run the other BASELINE configs without
// editing the reference, this one translation unit sees the private members.
#include <chrono>
#include <cmath>
#include <cstring>
#include <exception>
#include <iostream>
#include <string>
#include <vector>

#define private public
#include "Camera.h"
#include "Cmem.h"
#include "DataCube.h"
#undef private

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) {
    g_err = e.what();
    return 1;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ---- DataCube built in memory through its public members (include/DataCube.h:30-42)
void* ref_cube_create(const float* zyx, int nx, int ny, int nz, const float* origin,
                      const float* spacing, int trilinear) {
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

// ---- DataCube parsed from an ASCII .dat by the reference loader (src/DataCube.cpp:28-138)
void* ref_cube_load(const char* path, int trilinear) {
    try {
        DataCube* c = new DataCube();
        c->Load(path, false);
        c->SetTrilinear(trilinear != 0);
        return c;
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

void ref_cube_set_trilinear(void* cube, int trilinear) {
    static_cast<DataCube*>(cube)->SetTrilinear(trilinear != 0);
}

// dims = {nx, ny, nz, slices.size()}, geom = {origin xyz, spacing xyz}
void ref_cube_info(void* cube, int* dims, float* geom) {
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

// copy slices[k][j][i] into a flat z-major buffer (for loader parity tests)
int ref_cube_export(void* cube, float* zyx) {
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

void ref_cube_free(void* cube) { delete static_cast<DataCube*>(cube); }

// ---- CMEM
void* ref_cmem_create_default() {
    CMEM* m = new CMEM();
    m->Init();
    return m;
}

// The 17-argument Init (include/Cmem.h:21-39); has_v/has_b = 0 passes empty vectors.
void* ref_cmem_create(int has_v, const int* v, int has_b, const int* b, float dipole, float p0,
                      int p1, int p2, int p3, float B, float alpha, float beta, float bs, float A1,
                      float A2, float ay_bs, float az_bs, float density) {
    CMEM* m = new CMEM();
    std::vector<int> vv, bb;
    if (has_v) vv.assign(v, v + 3);
    if (has_b) bb.assign(b, b + 3);
    m->Init(true, vv, bb, dipole, p0, p1, p2, p3, B, alpha, beta, bs, A1, A2, ay_bs, az_bs,
            density);
    return m;
}

// derived constants, in the order of birp_cmem_params (include/birp_b200.h)
void ref_cmem_constants(void* cmem, float* out27) {
    CMEM* m = static_cast<CMEM*>(cmem);
    int n = 0;
    out27[n++] = m->p0;
    out27[n++] = m->p1;
    out27[n++] = m->p2;
    out27[n++] = m->p3;
    out27[n++] = m->bs;
    out27[n++] = m->A1;
    out27[n++] = m->A2;
    out27[n++] = m->B;
    out27[n++] = m->alpha;
    out27[n++] = m->beta;
    out27[n++] = m->ay_bs;
    out27[n++] = m->az_bs;
    out27[n++] = m->r0_lin;
    out27[n++] = m->charlie;
    out27[n++] = m->dn;
    out27[n++] = m->ds;
    out27[n++] = m->theta_n;
    out27[n++] = m->theta_s;
    out27[n++] = m->a[5];
    out27[n++] = m->a[21];
    out27[n++] = m->beta_c[0];
    out27[n++] = m->beta_c[1];
    out27[n++] = m->beta_c[2];
    out27[n++] = m->beta_c[3];
    out27[n++] = m->pressure_dynamic;
    out27[n++] = m->pressure_magnetic;
    out27[n++] = m->density;
}

void ref_cmem_free(void* cmem) { delete static_cast<CMEM*>(cmem); }

// ---- Space::GetSample on a batch of points (xyz interleaved)
int ref_space_sample(void* space_ptr, int is_cmem, const float* xyz, long n, float* out) {
    try {
        Space* s = is_cmem ? static_cast<Space*>(static_cast<CMEM*>(space_ptr))
                           : static_cast<Space*>(static_cast<DataCube*>(space_ptr));
        for (long i = 0; i < n; ++i) out[i] = s->GetSample(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// ---- one frame through the reference Camera.
// W/H <= 0 keeps the reference's own round(fov/px) sizes (src/Camera.cpp:38-39).
// nsamples = 200 and maxdist = 44 reproduce Camera::Render() (src/Camera.cpp:63-67).
// image is ACCUMULATED INTO, like the reference's image member (src/Camera.cpp:223).
int ref_render3(void* space_ptr, int is_cmem, float px_deg, float fov_h, float fov_w, int W, int H,
                const float* pos, const float* aim, int nsamples, int maxdist, float* image,
                int* out_wh, double* seconds);
int ref_render(void* space_ptr, int is_cmem, float px_deg, float fov_h, float fov_w, int W, int H,
               const float* pos, float aim_x, int nsamples, int maxdist, float* image,
               int* out_wh, double* seconds) {
    const float aim[3] = {aim_x, 0.f, 0.f};
    return ref_render3(space_ptr, is_cmem, px_deg, fov_h, fov_w, W, H, pos, aim, nsamples, maxdist, image, out_wh, seconds);
}

// the same with the whole aim vector of Camera::SetAim
int ref_render3(void* space_ptr, int is_cmem, float px_deg, float fov_h, float fov_w, int W, int H,
                const float* pos, const float* aim, int nsamples, int maxdist, float* image,
                int* out_wh, double* seconds) {
    try {
        Space* s = is_cmem ? static_cast<Space*>(static_cast<CMEM*>(space_ptr))
                           : static_cast<Space*>(static_cast<DataCube*>(space_ptr));
        Camera cam(*s, px_deg, fov_h, fov_w);
        if (W > 0 && H > 0) {
            cam.image_dimension_x = W;
            cam.image_dimension_y = H;
            cam.image.assign(H, std::vector<float>(W, 0.f));
        }
        const int w = cam.image_dimension_x, h = cam.image_dimension_y;
        if (out_wh) {
            out_wh[0] = w;
            out_wh[1] = h;
        }
        if (!image) return 0;  // size query only
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i) cam.image[j][i] = image[(size_t)j * w + i];
        float px = pos[0], py = pos[1], pz = pos[2];
        cam.SetPosition(px, py, pz);
        cam.SetAim(aim[0], aim[1], aim[2]);
        auto t0 = std::chrono::steady_clock::now();
        if (nsamples == 200 && maxdist == 44) {
            cam.Render();
        } else {
            cam.ray_samples = nsamples;
            cam.GenerateRayDistWidth(maxdist);
            cam.Integrate();
        }
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i) image[(size_t)j * w + i] = cam.image[j][i];
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// Camera::Orient() alone (src/Camera.cpp:309-417), for pinning the host-side frame setup.
int ref_orient3(const float* pos, const float* aim, float* angle_out);
int ref_orient(const float* pos, float aim_x, float* angle_out) {
    const float aim[3] = {aim_x, 0.f, 0.f};
    return ref_orient3(pos, aim, angle_out);
}

int ref_orient3(const float* pos, const float* aim, float* angle_out) {
    try {
        struct Dummy : Space {
            void Init() override {}
            float GetSample(float, float, float) override { return 0.f; }
        } dummy;
        Camera cam(dummy, 1.f, 1.f, 1.f);
        float px = pos[0], py = pos[1], pz = pos[2];
        cam.SetPosition(px, py, pz);
        cam.SetAim(aim[0], aim[1], aim[2]);
        *angle_out = cam.Orient();
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

// The sample points the reference Camera visits: a recording Space returns 0 and
// stores every (x,y,z) it is asked for.  Used to pin ray generation bit-for-bit.
int ref_trace_points(float px_deg, float fov_h, float fov_w, int W, int H, const float* pos,
                     float aim_x, int nsamples, int maxdist, float* xyz_out /* H*W*nsamples*3 */) {
    try {
        struct Recorder : Space {
            float* out;
            size_t n = 0;
            void Init() override {}
            float GetSample(float x, float y, float z) override {
                out[n++] = x;
                out[n++] = y;
                out[n++] = z;
                return 0.f;
            }
        } rec;
        rec.out = xyz_out;
        Camera cam(rec, px_deg, fov_h, fov_w);
        if (W > 0 && H > 0) {
            cam.image_dimension_x = W;
            cam.image_dimension_y = H;
            cam.image.assign(H, std::vector<float>(W, 0.f));
        }
        float px = pos[0], py = pos[1], pz = pos[2];
        cam.SetPosition(px, py, pz);
        cam.SetAim(aim_x, 0.f, 0.f);
        cam.ray_samples = nsamples;
        cam.GenerateRayDistWidth(maxdist);
        cam.Integrate();
    } catch (const std::exception& e) {
        return fail(e);
    }
    return 0;
}

}  // extern "C"
