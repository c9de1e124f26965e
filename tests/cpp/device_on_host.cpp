// TEST INFRASTRUCTURE: the device code of the march kernels (birp_cpp_b200/csrc/birp_device.cuh) compiled for the
// HOST by g++ through tests/cpp/cuda_on_host.h, behind a small C interface for tests/test_device_on_host.py.
//
// What runs here is the source the GPU runs: cube_sample / cube_sample_hot (every lerp class, division mode and
// corner fetch), scale_fast / scale_sample, march_steps and march_steps_span (the hot loop with its cold exits,
// the bisection for the end of a ray's hot run), pixel_direction, and the CMEM samplers with their fp64 blocks.
// What is restated here, because it lives in the __global__ kernels (birp_kernels.cu: march_cube_kernel,
// pack_quads_kernel, pack_octs_kernel, scan_positive_kernel) or in the upload code (birp_capi.cu: finish_cube), is
// the glue around it: the per-ray prologue, the staging of the ray table in chunks of kSmemDist entries, the
// copies of the grid, the classification of the voxels.  The window limits come from the product's own host
// functions (birp_hostmath.cpp, linked from libbirp_b200.so).
//
// This is a checker of device SOURCE on the CPU; it is not a product path (nothing in birp_cpp_b200/ uses it).
#include "cuda_on_host.h"

#include <thread>
#include <vector>

#include "../../birp_cpp_b200/csrc/birp_device.cuh"
#include "../../birp_cpp_b200/csrc/birp_host.h"

namespace {

thread_local const float* t_table = nullptr;

struct HostTex {
    const float* v;
    int nx, ny;
};

struct HostCube {
    birp::CubeView view;
    std::vector<float> grid, quad, oct, oct_marked, poly;  // oct_marked: the oct copy as the GPU builds it for a cube
                                                           // outside the positive class (CubeView::oct_marked)
    HostTex tex;
};

}  // namespace

float birp_host_table_entry(unsigned sdist) {
    float t;
    std::memcpy(&t, reinterpret_cast<const char*>(t_table) + sdist, sizeof(t));
    return t;
}

// tld4 on the layered texture: the four texels around the footprint centre (u, v) = (x0 + 1, y0 + 1) of layer `layer`
void birp_host_gather4(cudaTextureObject_t tex, int layer, float u, float v, float& t01, float& t11, float& t10,
                       float& t00) {
    const HostTex* t = reinterpret_cast<const HostTex*>(tex);
    const int x0 = (int)u - 1, y0 = (int)v - 1;
    const float* p = t->v + ((size_t)layer * t->ny + y0) * t->nx + x0;
    t00 = p[0];
    t10 = p[1];
    t01 = p[t->nx];
    t11 = p[t->nx + 1];
}

namespace {

using birp::CubeView;
using birp::RayState;

// march_cube_kernel's body for one ray (birp_kernels.cu), with the table staged chunk by chunk
template <bool TRI, int DIV, int LERP, int FETCH, int SPAN>
void march_ray(const CubeView& c, const birp_frame& fr, int i, int j, int W, int H,
               const std::vector<std::vector<float>>& staged, const std::vector<int>& chunk_len, float ray_width,
               int accumulate, float* pix, unsigned long long* valid, int* status) {
    RayState r;
    birp::pixel_direction(fr, i, j, (float)W, (float)H, r.dx, r.dy, r.dz);
    r.px = fr.pos[0];
    r.py = fr.pos[1];
    r.pz = fr.pos[2];
    r.acc = accumulate ? *pix : 0.f;
    bool alive = true;
    unsigned nlive4 = 0;
    const double K = BIRP_SCALE_K;
    for (size_t k = 0; k < staged.size() && alive; ++k) {
        t_table = staged[k].data();
        const unsigned send = 4u * (unsigned)chunk_len[k];
        nlive4 += SPAN ? birp::march_steps_span<TRI, DIV, LERP, FETCH, SPAN == 2>(c, r, 0u, send, ray_width, K, status, alive)
                       : birp::march_steps<TRI, DIV, LERP, FETCH>(c, r, 0u, send, ray_width, K, status, alive);
    }
    *pix = r.acc;
    *valid += nlive4 >> 2;
}

struct RenderJob {
    const HostCube* cube;
    const birp_frame* frames;
    int nframes, W, H, nsamples, maxdist, accumulate;
    float* images;
    unsigned long long valid = 0;
    int status = 0;
};

template <bool TRI, int DIV, int LERP, int FETCH, int SPAN>
void render_variant(RenderJob& job) {
    CubeView c = job.cube->view;
    if (!TRI) c.lox = c.loy = c.loz = BIRP_TINY_INGRESS;  // launch_render_cube: nearest keeps the first cells hot
    if (LERP == 3) {  // the mark-aware lerp form reads the marked oct copy
        c.oct = job.cube->oct_marked.data();
        c.oct_marked = 1;
    }
    std::vector<float> dist(job.nsamples);
    float ray_width = 0.f;
    birp::host_ray_table(job.nsamples, job.maxdist, dist.data(), &ray_width);
    // the shared-memory table as the CTA sees it in each chunk: kSmemDist entries + the look-ahead padding entry;
    // a shorter last chunk leaves the tail of the previous one in place
    std::vector<std::vector<float>> staged;
    std::vector<int> chunk_len;
    std::vector<float> cur(birp::kSmemDist + 1, 0.f);
    for (int base = 0; base < job.nsamples; base += birp::kSmemDist) {
        const int n = std::min(birp::kSmemDist, job.nsamples - base);
        std::copy(dist.begin() + base, dist.begin() + base + n, cur.begin());
        staged.push_back(cur);
        chunk_len.push_back(n);
    }
    const int nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<unsigned long long> valid(nthreads, 0);
    std::vector<int> status(nthreads, 0);
    std::vector<std::thread> pool;
    const long rows_total = (long)job.nframes * job.H;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            for (long row = t; row < rows_total; row += nthreads) {
                const int f = (int)(row / job.H), j = (int)(row % job.H);
                for (int i = 0; i < job.W; ++i)
                    march_ray<TRI, DIV, LERP, FETCH, SPAN>(c, job.frames[f], i, j, job.W, job.H, staged, chunk_len, ray_width,
                                                           job.accumulate,
                                                           job.images + ((size_t)f * job.H + j) * job.W + i, &valid[t],
                                                           &status[t]);
            }
        });
    for (auto& th : pool) th.join();
    for (int t = 0; t < nthreads; ++t) job.valid += valid[t], job.status |= status[t];
}

template <bool TRI, int DIV, int LERP, int FETCH>
void render_span(RenderJob& job, int span) {
    // span 2 = the lean march (BIRP_LEAN_MARCH): built, like the kernel, for positive cubes and oct records only;
    // the caller is responsible for launch_render_cube's conditions (grid size, value range)
    if (span == 2 && TRI && LERP == 1 && FETCH == 5) render_variant<TRI, DIV, LERP, FETCH, (TRI && LERP == 1 && FETCH == 5) ? 2 : 1>(job);
    else if (span) render_variant<TRI, DIV, LERP, FETCH, 1>(job);
    else render_variant<TRI, DIV, LERP, FETCH, 0>(job);
}

template <bool TRI, int DIV, int LERP>
bool render_fetch(RenderJob& job, int fetch, int span) {
    if (!TRI) return render_span<false, DIV, LERP, 0>(job, span), true;
    switch (fetch) {
        case 0: render_span<TRI, DIV, LERP, 0>(job, span); return true;
        case 1: render_span<TRI, DIV, LERP, 1>(job, span); return true;
        case 2: render_span<TRI, DIV, LERP, 2>(job, span); return true;
        case 4: render_span<TRI, DIV, LERP, 4>(job, span); return true;
        case 5: render_span<TRI, DIV, LERP, 5>(job, span); return true;
    }
    return false;
}

template <bool TRI, int DIV>
bool render_lerp(RenderJob& job, int lerp, int fetch, int span) {
    switch (lerp) {
        case 0: return render_fetch<TRI, DIV, 0>(job, fetch, span);
        case 1: return render_fetch<TRI, DIV, 1>(job, fetch, span);
        case 2: return render_fetch<TRI, DIV, 2>(job, fetch, span);
        case 3:  // marked oct records: exists, like the kernel, for the oct fetch and the span march only
            if (!TRI || fetch != 5 || span != 1) return false;
            render_variant<TRI, DIV, 3, 5, 1>(job);
            return true;
    }
    return false;
}

template <bool TRI>
bool render_div(RenderJob& job, int div, int lerp, int fetch, int span) {
    switch (div) {
        case 0: return render_lerp<TRI, 0>(job, lerp, fetch, span);
        case 1: return render_lerp<TRI, 1>(job, lerp, fetch, span);
        case 2: return render_lerp<TRI, 2>(job, lerp, fetch, span);
        case 3: return render_lerp<TRI, 3>(job, lerp, fetch, span);
    }
    return false;
}

// march_cube_fast_kernel over whole frames (BIRP_FAST_MARCH): per-ray prologue and epilogue restated, the march is
// the device source (march_steps_fast).  variant: fast_sample<VARIANT>; signs < 0: as the kernel chooses (class 0).
template <bool SIGNS, int VARIANT, bool KSTEP>
void fast_variant(RenderJob& job) {
    const CubeView c = job.cube->view;
    std::vector<float> dist(job.nsamples);
    float ray_width = 0.f;
    birp::host_ray_table(job.nsamples, job.maxdist, dist.data(), &ray_width);
    std::vector<std::vector<float>> staged;
    std::vector<int> chunk_len;
    std::vector<float> cur(birp::kSmemDist + 1, 0.f);
    for (int base = 0; base < job.nsamples; base += birp::kSmemDist) {
        const int n = std::min(birp::kSmemDist, job.nsamples - base);
        std::copy(dist.begin() + base, dist.begin() + base + n, cur.begin());
        staged.push_back(cur);
        chunk_len.push_back(n);
    }
    const int nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<unsigned long long> valid(nthreads, 0);
    std::vector<int> status(nthreads, 0);
    std::vector<std::thread> pool;
    const long rows_total = (long)job.nframes * job.H;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            for (long row = t; row < rows_total; row += nthreads) {
                const int f = (int)(row / job.H), j = (int)(row % job.H);
                for (int i = 0; i < job.W; ++i) {
                    RayState r;
                    birp::pixel_direction(job.frames[f], i, j, (float)job.W, (float)job.H, r.dx, r.dy, r.dz);
                    r.px = job.frames[f].pos[0], r.py = job.frames[f].pos[1], r.pz = job.frames[f].pos[2];
                    r.acc = 0.f;
                    const birp::FastRay q = birp::fast_ray(c, r, KSTEP ? (double)job.maxdist / (double)job.nsamples : 1.0);
                    bool alive = true;
                    unsigned nlive4 = 0;
                    for (size_t k = 0; k < staged.size() && alive; ++k) {
                        t_table = staged[k].data();
                        nlive4 += birp::march_steps_fast<SIGNS, VARIANT, KSTEP>(c, r, q, 0u, 4u * (unsigned)chunk_len[k],
                                                                                ray_width, &status[t], alive, 0u,
                                                                                (int)k * birp::kSmemDist);
                    }
                    float* pix = job.images + ((size_t)f * job.H + j) * job.W + i;
                    const float total = (float)((double)(r.acc * ray_width) * BIRP_SCALE_K);
                    *pix = job.accumulate ? *pix + total : total;
                    valid[t] += nlive4 >> 2;
                }
            }
        });
    for (auto& th : pool) th.join();
    for (int t = 0; t < nthreads; ++t) job.valid += valid[t], job.status |= status[t];
}

}  // namespace

extern "C" {

// The device view of a cube as birp_capi.cu (finish_cube) and the pack / scan kernels build it.  `divmode` is the
// caller's choice: on the GPU it comes from the exhaustive device-side comparison with IEEE division, which
// doh_verify_div repeats here on a stride of the ingresses.
void* doh_cube_create(const float* zyx, int nx, int ny, int nz, const float origin[3], const float spacing[3],
                      int divmode) {
    HostCube* h = new HostCube();
    const size_t n = (size_t)nx * ny * nz, slab = (size_t)nx * ny;
    h->grid.assign(zyx, zyx + n);
    CubeView& v = h->view;
    std::memset(&v, 0, sizeof(v));
    v.v = h->grid.data();
    v.nx = nx, v.ny = ny, v.nz = nz;
    v.slab = nx * ny;
    v.slab16 = 16LL * nx * ny;
    v.fnx = (float)nx, v.fny = (float)ny, v.fnz = (float)nz;
    v.ox = origin[0], v.oy = origin[1], v.oz = origin[2];
    v.sx = spacing[0], v.sy = spacing[1], v.sz = spacing[2];
    v.rx64 = 1.0 / (double)spacing[0], v.ry64 = 1.0 / (double)spacing[1], v.rz64 = 1.0 / (double)spacing[2];
    const volatile float rx = 1.0f / spacing[0], ry = 1.0f / spacing[1], rz = 1.0f / spacing[2];
    v.rx = rx, v.ry = ry, v.rz = rz;
    v.lx = birp::host_bounds_limit(spacing[0], nx);
    v.ly = birp::host_bounds_limit(spacing[1], ny);
    v.lz = birp::host_bounds_limit(spacing[2], nz);
    v.lox = birp::host_first_cell_limit(spacing[0]);
    v.loy = birp::host_first_cell_limit(spacing[1]);
    v.loz = birp::host_first_cell_limit(spacing[2]);
    v.divmode = divmode;
    int bad = 0;  // scan_positive_kernel
    for (size_t t = 0; t < n; ++t) {
        const float x = zyx[t];
        const bool pos = x > 0.f && x < CUDART_INF_F;
        uint32_t bits;
        std::memcpy(&bits, &x, 4);
        if (!pos) bad |= 1;
        if (!pos && bits != 0u) bad |= 2;
    }
    v.positive = (bad == 0) ? 1 : ((bad & 2) == 0) ? 2 : 0;
    h->tex = HostTex{h->grid.data(), nx, ny};
    v.tex = reinterpret_cast<cudaTextureObject_t>(&h->tex);
    if (nx >= 2 && ny >= 2 && nz >= 2) {
        h->quad.resize(4 * n);  // pack_quads_kernel
        h->oct.resize(8 * n);   // pack_octs_kernel
        for (size_t t = 0; t < n; ++t) {
            const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)(t / slab);
            const int dx = (x + 1 < nx) ? 1 : 0;
            const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
            const size_t u = t + ((z + 1 < nz) ? slab : 0);
            const float lo[4] = {zyx[t], zyx[t + dx], zyx[t + dy], zyx[t + dy + dx]};
            const float hi[4] = {zyx[u], zyx[u + dx], zyx[u + dy], zyx[u + dy + dx]};
            std::memcpy(&h->quad[4 * t], lo, sizeof(lo));
            std::memcpy(&h->oct[8 * t], lo, sizeof(lo));
            std::memcpy(&h->oct[8 * t + 4], hi, sizeof(hi));
        }
        h->oct_marked = h->oct;  // pack_octs_kernel with mark = 1
        for (size_t t = 0; t < n; ++t) {
            const float* r = &h->oct[8 * t];
            const int cls = birp::oct_record_class(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
            if (cls == 1) h->oct_marked[8 * t] = -r[0];
            if (cls == 2) h->oct_marked[8 * t] = __int_as_float(0x7fc00000);
        }
        h->poly.resize(8 * n);  // pack_poly_kernel
        for (size_t t = 0; t < n; ++t) {
            const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)(t / slab);
            const int dx = (x + 1 < nx) ? 1 : 0;
            const size_t dy = (y + 1 < ny) ? (size_t)nx : 0;
            const size_t u = t + ((z + 1 < nz) ? slab : 0);
            birp::poly_coefficients(zyx[t], zyx[t + dx], zyx[t + dy], zyx[t + dy + dx], zyx[u], zyx[u + dx], zyx[u + dy],
                                    zyx[u + dy + dx], &h->poly[8 * t]);
        }
        v.poly = h->poly.data();
        v.poly_bytes = 32;
        const unsigned B = __float_as_uint(BIRP_RINT_MAGIC);
        v.poly_bias = (int)((B * (unsigned)ny + B) * (unsigned)nx + B);  // (march_cube_fast_kernel sets these two itself)
        v.quad = reinterpret_cast<const float4*>(h->quad.data());
        v.quad_up = v.quad + slab;  // (march_cube_kernel sets these three itself)
        v.quad_bytes = 16;
        v.oct = h->oct.data();
    }
    return h;
}

void doh_cube_free(void* cube) { delete static_cast<HostCube*>(cube); }

// 1 all voxels finite and > 0, 2 finite and > 0 or +0, 0 otherwise
int doh_cube_class(const void* cube) { return static_cast<const HostCube*>(cube)->view.positive; }

void doh_cube_window(const void* cube, float lo[3], float hi[3]) {
    const CubeView& v = static_cast<const HostCube*>(cube)->view;
    lo[0] = v.lox, lo[1] = v.loy, lo[2] = v.loz;
    hi[0] = v.lx, hi[1] = v.ly, hi[2] = v.lz;
}

// verify_div_kernel on every `stride`-th fp32 ingress of the hot window of one axis: bad[0] / bad[1] / bad[2] = 1 when
// div_spacing<0> / <1> / <3> (the multiply alone) ever differs from IEEE division
void doh_verify_div(float s, int n, unsigned stride, int bad[3]) {
    const volatile float rv = 1.0f / s;
    const float r = rv;
    const float limit = birp::host_bounds_limit(s, n);
    uint32_t first, last;
    const float tiny = BIRP_TINY_INGRESS;
    std::memcpy(&first, &tiny, 4);
    std::memcpy(&last, &limit, 4);
    bad[0] = bad[1] = bad[2] = 0;
    for (uint64_t u = first; u < last; u += stride) {
        const uint32_t b = (uint32_t)u;
        float a;
        std::memcpy(&a, &b, 4);
        const float want = a / s;
        if (birp::div_spacing<0>(a, s, r) != want) bad[0] = 1;
        if (birp::div_spacing<1>(a, s, r) != want) bad[1] = 1;
        if (birp::div_spacing<3>(a, s, r) != want) bad[2] = 1;
    }
}

// DataCube::GetSample on points: literal device transcription (cube_sample) and, where the point lies in the hot
// window, the hot form (cube_sample_hot<lerp class, divmode, fetch>); hot_mask[p] says which points were hot.
// lerp < 0: the class of the cube.
int doh_sample_cube(const void* cube, int trilinear, int lerp, int fetch, const float* xyz, long n, float* literal,
                    float* hot, unsigned char* hot_mask) {
    const HostCube* h = static_cast<const HostCube*>(cube);
    CubeView c = h->view;
    if (!trilinear) c.lox = c.loy = c.loz = BIRP_TINY_INGRESS;
    if (lerp < 0) lerp = c.positive;
    int status = 0;
    for (long p = 0; p < n; ++p) {
        const float x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
        bool nan_index = false;
        literal[p] = trilinear ? birp::cube_sample<true>(c, x, y, z, nan_index) : birp::cube_sample<false>(c, x, y, z, nan_index);
        if (nan_index) status |= birp::kStatusNanIndex;
        const float xi = x - c.ox, yi = y - c.oy, zi = z - c.oz;
        const bool is_hot = birp::cube_hot_window(c, xi, yi, zi);
        hot_mask[p] = is_hot;
        hot[p] = 0.f;
        if (!is_hot) continue;
#define DOH_HOT(TRI, DIV, LERP, FETCH) hot[p] = birp::cube_sample_hot<TRI, DIV, LERP, FETCH>(c, xi, yi, zi)
#define DOH_HOT_FETCH(TRI, DIV, LERP)              \
    switch (TRI ? fetch : 0) {                     \
        case 1: DOH_HOT(TRI, DIV, LERP, 1); break; \
        case 2: DOH_HOT(TRI, DIV, LERP, 2); break; \
        case 4: DOH_HOT(TRI, DIV, LERP, 4); break; \
        case 5: DOH_HOT(TRI, DIV, LERP, 5); break; \
        default: DOH_HOT(TRI, DIV, LERP, 0); break; \
    }
#define DOH_HOT_LERP(TRI, DIV)                        \
    switch (lerp) {                                   \
        case 1: DOH_HOT_FETCH(TRI, DIV, 1); break;    \
        case 2: DOH_HOT_FETCH(TRI, DIV, 2); break;    \
        default: DOH_HOT_FETCH(TRI, DIV, 0); break;   \
    }
#define DOH_HOT_DIV(TRI)                         \
    switch (c.divmode) {                         \
        case 0: DOH_HOT_LERP(TRI, 0); break;     \
        case 1: DOH_HOT_LERP(TRI, 1); break;     \
        case 3: DOH_HOT_LERP(TRI, 3); break;     \
        default: DOH_HOT_LERP(TRI, 2); break;    \
    }
        if (trilinear) { DOH_HOT_DIV(true) } else { DOH_HOT_DIV(false) }
    }
    return status;
}

// march_cube_kernel over whole frames.  lerp / divmode < 0: the cube's.  Returns -1 for an unknown variant, else
// the status word (kStatusNanIndex when the reference would throw).
int doh_render_cube(const void* cube, int trilinear, int lerp, int divmode, int fetch, int span, const birp_frame* frames,
                    int nframes, int W, int H, int nsamples, int maxdist, int accumulate, float* images,
                    unsigned long long* valid) {
    const HostCube* h = static_cast<const HostCube*>(cube);
    RenderJob job{h, frames, nframes, W, H, nsamples, maxdist, accumulate, images};
    if (lerp < 0) lerp = h->view.positive;
    if (divmode < 0) divmode = h->view.divmode;
    if (!trilinear && lerp == 2) lerp = 0;  // launch_cube_div: nearest has the classes 1 and 0 only
    const bool ok = trilinear ? render_div<true>(job, divmode, lerp, fetch, span) : render_div<false>(job, divmode, lerp, 0, span);
    if (!ok) return -1;
    if (valid) *valid = job.valid;
    return job.status;
}

int doh_render_cube_fast(const void* cube, int variant, int signs, int kstep, const birp_frame* frames, int nframes, int W,
                                    int H, int nsamples, int maxdist, int accumulate, float* images,
                                    unsigned long long* valid) {
    const HostCube* h = static_cast<const HostCube*>(cube);
    RenderJob job{h, frames, nframes, W, H, nsamples, maxdist, accumulate, images};
    if (signs < 0) signs = h->view.positive == 0;
#define DOH_FAST(V)                              \
    case V:                                      \
        if (signs) kstep ? fast_variant<true, V, true>(job) : fast_variant<true, V, false>(job);    \
        else kstep ? fast_variant<false, V, true>(job) : fast_variant<false, V, false>(job);         \
        break;
    switch (variant) {
        DOH_FAST(0) DOH_FAST(1) DOH_FAST(2) DOH_FAST(3) DOH_FAST(4)
        default: return -1;
    }
#undef DOH_FAST
    if (valid) *valid = job.valid;
    return job.status;
}

// the exactness flag of the hot loop's scale: for each s, y = scale_fast's value, exact = its claim, chain = the
// literal three-step fp64 chain of the reference (scale_sample_chain), full = scale_sample (what the cold path uses)
void doh_scale(const float* s, long n, float ray_width, int maybe_zero, float* y, unsigned char* exact, float* chain,
               float* full) {
    for (long p = 0; p < n; ++p) {
        float v;
        const bool e = maybe_zero ? birp::scale_fast<true>(s[p], ray_width, BIRP_SCALE_K, v)
                                  : birp::scale_fast<false>(s[p], ray_width, BIRP_SCALE_K, v);
        y[p] = v;
        exact[p] = e;
        chain[p] = birp::scale_sample_chain(s[p] * ray_width);
        full[p] = birp::scale_sample(s[p], ray_width);
    }
}

// The scale shortcut over EVERY fp32 product (all 2^32 bit patterns of p = sample * ray_width, here sample = p and
// ray_width = 1): count[0] = products the hot loop accepts (scale_fast says exact), count[1] = accepted products
// whose value differs from the reference's three-step chain (must be 0), count[2] = accepted products that are
// negative or NaN after the chain (must be 0: those end the ray or poison the pixel, which only the cold path
// handles), count[3] = finite products in [2^-120, 2^120) that the flag sends to the cold path (the near-midpoint
// ones), count[4] = products whose chain value lies in [2^-120, 2^120) and differs from the single multiply,
// whatever the flag says (0 means: inside that range the single multiply IS the chain, for every input -- what
// scale_fast<.., RANGE_SAFE> relies on).  `stride` > 1 checks every stride-th pattern.
void doh_scale_exhaustive(int maybe_zero, unsigned stride, int nthreads, unsigned long long count[5]) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::vector<unsigned long long>> part(nthreads, std::vector<unsigned long long>(5, 0));
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            auto& c = part[t];
            for (uint64_t u = (uint64_t)t * stride; u < (1ull << 32); u += (uint64_t)nthreads * stride) {
                const uint32_t b = (uint32_t)u;
                float p;
                std::memcpy(&p, &b, 4);
                float y;
                const bool e = maybe_zero ? birp::scale_fast<true>(p, 1.0f, BIRP_SCALE_K, y)
                                          : birp::scale_fast<false>(p, 1.0f, BIRP_SCALE_K, y);
                const float chain = birp::scale_sample_chain(p);
                if (chain >= BIRP_SCALE_SAFE_LO && chain < BIRP_SCALE_SAFE_HI) {
                    uint32_t by, bc;
                    std::memcpy(&by, &y, 4);
                    std::memcpy(&bc, &chain, 4);
                    if (by != bc) ++c[4];
                }
                if (e) {
                    ++c[0];
                    uint32_t by, bc;
                    std::memcpy(&by, &y, 4);
                    std::memcpy(&bc, &chain, 4);
                    if (by != bc) ++c[1];
                    if (!(chain >= 0.f)) ++c[2];
                } else if (chain >= BIRP_SCALE_SAFE_LO && chain < BIRP_SCALE_SAFE_HI) {
                    ++c[3];
                }
            }
        });
    for (auto& th : pool) th.join();
    for (int k = 0; k < 5; ++k) {
        count[k] = 0;
        for (int t = 0; t < nthreads; ++t) count[k] += part[t][k];
    }
}

// the seven-lerp building blocks on raw operands: std::lerp transcription, positive form, non-negative form
void doh_lerps(const float* a, const float* b, const float* t, long n, float* ref, float* positive, float* nonneg) {
    for (long p = 0; p < n; ++p) {
        ref[p] = birp::lerp_ref(a[p], b[p], t[p]);
        positive[p] = a[p] + t[p] * (b[p] - a[p]);
        nonneg[p] = birp::lerp_nonneg(a[p], b[p], t[p], 1.f - t[p]);
    }
}

// CMEM::GetSample on points: mode 0 fp32 libm calls, 1 fp64 libm rounded once, 2 the library's own fp64 blocks,
// 3 the same behind the region filter (the default of the kernels)
void doh_sample_cmem(const birp_cmem_params* params, const float* xyz, long n, int mode, float* out) {
    const birp::CmemDev D = birp::make_cmem_dev(*params);
    for (long p = 0; p < n; ++p) {
        const float x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
        switch (mode) {
            case 0: out[p] = birp::cmem_sample_mode<0>(D, x, y, z); break;
            case 1: out[p] = birp::cmem_sample_mode<1>(D, x, y, z); break;
            case 2: out[p] = birp::cmem_sample_mode<2>(D, x, y, z); break;
            default: out[p] = birp::cmem_sample_mode<3>(D, x, y, z); break;
        }
    }
}

// march_kernel<CmemField<mode>> over whole frames (birp_kernels.cu; the loop is restated, the sampler and the scale
// are the device source)
void doh_render_cmem(const birp_cmem_params* params, int mode, const birp_frame* frames, int nframes, int W, int H,
                     int nsamples, int maxdist, float* images) {
    const birp::CmemDev D = birp::make_cmem_dev(*params);
    std::vector<float> dist(nsamples);
    float ray_width = 0.f;
    birp::host_ray_table(nsamples, maxdist, dist.data(), &ray_width);
    const int nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    const long rows_total = (long)nframes * H;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            for (long row = t; row < rows_total; row += nthreads) {
                const int f = (int)(row / H), j = (int)(row % H);
                const birp_frame& fr = frames[f];
                for (int i = 0; i < W; ++i) {
                    float dx, dy, dz;
                    birp::pixel_direction(fr, i, j, (float)W, (float)H, dx, dy, dz);
                    float acc = 0.f;
                    for (int k = 0; k < nsamples; ++k) {
                        const float tk = dist[k];
                        const float x = dx * tk + fr.pos[0], y = dy * tk + fr.pos[1], z = dz * tk + fr.pos[2];
                        float s;
                        switch (mode) {
                            case 0: s = birp::cmem_sample_mode<0>(D, x, y, z); break;
                            case 1: s = birp::cmem_sample_mode<1>(D, x, y, z); break;
                            case 2: s = birp::cmem_sample_mode<2>(D, x, y, z); break;
                            case 4: s = birp::cmem_sample_mode<4>(D, x, y, z); break;
                            default: s = birp::cmem_sample_mode<3>(D, x, y, z); break;
                        }
                        if (s < 0.f) break;
                        acc += s;  // (CmemField::kScaleOnce: raw samples summed, the scale applied to the sum)
                    }
                    images[((size_t)f * H + j) * W + i] = (float)((double)(acc * ray_width) * BIRP_SCALE_K);
                }
            }
        });
    for (auto& th : pool) th.join();
}

// cmem_table_kernel on the host: the same per-cell function (cmem_table_cell), over all cells.
// table: kCmemTableT * kCmemTableP x {mp_lo, mp_hi, bs_lo, bs_hi}.
void doh_cmem_table(const birp_cmem_params* params, float* table) {
    const birp::CmemDev D = birp::make_cmem_dev(*params);
    const int ncell = birp::kCmemTableT * birp::kCmemTableP;
    const int nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            for (int cell = t; cell < ncell; cell += nthreads)
                birp::cmem_table_cell(D, cell / birp::kCmemTableP, cell % birp::kCmemTableP, table + 4 * cell);
        });
    for (auto& th : pool) th.join();
}

// CMEM::GetSample with the region table in front of the filter (mode 3 + table), and whether the table decided
void doh_sample_cmem_table(const birp_cmem_params* params, const float* table, const float* xyz, long n, float* out,
                           unsigned char* decided) {
    birp::CmemDev D = birp::make_cmem_dev(*params);
    D.table = reinterpret_cast<const float4*>(table);
    for (long p = 0; p < n; ++p) {
        const float x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
        out[p] = birp::cmem_sample_mode<3>(D, x, y, z);
        float r, q_t, q_p;
        birp::cmem_cosines_estimate(x, y, z, r, q_t, q_p);
        const float* tb = table + 4 * birp::cmem_table_index(q_t, q_p);
        decided[p] = (r < tb[0] || (r > tb[1] && (r > tb[3] || r < tb[2]))) ? 1 : 0;
    }
}

// which samples the region filter decides on its own (1) and which go to the full evaluation (0)
void doh_cmem_filter_decides(const birp_cmem_params* params, const float* xyz, long n, unsigned char* decided) {
    const birp::CmemDev D = birp::make_cmem_dev(*params);
    for (long p = 0; p < n; ++p) {
        float value;
        decided[p] = birp::cmem_try_filter<false>(D, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], value) == birp::kCmemDecided;
    }
}

// r / 10 as the CMEM kernel takes it (div10 for an ordinary r, the division otherwise) against the division, for the
// bit patterns first, first + stride, ... below last: bad[0] = mismatches of the kernel's form, bad[1] = of div10 alone
void doh_div10_check(unsigned first, unsigned last, unsigned stride, unsigned long long bad[2]) {
    unsigned long long n0 = 0, n1 = 0;
    for (unsigned long long u = first; u < last; u += stride) {
        const float r = __uint_as_float((unsigned)u);
        const float want = r / 10.f;
        const float fast = birp::div10(r);
        const float used = birp::cmem_r_is_ordinary(r) ? fast : want;
        if (want == want && __float_as_uint(used) != __float_as_uint(want)) ++n0;
        if (want == want && __float_as_uint(fast) != __float_as_uint(want)) ++n1;
    }
    bad[0] = n0;
    bad[1] = n1;
}

}  // extern "C"
