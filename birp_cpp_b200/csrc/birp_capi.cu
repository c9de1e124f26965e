// C ABI of libbirp_b200.so (declared in include/birp_b200.h): device contexts, cube
// residency, pinned staging, chunked host<->device pipelines around the kernels.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "birp_host.h"
#include "birp_internal.h"

struct birp_cube {
    birp::CubeView view;
    float* d_data;
    int device;
    size_t voxels;
    cudaArray_t tex_array = nullptr;  // layered copy for the gather path (optional)
    float4* d_quad = nullptr;         // quad copy for the 16-byte corner fetch (optional)
    float* d_oct = nullptr;           // oct copy for the 32-byte corner fetch (optional)
    float* d_poly = nullptr;          // poly copy for the tolerance-mode march (built at the first BIRP_FAST_MARCH render)
    bool poly_tried = false;
};

static_assert(birp::kTinyIngress == BIRP_TINY_INGRESS, "birp_host.h and birp_device.cuh disagree on 2^-100");
using birp::host_bounds_limit;
using birp::host_first_cell_limit;

namespace {

thread_local std::string g_error;

int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define BIRP_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return set_error(BIRP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                   \
    } while (0)

struct RayTable {
    float* d_dist = nullptr;
    float ray_width = 0.f;
    bool monotone = false;  // dist[] never decreases, or never increases (no NaN): RenderGeom::monotone_table
    double step = 0.0;      // maxdist / nsamples: the table is host_ray_table's uniform one (RenderGeom::fast_step)
};

// Per-device state.  Created lazily on first use from the thread that owns the device.
struct DeviceCtx {
    int device = -1;
    cudaStream_t compute = nullptr;
    cudaStream_t compute2 = nullptr;  // odd chunks of a host-buffer call: the next chunk's kernel fills the SMs the
                                      // previous one's last CTAs leave idle
    cudaStream_t copy = nullptr;
    cudaEvent_t frames_ready = nullptr;
    cudaEvent_t rendered[2] = {nullptr, nullptr};
    cudaEvent_t drained[2] = {nullptr, nullptr};
    std::map<std::pair<int, int>, RayTable> tables;
    birp_frame* d_frames = nullptr;
    size_t frames_cap = 0;
    float* d_images[2] = {nullptr, nullptr};
    size_t images_cap[2] = {0, 0};  // floats, per buffer (a failed grow of one leaves the other consistent)
    float* d_scratch = nullptr;
    size_t scratch_cap = 0;  // floats
    int* d_status = nullptr;
    int* h_status = nullptr;  // pinned
    int* d_flags = nullptr;   // 4 ints of scratch for the upload-time analysis kernels
    // spacing bits + limit bits -> cheapest exact div_spacing mode (0, 1, 2), proven on this device
    std::map<std::pair<unsigned, unsigned>, int> divmode_cache;
    // CMEM parameter record (raw bytes) -> may the region filter be used?  (self-test at first use, see cmem_flags)
    struct CmemState {
        bool filter_ok = false;       // the fp32 estimates stay within a quarter of their margin
        float4* d_table = nullptr;    // region table (cmem_table_kernel), or nullptr when it failed its self-test
        unsigned long long table_decided = 0, table_probes = 0;  // of the self-test's positions: how many the table decided
    };
    std::map<std::string, CmemState> cmem_state;
    int tile_w_log2 = -1;     // BIRP_TILE_W_LOG2 override (experiments)
    int cta_warps = 0;        // warps per CTA, 0 = by frame size (experiments: birp_set_tuning("cta_warps", 2 | 4 | 8))
    int tile_swizzle = 0;     // BIRP_TILE_SWIZZLE override (experiments)
    int fast_variant = -1;    // BIRP_FAST_VARIANT override (experiments)
    int fast_coop = 0;        // experiment: warp-per-ray tolerance-mode march
    int fast_kstep = 1;       // BIRP_FAST_KSTEP=0: the tolerance-mode march reads the ray table (experiments)
    std::mutex mu;
};

std::mutex g_ctx_mu;
std::map<int, DeviceCtx*> g_ctx;

int get_ctx(DeviceCtx** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return set_error(BIRP_ERR_CUDA, "no usable CUDA device: %s (this library has no CPU path)",
                         cudaGetErrorString(e));
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(dev);
    if (it != g_ctx.end()) {
        *out = it->second;
        return BIRP_OK;
    }
    DeviceCtx* c = new DeviceCtx();
    c->device = dev;
    BIRP_CUDA(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    BIRP_CUDA(cudaStreamCreateWithFlags(&c->compute2, cudaStreamNonBlocking));
    BIRP_CUDA(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    BIRP_CUDA(cudaEventCreateWithFlags(&c->frames_ready, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        BIRP_CUDA(cudaEventCreateWithFlags(&c->rendered[i], cudaEventDisableTiming));
        BIRP_CUDA(cudaEventCreateWithFlags(&c->drained[i], cudaEventDisableTiming));
    }
    BIRP_CUDA(cudaMalloc(&c->d_status, sizeof(int)));
    BIRP_CUDA(cudaMemset(c->d_status, 0, sizeof(int)));
    BIRP_CUDA(cudaHostAlloc(&c->h_status, 4 * sizeof(int), cudaHostAllocDefault));
    BIRP_CUDA(cudaMalloc(&c->d_flags, 4 * sizeof(int)));
    if (const char* t = getenv("BIRP_TILE_W_LOG2")) c->tile_w_log2 = atoi(t);
    if (const char* t = getenv("BIRP_TILE_SWIZZLE")) c->tile_swizzle = atoi(t);
    if (const char* t = getenv("BIRP_FAST_VARIANT")) c->fast_variant = atoi(t);
    if (const char* t = getenv("BIRP_FAST_KSTEP")) c->fast_kstep = atoi(t);
    if (const char* t = getenv("BIRP_CTA_WARPS")) c->cta_warps = atoi(t);
    g_ctx[dev] = c;
    *out = c;
    return BIRP_OK;
}

int get_ray_table(DeviceCtx* c, int nsamples, int maxdist, cudaStream_t stream, RayTable* out) {
    auto key = std::make_pair(nsamples, maxdist);
    auto it = c->tables.find(key);
    if (it != c->tables.end()) {
        *out = it->second;
        return BIRP_OK;
    }
    std::vector<float> h(nsamples);
    RayTable t;
    birp::host_ray_table(nsamples, maxdist, h.data(), &t.ray_width);
    bool up = true, down = true;
    for (int i = 0; i < nsamples; ++i) {
        if (h[i] != h[i]) up = down = false;
        if (i > 0 && !(h[i] >= h[i - 1])) up = false;
        if (i > 0 && !(h[i] <= h[i - 1])) down = false;
    }
    t.monotone = up || down;
    t.step = (maxdist > 0 && nsamples < (1 << 22)) ? (double)maxdist / (double)nsamples : 0.0;
    BIRP_CUDA(cudaMalloc(&t.d_dist, sizeof(float) * nsamples));
    // synchronous copy from pageable memory: the table is ready for any stream afterwards
    BIRP_CUDA(cudaMemcpy(t.d_dist, h.data(), sizeof(float) * nsamples, cudaMemcpyHostToDevice));
    (void)stream;
    c->tables[key] = t;
    *out = t;
    return BIRP_OK;
}

template <class T>
int ensure(T** p, size_t* cap, size_t need) {
    if (*cap >= need) return BIRP_OK;
    if (*p) BIRP_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    BIRP_CUDA(cudaMalloc(p, need * sizeof(T)));
    *cap = need;
    return BIRP_OK;
}

int check_geom(const void* frames, int nframes, int W, int H, int* row_begin, int* row_end, int nsamples,
               const void* images) {
    if (!frames || !images) return set_error(BIRP_ERR_ARG, "null frames/images pointer");
    if (nframes < 0 || W <= 0 || H <= 0 || nsamples <= 0)
        return set_error(BIRP_ERR_ARG, "bad sizes: nframes=%d W=%d H=%d nsamples=%d", nframes, W, H, nsamples);
    if (*row_end <= 0 || *row_end > H) *row_end = H;
    if (*row_begin < 0) *row_begin = 0;
    if (*row_begin > *row_end) return set_error(BIRP_ERR_ARG, "row_begin > row_end");
    return BIRP_OK;
}

int read_status(DeviceCtx* c, cudaStream_t stream) {
    BIRP_CUDA(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    BIRP_CUDA(cudaStreamSynchronize(stream));
    const int st = *c->h_status;
    if (st) {
        BIRP_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int), stream));
        BIRP_CUDA(cudaStreamSynchronize(stream));
        if (st & birp::kStatusPeerTimeout)
            return set_error(BIRP_ERR_CUDA, "birp_peer_wait: a rank did not signal within the time limit");
        if (st & birp::kStatusSingular) return set_error(BIRP_ERR_SINGULAR, "Determinant is 0");
        if (st & birp::kStatusNanIndex)
            return set_error(BIRP_ERR_RANGE,
                             "NaN sample coordinate inside the cube bounds test "
                             "(the reference throws std::out_of_range here)");
    }
    return BIRP_OK;
}

int run_cmem_filter_selftest(DeviceCtx* c, const birp_cmem_params& params, long n, double out[3]) {
    int rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)16);
    if (rc) return rc;
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(c->d_scratch);
    BIRP_CUDA(cudaMemsetAsync(d_out, 0, 3 * sizeof(unsigned long long), c->compute));
    cudaError_t e = birp::launch_cmem_filter_selftest(params, n, d_out, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    unsigned long long h[3];
    BIRP_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    std::memcpy(&out[0], &h[0], sizeof(double));
    std::memcpy(&out[1], &h[1], sizeof(double));
    out[2] = (double)h[2] / (double)n;
    return BIRP_OK;
}

// The region filter of the CMEM kernel is exact as long as its fp32 radius estimates stay within their margin.
// That is measured over 10^8 directions for the parameter sets of the tests; for ANY parameter set the same
// self-test runs here once, at first use (4M directions, a few ms, cached per device): if the estimates use
// more than a quarter of the margin the filter is switched off for that set (BIRP_CMEM_NO_FILTER, same results,
// slower) -- the CMEM counterpart of the division / texture / quad self-checks at cube upload.
int cmem_flags(DeviceCtx* c, const birp_cmem_params& params, int flags, int* out_flags, const float4** d_table = nullptr) {
    *out_flags = flags;
    if (d_table) *d_table = nullptr;
    if (flags & (BIRP_CMEM_LIBM32 | BIRP_CMEM_LIBM64 | BIRP_CMEM_NO_FILTER)) return BIRP_OK;
    const std::string key(reinterpret_cast<const char*>(&params), sizeof(params));
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->cmem_state.find(key);
    if (it == c->cmem_state.end()) {
        DeviceCtx::CmemState st;
        double u[3];
        int rc = run_cmem_filter_selftest(c, params, 1L << 22, u);
        if (rc) return rc;
        st.filter_ok = u[0] <= 0.25 && u[1] <= 0.25;  // (NaNs compare false: no filter)
        // The region table of this parameter set: built once (256 x 64 cells x 169 exact evaluations, a few ms), then
        // the march with it is compared with the unfiltered evaluation on 4M pseudo-random positions; a single
        // differing sample and the table is dropped for this set.
        if (st.filter_ok && !getenv("BIRP_DISABLE_CMEM_TABLE")) {
            cudaError_t e = cudaMalloc(&st.d_table, sizeof(float4) * birp::kCmemTableT * birp::kCmemTableP);
            if (e == cudaSuccess) e = birp::launch_cmem_table(params, st.d_table, c->compute);
            unsigned long long h[2] = {1, 0};
            const long n = 1L << 22;
            if (e == cudaSuccess) {
                rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)16);
                if (rc) return rc;
                unsigned long long* d_out = reinterpret_cast<unsigned long long*>(c->d_scratch);
                e = cudaMemsetAsync(d_out, 0, 2 * sizeof(unsigned long long), c->compute);
                if (e == cudaSuccess) e = birp::launch_cmem_table_selftest(params, st.d_table, n, d_out, c->compute);
                if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->compute);
                if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
            }
            if (e != cudaSuccess || h[0] != 0) {
                cudaGetLastError();
                if (st.d_table) cudaFree(st.d_table);
                st.d_table = nullptr;
            } else {
                st.table_decided = h[1];
                st.table_probes = (unsigned long long)n;
            }
        }
        it = c->cmem_state.emplace(key, st).first;
    }
    if (!it->second.filter_ok) *out_flags |= (flags & BIRP_CMEM_FAST) ? BIRP_CMEM_LIBM32 : BIRP_CMEM_NO_FILTER;
    if (!it->second.filter_ok) *out_flags &= ~BIRP_CMEM_FAST;
    if (d_table && !(flags & BIRP_CMEM_NO_TABLE)) *d_table = it->second.d_table;
    return BIRP_OK;
}

// Shared host-buffer pipeline: frames are rendered in chunks of up to ~64 MiB of pixels,
// chunk c rendering on the compute stream while chunk c-1 drains to the host on the copy
// stream (two device image buffers).
// Poses of a sweep whose frame records are still to be made (birp_render_*_sweep): the records of chunk c are
// computed on the host threads while the GPU renders chunk c-1.
struct SweepPoses {
    const float* poses = nullptr;  // xyz interleaved, nframes of them; nullptr: `frames` is already filled
    float aim_x = 0.f, fov_x_deg = 0.f, fov_y_deg = 0.f;
    birp_frame* frames_out = nullptr;  // host buffer the records are written to (caller's or a scratch vector)
};

// Where the frames of a host-buffer call land and who hands them out.
//   slot of frame f in the caller's image buffer = out_offset + f * out_stride (1-GPU calls: 0 / 1; a sweep whose
//   frames are interleaved over several processes writes straight into the shared result with stride = #processes);
//   feed: the next unrendered frame of the sweep, shared by the device threads of a multi-GPU sweep (each takes a
//   chunk at a time: dynamic load balance, the frames of an orbit do not cost the same); nullptr = this call renders
//   every frame itself;
//   host_threads: threads for the per-chunk frame setup (0 = choose).
struct HostPlan {
    long long out_offset = 0, out_stride = 1;
    // interleaved row bands of every frame (RenderGeom::row_band ...): only those bands are rendered and copied
    int row_band = 0, row_phase = 0, row_nphases = 1;
    std::atomic<long long>* feed = nullptr;
    size_t feed_chunk = 0;
    int host_threads = 0;
};

template <class Launch>
int render_host(DeviceCtx* c, const birp_frame* frames, int nframes, int W, int H, int row_begin, int row_end,
                int nsamples, int maxdist, int flags, float* images, Launch&& launch,
                const SweepPoses* sweep = nullptr, const HostPlan* plan_in = nullptr) {
    std::lock_guard<std::mutex> lk(c->mu);
    const HostPlan plan = plan_in ? *plan_in : HostPlan();
    RayTable table;
    int rc = get_ray_table(c, nsamples, maxdist, c->compute, &table);
    if (rc) return rc;
    if (nframes == 0 || row_begin == row_end) return BIRP_OK;

    const size_t frame_px = (size_t)W * H;
    const size_t band_px = (size_t)W * (row_end - row_begin);
    const size_t band_off = (size_t)W * row_begin;
    size_t chunk = (size_t(16) << 20) / frame_px;  // 16 Mi pixels = 64 MiB per buffer
    // at least four chunks for a batch of many small frames, so that the D2H copy of one chunk (and the host
    // set-up of the next, for a sweep) overlaps the render of another
    // (four: eight pieces of a 10 000-frame sweep were measured slower, 283.9 k against 288.8 k frames/s end to end)
    if ((size_t)nframes >= 256 && chunk > ((size_t)nframes + 3) / 4) chunk = ((size_t)nframes + 3) / 4;
    if (plan.feed && plan.feed_chunk) chunk = plan.feed_chunk < chunk ? plan.feed_chunk : chunk;
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)nframes) chunk = nframes;
    // a batch of many frames: a small FIRST chunk, so that the GPU starts after the host set-up / H2D of a few frames
    // instead of a quarter of the batch, and a small LAST one, whose D2H is the piece nothing overlaps -- at 8 GPUs the
    // last chunks of all ranks drain at the same time through the box's ~94 GB/s of aggregate D2H bandwidth
    // (tools/d2h_aggregate.py), so a quarter of the sweep there was 0.7 ms of a 3.5 ms step
    const size_t first_chunk = (!plan.feed && chunk >= 128) ? (chunk / 8 > 32 ? chunk / 8 : 32) : chunk;
    rc = ensure(&c->d_frames, &c->frames_cap, (size_t)nframes);
    if (rc) return rc;
    for (int b = 0; b < 2; ++b) {
        rc = ensure(&c->d_images[b], &c->images_cap[b], chunk * frame_px);
        if (rc) return rc;
    }
    // From here on work is enqueued on two streams that read and write the caller's buffers: whatever way this
    // function is left -- a singular pose in a later chunk, a CUDA error -- nothing may still be in flight when it
    // returns ("synchronous on return"), and a pending device status must not leak into the next call.
    struct DrainOnExit {
        DeviceCtx* c;
        ~DrainOnExit() {
            cudaStreamSynchronize(c->compute);
            cudaStreamSynchronize(c->compute2);
            cudaStreamSynchronize(c->copy);
        }
    } drain{c};
    if (!sweep && !plan.feed) {
        BIRP_CUDA(cudaMemcpyAsync(c->d_frames, frames, sizeof(birp_frame) * nframes, cudaMemcpyHostToDevice,
                                  c->compute));
        BIRP_CUDA(cudaEventRecord(c->frames_ready, c->compute));
        BIRP_CUDA(cudaStreamWaitEvent(c->compute2, c->frames_ready, 0));
    }
    const bool full_rows = (row_begin == 0 && row_end == H);
    const size_t frame_bytes = sizeof(float) * frame_px;
    // interleaved row bands: per frame, the own bands are `own_full` whole bands at a pitch of row_nphases bands
    // (one 2-D copy) plus, when the image's last band is short and ours, that piece
    const bool interleaved = plan.row_band > 0 && plan.row_nphases > 1;
    auto copy_interleaved = [&](bool to_device, float* d_img, float* h_first, size_t h_pitch, int nf, cudaStream_t st) -> int {
        const int rows_total = row_end - row_begin;
        const int total_bands = (rows_total + plan.row_band - 1) / plan.row_band;
        const int last_rows = rows_total - (total_bands - 1) * plan.row_band;  // rows of the image's last band
        int own = (total_bands > plan.row_phase) ? (total_bands - plan.row_phase + plan.row_nphases - 1) / plan.row_nphases : 0;
        const bool own_last = own > 0 && ((total_bands - 1) % plan.row_nphases) == plan.row_phase;
        int own_full = own;
        if (own_last && last_rows < plan.row_band) --own_full;
        const size_t band_bytes = sizeof(float) * (size_t)plan.row_band * W;
        const size_t first = (size_t)W * ((size_t)row_begin + (size_t)plan.row_phase * plan.row_band);
        for (int f = 0; f < nf; ++f) {
            float* d = d_img + (size_t)f * frame_px;
            float* h = reinterpret_cast<float*>(reinterpret_cast<char*>(h_first) + (size_t)f * h_pitch);
            if (own_full > 0) {
                BIRP_CUDA(cudaMemcpy2DAsync(to_device ? d + first : h + first, band_bytes * plan.row_nphases,
                                            to_device ? h + first : d + first, band_bytes * plan.row_nphases, band_bytes,
                                            (size_t)own_full, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
            }
            if (own_full < own) {
                const size_t off = (size_t)W * ((size_t)row_begin + (size_t)(total_bands - 1) * plan.row_band);
                BIRP_CUDA(cudaMemcpyAsync(to_device ? d + off : h + off, to_device ? h + off : d + off,
                                          sizeof(float) * (size_t)last_rows * W,
                                          to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
            }
        }
        return BIRP_OK;
    };
    int nchunks = 0;
    size_t cursor = 0;  // next frame of a call that renders everything itself
    for (;; ++nchunks) {
        size_t f0;
        int nf;
        if (plan.feed) {
            const long long got = plan.feed->fetch_add((long long)chunk);
            if (got >= (long long)nframes) break;
            f0 = (size_t)got;
            nf = (int)((f0 + chunk <= (size_t)nframes) ? chunk : (size_t)nframes - f0);
        } else {
            if (cursor >= (size_t)nframes) break;
            f0 = cursor;
            // (and a small LAST chunk: its D2H is the other exposed piece)
            const size_t rem = (size_t)nframes - f0;
            size_t want = (nchunks == 0) ? first_chunk : chunk;
            if (first_chunk < chunk && nchunks > 0 && rem > first_chunk && rem <= chunk + first_chunk) want = rem - first_chunk;
            nf = (int)(want <= rem ? want : rem);
            cursor += (size_t)nf;
        }
        const int b = nchunks & 1;
        float* d_img = c->d_images[b];
        cudaStream_t cs = b ? c->compute2 : c->compute;  // (chunks alternate between two compute streams)
        if (sweep) {
            // frame records of this chunk: host threads now (the GPU is busy with the previous chunk), then H2D
            int src = birp::host_frame_setup_batch(sweep->poses + 3 * f0, nf, sweep->aim_x, sweep->fov_x_deg,
                                                   sweep->fov_y_deg, W, H, sweep->frames_out + f0, plan.host_threads);
            if (src) {
                cudaStreamSynchronize(c->compute);
                cudaStreamSynchronize(c->compute2);
                read_status(c, c->compute);  // (clears a status word set by the chunks already rendered)
                if (src == BIRP_ERR_SINGULAR) return set_error(src, "Determinant is 0");
                return src;
            }
            BIRP_CUDA(cudaMemcpyAsync(c->d_frames + f0, sweep->frames_out + f0, sizeof(birp_frame) * nf,
                                      cudaMemcpyHostToDevice, cs));
        } else if (plan.feed) {
            BIRP_CUDA(cudaMemcpyAsync(c->d_frames + f0, frames + f0, sizeof(birp_frame) * nf, cudaMemcpyHostToDevice,
                                      cs));
        }
        // caller's image slot of the chunk's first frame, and the distance between consecutive frames there
        float* h_first = images + (size_t)(plan.out_offset + (long long)f0 * plan.out_stride) * frame_px;
        const size_t h_pitch = (size_t)plan.out_stride * frame_bytes;
        if (nchunks >= 2) BIRP_CUDA(cudaStreamWaitEvent(cs, c->drained[b], 0));
        if (flags & BIRP_ACCUMULATE) {
            if (interleaved) {
                rc = copy_interleaved(true, d_img, h_first, h_pitch, nf, cs);
                if (rc) return rc;
            } else if (full_rows) {
                BIRP_CUDA(cudaMemcpy2DAsync(d_img, frame_bytes, h_first, h_pitch, frame_bytes, (size_t)nf,
                                            cudaMemcpyHostToDevice, cs));
            } else {
                BIRP_CUDA(cudaMemcpy2DAsync(d_img + band_off, frame_bytes, h_first + band_off, h_pitch,
                                            sizeof(float) * band_px, (size_t)nf, cudaMemcpyHostToDevice, cs));
            }
        }
        // A chunk of one or two LARGE frames (a 4096 x 4096 image is 64 MiB) is rendered and drained in row pieces:
        // the copy of piece p overlaps the render of piece p + 1 instead of following the whole frame.
        int npieces = 1;
        if (!interleaved && nf <= 2 && band_px >= (size_t(4) << 20) && (row_end - row_begin) >= 64) npieces = 4;
        for (int piece = 0; piece < npieces; ++piece) {
            // piece boundaries on multiples of 8 rows (the warp's pixel tile is 4 x 8 or 8 x 4)
            const int rows_all = row_end - row_begin;
            const int r0 = row_begin + (int)(((long long)rows_all * piece / npieces) / 8 * 8);
            const int r1 = (piece + 1 == npieces) ? row_end : row_begin + (int)(((long long)rows_all * (piece + 1) / npieces) / 8 * 8);
            if (r1 <= r0) continue;
            birp::RenderGeom g{nf, W, H, r0, r1, nsamples, table.ray_width};
            g.monotone_table = table.monotone;
            g.tile_w_log2 = c->tile_w_log2;
            g.fast_variant = c->fast_variant;
            g.tile_swizzle = c->tile_swizzle;
            g.cta_warps = c->cta_warps;
            g.fast_kstep = c->fast_kstep;
            g.fast_coop = c->fast_coop;
            g.fast_step = table.step;
            g.row_band = plan.row_band, g.row_phase = plan.row_phase, g.row_nphases = plan.row_nphases;
            cudaError_t e = launch(c->d_frames + f0, g, table.d_dist, d_img, cs);
            if (e != cudaSuccess)
                return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
            BIRP_CUDA(cudaEventRecord(c->rendered[b], cs));
            BIRP_CUDA(cudaStreamWaitEvent(c->copy, c->rendered[b], 0));
            const size_t p_off = (size_t)W * r0, p_px = (size_t)W * (r1 - r0);
            if (interleaved) {
                rc = copy_interleaved(false, d_img, h_first, h_pitch, nf, c->copy);
                if (rc) return rc;
            } else if (r0 == 0 && r1 == H) {
                if (plan.out_stride == 1) {
                    BIRP_CUDA(cudaMemcpyAsync(h_first, d_img, frame_bytes * nf, cudaMemcpyDeviceToHost, c->copy));
                } else {
                    BIRP_CUDA(cudaMemcpy2DAsync(h_first, h_pitch, d_img, frame_bytes, frame_bytes, (size_t)nf,
                                                cudaMemcpyDeviceToHost, c->copy));
                }
            } else {
                // the rows of this piece / band in every frame: rows of p_px floats, one per frame
                BIRP_CUDA(cudaMemcpy2DAsync(h_first + p_off, h_pitch, d_img + p_off, frame_bytes, sizeof(float) * p_px,
                                            (size_t)nf, cudaMemcpyDeviceToHost, c->copy));
            }
        }
        BIRP_CUDA(cudaEventRecord(c->drained[b], c->copy));
    }
    BIRP_CUDA(cudaStreamSynchronize(c->copy));  // (the copy stream waited for every chunk's kernel)
    BIRP_CUDA(cudaStreamSynchronize(c->compute2));
    return read_status(c, c->compute);
}

}  // namespace

extern "C" {

const char* birp_version(void) { return "birp_b200 0.1 (sm_100a)"; }
const char* birp_last_error(void) { return g_error.c_str(); }

int birp_device_count(int* count) {
    if (!count) return set_error(BIRP_ERR_ARG, "null count");
    *count = 0;
    BIRP_CUDA(cudaGetDeviceCount(count));
    return BIRP_OK;
}

int birp_set_device(int ordinal) {
    BIRP_CUDA(cudaSetDevice(ordinal));
    return BIRP_OK;
}

int birp_host_alloc(size_t bytes, void** out) {
    if (!out) return set_error(BIRP_ERR_ARG, "null out");
    BIRP_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return BIRP_OK;
}

int birp_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return set_error(BIRP_ERR_ARG, "null buffer");
    BIRP_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return BIRP_OK;
}

int birp_host_unregister(void* p) {
    if (p) BIRP_CUDA(cudaHostUnregister(p));
    return BIRP_OK;
}

int birp_host_free(void* p) {
    if (p) BIRP_CUDA(cudaFreeHost(p));
    return BIRP_OK;
}

int birp_frame_setup(const float pos[3], float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                     birp_frame* out) {
    if (!pos || !out || W <= 0 || H <= 0) return set_error(BIRP_ERR_ARG, "bad frame_setup arguments");
    int rc = birp::host_frame_setup(pos, aim_x, fov_x_deg, fov_y_deg, W, H, out);
    if (rc == BIRP_ERR_SINGULAR) return set_error(rc, "Determinant is 0");
    return rc;
}

int birp_frame_setup_aim(const float pos[3], const float aim[3], float fov_x_deg, float fov_y_deg, int W, int H,
                         birp_frame* out) {
    if (!pos || !aim || !out || W <= 0 || H <= 0) return set_error(BIRP_ERR_ARG, "bad frame_setup arguments");
    int rc = birp::host_frame_setup_aim(pos, aim, fov_x_deg, fov_y_deg, W, H, out);
    if (rc == BIRP_ERR_SINGULAR) return set_error(rc, "Determinant is 0");
    return rc;
}

int birp_frame_setup_batch(const float* poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                           birp_frame* out, int nthreads) {
    if (n < 0 || (n > 0 && (!poses || !out)) || W <= 0 || H <= 0)
        return set_error(BIRP_ERR_ARG, "bad frame_setup_batch arguments");
    int rc = birp::host_frame_setup_batch(poses, n, aim_x, fov_x_deg, fov_y_deg, W, H, out, nthreads);
    if (rc == BIRP_ERR_SINGULAR) return set_error(rc, "Determinant is 0");
    return rc;
}

int birp_frame_setup_device(const float* d_poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                            birp_frame* d_frames, void* stream) {
    if (n < 0 || (n > 0 && (!d_poses || !d_frames)) || W <= 0 || H <= 0)
        return set_error(BIRP_ERR_ARG, "bad frame_setup_device arguments");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    float tx, ty, aspect;
    birp::host_fov_terms(fov_x_deg, fov_y_deg, W, H, &tx, &ty, &aspect);  // pose-independent: host tanf, exact
    cudaError_t e = birp::launch_frame_setup(d_poses, n, aim_x, tx, ty, aspect, d_frames, c->d_status,
                                             (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return BIRP_OK;
}

static int orbit_constants(int n, double perigee, double apogee, double inclination_deg, double k[9]) {
    if (n < 0 || !(perigee > 0.0) || !(apogee >= perigee)) return set_error(BIRP_ERR_ARG, "bad orbit arguments");
    const double pe = perigee, ap = apogee;
    k[0] = 0.5 * (pe + ap);
    k[1] = (ap - pe) / (ap + pe);
    k[2] = std::sqrt(1.0 - k[1] * k[1]);
    const double inc = inclination_deg * (M_PI / 180.0);
    k[3] = std::cos(inc), k[4] = 0.0, k[5] = std::sin(inc);                             // apogee direction
    k[6] = -0.1 * std::sin(inc), k[7] = 0.99498743710662, k[8] = 0.1 * std::cos(inc);  // in-plane, orthogonal
    return BIRP_OK;
}

int birp_orbit_poses(int n, double perigee, double apogee, double inclination_deg, float* poses) {
    double k[9];
    int rc = orbit_constants(n, perigee, apogee, inclination_deg, k);
    if (rc) return rc;
    if (n > 0 && !poses) return set_error(BIRP_ERR_ARG, "null poses");
    for (int i = 0; i < n; ++i) {
        const double M = 2.0 * M_PI * ((double)i + 0.5) / (double)n;
        double E = M;
        for (int it = 0; it < 60; ++it) E -= (E - k[1] * std::sin(E) - M) / (1.0 - k[1] * std::cos(E));
        const double xp = k[0] * (std::cos(E) - k[1]);
        const double yp = k[0] * k[2] * std::sin(E);
        for (int a = 0; a < 3; ++a) poses[3 * i + a] = (float)((-xp) * k[3 + a] + yp * k[6 + a]);
    }
    return BIRP_OK;
}

int birp_orbit_poses_device(int n, double perigee, double apogee, double inclination_deg, float* d_poses, void* stream) {
    double k[9];
    int rc = orbit_constants(n, perigee, apogee, inclination_deg, k);
    if (rc) return rc;
    if (n > 0 && !d_poses) return set_error(BIRP_ERR_ARG, "null poses");
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    cudaError_t e = birp::launch_orbit_poses(n, k, d_poses, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return BIRP_OK;
}

int birp_ray_table(int nsamples, int maxdist, float* dist_out, float* ray_width_out) {
    if (nsamples <= 0) return set_error(BIRP_ERR_ARG, "nsamples must be positive");
    birp::host_ray_table(nsamples, maxdist, dist_out, ray_width_out);
    return BIRP_OK;
}

int birp_cmem_init(int has_v, const int v[3], int has_b, const int b[3], float dipole, float p0, int p1,
                   int p2, int p3, float B, float alpha, float beta, float bs, float A1, float A2,
                   float ay_bs, float az_bs, float density, birp_cmem_params* out) {
    if (!out) return set_error(BIRP_ERR_ARG, "null out");
    if ((has_v && !v) || (has_b && !b)) return set_error(BIRP_ERR_ARG, "null v/b");
    birp::host_cmem_init(has_v, v, has_b, b, dipole, p0, p1, p2, p3, B, alpha, beta, bs, A1, A2, ay_bs, az_bs,
                         density, out);
    return BIRP_OK;
}

static void destroy_cube(birp_cube* cube);

// Upload-time analysis of a resident cube: reciprocal spacings, bounds limits, the cheapest
// division variant that is EXACT for this cube (exhaustive device check, cached per spacing),
// and whether all voxels are finite and positive.
static int finish_cube(DeviceCtx* c, birp_cube* cube) {
    birp::CubeView& v = cube->view;
    const float sp[3] = {v.sx, v.sy, v.sz};
    const int n[3] = {v.nx, v.ny, v.nz};
    float rcp[3], lim[3];
    int mode = 0;
    std::lock_guard<std::mutex> lk(c->mu);
    for (int a = 0; a < 3; ++a) {
        const volatile float r = 1.0f / sp[a];
        rcp[a] = r;
        lim[a] = host_bounds_limit(sp[a], n[a]);
        unsigned sbits, lbits;
        std::memcpy(&sbits, &sp[a], 4);
        std::memcpy(&lbits, &lim[a], 4);
        auto key = std::make_pair(sbits, lbits);
        auto it = c->divmode_cache.find(key);
        int m;
        if (it != c->divmode_cache.end()) {
            m = it->second;
        } else {
            BIRP_CUDA(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->compute));
            cudaError_t e = birp::launch_verify_div(sp[a], rcp[a], lbits, c->d_flags, c->compute);
            if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "verify_div launch: %s", cudaGetErrorString(e));
            BIRP_CUDA(cudaMemcpyAsync(c->h_status, c->d_flags, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
            BIRP_CUDA(cudaStreamSynchronize(c->compute));
            // cost rank of the cheapest exact form: 0 the multiply alone, 1 one correction step, 2 two, 3 IEEE division
            m = (c->h_status[2] == 0) ? 0 : (c->h_status[0] == 0) ? 1 : (c->h_status[1] == 0) ? 2 : 3;
            c->divmode_cache[key] = m;
        }
        if (m > mode) mode = m;
    }
    BIRP_CUDA(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->compute));
    BIRP_CUDA(cudaMemsetAsync(c->d_flags + 1, 0xff, sizeof(int), c->compute));  // running minimum of the bit patterns
    cudaError_t e = birp::launch_scan_positive(cube->d_data, cube->voxels, c->d_flags, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "scan_positive launch: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(c->h_status, c->d_flags, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    v.positive = (c->h_status[0] == 0) ? 1 : ((c->h_status[0] & 2) == 0) ? 2 : 0;
    // value range of a positive cube (every interpolated sample lies inside it): what the lean march relies on
    v.vmin = v.vmax = 0.f;
    if (v.positive == 1) {
        std::memcpy(&v.vmin, &c->h_status[1], 4);
        std::memcpy(&v.vmax, &c->h_status[2], 4);
    }
    c->h_status[0] = c->h_status[1] = c->h_status[2] = 0;

    // Layered-texture copy for the two-gather corner fetch (trilinear only).  Any failure here
    // just leaves the global-load path in place.
    v.tex = 0;
    if (v.nx >= 2 && v.ny >= 2 && v.nz >= 2 && v.nx <= 32768 && v.ny <= 32768 && v.nz <= 2048 &&
        !getenv("BIRP_DISABLE_TEXTURE")) {
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        cudaError_t te = cudaMalloc3DArray(&cube->tex_array, &desc, make_cudaExtent(v.nx, v.ny, v.nz), cudaArrayLayered);
        if (te == cudaSuccess) {
            cudaMemcpy3DParms cp = {};
            cp.srcPtr = make_cudaPitchedPtr(cube->d_data, (size_t)v.nx * sizeof(float), v.nx, v.ny);
            cp.dstArray = cube->tex_array;
            cp.extent = make_cudaExtent(v.nx, v.ny, v.nz);
            cp.kind = cudaMemcpyDeviceToDevice;
            te = cudaMemcpy3DAsync(&cp, c->compute);
        }
        cudaTextureObject_t tex = 0;
        if (te == cudaSuccess) {
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = cube->tex_array;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            te = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        }
        if (te == cudaSuccess) {
            v.tex = tex;
            cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->compute);
            te = birp::launch_verify_gather(v, 65536, c->d_flags, c->compute);
            if (te == cudaSuccess)
                te = cudaMemcpyAsync(c->h_status, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->compute);
            if (te == cudaSuccess) te = cudaStreamSynchronize(c->compute);
            if (te != cudaSuccess || c->h_status[0] != 0) {
                cudaDestroyTextureObject(tex);
                v.tex = 0;
            }
            *c->h_status = 0;
        }
        if (v.tex == 0) {
            cudaGetLastError();  // clear a sticky-free error from the optional path
            if (cube->tex_array) cudaFreeArray(cube->tex_array);
            cube->tex_array = nullptr;
        }
    }
    // Quad copy for the 16-byte corner fetch (trilinear only): 4x the grid.  Optional like the
    // texture: when memory is short or the self-check fails the kernels use the other fetch forms.
    v.quad = nullptr;
    if (v.nx >= 2 && v.ny >= 2 && v.nz >= 2 && !getenv("BIRP_DISABLE_QUADS")) {
        const size_t bytes = cube->voxels * sizeof(float4);
        size_t free_b = 0, total_b = 0;
        cudaError_t qe = cudaMemGetInfo(&free_b, &total_b);
        if (qe == cudaSuccess && free_b / 2 >= bytes) qe = cudaMalloc(&cube->d_quad, bytes);
        else if (qe == cudaSuccess) qe = cudaErrorMemoryAllocation;
        if (qe == cudaSuccess) {
            cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->compute);
            qe = birp::launch_pack_quads(v, cube->d_quad, c->d_flags, c->compute);
            if (qe == cudaSuccess)
                qe = cudaMemcpyAsync(c->h_status, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->compute);
            if (qe == cudaSuccess) qe = cudaStreamSynchronize(c->compute);
            if (qe == cudaSuccess && c->h_status[0] == 0) v.quad = cube->d_quad;
            *c->h_status = 0;
        }
        if (!v.quad) {
            cudaGetLastError();
            if (cube->d_quad) cudaFree(cube->d_quad);
            cube->d_quad = nullptr;
        }
    }
    // Oct copy for the 32-byte corner fetch (the default of the trilinear kernel): 8x the grid.  Optional like
    // the quads: when memory is short or the self-check fails the kernels use the quad copy.
    v.oct = nullptr;
    v.oct_marked = 0;
    if (v.nx >= 2 && v.ny >= 2 && v.nz >= 2 && !getenv("BIRP_DISABLE_OCTS")) {
        const size_t bytes = cube->voxels * 8 * sizeof(float);
        size_t free_b = 0, total_b = 0;
        cudaError_t oe = cudaMemGetInfo(&free_b, &total_b);
        if (oe == cudaSuccess && free_b / 2 >= bytes) oe = cudaMalloc(&cube->d_oct, bytes);
        else if (oe == cudaSuccess) oe = cudaErrorMemoryAllocation;
        if (oe == cudaSuccess) {
            cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->compute);
            v.oct_marked = (v.positive == 1) ? 0 : 1;
            oe = birp::launch_pack_octs(v, cube->d_oct, c->d_flags, v.oct_marked, c->compute);
            if (oe == cudaSuccess)
                oe = cudaMemcpyAsync(c->h_status, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->compute);
            if (oe == cudaSuccess) oe = cudaStreamSynchronize(c->compute);
            if (oe == cudaSuccess && c->h_status[0] == 0) v.oct = cube->d_oct;
            *c->h_status = 0;
        }
        if (!v.oct) {
            cudaGetLastError();
            if (cube->d_oct) cudaFree(cube->d_oct);
            cube->d_oct = nullptr;
        }
    }
    v.lox = host_first_cell_limit(sp[0]);
    v.loy = host_first_cell_limit(sp[1]);
    v.loz = host_first_cell_limit(sp[2]);
    v.rx = rcp[0];
    v.ry = rcp[1];
    v.rz = rcp[2];
    v.lx = lim[0];
    v.ly = lim[1];
    v.lz = lim[2];
    // (cost rank -> div_spacing<MODE>: 3 multiply only, 0 one step, 1 two steps, 2 IEEE)
    static const int kModeOfRank[4] = {3, 0, 1, 2};
    v.divmode = kModeOfRank[mode];
    return BIRP_OK;
}

static int make_cube(const float* src, bool src_on_device, int nx, int ny, int nz, const float origin[3],
                     const float spacing[3], cudaStream_t stream, birp_cube** out) {
    if (!src || !origin || !spacing || !out) return set_error(BIRP_ERR_ARG, "null pointer");
    if (nx <= 0 || ny <= 0 || nz <= 0) return set_error(BIRP_ERR_ARG, "bad cube size %dx%dx%d", nx, ny, nz);
    for (int a = 0; a < 3; ++a)
        if (!(spacing[a] > 0.f) || std::isinf(spacing[a]) || std::isnan(origin[a]) || std::isinf(origin[a]))
            return set_error(BIRP_ERR_ARG, "cube spacing must be positive and finite, origin finite");
    const size_t voxels = (size_t)nx * ny * nz;
    if (voxels >= (size_t(1) << 31)) return set_error(BIRP_ERR_ARG, "cube has >= 2^31 voxels");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    birp_cube* cube = new birp_cube();
    cube->device = c->device;
    cube->voxels = voxels;
    cudaError_t e = cudaMalloc(&cube->d_data, voxels * sizeof(float));
    if (e != cudaSuccess) {
        delete cube;
        return set_error(e == cudaErrorMemoryAllocation ? BIRP_ERR_NOMEM : BIRP_ERR_CUDA,
                         "cudaMalloc of %zu-voxel cube failed: %s", voxels, cudaGetErrorString(e));
    }
    if (src_on_device) {
        e = cudaMemcpyAsync(cube->d_data, src, voxels * sizeof(float), cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    } else {
        e = cudaMemcpy(cube->d_data, src, voxels * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(cube->d_data);
        delete cube;
        return set_error(BIRP_ERR_CUDA, "cube copy failed: %s", cudaGetErrorString(e));
    }
    birp::CubeView& v = cube->view;
    v.v = cube->d_data;
    v.nx = nx;
    v.ny = ny;
    v.nz = nz;
    v.slab = nx * ny;
    v.slab16 = 16LL * nx * ny;
    v.fnx = (float)nx;
    v.fny = (float)ny;
    v.fnz = (float)nz;
    v.ox = origin[0];
    v.oy = origin[1];
    v.oz = origin[2];
    v.sx = spacing[0];
    v.sy = spacing[1];
    v.sz = spacing[2];
    v.rx64 = 1.0 / (double)spacing[0];
    v.ry64 = 1.0 / (double)spacing[1];
    v.rz64 = 1.0 / (double)spacing[2];
    v.poly = nullptr;
    rc = finish_cube(c, cube);
    if (rc) {
        destroy_cube(cube);
        return rc;
    }
    *out = cube;
    return BIRP_OK;
}

int birp_cube_upload(const float* zyx, int nx, int ny, int nz, const float origin[3], const float spacing[3],
                     birp_cube** out) {
    return make_cube(zyx, false, nx, ny, nz, origin, spacing, nullptr, out);
}

int birp_cube_upload_device(const float* d_zyx, int nx, int ny, int nz, const float origin[3],
                            const float spacing[3], void* stream, birp_cube** out) {
    return make_cube(d_zyx, true, nx, ny, nz, origin, spacing, (cudaStream_t)stream, out);
}

// A resident cube copied to another GPU of the box: the linear grid travels device to device (cudaMemcpyPeer over
// NVLink), the upload-time analysis and the copies for the corner fetch are rebuilt there.
int birp_cube_replicate(const birp_cube* src, int device, birp_cube** out) {
    if (!src || !out) return set_error(BIRP_ERR_ARG, "null pointer");
    int prev = 0;
    BIRP_CUDA(cudaGetDevice(&prev));
    int ndev = 0;
    BIRP_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return set_error(BIRP_ERR_ARG, "no device %d", device);
    BIRP_CUDA(cudaSetDevice(device));
    const birp::CubeView& v = src->view;
    float* d_tmp = nullptr;
    int rc = BIRP_OK;
    cudaError_t e = cudaMalloc(&d_tmp, src->voxels * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyPeer(d_tmp, device, src->d_data, src->device, src->voxels * sizeof(float));
    if (e != cudaSuccess) {
        rc = set_error(e == cudaErrorMemoryAllocation ? BIRP_ERR_NOMEM : BIRP_ERR_CUDA, "cube replication to device %d: %s",
                       device, cudaGetErrorString(e));
    } else {
        const float origin[3] = {v.ox, v.oy, v.oz}, spacing[3] = {v.sx, v.sy, v.sz};
        rc = make_cube(d_tmp, true, v.nx, v.ny, v.nz, origin, spacing, nullptr, out);
    }
    if (d_tmp) cudaFree(d_tmp);
    cudaSetDevice(prev);
    return rc;
}

int birp_cube_device(const birp_cube* cube, int* device) {
    if (!cube || !device) return set_error(BIRP_ERR_ARG, "null pointer");
    *device = cube->device;
    return BIRP_OK;
}

int birp_cube_info(const birp_cube* cube, int info[4]) {
    if (!cube || !info) return set_error(BIRP_ERR_ARG, "null pointer");
    info[0] = cube->view.divmode;
    info[1] = cube->view.positive;
    info[2] = cube->view.tex != 0;
    info[3] = cube->view.quad != nullptr;
    return BIRP_OK;
}

int birp_cube_has_oct_records(const birp_cube* cube, int* present) {
    if (!cube || !present) return set_error(BIRP_ERR_ARG, "null pointer");
    *present = cube->view.oct != nullptr;
    return BIRP_OK;
}

static void destroy_cube(birp_cube* cube) {
    if (cube->view.tex) cudaDestroyTextureObject(cube->view.tex);
    if (cube->tex_array) cudaFreeArray(cube->tex_array);
    if (cube->d_quad) cudaFree(cube->d_quad);
    if (cube->d_oct) cudaFree(cube->d_oct);
    if (cube->d_poly) cudaFree(cube->d_poly);
    cudaFree(cube->d_data);
    delete cube;
}

// The poly copy of the tolerance-mode march (BIRP_FAST_MARCH): built once, at the first render that asks for it
// (8x the grid, only when it fits into half of the free device memory; without it the bit-exact march runs).
// The cube handle is const at the C ABI; the copy is a cache inside the opaque object.
static int ensure_poly(DeviceCtx* c, const birp_cube* cube_c, int flags) {
    if (!(flags & BIRP_FAST_MARCH) || !(flags & BIRP_TRILINEAR)) return BIRP_OK;
    birp_cube* cube = const_cast<birp_cube*>(cube_c);
    std::lock_guard<std::mutex> lk(c->mu);
    if (cube->poly_tried) return BIRP_OK;
    cube->poly_tried = true;
    const birp::CubeView& v = cube->view;
    if (v.nx < 2 || v.ny < 2 || v.nz < 2) return BIRP_OK;
    const size_t bytes = cube->voxels * 8 * sizeof(float);
    size_t free_b = 0, total_b = 0;
    cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
    if (e == cudaSuccess && free_b / 2 >= bytes) e = cudaMalloc(&cube->d_poly, bytes);
    else if (e == cudaSuccess) e = cudaErrorMemoryAllocation;
    if (e == cudaSuccess) e = birp::launch_pack_poly(v, cube->d_poly, c->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (cube->d_poly) cudaFree(cube->d_poly);
        cube->d_poly = nullptr;
        return BIRP_OK;  // optional copy: the exact march takes over
    }
    cube->view.poly = cube->d_poly;
    return BIRP_OK;
}

void birp_cube_free(birp_cube* cube) {
    if (cube) destroy_cube(cube);
}

int birp_render_cube(const birp_cube* cube, const birp_frame* frames, int nframes, int W, int H, int row_begin,
                     int row_end, int nsamples, int maxdist, int flags, float* images) {
    if (!cube) return set_error(BIRP_ERR_ARG, "null cube");
    int rc = check_geom(frames, nframes, W, H, &row_begin, &row_end, nsamples, images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    if (c->device != cube->device) return set_error(BIRP_ERR_ARG, "cube lives on device %d", cube->device);
    rc = ensure_poly(c, cube, flags);
    if (rc) return rc;
    const birp::CubeView view = cube->view;
    int* d_status = c->d_status;
    return render_host(c, frames, nframes, W, H, row_begin, row_end, nsamples, maxdist, flags, images,
                       [&](const birp_frame* d_fr, const birp::RenderGeom& g, const float* d_dist, float* d_img,
                           cudaStream_t s) {
                           return birp::launch_render_cube(view, d_fr, g, d_dist, flags, d_img, nullptr,
                                                           d_status, s);
                       });
}

int birp_render_cube_rows(const birp_cube* cube, const birp_frame* frames, int nframes, int W, int H, int band_rows,
                          int phase, int nphases, int nsamples, int maxdist, int flags, float* images) {
    if (!cube) return set_error(BIRP_ERR_ARG, "null cube");
    if (band_rows < 1 || nphases < 1 || phase < 0 || phase >= nphases)
        return set_error(BIRP_ERR_ARG, "bad row interleave: band %d phase %d of %d", band_rows, phase, nphases);
    int row_begin = 0, row_end = H;
    int rc = check_geom(frames, nframes, W, H, &row_begin, &row_end, nsamples, images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    if (c->device != cube->device) return set_error(BIRP_ERR_ARG, "cube lives on device %d", cube->device);
    rc = ensure_poly(c, cube, flags);
    if (rc) return rc;
    const birp::CubeView view = cube->view;
    int* d_status = c->d_status;
    HostPlan plan;
    plan.row_band = band_rows, plan.row_phase = phase, plan.row_nphases = nphases;
    return render_host(c, frames, nframes, W, H, row_begin, row_end, nsamples, maxdist, flags, images,
                       [&](const birp_frame* d_fr, const birp::RenderGeom& g, const float* d_dist, float* d_img,
                           cudaStream_t s) {
                           return birp::launch_render_cube(view, d_fr, g, d_dist, flags, d_img, nullptr,
                                                           d_status, s);
                       },
                       nullptr, &plan);
}

int birp_render_cmem(const birp_cmem_params* params, const birp_frame* frames, int nframes, int W, int H,
                     int row_begin, int row_end, int nsamples, int maxdist, int flags, float* images) {
    if (!params) return set_error(BIRP_ERR_ARG, "null params");
    int rc = check_geom(frames, nframes, W, H, &row_begin, &row_end, nsamples, images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    const birp_cmem_params P = *params;
    const float4* d_table = nullptr;
    rc = cmem_flags(c, P, flags, &flags, &d_table);
    if (rc) return rc;
    return render_host(c, frames, nframes, W, H, row_begin, row_end, nsamples, maxdist, flags, images,
                       [&](const birp_frame* d_fr, const birp::RenderGeom& g, const float* d_dist, float* d_img,
                           cudaStream_t s) {
                           return birp::launch_render_cmem(P, d_fr, g, d_dist, flags, d_img, nullptr, s, d_table);
                       });
}

static int check_sweep(const float* poses, int nframes, birp_frame* frames_out, std::vector<birp_frame>* scratch,
                       SweepPoses* sw, float aim_x, float fov_x_deg, float fov_y_deg) {
    if (nframes > 0 && !poses) return set_error(BIRP_ERR_ARG, "null poses");
    sw->poses = poses;
    sw->aim_x = aim_x;
    sw->fov_x_deg = fov_x_deg;
    sw->fov_y_deg = fov_y_deg;
    if (frames_out) {
        sw->frames_out = frames_out;
    } else {
        scratch->resize(nframes > 0 ? nframes : 1);
        sw->frames_out = scratch->data();
    }
    return BIRP_OK;
}

// one device's share of a sweep (the calling thread's device): everything birp_render_cube_sweep does, with the
// plan saying where the frames land and who hands them out
static int cube_sweep_on_device(const birp_cube* cube, const float* poses, int nframes, float aim_x, float fov_x_deg,
                                float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples, int maxdist,
                                int flags, birp_frame* frames_out, float* images, const HostPlan* plan) {
    if (!cube) return set_error(BIRP_ERR_ARG, "null cube");
    std::vector<birp_frame> scratch;
    SweepPoses sw;
    int rc = check_sweep(poses, nframes, frames_out, &scratch, &sw, aim_x, fov_x_deg, fov_y_deg);
    if (rc) return rc;
    rc = check_geom(sw.frames_out, nframes, W, H, &row_begin, &row_end, nsamples, images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    if (c->device != cube->device) return set_error(BIRP_ERR_ARG, "cube lives on device %d", cube->device);
    rc = ensure_poly(c, cube, flags);
    if (rc) return rc;
    const birp::CubeView view = cube->view;
    int* d_status = c->d_status;
    return render_host(c, sw.frames_out, nframes, W, H, row_begin, row_end, nsamples, maxdist, flags, images,
                       [&](const birp_frame* d_fr, const birp::RenderGeom& g, const float* d_dist, float* d_img,
                           cudaStream_t s) {
                           return birp::launch_render_cube(view, d_fr, g, d_dist, flags, d_img, nullptr, d_status, s);
                       },
                       &sw, plan);
}

static int cmem_sweep_on_device(const birp_cmem_params* params, const float* poses, int nframes, float aim_x,
                                float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples,
                                int maxdist, int flags, birp_frame* frames_out, float* images, const HostPlan* plan) {
    if (!params) return set_error(BIRP_ERR_ARG, "null params");
    std::vector<birp_frame> scratch;
    SweepPoses sw;
    int rc = check_sweep(poses, nframes, frames_out, &scratch, &sw, aim_x, fov_x_deg, fov_y_deg);
    if (rc) return rc;
    rc = check_geom(sw.frames_out, nframes, W, H, &row_begin, &row_end, nsamples, images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    const birp_cmem_params P = *params;
    const float4* d_table = nullptr;
    rc = cmem_flags(c, P, flags, &flags, &d_table);
    if (rc) return rc;
    return render_host(c, sw.frames_out, nframes, W, H, row_begin, row_end, nsamples, maxdist, flags, images,
                       [&](const birp_frame* d_fr, const birp::RenderGeom& g, const float* d_dist, float* d_img,
                           cudaStream_t s) {
                           return birp::launch_render_cmem(P, d_fr, g, d_dist, flags, d_img, nullptr, s, d_table);
                       },
                       &sw, plan);
}

int birp_render_cube_sweep(const birp_cube* cube, const float* poses, int nframes, float aim_x, float fov_x_deg,
                           float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples, int maxdist,
                           int flags, birp_frame* frames_out, float* images) {
    return cube_sweep_on_device(cube, poses, nframes, aim_x, fov_x_deg, fov_y_deg, W, H, row_begin, row_end, nsamples,
                                maxdist, flags, frames_out, images, nullptr);
}

int birp_render_cmem_sweep(const birp_cmem_params* params, const float* poses, int nframes, float aim_x,
                           float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples,
                           int maxdist, int flags, birp_frame* frames_out, float* images) {
    return cmem_sweep_on_device(params, poses, nframes, aim_x, fov_x_deg, fov_y_deg, W, H, row_begin, row_end, nsamples,
                                maxdist, flags, frames_out, images, nullptr);
}

static int check_stride(long long frame_offset, long long frame_stride) {
    if (frame_offset < 0 || frame_stride < 1) return set_error(BIRP_ERR_ARG, "bad frame offset/stride");
    return BIRP_OK;
}

int birp_render_cube_sweep_strided(const birp_cube* cube, const float* poses, int nframes, float aim_x, float fov_x_deg,
                                   float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples, int maxdist,
                                   int flags, birp_frame* frames_out, float* images, long long frame_offset,
                                   long long frame_stride) {
    int rc = check_stride(frame_offset, frame_stride);
    if (rc) return rc;
    HostPlan plan;
    plan.out_offset = frame_offset;
    plan.out_stride = frame_stride;
    return cube_sweep_on_device(cube, poses, nframes, aim_x, fov_x_deg, fov_y_deg, W, H, row_begin, row_end, nsamples,
                                maxdist, flags, frames_out, images, &plan);
}

int birp_render_cmem_sweep_strided(const birp_cmem_params* params, const float* poses, int nframes, float aim_x,
                                   float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end,
                                   int nsamples, int maxdist, int flags, birp_frame* frames_out, float* images,
                                   long long frame_offset, long long frame_stride) {
    int rc = check_stride(frame_offset, frame_stride);
    if (rc) return rc;
    HostPlan plan;
    plan.out_offset = frame_offset;
    plan.out_stride = frame_stride;
    return cmem_sweep_on_device(params, poses, nframes, aim_x, fov_x_deg, fov_y_deg, W, H, row_begin, row_end, nsamples,
                                maxdist, flags, frames_out, images, &plan);
}

// ---- one sweep over several GPUs of the box, from one process: a host thread per device, each taking the next
// chunk of frames from a shared counter (host frame setup -> H2D -> render -> D2H straight into the caller's buffer,
// double-buffered per device), so every GPU's copy engine and PCIe link carry their own share and all the images end
// up in ONE host buffer.  The first failure stops the hand-out; its code and message are returned.
}  // extern "C"
template <class Work>
static int sweep_multi(int ndevices, const int* devices, int nframes, int W, int H, Work&& work) {
    if (ndevices < 1 || !devices) return set_error(BIRP_ERR_ARG, "no devices");
    std::atomic<long long> feed{0};
    HostPlan plan;
    plan.feed = &feed;
    // a chunk per grab: about eight grabs per device (dynamic balance, and a short last chunk per device: the last
    // chunks of all devices drain together through the box's aggregate D2H bandwidth), at least 32 frames
    size_t chunk = ((size_t)nframes + (size_t)ndevices * 8 - 1) / ((size_t)ndevices * 8);
    if (chunk < 32) chunk = 32;
    plan.feed_chunk = chunk;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    plan.host_threads = (int)(hw / (unsigned)ndevices > 0 ? hw / (unsigned)ndevices : 1);
    if (plan.host_threads > 8) plan.host_threads = 8;
    (void)W, (void)H;
    std::vector<int> rcs(ndevices, BIRP_OK);
    std::vector<std::string> msgs(ndevices);
    std::vector<std::thread> pool;
    for (int d = 0; d < ndevices; ++d)
        pool.emplace_back([&, d] {
            cudaError_t e = cudaSetDevice(devices[d]);
            int rc = (e == cudaSuccess) ? work(d, &plan)
                                        : set_error(BIRP_ERR_CUDA, "cudaSetDevice(%d): %s", devices[d], cudaGetErrorString(e));
            if (rc != BIRP_OK) {
                feed.store((long long)nframes);  // nobody starts another chunk
                msgs[d] = g_error;
            }
            rcs[d] = rc;
        });
    for (auto& th : pool) th.join();
    for (int d = 0; d < ndevices; ++d)
        if (rcs[d] != BIRP_OK) {
            g_error = msgs[d];
            return rcs[d];
        }
    return BIRP_OK;
}
extern "C" {

int birp_render_cube_sweep_multi(const birp_cube* const* cubes, int ndevices, const float* poses, int nframes,
                                 float aim_x, float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end,
                                 int nsamples, int maxdist, int flags, birp_frame* frames_out, float* images) {
    if (!cubes || ndevices < 1) return set_error(BIRP_ERR_ARG, "no cubes");
    std::vector<int> devices(ndevices);
    for (int d = 0; d < ndevices; ++d) {
        if (!cubes[d]) return set_error(BIRP_ERR_ARG, "null cube");
        devices[d] = cubes[d]->device;
        for (int k = 0; k < d; ++k)
            if (devices[k] == devices[d]) return set_error(BIRP_ERR_ARG, "two cubes on device %d", devices[d]);
    }
    if (nframes <= 0) return nframes == 0 ? BIRP_OK : set_error(BIRP_ERR_ARG, "negative frame count");
    std::vector<birp_frame> scratch;
    if (!frames_out) {  // one record buffer for all device threads
        scratch.resize((size_t)nframes);
        frames_out = scratch.data();
    }
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = sweep_multi(ndevices, devices.data(), nframes, W, H, [&](int d, const HostPlan* plan) {
        return cube_sweep_on_device(cubes[d], poses, nframes, aim_x, fov_x_deg, fov_y_deg, W, H, row_begin, row_end,
                                    nsamples, maxdist, flags, frames_out, images, plan);
    });
    cudaSetDevice(prev);
    return rc;
}

int birp_render_cmem_sweep_multi(const birp_cmem_params* params, const int* devices, int ndevices, const float* poses,
                                 int nframes, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H, int row_begin,
                                 int row_end, int nsamples, int maxdist, int flags, birp_frame* frames_out, float* images) {
    if (!params) return set_error(BIRP_ERR_ARG, "null params");
    if (nframes <= 0) return nframes == 0 ? BIRP_OK : set_error(BIRP_ERR_ARG, "negative frame count");
    std::vector<birp_frame> scratch;
    if (!frames_out) {
        scratch.resize((size_t)nframes);
        frames_out = scratch.data();
    }
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = sweep_multi(ndevices, devices, nframes, W, H, [&](int, const HostPlan* plan) {
        return cmem_sweep_on_device(params, poses, nframes, aim_x, fov_x_deg, fov_y_deg, W, H, row_begin, row_end, nsamples,
                                    maxdist, flags, frames_out, images, plan);
    });
    cudaSetDevice(prev);
    return rc;
}

int birp_accumulate(const float* samples, long nrays, int nsamples, float ray_width, int flags, float* image) {
    if (!samples || !image) return set_error(BIRP_ERR_ARG, "null pointer");
    if (nrays < 0 || nsamples <= 0) return set_error(BIRP_ERR_ARG, "bad sizes");
    if (nrays == 0) return BIRP_OK;
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t ns = (size_t)nrays * nsamples;
    rc = ensure(&c->d_scratch, &c->scratch_cap, ns + (size_t)nrays);
    if (rc) return rc;
    float* d_samples = c->d_scratch;
    float* d_image = c->d_scratch + ns;
    BIRP_CUDA(cudaMemcpyAsync(d_samples, samples, ns * sizeof(float), cudaMemcpyHostToDevice, c->compute));
    if (flags & BIRP_ACCUMULATE)
        BIRP_CUDA(cudaMemcpyAsync(d_image, image, nrays * sizeof(float), cudaMemcpyHostToDevice, c->compute));
    cudaError_t e = birp::launch_accumulate(d_samples, nrays, nsamples, ray_width, flags, d_image, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(image, d_image, nrays * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    return BIRP_OK;
}

int birp_ray_directions(const birp_frame* frame, int W, int H, float* dirs) {
    if (!frame || !dirs) return set_error(BIRP_ERR_ARG, "null pointer");
    if (W <= 0 || H <= 0 || (long)W * H > (1L << 30)) return set_error(BIRP_ERR_ARG, "bad image size");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t n = (size_t)W * H * 3;
    rc = ensure(&c->d_scratch, &c->scratch_cap, n + 16);
    if (rc) return rc;
    birp_frame* d_frame = reinterpret_cast<birp_frame*>(c->d_scratch);
    float* d_dirs = c->d_scratch + 16;
    BIRP_CUDA(cudaMemcpyAsync(d_frame, frame, sizeof(birp_frame), cudaMemcpyHostToDevice, c->compute));
    cudaError_t e = birp::launch_ray_directions(d_frame, W, H, d_dirs, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(dirs, d_dirs, n * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    return BIRP_OK;
}

int birp_sample_cube(const birp_cube* cube, const float* xyz, long n, int flags, float* out) {
    if (!cube || !xyz || !out) return set_error(BIRP_ERR_ARG, "null pointer");
    if (n < 0) return set_error(BIRP_ERR_ARG, "negative count");
    if (n == 0) return BIRP_OK;
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    if (c->device != cube->device) return set_error(BIRP_ERR_ARG, "cube lives on device %d", cube->device);
    std::lock_guard<std::mutex> lk(c->mu);
    rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)n * 4);
    if (rc) return rc;
    float* d_xyz = c->d_scratch;
    float* d_out = c->d_scratch + (size_t)n * 3;
    BIRP_CUDA(cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->compute));
    cudaError_t e = birp::launch_sample_cube(cube->view, d_xyz, n, flags, d_out, c->d_status, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    return read_status(c, c->compute);
}

int birp_sample_cmem(const birp_cmem_params* params, const float* xyz, long n, int flags, float* out) {
    if (!params || !xyz || !out) return set_error(BIRP_ERR_ARG, "null pointer");
    if (n < 0) return set_error(BIRP_ERR_ARG, "negative count");
    if (n == 0) return BIRP_OK;
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    const float4* d_table = nullptr;
    rc = cmem_flags(c, *params, flags, &flags, &d_table);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)n * 4);
    if (rc) return rc;
    float* d_xyz = c->d_scratch;
    float* d_out = c->d_scratch + (size_t)n * 3;
    BIRP_CUDA(cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->compute));
    cudaError_t e = birp::launch_sample_cmem(*params, d_xyz, n, flags, d_out, c->compute, d_table);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    return BIRP_OK;
}

int birp_render_cube_device(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H,
                            int row_begin, int row_end, int nsamples, int maxdist, int flags, float* d_images,
                            unsigned long long* d_valid_samples, void* stream) {
    return birp_render_cube_device_strided(cube, d_frames, nframes, W, H, row_begin, row_end, nsamples, maxdist,
                                           flags, d_images, 0, 1, d_valid_samples, stream);
}

static int cube_device_launch(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H, int row_begin,
                              int row_end, int band_rows, int phase, int nphases, int nsamples, int maxdist, int flags,
                              float* d_images, long long frame_offset, long long frame_stride,
                              unsigned long long* d_valid_samples, void* stream);

int birp_render_cube_device_strided(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H,
                                    int row_begin, int row_end, int nsamples, int maxdist, int flags,
                                    float* d_images, long long frame_offset, long long frame_stride,
                                    unsigned long long* d_valid_samples, void* stream) {
    return cube_device_launch(cube, d_frames, nframes, W, H, row_begin, row_end, 0, 0, 1, nsamples, maxdist, flags, d_images,
                              frame_offset, frame_stride, d_valid_samples, stream);
}

int birp_render_cube_device_rows(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H,
                                 int band_rows, int phase, int nphases, int nsamples, int maxdist, int flags,
                                 float* d_images, long long frame_offset, long long frame_stride,
                                 unsigned long long* d_valid_samples, void* stream) {
    if (band_rows < 1 || nphases < 1 || phase < 0 || phase >= nphases)
        return set_error(BIRP_ERR_ARG, "bad row interleave: band %d phase %d of %d", band_rows, phase, nphases);
    return cube_device_launch(cube, d_frames, nframes, W, H, 0, H, band_rows, phase, nphases, nsamples, maxdist, flags,
                              d_images, frame_offset, frame_stride, d_valid_samples, stream);
}

static int cube_device_launch(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H, int row_begin,
                              int row_end, int band_rows, int phase, int nphases, int nsamples, int maxdist, int flags,
                              float* d_images, long long frame_offset, long long frame_stride,
                              unsigned long long* d_valid_samples, void* stream) {
    if (!cube) return set_error(BIRP_ERR_ARG, "null cube");
    if (frame_offset < 0 || frame_stride < 1) return set_error(BIRP_ERR_ARG, "bad frame offset/stride");
    int rc = check_geom(d_frames, nframes, W, H, &row_begin, &row_end, nsamples, d_images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    if (c->device != cube->device) return set_error(BIRP_ERR_ARG, "cube lives on device %d", cube->device);
    rc = ensure_poly(c, cube, flags);
    if (rc) return rc;
    RayTable table;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        rc = get_ray_table(c, nsamples, maxdist, (cudaStream_t)stream, &table);
        if (rc) return rc;
    }
    birp::RenderGeom g{nframes, W, H, row_begin, row_end, nsamples, table.ray_width};
    g.monotone_table = table.monotone;
    g.tile_w_log2 = c->tile_w_log2;
    g.fast_variant = c->fast_variant;
    g.tile_swizzle = c->tile_swizzle;
    g.cta_warps = c->cta_warps;
    g.fast_kstep = c->fast_kstep;
    g.fast_coop = c->fast_coop;
    g.fast_step = table.step;
    g.out_frame_offset = frame_offset;
    g.out_frame_stride = frame_stride;
    g.row_band = band_rows, g.row_phase = phase, g.row_nphases = nphases;
    cudaError_t e = birp::launch_render_cube(cube->view, d_frames, g, table.d_dist, flags, d_images,
                                             d_valid_samples, c->d_status, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return BIRP_OK;
}

int birp_render_cmem_device(const birp_cmem_params* params, const birp_frame* d_frames, int nframes, int W,
                            int H, int row_begin, int row_end, int nsamples, int maxdist, int flags,
                            float* d_images, unsigned long long* d_valid_samples, void* stream) {
    return birp_render_cmem_device_strided(params, d_frames, nframes, W, H, row_begin, row_end, nsamples, maxdist,
                                           flags, d_images, 0, 1, d_valid_samples, stream);
}

int birp_render_cmem_device_strided(const birp_cmem_params* params, const birp_frame* d_frames, int nframes, int W,
                                    int H, int row_begin, int row_end, int nsamples, int maxdist, int flags,
                                    float* d_images, long long frame_offset, long long frame_stride,
                                    unsigned long long* d_valid_samples, void* stream) {
    if (!params) return set_error(BIRP_ERR_ARG, "null params");
    if (frame_offset < 0 || frame_stride < 1) return set_error(BIRP_ERR_ARG, "bad frame offset/stride");
    int rc = check_geom(d_frames, nframes, W, H, &row_begin, &row_end, nsamples, d_images);
    if (rc) return rc;
    DeviceCtx* c;
    rc = get_ctx(&c);
    if (rc) return rc;
    RayTable table;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        rc = get_ray_table(c, nsamples, maxdist, (cudaStream_t)stream, &table);
        if (rc) return rc;
    }
    birp::RenderGeom g{nframes, W, H, row_begin, row_end, nsamples, table.ray_width};
    g.monotone_table = table.monotone;
    g.tile_w_log2 = c->tile_w_log2;
    g.fast_variant = c->fast_variant;
    g.tile_swizzle = c->tile_swizzle;
    g.cta_warps = c->cta_warps;
    g.fast_kstep = c->fast_kstep;
    g.fast_coop = c->fast_coop;
    g.fast_step = table.step;
    g.out_frame_offset = frame_offset;
    g.out_frame_stride = frame_stride;
    const float4* d_table = nullptr;
    rc = cmem_flags(c, *params, flags, &flags, &d_table);  // (first use of a parameter set: table + self-tests on the library's stream)
    if (rc) return rc;
    cudaError_t e = birp::launch_render_cmem(*params, d_frames, g, table.d_dist, flags, d_images,
                                             d_valid_samples, (cudaStream_t)stream, d_table);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return BIRP_OK;
}

// ---- device memory that other processes of the box can map (peer stores over NVLink)
int birp_device_alloc(size_t bytes, void** d_ptr) {
    if (!d_ptr) return set_error(BIRP_ERR_ARG, "null out pointer");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return set_error(e == cudaErrorMemoryAllocation ? BIRP_ERR_NOMEM : BIRP_ERR_CUDA, "cudaMalloc(%zu): %s", bytes,
                         cudaGetErrorString(e));
    return BIRP_OK;
}

int birp_device_free(void* d_ptr) {
    if (d_ptr) BIRP_CUDA(cudaFree(d_ptr));
    return BIRP_OK;
}

int birp_ipc_export(const void* d_ptr, unsigned char handle[64]) {
    if (!d_ptr || !handle) return set_error(BIRP_ERR_ARG, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    BIRP_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    std::memcpy(handle, &h, 64);
    return BIRP_OK;
}

int birp_ipc_import(const unsigned char handle[64], void** d_ptr) {
    if (!handle || !d_ptr) return set_error(BIRP_ERR_ARG, "null pointer");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    BIRP_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BIRP_OK;
}

int birp_ipc_close(void* d_ptr) {
    if (d_ptr) BIRP_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return BIRP_OK;
}

int birp_peer_signal(unsigned* d_flag, unsigned value, void* stream) {
    if (!d_flag) return set_error(BIRP_ERR_ARG, "null flag");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    cudaError_t e = birp::launch_peer_signal(d_flag, value, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return BIRP_OK;
}

int birp_peer_wait(const unsigned* d_flags, int n, unsigned value, void* stream) {
    if (!d_flags || n < 1 || n > 32) return set_error(BIRP_ERR_ARG, "bad flag array");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    cudaError_t e = birp::launch_peer_wait(d_flags, n, value, c->d_status, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return BIRP_OK;
}

int birp_check_device_status(void* stream) {
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    return read_status(c, (cudaStream_t)stream);
}

int birp_selftest_cmem_math(long n, double max_err[6]) {
    if (n <= 0 || !max_err) return set_error(BIRP_ERR_ARG, "bad arguments");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)16);
    if (rc) return rc;
    unsigned long long* d_err = reinterpret_cast<unsigned long long*>(c->d_scratch);
    BIRP_CUDA(cudaMemsetAsync(d_err, 0, 6 * sizeof(unsigned long long), c->compute));
    cudaError_t e = birp::launch_cmem_math_selftest(n, d_err, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(max_err, d_err, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    return BIRP_OK;
}

int birp_selftest_div10(unsigned long long mismatches[2]) {
    if (!mismatches) return set_error(BIRP_ERR_ARG, "bad arguments");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)16);
    if (rc) return rc;
    unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(c->d_scratch);
    BIRP_CUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned long long), c->compute));
    cudaError_t e = birp::launch_div10_selftest(d_bad, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    BIRP_CUDA(cudaMemcpyAsync(mismatches, d_bad, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    return BIRP_OK;
}

int birp_selftest_cmem_filter(const birp_cmem_params* params, long n, double out[3]) {
    if (!params || n <= 0 || !out) return set_error(BIRP_ERR_ARG, "bad arguments");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    return run_cmem_filter_selftest(c, *params, n, out);
}

int birp_selftest_cmem_table(const birp_cmem_params* params, long n, double out[3]) {
    if (!params || n <= 0 || !out) return set_error(BIRP_ERR_ARG, "bad arguments");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    int flags = 0;
    const float4* d_table = nullptr;
    rc = cmem_flags(c, *params, 0, &flags, &d_table);  // builds (and self-tests) the table at first use
    if (rc) return rc;
    out[0] = out[1] = 0.0;
    out[2] = d_table ? 1.0 : 0.0;
    if (!d_table) return BIRP_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    rc = ensure(&c->d_scratch, &c->scratch_cap, (size_t)16);
    if (rc) return rc;
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(c->d_scratch);
    BIRP_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(unsigned long long), c->compute));
    cudaError_t e = birp::launch_cmem_table_selftest(*params, d_table, n, d_out, c->compute);
    if (e != cudaSuccess) return set_error(BIRP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    unsigned long long h[2];
    BIRP_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->compute));
    BIRP_CUDA(cudaStreamSynchronize(c->compute));
    out[0] = (double)h[0];
    out[1] = (double)h[1] / (double)n;
    return BIRP_OK;
}

unsigned long long birp_launch_count(void) { return birp::launch_count(); }

int birp_set_tuning(const char* name, int value) {
    if (!name) return set_error(BIRP_ERR_ARG, "null name");
    DeviceCtx* c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const std::string n(name);
    if (n == "tile_w_log2") c->tile_w_log2 = value;
    else if (n == "tile_swizzle") c->tile_swizzle = value;
    else if (n == "cta_warps") c->cta_warps = value;
    else if (n == "fast_variant") c->fast_variant = value;
    else if (n == "fast_kstep") c->fast_kstep = value;
    else if (n == "fast_coop") c->fast_coop = value;
    else return set_error(BIRP_ERR_ARG, "unknown tuning knob '%s'", name);
    return BIRP_OK;
}

}  // extern "C"
