// Directory batch: parse ahead on worker threads, upload + render + export on the calling thread
// (one GPU context per process; the render of an SXI frame is ~0.1 ms, the ASCII parse of a
// 101x161x161 cube ~0.1 s, so the parse is what has to be spread out).
#include "Batch.h"

#include <algorithm>
#include <condition_variable>
#include <exception>
#include <filesystem>
#include <memory>
#include <mutex>
#include <thread>

#include <cmath>
#include <stdexcept>

#include "../../include/birp_b200.h"
#include "Camera.h"
#include "Cmem.h"
#include "DataCube.h"

int BirpDeviceCount() {
    int n = 0;
    if (birp_device_count(&n) != BIRP_OK || n < 1) return 1;
    if (const char* env = std::getenv("BIRP_DEVICES")) {
        const int want = std::atoi(env);
        if (want >= 1 && want < n) n = want;
    }
    return n;
}

namespace {
struct Slot {
    std::unique_ptr<DataCube> cube;
    std::string error;
    bool ready = false;
};
}  // namespace

std::vector<BatchResult> RenderBatchDirectory(const std::string& input_dir, const std::string& output_dir,
                                              float pixel_size_deg, float plot_fov_h, float plot_fov_w,
                                              float pos_x, float pos_y, float pos_z, float aim_x, bool trilinear,
                                              int parser_threads, int devices) {
    namespace fs = std::filesystem;
    std::vector<std::string> files;
    for (const auto& entry : fs::directory_iterator(input_dir))
        if (entry.is_regular_file() && entry.path().extension() != ".birpcache")  // (side-cars of DataCube::Load)
            files.push_back(entry.path().string());
    std::sort(files.begin(), files.end());

    const size_t n = files.size();
    std::vector<Slot> slots(n);
    std::mutex mu;
    std::condition_variable cv;
    size_t next_to_parse = 0, next_to_render = 0;
    if (parser_threads <= 0) parser_threads = (int)std::min<size_t>(8, std::max<size_t>(1, std::thread::hardware_concurrency()));
    // parsed-but-not-rendered cubes kept in memory (the render threads hold up to two per device on top)
    const size_t window = (size_t)parser_threads + 1 + 2 * (size_t)(devices > 0 ? devices : BirpDeviceCount());
    // each parser converts its file on its own share of the cores (a Load left to itself starts up to 16 workers:
    // 8 parsers x 16 workers would oversubscribe the host and multiply the parse buffers)
    const int load_threads = (int)std::max<size_t>(1, std::max<size_t>(1, std::thread::hardware_concurrency()) / (size_t)parser_threads);

    auto parser = [&]() {
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next_to_parse >= n || next_to_parse < next_to_render + window; });
                if (next_to_parse >= n) return;
                i = next_to_parse++;
            }
            Slot s;
            try {
                s.cube = std::make_unique<DataCube>();
                s.cube->LoadParallel(files[i], false, load_threads);
                s.cube->SetTrilinear(trilinear);
            } catch (const std::exception& e) {
                s.cube.reset();
                s.error = e.what();
            }
            s.ready = true;
            {
                std::lock_guard<std::mutex> lk(mu);
                slots[i] = std::move(s);
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < parser_threads; ++t) pool.emplace_back(parser);

    std::vector<BatchResult> results(n);
    std::string out_prefix = output_dir;
    if (!out_prefix.empty() && out_prefix.back() != '/') out_prefix += '/';
    if (devices <= 0) devices = BirpDeviceCount();
    size_t next_to_claim = 0;  // next cube a render thread takes (in order; next_to_render trails it for the window)
    size_t finished = 0;
    auto render = [&](int device) {
        birp_set_device(device);
        for (;;) {
            size_t i;
            Slot s;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (next_to_claim >= n) return;
                i = next_to_claim++;
                cv.wait(lk, [&] { return slots[i].ready; });
                s = std::move(slots[i]);
            }
            results[i].input = files[i];
            if (!s.cube) {
                results[i].error = s.error;
            } else {
                try {
                    Camera camera(*s.cube, pixel_size_deg, plot_fov_h, plot_fov_w);
                    camera.SetPosition(pos_x, pos_y, pos_z);
                    camera.SetAim(aim_x, 0.f, 0.f);
                    camera.Render();
                    const std::string base = files[i].substr(files[i].find_last_of("/\\") + 1);
                    camera.ToFITS(out_prefix + base);
                    results[i].output = out_prefix + base + ".fits";
                } catch (const std::exception& e) {
                    results[i].error = e.what();
                }
                s.cube.reset();  // (frees the device copy on this thread's device)
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                ++finished;
                next_to_render = finished;  // the parse-ahead window counts cubes that are done
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> renderers;
    for (int t = 0; t < 2 * devices; ++t) renderers.emplace_back(render, t % devices);
    for (auto& th : renderers) th.join();
    for (auto& th : pool) th.join();
    return results;
}

std::vector<float> RenderSweep(Space& space, const std::vector<float>& positions_xyz, float aim_x,
                               float pixel_size_deg, float plot_fov_h, float plot_fov_w, int* width, int* height,
                               int samples, int max_distance, int devices) {
    // image size and fields of view exactly as Camera's constructor derives them (src/Camera.cpp:32-39)
    const int H = (int)std::round(plot_fov_h / pixel_size_deg), W = (int)std::round(plot_fov_w / pixel_size_deg);
    if (width) *width = W;
    if (height) *height = H;
    if (positions_xyz.size() % 3 != 0) throw std::invalid_argument("RenderSweep: positions must be x, y, z triples");
    const int n = (int)(positions_xyz.size() / 3);
    std::vector<float> images((size_t)n * H * W, 0.f);
    if (n == 0 || W <= 0 || H <= 0) return images;
    int rc;
    if (devices <= 0) devices = BirpDeviceCount();
    if (n < 64 * devices) devices = 1;  // too few frames to be worth a second GPU
    if (auto* cube = dynamic_cast<DataCube*>(&space)) {
        const int flags = (cube->IsTrilinear() ? BIRP_TRILINEAR : BIRP_NEAREST) | (cube->FastMarch() ? BIRP_FAST_MARCH : 0);
        if (devices > 1) {
            std::vector<const birp_cube*> cubes(devices);
            for (int d = 0; d < devices; ++d) cubes[d] = cube->ResidentOn(d);
            rc = birp_render_cube_sweep_multi(cubes.data(), devices, positions_xyz.data(), n, aim_x, plot_fov_w,
                                              plot_fov_h, W, H, 0, H, samples, max_distance, flags, nullptr,
                                              images.data());
        } else {
            rc = birp_render_cube_sweep(cube->Resident(), positions_xyz.data(), n, aim_x, plot_fov_w, plot_fov_h, W, H,
                                        0, H, samples, max_distance, flags, nullptr, images.data());
        }
    } else if (auto* cmem = dynamic_cast<CMEM*>(&space)) {
        const int flags = cmem->FastMath() ? BIRP_CMEM_FAST : 0;
        if (devices > 1) {
            std::vector<int> ids(devices);
            for (int d = 0; d < devices; ++d) ids[d] = d;
            rc = birp_render_cmem_sweep_multi(&cmem->Params(), ids.data(), devices, positions_xyz.data(), n, aim_x,
                                              plot_fov_w, plot_fov_h, W, H, 0, H, samples, max_distance, flags, nullptr,
                                              images.data());
        } else {
            rc = birp_render_cmem_sweep(&cmem->Params(), positions_xyz.data(), n, aim_x, plot_fov_w, plot_fov_h, W, H, 0,
                                        H, samples, max_distance, flags, nullptr, images.data());
        }
    } else {
        throw std::invalid_argument("RenderSweep: the Space must be a DataCube or a CMEM");
    }
    if (rc == BIRP_ERR_SINGULAR) throw std::runtime_error("Determinant is 0");
    if (rc == BIRP_ERR_RANGE) throw std::out_of_range("vector::_M_range_check");
    if (rc != BIRP_OK) throw std::runtime_error(birp_last_error());
    return images;
}
