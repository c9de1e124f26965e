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

#include "Camera.h"
#include "DataCube.h"

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
                                              int parser_threads) {
    namespace fs = std::filesystem;
    std::vector<std::string> files;
    for (const auto& entry : fs::directory_iterator(input_dir))
        if (entry.is_regular_file()) files.push_back(entry.path().string());
    std::sort(files.begin(), files.end());

    const size_t n = files.size();
    std::vector<Slot> slots(n);
    std::mutex mu;
    std::condition_variable cv;
    size_t next_to_parse = 0, next_to_render = 0;
    if (parser_threads <= 0) parser_threads = (int)std::min<size_t>(8, std::max<size_t>(1, std::thread::hardware_concurrency()));
    const size_t window = (size_t)parser_threads + 1;  // parsed-but-not-rendered cubes kept in memory

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
                s.cube->Load(files[i], false);
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
    for (size_t i = 0; i < n; ++i) {
        Slot s;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return slots[i].ready; });
            s = std::move(slots[i]);
            next_to_render = i + 1;
        }
        cv.notify_all();
        results[i].input = files[i];
        if (!s.cube) {
            results[i].error = s.error;
            continue;
        }
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
    }
    for (auto& th : pool) th.join();
    return results;
}
