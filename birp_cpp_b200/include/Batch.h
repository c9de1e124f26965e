// Batch rendering of a directory of cubes with one camera pose -- what the reference's CLI does in
// its batch mode (src/main.cpp:242-264: for every file of the directory: Load, Render, ToFITS), with
// the ASCII parsing of the following cubes running on worker threads while the current cube is
// uploaded and rendered.  Not part of the reference's API ("next").
#pragma once
#include <string>
#include <vector>

struct BatchResult {
    std::string input;   // cube file
    std::string output;  // FITS file written ("" if the cube failed to load)
    std::string error;   // exception text if it failed
};

// Renders every regular file of `input_dir` (sorted by name) and writes `<output_dir>/<file name>.fits`,
// exactly the names the reference CLI produces.  `parser_threads` <= 0 picks a default.
std::vector<BatchResult> RenderBatchDirectory(const std::string& input_dir, const std::string& output_dir,
                                              float pixel_size_deg, float plot_fov_h, float plot_fov_w,
                                              float pos_x, float pos_y, float pos_z, float aim_x, bool trilinear,
                                              int parser_threads = 0);

// An orbit sweep: what the reference would do with one Camera per position -- Camera(space, px, fov_h, fov_w),
// SetPosition(pos), SetAim(aim_x, 0, 0), Render() -- for n positions in one call (birp_render_*_sweep: the frame
// records of a chunk are made on host threads while the GPU renders the previous chunk).  `space` must be a
// DataCube or a CMEM (a user-defined Space has to be sampled on the host, frame by frame: use Camera).
// positions: n x {x, y, z}; returns n images of Height x Width floats, image k at [k * H * W], row j = image y,
// top row first (Camera::image[j][i] flattened).  Throws what Camera::Render throws.
class Space;
std::vector<float> RenderSweep(Space& space, const std::vector<float>& positions_xyz, float aim_x,
                               float pixel_size_deg, float plot_fov_h, float plot_fov_w, int* width = nullptr,
                               int* height = nullptr, int samples = 200, int max_distance = 44);
