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
