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

// Renders every regular file of `input_dir` (sorted by name; side-cars of DataCube::Load are skipped) and writes
// `<output_dir>/<file name>.fits`, exactly the names the reference CLI produces.  `parser_threads` <= 0 picks a
// default.  `devices`: how many GPUs of the box take cubes (<= 0: all visible, or BIRP_DEVICES); the cubes go round
// the devices as they finish parsing, two render threads per device so that the upload of one cube overlaps the
// render and export of another.
std::vector<BatchResult> RenderBatchDirectory(const std::string& input_dir, const std::string& output_dir,
                                              float pixel_size_deg, float plot_fov_h, float plot_fov_w,
                                              float pos_x, float pos_y, float pos_z, float aim_x, bool trilinear,
                                              int parser_threads = 0, int devices = 0);

// how many GPUs multi-device calls use: min(visible devices, BIRP_DEVICES if set), at least 1
int BirpDeviceCount();

// An orbit sweep: what the reference would do with one Camera per position -- Camera(space, px, fov_h, fov_w),
// SetPosition(pos), SetAim(aim_x, 0, 0), Render() -- for n positions in one call (birp_render_*_sweep: the frame
// records of a chunk are made on host threads while the GPU renders the previous chunk).  `space` must be a
// DataCube or a CMEM (a user-defined Space has to be sampled on the host, frame by frame: use Camera).
// positions: n x {x, y, z}; returns n images of Height x Width floats, image k at [k * H * W], row j = image y,
// top row first (Camera::image[j][i] flattened).  Throws what Camera::Render throws.
class Space;
// `devices`: GPUs of the box that share the sweep (<= 0: all visible, or BIRP_DEVICES; 1: the calling thread's
// current device only).  With several, a DataCube is replicated device to device, every GPU takes chunks of frames
// from a shared counter and drains its images into the one result (birp_render_*_sweep_multi); the images are the
// single-GPU images, frame for frame.
std::vector<float> RenderSweep(Space& space, const std::vector<float>& positions_xyz, float aim_x,
                               float pixel_size_deg, float plot_fov_h, float plot_fov_w, int* width = nullptr,
                               int* height = nullptr, int samples = 200, int max_distance = 44, int devices = 0);
