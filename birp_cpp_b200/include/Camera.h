// Camera -- SMILE SXI line-of-sight renderer.
//
// Public surface = the reference's include/Camera.h:20-64: construct with a Space, set the
// pose, Render(), read `image[j][i]` (row j = image y, top row first) or export it.
// Render() does the once-per-frame orientation math on the host (birp_frame_setup) and hands
// the whole pixel x sample loop of the reference (src/Camera.cpp:140-232) to the GPU.
#pragma once
#include <string>
#include <vector>

#include "DataCube.h"

class Camera {
public:
    Space* dataCube;  // may be a DataCube, a CMEM or any other Space
    std::vector<std::vector<float>> image;

    Camera(Space& cube, float pixel_size_deg, float plot_fov_h, float plot_fov_w);
    ~Camera();

    void SetPosition(float&, float&, float&);  // GSE position, Earth radii
    void SetAim(float, float, float);          // all three reach Orient(); theta / phi use x only, like the reference
    void Render();                             // 200 samples out to 44 RE, accumulated into `image`
    int ToDat(std::string filename);           // CSV rows, "<filename>.dat"
    int ToFITS(std::string filename);          // single-image FITS, "<filename>.fits"

    // ---- additions (not in the reference) -----------------------------------------------
    // March used by Render(); the reference hard-codes 200 / 44 (src/Camera.cpp:31,65).
    void SetMarch(int samples, int max_distance);
    int Width() const { return image_dimension_x; }
    int Height() const { return image_dimension_y; }

private:
    int image_dimension_x;
    int image_dimension_y;
    float fov_x;
    float fov_y;
    struct {
        float x;
        float y;
        float z;
    } position{0.f, 0.f, 0.f};
    struct {
        float x;
        float y;
        float z;
    } aim{0.f, 0.f, 0.f};
    float pixel_size_deg;
    int ray_samples;
    int ray_max_distance;
};
