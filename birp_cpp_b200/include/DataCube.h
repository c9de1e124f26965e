// DataCube -- an MHD emissivity grid, resident on the GPU while it is being rendered.
//
// Public surface = the reference's include/DataCube.h:22-50 (same member names and types, so
// src/main.cpp and tests/DataCube_test.cpp compile unchanged).  `slices[z][y][x]` stays the
// host-side truth; Camera::Render uploads it once (flattened, x fastest) through
// birp_cube_upload and reuses the device copy until Load()/Invalidate() marks it stale.
#pragma once
#include <iostream>
#include <string>
#include <vector>

#include "Space.h"

struct birp_cube;  // include/birp_b200.h

class DataCube : public Space {
public:
    DataCube();
    ~DataCube() override;
    DataCube(const DataCube&) = delete;
    DataCube& operator=(const DataCube&) = delete;

    void Init() override;
    void Init(std::string& filename, bool debug);
    void Load(std::string filename, bool debug);  // reference ASCII layout, src/DataCube.cpp:28-138
    float GetSample(float x, float y, float z) override;  // evaluated by the device sampler
    void SetTrilinear(bool);

    std::vector<float> coords_x;
    std::vector<float> coords_y;
    std::vector<float> coords_z;
    struct {
        int x;
        int y;
        int z;
    } size;

    float spacing[3];
    float origin_coords[3];

    std::vector<std::vector<std::vector<float>>> slices;

    // ---- additions (not in the reference) -----------------------------------------------
    bool IsTrilinear() const { return trilinear; }
    // Call after editing `slices`, `coords_*`, `spacing` or `size` by hand once the cube has
    // been rendered or sampled; Load() does it itself.
    void Invalidate();
    // Device-resident copy (uploads on first use).  Throws std::runtime_error on CUDA failure.
    const birp_cube* Resident();

private:
    void InitSpacing();
    void InitOriginCoords();
    void InitSize();
    float GetAxisSpacing(std::vector<float> coords);
    bool trilinear = false;  // the reference leaves this uninitialised until SetTrilinear
    birp_cube* device_cube = nullptr;
};
