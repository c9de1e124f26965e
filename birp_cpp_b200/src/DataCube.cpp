// DataCube: reference-compatible ASCII loader and GPU residency.
#include "DataCube.h"

#include <charconv>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <stdexcept>

#include "../../include/birp_b200.h"
#include "Helper.h"

namespace {
// whitespace-separated floats of one line, appended to `out` (what explode_float(line, ' ') yields)
void append_floats(const std::string& line, std::vector<float>& out) {
    const char* p = line.data();
    const char* end = p + line.size();
    while (p < end) {
        while (p < end && *p == ' ') ++p;
        if (p >= end) break;
        const char* q = p;
        while (q < end && *q != ' ') ++q;
        float v;
        auto res = std::from_chars(p, q, v);
        if (res.ec != std::errc() || res.ptr != q) v = std::stof(std::string(p, q));  // "+1.0", "nan", ...
        out.push_back(v);
        p = q;
    }
}
}  // namespace

DataCube::DataCube() : size{0, 0, 0}, spacing{0.f, 0.f, 0.f}, origin_coords{0.f, 0.f, 0.f} {}

DataCube::~DataCube() { Invalidate(); }

void DataCube::Init() {}

void DataCube::Init(std::string& filename, bool debug = false) { this->Load(filename, debug); }

// Layout (reference src/DataCube.cpp:28-138): line 1 is a free-text header, line 10 holds
// "nx ny nz", the following lines list the x, then y, then z node coordinates (each axis may
// wrap over several lines), and the rest are the values, x fastest: every run of nx values is
// one "thread", ny threads make a slice.  Blank lines between slices are tolerated.
void DataCube::Load(std::string filename, bool debug) {
    Invalidate();
    coords_x.clear();
    coords_y.clear();
    coords_z.clear();
    slices.clear();

    std::ifstream in(filename);
    if (!in) throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");

    std::string line;
    std::vector<int> dims;
    std::vector<float> pending;  // coordinates, then values, not yet assigned
    std::vector<float>* axes[3] = {&coords_x, &coords_y, &coords_z};
    int axis = 0;
    int line_num = 0;
    bool have_geometry = false;
    slices.emplace_back();
    size_t k = 0;

    while (std::getline(in, line)) {
        ++line_num;
        if (line_num < 10) continue;
        if (line_num == 10) {
            dims = Helper::explode_int(line, ' ');
            if (dims.size() < 3 || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0)
                throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
            continue;
        }
        if (axis < 3) {  // still reading node coordinates
            append_floats(line, pending);
            if ((int)pending.size() == dims[axis]) {
                axes[axis]->swap(pending);
                pending.clear();
                ++axis;
            }
            continue;
        }
        if (!have_geometry) {
            InitSpacing();
            InitOriginCoords();
            InitSize();
            have_geometry = true;
        }
        if ((int)slices[k].size() == dims[1]) {  // slice complete: open the next one
            ++k;
            slices.emplace_back();
        }
        append_floats(line, pending);
        if ((int)pending.size() == dims[0]) {
            slices[k].push_back(pending);
            pending.clear();
        }
    }
    if (!have_geometry) throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");

    if (debug) {  // the reference's marker line for orientation checks (src/DataCube.cpp:127-133)
        for (int i = 0; i < 101; ++i) slices.at(80).at(79).at(i) = 1;
    }
}

void DataCube::InitSpacing() {
    spacing[0] = GetAxisSpacing(coords_x);
    spacing[1] = GetAxisSpacing(coords_y);
    spacing[2] = GetAxisSpacing(coords_z);
}

// first gap of the axis; like the reference, later gaps are not compared against it
float DataCube::GetAxisSpacing(std::vector<float> coords) {
    if (coords.size() < 2) throw std::invalid_argument("non-uniform grid given");
    return std::abs(coords[0] - coords[1]);
}

void DataCube::InitOriginCoords() {
    origin_coords[0] = coords_x[0];
    origin_coords[1] = coords_y[0];
    origin_coords[2] = coords_z[0];
}

void DataCube::InitSize() {
    size.x = (int)coords_x.size();
    size.y = (int)coords_y.size();
    size.z = (int)coords_z.size();
}

void DataCube::SetTrilinear(bool option) { trilinear = option; }

void DataCube::Invalidate() {
    if (device_cube) {
        birp_cube_free(device_cube);
        device_cube = nullptr;
    }
}

const birp_cube* DataCube::Resident() {
    if (device_cube) return device_cube;
    if (size.x <= 0 || size.y <= 0 || size.z <= 0 || coords_x.empty() || coords_y.empty() || coords_z.empty())
        throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
    // flatten slices[z][y][x]; .at() keeps the reference's out_of_range on a short cube
    std::vector<float> flat((size_t)size.x * size.y * size.z);
    size_t n = 0;
    for (int k = 0; k < size.z; ++k)
        for (int j = 0; j < size.y; ++j) {
            const std::vector<float>& row = slices.at(k).at(j);
            if ((int)row.size() < size.x) throw std::out_of_range("DataCube: short row in slices");
            for (int i = 0; i < size.x; ++i) flat[n++] = row[i];
        }
    const float origin[3] = {coords_x[0], coords_y[0], coords_z[0]};
    int rc = birp_cube_upload(flat.data(), size.x, size.y, size.z, origin, spacing, &device_cube);
    if (rc != BIRP_OK) throw std::runtime_error(std::string("DataCube upload: ") + birp_last_error());
    return device_cube;
}

float DataCube::GetSample(float x, float y, float z) {
    const float xyz[3] = {x, y, z};
    float out = 0.f;
    int rc = birp_sample_cube(Resident(), xyz, 1, trilinear ? BIRP_TRILINEAR : BIRP_NEAREST, &out);
    if (rc == BIRP_ERR_RANGE) throw std::out_of_range("vector::_M_range_check");
    if (rc != BIRP_OK) throw std::runtime_error(std::string("DataCube::GetSample: ") + birp_last_error());
    return out;
}
