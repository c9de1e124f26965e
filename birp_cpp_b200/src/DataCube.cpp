// DataCube: reference-compatible ASCII loader and GPU residency.
#include "DataCube.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/birp_b200.h"
#include "Helper.h"

namespace {
// whitespace-separated floats of one line [p, end), appended to `out` (what explode_float(line, ' ') yields)
void append_floats(const char* p, const char* end, std::vector<float>& out) {
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

// One worker's share of the value section: the floats of its lines back to back, and how many each line held.
struct ParsedLines {
    std::vector<float> values;
    std::vector<unsigned> per_line;
    std::exception_ptr error;
};

// lines of [p, end) the way std::getline splits them ('\n' ends a line; a last line without '\n' counts)
void parse_lines(const char* p, const char* end, ParsedLines& result) {
    // (filled locally and handed over at the end: the workers' result slots sit next to each other in memory,
    // and the vectors' end pointers change with every value)
    ParsedLines out;
    out.values.reserve((size_t)(end - p) / 8);
    try {
        while (p < end) {
            const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
            const char* le = nl ? nl : end;
            const size_t before = out.values.size();
            append_floats(p, le, out.values);
            out.per_line.push_back((unsigned)(out.values.size() - before));
            p = nl ? nl + 1 : end;
        }
    } catch (...) {
        out.error = std::current_exception();
    }
    result = std::move(out);
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
//
// The reference parses line by line with a small state machine (a new slice is opened when a line
// arrives and the current slice is full -- hence its extra, empty trailing slice; a thread is
// stored when the pending values number exactly nx).  That state machine is replayed here verbatim,
// but over PRE-PARSED lines: the text -> float conversion, which is all the time, runs on worker
// threads over line-aligned pieces of the value section.
void DataCube::Load(std::string filename, bool debug) {
    Invalidate();
    coords_x.clear();
    coords_y.clear();
    coords_z.clear();
    slices.clear();

    // the file, mapped read-only (no copy, no zero-fill of a 60 MB buffer)
    struct Mapping {
        const char* data = nullptr;
        size_t size = 0;
        ~Mapping() {
            if (data && size) munmap(const_cast<char*>(data), size);
        }
    } map;
    {
        const int fd = open(filename.c_str(), O_RDONLY);
        if (fd < 0) throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            close(fd);
            throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
        }
        map.size = (size_t)st.st_size;
        if (map.size) {
            void* m = mmap(nullptr, map.size, PROT_READ, MAP_PRIVATE, fd, 0);
            close(fd);
            if (m == MAP_FAILED) {
                map.size = 0;
                throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
            }
            map.data = static_cast<const char*>(m);
        } else {
            close(fd);
        }
    }
    const char* p = map.data;
    const char* const end = p + map.size;

    // ---- header and node coordinates: serial, a few lines
    std::vector<int> dims;
    std::vector<float> pending;  // coordinates, then values, not yet assigned
    std::vector<float>* axes[3] = {&coords_x, &coords_y, &coords_z};
    int axis = 0;
    int line_num = 0;
    while (p < end && axis < 3) {
        const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
        const char* le = nl ? nl : end;
        ++line_num;
        if (line_num == 10) {
            dims = Helper::explode_int(std::string(p, le), ' ');
            if (dims.size() < 3 || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0)
                throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
        } else if (line_num > 10) {
            append_floats(p, le, pending);
            if ((int)pending.size() == dims[axis]) {
                axes[axis]->swap(pending);
                pending.clear();
                ++axis;
            }
        }
        p = nl ? nl + 1 : end;
    }
    // the reference derives the geometry when the first value line arrives
    if (axis < 3 || p >= end) throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
    InitSpacing();
    InitOriginCoords();
    InitSize();

    // ---- value section: text -> float on worker threads, over pieces that start and end at line breaks
    const size_t bytes = (size_t)(end - p);
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 4;
    if (nthreads > 16) nthreads = 16;
    if (bytes < (size_t(1) << 20)) nthreads = 1;
    if (const char* env = std::getenv("BIRP_LOAD_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(env));
    std::vector<const char*> cut(nthreads + 1, end);
    cut[0] = p;
    for (unsigned t = 1; t < nthreads; ++t) {
        const char* guess = p + bytes / nthreads * t;
        if (guess < cut[t - 1]) guess = cut[t - 1];
        const char* nl = static_cast<const char*>(std::memchr(guess, '\n', (size_t)(end - guess)));
        cut[t] = nl ? nl + 1 : end;
    }
    std::vector<ParsedLines> parts(nthreads);
    if (nthreads == 1) {
        parse_lines(cut[0], cut[1], parts[0]);
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(parse_lines, cut[t], cut[t + 1], std::ref(parts[t]));
        for (auto& th : pool) th.join();
    }
    for (auto& part : parts)
        if (part.error) std::rethrow_exception(part.error);  // (std::stof on a malformed token, like the reference)

    // ---- the reference's per-line state machine (src/DataCube.cpp:96-121) over the parsed lines
    slices.emplace_back();
    size_t k = 0;
    for (const ParsedLines& part : parts) {
        const float* v = part.values.data();
        for (unsigned count : part.per_line) {
            if ((int)slices[k].size() == dims[1]) {  // slice complete: open the next one
                ++k;
                slices.emplace_back();
            }
            pending.insert(pending.end(), v, v + count);
            v += count;
            if ((int)pending.size() == dims[0]) {
                slices[k].push_back(pending);
                pending.clear();
            }
        }
    }

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
