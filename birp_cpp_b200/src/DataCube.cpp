// DataCube: reference-compatible ASCII loader (straight into one flat buffer), binary side-car, GPU residency.
#include "DataCube.h"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <functional>
#include <mutex>
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
        if (res.ec != std::errc() || res.ptr != q) v = std::stof(std::string(p, q));  // "+1.0", "nan", "1.5\r", ...
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

// lines of [p, end) the way std::getline splits them ('\n' ends a line; a last line without '\n' counts).
// Every VALUE line goes through Helper::trim in the reference (src/DataCube.cpp:99: isspace at both ends) before it
// is split at ' ': CRLF files, tab- or space-terminated lines and "\r"-only separator lines load there, so the same
// trim happens here.  (The coordinate lines are NOT trimmed by the reference, nor by the caller of append_floats
// for that section: a whitespace-only token there throws from std::stof in both.)
void parse_lines(const char* p, const char* end, ParsedLines& result) {
    // (filled locally and handed over at the end: the workers' result slots sit next to each other in memory,
    // and the vectors' end pointers change with every value)
    ParsedLines out;
    out.values.reserve((size_t)(end - p) / 8);
    try {
        while (p < end) {
            const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
            const char* ls = p;
            const char* le = nl ? nl : end;
            while (ls < le && std::isspace((unsigned char)*ls)) ++ls;
            while (le > ls && std::isspace((unsigned char)le[-1])) --le;
            const size_t before = out.values.size();
            append_floats(ls, le, out.values);
            out.per_line.push_back((unsigned)(out.values.size() - before));
            p = nl ? nl + 1 : end;
        }
    } catch (...) {
        out.error = std::current_exception();
    }
    result = std::move(out);
}

std::atomic<int> g_cache_writes{-1};  // -1: not decided yet (environment), 0 / 1
std::mutex g_cache_dir_mu;
std::string g_cache_dir;
bool g_cache_dir_set = false;

// where the side-car of `filename` lives: next to it, or -- DataCube::SetCacheDirectory / BIRP_CUBE_CACHE_DIR -- in one
// directory for all cubes, named after the file and a hash of its path (a directory of cubes that other tools
// iterate over, like the reference CLI's batch mode, stays free of foreign files)
std::string cache_path(const std::string& filename) {
    std::string dir;
    {
        std::lock_guard<std::mutex> lk(g_cache_dir_mu);
        if (!g_cache_dir_set) {
            if (const char* env = std::getenv("BIRP_CUBE_CACHE_DIR")) g_cache_dir = env;
            g_cache_dir_set = true;
        }
        dir = g_cache_dir;
    }
    if (dir.empty()) return filename + ".birpcache";
    std::uint64_t h = 1469598103934665603ull;  // FNV-1a of the path as given
    for (unsigned char ch : filename) h = (h ^ ch) * 1099511628211ull;
    char hex[20];
    std::snprintf(hex, sizeof(hex), "%016llx", (unsigned long long)h);
    const std::string base = filename.substr(filename.find_last_of("/\\") + 1);
    if (dir.back() != '/') dir += '/';
    return dir + base + "." + hex + ".birpcache";
}

// ---- side-car layout: header, node coordinates, flat grid; all little-endian fp32 / integers of this machine
struct CacheHeader {
    char magic[8];  // "BIRPCUB1"
    std::int64_t src_size;
    std::int64_t src_mtime_ns;
    std::int32_t nx, ny, nz;
    std::int32_t extra_slices;
};
constexpr char kCacheMagic[8] = {'B', 'I', 'R', 'P', 'C', 'U', 'B', '1'};
}  // namespace

// ------------------------------------------------------------------ LazySlices
void LazySlices::materialise() const {
    DataCube* c = owner_;
    if (!c->flat_valid) return;  // nested_ already is the truth
    const int nx = c->size.x, ny = c->size.y, nz = c->size.z;
    nested_.assign((size_t)nz + (size_t)c->extra_slices, Slice());
    const float* v = c->flat.data();
    for (int k = 0; k < nz; ++k) {
        Slice& s = nested_[k];
        s.resize(ny);
        for (int j = 0; j < ny; ++j, v += nx) s[j].assign(v, v + nx);
    }
}

LazySlices::Nested& LazySlices::get() {
    if (owner_->flat_valid) {
        materialise();
        // the caller may write through this reference: the nested vectors are the truth from now on
        owner_->flat_valid = false;
        std::vector<float>().swap(owner_->flat);
    }
    return nested_;
}

const LazySlices::Nested& LazySlices::get() const {
    if (owner_->flat_valid && nested_.empty()) materialise();  // a read keeps the flat buffer valid
    return nested_;
}

void LazySlices::clear() {
    nested_.clear();
    owner_->flat_valid = false;
    std::vector<float>().swap(owner_->flat);
}

// ------------------------------------------------------------------ DataCube
DataCube::DataCube() : size{0, 0, 0}, spacing{0.f, 0.f, 0.f}, origin_coords{0.f, 0.f, 0.f}, slices(this) {}

DataCube::~DataCube() { Invalidate(); }

void DataCube::Init() {}

void DataCube::Init(std::string& filename, bool debug = false) { this->Load(filename, debug); }

void DataCube::SetCacheWrites(bool on) { g_cache_writes.store(on ? 1 : 0); }

void DataCube::SetCacheDirectory(const std::string& dir) {
    std::lock_guard<std::mutex> lk(g_cache_dir_mu);
    g_cache_dir = dir;
    g_cache_dir_set = true;
}

bool DataCube::CacheWrites() {
    int v = g_cache_writes.load();
    if (v < 0) {
        const char* env = std::getenv("BIRP_CUBE_CACHE");
        v = (env && std::atoi(env) != 0) ? 1 : 0;
        g_cache_writes.store(v);
    }
    return v != 0;
}

void DataCube::Load(std::string filename, bool debug) { LoadParallel(filename, debug, 0); }

// Layout (reference src/DataCube.cpp:28-138): line 1 is a free-text header, line 10 holds
// "nx ny nz", the following lines list the x, then y, then z node coordinates (each axis may
// wrap over several lines), and the rest are the values, x fastest: every run of nx values is
// one "thread", ny threads make a slice.  Blank lines between slices are tolerated.
//
// The reference parses line by line with a small state machine (a new slice is opened when a line
// arrives and the current slice is full -- hence its extra, empty trailing slice; a thread is
// stored when the pending values number exactly nx).  The text -> float conversion, which is all the
// time, runs on worker threads over line-aligned pieces of the value section; the state
// machine is then replayed over the per-line value COUNTS alone.  When that replay shows a regular file --
// nz full slices of ny rows of nx values, nothing left over -- the workers' values, concatenated, ARE the grid
// x fastest: they are copied once into the flat buffer and the nested vectors are not built at all.  Any other
// file (short rows, lines that overshoot a row, missing slices) is replayed value by value into the nested
// vectors exactly as the reference would leave them.
void DataCube::LoadParallel(const std::string& filename, bool debug, int nthreads_wanted) {
    Invalidate();
    coords_x.clear();
    coords_y.clear();
    coords_z.clear();
    slices.clear();
    extra_slices = 0;
    loaded_from_cache = false;

    // the file, mapped read-only (no copy, no zero-fill of a 60 MB buffer)
    struct Mapping {
        const char* data = nullptr;
        size_t size = 0;
        ~Mapping() {
            if (data && size) munmap(const_cast<char*>(data), size);
        }
    } map;
    long long src_mtime_ns = 0;
    {
        const int fd = open(filename.c_str(), O_RDONLY);
        if (fd < 0) throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            close(fd);
            throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
        }
        map.size = (size_t)st.st_size;
        src_mtime_ns = (long long)st.st_mtim.tv_sec * 1000000000LL + (long long)st.st_mtim.tv_nsec;
        if (!debug && TryLoadCache(filename, (long long)map.size, src_mtime_ns)) {
            close(fd);
            map.size = 0;
            return;
        }
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
    if (const char* env = std::getenv("BIRP_LOAD_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(env));
    else if (nthreads_wanted > 0) nthreads = (unsigned)nthreads_wanted;
    else if (bytes < (size_t(1) << 20)) nthreads = 1;
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

    // ---- the reference's per-line state machine (src/DataCube.cpp:96-121) over the per-line value COUNTS
    const size_t nx = (size_t)dims[0], ny = (size_t)dims[1], nz = (size_t)dims[2];
    bool regular = true;
    {
        size_t k = 0, rows_in_slice = 0, npending = 0, total = 0;
        for (const ParsedLines& part : parts) {
            for (unsigned count : part.per_line) {
                if (rows_in_slice == ny) {  // slice complete: the arrival of any line opens the next one
                    ++k;
                    rows_in_slice = 0;
                }
                npending += count;
                total += count;
                if (npending == nx) {
                    ++rows_in_slice;
                    npending = 0;
                } else if (npending > nx) {
                    regular = false;  // a row that overshoots never completes in the reference
                }
            }
        }
        // regular: exactly nz full slices (k counts the slices opened after the first), nothing pending; lines after
        // the last full slice only open empty trailing slices
        const size_t full = k + (rows_in_slice == ny ? 1 : 0);
        regular = regular && npending == 0 && total == nx * ny * nz && full >= nz &&
                  (rows_in_slice == ny || rows_in_slice == 0);
        if (regular) extra_slices = (int)(k + 1 - nz);
    }

    if (regular && !debug) {
        flat.resize(nx * ny * nz);
        size_t off = 0;
        std::vector<size_t> offs(parts.size());
        for (size_t t = 0; t < parts.size(); ++t) {
            offs[t] = off;
            off += parts[t].values.size();
        }
        if (parts.size() == 1) {
            std::memcpy(flat.data(), parts[0].values.data(), parts[0].values.size() * sizeof(float));
        } else {
            std::vector<std::thread> pool;
            for (size_t t = 0; t < parts.size(); ++t)
                pool.emplace_back([&, t] {
                    std::memcpy(flat.data() + offs[t], parts[t].values.data(), parts[t].values.size() * sizeof(float));
                    std::vector<float>().swap(parts[t].values);
                });
            for (auto& th : pool) th.join();
        }
        flat_valid = true;
        if (CacheWrites()) WriteCache(filename, (long long)map.size, src_mtime_ns);
        return;
    }

    // ---- irregular file (or the debug marker): the nested vectors, value by value, as the reference leaves them
    LazySlices::Nested& nested = slices.nested_;
    nested.emplace_back();
    size_t k = 0;
    for (const ParsedLines& part : parts) {
        const float* v = part.values.data();
        for (unsigned count : part.per_line) {
            if (nested[k].size() == ny) {  // slice complete: open the next one
                ++k;
                nested.emplace_back();
            }
            pending.insert(pending.end(), v, v + count);
            v += count;
            if (pending.size() == nx) {
                nested[k].push_back(pending);
                pending.clear();
            }
        }
    }

    if (debug) {  // the reference's marker line for orientation checks (src/DataCube.cpp:127-133)
        for (int i = 0; i < 101; ++i) nested.at(80).at(79).at(i) = 1;
    }
}

bool DataCube::TryLoadCache(const std::string& filename, long long src_size, long long src_mtime_ns) {
    const std::string path = cache_path(filename);
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    bool ok = false;
    CacheHeader h;
    struct stat st;
    do {
        if (fstat(fd, &st) != 0) break;
        if (read(fd, &h, sizeof(h)) != (ssize_t)sizeof(h)) break;
        if (std::memcmp(h.magic, kCacheMagic, 8) != 0) break;
        if (h.src_size != src_size || h.src_mtime_ns != src_mtime_ns) break;  // the ASCII file changed: stale
        if (h.nx <= 0 || h.ny <= 0 || h.nz <= 0 || h.extra_slices < 0) break;
        const size_t voxels = (size_t)h.nx * h.ny * h.nz;
        const size_t want = sizeof(h) + ((size_t)h.nx + h.ny + h.nz + voxels) * sizeof(float);
        if ((size_t)st.st_size != want) break;
        std::vector<float> cx(h.nx), cy(h.ny), cz(h.nz), grid(voxels);
        auto read_all = [&](void* dst, size_t n) {
            char* d = static_cast<char*>(dst);
            while (n) {
                const ssize_t r = read(fd, d, n);
                if (r <= 0) return false;
                d += r;
                n -= (size_t)r;
            }
            return true;
        };
        if (!read_all(cx.data(), cx.size() * 4) || !read_all(cy.data(), cy.size() * 4) || !read_all(cz.data(), cz.size() * 4) ||
            !read_all(grid.data(), voxels * 4))
            break;
        coords_x.swap(cx);
        coords_y.swap(cy);
        coords_z.swap(cz);
        flat.swap(grid);
        extra_slices = h.extra_slices;
        ok = true;
    } while (false);
    close(fd);
    if (!ok) return false;
    InitSpacing();
    InitOriginCoords();
    InitSize();
    flat_valid = true;
    loaded_from_cache = true;
    return true;
}

void DataCube::WriteCache(const std::string& filename, long long src_size, long long src_mtime_ns) const {
    // best effort: written under a temporary name and renamed into place, so that a reader never sees half a file
    const std::string path = cache_path(filename);
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    std::FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return;
    CacheHeader h;
    std::memcpy(h.magic, kCacheMagic, 8);
    h.src_size = src_size;
    h.src_mtime_ns = src_mtime_ns;
    h.nx = size.x, h.ny = size.y, h.nz = size.z;
    h.extra_slices = extra_slices;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && std::fwrite(coords_x.data(), 4, coords_x.size(), f) == coords_x.size();
    ok = ok && std::fwrite(coords_y.data(), 4, coords_y.size(), f) == coords_y.size();
    ok = ok && std::fwrite(coords_z.data(), 4, coords_z.size(), f) == coords_z.size();
    ok = ok && std::fwrite(flat.data(), 4, flat.size(), f) == flat.size();
    ok = (std::fclose(f) == 0) && ok;
    if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
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
    for (birp_cube* r : replicas) birp_cube_free(r);
    replicas.clear();
    if (device_cube) {
        birp_cube_free(device_cube);
        device_cube = nullptr;
    }
}

const birp_cube* DataCube::ResidentOn(int device) {
    const birp_cube* first = Resident();
    int dev = -1;
    if (birp_cube_device(first, &dev) == BIRP_OK && dev == device) return first;
    for (birp_cube* r : replicas)
        if (birp_cube_device(r, &dev) == BIRP_OK && dev == device) return r;
    birp_cube* copy = nullptr;
    int rc = birp_cube_replicate(first, device, &copy);
    if (rc != BIRP_OK) throw std::runtime_error(std::string("DataCube replication: ") + birp_last_error());
    replicas.push_back(copy);
    return copy;
}

const birp_cube* DataCube::Resident() {
    if (device_cube) return device_cube;
    if (size.x <= 0 || size.y <= 0 || size.z <= 0 || coords_x.empty() || coords_y.empty() || coords_z.empty())
        throw std::invalid_argument("Datacube is invalid, inaccessible, or malformed.");
    const float origin[3] = {coords_x[0], coords_y[0], coords_z[0]};
    const size_t voxels = (size_t)size.x * size.y * size.z;
    int rc;
    if (flat_valid && flat.size() == voxels) {
        // nobody touched `slices` since Load(): the parsed buffer goes to the device as it is
        rc = birp_cube_upload(flat.data(), size.x, size.y, size.z, origin, spacing, &device_cube);
    } else {
        // flatten slices[z][y][x]; .at() keeps the reference's out_of_range on a short cube
        const LazySlices::Nested& nested = slices.nested_;
        std::vector<float> tmp(voxels);
        size_t n = 0;
        for (int k = 0; k < size.z; ++k)
            for (int j = 0; j < size.y; ++j) {
                const std::vector<float>& row = nested.at(k).at(j);
                if ((int)row.size() < size.x) throw std::out_of_range("DataCube: short row in slices");
                for (int i = 0; i < size.x; ++i) tmp[n++] = row[i];
            }
        rc = birp_cube_upload(tmp.data(), size.x, size.y, size.z, origin, spacing, &device_cube);
    }
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

std::vector<float> DataCube::GetSamples(const std::vector<float>& xyz) {
    if (xyz.size() % 3 != 0) throw std::invalid_argument("DataCube::GetSamples: points must be x, y, z triples");
    std::vector<float> out(xyz.size() / 3);
    if (out.empty()) return out;
    int rc = birp_sample_cube(Resident(), xyz.data(), (long)out.size(), trilinear ? BIRP_TRILINEAR : BIRP_NEAREST,
                              out.data());
    if (rc == BIRP_ERR_RANGE) throw std::out_of_range("vector::_M_range_check");
    if (rc != BIRP_OK) throw std::runtime_error(std::string("DataCube::GetSamples: ") + birp_last_error());
    return out;
}
