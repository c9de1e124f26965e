// DataCube -- an MHD emissivity grid, resident on the GPU while it is being rendered.
//
// Public surface = the reference's include/DataCube.h:22-50 (same member names, so src/main.cpp and
// tests/DataCube_test.cpp compile unchanged).  Load() parses the reference's ASCII layout straight into ONE flat
// buffer (x fastest, the layout the GPU wants); the nested `slices[z][y][x]` of the reference is only built when a
// caller touches it -- from then on it is the truth and the next upload re-flattens it -- so a 512^3 cube is held
// once on the host, not twice.  Camera::Render uploads the grid once through birp_cube_upload and reuses the device
// copy until Load()/Invalidate() marks it stale.
#pragma once
#include <cstddef>
#include <iostream>
#include <string>
#include <vector>

#include "Space.h"

struct birp_cube;  // include/birp_b200.h
class DataCube;

// `DataCube::slices`: behaves like the reference's std::vector<std::vector<std::vector<float>>> (indexing, at(),
// size(), iteration, assign / push_back / clear, conversion to a reference to the nested vectors); the nested
// vectors are materialised from the cube's flat buffer at the first access.  Any non-const access may modify them,
// so it makes them the truth (the flat buffer is dropped and rebuilt at the next upload).
class LazySlices {
public:
    using Row = std::vector<float>;
    using Slice = std::vector<Row>;
    using Nested = std::vector<Slice>;

    Nested& get();
    const Nested& get() const;
    operator Nested&() { return get(); }
    operator const Nested&() const { return get(); }

    Slice& operator[](std::size_t k) { return get()[k]; }
    const Slice& operator[](std::size_t k) const { return get()[k]; }
    Slice& at(std::size_t k) { return get().at(k); }
    const Slice& at(std::size_t k) const { return get().at(k); }
    Slice& back() { return get().back(); }
    Slice& front() { return get().front(); }
    std::size_t size() const { return get().size(); }
    bool empty() const { return get().empty(); }
    Nested::iterator begin() { return get().begin(); }
    Nested::iterator end() { return get().end(); }
    Nested::const_iterator begin() const { return get().begin(); }
    Nested::const_iterator end() const { return get().end(); }
    void clear();
    void push_back(const Slice& s) { get().push_back(s); }
    void push_back(Slice&& s) { get().push_back(std::move(s)); }
    template <class... A>
    Slice& emplace_back(A&&... a) {
        return get().emplace_back(std::forward<A>(a)...);
    }
    template <class... A>
    void assign(A&&... a) {
        clear();
        nested_.assign(std::forward<A>(a)...);
    }
    void resize(std::size_t n) { get().resize(n); }
    LazySlices& operator=(const Nested& n) {
        clear();
        nested_ = n;
        return *this;
    }
    LazySlices& operator=(Nested&& n) {
        clear();
        nested_ = std::move(n);
        return *this;
    }

private:
    friend class DataCube;
    explicit LazySlices(DataCube* owner) : owner_(owner) {}
    void materialise() const;
    DataCube* owner_;
    mutable Nested nested_;
};

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

    LazySlices slices;  // slices[z][y][x], like the reference's nested vectors (include/DataCube.h:42)

    // ---- additions (not in the reference) -----------------------------------------------
    bool IsTrilinear() const { return trilinear; }
    // Tolerance mode of the trilinear march (BIRP_FAST_MARCH): sample values with fused arithmetic, ~1e-6 relative per
    // pixel against the reference instead of bit-identical images, 1.5x - 1.9x the frame rate; where rays end stays
    // exact.  Off by default.
    void SetFastMarch(bool on) { fast_march = on; }
    bool FastMarch() const { return fast_march; }
    // Load with an explicit number of text -> float worker threads (0 = choose: up to 16, or BIRP_LOAD_THREADS);
    // callers that load several cubes at once (RenderBatchDirectory) divide the cores among them.
    void LoadParallel(const std::string& filename, bool debug, int nthreads);
    // Binary side-car next to the ASCII file ("<file>.birpcache": dimensions, node coordinates, the flat fp32 grid,
    // keyed by the size and modification time of the ASCII file).  Load() always uses a valid side-car instead of
    // parsing; it WRITES one after parsing only when this is switched on (or BIRP_CUBE_CACHE=1 is in the
    // environment): a library should not drop files next to the user's data unasked.
    static void SetCacheWrites(bool on);
    static bool CacheWrites();
    // Keep all side-cars in one directory ("<dir>/<file name>.<hash of the path>.birpcache") instead of next to the
    // cubes; "" = next to the cubes.  Also BIRP_CUBE_CACHE_DIR in the environment.
    static void SetCacheDirectory(const std::string& dir);
    // true when the last Load() was served by a side-car
    bool LoadedFromCache() const { return loaded_from_cache; }
    // Many GetSample calls at once (one upload of the points, one kernel, one download); xyz: n x {x, y, z}.
    // GetSample() itself does a full host <-> device round trip per call: fine for a probe, hostile to a loop.
    std::vector<float> GetSamples(const std::vector<float>& xyz);
    // Call after editing `slices`, `coords_*`, `spacing` or `size` by hand once the cube has
    // been rendered or sampled; Load() does it itself.
    void Invalidate();
    // Device-resident copy (uploads on first use).  Throws std::runtime_error on CUDA failure.
    const birp_cube* Resident();
    // The same grid on another GPU of the box (copied device to device from the first resident copy); the copies
    // live until Load() / Invalidate().  The calling thread's current device is left as it was.
    const birp_cube* ResidentOn(int device);
    // true while the nested `slices` have not been built (nobody touched them since Load)
    bool SlicesAreLazy() const { return flat_valid; }

private:
    friend class LazySlices;
    void InitSpacing();
    void InitOriginCoords();
    void InitSize();
    float GetAxisSpacing(std::vector<float> coords);
    bool TryLoadCache(const std::string& filename, long long src_size, long long src_mtime_ns);
    void WriteCache(const std::string& filename, long long src_size, long long src_mtime_ns) const;
    bool trilinear = false;  // the reference leaves this uninitialised until SetTrilinear
    bool fast_march = false;
    birp_cube* device_cube = nullptr;
    std::vector<birp_cube*> replicas;  // ResidentOn: copies on other devices
    // the grid as Load() parsed it: flat[(z * ny + y) * nx + x]; valid until somebody touches `slices`
    std::vector<float> flat;
    bool flat_valid = false;
    int extra_slices = 0;  // empty trailing slices the reference's state machine leaves behind
    bool loaded_from_cache = false;
};
