// TEST INFRASTRUCTURE (oracle/): stand-in for the EleFits 4.0 headers, which are not
// installed in this image.  It exists only so that the UNMODIFIED reference
// src/Camera.cpp compiles into oracle/_ref/; the reference uses EleFits solely inside
// Camera::ToFITS (reference src/Camera.cpp:252-307), which the oracle never calls.
// No arithmetic on the hot path goes through these types.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Euclid {
namespace Cfitsio {
struct CfitsioError : std::runtime_error {
    CfitsioError() : std::runtime_error("cfitsio (shim)") {}
};
}  // namespace Cfitsio
namespace Fits {
enum class FileMode { Read, Edit, Create, Overwrite, Temporary };

template <typename T>
struct Record {
    std::string keyword;
    T value;
    std::string unit;
    std::string comment;
};

template <typename T>
struct ShimRaster {
    std::vector<T> data;
    long width;
    long height;
};

template <typename T>
ShimRaster<T> makeRaster(std::vector<T>&& data, long width, long height) {
    return ShimRaster<T>{std::move(data), width, height};
}

struct ShimHeader {
    template <typename... Rs>
    void writeSeq(const Rs&...) const {}
};

class SifFile {
public:
    SifFile(const std::string&, FileMode) {}
    ShimHeader header() const { return {}; }
    template <typename R, typename Raster>
    void write(const R&, const Raster&) const {}
};
}  // namespace Fits
}  // namespace Euclid
