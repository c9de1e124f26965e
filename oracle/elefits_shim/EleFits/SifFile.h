// TEST INFRASTRUCTURE (oracle/): stand-in for the EleFits 4.0 headers, which are not
// installed in this image.  It exists only so that the UNMODIFIED reference
// src/Camera.cpp compiles into oracle/_ref/; the reference uses EleFits solely inside
// Camera::ToFITS (reference src/Camera.cpp:252-307).  No arithmetic on the hot path goes
// through these types.  SifFile writes a minimal single-image FITS file (primary HDU,
// BITPIX -32) so that the reference CLI built in oracle/_ref/ produces output files the
// tests can compare pixel for pixel with the product's.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Euclid {
namespace Cfitsio {
struct CfitsioError : std::runtime_error {
    CfitsioError() : std::runtime_error("cfitsio (shim)") {}
    explicit CfitsioError(const std::string& m) : std::runtime_error(m) {}
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

namespace shim_detail {
inline std::string card(const std::string& key, const std::string& value, const std::string& comment) {
    char buf[96];
    std::string v = value;
    if (!comment.empty()) v += " / " + comment;
    std::string k = key;
    for (auto& c : k) c = (char)std::toupper((unsigned char)c);
    std::snprintf(buf, sizeof(buf), "%-8.8s= %-70.70s", k.c_str(), v.c_str());
    return std::string(buf, 80);
}
// a TFLOAT keyword value as cfitsio formats it (putkey.c: ffpkye, decim = -7 -> ffr2e): "%.7G", an exponent form
// without a decimal point redone as "%.1E", a '.' appended to an integer-looking result; right-justified to column 30
inline std::string value_of(float x) {
    char v[40];
    std::snprintf(v, sizeof(v), "%.7G", (double)x);
    if (!std::strchr(v, '.') && std::strchr(v, 'E')) std::snprintf(v, sizeof(v), "%.1E", (double)x);
    if (!std::strchr(v, '.') && !std::strchr(v, 'E') && !std::strchr(v, 'N') && !std::strchr(v, 'I')) std::strcat(v, ".");
    char buf[48];
    std::snprintf(buf, sizeof(buf), "%20s", v);
    return buf;
}
// a string value (ffpkys): at least eight characters inside the quotes, padded to column 30 like a number
inline std::string value_of(const std::string& s) {
    std::string q = "'" + s;
    while (q.size() < 9) q += ' ';
    q += "'";
    while (q.size() < 20) q += ' ';
    return q;
}
inline std::string value_of(long x) {
    char buf[40];
    std::snprintf(buf, sizeof(buf), "%20ld", x);
    return buf;
}
template <typename T>
std::string record_card(const Record<T>& r) {
    std::string c = r.unit.empty() ? r.comment : "[" + r.unit + "] " + r.comment;
    while (!c.empty() && c.back() == ' ') c.pop_back();
    return card(r.keyword, value_of(r.value), c);
}
}  // namespace shim_detail

class SifFile;
struct ShimHeader {
    SifFile* file;
    template <typename... Rs>
    void writeSeq(const Rs&... records) const;
};

class SifFile {
public:
    SifFile(const std::string& path, FileMode mode) : path_(path) {
        if (mode == FileMode::Create) {
            std::ifstream probe(path);
            if (probe.good()) throw Cfitsio::CfitsioError("file exists: " + path);
        }
    }
    ShimHeader header() { return ShimHeader{this}; }
    template <typename R, typename Raster>
    void write(const R& record, const Raster& raster) {
        using namespace shim_detail;
        // the mandatory keywords and the two COMMENT cards of cfitsio's fits_create_img
        std::string hdr = card("SIMPLE", "                   T", "file does conform to FITS standard");
        hdr += card("BITPIX", value_of(-32L), "number of bits per data pixel");
        hdr += card("NAXIS", value_of(2L), "number of data axes");
        hdr += card("NAXIS1", value_of((long)raster.width), "length of data axis 1");
        hdr += card("NAXIS2", value_of((long)raster.height), "length of data axis 2");
        hdr += card("EXTEND", "                   T", "FITS dataset may contain extensions");
        {
            char buf[96];
            std::snprintf(buf, sizeof(buf), "COMMENT   %-70.70s",
                          "FITS (Flexible Image Transport System) format is defined in 'Astronomy");
            hdr += std::string(buf, 80);
            std::snprintf(buf, sizeof(buf), "COMMENT   %-70.70s",
                          "and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H");
            hdr += std::string(buf, 80);
        }
        for (const auto& c : cards_) hdr += c;
        hdr += record_card(record);
        hdr += "END";
        hdr.append((2880 - hdr.size() % 2880) % 2880, ' ');
        std::string data(raster.data.size() * 4, '\0');
        size_t n = 0;
        for (float v : raster.data) {
            uint32_t bits;
            std::memcpy(&bits, &v, 4);
            data[n++] = (char)(bits >> 24);
            data[n++] = (char)(bits >> 16);
            data[n++] = (char)(bits >> 8);
            data[n++] = (char)bits;
        }
        data.append((2880 - data.size() % 2880) % 2880, '\0');
        std::ofstream out(path_, std::ios::binary | std::ios::trunc);
        out.write(hdr.data(), (std::streamsize)hdr.size());
        out.write(data.data(), (std::streamsize)data.size());
    }
    void add_card(const std::string& c) { cards_.push_back(c); }

private:
    std::string path_;
    std::vector<std::string> cards_;
};

template <typename... Rs>
void ShimHeader::writeSeq(const Rs&... records) const {
    (file->add_card(shim_detail::record_card(records)), ...);
}
}  // namespace Fits
}  // namespace Euclid
