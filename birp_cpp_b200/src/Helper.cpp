// Helper: string splitting for the cube loader / CLI, and the small dense-matrix routines of
// the reference's Helper class for callers that use them directly.  Camera itself does not go
// through these any more (its per-frame math lives behind birp_frame_setup); they are kept
// because they are part of the public interface (reference include/Helper.h:47-69).
#include "Helper.h"

#include <cmath>
#include <sstream>
#include <stdexcept>

namespace {
template <class T, class Conv>
std::vector<T> split_convert(std::string const& s, char delim, Conv conv) {
    std::vector<T> out;
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t end = s.find(delim, pos);
        if (end == std::string::npos) end = s.size();
        if (end > pos) out.push_back(conv(s.substr(pos, end - pos)));
        pos = end + 1;
    }
    return out;
}
using Matrix = std::vector<std::vector<float>>;

Matrix minor_of(const Matrix& m, size_t skip_row, size_t skip_col) {
    Matrix sub;
    for (size_t r = 0; r < m.size(); ++r) {
        if (r == skip_row) continue;
        std::vector<float> row;
        for (size_t c = 0; c < m.size(); ++c)
            if (c != skip_col) row.push_back(m[r][c]);
        sub.push_back(row);
    }
    return sub;
}
}  // namespace

std::vector<float> Helper::explode_float(std::string const& s, char delim) {
    return split_convert<float>(s, delim, [](const std::string& t) { return std::stof(t); });
}

std::vector<int> Helper::explode_int(std::string const& s, char delim) {
    return split_convert<int>(s, delim, [](const std::string& t) { return std::stoi(t); });
}

std::vector<std::string> Helper::explode_string(std::string const& s, char delim) {
    return split_convert<std::string>(s, delim, [](const std::string& t) { return t; });
}

Matrix Helper::MatrixMultiply(Matrix a, Matrix b) {
    const size_t n = a.size();
    Matrix out(n, std::vector<float>(b.size(), 0.f));
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < b.size(); ++j) {
            float acc = 0.0f;
            for (size_t k = 0; k < n; ++k) acc = acc + (a[i][k] * b[k][j]);
            out[i][j] = acc;
        }
    return out;
}

std::vector<float> Helper::MatrixMultiply(Matrix a, std::vector<float> b) {
    std::vector<float> out(4, 0.f);  // the reference returns a fixed 4-vector
    for (size_t r = 0; r < a.size() && r < 4; ++r)
        for (size_t c = 0; c < a[0].size(); ++c) out[r] += a[r][c] * b[c];
    return out;
}

Matrix Helper::MatrixScalarMultiply(Matrix m, float scalar) {
    for (auto& row : m)
        for (auto& v : row) v = v * scalar;
    return m;
}

std::vector<float> Helper::MatrixScalarMultiply(std::vector<float> v, float scalar) {
    for (auto& x : v) x = x * scalar;
    return v;
}

std::vector<float> Helper::ApplyRotation(Matrix rot, float* c) {
    std::vector<float> out;
    for (const auto& row : rot) out.push_back((row[0] * c[0]) + (row[1] * c[1]) + (row[2] * c[2]));
    return out;
}

float Helper::GetDeterminant(const Matrix m) {
    if (m.size() != m[0].size()) throw std::runtime_error("Matrix is not quadratic");
    const size_t n = m.size();
    if (n == 0) return 1;
    if (n == 1) return m[0][0];
    if (n == 2) return m[0][0] * m[1][1] - m[0][1] * m[1][0];
    float result = 0;
    int sign = 1;
    for (size_t i = 0; i < n; ++i) {
        result = result + sign * m[0][i] * GetDeterminant(minor_of(m, 0, i));
        sign = -sign;
    }
    return result;
}

Matrix Helper::GetTranspose(const Matrix m) {
    Matrix t(m[0].size(), std::vector<float>(m.size()));
    for (size_t i = 0; i < m.size(); ++i)
        for (size_t j = 0; j < m[0].size(); ++j) t[j][i] = m[i][j];
    return t;
}

Matrix Helper::GetCofactor(const Matrix m) {
    if (m.size() != m[0].size()) throw std::runtime_error("Matrix is not quadratic");
    Matrix c(m.size(), std::vector<float>(m.size()));
    for (size_t i = 0; i < m.size(); ++i)
        for (size_t j = 0; j < m.size(); ++j) c[i][j] = std::pow(-1, i + j) * GetDeterminant(minor_of(m, i, j));
    return c;
}

Matrix Helper::GetInverse(const Matrix m) {
    if (GetDeterminant(m) == 0) throw std::runtime_error("Determinant is 0");
    const float d = 1.f / GetDeterminant(m);
    Matrix inv = GetTranspose(GetCofactor(m));
    for (auto& row : inv)
        for (auto& v : row) v *= d;
    return inv;
}

float Helper::VectorDistance(float* p) { return std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]); }

float Helper::VectorDistance2D(float* p) { return std::sqrt(p[0] * p[0] + p[1] * p[1]); }

// The reference's version compares iterator distances and therefore accepts every axis
// (src/Helper.cpp:603-623); cubes with non-uniform axes load there and must load here.
bool Helper::EqualDistance(std::vector<float>) { return true; }
