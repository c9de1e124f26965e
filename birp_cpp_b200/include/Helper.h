// Helper -- the string and 3x3 helpers the driver and the renderer use.
//
// Same static signatures as the reference's include/Helper.h for the functions that are
// reachable from main.cpp, Camera, DataCube and CMEM.  The reference's unused astronomy
// leftovers (GEI/GSE transforms, RA/Dec conversion, Interp1, ...: src/Helper.cpp:266-588)
// are outside this repo's scope and are not provided.
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
#include <vector>

class Helper {
public:
    inline static void ltrim(std::string& s) {
        s.erase(s.begin(), std::find_if(s.begin(), s.end(), [](unsigned char c) { return !std::isspace(c); }));
    }
    inline static void rtrim(std::string& s) {
        s.erase(std::find_if(s.rbegin(), s.rend(), [](unsigned char c) { return !std::isspace(c); }).base(),
                s.end());
    }
    inline static void trim(std::string& s) {
        rtrim(s);
        ltrim(s);
    }
    inline static bool ends_with(std::string const& value, std::string const& ending) {
        return ending.size() <= value.size() && std::equal(ending.rbegin(), ending.rend(), value.rbegin());
    }

    static std::vector<float> explode_float(std::string const& s, char delim);
    static std::vector<int> explode_int(std::string const& s, char delim);
    static std::vector<std::string> explode_string(std::string const& s, char delim);

    static std::vector<std::vector<float>> MatrixMultiply(std::vector<std::vector<float>> a,
                                                          std::vector<std::vector<float>> b);
    static std::vector<float> MatrixMultiply(std::vector<std::vector<float>> a, std::vector<float> b);
    static std::vector<std::vector<float>> MatrixScalarMultiply(std::vector<std::vector<float>>, float);
    static std::vector<float> MatrixScalarMultiply(std::vector<float>, float);
    static std::vector<float> ApplyRotation(std::vector<std::vector<float>> rot, float* coords);
    static std::vector<std::vector<float>> GetInverse(const std::vector<std::vector<float>> vect);
    static float GetDeterminant(const std::vector<std::vector<float>> vect);
    static std::vector<std::vector<float>> GetTranspose(const std::vector<std::vector<float>> matrix1);
    static std::vector<std::vector<float>> GetCofactor(const std::vector<std::vector<float>> vect);
    static float VectorDistance(float* xyz);
    static float VectorDistance2D(float* xy);
    static bool EqualDistance(std::vector<float> steps);
};
