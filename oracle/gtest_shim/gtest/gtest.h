// TEST INFRASTRUCTURE (oracle/): a few-line stand-in for GoogleTest, which is not installed in
// this image, so that the reference's own test files (tests/CMEM_test.cpp, tests/Space_test.cpp)
// compile UNCHANGED.  Supports exactly what those files use: TEST, EXPECT_EQ, GTEST_ASSERT_EQ, ASSERT_TRUE,
// testing::InitGoogleTest, RUN_ALL_TESTS.  EXPECT_EQ on floats is an exact comparison, as in gtest.
#pragma once
#include <cstdio>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
struct ShimCase {
    std::string name;
    std::function<void(int&)> body;
};
inline std::vector<ShimCase>& shim_cases() {
    static std::vector<ShimCase> cases;
    return cases;
}
struct ShimRegistrar {
    ShimRegistrar(const char* suite, const char* name, std::function<void(int&)> body) {
        shim_cases().push_back({std::string(suite) + "." + name, std::move(body)});
    }
};
inline void InitGoogleTest() {}
inline void InitGoogleTest(int*, char**) {}
inline int shim_run_all() {
    int failed = 0;
    for (auto& c : shim_cases()) {
        int failures = 0;
        std::printf("[ RUN      ] %s\n", c.name.c_str());
        c.body(failures);
        std::printf(failures ? "[  FAILED  ] %s\n" : "[       OK ] %s\n", c.name.c_str());
        failed += failures ? 1 : 0;
    }
    std::printf("[==========] %zu tests ran, %d failed\n", shim_cases().size(), failed);
    return failed ? 1 : 0;
}
template <class A, class B>
bool shim_eq(const A& a, const B& b, const char* ea, const char* eb, const char* file, int line, int& failures) {
    if (a == b) return true;
    std::ostringstream os;
    os.precision(9);
    os << file << ":" << line << ": Failure\nExpected equality of these values:\n  " << ea << "\n    Which is: " << a
       << "\n  " << eb << "\n    Which is: " << b << "\n";
    std::fputs(os.str().c_str(), stdout);
    ++failures;
    return false;
}
}  // namespace testing

#define TEST(suite, name)                                                                         \
    static void shim_##suite##_##name(int& shim_failures);                                        \
    static ::testing::ShimRegistrar shim_reg_##suite##_##name(#suite, #name, shim_##suite##_##name); \
    static void shim_##suite##_##name(int& shim_failures)
#define EXPECT_EQ(a, b) ::testing::shim_eq((a), (b), #a, #b, __FILE__, __LINE__, shim_failures)
#define GTEST_ASSERT_EQ(a, b) EXPECT_EQ(a, b)
#define ASSERT_TRUE(c)                                                           \
    do {                                                                         \
        if (!(c)) {                                                              \
            std::printf("%s:%d: Failure\nValue of: %s\n  Actual: false\n", __FILE__, __LINE__, #c); \
            ++shim_failures;                                                     \
            return;                                                              \
        }                                                                        \
    } while (0)
#define RUN_ALL_TESTS() ::testing::shim_run_all()
