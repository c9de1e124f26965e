// TEST INFRASTRUCTURE: gtest_main stand-in for the reference's test files that define no main().
#include <gtest/gtest.h>
int main() {
    testing::InitGoogleTest();
    return RUN_ALL_TESTS();
}
