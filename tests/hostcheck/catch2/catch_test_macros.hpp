// Test infrastructure: a minimal stand-in for <catch2/catch_test_macros.hpp> (Catch2 is not installed in this image and the
// reference's CMake fetches it from the network) so that the reference's own unit test, /root/reference/cpp/tests/test_c4.cpp,
// compiles UNCHANGED against this tree's headers and runs on the GPU box.  TEST_CASE registers a function, REQUIRE ends the
// program with a message, main() runs every case.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mini_catch {
struct Case { const char* name; void (*fn)(); };
inline std::vector<Case>& cases() { static std::vector<Case> all; return all; }
struct Registrar { Registrar(const char* name, void (*fn)()) { cases().push_back({ name, fn }); } };
}  // namespace mini_catch

#define MINI_CATCH_CAT2(a, b) a##b
#define MINI_CATCH_CAT(a, b) MINI_CATCH_CAT2(a, b)
#define TEST_CASE(name)                                                                                              \
    static void MINI_CATCH_CAT(mini_catch_case_, __LINE__)();                                                        \
    static mini_catch::Registrar MINI_CATCH_CAT(mini_catch_reg_, __LINE__)(name, MINI_CATCH_CAT(mini_catch_case_, __LINE__)); \
    static void MINI_CATCH_CAT(mini_catch_case_, __LINE__)()
#define REQUIRE(...)                                                                                  \
    do {                                                                                              \
        if (!(__VA_ARGS__)) {                                                                         \
            std::fprintf(stderr, "REQUIRE failed: %s (line %d)\n", #__VA_ARGS__, __LINE__);           \
            std::exit(1);                                                                             \
        }                                                                                             \
    } while (0)

int main() {
    for (const mini_catch::Case& c : mini_catch::cases()) {
        c.fn();
        std::printf("passed: %s\n", c.name);
    }
    std::printf("all %zu test case(s) passed\n", mini_catch::cases().size());
    return 0;
}
