// Tiny stand-in for the googletest macros used by the reference's Dem unit tests (src/tests/unit_tests/dem/*.cpp), so
// the scenarios below read like the originals.  Each test binary returns 0 on success.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <exception>

#define ASSERT_TRUE(c)                                                                  \
    do {                                                                                \
        if (!(c)) {                                                                     \
            std::printf("\nFAILED %s:%d: ASSERT_TRUE(%s)\n", __FILE__, __LINE__, #c);   \
            std::exit(1);                                                               \
        }                                                                               \
    } while (0)
#define ASSERT_NEAR(a, b, tol)                                                                                    \
    do {                                                                                                          \
        const double a_ = (double)(a), b_ = (double)(b), t_ = (double)(tol);                                      \
        if (!(std::fabs(a_ - b_) <= t_)) {                                                                        \
            std::printf("\nFAILED %s:%d: ASSERT_NEAR(%s = %.9g, %s = %.9g, %g)\n", __FILE__, __LINE__, #a, a_, #b, b_, t_); \
            std::exit(1);                                                                                         \
        }                                                                                                         \
    } while (0)
#define RUN_TEST(fn)                                              \
    int main(int argc, char** argv) {                             \
        try {                                                     \
            fn(argc, argv);                                       \
        } catch (const std::exception& e) {                       \
            std::printf("\nEXCEPTION: %s\n", e.what());           \
            return 2;                                             \
        }                                                         \
        std::printf("\nPASSED %s\n", #fn);                        \
        return 0;                                                 \
    }
