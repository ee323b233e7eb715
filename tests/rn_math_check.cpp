// Host check of ria_b200/csrc/rn_math.h: the restated glibc float functions must return the same
// bits as the container's libm (which is what the reference calls).  Prints mismatch counts;
// tests/test_rn_math_cpu.py asserts on them.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../ria_b200/csrc/rn_math.h"

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 4000000;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::normal_distribution<double> g(0.0, 1.0);
    long bad_sin = 0, bad_cos = 0, bad_sincos = 0, bad_atan2 = 0, bad_large = 0;
    for (long i = 0; i < n; ++i) {
        const int kind = i % 5;
        const float x = static_cast<float>(kind == 0 ? u(rng) * M_PI : kind == 1 ? u(rng) * 119.9 : kind == 2 ? u(rng) * 0.11
                                         : kind == 3 ? u(rng) * 1e-3 : u(rng) * 0.79);
        float s, c;
        glibc_sincosf(x, &s, &c);
        float rs, rc;
        sincosf(x, &rs, &rc);
        bad_sincos += (rn_fbits(s) != rn_fbits(rs)) + (rn_fbits(c) != rn_fbits(rc));
        glibc_sincosf_uniform(x, &s, &c);
        bad_sincos += (rn_fbits(s) != rn_fbits(rs)) + (rn_fbits(c) != rn_fbits(rc));
        bad_sin += rn_fbits(glibc_sinf(x)) != rn_fbits(sinf(x));
        bad_cos += rn_fbits(glibc_cosf(x)) != rn_fbits(cosf(x));
    }
    // |x| >= 120: reduce_large, bit for bit (chirp / ZC preamble phases reach several 10^3 rad)
    for (long i = 0; i < n / 4; ++i) {
        const int kind = i % 4;
        float x = static_cast<float>(kind == 0 ? u(rng) * 5000.0 : kind == 1 ? u(rng) * 3.0e5 : kind == 2 ? u(rng) * 1.0e9
                                               : std::ldexp(u(rng), static_cast<int>(i % 120) + 7));
        if (std::fabs(x) < 120.0f) x = std::copysign(120.0f + std::fabs(x), x);
        float s, c, rs, rc;
        glibc_sincosf(x, &s, &c);
        sincosf(x, &rs, &rc);
        bad_large += (rn_fbits(s) != rn_fbits(rs)) + (rn_fbits(c) != rn_fbits(rc));
        bad_large += rn_fbits(glibc_sinf(x)) != rn_fbits(sinf(x));
        bad_large += rn_fbits(glibc_cosf(x)) != rn_fbits(cosf(x));
    }
    for (long i = 0; i < n; ++i) {
        const int kind = i % 5;
        float y = static_cast<float>(g(rng)), x = static_cast<float>(g(rng));
        if (kind == 1) y *= 1e-3f;             // small angles (decision-directed phase errors)
        if (kind == 2) x *= 1e-3f;             // near +-pi/2
        if (kind == 3) { x = static_cast<float>(u(rng)); y = x * static_cast<float>(1.0 + 1e-3 * u(rng)); }   // octant edges
        if (kind == 4) { y *= static_cast<float>(std::exp(12 * u(rng))); x *= static_cast<float>(std::exp(12 * u(rng))); }
        bad_atan2 += rn_fbits(glibc_atan2f(y, x)) != rn_fbits(atan2f(y, x));
    }
    int special_bad = 0;
    const float sv[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, 1e-30f, -1e-30f, 3e38f, -3e38f, 1e-45f, 0.4375f, 0.6875f, 1.1875f, 2.4375f};
    for (float y : sv) for (float x : sv) {
        const float r = glibc_atan2f(y, x), gr = atan2f(y, x);
        if (rn_fbits(r) != rn_fbits(gr)) { ++special_bad; fprintf(stderr, "atan2(%g,%g) = %a vs %a\n", y, x, r, gr); }
    }
    for (float x : sv) {
        if (std::isinf(x)) continue;
        float s, c;
        glibc_sincosf(x, &s, &c);
        if (std::fabs(x) < 120.0f && (rn_fbits(s) != rn_fbits(sinf(x)) || rn_fbits(c) != rn_fbits(cosf(x)))) {
            ++special_bad; fprintf(stderr, "sincos(%g)\n", x);
        }
    }
    // logf on (0, 1]: every 61st normal float plus the neighbourhood of 1 (the polar method's r2)
    long bad_log = 0;
    for (uint32_t u = 0x00800000u; u <= 0x3f800000u; u += (u > 0x3f7f0000u ? 1u : 61u)) {
        const float x = rn_ffrom(u);
        bad_log += rn_fbits(glibc_logf(x)) != rn_fbits(logf(x));
    }
    printf("n %ld bad_sin %ld bad_cos %ld bad_sincos %ld bad_atan2 %ld bad_large %ld special_bad %d bad_log %ld\n",
           n, bad_sin, bad_cos, bad_sincos, bad_atan2, bad_large, special_bad, bad_log);
    return 0;
}
