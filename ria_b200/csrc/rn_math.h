// glibc's float transcendentals, restated so that the device returns the same bits as the host.
//
// The reference calls std::atan2 / std::arg / std::sin / std::cos / std::polar / std::exp on
// floats, i.e. glibc's atan2f, sinf, cosf and sincosf.  Neither is correctly rounded (atan2f is
// off by one ulp in ~14 % of calls, sinf/cosf in ~1 %), and CUDA's own float versions differ from
// them by 1-2 ulp -- which the QAM demapper's 2/noise_var scale amplifies past the 1e-4 soft-bit
// contract.  So the algorithms glibc 2.39 (x86-64) runs are restated here operation by operation:
//   glibc_atan2f / glibc_atanf   sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c (fdlibm, pure fp32,
//                                no FMA: x86-64 glibc has no FMA variant of these);
//   glibc_sinf / glibc_cosf / glibc_sincosf
//                                sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h (ARM
//                                optimized-routines: double polynomial, the x86-64 ifunc picks
//                                the FMA build on every CPU with FMA3, so a + b*c is one fma).
// glibc is not part of /root/reference; the restatement is pinned bit-for-bit against the
// container's libm on 2 x 10^7 arguments per function (tests/test_rn_math_cpu.py), and the same
// source compiles for the device.  |x| >= 120 takes sincosf.h's reduce_large (the preamble generators'
// phases of several thousand radians go through it).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RN_HD __host__ __device__ __forceinline__
#else
#define RN_HD static inline
#endif

RN_HD uint32_t rn_fbits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RN_HD float rn_ffrom(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
// fp32 a*b + c with two roundings, whatever the compiler's contraction setting
RN_HD float rn_mad(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(__fmul_rn(a, b), c);
#else
    volatile float p = a * b;
    return p + c;
#endif
}
RN_HD float rn_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
RN_HD float rn_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
RN_HD float rn_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

// ------------------------------------------------------------------------------------------
// atanf / atan2f (fdlibm)
// ------------------------------------------------------------------------------------------
RN_HD float glibc_atanf(float x) {
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const int32_t hx = static_cast<int32_t>(rn_fbits(x));
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.0f, lo = 0.0f;
    if (ix >= 0x4c000000) {                      // |x| >= 2^25
        if (ix > 0x7f800000) return rn_add(x, x);
        return (hx > 0) ? rn_add(hi3, lo3) : rn_add(-hi3, -lo3);
    }
    if (ix < 0x3ee00000) {                       // |x| < 0.4375
        if (ix < 0x31000000) return x;           // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {                   // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; hi = hi0; lo = lo0; x = rn_div(rn_add(rn_mul(2.0f, x), -1.0f), rn_add(2.0f, x)); }
            else                 { id = 1; hi = hi1; lo = lo1; x = rn_div(rn_add(x, -1.0f), rn_add(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; hi = hi2; lo = lo2; x = rn_div(rn_add(x, -1.5f), rn_mad(1.5f, x, 1.0f)); }
            else                 { id = 3; hi = hi3; lo = lo3; x = rn_div(-1.0f, x); }
        }
    }
    const float z = rn_mul(x, x);
    const float w = rn_mul(z, z);
    // break sum from i=0 to 10 aT[i] z^(i+1) into odd and even poly
    const float s1 = rn_mul(z, rn_mad(w, rn_mad(w, rn_mad(w, rn_mad(w, rn_mad(w, aT10, aT8), aT6), aT4), aT2), aT0));
    const float s2 = rn_mul(w, rn_mad(w, rn_mad(w, rn_mad(w, rn_mad(w, aT9, aT7), aT5), aT3), aT1));
    const float xs = rn_mul(x, rn_add(s1, s2));
    if (id < 0) return rn_add(x, -xs);
    const float zz = rn_add(hi, -rn_add(rn_add(xs, -lo), -x));
    return (hx < 0) ? -zz : zz;
}

RN_HD float glibc_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = static_cast<int32_t>(rn_fbits(x)), hy = static_cast<int32_t>(rn_fbits(y));
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return rn_add(x, y);           // NaN
    if (hx == 0x3f800000) return glibc_atanf(y);                          // x = 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                     // 2*sign(x) + sign(y)
    if (iy == 0) {
        switch (m) { case 0: case 1: return y; case 2: return rn_add(pi, tiny); default: return rn_add(-pi, -tiny); }
    }
    if (ix == 0) return (hy < 0) ? rn_add(-pi_o_2, -tiny) : rn_add(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return rn_add(pi_o_4, tiny);
                case 1: return rn_add(-pi_o_4, -tiny);
                case 2: return rn_add(rn_mul(3.0f, pi_o_4), tiny);
                default: return rn_add(rn_mul(-3.0f, pi_o_4), -tiny);
            }
        }
        switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return rn_add(pi, tiny); default: return rn_add(-pi, -tiny); }
    }
    if (iy == 0x7f800000) return (hy < 0) ? rn_add(-pi_o_2, -tiny) : rn_add(pi_o_2, tiny);
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = rn_add(pi_o_2, rn_mul(0.5f, pi_lo));                   // |y/x| > 2^60
    else if (hx < 0 && k < -60) z = 0.0f;                                  // |y|/x < -2^60
    else z = glibc_atanf(fabsf(rn_div(y, x)));
    switch (m) {
        case 0: return z;
        case 1: return rn_ffrom(rn_fbits(z) ^ 0x80000000u);
        case 2: return rn_add(pi, -rn_add(z, -pi_lo));
        default: return rn_add(rn_add(z, -pi_lo), -pi);
    }
}

// ------------------------------------------------------------------------------------------
// sinf / cosf / sincosf (ARM optimized-routines as built by glibc with FMA)
// ------------------------------------------------------------------------------------------
// sincosf.h: __sincosf_table[0]; table[1] is the same with c0..c4 negated (flip = -1)
#define RN_SC_C0 0x1p0
#define RN_SC_C1 -0x1.ffffffd0c621cp-2
#define RN_SC_C2 0x1.55553e1068f19p-5
#define RN_SC_C3 -0x1.6c087e89a359dp-10
#define RN_SC_C4 0x1.99343027bf8c3p-16
#define RN_SC_S1 -0x1.555545995a603p-3
#define RN_SC_S2 0x1.1107605230bc4p-7
#define RN_SC_S3 -0x1.994eb3774cf24p-13
#define RN_SC_HPI_INV 0x1.45F306DC9C883p+23
#define RN_SC_HPI 0x1.921FB54442D18p0

RN_HD uint32_t rn_abstop12(float x) { return (rn_fbits(x) >> 20) & 0x7ff; }

// sine polynomial of sincosf_poly / sinf_poly
RN_HD double rn_sc_sin(double x, double x2) {
    const double x3 = x * x2;
    const double s1 = fma(x2, RN_SC_S3, RN_SC_S2);
    const double x5 = x3 * x2;
    const double s = fma(x3, RN_SC_S1, x);
    return fma(x5, s1, s);
}
// cosine polynomial; flip = +1 for table[0], -1 for table[1] (negated coefficients)
RN_HD double rn_sc_cos(double x2, double flip) {
    const double x4 = x2 * x2;
    const double c2 = fma(x2, flip * RN_SC_C4, flip * RN_SC_C3);
    const double c1 = fma(x2, flip * RN_SC_C1, flip * RN_SC_C0);
    const double x6 = x4 * x2;
    const double c = fma(x4, flip * RN_SC_C2, c1);
    return fma(x6, c2, c);
}
// reduce_fast: n = round(x / (pi/2)), returns x - n pi/2
RN_HD double rn_sc_reduce(double x, int* np) {
    const double r = x * RN_SC_HPI_INV;
    const int n = (static_cast<int32_t>(r) + 0x800000) >> 24;
    *np = n;
    return fma(-static_cast<double>(n), RN_SC_HPI, x);
}

// reduce_large (sincosf.h): |y| >= 120.  The 24 significand bits (shifted by exponent & 7) times a
// 96-bit window of 2/pi selected by the exponent; the top two bits of the product are the quadrant,
// the rest is the reduced angle as a signed 62-bit fixed-point fraction of pi/2.
RN_HD double rn_sc_reduce_large(uint32_t xi, int* np) {
    // 2/pi = 0.A2F9836E 4E441529 FC2757D1 F534DDC0 DB629599 3C439041 ... (hex), 32-bit windows 8 bits apart
    static const uint32_t inv_pio4[24] = {
        0x000000a2u, 0x0000a2f9u, 0x00a2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
        0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
        0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};
    const uint32_t* arr = &inv_pio4[(xi >> 26) & 15];
    const int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    uint64_t res0 = static_cast<uint32_t>(xi * arr[0]);
    const uint64_t res1 = static_cast<uint64_t>(xi) * arr[4];
    const uint64_t res2 = static_cast<uint64_t>(xi) * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    *np = static_cast<int>(n);
    return static_cast<double>(static_cast<int64_t>(res0)) * 0x1.921FB54442D18p-62;
}

RN_HD void glibc_sincosf(float y, float* sinp, float* cosp) {
    double x = static_cast<double>(y);
    if (rn_abstop12(y) < rn_abstop12(0x1.921FB6p-1f)) {                     // |y| < pi/4
        const double x2 = x * x;
        if (rn_abstop12(y) < rn_abstop12(0x1p-12f)) { *sinp = y; *cosp = 1.0f; return; }
        *sinp = static_cast<float>(rn_sc_sin(x, x2));
        *cosp = static_cast<float>(rn_sc_cos(x2, 1.0));
    } else if (rn_abstop12(y) < rn_abstop12(120.0f)) {
        int n;
        x = rn_sc_reduce(x, &n);
        // sign[n & 3] = {1, -1, -1, 1}; table[1] (n & 2) negates the cosine polynomial
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        const double flip = (n & 2) ? -1.0 : 1.0;
        const double x2 = x * x;
        const float ps = static_cast<float>(rn_sc_sin(x * s, x2));
        const float pc = static_cast<float>(rn_sc_cos(x2, flip));
        if (n & 1) { *sinp = pc; *cosp = ps; } else { *sinp = ps; *cosp = pc; }
    } else if (rn_abstop12(y) < rn_abstop12(INFINITY)) {
        const uint32_t xi = rn_fbits(y);
        const int sign = static_cast<int>(xi >> 31);
        int n;
        x = rn_sc_reduce_large(xi, &n);
        const int m = n + sign;
        const double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
        const double flip = (m & 2) ? -1.0 : 1.0;
        const double x2 = x * x;
        const float ps = static_cast<float>(rn_sc_sin(x * s, x2));
        const float pc = static_cast<float>(rn_sc_cos(x2, flip));
        if (n & 1) { *sinp = pc; *cosp = ps; } else { *sinp = ps; *cosp = pc; }
    } else {
        *sinp = *cosp = y - y;                                             // inf / NaN -> NaN
    }
}

// Same results as glibc_sincosf for |y| < 120 without the |y| < pi/4 fork: reduce_fast returns
// n = 0 and x unchanged there, and below 2^-12 the polynomials round to (y, 1) by themselves, so
// one straight-line path serves all lanes of a warp.  (|y| >= 120 goes to glibc_sincosf.)
RN_HD void glibc_sincosf_uniform(float y, float* sinp, float* cosp) {
    if (!(rn_abstop12(y) < rn_abstop12(120.0f))) { glibc_sincosf(y, sinp, cosp); return; }
    int n;
    const double x = rn_sc_reduce(static_cast<double>(y), &n);
    const double x2 = x * x;
    // glibc evaluates sin_poly(x * s) with s = sign[n & 3] = {1, -1, -1, 1} and the cosine polynomial with all
    // coefficients negated when n & 2.  Both polynomials are built from products and fused multiply-adds only, and
    // round-to-nearest is symmetric, so negating the argument of the odd polynomial / every coefficient of the even
    // one negates the result bit for bit: evaluate once with positive signs and flip the float's sign bit instead.
    const float ps0 = static_cast<float>(rn_sc_sin(x, x2));
    const float pc0 = static_cast<float>(rn_sc_cos(x2, 1.0));
    const float ps = rn_ffrom(rn_fbits(ps0) ^ (static_cast<uint32_t>((n + 1) & 2) << 30));
    const float pc = rn_ffrom(rn_fbits(pc0) ^ (static_cast<uint32_t>(n & 2) << 30));
    *sinp = (n & 1) ? pc : ps;
    *cosp = (n & 1) ? ps : pc;
}

RN_HD float glibc_sinf(float y) {
    double x = static_cast<double>(y);
    if (rn_abstop12(y) < rn_abstop12(0x1.921FB6p-1f)) {
        if (rn_abstop12(y) < rn_abstop12(0x1p-12f)) return y;
        return static_cast<float>(rn_sc_sin(x, x * x));
    }
    if (rn_abstop12(y) < rn_abstop12(120.0f)) {
        int n;
        x = rn_sc_reduce(x, &n);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        const double flip = (n & 2) ? -1.0 : 1.0;
        return static_cast<float>((n & 1) ? rn_sc_cos(x * x, flip) : rn_sc_sin(x * s, x * x));
    }
    if (!(rn_abstop12(y) < rn_abstop12(INFINITY))) return y - y;
    const uint32_t xi = rn_fbits(y);
    const int sign = static_cast<int>(xi >> 31);
    int n;
    x = rn_sc_reduce_large(xi, &n);
    const int m = n + sign;
    const double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
    const double flip = (m & 2) ? -1.0 : 1.0;
    return static_cast<float>((n & 1) ? rn_sc_cos(x * x, flip) : rn_sc_sin(x * s, x * x));
}

RN_HD float glibc_cosf(float y) {
    double x = static_cast<double>(y);
    if (rn_abstop12(y) < rn_abstop12(0x1.921FB6p-1f)) {
        if (rn_abstop12(y) < rn_abstop12(0x1p-12f)) return 1.0f;
        return static_cast<float>(rn_sc_cos(x * x, 1.0));
    }
    if (rn_abstop12(y) < rn_abstop12(120.0f)) {
        int n;
        x = rn_sc_reduce(x, &n);
        // cosf: sign[(n + 1) & 3], table[1] when (n + 1) & 2, polynomial selected by n ^ 1
        const int m = n + 1;
        const double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
        const double flip = (m & 2) ? -1.0 : 1.0;
        return static_cast<float>(((n ^ 1) & 1) ? rn_sc_cos(x * x, flip) : rn_sc_sin(x * s, x * x));
    }
    if (!(rn_abstop12(y) < rn_abstop12(INFINITY))) return y - y;
    const uint32_t xi = rn_fbits(y);
    const int sign = static_cast<int>(xi >> 31);
    int n;
    x = rn_sc_reduce_large(xi, &n);
    const int m = n + sign;
    const double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
    const double flip = (m & 2) ? -1.0 : 1.0;
    return static_cast<float>(((n ^ 1) & 1) ? rn_sc_cos(x * x, flip) : rn_sc_sin(x * s, x * x));
}

// ------------------------------------------------------------------------------------------
// logf (ARM optimized-routines, glibc sysdeps/ieee754/flt-32/e_logf.c + e_logf_data.c)
// ------------------------------------------------------------------------------------------
// libstdc++'s std::normal_distribution<float> (Marsaglia polar) calls std::log(float) == logf;
// the LDPC retry ladder of v2::decodeFixedFrame perturbs soft bits with it
// (src/protocol/frame_v2.cpp:1389-1546), so the perturbation is only reproducible with glibc's
// logf.  The table and polynomial below are __logf_data of the container's libm (2.39), read out
// of the binary; the restatement matches libm on every normal float in (0, 1] -- the only
// arguments the polar method produces -- with or without FMA contraction (tests/rn_math_check.cpp).
// Only positive normal arguments are handled (r2 in (0, 1]); others return NaN.
RN_HD float glibc_logf(float x) {
    static const double tab[32] = {
        0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
        0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
        0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
        0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
        0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0,
        0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
        0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
        0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};
    const uint32_t ix = rn_fbits(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return rn_ffrom(0x7fc00000u);
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = static_cast<int>((tmp >> 19) & 15u);
    const int k = static_cast<int32_t>(tmp) >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    const double invc = tab[2 * i], logc = tab[2 * i + 1];
    const double z = static_cast<double>(rn_ffrom(iz));
    const double r = fma(z, invc, -1.0);
    const double y0 = fma(static_cast<double>(k), 0x1.62e42fefa39efp-1, logc);
    const double r2 = r * r;
    double y = fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = fma(-0x1.00ea348b88334p-2, r2, y);
    y = fma(y, r2, y0 + r);
    return static_cast<float>(y);
}
