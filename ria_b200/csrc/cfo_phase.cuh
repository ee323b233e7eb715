// Exact evaluation of the reference's fp32 CFO correction-phase accumulator, 32 samples at a time.
//
// Both demodulators rotate every sample by e^{j phase} and then do
//     phase += phase_inc;  if (phase > M_PI) phase -= 2.0f*M_PI;  else if (phase < -M_PI) phase += 2.0f*M_PI;
// in fp32 with a double-promoted wrap (src/ofdm/channel_equalizer.cpp:132-144;
// src/psk/multi_carrier_dpsk.hpp:913-921 uses two independent ifs, which is equivalent because
// only one can fire).  The rounding of that accumulator is visible in the soft bits at the 1e-4
// level over 10^4..10^5 samples, so it has to be reproduced exactly -- but not sequentially:
// while phase_k stays inside one binade (same exponent, same sign, no wrap), fl(phase + inc) =
// phase + S*ulp with the constant integer S = rint(inc / ulp), provided inc/ulp is not an exact
// tie.  A warp therefore proposes phase_k = phase_0 + k*S*ulp for its 32 lanes (exact in double),
// checks that the last value is still strictly inside the binade and the wrap range, and only
// otherwise falls back to stepping the 32 samples one by one.  Crossings happen once every few
// hundred to few thousand samples, so almost every block takes the closed form.
#pragma once

#include <cuda_runtime.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

// one reference step
__device__ __forceinline__ float cfo_phase_step(float ph, float inc) {
    ph = __fadd_rn(ph, inc);
    if (static_cast<double>(ph) > M_PI) ph = static_cast<float>(static_cast<double>(ph) - 2.0f * M_PI);
    else if (static_cast<double>(ph) < -M_PI) ph = static_cast<float>(static_cast<double>(ph) + 2.0f * M_PI);
    return ph;
}

// Warp-collective.  `base` = phase before sample 0 of the block (identical in all lanes).
// Returns the phase before sample `lane` of the block; *next = phase before sample 32.
__device__ __forceinline__ float cfo_phase_block32(float base, float inc, int lane, float* next) {
    bool fast = false;
    double step = 0.0;
    const unsigned bits = __float_as_uint(base);
    const int e = static_cast<int>((bits >> 23) & 0xFF);
    if (e > 0 && e < 255 && inc != 0.0f && fabs(static_cast<double>(base)) <= M_PI) {
        const double ulp = __longlong_as_double(static_cast<long long>(e - 127 - 23 + 1023) << 52);   // 2^(e-150)
        const double q = static_cast<double>(inc) / ulp;        // exact (power-of-two divisor)
        const double S = rint(q);
        if (fabs(q - trunc(q)) != 0.5 && fabs(S) < 4194304.0) {
            step = S * ulp;
            const double last = static_cast<double>(base) + 32.0 * step;      // exact
            const double lo = __longlong_as_double(static_cast<long long>(e - 127 + 1023) << 52);   // 2^(e-127)
            const double mag = fabs(last);
            // strictly inside the binade of `base`, same sign, and never beyond the wrap range
            fast = (mag > lo) && (mag < 2.0 * lo) && ((last < 0.0) == (base < 0.0f)) && (mag <= M_PI);
        }
    }
    if (fast) {
        *next = static_cast<float>(static_cast<double>(base) + 32.0 * step);
        return static_cast<float>(static_cast<double>(base) + static_cast<double>(lane) * step);
    }
    float ph = base, mine = base;
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
        if (k == lane) mine = ph;
        ph = cfo_phase_step(ph, inc);
    }
    *next = ph;
    return mine;
}

// Closed-form parameters of one 32-sample block: returns true and the exact per-sample step
// (S * ulp, representable in fp32) when all 32 phases of the block follow base + k * step.
__device__ __forceinline__ bool cfo_block_step(float base, float inc, double* step_out) {
    const unsigned bits = __float_as_uint(base);
    const int e = static_cast<int>((bits >> 23) & 0xFF);
    if (e > 0 && e < 255 && inc != 0.0f && fabs(static_cast<double>(base)) <= M_PI) {
        const double ulp = __longlong_as_double(static_cast<long long>(e - 127 - 23 + 1023) << 52);   // 2^(e-150)
        const double q = static_cast<double>(inc) / ulp;
        const double S = rint(q);
        if (fabs(q - trunc(q)) != 0.5 && fabs(S) < 4194304.0) {
            const double step = S * ulp;
            const double last = static_cast<double>(base) + 32.0 * step;      // exact
            const double lo = __longlong_as_double(static_cast<long long>(e - 127 + 1023) << 52);   // 2^(e-127)
            const double mag = fabs(last);
            if ((mag > lo) && (mag < 2.0 * lo) && ((last < 0.0) == (base < 0.0f)) && (mag <= M_PI)) {
                *step_out = step;
                return true;
            }
        }
    }
    return false;
}

// How many whole 32-sample blocks, starting at `base`, follow the closed form base + k * step:
// the phase after the last of them must still be strictly inside base's binade, on the same side
// of zero and inside the wrap range (the sequence is monotone, so the end point decides).
// Requires cfo_block_step(base, inc, &step) == true, hence the result is >= 1.
__device__ __forceinline__ int cfo_closed_form_blocks(float base, double step, int max_blocks) {
    const int e = static_cast<int>((__float_as_uint(base) >> 23) & 0xFF);
    const double lo = __longlong_as_double(static_cast<long long>(e - 127 + 1023) << 52);     // 2^(e-127)
    const double b = fabs(static_cast<double>(base));
    const double d = 32.0 * fabs(step);
    const bool outward = (step < 0.0) == (base < 0.0f);          // |phase| grows
    const double hi = fmin(2.0 * lo, M_PI);
    // outward: b + m d < 2 lo and <= pi;  inward: b - m d > lo
    double room = outward ? (hi - b) : (b - lo);
    int m = static_cast<int>(fmin(floor(room / d), static_cast<double>(max_blocks)));
    if (m < 1) m = 1;
    for (;;) {                                                   // settle the strict inequalities exactly
        const double last = outward ? b + m * d : b - m * d;
        const bool ok = (last > lo) && (last < 2.0 * lo) && (last <= M_PI);
        if (ok || m == 1) break;
        --m;
    }
    return m;
}

// Warp-collective scan of one frame: out[b] = {phase before sample 32 b, exact per-sample step of
// that block, or NaN when the block has to be stepped (binade crossing, wrap, tie)}.  Between two
// such events the closed form holds for a whole stretch of blocks, which the lanes fill in parallel.
// A consumer evaluates phase_k = fl(base + k * step), k < 32 (cfo_block_phase below).
__device__ __forceinline__ void cfo_scan_frame(float base, float inc, int n_blocks, float2* __restrict__ out, int lane) {
    int b = 0;
    while (b < n_blocks) {
        double step = 0.0;
        if (!cfo_block_step(base, inc, &step)) {
            if (lane == 0) out[b] = make_float2(base, __int_as_float(0x7fc00000));
#pragma unroll 1
            for (int k = 0; k < 32; ++k) base = cfo_phase_step(base, inc);
            ++b;
            continue;
        }
        const int m = cfo_closed_form_blocks(base, step, n_blocks - b);
        const float stepf = static_cast<float>(step);
        for (int j = lane; j < m; j += 32)
            out[b + j] = make_float2(static_cast<float>(static_cast<double>(base) + (32.0 * j) * step), stepf);
        base = static_cast<float>(static_cast<double>(base) + (32.0 * m) * step);
        b += m;
    }
}
__device__ __forceinline__ float cfo_block_phase(float2 blk, float inc, int k) {
    if (blk.y == blk.y) return static_cast<float>(static_cast<double>(blk.x) + static_cast<double>(k) * static_cast<double>(blk.y));
    float ph = blk.x;
#pragma unroll 1
    for (int i = 0; i < k; ++i) ph = cfo_phase_step(ph, inc);
    return ph;
}

// Per-thread variant: phase before sample k (0 <= k < 32) of a block whose first sample sees
// `base`.  Same closed form, checked over the k steps actually taken.
__device__ __forceinline__ float cfo_phase_at(float base, float inc, int k) {
    const unsigned bits = __float_as_uint(base);
    const int e = static_cast<int>((bits >> 23) & 0xFF);
    if (e > 0 && e < 255 && inc != 0.0f && fabs(static_cast<double>(base)) <= M_PI) {
        const double ulp = __longlong_as_double(static_cast<long long>(e - 127 - 23 + 1023) << 52);   // 2^(e-150)
        const double q = static_cast<double>(inc) / ulp;
        const double S = rint(q);
        if (fabs(q - trunc(q)) != 0.5 && fabs(S) < 4194304.0) {
            const double last = static_cast<double>(base) + static_cast<double>(k) * (S * ulp);       // exact
            const double lo = __longlong_as_double(static_cast<long long>(e - 127 + 1023) << 52);     // 2^(e-127)
            const double mag = fabs(last);
            if ((mag > lo) && (mag < 2.0 * lo) && ((last < 0.0) == (base < 0.0f)) && (mag <= M_PI))
                return static_cast<float>(last);
        }
    }
    float ph = base;
#pragma unroll 1
    for (int i = 0; i < k; ++i) ph = cfo_phase_step(ph, inc);
    return ph;
}

}  // namespace ria
