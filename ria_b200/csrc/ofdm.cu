// Batched OFDM presynced receive chain for sm_100a: mix-down + CFO correction + CP strip +
// 1024-point FFT + LTS channel estimate + pilot tracking + MMSE equalise + soft demap.
//
// Replaces, per frame, OFDMChirpWaveform::process (src/waveform/ofdm_chirp_waveform.cpp:391-468)
//   = OFDMDemodulator::setFrequencyOffsetWithPhase + processPresynced
//     (src/ofdm/demodulator.cpp:1212-1221, 1250-1414), which runs
//       Impl::toBaseband / extractSymbol          src/ofdm/channel_equalizer.cpp:99-187
//       FFT::forward (radix-2 DIT)                src/dsp/fft.cpp:96-146
//       Impl::estimateChannelFromLTS              channel_equalizer.cpp:193-643
//       Impl::updateChannelEstimate               channel_equalizer.cpp:645-1043
//       Impl::equalize / hardDecision             channel_equalizer.cpp:1168-1451
//       Impl::demodulateSymbol + soft_demap::*    demodulator.cpp:208-508, soft_demap.hpp:22-263
//
// Mapping (B200).  The only sequential dependence between the symbols of a frame is the
// per-carrier equaliser state; the FFTs are independent (unless the LTS finds a residual CFO,
// see below).  The batch therefore runs as a short pipeline of kernels over chunks of frames:
//   ofdm_phase_scan_kernel  (only when a CFO is given) one warp per frame: value of the fp32
//                           CFO phase accumulator at every 32nd sample, exact (cfo_phase.cuh);
//   ofdm_fft_kernel         128-thread CTAs, work item = (symbol index, 32 frames): the mixer
//                           phasors of that symbol are staged once in shared memory (the plain
//                           variant keeps its share in registers), then every frame's samples are
//                           read from HBM exactly once (coalesced, CP never loaded, next frame
//                           prefetched into registers), mixed and transformed with the reference's
//                           radix-2 DIT butterflies in the reference's order -- three stages per
//                           pass in registers, exchanged through an additively padded shared-memory
//                           tile (conflict-free for every pass, per-thread base + constant offsets).
//                           The last four stages are evaluated only for the <= 63 used bins (output
//                           pruning: 7 half-butterflies per thread instead of 12 + 1 full ones).
//                           Every used bin is bit-identical to the reference FFT.
//   ofdm_carrier2_kernel    two frames per warp (half-warp groups), no block barriers: LTS
//                           estimate, then the data symbols in order (pilot tracking, MMSE,
//                           LLRs).  Reductions whose fp32 summation order is observable (they
//                           feed thresholds) are done in the reference's order by the group's
//                           first lane.  ofdm_carrier_kernel (one frame per warp) is the A/B path.
//   ofdm_presynced_kernel   the monolithic per-frame kernel (FFT + carriers in one CTA); runs
//                           only for the frames whose LTS reports a residual CFO in (0.3, 5) Hz:
//                           the reference then re-mixes the whole frame with the corrected CFO
//                           (channel_equalizer.cpp:304-382), which the carrier kernel cannot do.
// Float-order fidelity: compiled with --fmad=false; complex multiply/divide/abs follow what
// libstdc++/libgcc do on x86-64 (division and abs go through double, see cdiv/cabs below).

#include "ofdm_tables.h"
#include "cfo_phase.cuh"
#include "rn_math.h"

#include <cfloat>
#include <cstddef>
#include <cstdlib>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

namespace {

constexpr int kThreads = 128;
constexpr int kFft = 1024;
constexpr int kMaxSymLen = 1160;    // fft 1024 + CP <= 128 + guard <= 8
constexpr int kTwCount = 1024;      // stage-major twiddle table (1022 used)
constexpr int kFftGroup = 32;       // frames per (symbol, group) work item of the FFT kernel
constexpr int kCarWarps = 4;        // frames per CTA of the carrier kernel

// ------------------------------- complex helpers -------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    // (a.x + i a.y)(b.x + i b.y) as GCC expands it without -ffast-math: no FMA, this order
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s)); }
__device__ __forceinline__ float2 cdivf(float2 a, float s) { return make_float2(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s)); }
__device__ __forceinline__ float cnorm(float2 a) { return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }
// std::abs(std::complex<float>) = hypotf; glibc evaluates it in double and rounds once
// (verified bit-exact against the reference toolchain on 2M random inputs).
__device__ __noinline__ float cabs(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}
// complex<float> / complex<float> = libgcc __divsc3: plain formula evaluated in double
// (verified bit-exact against the reference toolchain on 2M random inputs).
__device__ __noinline__ float2 cdiv(float2 a, float2 b) {
    const double aa = a.x, bb = a.y, cc = b.x, dd = b.y;
    const double den = cc * cc + dd * dd;
    return make_float2(static_cast<float>((aa * cc + bb * dd) / den),
                       static_cast<float>((bb * cc - aa * dd) / den));
}
// Transcendentals: the reference calls glibc's float atan2f / sinf / cosf, which are not correctly
// rounded; rn_math.h restates the algorithms glibc runs so the device returns the same bits.
// (real calls: one copy of each in the instruction cache instead of one per call site)
__device__ __noinline__ float atan2_rn(float y, float x) { return glibc_atan2f(y, x); }
__device__ __forceinline__ float sin_rn(float x) { return glibc_sinf(x); }
__device__ __forceinline__ float cos_rn(float x) { return glibc_cosf(x); }
__device__ __forceinline__ float carg(float2 a) { return atan2_rn(a.y, a.x); }
__device__ __noinline__ float2 cexpj(float th) {
    float s, c;
    glibc_sincosf(th, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }   // std::max(a,b)
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }   // std::min(a,b)

// clipLLR, src/ofdm/soft_demap.hpp:22-29
__device__ __forceinline__ float clip_llr(float llr) {
    float c = std_max(-20.0f, std_min(20.0f, llr));
    if (fabsf(c) < 0.01f) c = (c >= 0.0f) ? 0.01f : -0.01f;
    return c;
}

// ------------------------------- per-frame equaliser state ---------------------------------
struct FrameScalars {
    float cfo_hz, cfo_phase, phase_start;
    float noise_var, snr_lin, slope;
    float avg_h_power;
    float signal_power;
    float2 cpc;                 // carrier_phase_correction
    int cpc_init;
    int snr_count;              // snr_symbol_count
    int rerun;
    int apply_cpe;
    float2 cpe;
    int noise_count;
    int have_prev_pilot;
    int have_dd;
    float last_fading;          // last_fading_index (LTS / pilot magnitude CV): gates the D8PSK two-pass demap
    float2 d8_corr;             // D8PSK two-pass: common phase correction of the current symbol
};

constexpr int kMaxPilots = 32;      // pilot_spacing >= 2 (checked by ofdm_config_error)

struct alignas(16) CarState {
    float2 bin[kMaxCarriers];
    float2 H[kMaxCarriers];
    float2 hps[2][kMaxCarriers];        // LTS estimates; after the LTS: de-slope / re-slope phasors
    float2 tmpc[kMaxCarriers];
    float tmpf[kMaxCarriers];
    float tmpg[kMaxCarriers];           // also |H|^2 of the data carriers during equalisation
    int   flag[kMaxCarriers];
    float2 pil_ls[kMaxPilots];
    float2 prev_pilot[kMaxPilots];
    float2 prev_eq[kMaxCarriers];       // differential reference; coherent: de-sloped pilot estimates
    float ema[kMaxCarriers];
    float var[kMaxCarriers];
    float dd[kMaxCarriers];
    FrameScalars s;
};
static_assert(offsetof(CarState, tmpf) == offsetof(CarState, tmpc) + 4 * kMaxPilots * sizeof(float) &&
              offsetof(CarState, tmpg) == offsetof(CarState, tmpf) + 2 * kMaxPilots * sizeof(float),
              "the reduction rows are laid over tmpc | tmpf | tmpg");

// what every stage needs to know about the batch
struct KernelArgs {
    const float* samples; long long frame_stride; int frame_len;
    const float* cfo_hz; const float* phase; long long n_frames;
    float* llr; int llr_stride; int* n_llr;
    float* snr_db; float* cfo_out; float* fading;
    float* bins_tap; float* h_lts_tap;
    const float2* tw_g; const float2* nco_g; const OfdmCarrierTable* car_g;
    int cp, sym_len, modulation, differential, bits_per_carrier, sample_rate;
    unsigned int* counter;
    // staged pipeline
    float2* bins;                 // [frame][symbol][carrier] carrier bins written by the FFT kernel
    long long bins_frame0;        // frame index bins[0] belongs to
    const float2* block_phase;    // [frame - frame_begin][block] {CFO accumulator at sample 32*block, per-sample step or NaN}
    int n_blocks;
    int pruned;                   // all used bins have distinct residues mod 64
    int* rerun_list;              // chunk-local ids of the frames whose LTS asks for the residual-CFO re-run
    unsigned int* rerun_count;
    float* rerun_cfo;             // [chunk] corrected CFO of those frames (input CFO + LTS residual)
    int second_pass;              // this launch re-runs the listed frames with rerun_cfo
    long long frame_begin, frame_end;   // chunk of the batch this launch covers
};

// Group of G threads working on one frame: a CTA (G = 128), a warp (G = 32) or a half-warp (G = 16: two frames share
// a warp, each with its own member mask, so that the two halves may sit in different branches).
template <int G> __device__ __forceinline__ unsigned gmask() {
    return (G == 16) ? (0xFFFFu << (threadIdx.x & 16)) : 0xffffffffu;
}
template <int G> __device__ __forceinline__ void gsync() {
    if (G == 16) __syncwarp(gmask<16>()); else if (G == 32) __syncwarp(); else __syncthreads();
}
// value of lane `src` of the group's first warp (G >= 32) / of the half-warp (G = 16)
template <int G> __device__ __forceinline__ float gshfl(float v, int src) {
    return __shfl_sync(gmask<G>(), v, src, G < 32 ? G : 32);
}

// hardDecision, channel_equalizer.cpp:1168-1230
__device__ float2 hard_decision(float2 s, int mod) {
    switch (mod) {
        case RIA_BPSK: return make_float2(s.x > 0 ? 1.0f : -1.0f, 0.0f);
        case RIA_QAM16: {
            auto sl = [](float x) { return x < -0.4f ? -0.9487f : (x < 0.0f ? -0.3162f : (x < 0.4f ? 0.3162f : 0.9487f)); };
            return make_float2(sl(s.x), sl(s.y));
        }
        case RIA_QAM32: {
            const float d = 0.1961161351381840f;
            auto si = [d](float x) { return x < -2 * d ? -3 * d : (x < 0 ? -d : (x < 2 * d ? d : 3 * d)); };
            auto sq = [d](float x) {
                return x < -6 * d ? -7 * d : x < -4 * d ? -5 * d : x < -2 * d ? -3 * d : x < 0 ? -d
                     : x < 2 * d ? d : x < 4 * d ? 3 * d : x < 6 * d ? 5 * d : 7 * d; };
            return make_float2(si(s.x), sq(s.y));
        }
        case RIA_QAM64: {
            const float d = 0.1543f;
            // the reference's chain of seven ascending thresholds (first true wins) as a three-level search over the
            // same threshold and level expressions: identical result for every x, NaN included (all compares false)
            auto sl = [d](float x) {
                return x < 0 ? (x < -4 * d ? (x < -6 * d ? -7 * d : -5 * d) : (x < -2 * d ? -3 * d : -d))
                             : (x < 4 * d ? (x < 2 * d ? d : 3 * d) : (x < 6 * d ? 5 * d : 7 * d)); };
            return make_float2(sl(s.x), sl(s.y));
        }
        default:   // QPSK and everything else
            return make_float2(s.x > 0 ? 0.7071f : -0.7071f, s.y > 0 ? 0.7071f : -0.7071f);
    }
}

// getCEErrorMargin, soft_demap.hpp:298-326
__device__ __forceinline__ float ce_margin(int mod) {
    switch (mod) {
        case RIA_D8PSK: case RIA_QAM8: return 1.1f;
        case RIA_QAM16: return 1.2f;
        case RIA_QAM32: return 1.5f;
        case RIA_QAM64: return 1.8f;
        case RIA_QAM256: return 2.5f;
        default: return 1.0f;
    }
}

// QAM32 max-log demap, soft_demap.hpp:68-121
// demapD8PSK (soft_demap.hpp:238-263): natural-binary 8-phase differential, sin-based LLRs
__device__ __noinline__ void demap_d8psk(float2 sym, float sym_mag, float2 prev, float nv, float* out) {
    const float2 diff = cmul(sym, cconj(prev));
    const float pd = atan2_rn(diff.y, diff.x);
    const float sp = sym_mag * cabs(prev);
    if (sp < 1e-6f) { out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; return; }      // neutral, NOT clipped to +-0.01
    const float conf = sp / (2.0f * nv);
    out[0] = clip_llr(conf * sin_rn(pd));
    out[1] = clip_llr(conf * sin_rn(2.0f * pd));
    out[2] = clip_llr(conf * sin_rn(4.0f * pd));
}

__device__ void demap_qam32(float2 sym, float nv, float* out) {
    const float I_LEVELS[4] = {-3, -1, 1, 3};
    const int I_GRAY[4] = {0, 1, 3, 2};
    const float Q_LEVELS[8] = {-7, -5, -3, -1, 1, 3, 5, 7};
    const int Q_GRAY[8] = {0, 1, 3, 2, 6, 7, 5, 4};
    const float scale = 0.1961161351381840f;
    const float sf = 2.0f / nv;
    float d0[5], d1[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) { d0[b] = 1e10f; d1[b] = 1e10f; }
    for (int qi = 0; qi < 8; ++qi) {
        for (int ii = 0; ii < 4; ++ii) {
            const float px = I_LEVELS[ii] * scale, py = Q_LEVELS[qi] * scale;
            const int bits = (Q_GRAY[qi] << 2) | I_GRAY[ii];
            const float dx = sym.x - px, dy = sym.y - py;
            const float dist = dx * dx + dy * dy;
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                if (bits & (1 << (4 - b))) { if (dist < d1[b]) d1[b] = dist; }
                else                       { if (dist < d0[b]) d0[b] = dist; }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 5; ++b) out[b] = clip_llr(sf * (d1[b] - d0[b]));
}

// fp32 sum of a[0..n) in index order (the order is observable).  One lane walks the chain;
// loads come four at a time.
__device__ __forceinline__ float ordered_sum(const float* a, int n) {
    float s = 0.f;
    int i = 0;
#pragma unroll 2
    for (; i + 4 <= n; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(a + i);
        s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
    }
#pragma unroll 1
    for (; i < n; ++i) s = __fadd_rn(s, a[i]);
    return s;
}
__device__ __forceinline__ int count_flags(const int* a, int n) {
    int s = 0, i = 0;
#pragma unroll 1
    for (; i + 4 <= n; i += 4) { const int4 v = *reinterpret_cast<const int4*>(a + i); s += v.x + v.y + v.z + v.w; }
#pragma unroll 1
    for (; i < n; ++i) s += a[i];
    return s;
}
__device__ __forceinline__ float2 ordered_csum(const float2* a, int n) {
    float2 s = make_float2(0.f, 0.f);
    int i = 0;
#pragma unroll 1
    for (; i + 2 <= n; i += 2) {
        const float4 v = *reinterpret_cast<const float4*>(a + i);
        s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y);
        s.x = __fadd_rn(s.x, v.z); s.y = __fadd_rn(s.y, v.w);
    }
#pragma unroll 1
    for (; i < n; ++i) s = cadd(s, a[i]);
    return s;
}

// Several ordered sums of the same length at once: lane r < n_rows walks row r (rows are
// kMaxPilots floats apart), so k sums cost one pass instead of k.  Result of row r in lane r.
__device__ __forceinline__ float ordered_sum_rows(const float* rows, int n_rows, int n, int lane) {
    float s = 0.f;
    if (lane < n_rows) {
        const float* a = rows + lane * kMaxPilots;
#pragma unroll 2
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, a[i]);
    }
    return s;
}

// sqrt(var) / mean of n magnitudes with the reference's loops (mean first, then the squared deviations, both in
// index order): last_fading_index after the LTS (channel_equalizer.cpp:623-638) and after a pilot update (:1007-1023)
__device__ __noinline__ float magnitude_cv(const float* mags, int n) {
    float mean = 0.0f;
    for (int i = 0; i < n; ++i) mean = __fadd_rn(mean, mags[i]);
    mean = __fdiv_rn(mean, static_cast<float>(n));
    float var = 0.0f;
    for (int i = 0; i < n; ++i) { const float d = __fsub_rn(mags[i], mean); var = __fadd_rn(var, __fmul_rn(d, d)); }
    var = __fdiv_rn(var, static_cast<float>(n));
    return (mean > 0.01f) ? __fdiv_rn(sqrtf(var), mean) : 0.0f;
}

// =============================== carrier-domain processing =================================
// Everything below works on cs.bin[] (the carrier bins of the current symbol) with a group of
// G threads, g = index inside the group.  It is shared by the warp-per-frame carrier kernel
// (G = 32) and the monolithic kernel (G = 128).

template <int G>
__device__ __forceinline__ void frame_reset(CarState& cs, const KernelArgs& a, long long f, int g) {
    // demodulator.cpp:1264-1300
    if (g == 0) {
        cs.s.cfo_hz = a.second_pass ? a.rerun_cfo[f - a.frame_begin] : (a.cfo_hz ? a.cfo_hz[f] : 0.0f);
        cs.s.cfo_phase = a.phase ? a.phase[f] : 0.0f;
        cs.s.phase_start = cs.s.cfo_phase;
        cs.s.noise_var = 0.1f;
        cs.s.snr_lin = 1.0f;
        cs.s.slope = 0.0f;
        cs.s.cpc = make_float2(1.0f, 0.0f);
        cs.s.cpc_init = 0;
        cs.s.snr_count = 0;
        cs.s.rerun = 0;
        cs.s.have_prev_pilot = 0;
        cs.s.have_dd = 0;
    }
    #pragma unroll 1
    for (int c = g; c < kMaxCarriers; c += G) cs.H[c] = make_float2(1.0f, 0.0f);
    gsync<G>();
}

// frames too short to hold the two training symbols: processPresynced returns false without output
template <int G>
__device__ __forceinline__ void frame_too_short(const KernelArgs& a, long long f, int g) {
    if (g == 0) {
        a.n_llr[f] = 0;
        if (a.snr_db) a.snr_db[f] = 0.0f;
        if (a.cfo_out) a.cfo_out[f] = a.cfo_hz ? a.cfo_hz[f] : 0.0f;
        if (a.fading) a.fading[f] = 0.0f;
    }
    float* llr_out = a.llr + f * a.llr_stride;
    #pragma unroll 1
    for (int i = g; i < a.llr_stride; i += G) llr_out[i] = 0.0f;
}

// H_s = rx / tx on LTS symbol s (channel_equalizer.cpp:264, 278)
template <int G>
__device__ __forceinline__ void lts_symbol(CarState& cs, const OfdmCarrierTable& car, int s, int g) {
    const int nc = car.num_carriers;
    #pragma unroll 1
    for (int c = g; c < nc; c += G) {
        const int sub = car.sub_idx[c];
        const float2 tx = car.is_pilot[c] ? make_float2(car.pilot_sign[sub], 0.0f) : car.tx_data[sub];
        cs.hps[s][c] = cdiv(cs.bin[c], tx);
    }
    gsync<G>();
}

// residual CFO from the phase rotation between the two LTS symbols (:304-382).  Returns true when
// the reference re-runs the LTS with the corrected CFO; cs.s.cfo_hz / cfo_phase are then updated.
template <int G>
__device__ __forceinline__ bool lts_residual(CarState& cs, const OfdmCarrierTable& car, const KernelArgs& a, int g) {
    const int nd = car.n_data;
    #pragma unroll 1
    for (int i = g; i < nd; i += G) {
        const int c = car.data_car[i];
        const float2 h0 = cs.hps[0][c], h1 = cs.hps[1][c];
        int ok = 0;
        if (cabs(h0) > 0.01f && cabs(h1) > 0.01f) {
            const float2 diff = cmul(h1, cconj(h0));
            const float mag = cabs(diff);
            if (mag > 1e-6f) { cs.tmpc[i] = cdivf(diff, mag); ok = 1; }
        }
        if (!ok) cs.tmpc[i] = make_float2(0.f, 0.f);       // adding zero leaves the ordered sum unchanged
        cs.flag[i] = ok;
    }
    gsync<G>();
    if (g == 0) {
        const float2 sum = ordered_csum(cs.tmpc, nd);
        const int valid = count_flags(cs.flag, nd);
        int rerun = 0;
        if (valid > 10) {
            const float avg_phase = atan2_rn(sum.y, sum.x);
            const float symbol_duration = static_cast<float>(a.sym_len) / static_cast<float>(static_cast<unsigned>(a.sample_rate));
            const float residual = static_cast<float>(static_cast<double>(avg_phase) /
                                                      (2.0f * M_PI * static_cast<double>(symbol_duration)));
            if (fabsf(residual) > 0.3f && fabsf(residual) < 5.0f) {
                cs.s.cfo_hz = __fadd_rn(cs.s.cfo_hz, residual);
                cs.s.cfo_phase = cs.s.phase_start;            // :341
                rerun = 1;
            }
        }
        cs.s.rerun = rerun;
    }
    gsync<G>();
    return cs.s.rerun != 0;
}

// channel estimate := last LTS symbol, phase slope, noise variance / SNR (:387-485, :642)
template <int G>
__device__ __forceinline__ void lts_finish(CarState& cs, const OfdmCarrierTable& car, const KernelArgs& a, long long f, int g) {
    const int nc = car.num_carriers, nd = car.n_data;
    #pragma unroll 1
    for (int c = g; c < nc; c += G) cs.H[c] = cs.hps[1][c];
    gsync<G>();
    // phase slope across adjacent carriers (:412-437)
    #pragma unroll 1
    for (int c = g; c < nc - 1; c += G) {
        const float2 h0 = cs.H[c], h1 = cs.H[c + 1];
        int ok = 0;
        if (cabs(h0) > 0.01f && cabs(h1) > 0.01f) {
            const float2 diff = cmul(h1, cconj(h0));
            const float mag = cabs(diff);
            if (mag > 1e-6f) { cs.tmpc[c] = cdivf(diff, mag); ok = 1; }
        }
        if (!ok) cs.tmpc[c] = make_float2(0.f, 0.f);
        cs.flag[c] = ok;
    }
    gsync<G>();
    if (g == 0) {
        const float2 sum = ordered_csum(cs.tmpc, nc - 1);
        const int cnt = count_flags(cs.flag, nc - 1);
        if (cnt > 0) cs.s.slope = carg(cdivf(sum, static_cast<float>(cnt)));
    }
    gsync<G>();
    // noise variance / SNR from the two LTS estimates (:457-485)
    #pragma unroll 1
    for (int i = g; i < nd; i += G) {
        const int c = car.data_car[i];
        const float2 h0 = cs.hps[0][c], h1 = cs.hps[1][c];
        int ok = 0;
        if (cabs(h0) > 1e-6f && cabs(h1) > 1e-6f) {
            cs.tmpf[i] = cnorm(csub(h1, h0));
            cs.tmpg[i] = __fdiv_rn(__fadd_rn(cnorm(h0), cnorm(h1)), 2.0f);
            ok = 1;
        } else { cs.tmpf[i] = 0.f; cs.tmpg[i] = 0.f; }
        cs.flag[i] = ok;
    }
    gsync<G>();
    if (g == 0) {
        const float noise_sum = ordered_sum(cs.tmpf, nd), signal_sum = ordered_sum(cs.tmpg, nd);
        const int cnt = count_flags(cs.flag, nd);
        if (cnt > 0) {
            const float nvar = noise_sum / (4.0f * cnt);
            const float sp = signal_sum / cnt;
            float snr = sp / std_max(nvar, 1e-10f);
            snr = std_max(3.16f, std_min(10000.0f, snr));
            cs.s.noise_var = nvar;
            cs.s.snr_lin = snr;
        }
        cs.s.snr_count = 2;                                       // :642
    }
    if (a.modulation == RIA_D8PSK) {                                  // LTS fading index (:619-639); only D8PSK reads it
        gsync<G>();
        #pragma unroll 1
        for (int i = g; i < nd; i += G) cs.tmpf[i] = cabs(cs.H[car.data_car[i]]);
        gsync<G>();
        if (g == 0) cs.s.last_fading = magnitude_cv(cs.tmpf, nd);
    }
    if (a.h_lts_tap)
        #pragma unroll 1
        for (int c = g; c < nc; c += G) reinterpret_cast<float2*>(a.h_lts_tap)[f * nc + c] = cs.H[c];
    gsync<G>();
    // The de-slope / re-slope phasors of the pilot interpolation (:897, :935) depend only on the
    // LTS slope and the carrier number: evaluate them once per frame.  hps[] is dead from here on.
    {
        const int np = car.n_pilot;
        float2* rot_p = cs.hps[0];
        float2* rot_d = cs.hps[1];
        #pragma unroll 1
        for (int i = g; i < np; i += G) rot_p[i] = cexpj(-cs.s.slope * static_cast<float>(car.car_k[car.pilot_car[i]]));
        #pragma unroll 1
        for (int i = g; i < nd; i += G) rot_d[i] = cexpj(cs.s.slope * static_cast<float>(car.car_k[car.data_car[i]]));
    }
    gsync<G>();
}

// One data symbol: updateChannelEstimate + equalize + demodulateSymbol on cs.bin[].
// `sd` = index of the data symbol (0 = first after the LTS), llr_out = soft bits of the frame.
// MOD >= 0 fixes the modulation at compile time (the carrier kernel is instantiated per modulation
// so that each instance only carries its own demapper); MOD < 0 reads it from the arguments.
__host__ __device__ constexpr bool mod_is_differential(int m) { return m == RIA_DBPSK || m == RIA_DQPSK || m == RIA_D8PSK; }
__host__ __device__ constexpr int mod_bits(int m) {
    return m == RIA_DBPSK || m == RIA_BPSK ? 1 : m == RIA_DQPSK || m == RIA_QPSK ? 2 : m == RIA_D8PSK || m == RIA_QAM8 ? 3
         : m == RIA_QAM16 ? 4 : m == RIA_QAM32 ? 5 : m == RIA_QAM64 ? 6 : m == RIA_QAM256 ? 8 : 0;
}

template <int G, int MOD>
__device__ __forceinline__ void data_symbol(CarState& cs, const OfdmCarrierTable& car, const KernelArgs& a,
                                            float* __restrict__ llr_out, int sd, int g) {
    const int nc = car.num_carriers, nd = car.n_data, np = car.n_pilot;
    const int mod = (MOD >= 0) ? MOD : a.modulation;
    const bool differential = (MOD >= 0) ? mod_is_differential(MOD) : (a.differential != 0);
    const int bpc = (MOD >= 0) ? mod_bits(MOD) : a.bits_per_carrier;
    const int llr_per_sym = nd * bpc;
    const bool first = (sd == 0);                             // soft_bits.empty()

    // ----------------- updateChannelEstimate (channel_equalizer.cpp:645-1043) -----------------
    if (np > 0) {
        float alpha;
        if (first) alpha = 1.0f; else if (differential) alpha = 0.5f; else alpha = 0.9f;
        // Reduction rows (kMaxPilots floats each, laid over tmpc/tmpf/tmpg): 0,1 = CPE sum re/im,
        // 2 = CPE weight, 3 = pilot power, 4 = temporal noise power.  Lane r of the first warp
        // walks row r in index order, so the five ordered sums cost one pass.
        float* red = reinterpret_cast<float*>(cs.tmpc);
        float cpe_x = 0.f, cpe_y = 0.f, cpe_w = 0.f;
        if (differential) {
            #pragma unroll 1
            for (int i = g; i < np; i += G) {
                const int c = car.pilot_car[i];
                // rx / (+-1, 0) (:687): libgcc's complex division by a unit real is a sign change
                const float2 rx = cs.bin[c];
                cs.pil_ls[i] = (car.pilot_sign[i] < 0.0f) ? make_float2(-rx.x, -rx.y) : rx;
            }
            gsync<G>();
            // carrier phase recovery on the first symbol that yields a usable average (:699-714)
            if (g == 0 && !cs.s.cpc_init) {
                const float2 h_sum = ordered_csum(cs.pil_ls, np);
                const float2 h_avg = cdivf(h_sum, static_cast<float>(np));
                const float avg_mag = cabs(h_avg);
                if (avg_mag > 0.01f) { cs.s.cpc = cdivf(cconj(h_avg), avg_mag); cs.s.cpc_init = 1; }
            }
            gsync<G>();
            #pragma unroll 1
            for (int i = g; i < np; i += G) {
                const float2 ch = cmul(cs.pil_ls[i], cs.s.cpc);
                cs.pil_ls[i] = ch;
                float term = 0.f;
                int ok = 0;
                if (cs.s.have_prev_pilot) {
                    const float2 ph = cs.prev_pilot[i];
                    if (cnorm(ph) > 1e-6f && cnorm(ch) > 1e-6f) { term = cnorm(csub(ch, ph)); ok = 1; }
                }
                red[3 * kMaxPilots + i] = cnorm(ch);
                red[4 * kMaxPilots + i] = term;
                cs.flag[i] = ok;
            }
        } else {
            // LS estimate at the pilots, CPE terms (common phase of pilot LS vs current H, :720-756),
            // pilot power and temporal noise terms (:778-800) in one pass over the pilots
            #pragma unroll 1
            for (int i = g; i < np; i += G) {
                const int c = car.pilot_car[i];
                const float2 rx = cs.bin[c];
                const float2 ls = (car.pilot_sign[i] < 0.0f) ? make_float2(-rx.x, -rx.y) : rx;   // rx / (+-1, 0), :687
                cs.pil_ls[i] = ls;
                const float2 h_old = cs.H[c];
                const float h_old_mag = cabs(h_old);
                float2 t = make_float2(0.f, 0.f);      // adding zero leaves the ordered sums unchanged
                float w = 0.f;
                if (h_old_mag > 0.01f) {
                    const float2 ratio = cmul(ls, cconj(h_old));
                    const float mag = cabs(ratio);
                    if (mag > 1e-6f) { t = cscale(cdivf(ratio, mag), h_old_mag); w = h_old_mag; }
                }
                float term = 0.f;
                int ok = 0;
                if (cs.s.have_prev_pilot) {
                    const float2 ph = cs.prev_pilot[i];
                    if (cnorm(ph) > 1e-6f && cnorm(ls) > 1e-6f) { term = cnorm(csub(ls, ph)); ok = 1; }
                }
                red[0 * kMaxPilots + i] = t.x;
                red[1 * kMaxPilots + i] = t.y;
                red[2 * kMaxPilots + i] = w;
                red[3 * kMaxPilots + i] = cnorm(ls);
                red[4 * kMaxPilots + i] = term;
                cs.flag[i] = ok;
            }
        }
        gsync<G>();
        if (g < 32) {       // the group's first warp (all of it)
            const float rs = ordered_sum_rows(red, 5, np, g);
            cpe_x = gshfl<G>(rs, 0);
            cpe_y = gshfl<G>(rs, 1);
            cpe_w = gshfl<G>(rs, 2);
            const float sp = gshfl<G>(rs, 3);
            float npow = gshfl<G>(rs, 4);
            if (g == 0) {
                // pilot power and temporal noise count (:778-800)
                cs.s.signal_power = sp / static_cast<float>(np);
                int ncount = count_flags(cs.flag, np);
                if (ncount == 0) { npow = cs.s.signal_power / 31.6f; ncount = 1; }
                // the SNR EMA only looks at (noise_count > 1) and (noise_power_sum > 0) (:1025-1040)
                cs.s.noise_count = (npow > 0.0f) ? ncount : 0;
                if (!differential) {
                    int apply = 0;
                    if (cpe_w > 0.01f) {
                        const float ph = atan2_rn(cpe_y, cpe_x);
                        if (fabsf(ph) > 0.001f) { cs.s.cpe = cexpj(ph); apply = 1; }
                    }
                    cs.s.apply_cpe = apply;
                }
            }
        }
        gsync<G>();
        if (!differential) {
            if (cs.s.apply_cpe)
                #pragma unroll 1
                for (int c = g; c < nc; c += G) cs.H[c] = cmul(cs.H[c], cs.s.cpe);
            gsync<G>();
        }
        // smoothed update at the pilots (:801-820)
        #pragma unroll 1
        for (int i = g; i < np; i += G) {
            const int c = car.pilot_car[i];
            const float2 h_old = cs.H[c];
            const float2 ls = cs.pil_ls[i];
            if (differential) {
                const float new_mag = alpha * cabs(ls) + (1.0f - alpha) * cabs(h_old);
                const float2 e = cexpj(carg(h_old));
                cs.H[c] = make_float2(new_mag * e.x, new_mag * e.y);              // std::polar
            } else {
                cs.H[c] = cadd(cscale(ls, alpha), cscale(h_old, 1.0f - alpha));
            }
            cs.prev_pilot[i] = ls;                                                // :882
        }
        gsync<G>();
        // interpolation to the data carriers (:885-957)
        if (!differential) {
            #pragma unroll 1
            for (int i = g; i < np; i += G) {
                cs.prev_eq[i] = cmul(cs.H[car.pilot_car[i]], cs.hps[0][i]);
            }
            gsync<G>();
            #pragma unroll 1
            for (int i = g; i < nd; i += G) {
                const int c = car.data_car[i];
                const int lo = car.interp_lo[i], hi = car.interp_hi[i];
                const float al = car.interp_alpha[i];
                float2 ih = make_float2(0.f, 0.f);
                if (lo >= 0 && hi >= 0) ih = cadd(cscale(cs.prev_eq[lo], 1.0f - al), cscale(cs.prev_eq[hi], al));
                else if (lo >= 0) ih = cs.prev_eq[lo];
                else if (hi >= 0) ih = cs.prev_eq[hi];
                float2 h = cmul(ih, cs.hps[1][i]);
                // decision-directed phase refinement from the previous symbol (:964-975)
                if (cs.s.have_dd && cs.s.snr_count >= 3) {
                    const float corr = cs.dd[i];
                    if (fabsf(corr) > 0.001f) h = cmul(h, cexpj(corr * 0.3f));
                }
                cs.H[c] = h;
            }
        } else {
            #pragma unroll 1
            for (int i = g; i < nd; i += G) {
                const int c = car.data_car[i];
                const int lo = car.interp_lo[i], hi = car.interp_hi[i];
                const float al = car.interp_alpha[i];
                float im = 0.0f;
                if (lo >= 0 && hi >= 0) {
                    const float m1 = cabs(cs.H[car.pilot_car[lo]]), m2 = cabs(cs.H[car.pilot_car[hi]]);
                    im = (1.0f - al) * m1 + al * m2;
                } else if (lo >= 0) im = cabs(cs.H[car.pilot_car[lo]]);
                else if (hi >= 0) im = cabs(cs.H[car.pilot_car[hi]]);
                const float2 e = cexpj(carg(cs.H[c]));
                cs.H[c] = make_float2(im * e.x, im * e.y);
            }
        }
        if (mod == RIA_D8PSK) {                                        // fading index from the pilot magnitudes (:1007-1023)
            gsync<G>();
            #pragma unroll 1
            for (int i = g; i < np; i += G) cs.tmpf[i] = cabs(cs.pil_ls[i]);
            gsync<G>();
            if (g == 0) cs.s.last_fading = magnitude_cv(cs.tmpf, np);
        }
        gsync<G>();     // every thread has read snr_count / have_dd before thread 0 moves them on
        if (g == 0) {
            if (!differential && cs.s.noise_count > 1) {          // :1035-1039
                float inst = cs.s.signal_power / std_max(cs.s.noise_var, 1e-6f);
                inst = std_max(0.1f, std_min(10000.0f, inst));
                cs.s.snr_lin = 0.3f * inst + (1.0f - 0.3f) * cs.s.snr_lin;
            }
            cs.s.snr_count += 1;
            cs.s.have_prev_pilot = 1;
        }
        gsync<G>();
    }

    // ----------------------- equalize (channel_equalizer.cpp:1259-1451) -----------------------
    #pragma unroll 1
    for (int i = g; i < nd; i += G) cs.tmpg[i] = cnorm(cs.H[car.data_car[i]]);
    gsync<G>();
    if (g == 0) {
        cs.s.avg_h_power = ordered_sum(cs.tmpg, nd) / static_cast<float>(nd);
    }
    gsync<G>();
    const bool dd_mod = !differential && (mod == RIA_QPSK || mod == RIA_BPSK || mod == RIA_QAM16 ||
                                          mod == RIA_QAM32 || mod == RIA_QAM64);
    // D8PSK on a fading channel takes the two-pass demapper (demodulator.cpp:286-296, TWO_PASS_FADING_THRESHOLD 0.30)
    const bool d8_two_pass = (mod == RIA_D8PSK) && cs.s.last_fading > 0.30f;
    #pragma unroll 1
    for (int i = g; i < nd; i += G) {
        const int c = car.data_car[i];
        const float2 rx = cs.bin[c], h = cs.H[c];
        const float h_power = cs.tmpg[i];
        const float fade_threshold = 0.25f * cs.s.avg_h_power;
        float2 e;
        float nvv;
        float e_mag = -1.0f;          // |e|, evaluated at most once (DD gate here, |eq| EMA below)
        if (differential) {
            float snv = cs.s.noise_var;
            if (snv < 1e-6f) snv = cs.s.avg_h_power / 31.6f;
            const float den = h_power + snv;
            if (den < 1e-10f) { e = make_float2(0.f, 0.f); nvv = 100.0f; }
            else { e = cdivf(cmul(rx, cconj(h)), den); nvv = snv / (h_power + snv); }
            if (h_power < fade_threshold) nvv = 100.0f;
            nvv = std_max(1e-6f, std_min(100.0f, nvv));
        } else {
            const float den = h_power + cs.s.noise_var;
            if (den < 1e-10f) { e = make_float2(0.f, 0.f); nvv = 100.0f; }
            else {
                e = cdivf(cmul(cconj(h), rx), den);
                nvv = cs.s.noise_var / den;
                nvv = std_max(1e-6f, std_min(100.0f, nvv));
            }
            if (h_power < fade_threshold) nvv = 100.0f;
            // decision-directed phase error for the next symbol (:1413-1448)
            if (dd_mod && cs.s.snr_count >= 2) {
                float mag_thr = 0.3f, ph_thr = 0.61f;
                if (mod == RIA_QAM16) { mag_thr = 0.25f; ph_thr = 0.44f; }
                else if (mod == RIA_QAM32 || mod == RIA_QAM64) { mag_thr = 0.20f; ph_thr = 0.35f; }
                float ddv = 0.0f;
                e_mag = cabs(e);
                if (!(e_mag < mag_thr)) {
                    const float2 dec = hard_decision(e, mod);
                    const float perr = carg(cmul(e, cconj(dec)));
                    if (fabsf(perr) < ph_thr) ddv = -perr;
                }
                cs.dd[i] = ddv;
            }
        }
        // ----------------------- demodulateSymbol (demodulator.cpp:208-508) -----------------------
        // (same carrier, same thread: the equalised symbol and its noise variance stay in registers)
        const float2 sym = e;
        // per-carrier |eq| EMA / variance (:240-254)
        const float mag = (e_mag >= 0.0f) ? e_mag : cabs(sym);
        float ema, var;
        if (first) { ema = mag; var = 0.0f; }
        else {
            ema = cs.ema[i]; var = cs.var[i];
            const float delta = mag - ema;
            ema += 0.3f * delta;
            var += 0.3f * (delta * delta - var);
        }
        cs.ema[i] = ema; cs.var[i] = var;
        float nv = nvv * ce_margin(mod);
        {
            const float mean_sq = ema * ema + 1e-6f;
            const float norm_var = var / mean_sq;
            nv *= (1.0f + 10.0f * norm_var);
        }
        float* out = llr_out + sd * llr_per_sym + i * bpc;
        switch (mod) {
            case RIA_DBPSK: {                                 // demapDBPSK, soft_demap.hpp:172-193
                const float2 prev = first ? make_float2(1.0f, 0.0f) : cs.prev_eq[i];
                const float2 diff = cmul(sym, cconj(prev));
                const float pd = atan2_rn(diff.y, diff.x);
                const float sp = mag * cabs(prev);
                float l = 0.0f;
                if (!(sp < 1e-6f)) {
                    const float dnv = 2.0f * nv;
                    const float conf = 2.0f * sp / dnv;
                    l = clip_llr(conf * cos_rn(pd));
                }
                out[0] = l;
                cs.prev_eq[i] = sym;
                break;
            }
            case RIA_DQPSK: {                                 // demapDQPSK, soft_demap.hpp:199-235
                const float2 prev = first ? make_float2(1.0f, 0.0f) : cs.prev_eq[i];
                const float2 diff = cmul(sym, cconj(prev));
                const float dmag = cabs(diff);
                float l0 = 0.0f, l1 = 0.0f;
                if (!(dmag < 1e-6f)) {
                    const float dnv = 2.0f * nv;
                    const float sp = mag * cabs(prev);
                    const float snr = sp / dnv;
                    const float scale = 2.0f * sqrtf(snr);
                    const float pi = 3.14159265358979f;
                    const float ph = atan2_rn(diff.y, diff.x);
                    l0 = clip_llr(scale * sin_rn(ph + pi / 4));
                    l1 = clip_llr(scale * (fabsf(diff.x) - fabsf(diff.y)) / dmag);
                }
                out[0] = l0; out[1] = l1;
                cs.prev_eq[i] = sym;
                break;
            }
            case RIA_D8PSK: {                                 // demapD8PSK, soft_demap.hpp:238-263
                if (d8_two_pass) {                            // needs the phase statistics of ALL carriers first
                    cs.tmpc[i] = sym;
                    cs.dd[i] = nv;
                    break;
                }
                const float2 prev = first ? make_float2(1.0f, 0.0f) : cs.prev_eq[i];
                demap_d8psk(sym, mag, prev, nv, out);
                cs.prev_eq[i] = sym;
                break;
            }
            case RIA_BPSK:                                    // soft_demap.hpp:37-39
                out[0] = clip_llr(-2.0f * sym.x / nv);
                break;
            case RIA_QPSK: {                                  // soft_demap.hpp:42-45
                const float scale = -2.0f * 0.7071067811865476f / nv;
                out[0] = clip_llr(sym.x * scale); out[1] = clip_llr(sym.y * scale);
                break;
            }
            case RIA_QAM16: {                                 // soft_demap.hpp:49-64
                const float scale = 2.0f / nv, T = 0.6324555320336759f;
                out[0] = clip_llr(-scale * sym.x);
                out[1] = clip_llr(scale * (fabsf(sym.x) - T));
                out[2] = clip_llr(-scale * sym.y);
                out[3] = clip_llr(scale * (fabsf(sym.y) - T));
                break;
            }
            case RIA_QAM32: {
                float l[5];
                demap_qam32(sym, nv, l);
#pragma unroll
                for (int b = 0; b < 5; ++b) out[b] = l[b];
                break;
            }
            case RIA_QAM64: {                                 // soft_demap.hpp:125-142
                const float scale = 2.0f / nv, D2 = 0.3086067f, D4 = 0.6172134f;
                const float I = sym.x, Q = sym.y;
                out[0] = clip_llr(-scale * I);
                out[1] = clip_llr(scale * (fabsf(I) - D4));
                out[2] = clip_llr(scale * (fabsf(fabsf(I) - D4) - D2));
                out[3] = clip_llr(-scale * Q);
                out[4] = clip_llr(scale * (fabsf(Q) - D4));
                out[5] = clip_llr(scale * (fabsf(fabsf(Q) - D4) - D2));
                break;
            }
            case RIA_QAM256: {                                // soft_demap.hpp:145-164
                const float scale = 2.0f / nv, D2 = 0.1290994f, D4 = 0.2581989f, D8 = 0.5163978f;
                const float I = sym.x, Q = sym.y;
                out[0] = clip_llr(-scale * I);
                out[1] = clip_llr(scale * (fabsf(I) - D8));
                out[2] = clip_llr(scale * (fabsf(fabsf(I) - D8) - D4));
                out[3] = clip_llr(scale * (fabsf(fabsf(fabsf(I) - D8) - D4) - D2));
                out[4] = clip_llr(-scale * Q);
                out[5] = clip_llr(scale * (fabsf(Q) - D8));
                out[6] = clip_llr(scale * (fabsf(fabsf(Q) - D8) - D4));
                out[7] = clip_llr(scale * (fabsf(fabsf(fabsf(Q) - D8) - D4) - D2));
                break;
            }
            default: break;
        }
    }
    if (d8_two_pass) {
        // demodulateD8PSKTwoPass (demodulator.cpp:533-624).  Pass 1: common phase error against the embedded DQPSK grid
        // (45, 135, 225, 315 degrees), a power-weighted circular mean whose three sums run in carrier order.
        gsync<G>();
        float* w_row = reinterpret_cast<float*>(cs.flag);
        #pragma unroll 1
        for (int i = g; i < nd; i += G) {
            const float2 eq = cs.tmpc[i];
            const float2 prev = first ? make_float2(1.0f, 0.0f) : cs.prev_eq[i];
            const float sp = cabs(eq) * cabs(prev);
            float ts = 0.0f, tc = 0.0f, tw = 0.0f;                 // adding zero leaves the ordered sums unchanged
            if (sp > 0.1f) {
                const float2 diff = cmul(eq, cconj(prev));
                const float phase = atan2_rn(diff.y, diff.x);
                const float pmo = static_cast<float>(static_cast<double>(phase) - M_PI / 4.0f);
                int q = static_cast<int>(round(static_cast<double>(pmo * 2.0f) / M_PI));
                q = ((q % 4) + 4) % 4;
                const float expected = static_cast<float>(q * M_PI / 2.0f + M_PI / 4.0f);
                float err = phase - expected;
                while (err > M_PI) err = static_cast<float>(err - 2 * M_PI);
                while (err < -M_PI) err = static_cast<float>(err + 2 * M_PI);
                ts = sp * sin_rn(err); tc = sp * cos_rn(err); tw = sp;
            }
            cs.tmpf[i] = ts; cs.tmpg[i] = tc; w_row[i] = tw;
        }
        gsync<G>();
        if (g == 0) {
            const float sin_sum = ordered_sum(cs.tmpf, nd), cos_sum = ordered_sum(cs.tmpg, nd), w_sum = ordered_sum(w_row, nd);
            const float mean_error = (w_sum > 0.1f) ? atan2_rn(sin_sum, cos_sum) : 0.0f;
            float2 corr = make_float2(1.0f, 0.0f);
            if (fabsf(mean_error) > 0.05f && fabsf(mean_error) < 0.26f) {      // half of errors between 3 and 15 degrees
                const float ce = mean_error * 0.5f;
                corr = make_float2(cos_rn(-ce), sin_rn(-ce));
            }
            cs.s.d8_corr = corr;
            cs.s.snr_count += 1;                                               // :292
        }
        gsync<G>();
        // Pass 2: rotate, demap, and keep the ROTATED symbol as the next differential reference (:588-612)
        #pragma unroll 1
        for (int i = g; i < nd; i += G) {
            const float2 prev = first ? make_float2(1.0f, 0.0f) : cs.prev_eq[i];
            const float2 csym = cmul(cs.tmpc[i], cs.s.d8_corr);
            demap_d8psk(csym, cabs(csym), prev, cs.dd[i], llr_out + sd * llr_per_sym + i * bpc);
            cs.prev_eq[i] = csym;
        }
    }
    gsync<G>();      // everybody has read snr_count / have_dd of this symbol
    if (g == 0 && dd_mod && cs.s.snr_count >= 2) cs.s.have_dd = 1;
    gsync<G>();
}

// per-frame outputs: soft-bit count, zero padding, SNR, CFO, fading index
template <int G>
__device__ __forceinline__ void frame_outputs(CarState& cs, const OfdmCarrierTable& car, const KernelArgs& a,
                                              long long f, int n_data_sym, int g) {
    const int nd = car.n_data;
    const int n_llr = (n_data_sym > 0 ? n_data_sym : 0) * nd * a.bits_per_carrier;
    float* llr_out = a.llr + f * a.llr_stride;
    #pragma unroll 1
    for (int i = n_llr + g; i < a.llr_stride; i += G) llr_out[i] = 0.0f;
    #pragma unroll 1
    for (int i = g; i < nd; i += G) cs.tmpf[i] = cabs(cs.H[car.data_car[i]]);
    gsync<G>();
    if (g == 0) {
        a.n_llr[f] = n_llr;
        if (a.snr_db) a.snr_db[f] = 10.0f * log10f(cs.s.snr_lin);                  // getEstimatedSNR
        if (a.cfo_out) a.cfo_out[f] = cs.s.cfo_hz;                                 // getFrequencyOffset
        if (a.fading) {                                                            // getFadingIndex, :1168-1199
            float sum = 0.f;
            for (int i = 0; i < nd; ++i) sum += cs.tmpf[i];
            const float mean = sum / static_cast<float>(nd);
            float fi = 0.0f;
            if (!(mean < 0.001f)) {
                float vs = 0.f;
                for (int i = 0; i < nd; ++i) { const float d = cs.tmpf[i] - mean; vs += d * d; }
                fi = sqrtf(vs / static_cast<float>(nd)) / mean;
            }
            a.fading[f] = fi;
        }
    }
    gsync<G>();
}

// ======================================= FFT ================================================
// The 1024-point tile lives in shared memory as float2, element j at fft_addr(j): the 16 rows of 64 elements
// (row = j9..j6) are shifted by 2 * (row & 7) + 17 * (row >> 3) elements (monotone, so rows do not overlap; 32 extra
// elements).  For each pass the 16 lanes of every half-warp then hit 16 different bank pairs, so all 64-bit
// loads/stores are conflict-free, and -- unlike an XOR swizzle -- every access of a pass is a per-thread base
// plus a compile-time offset (the eight elements of a thread stay in one row, or step one row at a time in
// pass 3: + 64 + 2 elements), which removes the address arithmetic from the per-frame loop.
constexpr int kFftTilePad = 32;
__device__ __forceinline__ int fft_addr(int j) { return j + 2 * ((j >> 6) & 7) + 17 * (j >> 9); }
constexpr int kFftRowStep = 64 + 2;        // fft_addr(j + 64) - fft_addr(j) while (j >> 6) & 7 < 7
// stage-major twiddle table: tw_L[k] = W[k * (1024 / L)], k < L/2, stored at offset L/2 - 2
__device__ __forceinline__ int tw_off(int L) { return (L >> 1) - 2; }

// Blackwell packed fp32 (FADD2 / FMUL2): two IEEE round-to-nearest operations per issue slot.
// ptxas contracts a packed multiply that feeds a packed add into FFMA2 even under --fmad=false,
// so products always go through a scalar add (tests/test_abi_cpu.py checks the SASS for FFMA).
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2s(float s, float2 b) {            // (s * b.x, s * b.y)
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(b.x), "f"(b.y));
    return r;
}
// same products and the same sums as cmul(w, b)
__device__ __forceinline__ float2 cmul_p(float2 w, float2 b) {
    const float2 m = mul2s(w.x, b), n = mul2s(w.y, b);
    return make_float2(__fsub_rn(m.x, n.y), __fadd_rn(m.y, n.x));
}

// one radix-2 DIT butterfly exactly as fft.cpp:115-119:  t = w * b;  b = a - t;  a = a + t
__device__ __forceinline__ void bfly(float2& a, float2& b, float2 w) {
    const float2 t = cmul_p(w, b);
    b = sub2(a, t);
    a = add2(a, t);
}
// k = 0: w = (1, -0); multiplying by it is value-exact, so the product is skipped
__device__ __forceinline__ void bfly0(float2& a, float2& b) {
    const float2 t = b;
    b = sub2(a, t);
    a = add2(a, t);
}
// one output of a butterfly: a + t (upper half of the block) or a - t (lower half).  a - t is
// a + (-w) * b exactly, so the sign is folded into the twiddle.
__device__ __forceinline__ float2 bfly_half(float2 a, float2 b, float2 w_signed) {
    return add2(a, cmul_p(w_signed, b));
}

struct FftTile {
    float2 x[kFft + kFftTilePad];
};

// Tile offsets (in elements) of the first element a thread touches in each pass; per-thread constants.  `pin` makes
// them opaque to the compiler so they stay in registers instead of being recomputed for every frame.
struct FftOffsets { int p1, p2, p3; };
__device__ __forceinline__ FftOffsets fft_offsets(int tid, bool pin) {
    const int lane = tid & 31, warp = tid >> 5;
    FftOffsets o;
    o.p1 = fft_addr(static_cast<int>(__brev(static_cast<unsigned>(tid)) >> 25) << 3);          // rows of 8 at brev7(tid)
    o.p2 = fft_addr(((lane >> 1) << 6) | (warp << 1) | (lane & 1));                             // j9..j6 | j2 j1 j0
    o.p3 = fft_addr(((warp >> 1) << 9) | ((warp & 1) << 5) | lane);                             // j9 | j5..j0
    if (pin) asm volatile("" : "+r"(o.p1), "+r"(o.p2), "+r"(o.p3));
    return o;
}

// twiddles of pass 2 (stages 16, 32, 64): per-thread constants, index = j2 j1 j0 of the thread
struct Pass2Tw { float2 w16, w32a, w32b, w64[4], w4, w8[3]; };       // + the (uniform) twiddles of stages 4 and 8
__device__ __forceinline__ Pass2Tw load_pass2_tw(const float2* __restrict__ tw, int tid) {
    const int lo = ((tid >> 5) << 1) | (tid & 1);
    Pass2Tw w;
    w.w16 = tw[tw_off(16) + lo];
    w.w32a = tw[tw_off(32) + lo]; w.w32b = tw[tw_off(32) + 8 + lo];
#pragma unroll
    for (int q = 0; q < 4; ++q) w.w64[q] = tw[tw_off(64) + 8 * q + lo];
    w.w4 = tw[tw_off(4) + 1];
#pragma unroll
    for (int q = 0; q < 3; ++q) w.w8[q] = tw[tw_off(8) + 1 + q];
    return w;
}

// passes 1 and 2 (stages L = 2 .. 64) on the 8 mixed samples each thread holds; v[t] = baseband
// sample (tid + 128 * brev3(t)) of the FFT window.  Leaves the stage-64 result in the tile.
__device__ __forceinline__ void fft_stages_2_to_64(FftTile& ft, const float2* __restrict__ tw, float2 (&v)[8], int tid,
                                                   const FftOffsets& fo, const Pass2Tw& w2) {
    // ---- pass 1: stages L = 2, 4, 8 on data[8g .. 8g+7], g = brev7(tid) ----
    bfly0(v[0], v[1]); bfly0(v[2], v[3]); bfly0(v[4], v[5]); bfly0(v[6], v[7]);
    {
        bfly0(v[0], v[2]); bfly(v[1], v[3], w2.w4);
        bfly0(v[4], v[6]); bfly(v[5], v[7], w2.w4);
        bfly0(v[0], v[4]); bfly(v[1], v[5], w2.w8[0]); bfly(v[2], v[6], w2.w8[1]); bfly(v[3], v[7], w2.w8[2]);
    }
    {
        float2* p = ft.x + fo.p1;
#pragma unroll
        for (int t = 0; t < 8; ++t) p[t] = v[t];
    }
    __syncthreads();
    // ---- pass 2: stages L = 16, 32, 64; thread owns bits j5..j3 ----
    {
        float2* p = ft.x + fo.p2;
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[u << 3];
        bfly(v[0], v[1], w2.w16); bfly(v[2], v[3], w2.w16); bfly(v[4], v[5], w2.w16); bfly(v[6], v[7], w2.w16);
        bfly(v[0], v[2], w2.w32a); bfly(v[4], v[6], w2.w32a);
        bfly(v[1], v[3], w2.w32b); bfly(v[5], v[7], w2.w32b);
        bfly(v[0], v[4], w2.w64[0]);
        bfly(v[1], v[5], w2.w64[1]);
        bfly(v[2], v[6], w2.w64[2]);
        bfly(v[3], v[7], w2.w64[3]);
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u << 3] = v[u];
    }
    __syncthreads();
}

// pass 3 (stages L = 128, 256, 512, all bins) and pass 4 (stage 1024, carriers only): the
// general path, used by the monolithic kernel and when the bins do not allow pruning.
__device__ __forceinline__ void fft_stages_128_to_1024_full(FftTile& ft, const float2* __restrict__ tw,
                                                            const OfdmCarrierTable& car, float2* __restrict__ bin_out,
                                                            int tid, const FftOffsets& fo) {
    const int lane = tid & 31, warp = tid >> 5;
    float2 v[8];
    {
        const int lo6 = ((warp & 1) << 5) | lane;                     // j5..j0
        float2* p = ft.x + fo.p3;
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[u * kFftRowStep];
        const float2 w128 = tw[tw_off(128) + lo6];
        bfly(v[0], v[1], w128); bfly(v[2], v[3], w128); bfly(v[4], v[5], w128); bfly(v[6], v[7], w128);
        const float2 w256a = tw[tw_off(256) + lo6], w256b = tw[tw_off(256) + 64 + lo6];
        bfly(v[0], v[2], w256a); bfly(v[4], v[6], w256a);
        bfly(v[1], v[3], w256b); bfly(v[5], v[7], w256b);
        bfly(v[0], v[4], tw[tw_off(512) + lo6]);
        bfly(v[1], v[5], tw[tw_off(512) + 64 + lo6]);
        bfly(v[2], v[6], tw[tw_off(512) + 128 + lo6]);
        bfly(v[3], v[7], tw[tw_off(512) + 192 + lo6]);
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u * kFftRowStep] = v[u];
    }
    __syncthreads();
    if (tid < car.num_carriers) {
        const int f = car.fft_idx[tid];
        const int k = f & 511;
        float2 x0 = ft.x[fft_addr(k)];
        float2 x1 = ft.x[fft_addr(k + 512)];
        bfly(x0, x1, tw[tw_off(1024) + k]);
        bin_out[tid] = (f < 512) ? x0 : x1;
    }
    __syncthreads();
}

// Per-thread constants of the pruned last four stages.  Thread (j9 = warp >> 1, r = j5..j0)
// follows the one bin k with k mod 64 == r (if any) through stages 128..512 of its half, and the
// j9 = 0 thread finishes stage 1024.  Bit log2(L/2) of k selects the a - t output of stage L;
// the twiddles carry that sign.
struct PrunedPlan {
    int carrier;            // logical carrier of bin k, -1 = residue unused
    float2 w128, w256, w512, w1024;
};

__device__ __forceinline__ float2 signed_tw(float2 w, bool neg) { return neg ? make_float2(-w.x, -w.y) : w; }

__device__ __forceinline__ PrunedPlan make_pruned_plan(const float2* __restrict__ tw, const short* res_car,
                                                       const short* res_k, int r) {
    PrunedPlan p;
    p.carrier = res_car[r];
    const int k = p.carrier >= 0 ? res_k[r] : r;
    p.w128 = signed_tw(tw[tw_off(128) + (k & 63)], k & 64);
    p.w256 = signed_tw(tw[tw_off(256) + (k & 127)], k & 128);
    p.w512 = signed_tw(tw[tw_off(512) + (k & 255)], k & 256);
    p.w1024 = signed_tw(tw[tw_off(1024) + (k & 511)], k & 512);
    return p;
}

__device__ __forceinline__ void fft_stages_128_to_1024_pruned(FftTile& ft, float2* xch, const PrunedPlan& p,
                                                              float2* __restrict__ bin_out, int tid, const FftOffsets& fo) {
    const int lane = tid & 31, warp = tid >> 5;
    const int r = ((warp & 1) << 5) | lane;
    const int j9 = warp >> 1;
    float2 x = make_float2(0.f, 0.f);
    if (p.carrier >= 0) {
        const float2* pt = ft.x + fo.p3;
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = pt[u * kFftRowStep];
        const float2 y0 = bfly_half(v[0], v[1], p.w128);
        const float2 y1 = bfly_half(v[2], v[3], p.w128);
        const float2 y2 = bfly_half(v[4], v[5], p.w128);
        const float2 y3 = bfly_half(v[6], v[7], p.w128);
        const float2 z0 = bfly_half(y0, y1, p.w256);
        const float2 z1 = bfly_half(y2, y3, p.w256);
        x = bfly_half(z0, z1, p.w512);
        if (j9) xch[r] = x;
    }
    __syncthreads();
    if (p.carrier >= 0 && j9 == 0) bin_out[p.carrier] = bfly_half(x, xch[r], p.w1024);
}

// mixer + CFO phase increment exactly as channel_equalizer.cpp:103
__device__ __forceinline__ float cfo_phase_inc(float cfo_hz, int sample_rate) {
    return static_cast<float>(-2.0f * M_PI * static_cast<double>(cfo_hz) /
                              static_cast<double>(static_cast<unsigned>(sample_rate)));
}

// ------------------------------- stage 0: CFO phase scan -----------------------------------
// One warp per frame: value of the fp32 phase accumulator (channel_equalizer.cpp:132-144) before
// every 32nd sample of the frame.  Frames without a usable CFO are skipped (never read).
__global__ void ofdm_phase_scan_kernel(const float* __restrict__ cfo_hz, const float* __restrict__ phase0,
                                       long long frame_begin, long long n_local, int n_blocks, int sample_rate,
                                       float2* __restrict__ block_phase /*[n_local][n_blocks]*/,
                                       const int* __restrict__ list, const unsigned int* __restrict__ list_count,
                                       const float* __restrict__ list_cfo) {
    long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    float cfo;
    if (list) {                                   // second pass: the listed frames with their corrected CFO
        if (w >= static_cast<long long>(*list_count)) return;
        w = list[w];
        cfo = list_cfo[w];
    } else {
        if (w >= n_local || !cfo_hz) return;
        cfo = cfo_hz[frame_begin + w];
    }
    const long long f = frame_begin + w;
    if (!(fabsf(cfo) > 0.01f)) return;
    const float inc = cfo_phase_inc(cfo, sample_rate);
    float base = phase0 ? phase0[f] : 0.0f;
    float2* out = block_phase + w * n_blocks;
    cfo_scan_frame(base, inc, n_blocks, out, lane);
}

// ------------------------------- stage 1: mix + FFT ----------------------------------------
struct FftSmem {
    float2 tw[kTwCount];
    float2 nco[kFft];           // conj(mixer phasor) over the FFT window of the current symbol
    FftTile ft;
    float2 xch[64];             // before the item loop: residue -> carrier / bin maps (short[64] each) of the pruned plan
    OfdmCarrierTable car;
};

// MODE 0: first pass, no CFO vector; 1: first pass with the CFO handed in; 2: second pass (listed
// frames, corrected CFO).  Separate instances keep the rotation code and the list indirection out
// of the plain transform (no spills at 64 registers).
// Five CTAs per SM: the kernel is bound by shared-memory wavefronts and the fp32 pipe, not by occupancy, so the
// registers go to what is constant per thread -- the mixer phasors of the item's symbol (16), the pass-2 twiddles
// (14), the pruned plan (8) and the tile offsets -- instead of being re-read from shared memory for every frame.
// The CFO variants are bound by instruction issue (glibc's sincosf per sample) and keep eight CTAs at 64 registers,
// re-reading those constants from shared memory.
template <bool PRUNED, int MODE>
__global__ void __launch_bounds__(kThreads, MODE == 0 ? 5 : 8)
ofdm_fft_kernel(const KernelArgs a) {
    constexpr bool kSecond = (MODE == 2);
    constexpr bool kCfo = (MODE != 0);
    constexpr bool kRegConst = (MODE == 0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FftSmem& sm = *reinterpret_cast<FftSmem*>(smem_raw);
    __shared__ unsigned int item_sh;
    const int tid = threadIdx.x;
    if (kSecond && *a.rerun_count == 0) return;               // the usual case: nothing to re-run

    for (int i = tid; i < kTwCount; i += kThreads) sm.tw[i] = a.tw_g[i];
    {
        const int* src = reinterpret_cast<const int*>(a.car_g);
        int* dst = reinterpret_cast<int*>(&sm.car);
        for (int i = tid; i < static_cast<int>(sizeof(OfdmCarrierTable) / 4); i += kThreads) dst[i] = src[i];
    }
    short* res_car = reinterpret_cast<short*>(sm.xch);
    short* res_k = res_car + 64;
    if (tid < 64) { res_car[tid] = -1; res_k[tid] = 0; }
    __syncthreads();
    const int nc = sm.car.num_carriers;
    if (tid < nc) {
        const int k = sm.car.fft_idx[tid];
        res_car[k & 63] = static_cast<short>(tid);
        res_k[k & 63] = static_cast<short>(k);
    }
    __syncthreads();
    PrunedPlan plan;
    if (PRUNED) plan = make_pruned_plan(sm.tw, res_car, res_k, (((tid >> 5) & 1) << 5) | (tid & 31));

    const int n_sym = a.frame_len / a.sym_len;
    // first pass: every frame of the chunk; second pass: the frames the carrier stage listed
    const long long n_local = kSecond ? static_cast<long long>(*a.rerun_count) : a.frame_end - a.frame_begin;
    const unsigned n_groups = static_cast<unsigned>((n_local + kFftGroup - 1) / kFftGroup);
    const unsigned n_items = n_groups * static_cast<unsigned>(n_sym);
    const long long out_step = static_cast<long long>(n_sym) * nc;
    const FftOffsets fo = fft_offsets(tid, true);
    Pass2Tw w2{};
    if (kRegConst) w2 = load_pass2_tw(sm.tw, tid);      // sm.tw is complete: staged before the barriers above
    int cur_sym = -1;

    for (;;) {
        if (tid == 0) item_sh = atomicAdd(a.counter, 1u);
        __syncthreads();
        const unsigned item = item_sh;
        __syncthreads();
        if (item >= n_items) break;
        const int s = static_cast<int>(item / n_groups);
        const unsigned grp = item % n_groups;
        const int win = s * a.sym_len + a.cp;               // first sample of the FFT window
        if (s != cur_sym) {
            const float2* src = a.nco_g + win;
            for (int i = tid; i < kFft; i += kThreads) { const float2 o = __ldg(src + i); sm.nco[i] = make_float2(o.x, -o.y); }
            cur_sym = s;
            __syncthreads();
        }
        const long long i0 = static_cast<long long>(grp) * kFftGroup;
        const long long i1 = (i0 + kFftGroup < n_local) ? i0 + kFftGroup : n_local;
        auto frame_of = [&](long long i) -> long long { return a.frame_begin + (kSecond ? a.rerun_list[i] : i); };

        // first pass: the frames of an item are consecutive, so the sample and bin pointers just step
        const float* psrc = a.samples + frame_of(i0) * a.frame_stride + win + tid;
        float2* out = a.bins + ((frame_of(i0) - a.bins_frame0) * n_sym + s) * nc;
        float nxt[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) nxt[q] = __ldcs(psrc + 128 * q);
        float2 oscr[8];
        if (kRegConst) {
#pragma unroll
            for (int q = 0; q < 8; ++q) oscr[q] = sm.nco[tid + 128 * q];
        }
        for (long long i = i0; i < i1; ++i) {
            const long long f = frame_of(i);
            // ---- mix: v[t] = baseband sample (tid + 128 * brev3(t)) of the FFT window ----
            float2 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int q = ((t & 1) << 2) | (t & 2) | ((t & 4) >> 2);      // brev3
                const float2 osc = kRegConst ? oscr[q] : sm.nco[tid + 128 * q];
                // samples[i] * conj(osc)  ->  (osc.re * s, (-osc.im) * s), one packed multiply
                v[t] = mul2s(nxt[q], osc);
            }
            if (i + 1 < i1) {
                if (kSecond) psrc = a.samples + frame_of(i + 1) * a.frame_stride + win + tid;
                else psrc += a.frame_stride;
#pragma unroll
                for (int q = 0; q < 8; ++q) nxt[q] = __ldcs(psrc + 128 * q);
            }
            if (kCfo) {
                const float cfo = kSecond ? a.rerun_cfo[f - a.frame_begin] : a.cfo_hz[f];
                if (fabsf(cfo) > 0.01f) {
                    const float inc = cfo_phase_inc(cfo, a.sample_rate);
                    const float2* bp = a.block_phase + (f - a.frame_begin) * a.n_blocks;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int q = ((t & 1) << 2) | (t & 2) | ((t & 4) >> 2);
                        const int n = win + tid + 128 * q;
                        const float ph = cfo_block_phase(__ldg(bp + (n >> 5)), inc, n & 31);
                        float sn, cs;
                        glibc_sincosf_uniform(ph, &sn, &cs);
                        v[t] = cmul(v[t], make_float2(cs, sn));
                    }
                }
            }
            fft_stages_2_to_64(sm.ft, sm.tw, v, tid, fo, kRegConst ? w2 : load_pass2_tw(sm.tw, tid));
            if (PRUNED) fft_stages_128_to_1024_pruned(sm.ft, sm.xch, plan, out, tid, fo);
            else        fft_stages_128_to_1024_full(sm.ft, sm.tw, sm.car, out, tid, fo);
            if (kSecond) { if (i + 1 < i1) out = a.bins + ((frame_of(i + 1) - a.bins_frame0) * n_sym + s) * nc; }
            else out += out_step;
        }
    }
}

// ------------------------------- stage 2: carriers -----------------------------------------
struct CarSmem {
    OfdmCarrierTable car;
    CarState cs[kCarWarps];
};

template <int MOD>
__global__ void __launch_bounds__(kCarWarps * 32)
ofdm_carrier_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CarSmem& sm = *reinterpret_cast<CarSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.second_pass && *a.rerun_count == 0) return;
    {
        const int* src = reinterpret_cast<const int*>(a.car_g);
        int* dst = reinterpret_cast<int*>(&sm.car);
        for (int i = tid; i < static_cast<int>(sizeof(OfdmCarrierTable) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const OfdmCarrierTable& car = sm.car;
    CarState& cs = sm.cs[warp];
    const int nc = car.num_carriers;
    const int n_sym = a.frame_len / a.sym_len;
    const int n_data_sym = n_sym - 2;

    for (;;) {
        long long f = -1;
        if (lane == 0) {
            const long long i = static_cast<long long>(atomicAdd(a.counter, 1u));
            if (a.second_pass) { if (i < static_cast<long long>(*a.rerun_count)) f = a.frame_begin + a.rerun_list[i]; }
            else if (a.frame_begin + i < a.frame_end) f = a.frame_begin + i;
        }
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f < 0) break;
        if (n_sym < 2) { frame_too_short<32>(a, f, lane); continue; }
        const float2* fb = a.bins + (f - a.bins_frame0) * n_sym * nc;

        frame_reset<32>(cs, a, f, lane);
        // ---- LTS: estimateChannelFromLTS (channel_equalizer.cpp:193-643) ----
        for (int s = 0; s < 2; ++s) {
            for (int c = lane; c < nc; c += 32) cs.bin[c] = fb[s * nc + c];
            __syncwarp();
            lts_symbol<32>(cs, car, s, lane);
        }
        // The reference checks the residual CFO only on its first LTS pass (:304-382); when it asks
        // for a re-run the whole frame is mixed again with the corrected CFO, so the frame goes
        // back to the phase-scan / FFT stages and returns here as a second-pass frame.
        if (!a.second_pass && lts_residual<32>(cs, car, a, lane)) {
            if (lane == 0) {
                const int fl = static_cast<int>(f - a.frame_begin);
                a.rerun_cfo[fl] = cs.s.cfo_hz;
                a.rerun_list[atomicAdd(a.rerun_count, 1u)] = fl;
            }
            continue;
        }
        lts_finish<32>(cs, car, a, f, lane);
        // ---- data symbols (demodulator.cpp:1361-1382) ----
        float* llr_out = a.llr + f * a.llr_stride;
        float2 nb0 = make_float2(0.f, 0.f), nb1 = nb0;
        if (n_data_sym > 0) {
            if (lane < nc) nb0 = fb[2 * nc + lane];
            if (lane + 32 < nc) nb1 = fb[2 * nc + lane + 32];
        }
        for (int sd = 0; sd < n_data_sym; ++sd) {
            if (lane < nc) cs.bin[lane] = nb0;
            if (lane + 32 < nc) cs.bin[lane + 32] = nb1;
            __syncwarp();
            if (sd + 1 < n_data_sym) {          // prefetch the next symbol's bins
                if (lane < nc) nb0 = fb[(3 + sd) * nc + lane];
                if (lane + 32 < nc) nb1 = fb[(3 + sd) * nc + lane + 32];
            }
            data_symbol<32, MOD>(cs, car, a, llr_out, sd, lane);
        }
        frame_outputs<32>(cs, car, a, f, n_data_sym, lane);
    }
}

// Two frames per warp (G = 16): 15 pilots fill one pass of a half-warp and 44 data carriers three, against one half-empty
// and two passes of a whole warp, and the serial sections (ordered sums, phase estimates on the group's first lane) run
// for two frames at once.  The halves advance in lock step through the frame loop; a half whose frame leaves early
// (too short, residual-CFO re-run, end of the batch) sits out the rest of the iteration.
constexpr int kCar2Warps = 2;       // 4 frames per CTA
// The carrier tables (2 KB, read-only, the same for every CTA) are read from global memory through L1 here: without
// them in shared memory a tenth CTA fits on the SM.
struct Car2Smem {
    CarState cs[2 * kCar2Warps];
};

template <int MOD>
__global__ void __launch_bounds__(kCar2Warps * 32)
ofdm_carrier2_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Car2Smem& sm = *reinterpret_cast<Car2Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, g = tid & 15, grp = tid >> 4;
    if (a.second_pass && *a.rerun_count == 0) return;
    const OfdmCarrierTable& car = *a.car_g;
    // The two groups of a warp touch the same element of their states in the same instruction.  sizeof(CarState) is
    // 96 mod 128 bytes, so neighbouring states would overlap in 8 of their 16 banks (2-way conflict on every 32-bit
    // access); two states apart the offset is 64 mod 128 bytes and the halves of the warp use disjoint banks.
    static_assert((2 * sizeof(CarState)) % 128 == 64, "group states of one warp must sit 16 banks apart");
    CarState& cs = sm.cs[((grp & 1) << 1) | (grp >> 1)];
    const int nc = car.num_carriers;
    const int n_sym = a.frame_len / a.sym_len;
    const int n_data_sym = n_sym - 2;
    constexpr int kPer = kMaxCarriers / 16;

    for (;;) {
        __syncwarp();
        long long f = -1;
        if (g == 0) {
            const long long i = static_cast<long long>(atomicAdd(a.counter, 1u));
            if (a.second_pass) { if (i < static_cast<long long>(*a.rerun_count)) f = a.frame_begin + a.rerun_list[i]; }
            else if (a.frame_begin + i < a.frame_end) f = a.frame_begin + i;
        }
        f = __shfl_sync(0xffffffffu, f, lane & 16);
        if (__all_sync(0xffffffffu, f < 0)) break;
        bool active = f >= 0;
        if (active && n_sym < 2) { frame_too_short<16>(a, f, g); active = false; }
        const float2* fb = a.bins + (active ? (f - a.bins_frame0) : 0) * n_sym * nc;
        if (active) {
            frame_reset<16>(cs, a, f, g);
            // ---- LTS: estimateChannelFromLTS (channel_equalizer.cpp:193-643) ----
            for (int s = 0; s < 2; ++s) {
                for (int c = g; c < nc; c += 16) cs.bin[c] = __ldcs(fb + s * nc + c);      // streamed: keep L1 for the tables
                gsync<16>();
                lts_symbol<16>(cs, car, s, g);
            }
            // residual-CFO re-run: the frame goes back to the phase-scan / FFT stages (see ofdm_carrier_kernel)
            if (!a.second_pass && lts_residual<16>(cs, car, a, g)) {
                if (g == 0) {
                    const int fl = static_cast<int>(f - a.frame_begin);
                    a.rerun_cfo[fl] = cs.s.cfo_hz;
                    a.rerun_list[atomicAdd(a.rerun_count, 1u)] = fl;
                }
                active = false;
            }
        }
        if (active) {
            lts_finish<16>(cs, car, a, f, g);
            // ---- data symbols (demodulator.cpp:1361-1382) ----
            float* llr_out = a.llr + f * a.llr_stride;
            float2 nb[kPer];
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                nb[q] = make_float2(0.f, 0.f);
                if (n_data_sym > 0 && g + 16 * q < nc) nb[q] = __ldcs(fb + 2 * nc + g + 16 * q);
            }
            for (int sd = 0; sd < n_data_sym; ++sd) {
#pragma unroll
                for (int q = 0; q < kPer; ++q) if (g + 16 * q < nc) cs.bin[g + 16 * q] = nb[q];
                gsync<16>();
                if (sd + 1 < n_data_sym) {          // prefetch the next symbol's bins
#pragma unroll
                    for (int q = 0; q < kPer; ++q) if (g + 16 * q < nc) nb[q] = __ldcs(fb + (3 + sd) * nc + g + 16 * q);
                }
                data_symbol<16, MOD>(cs, car, a, llr_out, sd, g);
            }
            frame_outputs<16>(cs, car, a, f, n_data_sym, g);
        }
    }
}

// ------------------------------- monolithic kernel -----------------------------------------
// FFT + carriers of one frame in one CTA.  Serves the frames the carrier kernel handed over
// (residual CFO re-run); `rerun_list == nullptr` runs every frame of [frame_begin, frame_end).
struct MonoSmem {
    float2 tw[kTwCount];
    FftTile ft;
    float cph[kMaxSymLen];
    OfdmCarrierTable car;
    CarState cs;
};

// Mix + CFO-correct + FFT one symbol (absolute symbol index `sym` inside the frame) and leave
// the carrier bins in sm.cs.bin[c].  channel_equalizer.cpp:99-187 + fft.cpp:96-128.
__device__ void mono_fft_symbol(MonoSmem& sm, const KernelArgs& a, const float* __restrict__ frame, int sym) {
    const int tid = threadIdx.x;
    const int base = sym * a.sym_len;
    const bool cfo_on = fabsf(sm.cs.s.cfo_hz) > 0.01f;

    if (cfo_on) {
        // freq_correction_phase is an fp32 accumulator with a double-promoted wrap
        // (channel_equalizer.cpp:103, 132-144); its rounding is part of the result
        if (tid < 32) {
            const float inc = cfo_phase_inc(sm.cs.s.cfo_hz, a.sample_rate);
            float ph0 = sm.cs.s.cfo_phase;
            for (int b = 0; b < a.sym_len; b += 32) {
                float next;
                const float mine = cfo_phase_block32(ph0, inc, tid, &next);
                if (b + tid < a.sym_len) sm.cph[b + tid] = mine;
                if (b + 32 <= a.sym_len) ph0 = next;
                else ph0 = __shfl_sync(0xffffffffu, mine, a.sym_len - b);     // ragged tail: phase before sample sym_len
            }
            if (tid == 0) sm.cs.s.cfo_phase = ph0;
        }
        __syncthreads();
    }

    float2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int q = ((t & 1) << 2) | (t & 2) | ((t & 4) >> 2);      // brev3
        const int i = a.cp + tid + 128 * q;                           // index inside the symbol
        const float s = __ldcs(frame + base + i);
        const float2 osc = __ldg(a.nco_g + base + i);
        // samples[i] * conj(osc)  ->  (osc.re * s, (-osc.im) * s)
        float2 m = make_float2(__fmul_rn(osc.x, s), __fmul_rn(-osc.y, s));
        if (cfo_on) m = cmul(m, cexpj(sm.cph[i]));
        v[t] = m;
    }
    const FftOffsets fo = fft_offsets(tid, false);
    fft_stages_2_to_64(sm.ft, sm.tw, v, tid, fo, load_pass2_tw(sm.tw, tid));
    fft_stages_128_to_1024_full(sm.ft, sm.tw, sm.car, sm.cs.bin, tid, fo);
}

__global__ void __launch_bounds__(kThreads, 4)
ofdm_presynced_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MonoSmem& sm = *reinterpret_cast<MonoSmem*>(smem_raw);
    __shared__ long long frame_sh;
    const int tid = threadIdx.x;
    if (a.rerun_list && *a.rerun_count == 0) return;          // the usual case: nothing handed over

    for (int i = tid; i < kTwCount; i += kThreads) sm.tw[i] = a.tw_g[i];
    {
        const int* src = reinterpret_cast<const int*>(a.car_g);
        int* dst = reinterpret_cast<int*>(&sm.car);
        for (int i = tid; i < static_cast<int>(sizeof(OfdmCarrierTable) / 4); i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    const OfdmCarrierTable& car = sm.car;
    CarState& cs = sm.cs;
    const int nc = car.num_carriers;
    const int n_sym = a.frame_len / a.sym_len;
    const int n_data_sym = n_sym - 2;

    for (;;) {
        if (tid == 0) {
            const unsigned i = atomicAdd(a.counter, 1u);
            long long f = -1;
            if (a.rerun_list) { if (i < *a.rerun_count) f = a.frame_begin + a.rerun_list[i]; }
            else if (a.frame_begin + i < a.frame_end) f = a.frame_begin + i;
            frame_sh = f;
        }
        __syncthreads();
        const long long f = frame_sh;
        __syncthreads();
        if (f < 0) break;
        const float* frame = a.samples + f * a.frame_stride;
        if (n_sym < 2) { frame_too_short<kThreads>(a, f, tid); continue; }

        frame_reset<kThreads>(cs, a, f, tid);
        // ---- LTS: estimateChannelFromLTS (channel_equalizer.cpp:193-643) ----
        for (int pass = 0; pass < 2; ++pass) {
            for (int s = 0; s < 2; ++s) {
                mono_fft_symbol(sm, a, frame, s);
                if (a.bins_tap && tid < nc)
                    reinterpret_cast<float2*>(a.bins_tap)[(f * n_sym + s) * nc + tid] = cs.bin[tid];
                lts_symbol<kThreads>(cs, car, s, tid);
            }
            if (pass == 1) break;
            if (!lts_residual<kThreads>(cs, car, a, tid)) break;
        }
        lts_finish<kThreads>(cs, car, a, f, tid);
        // ---- data symbols (demodulator.cpp:1361-1382) ----
        float* llr_out = a.llr + f * a.llr_stride;
        for (int sd = 0; sd < n_data_sym; ++sd) {
            mono_fft_symbol(sm, a, frame, 2 + sd);
            if (a.bins_tap && tid < nc)
                reinterpret_cast<float2*>(a.bins_tap)[(f * n_sym + 2 + sd) * nc + tid] = cs.bin[tid];
            data_symbol<kThreads, -1>(cs, car, a, llr_out, sd, tid);
        }
        frame_outputs<kThreads>(cs, car, a, f, n_data_sym, tid);
    }
}

// stage-major twiddle table from the reference twiddles
void build_stage_twiddles(const std::vector<float2>& W, std::vector<float2>& out) {
    out.assign(kTwCount, make_float2(0.f, 0.f));
    for (int L = 4; L <= kFft; L <<= 1)
        for (int k = 0; k < L / 2; ++k) out[(L >> 1) - 2 + k] = W[static_cast<size_t>(k) * (kFft / L)];
}

// output pruning needs every used bin to have its own residue mod 64
bool bins_allow_pruning(const OfdmCarrierTable& car) {
    bool seen[64] = {};
    for (int c = 0; c < car.num_carriers; ++c) {
        const int r = car.fft_idx[c] & 63;
        if (seen[r]) return false;
        seen[r] = true;
    }
    return true;
}

int ensure_ofdm_scratch(ria_ctx* ctx, size_t bytes) {
    if (bytes > ctx->ofdm_scratch_bytes) {
        if (ctx->ofdm_scratch) RIA_CUDA(ctx, cudaFree(ctx->ofdm_scratch));
        ctx->ofdm_scratch = nullptr; ctx->ofdm_scratch_bytes = 0;
        RIA_CUDA(ctx, cudaMalloc(&ctx->ofdm_scratch, bytes));
        ctx->ofdm_scratch_bytes = bytes;
    }
    return RIA_OK;
}

template <typename K>
int blocks_per_sm(ria_ctx* ctx, K kernel, int threads, size_t smem, int* out) {
    const void* key = reinterpret_cast<const void*>(kernel);
    auto hit = ctx->occ_cache.find(key);
    if (hit != ctx->occ_cache.end()) { *out = hit->second; return RIA_OK; }
    RIA_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    RIA_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RIA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kernel, threads, smem));
    if (*out < 1) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: kernel does not fit on this device");
    ctx->occ_cache[key] = *out;
    return RIA_OK;
}

}  // namespace

void ofdm_tables_free(OfdmTablesDev* t) {
    if (!t) return;
    if (t->twiddle) cudaFree(t->twiddle);
    if (t->twiddle_nat) cudaFree(t->twiddle_nat);
    if (t->nco) cudaFree(t->nco);
    if (t->car) cudaFree(t->car);
    delete t;
}

int ofdm_tables_dev(ria_ctx* ctx, const ria_modem_config& cfg, int need_nco, const OfdmTablesDev** out) {
    if (const char* err = ofdm_config_error(cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: %s", err);
    if (ofdm_symbol_samples(cfg) > kMaxSymLen) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: symbol too long");
    OfdmTablesDev* hit = nullptr;
    for (OfdmTablesDev* t : ctx->ofdm_tables)
        if (std::memcmp(&t->cfg, &cfg, sizeof cfg) == 0) { hit = t; break; }
    if (!hit) { hit = new OfdmTablesDev(); hit->cfg = cfg; ctx->ofdm_tables.push_back(hit); }
    if (!hit->ready || hit->nco_len < need_nco) {
        int nco_len = need_nco < 65536 ? 65536 : need_nco;
        OfdmTablesHost h;
        ofdm_build_tables(cfg, nco_len, h);
        std::vector<float2> tw;
        build_stage_twiddles(h.twiddle, tw);
        if (hit->twiddle) cudaFree(hit->twiddle);
        if (hit->twiddle_nat) cudaFree(hit->twiddle_nat);
        if (hit->nco) cudaFree(hit->nco);
        if (hit->car) cudaFree(hit->car);
        hit->twiddle = nullptr; hit->twiddle_nat = nullptr; hit->nco = nullptr; hit->car = nullptr; hit->ready = false;
        RIA_CUDA(ctx, cudaMalloc(&hit->twiddle, tw.size() * sizeof(float2)));
        RIA_CUDA(ctx, cudaMalloc(&hit->twiddle_nat, h.twiddle.size() * sizeof(float2)));
        RIA_CUDA(ctx, cudaMemcpy(hit->twiddle_nat, h.twiddle.data(), h.twiddle.size() * sizeof(float2), cudaMemcpyHostToDevice));
        RIA_CUDA(ctx, cudaMalloc(&hit->nco, h.nco.size() * sizeof(float2)));
        RIA_CUDA(ctx, cudaMalloc(&hit->car, sizeof(OfdmCarrierTable)));
        RIA_CUDA(ctx, cudaMemcpy(hit->twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        RIA_CUDA(ctx, cudaMemcpy(hit->nco, h.nco.data(), h.nco.size() * sizeof(float2), cudaMemcpyHostToDevice));
        RIA_CUDA(ctx, cudaMemcpy(hit->car, &h.car, sizeof(OfdmCarrierTable), cudaMemcpyHostToDevice));
        hit->car_host = h.car;
        hit->cp = h.cp; hit->sym_len = h.sym_len; hit->nco_len = nco_len;
        hit->ready = true;
    }
    *out = hit;
    return RIA_OK;
}

}  // namespace ria

extern "C" int ria_ofdm_presynced_batch_taps_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                                 const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                                 const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                                 float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                                 float* snr_db_dev, float* cfo_out_dev, float* fading_dev,
                                                 float* bins_dev, float* h_lts_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len < 0 || frame_stride < frame_len) return set_error(ctx, RIA_E_INVAL, "ofdm: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !llr_dev || !n_llr_dev) return set_error(ctx, RIA_E_INVAL, "ofdm: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const OfdmTablesDev* t = nullptr;
    int rc = ofdm_tables_dev(ctx, *cfg, frame_len, &t);
    if (rc != RIA_OK) return rc;
    const int n_sym = frame_len / t->sym_len;
    const int nd = t->car_host.n_data;
    const int nc = t->car_host.num_carriers;
    const int bpc = ofdm_bits_per_carrier(cfg->modulation);
    const int n_llr = (n_sym > 2 ? n_sym - 2 : 0) * nd * bpc;
    if (llr_stride < n_llr) return set_error(ctx, RIA_E_INVAL, "ofdm: llr_stride %d < %d soft bits per frame", llr_stride, n_llr);

    KernelArgs a{};
    a.samples = samples_dev; a.frame_stride = frame_stride; a.frame_len = frame_len;
    a.cfo_hz = cfo_hz_dev; a.phase = phase_dev; a.n_frames = n_frames;
    a.llr = llr_dev; a.llr_stride = llr_stride; a.n_llr = n_llr_dev;
    a.snr_db = snr_db_dev; a.cfo_out = cfo_out_dev; a.fading = fading_dev;
    a.bins_tap = bins_dev; a.h_lts_tap = h_lts_dev;
    a.tw_g = t->twiddle; a.nco_g = t->nco; a.car_g = t->car;
    a.cp = t->cp; a.sym_len = t->sym_len; a.modulation = static_cast<int>(cfg->modulation);
    a.differential = ofdm_is_differential(cfg->modulation) ? 1 : 0;
    a.bits_per_carrier = bpc; a.sample_rate = static_cast<int>(cfg->sample_rate);
    a.pruned = bins_allow_pruning(t->car_host) ? 1 : 0;
    cudaStream_t st = ctx->stream;

    int mono_per_sm = 0;
    rc = blocks_per_sm(ctx, ofdm_presynced_kernel, kThreads, sizeof(MonoSmem), &mono_per_sm);
    if (rc != RIA_OK) return rc;

    // RIA_OFDM_MONOLITHIC=1: every frame through the one-CTA-per-frame kernel (A/B testing)
    const char* mono_env = std::getenv("RIA_OFDM_MONOLITHIC");
    if (n_sym < 2 || (mono_env && mono_env[0] == '1')) {
        a.frame_begin = 0; a.frame_end = n_frames;
        a.counter = ctx->work_counter + 1;
        long long grid = static_cast<long long>(ctx->sm_count) * mono_per_sm;
        if (grid > n_frames) grid = n_frames;
        RIA_CUDA(ctx, cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
        time_begin(ctx, KK_OFDM_DEMOD);
        ofdm_presynced_kernel<<<static_cast<unsigned>(grid), kThreads, sizeof(MonoSmem), st>>>(a);
        time_end(ctx);
        RIA_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        return RIA_OK;
    }

    int car_per_sm = 0;
    void (*fft_kernels[3])(const KernelArgs);
    if (a.pruned) { fft_kernels[0] = ofdm_fft_kernel<true, 0>; fft_kernels[1] = ofdm_fft_kernel<true, 1>; fft_kernels[2] = ofdm_fft_kernel<true, 2>; }
    else          { fft_kernels[0] = ofdm_fft_kernel<false, 0>; fft_kernels[1] = ofdm_fft_kernel<false, 1>; fft_kernels[2] = ofdm_fft_kernel<false, 2>; }
    int fft_per_sm_mode[3] = {0, 0, 0};          // the plain variant runs fewer, register-heavier CTAs than the CFO ones
    for (int m = 0; m < 3; ++m) {
        rc = blocks_per_sm(ctx, fft_kernels[m], kThreads, sizeof(FftSmem), &fft_per_sm_mode[m]);
        if (rc != RIA_OK) return rc;
    }
    void (*carrier_kernel)(const KernelArgs) = nullptr;
    static const bool carrier_g32 = getenv("RIA_CARRIER_G32") != nullptr;      // A/B: one frame per warp
    const int car_threads = carrier_g32 ? kCarWarps * 32 : kCar2Warps * 32;
    if (!carrier_g32) switch (cfg->modulation) {
        case RIA_DBPSK:  carrier_kernel = ofdm_carrier2_kernel<RIA_DBPSK>; break;
        case RIA_DQPSK:  carrier_kernel = ofdm_carrier2_kernel<RIA_DQPSK>; break;
        case RIA_D8PSK:  carrier_kernel = ofdm_carrier2_kernel<RIA_D8PSK>; break;
        case RIA_BPSK:   carrier_kernel = ofdm_carrier2_kernel<RIA_BPSK>; break;
        case RIA_QPSK:   carrier_kernel = ofdm_carrier2_kernel<RIA_QPSK>; break;
        case RIA_QAM16:  carrier_kernel = ofdm_carrier2_kernel<RIA_QAM16>; break;
        case RIA_QAM32:  carrier_kernel = ofdm_carrier2_kernel<RIA_QAM32>; break;
        case RIA_QAM64:  carrier_kernel = ofdm_carrier2_kernel<RIA_QAM64>; break;
        case RIA_QAM256: carrier_kernel = ofdm_carrier2_kernel<RIA_QAM256>; break;
        default: return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: modulation %u has no demapper", cfg->modulation);
    }
    else switch (cfg->modulation) {
        case RIA_DBPSK:  carrier_kernel = ofdm_carrier_kernel<RIA_DBPSK>; break;
        case RIA_DQPSK:  carrier_kernel = ofdm_carrier_kernel<RIA_DQPSK>; break;
        case RIA_D8PSK:  carrier_kernel = ofdm_carrier_kernel<RIA_D8PSK>; break;
        case RIA_BPSK:   carrier_kernel = ofdm_carrier_kernel<RIA_BPSK>; break;
        case RIA_QPSK:   carrier_kernel = ofdm_carrier_kernel<RIA_QPSK>; break;
        case RIA_QAM16:  carrier_kernel = ofdm_carrier_kernel<RIA_QAM16>; break;
        case RIA_QAM32:  carrier_kernel = ofdm_carrier_kernel<RIA_QAM32>; break;
        case RIA_QAM64:  carrier_kernel = ofdm_carrier_kernel<RIA_QAM64>; break;
        case RIA_QAM256: carrier_kernel = ofdm_carrier_kernel<RIA_QAM256>; break;
        default: return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: modulation %u has no demapper", cfg->modulation);
    }
    const size_t car_smem = carrier_g32 ? sizeof(CarSmem) : sizeof(Car2Smem);
    rc = blocks_per_sm(ctx, carrier_kernel, car_threads, car_smem, &car_per_sm);
    if (rc != RIA_OK) return rc;

    // chunk of frames whose carrier bins live in scratch between the stages
    const int64_t chunk = n_frames < 65536 ? n_frames : 65536;
    const int n_blocks = (n_sym * t->sym_len + 31) / 32;
    const size_t bins_b = bins_dev ? 0 : static_cast<size_t>(chunk) * n_sym * nc * sizeof(float2);
    const size_t phase_b = static_cast<size_t>(chunk) * n_blocks * sizeof(float2);
    const size_t list_b = static_cast<size_t>(chunk) * sizeof(int);
    rc = ensure_ofdm_scratch(ctx, bins_b + phase_b + 2 * list_b + 256);
    if (rc != RIA_OK) return rc;
    unsigned char* sp = static_cast<unsigned char*>(ctx->ofdm_scratch);
    float2* bins_scratch = reinterpret_cast<float2*>(sp);
    float2* block_phase = reinterpret_cast<float2*>(sp + bins_b);
    a.rerun_list = reinterpret_cast<int*>(sp + bins_b + phase_b);
    a.rerun_cfo = reinterpret_cast<float*>(sp + bins_b + phase_b + list_b);
    a.block_phase = block_phase;
    a.n_blocks = n_blocks;
    // [0] fft items, [1] carrier frames, [2] re-run count, [3] fft items and [4] carrier frames of the second pass
    unsigned int* ctr = ctx->work_counter + 8;
    a.rerun_count = ctr + 2;

    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = (n_frames - off < chunk) ? (n_frames - off) : chunk;
        a.frame_begin = off; a.frame_end = off + n;
        if (bins_dev) { a.bins = reinterpret_cast<float2*>(bins_dev); a.bins_frame0 = 0; }
        else          { a.bins = bins_scratch; a.bins_frame0 = off; }
        RIA_CUDA(ctx, cudaMemsetAsync(ctr, 0, 5 * sizeof(unsigned int), st));
        const long long items = ((n + kFftGroup - 1) / kFftGroup) * n_sym;
        long long car_grid = static_cast<long long>(ctx->sm_count) * car_per_sm;
        if (car_grid > (n + kCarWarps - 1) / kCarWarps) car_grid = (n + kCarWarps - 1) / kCarWarps;
        // pass 0: every frame with the CFO it was handed; pass 1: the frames whose LTS asked for
        // the residual-CFO re-run, with the corrected CFO (kernels exit at once when there are none)
        for (int pass = 0; pass < 2; ++pass) {
            a.second_pass = pass;
            if (pass == 1 || cfo_hz_dev) {
                time_begin(ctx, KK_OFDM_PHASE);
                ofdm_phase_scan_kernel<<<static_cast<unsigned>((n + 3) / 4), 128, 0, st>>>(
                    cfo_hz_dev, phase_dev, off, n, n_blocks, a.sample_rate, block_phase,
                    pass ? a.rerun_list : nullptr, a.rerun_count, a.rerun_cfo);
                time_end(ctx);
                ctx->launches += 1;
            }
            a.counter = ctr + (pass ? 3 : 0);
            const int mode = pass ? 2 : (cfo_hz_dev ? 1 : 0);
            long long fft_grid = static_cast<long long>(ctx->sm_count) * fft_per_sm_mode[mode];
            if (fft_grid > items) fft_grid = items;
            time_begin(ctx, KK_OFDM_FFT);
            fft_kernels[mode]<<<static_cast<unsigned>(fft_grid), kThreads, sizeof(FftSmem), st>>>(a);
            time_end(ctx);
            a.counter = ctr + (pass ? 4 : 1);
            time_begin(ctx, KK_OFDM_CARRIER);
            carrier_kernel<<<static_cast<unsigned>(car_grid), car_threads, car_smem, st>>>(a);
            time_end(ctx);
            ctx->launches += 2;
        }
        RIA_CUDA(ctx, cudaGetLastError());
    }
    return RIA_OK;
}

extern "C" int ria_ofdm_presynced_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                            const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                            const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                            float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                            float* snr_db_dev, float* cfo_out_dev, float* fading_dev) {
    return ria_ofdm_presynced_batch_taps_dev(ctx, cfg, samples_dev, frame_stride, frame_len, cfo_hz_dev, phase_dev,
                                             n_frames, llr_dev, llr_stride, n_llr_dev, snr_db_dev, cfo_out_dev,
                                             fading_dev, nullptr, nullptr);
}

extern "C" int ria_ofdm_presynced_batch_host(ria_ctx* ctx, const ria_modem_config* cfg,
                                             const float* samples, int64_t frame_stride, int32_t frame_len,
                                             const float* cfo_hz, const float* phase, int64_t n_frames,
                                             float* llr, int32_t llr_stride, int32_t* n_llr,
                                             float* snr_db, float* cfo_out, float* fading) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len < 0 || frame_stride < frame_len || llr_stride < 0) return set_error(ctx, RIA_E_INVAL, "ofdm: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !llr || !n_llr) return set_error(ctx, RIA_E_INVAL, "ofdm: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = 4096;
    const size_t in_b = static_cast<size_t>(chunk) * frame_len * 4;
    const size_t llr_b = static_cast<size_t>(chunk) * llr_stride * 4;
    const size_t aux_b = static_cast<size_t>(chunk) * 4 * 6;
    int rc = ensure_stage(ctx, 0, in_b + llr_b + aux_b + 1024, 0);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_samp = reinterpret_cast<float*>(base);
    float* d_llr = reinterpret_cast<float*>(base + in_b);
    float* d_aux = reinterpret_cast<float*>(base + in_b + llr_b);
    float *d_cfo = d_aux, *d_ph = d_aux + chunk, *d_snr = d_aux + 2 * chunk, *d_co = d_aux + 3 * chunk, *d_fa = d_aux + 4 * chunk;
    int32_t* d_nl = reinterpret_cast<int32_t*>(d_aux + 5 * chunk);
    cudaStream_t s = ctx->stream;
    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = (n_frames - off < chunk) ? (n_frames - off) : chunk;
        RIA_CUDA(ctx, cudaMemcpy2DAsync(d_samp, static_cast<size_t>(frame_len) * 4, samples + off * frame_stride,
                                        static_cast<size_t>(frame_stride) * 4, static_cast<size_t>(frame_len) * 4,
                                        static_cast<size_t>(n), cudaMemcpyHostToDevice, s));
        if (cfo_hz) RIA_CUDA(ctx, cudaMemcpyAsync(d_cfo, cfo_hz + off, n * 4, cudaMemcpyHostToDevice, s));
        if (phase) RIA_CUDA(ctx, cudaMemcpyAsync(d_ph, phase + off, n * 4, cudaMemcpyHostToDevice, s));
        rc = ria_ofdm_presynced_batch_dev(ctx, cfg, d_samp, frame_len, frame_len, cfo_hz ? d_cfo : nullptr,
                                          phase ? d_ph : nullptr, n, d_llr, llr_stride, d_nl, d_snr, d_co, d_fa);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(llr + off * llr_stride, d_llr, static_cast<size_t>(n) * llr_stride * 4, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(n_llr + off, d_nl, n * 4, cudaMemcpyDeviceToHost, s));
        if (snr_db) RIA_CUDA(ctx, cudaMemcpyAsync(snr_db + off, d_snr, n * 4, cudaMemcpyDeviceToHost, s));
        if (cfo_out) RIA_CUDA(ctx, cudaMemcpyAsync(cfo_out + off, d_co, n * 4, cudaMemcpyDeviceToHost, s));
        if (fading) RIA_CUDA(ctx, cudaMemcpyAsync(fading + off, d_fa, n * 4, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return RIA_OK;
}
