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
// Mapping (B200): one CTA of 128 threads per frame (persistent CTAs, atomic frame counter);
// the symbols of a frame are inherently sequential (channel estimate, EMA and differential
// reference carry over), so parallelism is frames x (samples | carriers) inside a symbol.
//   * samples are read from HBM exactly once, coalesced, straight into the FFT registers
//     (thread t owns samples t + 128 q, which is precisely the radix-2 DIT bit-reversed group it
//     needs for the first three stages); the cyclic prefix is never loaded;
//   * the FFT performs the reference's radix-2 DIT butterflies in the reference's order with
//     the reference's twiddle values, three stages per pass in registers, exchanging through
//     an XOR-swizzled shared-memory tile (conflict-free for every pass), so every used bin is
//     bit-identical to the reference FFT;  the last stage is evaluated only for the carriers;
//   * the 59 carriers are handled by threads 0..63; reductions whose fp32 summation order is
//     observable (they feed thresholds) are done in the reference's order.
// Float-order fidelity: compiled with --fmad=false; complex multiply/divide/abs follow what
// libstdc++/libgcc do on x86-64 (division and abs go through double, see cdiv/cabs below).

#include "ofdm_tables.h"

#include <cfloat>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

namespace {

constexpr int kThreads = 128;
constexpr int kFft = 1024;
constexpr int kMaxSymLen = 1160;    // fft 1024 + CP <= 128 + guard <= 8
constexpr int kTwCount = 1024;     // stage-major twiddle table (1022 used)

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
__device__ __forceinline__ float cabs(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}
// complex<float> / complex<float> = libgcc __divsc3: plain formula evaluated in double
// (verified bit-exact against the reference toolchain on 2M random inputs).
__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {
    const double aa = a.x, bb = a.y, cc = b.x, dd = b.y;
    const double den = cc * cc + dd * dd;
    return make_float2(static_cast<float>((aa * cc + bb * dd) / den),
                       static_cast<float>((bb * cc - aa * dd) / den));
}
// Transcendentals: the reference calls glibc's float cosf/sinf/atan2f, which are (almost always)
// correctly rounded; CUDA's float versions are 1-2 ulp.  Those ulps are amplified by the QAM
// demapper's 2/noise_var scale into LLR differences above the 1e-4 contract, so the device
// evaluates them in double and rounds once (B200 has half-rate FP64, this is affordable).
__device__ __forceinline__ float atan2_rn(float y, float x) { return static_cast<float>(atan2(static_cast<double>(y), static_cast<double>(x))); }
__device__ __forceinline__ float sin_rn(float x) { return static_cast<float>(sin(static_cast<double>(x))); }
__device__ __forceinline__ float cos_rn(float x) { return static_cast<float>(cos(static_cast<double>(x))); }
__device__ __forceinline__ float carg(float2 a) { return atan2_rn(a.y, a.x); }
__device__ __forceinline__ float2 cexpj(float th) {
    double s, c;
    sincos(static_cast<double>(th), &s, &c);
    return make_float2(static_cast<float>(c), static_cast<float>(s));
}
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }   // std::max(a,b)
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }   // std::min(a,b)

// clipLLR, src/ofdm/soft_demap.hpp:22-29
__device__ __forceinline__ float clip_llr(float llr) {
    float c = std_max(-20.0f, std_min(20.0f, llr));
    if (fabsf(c) < 0.01f) c = (c >= 0.0f) ? 0.01f : -0.01f;
    return c;
}

// ------------------------------- shared memory layout --------------------------------------
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
};

struct Smem {
    float2 tw[kTwCount];
    float re[kFft];
    float im[kFft];
    float cph[kMaxSymLen];
    OfdmCarrierTable car;
    float2 bin[kMaxCarriers];
    float2 H[kMaxCarriers];
    float2 hps[2][kMaxCarriers];
    float2 tmpc[kMaxCarriers];
    float tmpf[kMaxCarriers];
    float tmpg[kMaxCarriers];
    int   flag[kMaxCarriers];
    float2 pil_ls[kMaxCarriers];
    float2 prev_pilot[kMaxCarriers];
    float2 desloped[kMaxCarriers];
    float2 eq[kMaxCarriers];
    float2 prev_eq[kMaxCarriers];
    float cnv[kMaxCarriers];
    float hpow[kMaxCarriers];
    float ema[kMaxCarriers];
    float var[kMaxCarriers];
    float dd[kMaxCarriers];
    FrameScalars s;
};

// XOR-swizzled 32x32 transpose: conflict-free for all four FFT passes.
__device__ __forceinline__ int fft_addr(int j) { return ((j & 31) << 5) | (((j >> 5) ^ j) & 31); }
// stage-major twiddle table: tw_L[k] = W[k * (1024 / L)], k < L/2, stored at offset L/2 - 2
__device__ __forceinline__ int tw_off(int L) { return (L >> 1) - 2; }

// one radix-2 DIT butterfly exactly as fft.cpp:115-119:  t = w * b;  b = a - t;  a = a + t
__device__ __forceinline__ void bfly(float2& a, float2& b, float2 w) {
    const float2 t = cmul(w, b);
    b = csub(a, t);
    a = cadd(a, t);
}
// k = 0: w = (1, -0); multiplying by it is value-exact, so the product is skipped
__device__ __forceinline__ void bfly0(float2& a, float2& b) {
    const float2 t = b;
    b = csub(a, t);
    a = cadd(a, t);
}

struct KernelArgs {
    const float* samples; long long frame_stride; int frame_len;
    const float* cfo_hz; const float* phase; long long n_frames;
    float* llr; int llr_stride; int* n_llr;
    float* snr_db; float* cfo_out; float* fading;
    float* bins_tap; float* h_lts_tap;
    const float2* tw_g; const float2* nco_g; const OfdmCarrierTable* car_g;
    int cp, sym_len, modulation, differential, bits_per_carrier, sample_rate;
    unsigned int* counter;
};

// Mix + CFO-correct + FFT one symbol (absolute symbol index `sym` inside the frame) and leave
// the carrier bins in sm.bin[c].  channel_equalizer.cpp:99-187 + fft.cpp:96-128.
__device__ void fft_symbol(Smem& sm, const KernelArgs& a, const float* __restrict__ frame, int sym) {
    const int tid = threadIdx.x;
    const int base = sym * a.sym_len;
    const bool cfo_on = fabsf(sm.s.cfo_hz) > 0.01f;

    if (cfo_on) {
        // freq_correction_phase is an fp32 accumulator with a double-promoted wrap
        // (channel_equalizer.cpp:103, 132-144); its rounding is part of the result, so the scan
        // is done sequentially, once per symbol, by one thread.
        if (tid == 0) {
            const float inc = static_cast<float>(-2.0f * M_PI * static_cast<double>(sm.s.cfo_hz) /
                                                 static_cast<double>(static_cast<unsigned>(a.sample_rate)));
            float ph = sm.s.cfo_phase;
            for (int i = 0; i < a.sym_len; ++i) {
                sm.cph[i] = ph;
                ph = __fadd_rn(ph, inc);
                if (static_cast<double>(ph) > M_PI) ph = static_cast<float>(static_cast<double>(ph) - 2.0f * M_PI);
                else if (static_cast<double>(ph) < -M_PI) ph = static_cast<float>(static_cast<double>(ph) + 2.0f * M_PI);
            }
            sm.s.cfo_phase = ph;
        }
        __syncthreads();
    }

    // ---- load + mix: v[t] = baseband sample (tid + 128 * brev3(t)) of the FFT window ----
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
    // ---- pass 1: stages L = 2, 4, 8 on data[8g .. 8g+7], g = brev7(tid) ----
    bfly0(v[0], v[1]); bfly0(v[2], v[3]); bfly0(v[4], v[5]); bfly0(v[6], v[7]);
    {
        const float2 w1 = sm.tw[tw_off(4) + 1];
        bfly0(v[0], v[2]); bfly(v[1], v[3], w1);
        bfly0(v[4], v[6]); bfly(v[5], v[7], w1);
        const float2 x1 = sm.tw[tw_off(8) + 1], x2 = sm.tw[tw_off(8) + 2], x3 = sm.tw[tw_off(8) + 3];
        bfly0(v[0], v[4]); bfly(v[1], v[5], x1); bfly(v[2], v[6], x2); bfly(v[3], v[7], x3);
    }
    {
        const int g = __brev(static_cast<unsigned>(tid)) >> 25;      // brev7
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int ad = fft_addr((g << 3) | t);
            sm.re[ad] = v[t].x; sm.im[ad] = v[t].y;
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    // ---- pass 2: stages L = 16, 32, 64; thread owns bits j5..j3 ----
    {
        const int lo = (warp << 1) | (lane & 1);                      // j2 j1 j0
        const int hi = lane >> 1;                                     // j9..j6
        const int jb = (hi << 6) | lo;
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int ad = fft_addr(jb | (u << 3)); v[u] = make_float2(sm.re[ad], sm.im[ad]); }
        const float2 w16 = sm.tw[tw_off(16) + lo];
        if (lo == 0) { bfly0(v[0], v[1]); bfly0(v[2], v[3]); bfly0(v[4], v[5]); bfly0(v[6], v[7]); }
        else { bfly(v[0], v[1], w16); bfly(v[2], v[3], w16); bfly(v[4], v[5], w16); bfly(v[6], v[7], w16); }
        const float2 w32a = sm.tw[tw_off(32) + lo], w32b = sm.tw[tw_off(32) + 8 + lo];
        if (lo == 0) { bfly0(v[0], v[2]); bfly0(v[4], v[6]); } else { bfly(v[0], v[2], w32a); bfly(v[4], v[6], w32a); }
        bfly(v[1], v[3], w32b); bfly(v[5], v[7], w32b);
        if (lo == 0) bfly0(v[0], v[4]); else bfly(v[0], v[4], sm.tw[tw_off(64) + lo]);
        bfly(v[1], v[5], sm.tw[tw_off(64) + 8 + lo]);
        bfly(v[2], v[6], sm.tw[tw_off(64) + 16 + lo]);
        bfly(v[3], v[7], sm.tw[tw_off(64) + 24 + lo]);
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int ad = fft_addr(jb | (u << 3)); sm.re[ad] = v[u].x; sm.im[ad] = v[u].y; }
    }
    __syncthreads();
    // ---- pass 3: stages L = 128, 256, 512; thread owns bits j8..j6 ----
    {
        const int lo6 = ((warp & 1) << 5) | lane;                     // j5..j0
        const int jb = ((warp >> 1) << 9) | lo6;
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int ad = fft_addr(jb | (u << 6)); v[u] = make_float2(sm.re[ad], sm.im[ad]); }
        const float2 w128 = sm.tw[tw_off(128) + lo6];
        if (lo6 == 0) { bfly0(v[0], v[1]); bfly0(v[2], v[3]); bfly0(v[4], v[5]); bfly0(v[6], v[7]); }
        else { bfly(v[0], v[1], w128); bfly(v[2], v[3], w128); bfly(v[4], v[5], w128); bfly(v[6], v[7], w128); }
        const float2 w256a = sm.tw[tw_off(256) + lo6], w256b = sm.tw[tw_off(256) + 64 + lo6];
        if (lo6 == 0) { bfly0(v[0], v[2]); bfly0(v[4], v[6]); } else { bfly(v[0], v[2], w256a); bfly(v[4], v[6], w256a); }
        bfly(v[1], v[3], w256b); bfly(v[5], v[7], w256b);
        if (lo6 == 0) bfly0(v[0], v[4]); else bfly(v[0], v[4], sm.tw[tw_off(512) + lo6]);
        bfly(v[1], v[5], sm.tw[tw_off(512) + 64 + lo6]);
        bfly(v[2], v[6], sm.tw[tw_off(512) + 128 + lo6]);
        bfly(v[3], v[7], sm.tw[tw_off(512) + 192 + lo6]);
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int ad = fft_addr(jb | (u << 6)); sm.re[ad] = v[u].x; sm.im[ad] = v[u].y; }
    }
    __syncthreads();
    // ---- pass 4: stage L = 1024, only for the carriers ----
    if (tid < sm.car.num_carriers) {
        const int f = sm.car.fft_idx[tid];
        const int k = f & 511;
        const int a0 = fft_addr(k), a1 = fft_addr(k + 512);
        float2 x0 = make_float2(sm.re[a0], sm.im[a0]);
        float2 x1 = make_float2(sm.re[a1], sm.im[a1]);
        if (k == 0) bfly0(x0, x1); else bfly(x0, x1, sm.tw[tw_off(1024) + k]);
        sm.bin[tid] = (f < 512) ? x0 : x1;
    }
    __syncthreads();
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
            auto sl = [d](float x) {
                return x < -6 * d ? -7 * d : x < -4 * d ? -5 * d : x < -2 * d ? -3 * d : x < 0 ? -d
                     : x < 2 * d ? d : x < 4 * d ? 3 * d : x < 6 * d ? 5 * d : 7 * d; };
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

__global__ void __launch_bounds__(kThreads, 8)
ofdm_presynced_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;

    for (int i = tid; i < kTwCount; i += kThreads) sm.tw[i] = a.tw_g[i];
    {
        const int* src = reinterpret_cast<const int*>(a.car_g);
        int* dst = reinterpret_cast<int*>(&sm.car);
        for (int i = tid; i < static_cast<int>(sizeof(OfdmCarrierTable) / 4); i += kThreads) dst[i] = src[i];
    }
    __shared__ long long frame_sh;
    __syncthreads();

    const int nc = sm.car.num_carriers, nd = sm.car.n_data, np = sm.car.n_pilot;
    const int mod = a.modulation;
    const bool differential = a.differential != 0;
    const int bpc = a.bits_per_carrier;
    const int n_sym_total = a.frame_len / a.sym_len;
    const int n_data_sym = n_sym_total - 2;
    const int llr_per_sym = nd * bpc;

    for (;;) {
        if (tid == 0) frame_sh = static_cast<long long>(atomicAdd(a.counter, 1u));
        __syncthreads();
        const long long f = frame_sh;
        __syncthreads();
        if (f >= a.n_frames) break;
        const float* frame = a.samples + f * a.frame_stride;
        float* llr_out = a.llr + f * a.llr_stride;

        if (n_sym_total < 2) {       // processPresynced returns false without output
            if (tid == 0) {
                a.n_llr[f] = 0;
                if (a.snr_db) a.snr_db[f] = 0.0f;
                if (a.cfo_out) a.cfo_out[f] = a.cfo_hz ? a.cfo_hz[f] : 0.0f;
                if (a.fading) a.fading[f] = 0.0f;
            }
            for (int i = tid; i < a.llr_stride; i += kThreads) llr_out[i] = 0.0f;
            continue;
        }

        // ---- state reset (demodulator.cpp:1264-1300) ----
        if (tid == 0) {
            sm.s.cfo_hz = a.cfo_hz ? a.cfo_hz[f] : 0.0f;
            sm.s.cfo_phase = a.phase ? a.phase[f] : 0.0f;
            sm.s.phase_start = sm.s.cfo_phase;
            sm.s.noise_var = 0.1f;
            sm.s.snr_lin = 1.0f;
            sm.s.slope = 0.0f;
            sm.s.cpc = make_float2(1.0f, 0.0f);
            sm.s.cpc_init = 0;
            sm.s.snr_count = 0;
            sm.s.rerun = 0;
            sm.s.have_prev_pilot = 0;
            sm.s.have_dd = 0;
        }
        if (tid < kMaxCarriers) sm.H[tid] = make_float2(1.0f, 0.0f);
        __syncthreads();

        // =============== LTS: estimateChannelFromLTS (channel_equalizer.cpp:193-643) ===============
        for (int pass = 0; pass < 2; ++pass) {
            for (int s = 0; s < 2; ++s) {
                fft_symbol(sm, a, frame, s);
                if (tid < nc) {
                    const int sub = sm.car.sub_idx[tid];
                    const float2 tx = sm.car.is_pilot[tid] ? make_float2(sm.car.pilot_sign[sub], 0.0f)
                                                           : sm.car.tx_data[sub];
                    sm.hps[s][tid] = cdiv(sm.bin[tid], tx);          // H = rx / tx  (:264, :278)
                    if (a.bins_tap)
                        reinterpret_cast<float2*>(a.bins_tap)[(f * n_sym_total + s) * nc + tid] = sm.bin[tid];
                }
                __syncthreads();
            }
            if (pass == 1) break;
            // residual CFO from the phase rotation between the two LTS symbols (:304-382)
            if (tid < nd) {
                const int c = sm.car.data_car[tid];
                const float2 h0 = sm.hps[0][c], h1 = sm.hps[1][c];
                int ok = 0;
                if (cabs(h0) > 0.01f && cabs(h1) > 0.01f) {
                    const float2 diff = cmul(h1, cconj(h0));
                    const float mag = cabs(diff);
                    if (mag > 1e-6f) { sm.tmpc[tid] = cdivf(diff, mag); ok = 1; }
                }
                sm.flag[tid] = ok;
            }
            __syncthreads();
            if (tid == 0) {
                float2 sum = make_float2(0.f, 0.f);
                int valid = 0;
                for (int i = 0; i < nd; ++i) if (sm.flag[i]) { sum = cadd(sum, sm.tmpc[i]); ++valid; }
                int rerun = 0;
                if (valid > 10) {
                    const float avg_phase = atan2_rn(sum.y, sum.x);
                    const float symbol_duration = static_cast<float>(a.sym_len) / static_cast<float>(static_cast<unsigned>(a.sample_rate));
                    const float residual = static_cast<float>(static_cast<double>(avg_phase) /
                                                              (2.0f * M_PI * static_cast<double>(symbol_duration)));
                    if (fabsf(residual) > 0.3f && fabsf(residual) < 5.0f) {
                        sm.s.cfo_hz = __fadd_rn(sm.s.cfo_hz, residual);
                        sm.s.cfo_phase = sm.s.phase_start;            // :341
                        rerun = 1;
                    }
                }
                sm.s.rerun = rerun;
            }
            __syncthreads();
            if (!sm.s.rerun) break;
        }
        // channel estimate := last LTS symbol (:387-404)
        if (tid < nc) {
            sm.H[tid] = sm.hps[1][tid];
        }
        __syncthreads();
        // phase slope across adjacent carriers (:412-437)
        if (tid < nc - 1) {
            const float2 h0 = sm.H[tid], h1 = sm.H[tid + 1];
            int ok = 0;
            if (cabs(h0) > 0.01f && cabs(h1) > 0.01f) {
                const float2 diff = cmul(h1, cconj(h0));
                const float mag = cabs(diff);
                if (mag > 1e-6f) { sm.tmpc[tid] = cdivf(diff, mag); ok = 1; }
            }
            sm.flag[tid] = ok;
        }
        __syncthreads();
        if (tid == 0) {
            float2 sum = make_float2(0.f, 0.f);
            int cnt = 0;
            for (int i = 0; i < nc - 1; ++i) if (sm.flag[i]) { sum = cadd(sum, sm.tmpc[i]); ++cnt; }
            if (cnt > 0) sm.s.slope = carg(cdivf(sum, static_cast<float>(cnt)));
        }
        __syncthreads();
        // noise variance / SNR from the two LTS estimates (:457-485)
        if (tid < nd) {
            const int c = sm.car.data_car[tid];
            const float2 h0 = sm.hps[0][c], h1 = sm.hps[1][c];
            int ok = 0;
            if (cabs(h0) > 1e-6f && cabs(h1) > 1e-6f) {
                sm.tmpf[tid] = cnorm(csub(h1, h0));
                sm.tmpg[tid] = __fdiv_rn(__fadd_rn(cnorm(h0), cnorm(h1)), 2.0f);
                ok = 1;
            }
            sm.flag[tid] = ok;
        }
        __syncthreads();
        if (tid == 0) {
            float noise_sum = 0.f, signal_sum = 0.f;
            int cnt = 0;
            for (int i = 0; i < nd; ++i) if (sm.flag[i]) { noise_sum += sm.tmpf[i]; signal_sum += sm.tmpg[i]; ++cnt; }
            if (cnt > 0) {
                const float nvar = noise_sum / (4.0f * cnt);
                const float sp = signal_sum / cnt;
                float snr = sp / std_max(nvar, 1e-10f);
                snr = std_max(3.16f, std_min(10000.0f, snr));
                sm.s.noise_var = nvar;
                sm.s.snr_lin = snr;
            }
            sm.s.snr_count = 2;                                       // :642
        }
        if (a.h_lts_tap && tid < nc) reinterpret_cast<float2*>(a.h_lts_tap)[f * nc + tid] = sm.H[tid];
        __syncthreads();

        // ======================= data symbols (demodulator.cpp:1361-1382) =======================
        for (int sd = 0; sd < n_data_sym; ++sd) {
            fft_symbol(sm, a, frame, 2 + sd);
            if (a.bins_tap && tid < nc)
                reinterpret_cast<float2*>(a.bins_tap)[(f * n_sym_total + 2 + sd) * nc + tid] = sm.bin[tid];
            const bool first = (sd == 0);                             // soft_bits.empty()

            // ----------------- updateChannelEstimate (channel_equalizer.cpp:645-1043) -----------------
            if (np > 0) {
                float alpha;
                if (first) alpha = 1.0f; else if (differential) alpha = 0.5f; else alpha = 0.9f;
                if (tid < np) {
                    const int c = sm.car.pilot_car[tid];
                    sm.pil_ls[tid] = cdiv(sm.bin[c], make_float2(sm.car.pilot_sign[tid], 0.0f));   // :687
                }
                __syncthreads();
                if (differential) {
                    // carrier phase recovery on the first symbol that yields a usable average (:699-714)
                    if (tid == 0 && !sm.s.cpc_init) {
                        float2 h_sum = make_float2(0.f, 0.f);
                        for (int i = 0; i < np; ++i) h_sum = cadd(h_sum, sm.pil_ls[i]);
                        const float2 h_avg = cdivf(h_sum, static_cast<float>(np));
                        const float avg_mag = cabs(h_avg);
                        if (avg_mag > 0.01f) { sm.s.cpc = cdivf(cconj(h_avg), avg_mag); sm.s.cpc_init = 1; }
                    }
                    __syncthreads();
                    if (tid < np) sm.pil_ls[tid] = cmul(sm.pil_ls[tid], sm.s.cpc);
                    __syncthreads();
                } else {
                    // CPE: common phase of pilot LS vs current H, applied to every carrier (:720-756)
                    if (tid < np) {
                        const int c = sm.car.pilot_car[tid];
                        const float2 h_old = sm.H[c];
                        const float h_old_mag = cabs(h_old);
                        int ok = 0;
                        if (h_old_mag > 0.01f) {
                            const float2 ratio = cmul(sm.pil_ls[tid], cconj(h_old));
                            const float mag = cabs(ratio);
                            if (mag > 1e-6f) { sm.tmpc[tid] = cscale(cdivf(ratio, mag), h_old_mag); sm.tmpf[tid] = h_old_mag; ok = 1; }
                        }
                        sm.flag[tid] = ok;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        float2 cpe_sum = make_float2(0.f, 0.f);
                        float w = 0.f;
                        for (int i = 0; i < np; ++i) if (sm.flag[i]) { cpe_sum = cadd(cpe_sum, sm.tmpc[i]); w += sm.tmpf[i]; }
                        int apply = 0;
                        if (w > 0.01f) {
                            const float ph = carg(cpe_sum);
                            if (fabsf(ph) > 0.001f) { sm.s.cpe = cexpj(ph); apply = 1; }
                        }
                        sm.s.apply_cpe = apply;
                    }
                    __syncthreads();
                    if (sm.s.apply_cpe && tid < nc) sm.H[tid] = cmul(sm.H[tid], sm.s.cpe);
                    __syncthreads();
                }
                // pilot power, temporal noise count, smoothed update at the pilots (:778-820)
                if (tid == 0) {
                    float sp = 0.f;
                    for (int i = 0; i < np; ++i) sp += cnorm(sm.pil_ls[i]);
                    sm.s.signal_power = sp / static_cast<float>(np);
                    int ncount = 0;
                    float npow = 0.f;
                    if (sm.s.have_prev_pilot)
                        for (int i = 0; i < np; ++i) {
                            const float2 ph = sm.prev_pilot[i], ch = sm.pil_ls[i];
                            if (cnorm(ph) > 1e-6f && cnorm(ch) > 1e-6f) { npow += cnorm(csub(ch, ph)); ++ncount; }
                        }
                    if (ncount == 0) { npow = sm.s.signal_power / 31.6f; ncount = 1; }
                    // the SNR EMA only looks at (noise_count > 1) and (noise_power_sum > 0) (:1025-1040)
                    sm.s.noise_count = (npow > 0.0f) ? ncount : 0;
                }
                __syncthreads();
                if (tid < np) {
                    const int c = sm.car.pilot_car[tid];
                    const float2 h_old = sm.H[c];
                    const float2 ls = sm.pil_ls[tid];
                    if (differential) {
                        const float new_mag = alpha * cabs(ls) + (1.0f - alpha) * cabs(h_old);
                        const float2 e = cexpj(carg(h_old));
                        sm.H[c] = make_float2(new_mag * e.x, new_mag * e.y);              // std::polar
                    } else {
                        sm.H[c] = cadd(cscale(ls, alpha), cscale(h_old, 1.0f - alpha));
                    }
                    sm.prev_pilot[tid] = ls;                                              // :882
                }
                __syncthreads();
                // interpolation to the data carriers (:885-957)
                if (!differential) {
                    if (tid < np) {
                        const int c = sm.car.pilot_car[tid];
                        const float ph = -sm.s.slope * static_cast<float>(sm.car.car_k[c]);
                        sm.desloped[tid] = cmul(sm.H[c], cexpj(ph));
                    }
                    __syncthreads();
                    if (tid < nd) {
                        const int c = sm.car.data_car[tid];
                        const int lo = sm.car.interp_lo[tid], hi = sm.car.interp_hi[tid];
                        const float al = sm.car.interp_alpha[tid];
                        float2 ih = make_float2(0.f, 0.f);
                        if (lo >= 0 && hi >= 0) ih = cadd(cscale(sm.desloped[lo], 1.0f - al), cscale(sm.desloped[hi], al));
                        else if (lo >= 0) ih = sm.desloped[lo];
                        else if (hi >= 0) ih = sm.desloped[hi];
                        const float ph = sm.s.slope * static_cast<float>(sm.car.car_k[c]);
                        float2 h = cmul(ih, cexpj(ph));
                        // decision-directed phase refinement from the previous symbol (:964-975)
                        if (sm.s.have_dd && sm.s.snr_count >= 3) {
                            const float corr = sm.dd[tid];
                            if (fabsf(corr) > 0.001f) h = cmul(h, cexpj(corr * 0.3f));
                        }
                        sm.H[c] = h;
                    }
                } else {
                    if (tid < nd) {
                        const int c = sm.car.data_car[tid];
                        const int lo = sm.car.interp_lo[tid], hi = sm.car.interp_hi[tid];
                        const float al = sm.car.interp_alpha[tid];
                        float im = 0.0f;
                        if (lo >= 0 && hi >= 0) {
                            const float m1 = cabs(sm.H[sm.car.pilot_car[lo]]), m2 = cabs(sm.H[sm.car.pilot_car[hi]]);
                            im = (1.0f - al) * m1 + al * m2;
                        } else if (lo >= 0) im = cabs(sm.H[sm.car.pilot_car[lo]]);
                        else if (hi >= 0) im = cabs(sm.H[sm.car.pilot_car[hi]]);
                        const float2 e = cexpj(carg(sm.H[c]));
                        sm.H[c] = make_float2(im * e.x, im * e.y);
                    }
                }
                if (tid == 0) {
                    if (!differential && sm.s.noise_count > 1) {          // :1035-1039
                        float inst = sm.s.signal_power / std_max(sm.s.noise_var, 1e-6f);
                        inst = std_max(0.1f, std_min(10000.0f, inst));
                        sm.s.snr_lin = 0.3f * inst + (1.0f - 0.3f) * sm.s.snr_lin;
                    }
                    sm.s.snr_count += 1;
                    sm.s.have_prev_pilot = 1;
                }
                __syncthreads();
            }

            // ----------------------- equalize (channel_equalizer.cpp:1259-1451) -----------------------
            if (tid < nd) sm.hpow[tid] = cnorm(sm.H[sm.car.data_car[tid]]);
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int i = 0; i < nd; ++i) s += sm.hpow[i];
                sm.s.avg_h_power = s / static_cast<float>(nd);
            }
            __syncthreads();
            if (tid < nd) {
                const int c = sm.car.data_car[tid];
                const float2 rx = sm.bin[c], h = sm.H[c];
                const float h_power = sm.hpow[tid];
                const float fade_threshold = 0.25f * sm.s.avg_h_power;
                float2 e;
                float nvv;
                if (differential) {
                    float snv = sm.s.noise_var;
                    if (snv < 1e-6f) snv = sm.s.avg_h_power / 31.6f;
                    const float den = h_power + snv;
                    if (den < 1e-10f) { e = make_float2(0.f, 0.f); nvv = 100.0f; }
                    else { e = cdivf(cmul(rx, cconj(h)), den); nvv = snv / (h_power + snv); }
                    if (h_power < fade_threshold) nvv = 100.0f;
                    nvv = std_max(1e-6f, std_min(100.0f, nvv));
                } else {
                    const float den = h_power + sm.s.noise_var;
                    if (den < 1e-10f) { e = make_float2(0.f, 0.f); nvv = 100.0f; }
                    else {
                        e = cdivf(cmul(cconj(h), rx), den);
                        nvv = sm.s.noise_var / den;
                        nvv = std_max(1e-6f, std_min(100.0f, nvv));
                    }
                    if (h_power < fade_threshold) nvv = 100.0f;
                    // decision-directed phase error for the next symbol (:1413-1448)
                    const bool dd_mod = (mod == RIA_QPSK || mod == RIA_BPSK || mod == RIA_QAM16 ||
                                         mod == RIA_QAM32 || mod == RIA_QAM64);
                    if (dd_mod && sm.s.snr_count >= 2) {
                        float mag_thr = 0.3f, ph_thr = 0.61f;
                        if (mod == RIA_QAM16) { mag_thr = 0.25f; ph_thr = 0.44f; }
                        else if (mod == RIA_QAM32 || mod == RIA_QAM64) { mag_thr = 0.20f; ph_thr = 0.35f; }
                        float ddv = 0.0f;
                        if (!(cabs(e) < mag_thr)) {
                            const float2 dec = hard_decision(e, mod);
                            const float perr = carg(cmul(e, cconj(dec)));
                            if (fabsf(perr) < ph_thr) ddv = -perr;
                        }
                        sm.dd[tid] = ddv;
                    }
                }
                sm.eq[tid] = e;
                sm.cnv[tid] = nvv;
            }
            if (tid == 0 && !differential && sm.s.snr_count >= 2 &&
                (mod == RIA_QPSK || mod == RIA_BPSK || mod == RIA_QAM16 || mod == RIA_QAM32 || mod == RIA_QAM64))
                sm.s.have_dd = 1;
            __syncthreads();

            // ----------------------- demodulateSymbol (demodulator.cpp:208-508) -----------------------
            if (tid < nd) {
                const float2 sym = sm.eq[tid];
                // per-carrier |eq| EMA / variance (:240-254)
                const float mag = cabs(sym);
                float ema, var;
                if (first) { ema = mag; var = 0.0f; }
                else {
                    ema = sm.ema[tid]; var = sm.var[tid];
                    const float delta = mag - ema;
                    ema += 0.3f * delta;
                    var += 0.3f * (delta * delta - var);
                }
                sm.ema[tid] = ema; sm.var[tid] = var;
                float nv = sm.cnv[tid] * ce_margin(mod);
                {
                    const float mean_sq = ema * ema + 1e-6f;
                    const float norm_var = var / mean_sq;
                    nv *= (1.0f + 10.0f * norm_var);
                }
                float* out = llr_out + sd * llr_per_sym + tid * bpc;
                switch (mod) {
                    case RIA_DBPSK: {                                 // demapDBPSK, soft_demap.hpp:172-193
                        const float2 prev = first ? make_float2(1.0f, 0.0f) : sm.prev_eq[tid];
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
                        sm.prev_eq[tid] = sym;
                        break;
                    }
                    case RIA_DQPSK: {                                 // demapDQPSK, soft_demap.hpp:199-235
                        const float2 prev = first ? make_float2(1.0f, 0.0f) : sm.prev_eq[tid];
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
                        sm.prev_eq[tid] = sym;
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
            __syncthreads();
        }

        // ---- per-frame outputs ----
        const int n_llr = (n_data_sym > 0 ? n_data_sym : 0) * llr_per_sym;
        for (int i = n_llr + tid; i < a.llr_stride; i += kThreads) llr_out[i] = 0.0f;
        if (tid < nd) sm.tmpf[tid] = cabs(sm.H[sm.car.data_car[tid]]);
        __syncthreads();
        if (tid == 0) {
            a.n_llr[f] = n_llr;
            if (a.snr_db) a.snr_db[f] = 10.0f * log10f(sm.s.snr_lin);                  // getEstimatedSNR
            if (a.cfo_out) a.cfo_out[f] = sm.s.cfo_hz;                                 // getFrequencyOffset
            if (a.fading) {                                                            // getFadingIndex, :1168-1199
                float sum = 0.f;
                for (int i = 0; i < nd; ++i) sum += sm.tmpf[i];
                const float mean = sum / static_cast<float>(nd);
                float fi = 0.0f;
                if (!(mean < 0.001f)) {
                    float vs = 0.f;
                    for (int i = 0; i < nd; ++i) { const float d = sm.tmpf[i] - mean; vs += d * d; }
                    fi = sqrtf(vs / static_cast<float>(nd)) / mean;
                }
                a.fading[f] = fi;
            }
        }
        __syncthreads();
    }
}

// stage-major twiddle table from the reference twiddles
void build_stage_twiddles(const std::vector<float2>& W, std::vector<float2>& out) {
    out.assign(kTwCount, make_float2(0.f, 0.f));
    for (int L = 4; L <= kFft; L <<= 1)
        for (int k = 0; k < L / 2; ++k) out[(L >> 1) - 2 + k] = W[static_cast<size_t>(k) * (kFft / L)];
}

}  // namespace

void ofdm_tables_free(OfdmTablesDev* t) {
    if (!t) return;
    if (t->twiddle) cudaFree(t->twiddle);
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
        if (hit->nco) cudaFree(hit->nco);
        if (hit->car) cudaFree(hit->car);
        hit->twiddle = nullptr; hit->nco = nullptr; hit->car = nullptr; hit->ready = false;
        RIA_CUDA(ctx, cudaMalloc(&hit->twiddle, tw.size() * sizeof(float2)));
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
    a.counter = ctx->work_counter + 1;

    const size_t smem = sizeof(Smem);
    RIA_CUDA(ctx, cudaFuncSetAttribute(ofdm_presynced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    RIA_CUDA(ctx, cudaFuncSetAttribute(ofdm_presynced_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    RIA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ofdm_presynced_kernel, kThreads, smem));
    if (per_sm < 1) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: kernel does not fit");
    long long grid = static_cast<long long>(ctx->sm_count) * per_sm;
    if (grid > n_frames) grid = n_frames;
    RIA_CUDA(ctx, cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), ctx->stream));
    time_begin(ctx, KK_OFDM_DEMOD);
    ofdm_presynced_kernel<<<static_cast<unsigned>(grid), kThreads, smem, ctx->stream>>>(a);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
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
