// Batched MC-DPSK receive path for sm_100a: optional CFO correction (Hilbert FIR + rotation),
// per-(symbol, carrier) correlation demod, coherent 2x/4x despreading, differential decode and
// the two-pass LLR scaling.
//
// Replaces, per frame, MCDPSKWaveform::process (src/waveform/mc_dpsk_waveform.cpp:294-338)
//   = MultiCarrierDPSKDemodulator::setChirpDetected + process -> processGotChirp
//     (src/psk/multi_carrier_dpsk.hpp:323-340, 797-895), which runs
//       applyCFOCorrection   :901-926   (127-tap Hilbert, src/dsp/filters.cpp:266-317)
//       setReference         :507-518
//       demodulateOneSymbol  :931-946
//       demodulateSoft       :520-736
//   processTraining (:473-505) only refines cfo_hz_, which processGotChirp overwrites again when
//   the chirp was detected externally (:857-862) -- it has no effect on any output and is skipped.
//
// Kernels
//   mcdpsk_phase_scan_kernel  one warp per frame: phase of the fp32 CFO accumulator at every
//                             32-sample block boundary (exact, cfo_phase.cuh)
//   mcdpsk_cfo_kernel         Hilbert FIR + rotation -> corrected samples (HBM scratch)
//   mcdpsk_demod_kernel       one CTA per frame: correlate, despread, differential phases, LLRs
//
// Float-order fidelity: every per-(symbol, carrier) correlation is accumulated over the 512
// samples in the reference's order without FMA, the mixer phasors come from a host table built
// with the reference's fp32 recurrence, and the ordered reductions of demodulateSoft are
// evaluated in the reference's order.

#include "cfo_phase.cuh"
#include "rn_math.h"
#include "ria_internal.h"

#include <cmath>

namespace ria {

struct McdpskTablesDev {
    bool ready = false;
    ria_mcdpsk_config cfg{};
    float2* mixer = nullptr;     // [carriers][sps] = polar(1, -phase_i)
    float* hilbert = nullptr;    // 127 taps
};

namespace {

constexpr int kMaxCar = 16;
constexpr int kSps = 512;
constexpr int kSymPad = kSps + 4;           // floats per staged symbol row: 16-byte aligned, bank shift 4 per row
constexpr int kGroup = 24;               // rx symbols correlated per pass (multiple of 1, 2 and 4)
constexpr int kDemodThreads = 256;
constexpr int kHilbertTaps = 127;
constexpr int kHilbertDelay = 63;

__device__ __forceinline__ float cabs_d(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}
// std::arg / std::sin / std::cos on floats are glibc's atan2f / sinf / cosf: restated bit for bit in rn_math.h
__device__ __noinline__ float atan2_rn(float y, float x) { return glibc_atan2f(y, x); }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }

// ---------------------------------------------------------------------------------------------
// CFO correction
// ---------------------------------------------------------------------------------------------
__global__ void mcdpsk_phase_scan_kernel(const float* __restrict__ cfo_hz, const float* __restrict__ phase0,
                                         long long frame0, long long n_frames, int n_blocks, float sample_rate,
                                         float2* __restrict__ block_phase /*[n][n_blocks]*/) {
    const long long fl = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (fl >= n_frames) return;
    const long long f = frame0 + fl;
    const float cfo = cfo_hz[f];
    if (!(fabsf(cfo) > 0.1f)) return;
    // phase_inc = -2.0f * M_PI * cfo_hz / sample_rate   (multi_carrier_dpsk.hpp:911)
    const float inc = static_cast<float>(-2.0f * M_PI * static_cast<double>(cfo) / static_cast<double>(sample_rate));
    cfo_scan_frame(phase0 ? phase0[f] : 0.0f, inc, n_blocks, block_phase + fl * n_blocks, lane);
}

// HilbertTransform(127) taps (filters.cpp:266-291); only the even-indexed taps are non-zero.  In
// constant memory every tap is an immediate operand of its multiply once the loop is unrolled.
__constant__ float c_hilbert[kHilbertTaps + 1];

// A CTA covers 1024 consecutive samples of one frame (+126 samples of history); each thread
// produces four consecutive outputs.  q_i = sum_k coeffs[k] * x[i - k] in ascending k
// (filters.cpp:299-306) -- the odd taps are exactly zero and add nothing -- so the four chains
// run side by side: every tap needs one new pair of inputs (the window slides by two) and one
// broadcast load of the tap; the adds stay ordered.
__global__ void __launch_bounds__(256, 4)
mcdpsk_cfo_kernel(const float* __restrict__ samples, long long frame_stride, int frame_len,
                  const int* __restrict__ start, long long frame0,
                  const float* __restrict__ cfo_hz, const float2* __restrict__ block_phase, int n_blocks,
                  float sample_rate, float* __restrict__ out, long long out_stride) {
    constexpr int kHist = kHilbertTaps - 1;          // 126
    __shared__ __align__(16) float x[1024 + kHist + 2];
    __shared__ float taps[kHilbertTaps + 1];
    const long long fl = blockIdx.y;                 // frame inside the chunk
    const long long f = frame0 + fl;
    const int chunk0 = blockIdx.x * 1024;
    const float cfo = cfo_hz[f];
    const int st = start ? start[f] : 0;
    if (st < 0 || static_cast<long long>(st) + frame_len > frame_stride) return;   // demod reports 0 soft bits
    const float* src = samples + f * frame_stride + st;
    float* dst = out + fl * out_stride;
    const bool active = fabsf(cfo) > 0.1f && frame_len >= 128;      // :838, :903
    if (!active) return;                       // the demodulator reads such frames straight from the input
    if (threadIdx.x < kHilbertTaps) taps[threadIdx.x] = c_hilbert[threadIdx.x];
    for (int i = threadIdx.x; i < 1024 + kHist; i += 256) {
        const int g = chunk0 - kHist + i;
        x[i] = (g >= 0 && g < frame_len) ? src[g] : 0.0f;          // delay line starts at zero
    }
    __syncthreads();
    const float inc = static_cast<float>(-2.0f * M_PI * static_cast<double>(cfo) / static_cast<double>(sample_rate));
    const int li = 4 * threadIdx.x;                  // first of this thread's four local samples
    const int gi = chunk0 + li;
    if (gi >= frame_len) return;
    // window w = x[i - k .. i - k + 3] for the current tap k, i = li (index in x: + kHist)
    const float* xp = x + li + kHist;
    float2 w01 = *reinterpret_cast<const float2*>(xp), w23 = *reinterpret_cast<const float2*>(xp + 2);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int k = 0; k < kHilbertTaps; k += 2) {
        const float c = taps[k];                           // one broadcast load per tap
        q0 = __fadd_rn(q0, __fmul_rn(c, w01.x)); q1 = __fadd_rn(q1, __fmul_rn(c, w01.y));
        q2 = __fadd_rn(q2, __fmul_rn(c, w23.x)); q3 = __fadd_rn(q3, __fmul_rn(c, w23.y));
        if (k + 2 < kHilbertTaps) { w23 = w01; w01 = *reinterpret_cast<const float2*>(xp - k - 2); }
    }
    const float q[4] = {q0, q1, q2, q3};
    const float2* bp = block_phase + fl * n_blocks;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g = gi + j;
        if (g >= frame_len) break;
        const float re = xp[j - kHilbertDelay];            // delayed input (filters.cpp:309-310)
        const float ph = cfo_block_phase(bp[g >> 5], inc, g & 31);
        float sr, cr;
        glibc_sincosf_uniform(ph, &sr, &cr);               // Complex(std::cos(phase), std::sin(phase)), :913
        // real part of analytic * rotation (:916-917)
        dst[g] = __fsub_rn(__fmul_rn(re, cr), __fmul_rn(q[j], sr));
    }
}

// ---------------------------------------------------------------------------------------------
// demod
// ---------------------------------------------------------------------------------------------
struct DemodArgs {
    const float* samples; long long frame_stride; int frame_len;
    const int* start;          // per-frame first sample inside its row (nullptr = 0), only for uncorrected input
    long long row_len;         // samples available per row (start + frame_len must fit)
    long long frame0;          // first frame of this chunk
    const float* corrected;    // chunk-local CFO-corrected samples (frames with |cfo| > 0.1 Hz), or nullptr
    long long corrected_stride;
    const float* cfo_hz; long long n_frames;
    float* llr; int llr_stride; int* n_llr; float* fading; float* cfo_out;
    float* scratch;            // per frame: mags[max_ds][C] | phases[max_ds][C] | errs[max_ds][C]
    int max_ds;
    int stats_smem_off;        // > 0: the three per-frame arrays live in shared memory at this byte offset
                               // (the ordered reductions below walk them serially: from global memory that
                               // walk, not the correlations, was most of the kernel's time)
    const float2* mixer_g;
    int carriers, bits, spread, training;
    unsigned int* counter;
};

// Shared memory of the demodulator: the mixer table only takes the rows of the carriers in use,
// so that two CTAs fit on an SM for the 10-carrier modes.
struct DemodSmem {
    float sym[kGroup * kSymPad];
    float2 corr[kGroup * kMaxCar];
    float2 nrm[kGroup * kMaxCar];      // normalised data symbols of the current group
    float2 prev[kMaxCar];
    float car_sum[kMaxCar], car_sq[kMaxCar], rel[kMaxCar], car_mag[kMaxCar];
    float scale;
    int valid_symbols;
    long long frame;
    float2 mixer[1];           // [kSps][carriers] (sample-major, see correlate), allocated with the launch
};
static size_t demod_smem_bytes(int carriers) {
    return sizeof(DemodSmem) + static_cast<size_t>(carriers) * kSps * sizeof(float2);
}

// Packed fp32 multiply (Blackwell FMUL2): (s * m.x, s * m.y) in one issue slot.  The products go
// through scalar adds, because ptxas would contract a packed multiply feeding a packed add into
// FFMA2 (see ofdm.cu); the sums are the reference's `sum += samples[i] * mixer` (:940-943).
__device__ __forceinline__ float2 mul2s(float s, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(b.x), "f"(b.y));
    return r;
}
// 512-term complex correlation of one staged symbol with one carrier's mixer phasors, in sample
// order.  The kernel is bound by shared-memory wavefronts, not by issue slots: with one mixer row
// per carrier a warp's 128-bit phasor load touched ten rows and cost ~6 wavefronts (ncu), and
// there are twice as many phasor loads as sample loads.  The table is therefore sample-major,
// [sample][carrier]: at a given step every lane of the warp reads the same sample index, so a
// 64-bit phasor load covers one contiguous 8*C-byte span (one wavefront), and the sample loads are
// broadcasts.  Per four samples and warp: 4 + ~2 wavefronts instead of ~14.
__device__ __forceinline__ float2 correlate(const float* __restrict__ sy, const float2* __restrict__ mx, int C) {
    float2 acc = make_float2(0.f, 0.f);
    const float4* sy4 = reinterpret_cast<const float4*>(sy);
#pragma unroll 4
    for (int i = 0; i < kSps / 4; ++i) {
        const float4 v = sy4[i];
        const float2 m0 = mx[0], m1 = mx[C], m2 = mx[2 * C], m3 = mx[3 * C];
        mx += 4 * C;
        float2 p;
        p = mul2s(v.x, m0); acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y);
        p = mul2s(v.y, m1); acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y);
        p = mul2s(v.z, m2); acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y);
        p = mul2s(v.w, m3); acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y);
    }
    return acc;
}

__global__ void __launch_bounds__(kDemodThreads)
mcdpsk_demod_kernel(const DemodArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DemodSmem& sm = *reinterpret_cast<DemodSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int C = a.carriers;
    for (int i = tid; i < C * kSps; i += kDemodThreads) {
        const int c = i / kSps, k = i - c * kSps;
        sm.mixer[k * C + c] = a.mixer_g[i];
    }
    __syncthreads();

    const int preamble = (a.training + 1) * kSps;
    const int data_samples = a.frame_len - preamble;
    const int n_rx = data_samples > 0 ? data_samples / kSps : 0;
    int n_ds = n_rx / a.spread;
    if (n_ds < 1) n_ds = 1;                                         // :527
    const int n_out = n_ds * C * a.bits;

    for (;;) {
        if (tid == 0) sm.frame = static_cast<long long>(atomicAdd(a.counter, 1u));
        __syncthreads();
        const long long fl = sm.frame;
        __syncthreads();
        if (fl >= a.n_frames) break;
        const long long f = a.frame0 + fl;
        float* llr_out = a.llr + f * a.llr_stride;
        const int st = a.start ? a.start[f] : 0;
        const bool in_row = st >= 0 && static_cast<long long>(st) + a.frame_len <= a.row_len;
        if (a.frame_len <= preamble || n_ds > a.max_ds || !in_row) {   // processGotChirp waits for more samples
            if (tid == 0) { a.n_llr[f] = 0; if (a.fading) a.fading[f] = 0.0f; if (a.cfo_out) a.cfo_out[f] = a.cfo_hz ? a.cfo_hz[f] : 0.0f; }
            for (int i = tid; i < a.llr_stride; i += kDemodThreads) llr_out[i] = 0.0f;
            continue;
        }
        const bool corrected = a.corrected && fabsf(a.cfo_hz[f]) > 0.1f && a.frame_len >= 128;      // :838, :903
        const float* frame = corrected ? a.corrected + fl * a.corrected_stride : a.samples + f * a.frame_stride + st;
        float* mags = a.stats_smem_off ? reinterpret_cast<float*>(smem_raw + a.stats_smem_off)
                                       : a.scratch + fl * (3LL * a.max_ds * kMaxCar);
        float* phases = mags + a.max_ds * kMaxCar;
        float* errs = phases + a.max_ds * kMaxCar;

        // ---- setReference (:507-518): symbol right after the training ----
        for (int i = tid; i < kSps; i += kDemodThreads) sm.sym[i] = frame[a.training * kSps + i];
        __syncthreads();
        if (tid < C) {
            const float2 acc = correlate(sm.sym, sm.mixer + tid, C);
            float2 p = make_float2(__fdiv_rn(acc.x, static_cast<float>(kSps)), __fdiv_rn(acc.y, static_cast<float>(kSps)));
            const float m = cabs_d(p);
            if (m > 0.001f) { const float m2 = cabs_d(p); p = make_float2(__fdiv_rn(p.x, m2), __fdiv_rn(p.y, m2)); }
            else p = make_float2(1.0f, 0.0f);
            sm.prev[tid] = p;
        }
        __syncthreads();

        // ---- demodulateSoft pass 1 (:548-602): correlate, despread, differential phase ----
        for (int g0 = 0; g0 < n_ds * a.spread; g0 += kGroup) {
            int g_n = n_rx - g0;                                   // rx symbols available in this group
            if (g_n > kGroup) g_n = kGroup;
            if (g_n < 0) g_n = 0;
            for (int i = tid; i < g_n * kSps; i += kDemodThreads) {
                const int s = i / kSps, k = i - s * kSps;
                sm.sym[s * kSymPad + k] = __ldcs(frame + preamble + (g0 + s) * kSps + k);
            }
            __syncthreads();
            for (int p = tid; p < kGroup * C; p += kDemodThreads) {
                const int s = p / C, c = p - s * C;
                float2 acc = make_float2(0.f, 0.f);
                if (s < g_n) {
                    acc = correlate(sm.sym + s * kSymPad, sm.mixer + c, C);     // sum += samples[i] * mixer  (:940-943)
                    acc = make_float2(__fdiv_rn(acc.x, static_cast<float>(kSps)), __fdiv_rn(acc.y, static_cast<float>(kSps)));
                }
                sm.corr[s * kMaxCar + c] = acc;
            }
            __syncthreads();
            // combine repetitions, normalise, differential decode.  The chain through `prev` is not a
            // recurrence -- prev is just the previous symbol's normalised value -- so every (data symbol,
            // carrier) of the group is evaluated by its own thread: first the normalised symbols, then the
            // differences against the neighbour (the group's first symbol against sm.prev).
            const int ds0 = g0 / a.spread;
            int ds_n = kGroup / a.spread;
            if (ds_n > n_ds - ds0) ds_n = n_ds - ds0;
            for (int p = tid; p < ds_n * C; p += kDemodThreads) {
                const int d = p / C, c = p - d * C;
                float2 comb = make_float2(0.f, 0.f);
                for (int rep = 0; rep < a.spread; ++rep) {
                    const int rx = (ds0 + d) * a.spread + rep;
                    if (rx >= n_rx) break;
                    const float2 cur = sm.corr[(d * a.spread + rep) * kMaxCar + c];
                    comb.x = __fadd_rn(comb.x, cur.x); comb.y = __fadd_rn(comb.y, cur.y);
                }
                comb.x = __fdiv_rn(comb.x, static_cast<float>(a.spread));
                comb.y = __fdiv_rn(comb.y, static_cast<float>(a.spread));
                const float mag = cabs_d(comb);
                sm.nrm[d * kMaxCar + c] = (mag > 0.0001f) ? make_float2(__fdiv_rn(comb.x, mag), __fdiv_rn(comb.y, mag))
                                                          : make_float2(1.0f, 0.0f);
                mags[(ds0 + d) * kMaxCar + c] = mag;
            }
            __syncthreads();
            for (int p = tid; p < ds_n * C; p += kDemodThreads) {
                const int d = p / C, c = p - d * C;
                const float2 nrm = sm.nrm[d * kMaxCar + c];
                const float2 prev = (d == 0) ? sm.prev[c] : sm.nrm[(d - 1) * kMaxCar + c];
                // diff = normalized * conj(prev)
                const float2 diff = make_float2(__fsub_rn(__fmul_rn(nrm.x, prev.x), __fmul_rn(nrm.y, -prev.y)),
                                                __fadd_rn(__fmul_rn(nrm.x, -prev.y), __fmul_rn(nrm.y, prev.x)));
                const float phase = atan2_rn(diff.y, diff.x);
                const float PI = static_cast<float>(M_PI);
                float err;
                if (a.bits == 2) {
                    const float shifted = phase - PI / 4.0f;
                    const float idx = roundf(shifted / (PI / 2.0f));
                    const float ideal = idx * PI / 2.0f + PI / 4.0f;
                    err = phase - ideal;
                } else {
                    const float idx = roundf(phase / PI);
                    const float ideal = idx * PI;
                    err = phase - ideal;
                }
                while (err > PI) err -= 2.0f * PI;
                while (err < -PI) err += 2.0f * PI;
                const int o = (ds0 + d) * kMaxCar + c;
                phases[o] = phase; errs[o] = err * err;
            }
            __syncthreads();
            if (tid < C && ds_n > 0) sm.prev[tid] = sm.nrm[(ds_n - 1) * kMaxCar + tid];
            __syncthreads();
        }

        // ---- ordered reductions (:541-570, 604-634, 636-698) ----
        if (tid == 0) {
            // noise_sum over (data symbol, carrier) in order
            float noise_sum = 0.f;
            for (int d = 0; d < n_ds; ++d) for (int c = 0; c < C; ++c) noise_sum += errs[d * kMaxCar + c];
            const int noise_count = n_ds * C;
            // silence detection on per-symbol total magnitude
            int valid = n_ds;
            if (n_ds >= 4) {
                float ref_mag = 0.f;
                for (int s = 0; s < 4; ++s) { float t = 0.f; for (int c = 0; c < C; ++c) t += mags[s * kMaxCar + c]; ref_mag += t; }
                ref_mag /= 4.0f;
                if (ref_mag > 0.001f) {
                    const float thr = ref_mag * 0.2f;
                    while (valid > 4) {
                        float t = 0.f;
                        for (int c = 0; c < C; ++c) t += mags[(valid - 1) * kMaxCar + c];
                        if (t < thr) --valid; else break;
                    }
                }
            }
            sm.valid_symbols = valid;
            float var = (noise_count > 0) ? noise_sum / noise_count : 0.5f;
            var = std_max(0.01f, var);
            float scale = 2.0f * sqrtf(1.0f / var);
            sm.scale = std_min(scale, 20.0f);
        }
        __syncthreads();
        if (tid < C) {       // per-carrier sums over the valid symbols, in symbol order
            float s1 = 0.f, s2 = 0.f;
            for (int d = 0; d < sm.valid_symbols; ++d) { const float m = mags[d * kMaxCar + tid]; s1 += m; s2 += m * m; }
            sm.car_sum[tid] = s1; sm.car_sq[tid] = s2;
        }
        __syncthreads();
        if (tid == 0) {
            const int valid = sm.valid_symbols;
            for (int c = 0; c < C; ++c) sm.rel[c] = 1.0f;
            if (a.bits == 1 && valid > 0) {
                float gsum = 0.f; int gcnt = 0;
                for (int c = 0; c < C; ++c) {
                    const float mean = sm.car_sum[c] / valid;
                    sm.car_mag[c] = mean;
                    if (mean > 1e-4f) { gsum += mean; ++gcnt; }
                }
                const float gmean = (gcnt > 0) ? (gsum / gcnt) : 0.0f;
                for (int c = 0; c < C; ++c) {
                    const float mean = sm.car_mag[c];
                    if (mean <= 1e-4f || gmean <= 1e-4f) { sm.rel[c] = 0.12f; continue; }
                    const float mean_sq = sm.car_sq[c] / valid;
                    const float var = std_max(0.0f, mean_sq - mean * mean);
                    const float cv = sqrtf(var) / (mean + 1e-6f);
                    const float ratio = mean / gmean;
                    const float mag_w = std_max(0.10f, std_min(1.25f, ratio));
                    const float stab = 1.0f / (1.0f + 1.5f * cv);
                    float damp = 1.0f;
                    if (ratio < 0.20f) damp = 0.25f; else if (ratio < 0.35f) damp = 0.50f;
                    const float w = mag_w * stab * damp;
                    sm.rel[c] = std_max(0.12f, std_min(1.25f, w));
                }
            }
            // fading indices (:702-733, 403-432)
            if (a.fading) {
                float cm[kMaxCar];
                for (int c = 0; c < C; ++c) cm[c] = (valid > 0) ? sm.car_sum[c] / valid : 0.0f;
                float temporal = 0.0f;
                if (valid >= 4) {
                    float cvs = 0.f; int vc = 0;
                    for (int c = 0; c < C; ++c) {
                        const float mean = sm.car_sum[c] / valid;
                        if (mean < 0.001f) continue;
                        const float mean_sq = sm.car_sq[c] / valid;
                        const float var = std_max(0.0f, mean_sq - mean * mean);
                        cvs += sqrtf(var) / mean; ++vc;
                    }
                    temporal = (vc > 0) ? cvs / vc : 0.0f;
                }
                float sum = 0.f;
                for (int c = 0; c < C; ++c) sum += cm[c];
                const float mean = sum / C;
                float freq_cv = 0.0f;
                if (!(mean < 0.001f)) {
                    float vs = 0.f;
                    for (int c = 0; c < C; ++c) { const float d = cm[c] - mean; vs += d * d; }
                    freq_cv = sqrtf(vs / C) / mean;
                }
                a.fading[f] = freq_cv + 1.0f * temporal;
            }
            a.n_llr[f] = n_out;
            if (a.cfo_out) {
                const float cfo = a.cfo_hz ? a.cfo_hz[f] : 0.0f;
                // applyCFOCorrection zeroes cfo_hz_ once the correction is in the samples (:924)
                a.cfo_out[f] = (fabsf(cfo) > 0.1f && a.frame_len >= 128) ? 0.0f : cfo;
            }
        }
        __syncthreads();
        // ---- pass 2: LLRs (:683-698) ----
        for (int p = tid; p < n_ds * C; p += kDemodThreads) {
            const int d = p / C, c = p - d * C;
            const float phase = phases[d * kMaxCar + c];
            const float cs = sm.scale * sm.rel[c];
            if (a.bits == 2) {
                const float sb0 = cs * glibc_sinf(phase);
                const float sb1 = cs * glibc_sinf(2.0f * phase);
                llr_out[p * 2] = std_max(-20.0f, std_min(20.0f, sb0));
                llr_out[p * 2 + 1] = std_max(-20.0f, std_min(20.0f, sb1));
            } else {
                const float sb = cs * glibc_cosf(phase);
                llr_out[p] = std_max(-20.0f, std_min(20.0f, sb));
            }
        }
        for (int i = n_out + tid; i < a.llr_stride; i += kDemodThreads) llr_out[i] = 0.0f;
        __syncthreads();
    }
}

const char* mcdpsk_config_error(const ria_mcdpsk_config& c) {
    if (c.samples_per_symbol != kSps) return "samples_per_symbol must be 512";
    if (c.num_carriers < 1 || c.num_carriers > kMaxCar) return "num_carriers must be in [1, 16]";
    if (c.bits_per_symbol != 1 && c.bits_per_symbol != 2) return "bits_per_symbol must be 1 (DBPSK) or 2 (DQPSK)";
    if (c.spreading != 1 && c.spreading != 2 && c.spreading != 4) return "spreading must be 1, 2 or 4";
    if (c.training_symbols > 64) return "bad training_symbols";
    if (!(c.sample_rate > 0.0f)) return "sample_rate must be > 0";
    return nullptr;
}

void build_tables(const ria_mcdpsk_config& cfg, std::vector<float2>& mixer, std::vector<float>& taps) {
    // carrier frequencies (multi_carrier_dpsk.hpp:68-79) and mixer phasors (:931-946)
    const int C = static_cast<int>(cfg.num_carriers);
    mixer.resize(static_cast<size_t>(C) * kSps);
    for (int c = 0; c < C; ++c) {
        float freq;
        if (C == 1) freq = (cfg.freq_low + cfg.freq_high) / 2.0f;
        else { const float spacing = (cfg.freq_high - cfg.freq_low) / (C - 1); freq = cfg.freq_low + c * spacing; }
        const float phase_inc = 2.0f * M_PI * freq / cfg.sample_rate;
        float phase = 0.0f;
        for (int i = 0; i < kSps; ++i) {
            mixer[static_cast<size_t>(c) * kSps + i] = make_float2(1.0f * std::cos(-phase), 1.0f * std::sin(-phase));
            phase += phase_inc;
        }
    }
    // HilbertTransform(127) coefficients (filters.cpp:266-291)
    taps.assign(kHilbertTaps, 0.0f);
    const int M = (kHilbertTaps - 1) / 2;
    for (int n = 0; n < kHilbertTaps; ++n) {
        const int k = n - M;
        float v;
        if (k == 0) v = 0;
        else if (k % 2 != 0) v = 2.0f / (M_PI * k);
        else v = 0;
        const float w = 2.0f * M_PI * n / (kHilbertTaps - 1);
        v *= 0.42f - 0.5f * std::cos(w) + 0.08f * std::cos(2.0f * w);
        taps[n] = v;
    }
}

}  // namespace

void mcdpsk_tables_free(McdpskTablesDev* t) {
    if (!t) return;
    if (t->mixer) cudaFree(t->mixer);
    if (t->hilbert) cudaFree(t->hilbert);
    delete t;
}

static int mcdpsk_tables_dev(ria_ctx* ctx, const ria_mcdpsk_config& cfg, const McdpskTablesDev** out) {
    if (const char* e = mcdpsk_config_error(cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "mcdpsk: %s", e);
    for (McdpskTablesDev* t : ctx->mcdpsk_tables)
        if (std::memcmp(&t->cfg, &cfg, sizeof cfg) == 0) { *out = t; return RIA_OK; }
    std::vector<float2> mixer;
    std::vector<float> taps;
    build_tables(cfg, mixer, taps);
    McdpskTablesDev* t = new McdpskTablesDev();
    t->cfg = cfg;
    ctx->mcdpsk_tables.push_back(t);
    RIA_CUDA(ctx, cudaMalloc(&t->mixer, mixer.size() * sizeof(float2)));
    RIA_CUDA(ctx, cudaMalloc(&t->hilbert, taps.size() * sizeof(float)));
    RIA_CUDA(ctx, cudaMemcpy(t->mixer, mixer.data(), mixer.size() * sizeof(float2), cudaMemcpyHostToDevice));
    RIA_CUDA(ctx, cudaMemcpy(t->hilbert, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    RIA_CUDA(ctx, cudaMemcpyToSymbol(c_hilbert, taps.data(), taps.size() * sizeof(float)));
    t->ready = true;
    *out = t;
    return RIA_OK;
}

}  // namespace ria

extern "C" int ria_mcdpsk_soft_bits_per_frame(const ria_mcdpsk_config* cfg, int32_t frame_len) {
    if (!cfg || ria::mcdpsk_config_error(*cfg)) return RIA_E_INVAL;
    const int preamble = (static_cast<int>(cfg->training_symbols) + 1) * static_cast<int>(cfg->samples_per_symbol);
    if (frame_len <= preamble) return 0;
    const int n_rx = (frame_len - preamble) / static_cast<int>(cfg->samples_per_symbol);
    int n_ds = n_rx / static_cast<int>(cfg->spreading);
    if (n_ds < 1) n_ds = 1;
    return n_ds * static_cast<int>(cfg->num_carriers) * static_cast<int>(cfg->bits_per_symbol);
}

extern "C" int ria_mcdpsk_process_batch_at_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg,
                                               const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                               const int32_t* start_dev,
                                               const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                               float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                               float* fading_dev, float* cfo_out_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len < 0 || frame_stride < frame_len) return set_error(ctx, RIA_E_INVAL, "mcdpsk: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !llr_dev || !n_llr_dev) return set_error(ctx, RIA_E_INVAL, "mcdpsk: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const McdpskTablesDev* t = nullptr;
    int rc = mcdpsk_tables_dev(ctx, *cfg, &t);
    if (rc != RIA_OK) return rc;
    const int n_out = ria_mcdpsk_soft_bits_per_frame(cfg, frame_len);
    if (llr_stride < n_out) return set_error(ctx, RIA_E_INVAL, "mcdpsk: llr_stride %d < %d soft bits per frame", llr_stride, n_out);
    const int C = static_cast<int>(cfg->num_carriers);
    const int max_ds = n_out / (C * static_cast<int>(cfg->bits_per_symbol)) + 1;
    const int n_blocks = (frame_len + 31) / 32;

    // Frames go through in chunks so that the scratch (per-frame mags/phases/errs and, with a CFO
    // vector, block phases + corrected samples) stays bounded for 10^5-frame batches.
    const size_t corr_stride = (static_cast<size_t>(frame_len) + 3) & ~size_t(3);
    const size_t per_frame = static_cast<size_t>(3) * max_ds * kMaxCar * sizeof(float) +
                             (cfo_hz_dev ? (2 * static_cast<size_t>(n_blocks) + corr_stride) * sizeof(float) : 0);
    int64_t chunk = static_cast<int64_t>((size_t(3) << 30) / (per_frame ? per_frame : 1));     // <= 3 GiB of scratch
    if (chunk > 65535) chunk = 65535;
    if (chunk > n_frames) chunk = n_frames;
    if (chunk < 1) chunk = 1;
    const size_t s_demod = static_cast<size_t>(chunk) * 3 * max_ds * kMaxCar * sizeof(float);
    const size_t s_phase = cfo_hz_dev ? static_cast<size_t>(chunk) * n_blocks * sizeof(float2) : 0;
    const size_t s_corr = cfo_hz_dev ? static_cast<size_t>(chunk) * corr_stride * sizeof(float) : 0;
    const size_t a1 = (s_demod + 255) & ~size_t(255), a2 = (s_phase + 255) & ~size_t(255);
    rc = ensure_scratch(ctx, a1 + a2 + s_corr + 256);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    float* d_scr = reinterpret_cast<float*>(base);
    float2* d_bph = reinterpret_cast<float2*>(base + a1);
    float* d_corr = reinterpret_cast<float*>(base + a1 + a2);

    size_t smem = demod_smem_bytes(C);
    int stats_smem_off = 0;
    {
        // per-frame mags / phases / errs in shared memory when two CTAs still fit on an SM
        const size_t off = (smem + 15) & ~size_t(15);
        const size_t with_stats = off + static_cast<size_t>(3) * max_ds * kMaxCar * sizeof(float);
        if (2 * (with_stats + 1024) <= ctx->smem_per_sm && with_stats <= ctx->smem_optin) {
            stats_smem_off = static_cast<int>(off);
            smem = with_stats;
        }
    }
    RIA_CUDA(ctx, cudaFuncSetAttribute(mcdpsk_demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    RIA_CUDA(ctx, cudaFuncSetAttribute(mcdpsk_demod_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    RIA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mcdpsk_demod_kernel, kDemodThreads, smem));
    if (per_sm < 1) return set_error(ctx, RIA_E_UNSUPPORTED, "mcdpsk: kernel does not fit");

    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = (n_frames - off < chunk) ? (n_frames - off) : chunk;
        DemodArgs a{};
        a.samples = samples_dev; a.frame_stride = frame_stride; a.frame_len = frame_len;
        a.start = start_dev; a.row_len = frame_stride; a.frame0 = off; a.corrected = nullptr; a.corrected_stride = 0;
        if (cfo_hz_dev) {
            const unsigned scan_blocks = static_cast<unsigned>((n * 32 + 127) / 128);
            time_begin(ctx, KK_MCDPSK_CFO);
            mcdpsk_phase_scan_kernel<<<scan_blocks, 128, 0, ctx->stream>>>(cfo_hz_dev, phase_dev, off, n, n_blocks,
                                                                          cfg->sample_rate, d_bph);
            dim3 grid(static_cast<unsigned>((frame_len + 1023) / 1024), static_cast<unsigned>(n));
            mcdpsk_cfo_kernel<<<grid, 256, 0, ctx->stream>>>(samples_dev, frame_stride, frame_len, start_dev, off,
                                                             cfo_hz_dev, d_bph, n_blocks, cfg->sample_rate,
                                                             d_corr, static_cast<long long>(corr_stride));
            time_end(ctx);
            RIA_CUDA(ctx, cudaGetLastError());
            ctx->launches += 2;
            a.corrected = d_corr; a.corrected_stride = static_cast<long long>(corr_stride);
        }
        a.cfo_hz = cfo_hz_dev; a.n_frames = n;
        a.llr = llr_dev; a.llr_stride = llr_stride; a.n_llr = n_llr_dev; a.fading = fading_dev; a.cfo_out = cfo_out_dev;
        a.scratch = d_scr; a.max_ds = max_ds; a.mixer_g = t->mixer; a.stats_smem_off = stats_smem_off;
        a.carriers = C; a.bits = static_cast<int>(cfg->bits_per_symbol); a.spread = static_cast<int>(cfg->spreading);
        a.training = static_cast<int>(cfg->training_symbols);
        a.counter = ctx->work_counter + 2;
        long long grid = static_cast<long long>(ctx->sm_count) * per_sm;
        if (grid > n) grid = n;
        RIA_CUDA(ctx, cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), ctx->stream));
        time_begin(ctx, KK_MCDPSK);
        mcdpsk_demod_kernel<<<static_cast<unsigned>(grid), kDemodThreads, smem, ctx->stream>>>(a);
        time_end(ctx);
        RIA_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return RIA_OK;
}

extern "C" int ria_mcdpsk_process_batch_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg,
                                            const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                            const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                            float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                            float* fading_dev, float* cfo_out_dev) {
    return ria_mcdpsk_process_batch_at_dev(ctx, cfg, samples_dev, frame_stride, frame_len, nullptr, cfo_hz_dev, phase_dev,
                                           n_frames, llr_dev, llr_stride, n_llr_dev, fading_dev, cfo_out_dev);
}
