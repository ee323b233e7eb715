// Batched OFDM light (training-only) preamble detection for connected-mode DATA frames.
//
// Replaces OFDMChirpWaveform::detectDataSync (src/waveform/ofdm_chirp_waveform.cpp:207-384):
// energy gate on the first 100 ms, 65-tap Blackman Hilbert FIR (src/dsp/filters.cpp:266-317),
// Schmidl-Cox style lag-one-symbol autocorrelation of the analytic signal on a step-8 grid with
// early exit at the first correlation above 0.95, +-4 sample refinement, burst-interleave marker
// from the sign of Re(P e^{-j phi_cfo}).
//
// The search is sequential in the reference only through its early exit; here every candidate
// offset is evaluated in parallel (each with the reference's in-order fp32 sums) and the scan
// semantics (running best, first > 0.95 stops) are applied afterwards by one thread, which gives
// the identical offset.

#include "ria_internal.h"
#include "rn_math.h"

#include <cmath>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {
namespace {

constexpr int kTaps = 65;
constexpr int kDelay = 32;
constexpr int kThreads = 256;          // two CTAs per SM (tile + energy plane ~84 KB each)
constexpr int kMaxCand = 1280;         // >= the widest coarse grid: 8 symbols / 8, symbols <= 1160 samples

__device__ __forceinline__ float cabs_d(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}

struct SyncArgs {
    const float* samples; long long frame_stride; int window;
    const float* known_cfo; float threshold; int sym; float sample_rate;
    const float* taps_g;
    int an_cap;                                 // analytic samples the shared-memory tile holds
    int an_words;                               // float2 slots of the tile; the energy plane (floats, same padding) follows
    ria_sync_result* out;
};

// The analytic signal of the span the search touches lives in shared memory, sample i of the span at
// an_pad(i) = i + (i >> 5).  In the coarse search a thread owns four neighbouring candidates (32 samples), so the threads
// of a warp read float2 words 33 apart (and floats 33 apart in the energy plane): an odd stride, conflict-free.
__device__ __forceinline__ int an_pad(int i) { return i + (i >> 5); }

// Blackwell packed fp32; products only ever feed scalar adds (ptxas would contract a packed multiply into a packed add)
__device__ __forceinline__ float2 ds_mul2s(float s, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 ds_mul2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 ds_add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

// P = sum conj(s1) s2, energies, in order (:283-295); `offset` is relative to the tile
__device__ __forceinline__ void lag_corr(const float2* an, int offset, int sym, float* corr, float2* P_out) {
    float2 P = make_float2(0.f, 0.f);
    float e1 = 0.f, e2 = 0.f;
    for (int n = 0; n < sym; ++n) {
        const float2 s1 = an[an_pad(offset + n)], s2 = an[an_pad(offset + n + sym)];
        // conj(s1) * s2
        const float re = __fsub_rn(__fmul_rn(s1.x, s2.x), __fmul_rn(-s1.y, s2.y));
        const float im = __fadd_rn(__fmul_rn(s1.x, s2.y), __fmul_rn(-s1.y, s2.x));
        P.x = __fadd_rn(P.x, re); P.y = __fadd_rn(P.y, im);
        e1 = __fadd_rn(e1, __fadd_rn(__fmul_rn(s1.x, s1.x), __fmul_rn(s1.y, s1.y)));
        e2 = __fadd_rn(e2, __fadd_rn(__fmul_rn(s2.x, s2.x), __fmul_rn(s2.y, s2.y)));
    }
    const float denom = __fadd_rn(sqrtf(__fmul_rn(e1, e2)), 1e-10f);
    *corr = __fdiv_rn(cabs_d(P), denom);
    *P_out = P;
}

// The coarse grid: candidate c sits at tile offset 8 c and sums, in sample order (:283-295),
//   P(c) = sum_n conj(a[8c+n]) a[8c+n+sym],   e1(c) = sum_n |a[8c+n]|^2,   e2(c) = sum_n |a[8c+n+sym]|^2 = e1(c + sym/8).
// The TERMS depend on the sample index k = 8c + n alone, so they are formed once per sample -- T[k] = conj(a[k]) a[k+sym]
// (in place of a[k]) and E[k] = |a[k]|^2 -- with the reference's products (conj(s1) s2 = (s1x s2x + s1y s2y,
// s1x s2y - s1y s2x); a - (-t) = a + t exactly), and a candidate only ADDS its 'sym' terms in order: the same floats
// added in the same order as the reference's loop.  A thread owns four neighbouring candidates and slides a window of
// four 8-sample blocks over the term planes, loading each block once.
constexpr int kCandPerThread = 4;

struct Blk { float2 t[8]; float e[8]; };

__device__ __forceinline__ void load_blk(const float2* T, const float* E, int b, Blk& o) {
    const int base = an_pad(8 * b);            // 8 b is a multiple of 8: the block sits inside one padded group of 32
#pragma unroll
    for (int j = 0; j < 8; ++j) { o.t[j] = T[base + j]; o.e[j] = E[base + j]; }
}
__device__ __forceinline__ void add_blk(const Blk& b, float2& P, float& e) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { P = ds_add2(P, b.t[j]); e = __fadd_rn(e, b.e[j]); }
}

// candidates c0 .. c0+3: (P, e1) each
__device__ __forceinline__ void coarse_quad(const float2* T, const float* E, int c0, int n_blocks, float2 (&P)[4], float (&e)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { P[q] = make_float2(0.f, 0.f); e[q] = 0.f; }
    Blk w0, w1, w2, w3;
    load_blk(T, E, c0, w0); load_blk(T, E, c0 + 1, w1); load_blk(T, E, c0 + 2, w2);
    int b = 0;
    // step b adds block (c0 + q + b) to candidate q; the window rotates through four names
    for (; b + 4 <= n_blocks; b += 4) {
        load_blk(T, E, c0 + 3 + b, w3); add_blk(w0, P[0], e[0]); add_blk(w1, P[1], e[1]); add_blk(w2, P[2], e[2]); add_blk(w3, P[3], e[3]);
        load_blk(T, E, c0 + 4 + b, w0); add_blk(w1, P[0], e[0]); add_blk(w2, P[1], e[1]); add_blk(w3, P[2], e[2]); add_blk(w0, P[3], e[3]);
        load_blk(T, E, c0 + 5 + b, w1); add_blk(w2, P[0], e[0]); add_blk(w3, P[1], e[1]); add_blk(w0, P[2], e[2]); add_blk(w1, P[3], e[3]);
        load_blk(T, E, c0 + 6 + b, w2); add_blk(w3, P[0], e[0]); add_blk(w0, P[1], e[1]); add_blk(w1, P[2], e[2]); add_blk(w2, P[3], e[3]);
    }
    for (; b < n_blocks; ++b) {                // n_blocks % 4 leftover steps
        load_blk(T, E, c0 + 3 + b, w3);
        add_blk(w0, P[0], e[0]); add_blk(w1, P[1], e[1]); add_blk(w2, P[2], e[2]); add_blk(w3, P[3], e[3]);
        w0 = w1; w1 = w2; w2 = w3;
    }
}

// Four Hilbert outputs of one parity (see build_analytic): P = 0 even outputs, 1 odd outputs relative to the staged
// origin.  w[] holds src[4 g - 32 .. 4 g + 3]; tap t of output u reads w[31 + u - t + P].
template <int P>
__device__ __forceinline__ void hilbert4(const float* __restrict__ src, const float* taps, int g, float (&q)[4]) {
    float w[36];
#pragma unroll
    for (int v = 0; v < 9; ++v) {
        const float4 q4 = *reinterpret_cast<const float4*>(src + 4 * g - 32 + 4 * v);
        w[4 * v] = q4.x; w[4 * v + 1] = q4.y; w[4 * v + 2] = q4.z; w[4 * v + 3] = q4.w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = 0.f;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        const float tp = taps[2 * t + 1];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = __fadd_rn(q[u], __fmul_rn(tp, w[31 + u - t + P]));
    }
}

__global__ void __launch_bounds__(kThreads, 2)
ofdm_data_sync_kernel(const SyncArgs a) {
    extern __shared__ __align__(16) float2 an_tile[];          // analytic samples, then product terms, of one pass (an_pad)
    __shared__ float taps[kTaps + 3];
    __shared__ int s_i[kThreads];
    __shared__ float s_f[4];
    __shared__ int s_n[4];
    __shared__ float cand[3 * kMaxCand];                         // corr | P.re | P.im per candidate
    __shared__ float e_sum[kMaxCand + 520];                      // first energy sum per grid position
    const long long f = blockIdx.x;
    const int tid = threadIdx.x;
    const float* x = a.samples + f * a.frame_stride;
    const int N = a.window, sym = a.sym;
    const float known = a.known_cfo ? a.known_cfo[f] : 0.0f;

    ria_sync_result res;
    res.detected = 0; res.start_sample = -1; res.correlation = 0.0f; res.cfo_hz = known;
    res.snr_estimate = 0.0f; res.root = 0; res.frame_type = 0; res.aux = 0;
    if (N < sym * 3) { if (tid == 0) a.out[f] = res; return; }      // :222-224

    if (tid < kTaps) taps[tid] = a.taps_g[tid];
    float* E = reinterpret_cast<float*>(an_tile + a.an_words);       // energy plane; before that, staged samples
    // The SQUARES of the window's samples are staged once in the tile (free until the first pass builds its analytic
    // samples; the product is exact per sample): the energy gate and the signal-start search only ever add squares.
    // Windows longer than the tile fall back to global memory.
    float* sq = reinterpret_cast<float*>(an_tile);
    const bool staged = N + 8 <= 2 * a.an_words;
    if (staged) for (int i = tid; i < N; i += kThreads) { const float v = x[i]; sq[i] = __fmul_rn(v, v); }
    __syncthreads();
    // ---- energy gate (:232-260): a sequential sum of up to 4800 squares ----
    const int ns = min(N / 4, 4800);
    if (tid == 0) {
        float nf = 0.0f;
        int i = 0;
        if (staged) {
            for (; i + 4 <= ns; i += 4) {
                const float4 q = *reinterpret_cast<const float4*>(sq + i);
                nf = __fadd_rn(nf, q.x); nf = __fadd_rn(nf, q.y); nf = __fadd_rn(nf, q.z); nf = __fadd_rn(nf, q.w);
            }
            for (; i < ns; ++i) nf = __fadd_rn(nf, sq[i]);
        } else {
            for (; i < ns; ++i) nf = __fadd_rn(nf, __fmul_rn(x[i], x[i]));
        }
        nf = sqrtf(nf / ns);
        s_f[0] = nf;
        s_f[1] = nf * 3.0f + 0.01f;
    }
    __syncthreads();
    const bool in_noise = s_f[0] < 0.05f;
    int signal_start = 0;
    if (in_noise) {
        const float thr = s_f[1];
        int first = 0x7fffffff;
        const int lim = N - sym * 2;
        for (int i = tid; i < lim && i < first; i += kThreads) {
            float e = 0.0f;
            if (staged && i + 64 <= N) {
#pragma unroll 16
                for (int j = 0; j < 64; ++j) e = __fadd_rn(e, sq[i + j]);
            } else {
                for (int j = 0; j < 64; ++j) if (i + j < N) e = __fadd_rn(e, __fmul_rn(x[i + j], x[i + j]));
            }
            e = sqrtf(e / 64);
            if (e > thr) { first = i; break; }
        }
        s_i[tid] = first;
        __syncthreads();
        for (int s = kThreads / 2; s > 0; s >>= 1) { if (tid < s) s_i[tid] = min(s_i[tid], s_i[tid + s]); __syncthreads(); }
        signal_start = (s_i[0] == 0x7fffffff) ? 0 : s_i[0];
        __syncthreads();
    }
    const int search_window = sym * 4;
    const int actual = in_noise ? search_window : max(search_window, sym * 8);
    const int search_end = min(signal_start + actual, N - sym * 2);

    // Analytic samples [alo, hi) of the window -> the tile (tile index 0 = alo) (:266-270, filters.cpp:293-317).
    // Only the odd taps of the 65-tap Hilbert FIR are non-zero, so output i reads the samples of the OTHER parity:
    //   q(i) = sum_{t=0..31} taps[2t+1] x[i - 2t - 1]   (accumulated in this order),   re(i) = x[i - 32].
    // The span is staged de-interleaved in the energy plane (even samples | odd samples, zeros before sample 0), which
    // makes the 32 samples of an output a contiguous run and the runs of outputs i, i+2, i+4, i+6 overlap: a thread
    // produces those four outputs from nine 128-bit loads.  Threads 0..127 take the even outputs, 128..255 the odd ones.
    auto build_analytic = [&](int alo, int hi) {
        const int j0 = (alo - 64) & ~7;                               // staged sample 0 (may be negative: zeros)
        const int n_st = hi - j0;                                     // staged samples
        const int half = ((n_st + 1) / 2 + 7) & ~3;                   // floats per parity plane, 16-byte granules
        float* Ee = E;
        float* Eo = E + half;
        __syncthreads();
        for (int m = tid; m < half; m += kThreads) {
            const int je = j0 + 2 * m, jo = je + 1;
            Ee[m] = (je >= 0 && je < hi) ? x[je] : 0.0f;
            Eo[m] = (jo >= 0 && jo < hi) ? x[jo] : 0.0f;
        }
        __syncthreads();
        const int p = tid >> 7;                                       // output parity relative to j0 (j0 is even)
        const float* src = p ? Ee : Eo;                               // taps read the other parity
        const float* same = p ? Eo : Ee;                              // x[i - 32] has the parity of i
        const int n_groups = (n_st + 7) / 8;
        for (int g = 8 + (tid & 127); g < n_groups; g += 128) {       // local output index li = 8 g + 2 u + p >= 64
            float q[4];
            if (p) hilbert4<1>(src, taps, g, q); else hilbert4<0>(src, taps, g, q);
            const float4 re4 = *reinterpret_cast<const float4*>(same + 4 * g - 16);     // x[i - 32], u = 0..3
            const float re[4] = {re4.x, re4.y, re4.z, re4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = j0 + 8 * g + 2 * u + p;
                if (i >= alo && i < hi) an_tile[an_pad(i - alo)] = make_float2(i >= kDelay ? re[u] : 0.0f, q[u]);
            }
        }
        __syncthreads();
    };
    float2* an = an_tile;

    // ---- coarse candidates (:283-312), in passes of at most kPassCand candidates (four symbols of offsets): a pass
    // stages its span, forms its term planes and sums its candidates; the scan semantics (running best, first > 0.95
    // stops) are applied after each pass, so a hit in the first pass ends the search like the reference's break ----
    int n_cand = (search_end > signal_start) ? (search_end - signal_start + 7) / 8 : 0;
    if (n_cand > kMaxCand) n_cand = kMaxCand;
    const bool share = (sym % 8) == 0 && sym / 8 <= 520;            // second energy = first energy sym / 8 places on
    const int hop = sym / 8, n_blocks = sym / 8;
    // a pass of nc candidates needs 8 (nc + hop - 1) + sym terms (term planes) or 8 (nc - 1) + 2 sym analytic samples
    const int pass_cand = share ? (max(8, (a.an_cap - sym) / 8 + 1 - hop - 8) & ~3) : max(8, (a.an_cap - 2 * sym) / 8);
    __syncthreads();                                                // every thread has read the gate's s_f[0] / s_f[1]
    if (tid == 0) { s_f[2] = 0.0f; s_n[0] = 0; s_f[0] = 0.f; s_f[1] = 0.f; s_n[1] = 0; }
    __syncthreads();
    for (int c_lo = 0; c_lo < n_cand; c_lo += pass_cand) {
        const int nc = min(pass_cand, n_cand - c_lo);                // candidates of this pass
        const int alo = signal_start + 8 * c_lo;                     // tile origin
        if (share) {
            const int want = 8 * (nc + hop - 1) + sym;               // energy terms the pass sums
            const int hi = min(N, alo + want);
            build_analytic(alo, hi);
            const int have = hi - alo;
            const int n_e = have;
            const int n_t = min(max(have - sym, 0), 8 * (nc - 1) + sym);
            // in place, ascending chunks of <= sym samples: a chunk reads a[k], a[k + sym], then (barrier) writes T[k] over a[k]
            const int chunk = (sym < 4 * kThreads) ? ((sym < 2 * kThreads) ? kThreads : 2 * kThreads) : 4 * kThreads;
            for (int k0 = 0; k0 < n_e; k0 += chunk) {
                float2 tv[4]; float ev[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + tid + u * kThreads;
                    tv[u] = make_float2(0.f, 0.f); ev[u] = 0.f;
                    if (u * kThreads < chunk && k < n_e) {
                        const float2 s1 = an[an_pad(k)];
                        const float2 sq = ds_mul2(s1, s1);
                        ev[u] = __fadd_rn(sq.x, sq.y);
                        if (k < n_t) {
                            const float2 s2 = an[an_pad(k + sym)];
                            const float2 p1 = ds_mul2s(s1.x, s2);                              // (s1x s2x, s1x s2y)
                            const float2 p2 = ds_mul2s(s1.y, make_float2(s2.y, s2.x));         // (s1y s2y, s1y s2x)
                            tv[u] = make_float2(__fadd_rn(p1.x, p2.x), __fsub_rn(p1.y, p2.y));
                        }
                    }
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + tid + u * kThreads;
                    if (u * kThreads < chunk && k < n_e) { an[an_pad(k)] = tv[u]; E[an_pad(k)] = ev[u]; }
                }
                __syncthreads();
            }
            // zero what the sliding windows may touch past the planes (a block is loaded whole)
            for (int k = n_e + tid; k < a.an_cap; k += kThreads) { an[an_pad(k)] = make_float2(0.f, 0.f); E[an_pad(k)] = 0.f; }
            __syncthreads();
            const int n_quads = (nc + kCandPerThread - 1) / kCandPerThread;
            if (tid < n_quads) {
                float2 P[4]; float e[4];
                coarse_quad(an, E, kCandPerThread * tid, n_blocks, P, e);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = kCandPerThread * tid + q;
                    if (c < nc) { e_sum[c_lo + c] = e[q]; cand[kMaxCand + c_lo + c] = P[q].x; cand[2 * kMaxCand + c_lo + c] = P[q].y; }
                }
            } else {
                // the other warps: first energy sums of the grid positions behind the pass's last candidate (the second
                // energies of its last sym / 8 candidates)
                const int first_tail_thread = (n_quads + 31) & ~31;
                for (int c = nc + (tid - first_tail_thread); tid >= first_tail_thread && c < nc + hop; c += kThreads - first_tail_thread) {
                    float e1 = 0.f;
                    for (int b = 0; b < n_blocks; ++b) {
                        const float* eb = E + an_pad(8 * (c + b));
#pragma unroll
                        for (int j = 0; j < 8; ++j) e1 = __fadd_rn(e1, eb[j]);
                    }
                    e_sum[c_lo + c] = e1;
                }
            }
            __syncthreads();
            for (int c = c_lo + tid; c < c_lo + nc; c += kThreads) {
                const float denom = __fadd_rn(sqrtf(__fmul_rn(e_sum[c], e_sum[c + hop])), 1e-10f);
                cand[c] = __fdiv_rn(cabs_d(make_float2(cand[kMaxCand + c], cand[2 * kMaxCand + c])), denom);
            }
        } else {
            build_analytic(alo, min(N, alo + 8 * (nc - 1) + 2 * sym));
            for (int c = tid; c < nc; c += kThreads) {
                float corr; float2 P;
                lag_corr(an, 8 * c, sym, &corr, &P);
                cand[c_lo + c] = corr; cand[kMaxCand + c_lo + c] = P.x; cand[2 * kMaxCand + c_lo + c] = P.y;
            }
        }
        __syncthreads();
        if (tid == 0) {
            float best = s_f[2]; int bo = s_n[0]; float2 bp = make_float2(s_f[0], s_f[1]);
            int stop = 0;
            for (int c = c_lo; c < c_lo + nc; ++c) {
                const float corr = cand[c];
                if (corr > best) { best = corr; bo = signal_start + 8 * c; bp = make_float2(cand[kMaxCand + c], cand[2 * kMaxCand + c]); }
                if (corr > 0.95f) { stop = 1; break; }                // first high-confidence peak
            }
            s_f[2] = best; s_n[0] = bo; s_f[0] = bp.x; s_f[1] = bp.y; s_n[1] = stop;
        }
        __syncthreads();
        if (s_n[1]) break;
    }
    float best_corr = s_f[2]; int best_offset = s_n[0]; float2 best_p = make_float2(s_f[0], s_f[1]);
    __syncthreads();
    // ---- +-4 refinement (:318-350) ----
    if (best_corr > a.threshold) {
        const int r0 = max(signal_start, best_offset - 4), r1 = min(search_end, best_offset + 5);
        // The fine lags are one sample apart: their terms T[k] = conj(a[k]) a[k + sym], E[k] = |a[k]|^2 over
        // k in [r0, r1 + 2 sym) are formed once, then one thread per (lag, sum) walks its 'sym' terms in order
        // (lag_corr's three sums, :283-295), the three kinds of sum on three warps.  Layout inside the tile:
        // analytic samples | T plane | E plane.
        const int n_fine = r1 - r0;
        const int fine_hi = min(N, r1 + 2 * sym);
        const int n_an = fine_hi - r0;                               // analytic samples of the fine lags
        build_analytic(r0, fine_hi);
        float2* FT = an + an_pad(n_an) + 8;                          // n_fine - 1 + sym terms
        float* FE = reinterpret_cast<float*>(FT + (n_fine + sym + 8)); // n_fine - 1 + 2 sym terms
        for (int k = tid; k < n_an; k += kThreads) {
            const float2 s1 = an[an_pad(k)];
            FE[k] = __fadd_rn(__fmul_rn(s1.x, s1.x), __fmul_rn(s1.y, s1.y));
            if (k + sym < n_an) {
                const float2 s2 = an[an_pad(k + sym)];
                FT[k] = make_float2(__fsub_rn(__fmul_rn(s1.x, s2.x), __fmul_rn(-s1.y, s2.y)),
                                    __fadd_rn(__fmul_rn(s1.x, s2.y), __fmul_rn(-s1.y, s2.x)));
            }
        }
        __syncthreads();
        {
            const int kind = tid >> 5, d = tid & 31;                  // warp 0: P, warp 1: e1, warp 2: e2
            if (kind < 3 && d < n_fine && r0 + d != best_offset) {
                if (kind == 0) {
                    float2 P = make_float2(0.f, 0.f);
#pragma unroll 4
                    for (int n = 0; n < sym; ++n) { const float2 t = FT[d + n]; P.x = __fadd_rn(P.x, t.x); P.y = __fadd_rn(P.y, t.y); }
                    cand[kMaxCand + d] = P.x; cand[2 * kMaxCand + d] = P.y;
                } else {
                    const float* e = FE + d + (kind == 2 ? sym : 0);
                    float acc = 0.f;
#pragma unroll 4
                    for (int n = 0; n < sym; ++n) acc = __fadd_rn(acc, e[n]);
                    e_sum[2 * d + (kind - 1)] = acc;
                }
            }
        }
        __syncthreads();
        if (tid < n_fine && r0 + tid != best_offset) {
            const float denom = __fadd_rn(sqrtf(__fmul_rn(e_sum[2 * tid], e_sum[2 * tid + 1])), 1e-10f);
            cand[tid] = __fdiv_rn(cabs_d(make_float2(cand[kMaxCand + tid], cand[2 * kMaxCand + tid])), denom);
        }
        __syncthreads();
        if (tid == 0) {
            for (int o = r0; o < r1; ++o) {
                if (o == best_offset) continue;
                const float corr = cand[o - r0];
                if (corr > best_corr) { best_corr = corr; best_offset = o; best_p = make_float2(cand[kMaxCand + o - r0], cand[2 * kMaxCand + o - r0]); }
            }
        }
    }
    if (tid == 0) {
        res.correlation = best_corr;
        if (best_corr > a.threshold) {
            res.detected = 1;
            res.start_sample = best_offset;
            // burst-interleave marker (:366-372)
            const float cfo_phase = static_cast<float>(2.0f * M_PI * static_cast<double>(known) * sym / static_cast<double>(a.sample_rate));
            float s, c;
            glibc_sincosf(-cfo_phase, &s, &c);                  // std::cos / std::sin of a float = glibc, restated
            const float re = __fsub_rn(__fmul_rn(best_p.x, c), __fmul_rn(best_p.y, s));
            res.aux = (re < 0.0f) ? 1 : 0;
        }
        a.out[f] = res;
    }
}

}  // namespace
}  // namespace ria

extern "C" int ria_ofdm_data_sync_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                            const float* samples_dev, int64_t frame_stride, int32_t window,
                                            const float* known_cfo_dev, float threshold, int64_t n_frames,
                                            ria_sync_result* out_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || window < 0 || frame_stride < window) return set_error(ctx, RIA_E_INVAL, "data sync: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "data sync: null buffer");
    const int sym = ria_ofdm_symbol_samples(cfg);
    if (sym <= 0 || sym > 4096) return set_error(ctx, RIA_E_UNSUPPORTED, "data sync: unsupported symbol length");
    if ((sym * 8 + 7) / 8 > kMaxCand) return set_error(ctx, RIA_E_UNSUPPORTED, "data sync: search window too large");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    // HilbertTransform(65) coefficients (filters.cpp:266-291), host floats
    if (!ctx->hilbert65) {
        float taps[kTaps];
        const int M = (kTaps - 1) / 2;
        for (int n = 0; n < kTaps; ++n) {
            const int k = n - M;
            float v;
            if (k == 0) v = 0;
            else if (k % 2 != 0) v = 2.0f / (M_PI * k);
            else v = 0;
            const float w = 2.0f * M_PI * n / (kTaps - 1);
            v *= 0.42f - 0.5f * std::cos(w) + 0.08f * std::cos(2.0f * w);
            taps[n] = v;
        }
        RIA_CUDA(ctx, cudaMalloc(&ctx->hilbert65, sizeof taps));
        RIA_CUDA(ctx, cudaMemcpy(ctx->hilbert65, taps, sizeof taps, cudaMemcpyHostToDevice));
    }
    // shared-memory tile of one pass: four symbols of candidate offsets plus the two symbols the last candidate
    // correlates over (the widest search, eight symbols when the buffer starts inside a burst, :273-276, takes two
    // passes), padded one word per 32
    const int an_cap = 6 * sym + 64;
    const int an_words = an_cap + an_cap / 32 + 8;
    const size_t smem = static_cast<size_t>(an_words) * (sizeof(float2) + sizeof(float));
    if (smem + 48 * 1024 > ctx->smem_optin) return set_error(ctx, RIA_E_UNSUPPORTED, "data sync: symbol too long for the shared-memory tile");
    RIA_CUDA(ctx, cudaFuncSetAttribute(ofdm_data_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SyncArgs a{};
    a.samples = samples_dev; a.frame_stride = frame_stride; a.window = window;
    a.known_cfo = known_cfo_dev; a.threshold = threshold; a.sym = sym; a.sample_rate = static_cast<float>(cfg->sample_rate);
    a.taps_g = ctx->hilbert65;
    a.an_cap = an_cap;
    a.an_words = an_words;
    a.out = out_dev;
    time_begin(ctx, KK_OFDM_SYNC);
    ofdm_data_sync_kernel<<<static_cast<unsigned>(n_frames), kThreads, smem, ctx->stream>>>(a);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
