// Batched dual-chirp (up/down LFM) preamble detection with CFO estimation for sm_100a.
//
// Replaces sync::ChirpSync::detectDualChirp / detectChirpTemplateFFT (src/sync/chirp_sync.hpp:
// 352-512, 627-712) as called by OFDMChirpWaveform::detectSync / MCDPSKWaveform::detectSync
// (src/waveform/ofdm_chirp_waveform.cpp:163-205, mc_dpsk_waveform.cpp:177-224): matched filter
// against the complex (cos + j sin) up- and down-chirp templates through a 131072-point FFT,
// normalisation by the sliding signal energy from an fp32 prefix sum, first strict maximum, down
// chirp searched in a window that starts half a chirp after the up-chirp peak, CFO from the
// error of the gap between the two peaks, CFO-corrected positions.
//
// B200 mapping: the 2^17-point transform is three shared-memory stages 32 x 64 x 64 (four-step
// decomposition, in place, digit-permuted spectrum); because forward and inverse use mirrored
// stage orders no transpose or bit-reversal pass ever touches HBM, and the template spectra are
// stored in the same permuted order.  One forward transform of the window serves both templates
// (the reference transforms the down-chirp search slice separately; the correlation at a given
// lag is the same sum, only fp32 rounding differs at the 1e-6 level, far below the peak margins).
// The sliding energy uses the reference's sequential fp32 prefix sums, restarted at the start of
// the down-chirp search slice exactly like the reference.
//
// When the down-chirp search slice is shorter than two chirps (window truncated right after the
// up chirp) the reference falls back to a time-domain coarse/fine/parabolic search
// (detectChirpTemplate :745-817); td_detect reproduces it with the reference's sequential sums.

#include "ria_internal.h"

#include <cmath>
#include <cstdlib>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

struct ChirpTablesDev {
    ria_chirp_config cfg{};
    float2* tw1 = nullptr;        // [A][B*C]  W_N^{ka * m}
    float2* tw2 = nullptr;        // [B][C]    W_{BC}^{kb * c}
    float2* tmpl_up = nullptr;    // conj(FFT(up template)), permuted layout
    float2* tmpl_dn = nullptr;
    float2* tmpl_dn_time = nullptr;   // (cos, sin) of the down-chirp template, time domain
    float energy_up = 0.f, energy_dn = 0.f;
    int chirp_len = 0, gap = 0;
    // residue-decomposed transform (default path): twiddle tables, template spectra in [residue][digit-reversed]
    // order, and the per-CTA spill row of the second template product
    float2* res_tables = nullptr;     // wa_hi | wa_lo | wb | wc | wr_hi | wr_lo
    float2* tmpl_up_res = nullptr;    // [8][16384]
    float2* tmpl_dn_res = nullptr;
    float2* res_scratch = nullptr;    // [res_grid][16384]
    int res_grid = 0;
};

namespace {

constexpr int kA = 32, kB = 64, kC = 64;
constexpr int kN = kA * kB * kC;          // 131072 = ChirpSync::FFT_SIZE (:565)
constexpr int kFftThreads = 128;
constexpr int kTilesPerCta = 4;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---------------------------------------------------------------------------------------------
// One stage of the three-stage transform: a tile of small DFTs (length 32 or 64) evaluated as
// radix-8 x radix-4/8 in registers with one exchange through shared memory, followed by the
// pointwise twiddle of the four-step decomposition (or the final 1/N scale).
//   mode 1: over a (L = 32), element stride B*C      forward: then * tw1     inverse: then * scale
//   mode 2: over b (L = 64), element stride C        forward: then * tw2     inverse: then * conj(tw1)
//   mode 3: over c (L = 64), contiguous              forward: nothing        inverse: then * conj(tw2)
// Forward mode 1 can read the real input window directly (zero-padded, :647-650) and forward
// mode 3 can write the two template products (:656-659) instead of the spectrum, so neither the
// packed complex signal nor the bare spectrum ever exists in HBM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -j (forward) / +j (inverse)
template <bool INV> __device__ __forceinline__ float2 mul_mj(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// 8-point DFT, natural order in and out (decimation in frequency, outputs un-bit-reversed)
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
    const float h = 0.70710678118654752440f;
    float2 b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = cadd2(a[i], a[i + 4]); b[i + 4] = csub2(a[i], a[i + 4]); }
    // twiddles W8^i on the lower half: 1, (1 -+ j)/sqrt2, -+j, (-1 -+ j)/sqrt2
    {
        const float2 t5 = b[5], t7 = b[7];
        b[5] = INV ? make_float2(h * (t5.x - t5.y), h * (t5.x + t5.y)) : make_float2(h * (t5.x + t5.y), h * (t5.y - t5.x));
        b[6] = mul_mj<INV>(b[6]);
        b[7] = INV ? make_float2(-h * (t7.x + t7.y), h * (t7.x - t7.y)) : make_float2(h * (t7.y - t7.x), -h * (t7.x + t7.y));
    }
    float2 c[8];
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
        c[g] = cadd2(b[g], b[g + 2]);     c[g + 2] = csub2(b[g], b[g + 2]);
        c[g + 1] = cadd2(b[g + 1], b[g + 3]); c[g + 3] = mul_mj<INV>(csub2(b[g + 1], b[g + 3]));
    }
    float2 d[8];
#pragma unroll
    for (int g = 0; g < 8; g += 2) { d[g] = cadd2(c[g], c[g + 1]); d[g + 1] = csub2(c[g], c[g + 1]); }
    a[0] = d[0]; a[4] = d[1]; a[2] = d[2]; a[6] = d[3]; a[1] = d[4]; a[5] = d[5]; a[3] = d[6]; a[7] = d[7];
}
template <bool INV>
__device__ __forceinline__ void dft4(float2 (&a)[4]) {
    const float2 s0 = cadd2(a[0], a[2]), d0 = csub2(a[0], a[2]);
    const float2 s1 = cadd2(a[1], a[3]), d1 = mul_mj<INV>(csub2(a[1], a[3]));
    a[0] = cadd2(s0, s1); a[2] = csub2(s0, s1); a[1] = cadd2(d0, d1); a[3] = csub2(d0, d1);
}

// std::abs(std::complex<float>) = hypotf: the sum of squares in double, rounded once (:683)
__device__ __forceinline__ float cabs_d(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}

struct StageArgs {
    float2* data;                 // [batch][kN] in place
    const float* real_in;         // forward mode 1: real input rows instead of data (may be null)
    long long real_stride; int n_in;
    float2* prod_up; float2* prod_dn;            // forward mode 3: write v * tu, v * td instead of v
    const float2* tmpl_up; const float2* tmpl_dn;
    const float2* tw1; const float2* tw2;
    float scale;
    float* mag_out;               // inverse mode 1: write |x| as fp32 [batch][2 kN floats per window at mag_stride] instead of x
    long long mag_stride;
};

template <int MODE, bool INV>
__global__ void __launch_bounds__(kFftThreads)
fft_stage_kernel(const StageArgs a) {
    constexpr int L = (MODE == 1) ? kA : 64;
    constexpr int R2 = L / 8;                         // second radix: 4 or 8
    constexpr int T = (MODE == 1) ? 32 : 16;          // transforms per tile
    constexpr int kRow = (MODE == 3) ? 72 : 0;        // mode 3: per-transform row of 8 x 9
    __shared__ float2 buf[(MODE == 3) ? T * kRow : L * T];
    __shared__ float2 wl[L];                          // w_L^m, conjugated for the inverse
    const int tid = threadIdx.x;
    const size_t f = blockIdx.y;
    float2* x = a.data + f * kN;
    if (tid < L) {
        double s, c;
        sincospi(-2.0 * tid / L, &s, &c);
        wl[tid] = make_float2(static_cast<float>(c), static_cast<float>(INV ? -s : s));
    }
    __syncthreads();

    for (int qq = 0; qq < kTilesPerCta; ++qq) {
        const int q = blockIdx.x * kTilesPerCta + qq;
        // element (i, t) of the tile lives at x[elem0 + i * istride + t * tstride]
        size_t elem0; int istride, tstride;
        if (MODE == 1) { elem0 = static_cast<size_t>(q) * T; istride = kB * kC; tstride = 1; }
        else if (MODE == 2) {
            const int ka = q / (kC / T), c0 = (q % (kC / T)) * T;
            elem0 = static_cast<size_t>(ka) * (kB * kC) + c0; istride = kC; tstride = 1;
        } else { elem0 = static_cast<size_t>(q) * T * kC; istride = 1; tstride = kC; }

        // ---- pass 1: radix 8 over n1 (n = R2 * n1 + n2), twiddle w_L^(n2 k1) ----
        {
            int n2, t;
            if (MODE == 3) { n2 = tid & 7; t = tid >> 3; } else { t = tid % T; n2 = tid / T; }
            float2 v[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const size_t idx = elem0 + static_cast<size_t>(R2 * n1 + n2) * istride + static_cast<size_t>(t) * tstride;
                if (MODE == 1 && !INV && a.real_in) v[n1] = make_float2(idx < static_cast<size_t>(a.n_in) ? a.real_in[f * a.real_stride + idx] : 0.0f, 0.0f);
                else v[n1] = x[idx];
            }
            dft8<INV>(v);
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                const float2 w = wl[n2 * k1];                      // n2 * k1 < L
                const float2 y = (k1 == 0) ? v[0] : cmulf(v[k1], w);
                if (MODE == 3) buf[t * kRow + k1 * 9 + n2] = y;
                else buf[(k1 * R2 + n2) * T + t] = y;
            }
        }
        __syncthreads();
        // ---- pass 2: radix R2 over n2 -> X[k1 + 8 k2], pointwise factor, store ----
        for (int task = tid; task < 8 * T; task += kFftThreads) {
            int k1, t;
            if (MODE == 3) { k1 = task & 7; t = task >> 3; } else { t = task % T; k1 = task / T; }
            float2 v[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) v[n2] = (MODE == 3) ? buf[t * kRow + k1 * 9 + n2] : buf[(k1 * R2 + n2) * T + t];
            if (R2 == 8) dft8<INV>(reinterpret_cast<float2 (&)[8]>(v)); else dft4<INV>(reinterpret_cast<float2 (&)[4]>(v));
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const int i = k1 + 8 * k2;
                const size_t idx = elem0 + static_cast<size_t>(i) * istride + static_cast<size_t>(t) * tstride;
                float2 val = v[k2];
                if (MODE == 1) {
                    if (!INV) val = cmulf(val, a.tw1[idx]);
                    else { val.x *= a.scale; val.y *= a.scale; }
                } else if (MODE == 2) {
                    const int c = static_cast<int>(idx % kC);
                    if (!INV) val = cmulf(val, a.tw2[i * kC + c]);
                    else { float2 w = a.tw1[idx]; w.y = -w.y; val = cmulf(val, w); }
                } else if (INV) {
                    const int row = static_cast<int>(idx / kC);               // row = ka*B + kb
                    float2 w = a.tw2[(row % kB) * kC + i]; w.y = -w.y;
                    val = cmulf(val, w);
                }
                if (MODE == 3 && !INV && a.prod_up) {
                    a.prod_up[f * kN + idx] = cmulf(val, a.tmpl_up[idx]);
                    a.prod_dn[f * kN + idx] = cmulf(val, a.tmpl_dn[idx]);
                } else if (MODE == 1 && INV && a.mag_out) {
                    // the peak search only needs the magnitude: half the bytes, and the double-precision
                    // square root moves from the latency-bound peak kernel into this bandwidth-bound stage
                    a.mag_out[f * a.mag_stride + idx] = cabs_d(val);
                } else x[idx] = val;
            }
        }
        __syncthreads();
    }
}

// Forward stage 3, the two template products and inverse stage 3 in one kernel.  Forward and inverse use
// mirrored stage orders, so the innermost pair works on the same tiles (16 rows of 64 contiguous bins):
// the products never leave the SM, which removes one write and one read of both 1 MB product spectra
// (4 of ~18 MB of HBM traffic per window) and two launches.  Output: prod_up / prod_dn hold the result
// of inverse stage 3, ready for inverse stages 2 and 1.
__global__ void __launch_bounds__(kFftThreads)
fft_stage3_fused_kernel(const StageArgs a) {
    constexpr int L = 64, T = 16, kRow = 72;
    __shared__ float2 buf[T * kRow];
    __shared__ float2 pu[T * kRow];                    // products, natural order inside a row
    __shared__ float2 pd[T * kRow];
    __shared__ float2 wl[L];                           // w_64^m (forward); the inverse uses the conjugate
    const int tid = threadIdx.x;
    const size_t f = blockIdx.y;
    const float2* x = a.data + f * kN;
    if (tid < L) {
        double s, c;
        sincospi(-2.0 * tid / L, &s, &c);
        wl[tid] = make_float2(static_cast<float>(c), static_cast<float>(s));
    }
    __syncthreads();
    for (int qq = 0; qq < kTilesPerCta; ++qq) {
        const int q = blockIdx.x * kTilesPerCta + qq;
        const size_t elem0 = static_cast<size_t>(q) * T * kC;
        const int n2 = tid & 7, t = tid >> 3;          // pass-1 role
        const int k1 = tid & 7;                        // pass-2 role (same t)
        // ---- forward pass 1 ----
        {
            float2 v[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) v[n1] = x[elem0 + static_cast<size_t>(8 * n1 + n2) + static_cast<size_t>(t) * kC];
            dft8<false>(v);
#pragma unroll
            for (int k = 0; k < 8; ++k) buf[t * kRow + k * 9 + n2] = (k == 0) ? v[0] : cmulf(v[k], wl[n2 * k]);
        }
        __syncthreads();
        // ---- forward pass 2 + template products ----
        {
            float2 v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = buf[t * kRow + k1 * 9 + m];
            dft8<false>(v);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const int i = k1 + 8 * k2;
                const size_t idx = elem0 + static_cast<size_t>(i) + static_cast<size_t>(t) * kC;
                pu[t * kRow + i] = cmulf(v[k2], a.tmpl_up[idx]);
                pd[t * kRow + i] = cmulf(v[k2], a.tmpl_dn[idx]);
            }
        }
        __syncthreads();
        // ---- inverse stage 3 on both products ----
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const float2* prod = which ? pd : pu;
            float2* out = (which ? a.prod_dn : a.prod_up) + f * kN;
            {
                float2 v[8];
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) v[n1] = prod[t * kRow + 8 * n1 + n2];
                dft8<true>(v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float2 w = wl[n2 * k]; w.y = -w.y;
                    buf[t * kRow + k * 9 + n2] = (k == 0) ? v[0] : cmulf(v[k], w);
                }
            }
            __syncthreads();
            {
                float2 v[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = buf[t * kRow + k1 * 9 + m];
                dft8<true>(v);
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const int i = k1 + 8 * k2;
                    const size_t idx = elem0 + static_cast<size_t>(i) + static_cast<size_t>(t) * kC;
                    const int row = static_cast<int>(idx / kC);                // row = ka*B + kb
                    float2 w = a.tw2[(row % kB) * kC + i]; w.y = -w.y;
                    out[idx] = cmulf(v[k2], w);
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// "Slab" kernel: forward stages 2 and 3, both template products and inverse stages 3 and 2 in one pass
// over HBM.  For a fixed ka the (b, c) plane is one contiguous 32 KB slab of the spectrum, so a CTA loads
// it with fully coalesced accesses, runs the 64-point transforms over b and over c in shared memory,
// multiplies by the two template slabs and runs the two inverse transforms back, writing two slabs.
// With it the three 131072-point transforms are: stage 1 (real input) -> slab kernel -> inverse stage 1
// for each template = 7.5 MB of HBM traffic per window instead of 14.5 MB (and 21.5 MB originally).
// MEASURED (B200, profiles/r1_chirp_subbatch.txt): correct (same peak indices / correlations as the staged
// path) but not faster yet -- 4.37 vs 4.09 ms per 1024 windows: with half the bytes the slab kernel is
// bound by its nine block barriers and 2-way-conflicted strided shared-memory passes (about twice its
// issue bound), not by HBM.  It is therefore opt-in (RIA_CHIRP_SLAB=1) until its shared-memory passes are
// vectorised; tests/test_chirp_gpu.py runs it against the default path.
// (A variant that fused stages 1+2 over (a, b) planes of four c was measured first: 40 % fewer bytes
// but 32-byte global segments and two CTAs per SM made it slower still, 5.1 ms per 1024 windows.)
// ---------------------------------------------------------------------------------------------
constexpr int kSlabS = 68;                               // row stride (elements): both transform directions 2-way at worst
constexpr int kSlabElems = kB * kSlabS;                  // 4352 float2 = 34 KB per buffer
constexpr int kSlabThreads = 256;

// Four 64-point transforms at once (radix 8 x 8): the lanes of a warp are (hi, r), r = line 0..3.
// Element i of this lane's line is in_[base + i * step] (pass 1 reads `in_`, writes `wk`; pass 2 reads `wk`).
// On return v[k2] = X[hi + 8 k2] and every lane has finished reading `wk`.
template <bool INV>
__device__ __forceinline__ void dft64_line(const float2* in_, float2* wk, int base, int step, int hi,
                                           const float2* __restrict__ w64, float2 (&v)[8],
                                           const float2* __restrict__ mul = nullptr) {
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) v[n1] = in_[base + (8 * n1 + hi) * step];
    if (mul) {                                           // pointwise factor on the input (template product): mul[i] for element i
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) v[n1] = cmulf(v[n1], __ldg(mul + 8 * n1 + hi));
    }
    dft8<INV>(v);
    __syncwarp();                                        // in_ may alias wk: everyone has read before anyone writes
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        float2 w = w64[hi * k1];
        if (INV) w.y = -w.y;
        wk[base + (8 * k1 + hi) * step] = (k1 == 0) ? v[0] : cmulf(v[k1], w);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) v[n2] = wk[base + (8 * hi + n2) * step];
    dft8<INV>(v);
    __syncwarp();
}

__global__ void __launch_bounds__(kSlabThreads, 3)
fft_slab_kernel(const StageArgs a) {
    extern __shared__ __align__(16) float2 slab_smem[];
    float2* F = slab_smem;                               // forward spectrum of the slab
    float2* W = slab_smem + kSlabElems;                  // product / inverse work buffer
    __shared__ float2 w64[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hi = lane >> 2, r = lane & 3;
    const size_t f = blockIdx.y;
    const int ka = blockIdx.x;
    const size_t slab0 = static_cast<size_t>(ka) * (kB * kC);
    const float2* x = a.data + f * kN + slab0;
    if (tid < 64) { double sn, cs; sincospi(-2.0 * tid / 64, &sn, &cs); w64[tid] = make_float2(static_cast<float>(cs), static_cast<float>(sn)); }
    // ---- load the slab: 4096 contiguous points ----
    for (int e = tid; e < kB * kC / 2; e += kSlabThreads) {
        const float4 p = reinterpret_cast<const float4*>(x)[e];
        const int b = (2 * e) / kC, c = (2 * e) % kC;
        F[b * kSlabS + c] = make_float2(p.x, p.y);
        F[b * kSlabS + c + 1] = make_float2(p.z, p.w);
    }
    __syncthreads();
    float2 v[8];
    // ---- forward stage 2: over b (columns c), twiddle W_BC^(kb c) ----
#pragma unroll 1
    for (int g = warp; g < kC / 4; g += kSlabThreads / 32) {
        const int c = 4 * g + r;
        dft64_line<false>(F, F, c, kSlabS, hi, w64, v);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) { const int kb = hi + 8 * k2; F[kb * kSlabS + c] = cmulf(v[k2], a.tw2[kb * kC + c]); }
    }
    __syncthreads();
    // ---- forward stage 3: over c (rows kb) ----
#pragma unroll 1
    for (int g = warp; g < kB / 4; g += kSlabThreads / 32) {
        const int kb = 4 * g + r;
        dft64_line<false>(F, F, kb * kSlabS, 1, hi, w64, v);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) F[kb * kSlabS + hi + 8 * k2] = v[k2];
    }
    __syncthreads();
    // ---- per template: product, inverse stage 3 (over kc), inverse stage 2 (over kb), slab out ----
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        const float2* tm = (which ? a.tmpl_dn : a.tmpl_up) + slab0;
        float2* out = (which ? a.prod_dn : a.prod_up) + f * kN + slab0;
        // the template product rides on the loads of the first inverse pass (F is kept for the second template)
#pragma unroll 1
        for (int g = warp; g < kB / 4; g += kSlabThreads / 32) {
            const int kb = 4 * g + r;
            dft64_line<true>(F, W, kb * kSlabS, 1, hi, w64, v, tm + kb * kC);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const int c = hi + 8 * k2;
                float2 w = a.tw2[kb * kC + c]; w.y = -w.y;
                W[kb * kSlabS + c] = cmulf(v[k2], w);
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int g = warp; g < kC / 4; g += kSlabThreads / 32) {
            const int c = 4 * g + r;
            dft64_line<true>(W, W, c, kSlabS, hi, w64, v);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const int b = hi + 8 * k2;
                float2 w = a.tw1[slab0 + static_cast<size_t>(b) * kC + c]; w.y = -w.y;
                W[b * kSlabS + c] = cmulf(v[k2], w);
            }
        }
        __syncthreads();
        for (int e = tid; e < kB * kC / 2; e += kSlabThreads) {
            const int b = (2 * e) / kC, c = (2 * e) % kC;
            const float2 p0 = W[b * kSlabS + c], p1 = W[b * kSlabS + c + 1];
            reinterpret_cast<float4*>(out)[e] = make_float4(p0.x, p0.y, p1.x, p1.y);
        }
        __syncthreads();
    }
}

// =============================================================================================
// Residue-decomposed matched filter (default path): the 2^17-point transforms run out of shared memory.
//
// N = 8 x M, M = 16384.  With k = 8 q + r the forward transform splits into eight independent M-point transforms,
// one per residue r:   X[8 q + r] = sum_m W_M^{m q} z_r[m],   z_r[m] = W_N^{m r} sum_a x[M a + m] W_8^{a r}.
// A CTA owns one (window, r): it reads the real window once (the other seven CTAs of the window hit L2), forms z_r,
// runs the M-point transform in place in 128 KB of shared memory (four passes, radix 16 x 16 x 8 x 8, decimation in
// frequency; the spectrum stays in digit-reversed order, which the template spectra share), multiplies by the two
// template spectra and runs the mirrored decimation-in-time inverse for each, leaving v_r[m] in natural order.
// What reaches HBM is v_r for the two templates (2 MB per window).  The lag-domain result needs one radix-8
// butterfly over r,   R[M a + m] = sum_r v'_r[m] W_8^{-a r},   v'_r[m] = W_N^{-m r} v_r[m] / N,
// which chirp_combine_kernel applies while it forms the magnitudes for the peak search.
// Per window: 0.48 MB of samples + 2 MB out + 2 MB in + magnitudes instead of ~14.5 MB of spectrum passes.
//
// Shared-memory layout: point i lives at i ^ ((i >> 4) & 15).  For every pass the lanes of each half-warp then fall
// on sixteen different 8-byte bank pairs (thread mappings below), so all 64-bit accesses are conflict-free.
// Twiddles come from exact tables (pass B: W_1024^e, pass C: W_64^e) or, for the M-point pass-A factors W_M^{t k},
// from two exact 512-entry tables and one extra product (W_M^{(32 th + tl) k} = W_512^{th k} W_M^{tl k}).
// =============================================================================================
constexpr int kM = 16384;
constexpr int kResThreads = 512;

struct ResidueTables {            // device pointers, built once per context
    const float2* wa_hi;          // [16][32]  W_512^{th k}
    const float2* wa_lo;          // [16][32]  W_M^{tl k}
    const float2* wb;             // [16][64]  W_1024^{t k}
    const float2* wc;             // [64]      W_64^e
    const float2* wr_hi;          // [8][128]  W_N^{128 r mh}
    const float2* wr_lo;          // [8][128]  W_N^{r ml}
};

struct ResidueArgs {
    const float* samples; long long stride; int n_in;          // MODE 0: real windows
    const float2* cplx_in;                                      // MODE 1: one complex input of kN points
    const float2* tmpl_up; const float2* tmpl_dn;               // [8][kM] conj template spectra, residue / digit-reversed order
    float2* out_up; float2* out_dn;                             // MODE 0: [n_win][8][kM] v'_r[m];  MODE 1: out_up = [8][kM] conj spectrum
    float2* scratch;                                            // [gridDim.x][kM] second product, stays in L2
    ResidueTables t;
    int n_items;                                                // windows x 8
    float scale;                                                // 1 / N
};

__device__ __forceinline__ int res_sw(int i) { return i ^ ((i >> 4) & 15); }
// The file is compiled with --fmad=false (the bit-exact kernels need it); this transform is not bit-exact by
// construction (its order is not the reference's radix-2 order), so its complex products use explicit FMAs:
// four issue slots instead of six, and one rounding less per product.
__device__ __forceinline__ float2 cmulr(float2 a, float2 b) {   // a * b
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
}

// 16-point DFT, natural order in and out: 4 x 4 with the W_16 factors as constants
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    float2 b[4][4];                                             // b[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        float2 col[4] = {a[n2], a[4 + n2], a[8 + n2], a[12 + n2]};
        dft4<INV>(col);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) b[n2][k1] = col[k1];
    }
    // b[n2][k1] *= W_16^{n2 k1} (conjugated for the inverse)
    auto tw = [&](float2 v, float wr, float wi) {
        const float wim = INV ? -wi : wi;
        return make_float2(fmaf(v.x, wr, -(v.y * wim)), fmaf(v.x, wim, v.y * wr));
    };
    b[1][1] = tw(b[1][1], c1, -s1);  b[1][2] = tw(b[1][2], h, -h);    b[1][3] = tw(b[1][3], s1, -c1);
    b[2][1] = tw(b[2][1], h, -h);    b[2][2] = mul_mj<INV>(b[2][2]);  b[2][3] = tw(b[2][3], -h, -h);
    b[3][1] = tw(b[3][1], s1, -c1);  b[3][2] = tw(b[3][2], -h, -h);   b[3][3] = tw(b[3][3], -c1, s1);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 row[4] = {b[0][k1], b[1][k1], b[2][k1], b[3][k1]};
        dft4<INV>(row);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) a[k1 + 4 * k2] = row[k2];
    }
}

// ---- the four passes of the M-point transform on the swizzled tile D ----
// Forward (decimation in frequency): DFT first, twiddle after.  Inverse (decimation in time): conjugate twiddle
// first, inverse DFT after, passes in the opposite order.
// The swizzle only touches the low four address bits and reads bits 4..7, so inside a pass most of it is a constant of
// the thread (or of the unrolled loop index): each pass computes its swizzled base once and adds immediates.
template <bool INV>
__device__ __forceinline__ void res_pass_a(float2* D, const ResidueTables& t, int tid) {
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
        const int tp = tid + kResThreads * u;                   // t' in [0, 1024)
        const int th = tp >> 5, tl = tp & 31;
        float2* p = D + res_sw(tp);                             // + 1024 j leaves bits 0..7 alone
        float2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = p[1024 * j];
        if (INV) {
#pragma unroll
            for (int k = 1; k < 16; ++k) a[k] = cmulc(a[k], cmulr(__ldg(t.wa_hi + k * 32 + th), __ldg(t.wa_lo + k * 32 + tl)));
        }
        dft16<INV>(a);
        if (!INV) {
#pragma unroll
            for (int k = 1; k < 16; ++k) a[k] = cmulr(a[k], cmulr(__ldg(t.wa_hi + k * 32 + th), __ldg(t.wa_lo + k * 32 + tl)));
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) p[1024 * k] = a[k];
    }
}
template <bool INV>
__device__ __forceinline__ void res_pass_b(float2* D, const float2* wb, int tid) {
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
        const int tau = tid + kResThreads * u;
        const int tq = tau & 63, base = (tau >> 6) * 1024 + tq;
        // i = base + 64 j: bits 4, 5 come from tq, bits 6, 7 from j -> low nibble = (tq & 15) ^ (tq >> 4) ^ ((j & 3) << 2)
        const int lo = (tq & 15) ^ (tq >> 4);
        float2* hi = D + (base & ~15);
        const int l0 = lo, l1 = lo ^ 4, l2 = lo ^ 8, l3 = lo ^ 12;
        float2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = hi[64 * j + ((j & 3) == 0 ? l0 : (j & 3) == 1 ? l1 : (j & 3) == 2 ? l2 : l3)];
        if (INV) {
#pragma unroll
            for (int k = 1; k < 16; ++k) a[k] = cmulc(a[k], wb[k * 64 + tq]);
        }
        dft16<INV>(a);
        if (!INV) {
#pragma unroll
            for (int k = 1; k < 16; ++k) a[k] = cmulr(a[k], wb[k * 64 + tq]);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) hi[64 * k + ((k & 3) == 0 ? l0 : (k & 3) == 1 ? l1 : (k & 3) == 2 ? l2 : l3)] = a[k];
    }
}
template <bool INV>
__device__ __forceinline__ void res_pass_c1(float2* D, const float2* wc, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const int t3 = lane & 7;
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
        const int blk = ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (warp << 2) | (u << 6);
        // i = 64 blk + 8 j + t3: low nibble = t3 | (j & 1) << 3, bits 4..7 = (j >> 1) | (blk & 3) << 2
        const int v = t3 ^ ((blk & 3) << 2);                    // thread part of the swizzled low nibble
        float2* hi = D + blk * 64;
        float2 a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = hi[(8 * j & ~15) + (v ^ (((j & 1) << 3) ^ (j >> 1)))];
        if (INV) {
#pragma unroll
            for (int k = 1; k < 8; ++k) a[k] = cmulc(a[k], wc[t3 * k]);
        }
        dft8<INV>(a);
        if (!INV) {
#pragma unroll
            for (int k = 1; k < 8; ++k) a[k] = cmulr(a[k], wc[t3 * k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) hi[(8 * k & ~15) + (v ^ (((k & 1) << 3) ^ (k >> 1)))] = a[k];
    }
}
template <bool INV>
__device__ __forceinline__ void res_pass_c2(float2* D, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
        // g0 = lane bit 3, g1..g3 = lane bits 0..2, g4 = lane bit 4, g5.. = warp, u
        const int g = ((lane >> 3) & 1) | ((lane & 7) << 1) | (((lane >> 4) & 1) << 4) | (warp << 5) | (u << 9);
        // i = 8 g + j: low nibble = j | (g & 1) << 3, bits 4..7 = (g >> 1) & 15 -> thread part v, loop part j
        const int v = ((g & 1) << 3) ^ ((g >> 1) & 15);
        float2* hi = D + ((8 * g) & ~15);
        float2 a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = hi[v ^ j];
        dft8<INV>(a);
#pragma unroll
        for (int k = 0; k < 8; ++k) hi[v ^ k] = a[k];
    }
}

// MODE 0: correlate real windows against both templates.  MODE 1: conj spectrum of one complex input (template prep).
template <int MODE>
__global__ void __launch_bounds__(kResThreads, 1)
chirp_residue_kernel(const ResidueArgs a) {
    extern __shared__ __align__(16) float2 res_smem[];
    float2* D = res_smem;                                       // [kM], swizzled
    float2* wb = res_smem + kM;                                 // [1024]
    float2* wc = wb + 1024;                                     // [64]
    float2* wrh = wc + 64;                                      // [128]  W_N^{128 r mh}
    float2* wrl = wrh + 128;                                    // [128]  W_N^{r ml}
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += kResThreads) wb[i] = a.t.wb[i];
    if (tid < 64) wc[tid] = a.t.wc[tid];
    int loaded_r = -1;
    float2* my_scratch = a.scratch + static_cast<size_t>(blockIdx.x) * kM;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int w = item >> 3, r = item & 7;
        __syncthreads();                                        // previous item done with D and the r tables
        if (r != loaded_r) {
            if (tid < 128) { wrh[tid] = a.t.wr_hi[r * 128 + tid]; wrl[tid] = a.t.wr_lo[r * 128 + tid]; }
            loaded_r = r;
        }
        __syncthreads();
        // ---- load: z_r[m] = W_N^{m r} sum_a x[M a + m] W_8^{a r} ----
        if (MODE == 0) {
            // Real input: with W_8^{(a+4) r} = (-1)^r W_8^{a r} the eight-term sum folds to four real pairs,
            //   z = sum_{a<4} (x_a + sg x_{a+4}) (c_a - j s_a),  c_a + j s_a = e^{j 2 pi a r / 8}:  11 FMAs per point.
            // Consecutive lanes take consecutive m: coalesced 128-byte rows, conflict-free table reads and stores.
            const float* x = a.samples + static_cast<long long>(w) * a.stride;
            const float h = 0.70710678118654752440f;
            const float sg = (r & 1) ? -1.0f : 1.0f;
            // (c_a, s_a) for a = 1, 2, 3 (a = 0 is (1, 0)): e^{j pi a r / 4}
            const float ct[8] = {1.f, h, 0.f, -h, -1.f, -h, 0.f, h}, st[8] = {0.f, h, 1.f, h, 0.f, -h, -1.f, -h};
            const float c1 = ct[r & 7], s1 = st[r & 7], c2 = ct[(2 * r) & 7], s2 = st[(2 * r) & 7], c3 = ct[(3 * r) & 7], s3 = st[(3 * r) & 7];
#pragma unroll 4
            for (int m = tid; m < kM; m += kResThreads) {
                float xa[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { const int idx = kM * q + m; xa[q] = (idx < a.n_in) ? __ldg(x + idx) : 0.0f; }
                const float u0 = fmaf(sg, xa[4], xa[0]), u1 = fmaf(sg, xa[5], xa[1]), u2 = fmaf(sg, xa[6], xa[2]), u3 = fmaf(sg, xa[7], xa[3]);
                const float re = fmaf(u3, c3, fmaf(u2, c2, fmaf(u1, c1, u0)));
                const float im = -fmaf(u3, s3, fmaf(u2, s2, u1 * s1));
                D[res_sw(m)] = cmulr(make_float2(re, im), cmulr(wrh[m >> 7], wrl[m & 127]));
            }
        } else {
#pragma unroll 1
            for (int m = tid; m < kM; m += kResThreads) {
                float2 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = a.cplx_in[kM * q + m];
                dft8<false>(v);
                float2 zr = v[0];
#pragma unroll
                for (int q = 1; q < 8; ++q) if (r == q) zr = v[q];
                D[res_sw(m)] = cmulr(zr, cmulr(wrh[m >> 7], wrl[m & 127]));
            }
        }
        __syncthreads();
        // ---- forward M-point transform ----
        res_pass_a<false>(D, a.t, tid);      __syncthreads();
        res_pass_b<false>(D, wb, tid);       __syncthreads();
        res_pass_c1<false>(D, wc, tid);      __syncthreads();
        res_pass_c2<false>(D, tid);          __syncthreads();
        if (MODE == 1) {
            float2* out = a.out_up + static_cast<size_t>(r) * kM;
            for (int p = tid; p < kM; p += kResThreads) { const float2 z = D[res_sw(p)]; out[p] = make_float2(z.x, -z.y); }
            continue;
        }
        // ---- both template products: the up product stays in shared memory, the down product waits in L2 ----
        {
            const float2* tu = a.tmpl_up + static_cast<size_t>(r) * kM;
            const float2* td = a.tmpl_dn + static_cast<size_t>(r) * kM;
#pragma unroll 4
            for (int p = tid; p < kM; p += kResThreads) {
                const float2 z = D[res_sw(p)];
                my_scratch[p] = cmulr(z, __ldg(td + p));
                D[res_sw(p)] = cmulr(z, __ldg(tu + p));
            }
        }
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            __syncthreads();
            if (which == 1) {
#pragma unroll 4
                for (int p = tid; p < kM; p += kResThreads) D[res_sw(p)] = my_scratch[p];
                __syncthreads();
            }
            res_pass_c2<true>(D, tid);       __syncthreads();
            res_pass_c1<true>(D, wc, tid);   __syncthreads();
            res_pass_b<true>(D, wb, tid);    __syncthreads();
            res_pass_a<true>(D, a.t, tid);   __syncthreads();
            float2* out = (which ? a.out_dn : a.out_up) + (static_cast<size_t>(w) * 8 + r) * kM;
#pragma unroll 4
            for (int m = tid; m < kM; m += kResThreads) {
                const float2 v = cmulc(D[res_sw(m)], cmulr(wrh[m >> 7], wrl[m & 127]));
                out[m] = make_float2(v.x * a.scale, v.y * a.scale);
            }
        }
    }
}

// R[M a + m] = sum_r v'_r[m] W_8^{-a r} for both templates, magnitude only (what the peak search reads)
struct CombineArgs {
    const float2* v_up; const float2* v_dn;                     // [n_win][8][kM]
    float* mag_up; float* mag_dn; long long mag_stride;         // [n_win] rows of |R| per lag
    int n_lags;                                                 // lags the peak search can ask for (< kN)
};
__global__ void __launch_bounds__(256)
chirp_combine_kernel(const CombineArgs a) {
    const int m = blockIdx.x * 256 + threadIdx.x;               // < kM
    const size_t w = blockIdx.y;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        const float2* v = (which ? a.v_dn : a.v_up) + w * 8 * kM + m;
        float* mag = (which ? a.mag_dn : a.mag_up) + w * a.mag_stride;
        float2 x[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) x[r] = __ldcs(v + static_cast<size_t>(r) * kM);
        dft8<true>(x);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int lag = kM * q + m;
            if (lag < a.n_lags) mag[lag] = cabs_d(x[q]);
        }
    }
}

__global__ void residue_tables_kernel(float2* wa_hi, float2* wa_lo, float2* wb, float2* wc, float2* wr_hi, float2* wr_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto w = [](double num, double den) { double s, c; sincospi(-2.0 * num / den, &s, &c); return make_float2(static_cast<float>(c), static_cast<float>(s)); };
    if (i < 512) {                                              // [k][32]
        const int k = i >> 5, t = i & 31;
        wa_hi[i] = w(static_cast<double>(t) * k, 512.0);
        wa_lo[i] = w(static_cast<double>(t) * k, static_cast<double>(kM));
    }
    if (i < 1024) { const int k = i >> 6, t = i & 63; wb[i] = w(static_cast<double>(t) * k, 1024.0); }
    if (i < 64) wc[i] = w(static_cast<double>(i), 64.0);
    if (i < 1024) {
        const int r = i >> 7, j = i & 127;
        wr_hi[i] = w(128.0 * r * j, static_cast<double>(kN));
        wr_lo[i] = w(static_cast<double>(r) * j, static_cast<double>(kN));
    }
}

template <bool INV>
void fft_stages(const StageArgs& a, int batch, cudaStream_t s) {
    const dim3 g1(kB * kC / 32 / kTilesPerCta, batch), g2(kA * kC / 16 / kTilesPerCta, batch), g3(kA * kB / 16 / kTilesPerCta, batch);
    if (!INV) {
        fft_stage_kernel<1, false><<<g1, kFftThreads, 0, s>>>(a);
        fft_stage_kernel<2, false><<<g2, kFftThreads, 0, s>>>(a);
        fft_stage_kernel<3, false><<<g3, kFftThreads, 0, s>>>(a);
    } else {
        fft_stage_kernel<3, true><<<g3, kFftThreads, 0, s>>>(a);
        fft_stage_kernel<2, true><<<g2, kFftThreads, 0, s>>>(a);
        fft_stage_kernel<1, true><<<g1, kFftThreads, 0, s>>>(a);
    }
}
void fft_forward(float2* d, int batch, const ChirpTablesDev& t, cudaStream_t s) {
    StageArgs a{};
    a.data = d; a.tw1 = t.tw1; a.tw2 = t.tw2; a.scale = 1.0f;
    fft_stages<false>(a, batch, s);
}
void fft_inverse(float2* d, int batch, const ChirpTablesDev& t, cudaStream_t s, float* mag_out = nullptr, long long mag_stride = 0) {
    StageArgs a{};
    a.data = d; a.tw1 = t.tw1; a.tw2 = t.tw2; a.scale = 1.0f / kN; a.mag_out = mag_out; a.mag_stride = mag_stride;
    fft_stages<true>(a, batch, s);
}
// forward transform of real rows -> template products -> inverse stage 3 (fused), then inverse stages 2, 1:
// leaves the two correlations in pu / pd
// magnitudes in mag_up / mag_dn ([batch] rows of mag_stride floats; may live in `work`, which is free after stage 3)
void chirp_correlate(const float* samples, long long stride, int n_in, float2* work, float2* pu, float2* pd,
                     float* mag_up, float* mag_dn, long long mag_stride,
                     int batch, const ChirpTablesDev& t, cudaStream_t s, bool use_slab) {
    StageArgs a{};
    a.data = work; a.real_in = samples; a.real_stride = stride; a.n_in = n_in;
    a.prod_up = nullptr; a.prod_dn = nullptr; a.tmpl_up = t.tmpl_up; a.tmpl_dn = t.tmpl_dn;
    a.tw1 = t.tw1; a.tw2 = t.tw2; a.scale = 1.0f;
    const dim3 g1(kB * kC / 32 / kTilesPerCta, batch), g2(kA * kC / 16 / kTilesPerCta, batch), g3(kA * kB / 16 / kTilesPerCta, batch);
    if (use_slab) {
        const size_t slab_smem_bytes = 2 * static_cast<size_t>(kSlabElems) * sizeof(float2);
        fft_stage_kernel<1, false><<<g1, kFftThreads, 0, s>>>(a);
        a.prod_up = pu; a.prod_dn = pd;
        fft_slab_kernel<<<dim3(kA, batch), kSlabThreads, slab_smem_bytes, s>>>(a);
    } else {
        fft_stage_kernel<1, false><<<g1, kFftThreads, 0, s>>>(a);
        fft_stage_kernel<2, false><<<g2, kFftThreads, 0, s>>>(a);
        a.prod_up = pu; a.prod_dn = pd;
        fft_stage3_fused_kernel<<<g3, kFftThreads, 0, s>>>(a);
    }
    for (int which = 0; which < 2; ++which) {
        StageArgs b{};
        b.data = which ? pd : pu; b.tw1 = t.tw1; b.tw2 = t.tw2; b.scale = 1.0f / kN;
        b.mag_out = which ? mag_dn : mag_up; b.mag_stride = mag_stride;
        if (!use_slab) fft_stage_kernel<2, true><<<g2, kFftThreads, 0, s>>>(b);
        fft_stage_kernel<1, true><<<g1, kFftThreads, 0, s>>>(b);
    }
}

// forward transform of real rows with the two template products as output
void fft_forward_real_to_products(const float* samples, long long stride, int n_in, float2* work, float2* pu, float2* pd,
                                  int batch, const ChirpTablesDev& t, cudaStream_t s) {
    StageArgs a{};
    a.data = work; a.real_in = samples; a.real_stride = stride; a.n_in = n_in;
    a.prod_up = pu; a.prod_dn = pd; a.tmpl_up = t.tmpl_up; a.tmpl_dn = t.tmpl_dn;
    a.tw1 = t.tw1; a.tw2 = t.tw2; a.scale = 1.0f;
    fft_stages<false>(a, batch, s);
}

__global__ void chirp_twiddle_kernel(float2* tw1, float2* tw2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kN) {
        const int ka = i / (kB * kC), m = i % (kB * kC);
        double s, c;
        sincospi(-2.0 * (static_cast<double>(ka) * m) / kN, &s, &c);
        tw1[i] = make_float2(static_cast<float>(c), static_cast<float>(s));
    }
    if (i < kB * kC) {
        const int kb = i / kC, cc = i % kC;
        double s, c;
        sincospi(-2.0 * (static_cast<double>(kb) * cc) / (kB * kC), &s, &c);
        tw2[i] = make_float2(static_cast<float>(c), static_cast<float>(s));
    }
}

__global__ void conj_kernel(float2* d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kN) d[i].y = -d[i].y;
}

__device__ __forceinline__ void better(float& bv, int& bi, float v, int i) {
    if (v > bv || (v == bv && v > 0.0f && (bi < 0 || i < bi))) { bv = v; bi = i; }
}

struct PeakArgs {
    const float* samples; long long frame_stride; int window;
    const float* corr_up; const float* corr_dn; long long corr_stride;      // |R| per lag (fp32)
    float* cumsum;                 // [n][window + 1] scratch
    const float2* tmpl_dn_time;
    float threshold, energy_up, energy_dn;
    int chirp_len, gap;
    float sample_rate, f_start, f_end, duration_ms;
    ria_sync_result* out;
};

// Sequential fp32 prefix sum of s^2 over [start, start+len) (:666-669): c[i+1] = fl(c[i] + fl(s[i]^2)).
// The order of the adds is part of the result (the rounding of every step depends on the running sum), and
// round 1 had ONE thread walk the 170 k-step chain per window -- a single active lane costing full issue
// slots, 43 % of the kernel's instructions.  The chain is evaluated in parallel here, EXACTLY:
//
//  * While the running sum stays inside one binade (ulp u), fl(c + t) = c + u * rne(t / u): the increment does
//    not depend on c, except for exact ties (t / u = n + 1/2), where round-to-even looks at the PARITY of c's
//    mantissa.  A chunk of L elements is therefore a function parity -> increment (in ulps), and such functions
//    compose associatively: (a then b)(p) = a(p) + b(p xor (a(p) & 1)).
//  * Each thread runs its chunk twice with real float adds from two hypothetical sums of this binade with even
//    and odd mantissa (2^E and 2^E + u): that gives its function (n0, n1) without any integer emulation of the
//    rounding.  A block scan of the composition gives every thread the mantissa it really starts from.
//  * Each thread then runs its chunk a third time from that true start, again with real float adds -- these are
//    the stored values, so they are the reference's values whenever the start is right.
//  * The start of thread j is right iff it equals what thread j-1 really ended on.  That comparison is made for
//    every thread; everything before the first mismatch (a binade crossing, or a chunk that left the binade) is
//    final, the tile restarts there with the new binade.  Thread 0 always starts from a known-exact value, so
//    every pass makes progress and the result is exact by construction; prediction only affects speed.
// Chunk length grows with the position (binade crossings are exponentially rarer as the sum grows).
// Block-collective; ends with a barrier.
constexpr int kScanThreads = 256;
constexpr int kScanMaxChunk = 32;
constexpr int kScanTile = kScanThreads * kScanMaxChunk;          // 8192 elements
constexpr int kScanPad = kScanTile + kScanTile / 32;             // one pad word per 32: lane-strided chunks hit distinct banks

struct ScanSmem {
    float t[kScanPad];
    float c_end[kScanThreads];
    unsigned wn0[kScanThreads / 32], wn1[kScanThreads / 32];
    float c_start;
    int valid, first_nz;
};

__device__ __forceinline__ int scan_pad(int e) { return e + (e >> 5); }
__device__ __forceinline__ unsigned sat_add(unsigned a, unsigned b) { const unsigned s = a + b; return s > 0x40000000u ? 0x40000000u : s; }
// (a then b): parity -> ulps
__device__ __forceinline__ void scan_compose(unsigned a0, unsigned a1, unsigned b0, unsigned b1, unsigned& r0, unsigned& r1) {
    r0 = sat_add(a0, (a0 & 1u) ? b1 : b0);
    r1 = sat_add(a1, (a1 & 1u) ? b0 : b1);
}

__device__ void prefix_energy(const float* __restrict__ s, int start, int len, float* __restrict__ c, ScanSmem& sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { c[0] = 0.0f; sm.c_start = 0.0f; }
    int pos = 0;
    __syncthreads();
    while (pos < len) {
        const float c_start = sm.c_start;
        // ---- chunk length for this pass ----
        int L = kScanMaxChunk;
        if (pos < kScanTile) { L = 1; while (L < kScanMaxChunk && L * kScanThreads < pos) L <<= 1; }
        const int tile = min(L * kScanThreads, len - pos);
        // ---- stage t = s^2 ----
        for (int e = tid; e < L * kScanThreads; e += kScanThreads) {
            float v = 0.0f;
            if (e < tile) v = s[start + pos + e];
            sm.t[scan_pad(e)] = __fmul_rn(v, v);
        }
        if (tid == 0) { sm.valid = tile; sm.first_nz = tile; }
        __syncthreads();
        if (c_start == 0.0f || pos < 64 || (__float_as_uint(c_start) >> 23) < 24u) {      // (sums below 2^-103: 2^(23-E) overflows)
            // Start of the chain: zeros add nothing, and right after the first non-zero term the sum changes binade
            // with almost every add.  Skip the zeros in parallel, then let one thread take up to 64 steps.
            if (c_start == 0.0f) {
                for (int e = tid; e < tile; e += kScanThreads)
                    if (sm.t[scan_pad(e)] != 0.0f) { atomicMin(&sm.first_nz, e); break; }
                __syncthreads();
            } else if (tid == 0) sm.first_nz = 0;
            __syncthreads();
            const int z = sm.first_nz;                                   // c stays c_start over [0, z)
            for (int e = tid; e < z; e += kScanThreads) c[pos + e + 1] = c_start;
            const int n_seq = min(64, tile - z);
            if (tid == 0) {
                float acc = c_start;
                for (int e = z; e < z + n_seq; ++e) { acc = __fadd_rn(acc, sm.t[scan_pad(e)]); c[pos + e + 1] = acc; }
                sm.c_start = acc;
            }
            pos += z + n_seq;
            __syncthreads();
            continue;
        }
        // ---- pass 1: the chunk as a function parity -> ulps, from two hypothetical sums of this binade ----
        const unsigned cs_bits = __float_as_uint(c_start);
        const float h0 = __uint_as_float(cs_bits & 0x7f800000u);         // 2^E: mantissa 2^23, even
        const float h1 = __uint_as_float((cs_bits & 0x7f800000u) | 1u);  // 2^E + ulp: odd
        const float inv_u = __uint_as_float((150u + 127u - (cs_bits >> 23)) << 23);   // 2^(23 - E), E = exponent - 127
        const int base = tid * L;
        float a0 = h0, a1 = h1;
#pragma unroll 4
        for (int k = 0; k < L; ++k) {
            const float t = sm.t[scan_pad(base + k)];
            a0 = __fadd_rn(a0, t); a1 = __fadd_rn(a1, t);
        }
        // inside the binade the differences are exact multiples of u below 2^24; beyond it the values are
        // garbage and the start check below rejects everything that depends on them
        const float d0 = __fmul_rn(__fsub_rn(a0, h0), inv_u), d1 = __fmul_rn(__fsub_rn(a1, h1), inv_u);
        unsigned n0 = (d0 >= 0.0f && d0 < 16777216.0f) ? static_cast<unsigned>(d0) : 0x40000000u;
        unsigned n1 = (d1 >= 0.0f && d1 < 16777216.0f) ? static_cast<unsigned>(d1) : 0x40000000u;
        // ---- block scan of the composition (inclusive over lanes, then over warps) ----
        unsigned i0 = n0, i1 = n1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned p0 = __shfl_up_sync(0xffffffffu, i0, d), p1 = __shfl_up_sync(0xffffffffu, i1, d);
            if (lane >= d) { unsigned r0, r1; scan_compose(p0, p1, i0, i1, r0, r1); i0 = r0; i1 = r1; }
        }
        if (lane == 31) { sm.wn0[warp] = i0; sm.wn1[warp] = i1; }
        __syncthreads();
        // exclusive prefix of this thread: warps before it, then lanes before it
        unsigned e0 = 0u, e1 = 0u;
        for (int w = 0; w < warp; ++w) { unsigned r0, r1; scan_compose(e0, e1, sm.wn0[w], sm.wn1[w], r0, r1); e0 = r0; e1 = r1; }
        {
            unsigned q0 = __shfl_up_sync(0xffffffffu, i0, 1), q1 = __shfl_up_sync(0xffffffffu, i1, 1);
            if (lane == 0) { q0 = 0u; q1 = 0u; }
            unsigned r0, r1; scan_compose(e0, e1, q0, q1, r0, r1); e0 = r0; e1 = r1;
        }
        const unsigned inc = (cs_bits & 1u) ? e1 : e0;
        // adding ulps to a float inside its binade is an integer add on its bit pattern
        const float c_in = (inc < 0x01000000u) ? __uint_as_float(cs_bits + inc) : __uint_as_float(0x7fc00000u);
        // ---- pass 2: the real chain from the predicted start; these are the values that get stored ----
        float acc = c_in;
#pragma unroll 4
        for (int k = 0; k < L; ++k) {
            const int a = scan_pad(base + k);
            acc = __fadd_rn(acc, sm.t[a]);
            sm.t[a] = acc;
        }
        sm.c_end[tid] = acc;
        __syncthreads();
        // ---- which starts were right?  thread j's start must be what thread j-1 really ended on ----
        if (tid > 0 && base < tile) {
            if (__float_as_uint(sm.c_end[tid - 1]) != __float_as_uint(c_in)) atomicMin(&sm.valid, base);
        }
        __syncthreads();
        const int valid = sm.valid;                                       // >= L: thread 0 starts from the exact sum
        for (int e = tid; e < valid; e += kScanThreads) c[pos + e + 1] = sm.t[scan_pad(e)];
        if (tid == 0) sm.c_start = sm.c_end[(valid - 1) / L];
        pos += valid;
        __syncthreads();
    }
}

// normalised peak over pos < search_len (:677-689); corr index offset `off`
__device__ void peak_search(const float* corr, int off, const float* c, int search_len, int chirp_len,
                            float tmpl_energy, float* red_v, int* red_i, float* best, int* pos) {
    const int tid = threadIdx.x;
    float bv = 0.0f; int bi = -1;
    // four positions per trip with all twelve loads issued first: one position per trip left the loop
    // waiting on its own global loads (ncu: 60 % of the kernel's stall samples sat on these three loads)
    const int stride = blockDim.x;
    int p = tid;
    for (; p + 3 * stride < search_len; p += 4 * stride) {
        float mag[4], hi[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = p + u * stride;
            mag[u] = __ldg(corr + off + q); hi[u] = c[q + chirp_len]; lo[u] = c[q];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float e = __fsub_rn(hi[u], lo[u]);
            const float denom = sqrtf(__fmul_rn(e, tmpl_energy));
            const float nc = (denom > 1e-10f) ? __fdiv_rn(mag[u], denom) : 0.0f;
            better(bv, bi, nc, p + u * stride);
        }
    }
    for (; p < search_len; p += stride) {
        const float mag = corr[off + p];
        const float e = __fsub_rn(c[p + chirp_len], c[p]);
        const float denom = sqrtf(__fmul_rn(e, tmpl_energy));
        const float nc = (denom > 1e-10f) ? __fdiv_rn(mag, denom) : 0.0f;
        better(bv, bi, nc, p);
    }
    red_v[tid] = bv; red_i[tid] = bi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) { float v = red_v[tid]; int i = red_i[tid]; better(v, i, red_v[tid + s], red_i[tid + s]); red_v[tid] = v; red_i[tid] = i; }
        __syncthreads();
    }
    *best = red_v[0]; *pos = red_i[0];
    __syncthreads();
}

// computeComplexTemplateCorrelation (:823-846): sequential fp32 sums over the whole template
__device__ float td_corr(const float* s, int offset, const float2* tmpl, int CL, float tmpl_energy) {
    float ci = 0.0f, cq = 0.0f, en = 0.0f;
    for (int i = 0; i < CL; ++i) {
        const float v = s[offset + i];
        const float2 t = tmpl[i];
        ci = __fadd_rn(ci, __fmul_rn(v, t.x));
        cq = __fadd_rn(cq, __fmul_rn(v, t.y));
        en = __fadd_rn(en, __fmul_rn(v, v));
    }
    const float denom = sqrtf(__fmul_rn(en, tmpl_energy));
    if (denom < 1e-10f) return 0.0f;
    return __fdiv_rn(sqrtf(__fadd_rn(__fmul_rn(ci, ci), __fmul_rn(cq, cq))), denom);
}

// time-domain fallback of detectChirpTemplate (:745-817) on the slice s[0 .. len)
__device__ void td_detect(const float* s, int len, const float2* tmpl, int CL, float tmpl_energy, float threshold,
                          float* vals /*>= 640 floats of scratch*/, float* out_corr, int* out_pos) {
    __shared__ float sh_best;
    __shared__ int sh_pos;
    const int tid = threadIdx.x;
    const int search_len = len - CL;
    const int n_coarse = (search_len + 47) / 48;
    for (int k = tid; k < n_coarse; k += blockDim.x) vals[k] = td_corr(s, k * 48, tmpl, CL, tmpl_energy);
    __syncthreads();
    if (tid == 0) {
        float best = 0.0f; int pos = -1;
        for (int k = 0; k < n_coarse; ++k) if (vals[k] > best) { best = vals[k]; pos = k * 48; }
        sh_best = best; sh_pos = pos;
    }
    __syncthreads();
    float best = sh_best; int pos = sh_pos;
    if (pos < 0 || best < threshold * 0.3f) { *out_corr = best; *out_pos = -1; return; }
    const int fs = max(0, pos - 48), fe = min(search_len, pos + 48);
    __syncthreads();
    for (int k = tid; k <= fe - fs; k += blockDim.x) vals[k] = td_corr(s, fs + k, tmpl, CL, tmpl_energy);
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k <= fe - fs; ++k) if (vals[k] > best) { best = vals[k]; pos = fs + k; }
        sh_best = best; sh_pos = pos;
    }
    __syncthreads();
    best = sh_best; pos = sh_pos;
    __syncthreads();
    if (pos > 0 && pos < search_len - 1) {
        if (tid < 2) vals[tid] = td_corr(s, tid == 0 ? pos - 1 : pos + 1, tmpl, CL, tmpl_energy);
        __syncthreads();
        const float c0 = vals[0], c1 = best, c2 = vals[1];
        const float denom = 2.0f * (c0 - 2.0f * c1 + c2);
        if (fabsf(denom) > 1e-10f) {
            float delta = (c0 - c2) / denom;
            delta = fmaxf(-1.0f, fminf(1.0f, delta));
            pos = static_cast<int>(roundf(pos + delta));
        }
        __syncthreads();
    }
    *out_corr = best;
    *out_pos = (best >= threshold) ? pos : -1;
}

// One CTA per window: exact parallel energy prefix (above), normalised peak over all lags for the up chirp, then the
// same for the down chirp on the slice the up-chirp peak selects.  (Round 1 ran the prefix as one serial chain per
// window and needed 14 windows per SM in flight to hide its latency; the parallel prefix needs ~3.5x fewer issue
// slots and no latency hiding.)
constexpr int kPeakThreads = kScanThreads;
__global__ void __launch_bounds__(kPeakThreads, 5)
chirp_peak_kernel(const PeakArgs a) {
    __shared__ float red_v[256];
    __shared__ int red_i[256];
    __shared__ ScanSmem tile;
    const long long f = blockIdx.x;
    const int tid = threadIdx.x;
    const float* s = a.samples + f * a.frame_stride;
    float* c = a.cumsum + f * (static_cast<long long>(a.window) + 1);
    const float* cu = a.corr_up + static_cast<size_t>(f) * a.corr_stride;
    const float* cd = a.corr_dn + static_cast<size_t>(f) * a.corr_stride;
    const int CL = a.chirp_len;
    const int n_in = min(a.window, kN);

    ria_sync_result res;
    res.detected = 0; res.start_sample = -1; res.correlation = 0.0f; res.cfo_hz = 0.0f;
    res.snr_estimate = 0.0f; res.root = -1; res.frame_type = -1; res.aux = -1;

    if (a.window >= 2 * CL + a.gap) {                       // :373-377
        // ---- up chirp over the whole window ----
        prefix_energy(s, 0, n_in, c, tile);
        float up_corr; int up_pos;
        peak_search(cu, 0, c, n_in - CL, CL, a.energy_up, red_v, red_i, &up_corr, &up_pos);
        res.correlation = up_corr;
        if (up_pos >= 0 && !(up_corr < a.threshold)) {      // :708-712
            res.root = up_pos;
            // ---- down chirp search slice (:429-446) ----
            const long long ds = static_cast<long long>(up_pos) + CL / 2;
            const long long expected = static_cast<long long>(up_pos) + CL + a.gap;
            long long de = max(expected + 10000 + CL, ds + 2LL * CL + 1000);
            if (de > a.window) de = a.window;
            if (ds < a.window) {
                if (de <= ds + CL) { de = ds + 2LL * CL; if (de > a.window) de = a.window; }
                const int dlen = static_cast<int>(de - ds);
                if (dlen >= CL) {
                    float dn_corr; int dn_rel;
                    if (dlen >= 2 * CL) {
                        prefix_energy(s, static_cast<int>(ds), min(dlen, kN), c, tile);
                        peak_search(cd, static_cast<int>(ds), c, min(dlen, kN) - CL, CL, a.energy_dn, red_v, red_i, &dn_corr, &dn_rel);
                        if (dn_corr < a.threshold) dn_rel = -1;
                    } else {
                        // slice shorter than two chirps: the reference correlates in the time domain
                        td_detect(s + ds, dlen, a.tmpl_dn_time, CL, a.energy_dn, a.threshold, c, &dn_corr, &dn_rel);
                    }
                    if (dn_rel >= 0) {
                        const int dn_pos = dn_rel + static_cast<int>(ds);
                        res.snr_estimate = dn_corr;
                        res.frame_type = dn_pos;
                        // ---- CFO from the gap error (:462-490) ----
                        const float T = a.duration_ms / 1000.0f;
                        const float rate = (a.f_end - a.f_start) / T;
                        const float c2s = a.sample_rate / rate;
                        const int expected_gap = CL + a.gap;
                        const float gap_error = static_cast<float>(dn_pos - up_pos - expected_gap);
                        const float cfo = gap_error / (2.0f * c2s);
                        res.cfo_hz = cfo;
                        if (!(fabsf(cfo) > 100.0f)) {
                            const float upc = cfo * c2s, dnc = -cfo * c2s;
                            res.start_sample = static_cast<int>(roundf(up_pos + upc));
                            res.aux = static_cast<int>(roundf(dn_pos + dnc));
                            res.detected = 1;
                        }
                    }
                }
            }
        }
    }
    if (tid == 0) a.out[f] = res;
}

}  // namespace

void chirp_tables_free(ChirpTablesDev* t) {
    if (!t) return;
    if (t->tw1) cudaFree(t->tw1);
    if (t->tw2) cudaFree(t->tw2);
    if (t->tmpl_up) cudaFree(t->tmpl_up);
    if (t->tmpl_dn) cudaFree(t->tmpl_dn);
    if (t->tmpl_dn_time) cudaFree(t->tmpl_dn_time);
    if (t->res_tables) cudaFree(t->res_tables);
    if (t->tmpl_up_res) cudaFree(t->tmpl_up_res);
    if (t->tmpl_dn_res) cudaFree(t->tmpl_dn_res);
    if (t->res_scratch) cudaFree(t->res_scratch);
    delete t;
}

static size_t residue_smem_bytes() { return (static_cast<size_t>(kM) + 1024 + 64 + 256) * sizeof(float2); }
static ResidueTables residue_tables(const ChirpTablesDev& t) {
    const float2* tb = t.res_tables;
    return ResidueTables{tb, tb + 512, tb + 1024, tb + 2048, tb + 2112, tb + 3136};
}

static int chirp_tables_dev(ria_ctx* ctx, const ria_chirp_config& cfg, ChirpTablesDev** out) {
    for (ChirpTablesDev* t : ctx->chirp_tables)
        if (std::memcmp(&t->cfg, &cfg, sizeof cfg) == 0) { *out = t; return RIA_OK; }
    // ChirpSync::generateTemplate (:874-900): sin/cos templates and their energies, host floats
    const size_t chirp_len = static_cast<size_t>(cfg.sample_rate * cfg.duration_ms / 1000.0f);
    if (chirp_len == 0 || 2 * chirp_len > kN) return set_error(ctx, RIA_E_UNSUPPORTED, "chirp: unsupported chirp length");
    std::vector<float2> up(kN, make_float2(0.f, 0.f)), dn(kN, make_float2(0.f, 0.f));
    float e_up = 0.f, e_dn = 0.f;
    const float T = cfg.duration_ms / 1000.0f;
    const float k = (cfg.f_end - cfg.f_start) / T;
    for (size_t i = 0; i < chirp_len; ++i) {
        const float t = static_cast<float>(i) / cfg.sample_rate;
        const float pu = 2.0f * M_PI * (cfg.f_start * t + 0.5f * k * t * t);
        const float pd = 2.0f * M_PI * (cfg.f_end * t - 0.5f * k * t * t);
        up[i] = make_float2(std::cos(pu), std::sin(pu));
        dn[i] = make_float2(std::cos(pd), std::sin(pd));
        e_up += up[i].y * up[i].y;
        e_dn += dn[i].y * dn[i].y;
    }
    ChirpTablesDev* t = new ChirpTablesDev();
    t->cfg = cfg; t->chirp_len = static_cast<int>(chirp_len);
    t->gap = static_cast<int>(static_cast<size_t>(cfg.sample_rate * cfg.gap_ms / 1000.0f));
    t->energy_up = e_up; t->energy_dn = e_dn;
    ctx->chirp_tables.push_back(t);
    RIA_CUDA(ctx, cudaMalloc(&t->tw1, sizeof(float2) * kN));
    RIA_CUDA(ctx, cudaMalloc(&t->tw2, sizeof(float2) * kB * kC));
    RIA_CUDA(ctx, cudaMalloc(&t->tmpl_up, sizeof(float2) * kN));
    RIA_CUDA(ctx, cudaMalloc(&t->tmpl_dn, sizeof(float2) * kN));
    RIA_CUDA(ctx, cudaMalloc(&t->tmpl_dn_time, sizeof(float2) * chirp_len));
    cudaStream_t s = ctx->stream;
    chirp_twiddle_kernel<<<(kN + 255) / 256, 256, 0, s>>>(t->tw1, t->tw2);
    RIA_CUDA(ctx, cudaMemcpyAsync(t->tmpl_up, up.data(), sizeof(float2) * kN, cudaMemcpyHostToDevice, s));
    RIA_CUDA(ctx, cudaMemcpyAsync(t->tmpl_dn, dn.data(), sizeof(float2) * kN, cudaMemcpyHostToDevice, s));
    RIA_CUDA(ctx, cudaMemcpyAsync(t->tmpl_dn_time, dn.data(), sizeof(float2) * chirp_len, cudaMemcpyHostToDevice, s));
    RIA_CUDA(ctx, cudaStreamSynchronize(s));
    // residue path: tables, then the conjugated template spectra through the same kernel that transforms the windows
    {
        constexpr size_t kTab = 512 + 512 + 1024 + 64 + 1024 + 1024;
        RIA_CUDA(ctx, cudaMalloc(&t->res_tables, kTab * sizeof(float2)));
        RIA_CUDA(ctx, cudaMalloc(&t->tmpl_up_res, sizeof(float2) * kN));
        RIA_CUDA(ctx, cudaMalloc(&t->tmpl_dn_res, sizeof(float2) * kN));
        t->res_grid = ctx->sm_count > 0 ? ctx->sm_count : 148;
        RIA_CUDA(ctx, cudaMalloc(&t->res_scratch, sizeof(float2) * kM * static_cast<size_t>(t->res_grid)));
        float2* tb = t->res_tables;
        residue_tables_kernel<<<4, 256, 0, s>>>(tb, tb + 512, tb + 1024, tb + 2048, tb + 2112, tb + 3136);
        const size_t smem = residue_smem_bytes();
        RIA_CUDA(ctx, cudaFuncSetAttribute(chirp_residue_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        RIA_CUDA(ctx, cudaFuncSetAttribute(chirp_residue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        for (int which = 0; which < 2; ++which) {
            ResidueArgs ra{};
            ra.cplx_in = which ? t->tmpl_dn : t->tmpl_up;              // still the time-domain (cos, sin) templates here
            ra.out_up = which ? t->tmpl_dn_res : t->tmpl_up_res;
            ra.scratch = t->res_scratch;
            ra.t = residue_tables(*t);
            ra.n_items = 8;
            chirp_residue_kernel<1><<<8, kResThreads, smem, s>>>(ra);
        }
        RIA_CUDA(ctx, cudaGetLastError());
        ctx->launches += 3;
    }
    // template spectra, conjugated (:586-611), kept in the permuted layout of the stage transform
    fft_forward(t->tmpl_up, 1, *t, s);
    fft_forward(t->tmpl_dn, 1, *t, s);
    conj_kernel<<<(kN + 255) / 256, 256, 0, s>>>(t->tmpl_up);
    conj_kernel<<<(kN + 255) / 256, 256, 0, s>>>(t->tmpl_dn);
    RIA_CUDA(ctx, cudaGetLastError());
    RIA_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->launches += 9;
    *out = t;
    return RIA_OK;
}

}  // namespace ria

extern "C" int ria_chirp_config_default(ria_chirp_config* cfg) {
    if (!cfg) return RIA_E_INVAL;
    *cfg = ria_chirp_config{48000.0f, 300.0f, 2700.0f, 500.0f, 100.0f};     // ChirpConfig defaults (:30-39)
    return RIA_OK;
}

extern "C" int ria_chirp_detect_dual_batch_dev(ria_ctx* ctx, const ria_chirp_config* cfg,
                                               const float* samples_dev, int64_t frame_stride, int32_t window,
                                               float threshold, int64_t n_frames, ria_sync_result* out_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || window < 0 || frame_stride < window) return set_error(ctx, RIA_E_INVAL, "chirp: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "chirp: null buffer");
    if (n_frames > 65535) return set_error(ctx, RIA_E_INVAL, "chirp: at most 65535 windows per call");
    if (window > 131072) return set_error(ctx, RIA_E_UNSUPPORTED, "chirp: window must be <= 131072 samples");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    ChirpTablesDev* t = nullptr;
    int rc = chirp_tables_dev(ctx, *cfg, &t);
    if (rc != RIA_OK) return rc;
    // The three transforms make 7 passes over a 1 MB spectrum per window (~18 MB of HBM traffic per
    // window; measured 3.6 ms per 1024 windows = 5.1 TB/s, 77 % of the copy bandwidth).  Walking the batch
    // in sub-batches whose working set stays in the 126 MB L2 was measured and is SLOWER (16 windows per
    // sub-batch: 7.1 ms per 1024; 64: 4.6 ms): the stage grids become too small to fill the machine and
    // the serial prefix chain of the peak kernel is no longer hidden behind other windows.  The knob
    // stays for experiments (profiles/chirp_subbatch.py); the default is the whole call.
    static const int sub_env = [] { const char* e = std::getenv("RIA_CHIRP_SUBBATCH"); return e ? std::atoi(e) : 0; }();
    int sub = sub_env > 0 ? sub_env : static_cast<int>(n_frames);
    if (sub > n_frames) sub = static_cast<int>(n_frames);
    const size_t s_c = static_cast<size_t>(sub) * kN * sizeof(float2);
    const size_t s_cum = ((static_cast<size_t>(sub) * (static_cast<size_t>(window) + 1) * sizeof(float)) + 255) & ~size_t(255);
    rc = ensure_scratch(ctx, 3 * s_c + s_cum + 256);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    float2* d_sig = reinterpret_cast<float2*>(base);
    float2* d_pu = reinterpret_cast<float2*>(base + s_c);
    float2* d_pd = reinterpret_cast<float2*>(base + 2 * s_c);
    float* d_cum = reinterpret_cast<float*>(base + 3 * s_c);
    cudaStream_t s = ctx->stream;
    const int n_in = window < kN ? window : kN;
    time_begin(ctx, KK_CHIRP_SYNC);
    for (int64_t off = 0; off < n_frames; off += sub) {
        const int batch = static_cast<int>(n_frames - off < sub ? n_frames - off : sub);
        const float* in = samples_dev + off * frame_stride;
        static const bool unfused = [] { const char* e = std::getenv("RIA_CHIRP_UNFUSED"); return e && e[0] == '1'; }();
        static const bool slab_env = [] { const char* e = std::getenv("RIA_CHIRP_SLAB"); return e && e[0] == '1'; }();
        static bool slab_attr = false;
        if (slab_env && !slab_attr) {
            RIA_CUDA(ctx, cudaFuncSetAttribute(fft_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(2 * kSlabElems * sizeof(float2))));
            slab_attr = true;
        }
        const bool use_slab = slab_env;
        // |R_up| / |R_dn| go into the signal buffer, which is free once stage 3 has consumed it
        float* mag_up = reinterpret_cast<float*>(d_sig);
        float* mag_dn = mag_up + kN;
        const long long mag_stride = 2LL * kN;
        static const bool staged_env = [] { const char* e = std::getenv("RIA_CHIRP_STAGED"); return e && e[0] == '1'; }();
        if (!staged_env && !unfused && !use_slab) {
            // default: residue-decomposed transforms out of shared memory, then the radix-8 combine + magnitudes
            ResidueArgs ra{};
            ra.samples = in; ra.stride = frame_stride; ra.n_in = n_in;
            ra.tmpl_up = t->tmpl_up_res; ra.tmpl_dn = t->tmpl_dn_res;
            ra.out_up = d_pu; ra.out_dn = d_pd; ra.scratch = t->res_scratch;
            ra.t = residue_tables(*t);
            ra.n_items = batch * 8;
            ra.scale = 1.0f / kN;
            const int grid = ra.n_items < t->res_grid ? ra.n_items : t->res_grid;
            chirp_residue_kernel<0><<<grid, kResThreads, residue_smem_bytes(), s>>>(ra);
            CombineArgs ca{};
            ca.v_up = d_pu; ca.v_dn = d_pd; ca.mag_up = mag_up; ca.mag_dn = mag_dn; ca.mag_stride = mag_stride; ca.n_lags = n_in;
            chirp_combine_kernel<<<dim3(kM / 256, batch), 256, 0, s>>>(ca);
        } else if (unfused) {
            fft_forward_real_to_products(in, frame_stride, n_in, d_sig, d_pu, d_pd, batch, *t, s);
            fft_inverse(d_pu, batch, *t, s, mag_up, mag_stride);
            fft_inverse(d_pd, batch, *t, s, mag_dn, mag_stride);
        } else {
            chirp_correlate(in, frame_stride, n_in, d_sig, d_pu, d_pd, mag_up, mag_dn, mag_stride, batch, *t, s, use_slab);
        }
        PeakArgs a{};
        a.samples = in; a.frame_stride = frame_stride; a.window = window;
        a.corr_up = mag_up; a.corr_dn = mag_dn; a.corr_stride = mag_stride; a.cumsum = d_cum; a.tmpl_dn_time = t->tmpl_dn_time;
        a.threshold = threshold; a.energy_up = t->energy_up; a.energy_dn = t->energy_dn;
        a.chirp_len = t->chirp_len; a.gap = t->gap;
        a.sample_rate = cfg->sample_rate; a.f_start = cfg->f_start; a.f_end = cfg->f_end; a.duration_ms = cfg->duration_ms;
        a.out = out_dev + off;
        chirp_peak_kernel<<<batch, kPeakThreads, 0, s>>>(a);
        ctx->launches += (!staged_env && !unfused && !use_slab) ? 3 : unfused ? 10 : (use_slab ? 5 : 8);
    }
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    return RIA_OK;
}
