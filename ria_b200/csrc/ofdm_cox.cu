// OFDM_COX acquisition: batched OFDMDemodulator::searchForSync
// (src/ofdm/demodulator.cpp:1450-1542), which is what OFDMNvisWaveform::detectSync runs
// (src/waveform/ofdm_cox_waveform.cpp:121-153).  After it the waveform calls processPresynced at the returned
// LTS position with the returned CFO: that half is the presynced demodulator this library already has.
//
// The reference walks the window in steps of 64 samples:
//   hasMinimumEnergy(i, 2 symbols)                 ofdm_sync.cpp:20-50   (decimated energy against a tracked noise floor;
//                                                                         a quiet position skips one symbol ahead)
//   measureCorrelation(i) > threshold              ofdm_sync.cpp:118-163 (Schmidl-Cox metric of the two halves of one FFT
//                                                                         window: DC removal, analytic signal through a
//                                                                         1024-point FFT pair, |P| / sqrt(R1 R2))
//   plateau: measureCorrelation(i + 0, 8, .. 300)  demodulator.cpp:1494-1509 (>= 15 values >= 0.90, peak position)
//   refineLTSTiming(peak)                          ofdm_sync.cpp:386-484 (passband LTS template correlation at 3921 offsets,
//                                                                         earlier-LTS preference, 0.05 floor)
//   estimateCoarseCFO(peak)                        ofdm_sync.cpp:230-261
// and stops at the first position that passes everything.
//
// Which positions are visited depends only on the energies (the skip rule) and everything a candidate goes through
// is independent of earlier failures, so the batch runs as two launches without host synchronisation:
//   cox_scan_kernel    one CTA per window: decimated energies at every multiple of 32, then the noise-floor walk
//                      (one thread, it is a recurrence) -> list of visited positions + the noise floor at each
//   cox_decide_kernel  one CTA per window: the Schmidl-Cox metric at the visited positions, eight at a time (one
//                      per warp), until one is above threshold -> plateau metrics (one warp each) -> LTS search
//                      (all threads) -> CFO; on failure the walk goes on behind the candidate
// Every sum is accumulated in the reference's order with un-fused fp32 operations and the FFT is the reference's
// radix-2 (src/dsp/fft.cpp:96-128), so positions and CFO are the reference's bits.

#include "ofdm_tables.h"
#include "rn_math.h"

#include <algorithm>
#include <cmath>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

struct CoxTablesDev {
    ria_modem_config cfg{};
    float2* lts_iq = nullptr;       // passband LTS template (I, Q), cp + N samples (demodulator.cpp:108-141)
    int tmpl_len = 0;
    float energy_ref = 0.0f;        // ofdm_sync.cpp:405-411
};

void cox_tables_free(CoxTablesDev* t) {
    if (!t) return;
    if (t->lts_iq) cudaFree(t->lts_iq);
    delete t;
}

namespace {

constexpr int kN = 1024;                 // FFT size (ofdm_config_error admits nothing else)
constexpr int kQuickStep = 64;           // QUICK_SEARCH_STEP, demodulator.cpp:1484
constexpr int kEnergyStep = 16;          // SAMPLE_STEP_ENERGY_CHECK, demodulator_constants.hpp:139
constexpr int kPlateauWindow = 300;      // PLATEAU_SEARCH_WINDOW, demodulator_constants.hpp:56
constexpr float kPlateauThreshold = 0.90f;
constexpr int kMinPlateau = 15;
constexpr int kMinSearch = 4000;         // MIN_SEARCH_SAMPLES
constexpr int kMaxWindow = 65536;
constexpr int kMaxEnergy = kMaxWindow / 32;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float2 cmul_ref(float2 a, float2 b) {       // std::complex<float> operator* (finite operands)
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// Blackwell packed fp32 multiply (two IEEE round-to-nearest products per issue slot); the products feed scalar adds,
// never a packed add (ptxas would contract that pair into FFMA2 even under --fmad=false)
__device__ __forceinline__ float2 mul2s(float s, float2 b) {            // (s * b.x, s * b.y)
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float cnorm(float2 a) { return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }
__device__ __forceinline__ unsigned brev10(unsigned i) { return __brev(i) >> 22; }

// per-warp scratch: the transform tile and the staged real samples
struct WarpTile {
    float2 x[kN];
    float s[kN];
};

// Tile element i lives at xi(i): the four low bits are flipped by bit 4 and XORed with bits 9..6.  In the first four
// stages (partners 1, 2, 4, 8 apart: a half-warp touches 16 elements spread over one block of 32) the two halves of the
// spread land in complementary bank pairs; from stage 5 on a half-warp's 16 consecutive elements are only permuted;
// and in the bit-reversed sweeps (a half-warp's lanes differ in bits 9..6 only) the 16 elements land in 16 different
// bank pairs -- every 64-bit access of the transform is conflict-free.
__device__ __forceinline__ int xi(int i) { return i ^ (((i >> 4) & 1) * 15) ^ ((i >> 6) & 15); }

// Stage-major copy of the reference's twiddle table: stage `half` (butterfly span) holds W[k * (N / 2 / half)], k < half,
// at offset half - 1.  In natural order the lanes of a stage read the table with a stride of N / (2 half) entries, all
// in one bank for the middle stages (16- and 32-way conflicts: ncu showed 57 % of the kernel's wavefronts were replays).
__device__ __forceinline__ void stage_twiddles(float2* tws, const float2* __restrict__ tw_nat, int tid, int n_threads) {
    for (int i = tid; i < kN - 1; i += n_threads) {
        const int half = 1 << (31 - __clz(i + 1));
        tws[i] = tw_nat[(i + 1 - half) * (kN / 2 / half)];
    }
}

// The reference's butterflies in stage order on a bit-reversed tile (fft.cpp:107-120); one warp, 16 butterflies per lane
// and stage.
__device__ void warp_fft_stages(float2* x, const float2* tws, bool inverse, int lane) {
    for (int half = 1; half < kN; half <<= 1) {
        // four butterflies in flight per lane: the loads of all four are issued before the first store (the compiler
        // cannot move a load of x[] above a store to x[] by itself; the butterflies of one stage touch disjoint elements)
        for (int b0 = lane; b0 < (kN >> 1); b0 += 128) {
            int i0[4], i1[4];
            float2 u[4], v[4], w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = b0 + 32 * q;
                const int k = b & (half - 1);
                i0[q] = xi(((b - k) << 1) + k);
                i1[q] = xi(((b - k) << 1) + k + half);
                w[q] = tws[half - 1 + k];
                u[q] = x[i0[q]];
                v[q] = x[i1[q]];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (inverse) w[q].y = -w[q].y;
                const float2 t = cmul_ref(w[q], v[q]);
                x[i1[q]] = make_float2(__fsub_rn(u[q].x, t.x), __fsub_rn(u[q].y, t.y));
                x[i0[q]] = make_float2(__fadd_rn(u[q].x, t.x), __fadd_rn(u[q].y, t.y));
            }
        }
        __syncwarp();
    }
}

// measureSchmidlCoxCorrelation (ofdm_sync.cpp:118-163) / the first half of estimateCoarseCFO (:230-246) for the FFT
// window that starts at src: P = sum conj(a[i]) a[i + N/2], R1, R2 over the analytic signal (toAnalytic, :56-84).
// All lanes return the sums.
__device__ void warp_schmidl_cox(const float* __restrict__ src, bool remove_dc, WarpTile& t, const float2* tw, int lane,
                                 float2& P, float& R1, float& R2) {
    for (int i = lane; i < kN; i += 32) t.s[i] = src[i];
    __syncwarp();
    float dc = 0.0f;
    if (remove_dc) {
        if (lane == 0) {                                                // dc_sum is a sequential fp32 sum (:131-135)
            float acc = 0.0f;
            const float4* s4 = reinterpret_cast<const float4*>(t.s);
#pragma unroll 4
            for (int i = 0; i < kN / 4; ++i) {
                const float4 v = s4[i];
                acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
            }
            dc = __fdiv_rn(acc, static_cast<float>(kN));
        }
        dc = __shfl_sync(kFull, dc, 0);
    }
    for (int i = lane; i < kN; i += 32)
        t.x[xi(brev10(i))] = make_float2(remove_dc ? __fsub_rn(t.s[i], dc) : t.s[i], 0.0f);
    __syncwarp();
    warp_fft_stages(t.x, tw, false, lane);
    // positive frequencies x 2, negative ones dropped (:72-77), and the bit reversal of the inverse transform in one sweep
    for (int i = lane; i < kN; i += 32) {
        const int j = static_cast<int>(brev10(i));
        if (i > j) continue;
        float2 a = t.x[xi(i)], b = t.x[xi(j)];
        if (i >= 1 && i < kN / 2) a = make_float2(__fmul_rn(a.x, 2.0f), __fmul_rn(a.y, 2.0f));
        else if (i > kN / 2) a = make_float2(0.0f, 0.0f);
        if (j >= 1 && j < kN / 2) b = make_float2(__fmul_rn(b.x, 2.0f), __fmul_rn(b.y, 2.0f));
        else if (j > kN / 2) b = make_float2(0.0f, 0.0f);
        t.x[xi(i)] = b;
        t.x[xi(j)] = a;
    }
    __syncwarp();
    warp_fft_stages(t.x, tw, true, lane);
    const float scale = 1.0f / static_cast<float>(kN);
    for (int i = lane; i < kN; i += 32) {
        const float2 v = t.x[i];                                        // every element once: the placement does not matter
        t.x[i] = make_float2(__fmul_rn(v.x, scale), __fmul_rn(v.y, scale));
    }
    __syncwarp();
    // four sequential chains of N/2 terms (:149-153).  The terms are independent of the running sums: every lane
    // computes 16 of each kind into two term planes (the sample staging area and the upper half of the transform tile,
    // both dead by now), then lanes 0..3 walk one chain each with 128-bit loads.
    float* term_a = t.s;                                                // [0, 512): Re conj(a) b   [512, 1024): Im conj(a) b
    float* term_b = reinterpret_cast<float*>(t.x);                      // [0, 512): |a|^2          [512, 1024): |b|^2
    float2 av[kN / 64], bv[kN / 64];
#pragma unroll
    for (int q = 0; q < kN / 64; ++q) { av[q] = t.x[xi(lane + 32 * q)]; bv[q] = t.x[xi(lane + 32 * q + kN / 2)]; }
    __syncwarp();                                                       // every a / b is in registers before the tile is reused
#pragma unroll
    for (int q = 0; q < kN / 64; ++q) {
        const int i = lane + 32 * q;
        const float2 a = av[q], b = bv[q];
        term_a[i] = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
        term_a[i + kN / 2] = __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
        term_b[i] = cnorm(a);
        term_b[i + kN / 2] = cnorm(b);
    }
    __syncwarp();
    float acc = 0.0f;
    if (lane < 4) {
        const float4* src = reinterpret_cast<const float4*>((lane < 2 ? term_a : term_b) + (lane & 1) * (kN / 2));
#pragma unroll 4
        for (int i = 0; i < kN / 8; ++i) {
            const float4 v = src[i];
            acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
        }
    }
    P.x = __shfl_sync(kFull, acc, 0);
    P.y = __shfl_sync(kFull, acc, 1);
    R1 = __shfl_sync(kFull, acc, 2);
    R2 = __shfl_sync(kFull, acc, 3);
    __syncwarp();
}

__device__ __forceinline__ float cabs_hypotf(float2 a) {       // std::abs(complex<float>) = glibc hypotf (double inside)
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}

__device__ float warp_correlation(const float* __restrict__ win, int offset, int cp, WarpTile& t, const float2* tw, int lane) {
    float2 P; float R1, R2;
    warp_schmidl_cox(win + offset + cp, true, t, tw, lane, P, R1, R2);
    const float normalization = __fsqrt_rn(__fmul_rn(R1, R2));
    if (normalization < 1e-10f) return 0.0f;
    return __fdiv_rn(cabs_hypotf(P), normalization);
}

// ---------------------------------------------------------------------------------------------
// scan: energies + the noise-floor walk
// ---------------------------------------------------------------------------------------------
struct CoxArgs {
    const float* win; long long stride; int L; long long n;
    float* noise_floor;                 // nullable, in/out per window (Impl::noise_floor_energy)
    float threshold;
    int cp, sym, search_end, corr_window, total_len, max_visit;
    int* visit_off; float* visit_nf; int* n_visit;
    const float2* tw; const float2* lts_iq; int tmpl_len; float energy_ref;
    float sample_rate;
    ria_sync_result* out;
};

__global__ void __launch_bounds__(128)
cox_scan_kernel(const CoxArgs a) {
    __shared__ float energy[kMaxEnergy];
    const int tid = threadIdx.x;
    const int n_e = (a.search_end + 31) >> 5;
    const int count = (a.corr_window + kEnergyStep - 1) / kEnergyStep;
    for (long long w = blockIdx.x; w < a.n; w += gridDim.x) {
        const float* win = a.win + w * a.stride;
        for (int e = tid; e < n_e; e += blockDim.x) {
            const float* p = win + (e << 5);
            float sum_sq = 0.0f;
            for (int i = 0; i < a.corr_window; i += kEnergyStep) {
                const float s = p[i];
                sum_sq = __fadd_rn(sum_sq, __fmul_rn(s, s));
            }
            energy[e] = __fdiv_rn(sum_sq, static_cast<float>(count));
        }
        __syncthreads();
        if (tid == 0) {
            float nf = a.noise_floor ? a.noise_floor[w] : 0.0f;
            int* off = a.visit_off + w * a.max_visit;
            float* vnf = a.visit_nf + w * (a.max_visit + 1);
            int v = 0;
            for (int i = 0; i < a.search_end;) {
                const float e = energy[i >> 5];
                if (nf < 1e-20f) nf = __fmul_rn(e, 0.1f);
                if (e < nf) nf = e;
                else if (e < __fmul_rn(nf, 3.0f))
                    nf = __fadd_rn(__fmul_rn(1.0f - 0.01f, nf), __fmul_rn(0.01f, e));
                if (e >= __fmul_rn(nf, 4.0f)) { off[v] = i; vnf[v] = nf; ++v; i += kQuickStep; }
                else i += a.corr_window / 2;                     // skips half the correlation window (:1487-1490)
            }
            vnf[v] = nf;
            a.n_visit[w] = v;
        }
        __syncthreads();
    }
}

constexpr size_t kTwSmem = sizeof(float2) * kN;           // stage-major twiddles (N - 1 used)
constexpr size_t kTileSmem = kTwSmem + sizeof(WarpTile) * kWarps;

// ---------------------------------------------------------------------------------------------
// per window: candidates in order -> plateau -> LTS -> CFO
// ---------------------------------------------------------------------------------------------
constexpr int kSearchBackSyms = 3;       // SEARCH_BACK = 3 symbols, SEARCH_FWD = half a symbol (ofdm_sync.cpp:391-393)

__global__ void __launch_bounds__(kThreads)
cox_decide_kernel(const CoxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    unsigned char* region = smem_raw + kTwSmem;
    WarpTile* tiles = reinterpret_cast<WarpTile*>(region);
    __shared__ float pc[kPlateauWindow / 8 + 2];
    __shared__ float chunk_c[kWarps];
    __shared__ int sh_i[4];
    __shared__ float red_c[kWarps];
    __shared__ int red_o[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_twiddles(tw, a.tw, tid, kThreads);
    __syncthreads();

    const int psym = a.cp + kN;                             // preamble symbol: FFT + CP (demodulator.cpp:1463)
    const int back = kSearchBackSyms * psym, fwd = psym / 2;
    const int n_off = back + fwd + 1;
    // LTS phase aliases the transform tiles: window span, template, correlations
    float* span = reinterpret_cast<float*>(region);
    const int n_off_pad = ((n_off + 3) & ~3) + 8;           // four adjacent offsets per thread + the look-ahead load
    float2* tmpl = reinterpret_cast<float2*>(span + ((n_off_pad + a.tmpl_len + 3) & ~3));
    float* lcorr = reinterpret_cast<float*>(tmpl + a.tmpl_len);

    for (long long w = blockIdx.x; w < a.n; w += gridDim.x) {
        const float* win = a.win + w * a.stride;
        const int nv = a.n_visit[w];
        const int* voff = a.visit_off + w * a.max_visit;
        const float* vnf = a.visit_nf + w * (a.max_visit + 1);
        int cursor = 0;
        bool done = false;
        while (!done) {
            // the metric at the next kWarps visited positions, one per warp; the first one above threshold is the
            // candidate (positions behind a success are never evaluated, like the reference's break)
            {
                const int v = cursor + warp;
                float c = 0.0f;
                if (v < nv) c = warp_correlation(win, voff[v], a.cp, tiles[warp], tw, lane);
                if (lane == 0) chunk_c[warp] = c;
            }
            __syncthreads();
            if (tid == 0) {
                int hit = -1;
                for (int k = 0; k < kWarps && cursor + k < nv; ++k)
                    if (chunk_c[k] > a.threshold) { hit = k; break; }
                sh_i[0] = hit;
            }
            __syncthreads();
            if (sh_i[0] < 0) {
                cursor += kWarps;
                if (cursor >= nv) {                          // nothing (left) above threshold
                    if (tid == 0) {
                        ria_sync_result r{};
                        a.out[w] = r;
                        if (a.noise_floor) a.noise_floor[w] = vnf[nv];
                    }
                    done = true;
                }
                __syncthreads();
                continue;
            }
            const int v = cursor + sh_i[0];
            const float v_corr = chunk_c[sh_i[0]];
            const int i0 = voff[v];
            int nj = 0;
            for (int j = 0; j <= kPlateauWindow && i0 + j + a.total_len < a.L; j += 8) ++nj;
            for (int q = warp; q < nj; q += kWarps) {
                const float c = warp_correlation(win, i0 + 8 * q, a.cp, tiles[warp], tw, lane);
                if (lane == 0) pc[q] = c;
            }
            __syncthreads();
            if (tid == 0) {
                int plateau = 0, peak_pos = i0;
                float peak = v_corr;
                for (int q = 0; q < nj; ++q) {
                    const float c = pc[q];
                    if (c >= kPlateauThreshold) ++plateau;
                    if (c > peak) { peak = c; peak_pos = i0 + 8 * q; }
                }
                sh_i[1] = plateau; sh_i[2] = peak_pos;
            }
            __syncthreads();
            const int peak_pos = sh_i[2];
            bool ok = sh_i[1] >= kMinPlateau;
            long long best_offset = -1;
            if (ok) {
                // ---- refineLTSTiming(peak_pos) ----
                const int coarse = peak_pos + 4 * psym;
                if (coarse < back || coarse + fwd + a.tmpl_len > a.L) {
                    best_offset = coarse;                    // "not enough data": coarse timing is returned (:395-399)
                } else {
                    const int base = coarse - back;
                    const int span_len = n_off + a.tmpl_len - 1;
                    __syncthreads();
                    for (int i = tid; i < n_off_pad + a.tmpl_len; i += kThreads) span[i] = (i < span_len) ? win[base + i] : 0.0f;
                    for (int i = tid; i < a.tmpl_len; i += kThreads) tmpl[i] = a.lts_iq[i];
                    __syncthreads();
                    float my_best = 0.0f; int my_off = 0x7fffffff;
                    // Four ADJACENT offsets per thread: the seven samples they need for four template taps come from
                    // two 128-bit loads, the squares are formed once per sample, the two template products are one
                    // packed multiply.  Each offset keeps its three sums in the reference's order (:420-433).
                    const float4* span4 = reinterpret_cast<const float4*>(span);
                    const float4* tmpl4 = reinterpret_cast<const float4*>(tmpl);
                    for (int t4 = tid; 4 * t4 < n_off; t4 += kThreads) {
                        float cI[4] = {0.f, 0.f, 0.f, 0.f}, cQ[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
                        float4 lo = span4[t4];
#pragma unroll 2
                        for (int i4 = 0; i4 < a.tmpl_len / 4; ++i4) {
                            const float4 hi = span4[t4 + i4 + 1];
                            const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                            float v2[7];
#pragma unroll
                            for (int j = 0; j < 7; ++j) v2[j] = __fmul_rn(v[j], v[j]);
                            const float4 ta = tmpl4[2 * i4], tb = tmpl4[2 * i4 + 1];
                            const float2 tt[4] = {make_float2(ta.x, ta.y), make_float2(ta.z, ta.w), make_float2(tb.x, tb.y), make_float2(tb.z, tb.w)};
#pragma unroll
                            for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 p = mul2s(v[ii + q], tt[ii]);
                                    cI[q] = __fadd_rn(cI[q], p.x);
                                    cQ[q] = __fadd_rn(cQ[q], p.y);
                                    e[q] = __fadd_rn(e[q], v2[ii + q]);
                                }
                            }
                            lo = hi;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int o = 4 * t4 + q;
                            if (o >= n_off) break;
                            const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(cI[q], cI[q]), __fmul_rn(cQ[q], cQ[q])));
                            const float norm = __fsqrt_rn(__fmul_rn(e[q], a.energy_ref));
                            const float c = (norm > 1e-6f) ? __fdiv_rn(mag, norm) : 0.0f;
                            lcorr[o] = c;
                            if (c > my_best) { my_best = c; my_off = o; }   // ascending offsets: first maximum wins
                        }
                    }
                    // arg-max, ties to the earliest offset (strict > in the reference's ascending loop, :436-448)
                    for (int d = 16; d > 0; d >>= 1) {
                        const float oc = __shfl_down_sync(kFull, my_best, d);
                        const int oo = __shfl_down_sync(kFull, my_off, d);
                        if (oc > my_best || (oc == my_best && oo < my_off)) { my_best = oc; my_off = oo; }
                    }
                    if (lane == 0) { red_c[warp] = my_best; red_o[warp] = my_off; }
                    __syncthreads();
                    if (tid == 0) {
                        float bc = 0.0f; int bo = 0x7fffffff;
                        for (int k = 0; k < kWarps; ++k)
                            if (red_c[k] > bc || (red_c[k] == bc && red_o[k] < bo)) { bc = red_c[k]; bo = red_o[k]; }
                        int best = (bc > 0.0f) ? base + bo : coarse;         // best_offset starts at the coarse position
                        // the LTS is sent twice: prefer the earlier one when it is close (:455-467)
                        if (best >= psym) {
                            const int prev = best - psym;
                            if (prev >= base) {
                                const float prev_c = lcorr[prev - base];
                                if (prev_c >= __fmul_rn(bc, 0.92f)) { best = prev; bc = prev_c; }
                            }
                        }
                        sh_i[3] = (bc < 0.05f) ? -1 : best;                  // LTS_CORRELATION_THRESHOLD for fft >= 1024
                    }
                    __syncthreads();
                    best_offset = sh_i[3];
                    __syncthreads();                                         // tiles are scratch again
                }
                ok = best_offset >= 0;
            }
            if (ok) {
                // ---- estimateCoarseCFO(peak_pos) ----
                if (warp == 0) {
                    float cfo = 0.0f;
                    if (peak_pos + a.cp + kN <= a.L) {
                        float2 P; float R1, R2;
                        warp_schmidl_cox(win + peak_pos + a.cp, false, tiles[0], tw, lane, P, R1, R2);
                        const float phase = glibc_atan2f(P.y, P.x);
                        const double den = M_PI * static_cast<double>(kN);
                        cfo = static_cast<float>(static_cast<double>(__fmul_rn(phase, a.sample_rate)) / den);
                        const float max_cfo = static_cast<float>(static_cast<unsigned>(a.sample_rate) / static_cast<unsigned>(kN));
                        cfo = fmaxf(-max_cfo, fminf(max_cfo, cfo));
                    }
                    if (lane == 0) {
                        ria_sync_result r{};
                        r.detected = 1;
                        r.start_sample = static_cast<int32_t>(best_offset);
                        r.correlation = 0.9f;               // what OFDMNvisWaveform::detectSync reports (ofdm_cox_waveform.cpp:139)
                        r.cfo_hz = cfo;
                        r.aux = peak_pos;
                        a.out[w] = r;
                        if (a.noise_floor) a.noise_floor[w] = vnf[v];
                    }
                }
                done = true;
                __syncthreads();
                break;
            }
            cursor = v + 1;
            __syncthreads();
        }
    }
}

// ---- chain glue: what OFDMNvisWaveform::process derives from the search result (ofdm_cox_waveform.cpp:160-190) ----
// cfo[f] = CFO found (0 when nothing was found), phase[f] = -2 pi cfo pos / fs (a double expression rounded to float,
// wrapped to [-pi, pi] in double steps rounded to float), and the frame copied out of the window at the LTS position.
__global__ void cox_prepare_kernel(const ria_sync_result* __restrict__ sync, long long n, int window, int frame_len,
                                   double sample_rate, int* __restrict__ start, float* __restrict__ cfo, float* __restrict__ phase) {
    const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (f >= n) return;
    const ria_sync_result r = sync[f];
    const bool ok = r.detected && r.start_sample >= 0 && r.start_sample + frame_len <= window;
    start[f] = ok ? r.start_sample : -1;
    const float c = ok ? r.cfo_hz : 0.0f;
    cfo[f] = c;
    const double two_pi_f = static_cast<double>(-2.0f) * M_PI;
    float ph = static_cast<float>(__ddiv_rn(__dmul_rn(__dmul_rn(two_pi_f, static_cast<double>(c)),
                                                      static_cast<double>(ok ? r.start_sample : 0)), sample_rate));
    const double two_pi = static_cast<double>(2.0f) * M_PI;
    while (static_cast<double>(ph) > M_PI) ph = static_cast<float>(static_cast<double>(ph) - two_pi);
    while (static_cast<double>(ph) < -M_PI) ph = static_cast<float>(static_cast<double>(ph) + two_pi);
    phase[f] = ph;
}

__global__ void cox_gather_kernel(const float* __restrict__ win, long long stride, const int* __restrict__ start,
                                  int frame_len, float* __restrict__ frames) {
    const long long f = blockIdx.x;
    const int s0 = start[f];
    const float* src = win + f * stride + (s0 >= 0 ? s0 : 0);
    float* dst = frames + f * static_cast<long long>(frame_len);
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < frame_len; i += gridDim.y * blockDim.x)
        dst[i] = (s0 >= 0) ? src[i] : 0.0f;                      // nothing found: a silent frame (decodes to "not valid")
}

// tap: Impl::measureCorrelation at one offset per window (parity tests)
__global__ void __launch_bounds__(kThreads)
cox_corr_tap_kernel(const CoxArgs a, const int* offsets, float* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    WarpTile* tiles = reinterpret_cast<WarpTile*>(smem_raw + kTwSmem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_twiddles(tw, a.tw, tid, kThreads);
    __syncthreads();
    for (long long w = static_cast<long long>(blockIdx.x) * kWarps + warp; w < a.n; w += static_cast<long long>(gridDim.x) * kWarps) {
        const int off = offsets[w];
        float c = 0.0f;
        if (off >= 0 && off + a.cp + kN <= a.L) c = warp_correlation(a.win + w * a.stride, off, a.cp, tiles[warp], tw, lane);
        if (lane == 0) out[w] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// host side: the LTS passband template
// ---------------------------------------------------------------------------------------------
void host_ifft_ref(std::vector<float2>& x, const std::vector<float2>& tw) {
    // FFT::inverse without FFTW (fft.cpp:96-128), restated: bit reversal, butterflies with conj(w), 1/N
    const size_t n = x.size();
    int logn = 0; while ((size_t{1} << logn) < n) ++logn;
    for (size_t i = 0; i < n; ++i) {
        size_t j = 0;
        for (int b = 0; b < logn; ++b) if (i & (size_t{1} << b)) j |= size_t{1} << (logn - 1 - b);
        if (i < j) std::swap(x[i], x[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < half; ++k) {
                const float wr = tw[k * step].x, wi = -tw[k * step].y;
                const float2 d = x[i + k + half];
                const float tr = wr * d.x - wi * d.y;
                const float ti = wr * d.y + wi * d.x;
                const float2 u = x[i + k];
                x[i + k + half] = make_float2(u.x - tr, u.y - ti);
                x[i + k] = make_float2(u.x + tr, u.y + ti);
            }
    }
    const float scale = 1.0f / static_cast<float>(n);
    for (auto& v : x) { v.x *= scale; v.y *= scale; }
}

int cox_tables_dev(ria_ctx* ctx, const ria_modem_config& cfg, const CoxTablesDev** out) {
    for (CoxTablesDev* t : ctx->cox_tables)
        if (std::memcmp(&t->cfg, &cfg, sizeof cfg) == 0) { *out = t; return RIA_OK; }
    OfdmTablesHost h;
    const int cp = ofdm_cyclic_prefix(cfg), N = static_cast<int>(cfg.fft_size);
    ofdm_build_tables(cfg, cp + N, h);
    std::vector<float2> f(N, make_float2(0.0f, 0.0f));
    const OfdmCarrierTable& c = h.car;
    for (int i = 0; i < c.n_data; ++i) f[c.fft_idx[c.data_car[i]]] = c.tx_data[i];
    for (int i = 0; i < c.n_pilot; ++i) f[c.fft_idx[c.pilot_car[i]]] = make_float2(c.pilot_sign[i], 0.0f);
    host_ifft_ref(f, h.twiddle);
    std::vector<float2> iq(cp + N);
    float energy = 0.0f;
    for (int i = 0; i < cp + N; ++i) {
        const float2 b = (i < cp) ? f[N - cp + i] : f[i - cp];
        const float2 o = h.nco[i];
        iq[i] = make_float2(b.x * o.x - b.y * o.y, b.x * o.y + b.y * o.x);   // lts_baseband[i] * osc (:137-140)
    }
    for (int i = 0; i < cp + N; ++i) { energy += iq[i].x * iq[i].x; energy += iq[i].y * iq[i].y; }
    energy *= 0.5f;
    CoxTablesDev* t = new CoxTablesDev();
    t->cfg = cfg; t->tmpl_len = cp + N; t->energy_ref = energy;
    if (cudaMalloc(&t->lts_iq, iq.size() * sizeof(float2)) != cudaSuccess ||
        cudaMemcpy(t->lts_iq, iq.data(), iq.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        cox_tables_free(t);
        return set_error(ctx, RIA_E_CUDA, "ofdm cox: template upload failed");
    }
    ctx->cox_tables.push_back(t);
    *out = t;
    return RIA_OK;
}

size_t align256(size_t v) { return (v + 255) & ~size_t{255}; }

}  // namespace
}  // namespace ria

extern "C" int ria_ofdm_cox_search_sync_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                                  const float* samples_dev, int64_t window_stride, int32_t window,
                                                  float threshold, float* noise_floor_dev, int64_t n_windows,
                                                  ria_sync_result* out_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_windows < 0 || window < 0 || window_stride < window) return set_error(ctx, RIA_E_INVAL, "ofdm cox: bad sizes");
    if (n_windows == 0) return RIA_OK;
    if (!samples_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "ofdm cox: null buffer");
    if (const char* err = ofdm_config_error(*cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: %s", err);
    if (window > kMaxWindow) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm cox: window %d > %d samples", window, kMaxWindow);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int cp = ofdm_cyclic_prefix(*cfg);
    const int psym = static_cast<int>(cfg->fft_size) + cp;
    const int total_len = 6 * psym, corr_window = 2 * psym;
    if (window < kMinSearch || window < total_len + corr_window) {        // demodulator.cpp:1454-1469: nothing to search
        RIA_CUDA(ctx, cudaMemsetAsync(out_dev, 0, n_windows * sizeof(ria_sync_result), ctx->stream));
        return RIA_OK;
    }
    const OfdmTablesDev* ot = nullptr;
    int rc = ofdm_tables_dev(ctx, *cfg, 0, &ot);
    if (rc != RIA_OK) return rc;
    const CoxTablesDev* ct = nullptr;
    if ((rc = cox_tables_dev(ctx, *cfg, &ct)) != RIA_OK) return rc;

    CoxArgs a{};
    a.win = samples_dev; a.stride = window_stride; a.L = window; a.n = n_windows;
    a.noise_floor = noise_floor_dev; a.threshold = threshold;
    a.cp = cp; a.sym = psym; a.search_end = window - total_len - corr_window; a.corr_window = corr_window;
    a.total_len = total_len;
    a.max_visit = a.search_end / kQuickStep + 2;
    const size_t b_off = align256(static_cast<size_t>(n_windows) * a.max_visit * sizeof(int));
    const size_t b_nf = align256(static_cast<size_t>(n_windows) * (a.max_visit + 1) * sizeof(float));
    const size_t b_nv = align256(static_cast<size_t>(n_windows) * sizeof(int));
    if ((rc = ensure_scratch(ctx, b_off + b_nf + b_nv + 256)) != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    a.visit_off = reinterpret_cast<int*>(base);
    a.visit_nf = reinterpret_cast<float*>(base + b_off);
    a.n_visit = reinterpret_cast<int*>(base + b_off + b_nf);
    a.tw = ot->twiddle_nat; a.lts_iq = ct->lts_iq; a.tmpl_len = ct->tmpl_len; a.energy_ref = ct->energy_ref;
    a.sample_rate = static_cast<float>(cfg->sample_rate);
    a.out = out_dev;

    if (!ctx->occ_cache.count(reinterpret_cast<const void*>(cox_decide_kernel))) {
        RIA_CUDA(ctx, cudaFuncSetAttribute(cox_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTileSmem)));
        ctx->occ_cache[reinterpret_cast<const void*>(cox_decide_kernel)] = 2;
    }
    time_begin(ctx, KK_OFDM_SYNC);
    long long g = std::min<long long>(n_windows, static_cast<long long>(ctx->sm_count) * 8);
    cox_scan_kernel<<<static_cast<unsigned>(g), 128, 0, ctx->stream>>>(a);
    g = std::min<long long>(n_windows, static_cast<long long>(ctx->sm_count) * 2);
    cox_decide_kernel<<<static_cast<unsigned>(g), kThreads, kTileSmem, ctx->stream>>>(a);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return RIA_OK;
}

extern "C" int ria_ofdm_cox_correlation_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                                  const float* samples_dev, int64_t window_stride, int32_t window,
                                                  const int32_t* offsets_dev, int64_t n_windows, float* corr_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_windows < 0 || window < 0 || window_stride < window) return set_error(ctx, RIA_E_INVAL, "ofdm cox: bad sizes");
    if (n_windows == 0) return RIA_OK;
    if (!samples_dev || !offsets_dev || !corr_dev) return set_error(ctx, RIA_E_INVAL, "ofdm cox: null buffer");
    if (const char* err = ofdm_config_error(*cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: %s", err);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const OfdmTablesDev* ot = nullptr;
    int rc = ofdm_tables_dev(ctx, *cfg, 0, &ot);
    if (rc != RIA_OK) return rc;
    CoxArgs a{};
    a.win = samples_dev; a.stride = window_stride; a.L = window; a.n = n_windows;
    a.cp = ofdm_cyclic_prefix(*cfg); a.tw = ot->twiddle_nat;
    if (!ctx->occ_cache.count(reinterpret_cast<const void*>(cox_corr_tap_kernel))) {
        RIA_CUDA(ctx, cudaFuncSetAttribute(cox_corr_tap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTileSmem)));
        ctx->occ_cache[reinterpret_cast<const void*>(cox_corr_tap_kernel)] = 2;
    }
    const long long g = std::min<long long>((n_windows + kWarps - 1) / kWarps, static_cast<long long>(ctx->sm_count) * 2);
    cox_corr_tap_kernel<<<static_cast<unsigned>(g), kThreads, kTileSmem, ctx->stream>>>(a, offsets_dev, corr_dev);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}

// OFDMNvisWaveform::detectSync + process + the frame decoder for a batch of windows: search, then the presynced chain
// (ria_ofdm_rx_frames_dev) on the frame that starts at the LTS position found, with the CFO found and the initial mixer
// phase the waveform derives from them.  A window without sync yields a status with nothing valid; sync_dev (optional)
// receives the search results.
extern "C" int ria_ofdm_cox_rx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg, int rate, int use_channel_interleave,
                                          const float* samples_dev, int64_t window_stride, int32_t window, int32_t frame_len,
                                          float threshold, float* noise_floor_dev, int64_t n_windows,
                                          uint8_t* data_dev, ria_frame_status* status_dev, float* snr_db_dev,
                                          ria_sync_result* sync_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_windows < 0 || window < 0 || frame_len <= 0 || window_stride < window) return set_error(ctx, RIA_E_INVAL, "ofdm cox rx: bad sizes");
    if (n_windows == 0) return RIA_OK;
    if (!samples_dev || !data_dev || !status_dev) return set_error(ctx, RIA_E_INVAL, "ofdm cox rx: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    // buffers between the stages live in the chain scratch: sync results, start / cfo / phase, the gathered frames
    const size_t b_sync = align256(static_cast<size_t>(n_windows) * sizeof(ria_sync_result));
    const size_t b_i = align256(static_cast<size_t>(n_windows) * 4);
    const size_t b_fr = align256(static_cast<size_t>(n_windows) * frame_len * sizeof(float));
    const size_t need = b_sync + 3 * b_i + b_fr + 256;
    if (need > ctx->chain_scratch_bytes) {
        if (ctx->chain_scratch) RIA_CUDA(ctx, cudaFree(ctx->chain_scratch));
        ctx->chain_scratch = nullptr; ctx->chain_scratch_bytes = 0;
        RIA_CUDA(ctx, cudaMalloc(&ctx->chain_scratch, need));
        ctx->chain_scratch_bytes = need;
    }
    unsigned char* base = static_cast<unsigned char*>(ctx->chain_scratch);
    ria_sync_result* d_sync = sync_dev ? sync_dev : reinterpret_cast<ria_sync_result*>(base);
    int* d_start = reinterpret_cast<int*>(base + b_sync);
    float* d_cfo = reinterpret_cast<float*>(base + b_sync + b_i);
    float* d_phase = reinterpret_cast<float*>(base + b_sync + 2 * b_i);
    float* d_frames = reinterpret_cast<float*>(base + b_sync + 3 * b_i);
    int rc = ria_ofdm_cox_search_sync_batch_dev(ctx, cfg, samples_dev, window_stride, window, threshold, noise_floor_dev,
                                                n_windows, d_sync);
    if (rc != RIA_OK) return rc;
    cox_prepare_kernel<<<static_cast<unsigned>((n_windows + 255) / 256), 256, 0, ctx->stream>>>(
        d_sync, n_windows, window, frame_len, static_cast<double>(cfg->sample_rate), d_start, d_cfo, d_phase);
    cox_gather_kernel<<<dim3(static_cast<unsigned>(n_windows), 4), 256, 0, ctx->stream>>>(
        samples_dev, window_stride, d_start, frame_len, d_frames);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return ria_ofdm_rx_frames_dev(ctx, cfg, rate, use_channel_interleave, d_frames, frame_len, frame_len, d_cfo, d_phase,
                                  n_windows, data_dev, status_dev, snr_db_dev);
}

extern "C" int ria_ofdm_cox_rx_frames_host(ria_ctx* ctx, const ria_modem_config* cfg, int rate, int use_channel_interleave,
                                           const float* samples, int64_t window_stride, int32_t window, int32_t frame_len,
                                           float threshold, float* noise_floor, int64_t n_windows,
                                           uint8_t* data, ria_frame_status* status, float* snr_db, ria_sync_result* sync) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_windows < 0 || window <= 0 || frame_len <= 0 || window_stride < window) return set_error(ctx, RIA_E_INVAL, "ofdm cox rx host: bad sizes");
    if (n_windows == 0) return RIA_OK;
    if (!samples || !data || !status) return set_error(ctx, RIA_E_INVAL, "ofdm cox rx host: null buffer");
    const LdpcCodeHost* code = nullptr;
    try { code = &ldpc_code_host(rate); } catch (...) { return set_error(ctx, RIA_E_INVAL, "ofdm cox rx host: bad rate %d", rate); }
    const int out_row = 4 * (code->k / 8);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = n_windows < 8192 ? n_windows : 8192;
    const size_t b_in = align256(static_cast<size_t>(chunk) * window * sizeof(float));
    const size_t b_nf = align256(static_cast<size_t>(chunk) * 4);
    const size_t b_data = align256(static_cast<size_t>(chunk) * out_row);
    const size_t b_st = align256(static_cast<size_t>(chunk) * sizeof(ria_frame_status));
    const size_t b_sy = align256(static_cast<size_t>(chunk) * sizeof(ria_sync_result));
    int rc = ensure_stage(ctx, 0, b_in + 2 * b_nf + b_data + b_st + b_sy + 256, 0);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_nf = reinterpret_cast<float*>(base + b_in);
    float* d_snr = reinterpret_cast<float*>(base + b_in + b_nf);
    uint8_t* d_data = base + b_in + 2 * b_nf;
    ria_frame_status* d_st = reinterpret_cast<ria_frame_status*>(base + b_in + 2 * b_nf + b_data);
    ria_sync_result* d_sy = reinterpret_cast<ria_sync_result*>(base + b_in + 2 * b_nf + b_data + b_st);
    cudaStream_t s = ctx->stream;
    for (int64_t off = 0; off < n_windows; off += chunk) {
        const int64_t n = std::min<int64_t>(chunk, n_windows - off);
        RIA_CUDA(ctx, cudaMemcpy2DAsync(d_in, static_cast<size_t>(window) * sizeof(float), samples + off * window_stride,
                                        static_cast<size_t>(window_stride) * sizeof(float), static_cast<size_t>(window) * sizeof(float),
                                        static_cast<size_t>(n), cudaMemcpyHostToDevice, s));
        if (noise_floor) RIA_CUDA(ctx, cudaMemcpyAsync(d_nf, noise_floor + off, n * sizeof(float), cudaMemcpyHostToDevice, s));
        rc = ria_ofdm_cox_rx_frames_dev(ctx, cfg, rate, use_channel_interleave, d_in, window, window, frame_len, threshold,
                                        noise_floor ? d_nf : nullptr, n, d_data, d_st, d_snr, d_sy);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(data + off * out_row, d_data, static_cast<size_t>(n) * out_row, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(status + off, d_st, n * sizeof(ria_frame_status), cudaMemcpyDeviceToHost, s));
        if (snr_db) RIA_CUDA(ctx, cudaMemcpyAsync(snr_db + off, d_snr, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (sync) RIA_CUDA(ctx, cudaMemcpyAsync(sync + off, d_sy, n * sizeof(ria_sync_result), cudaMemcpyDeviceToHost, s));
        if (noise_floor) RIA_CUDA(ctx, cudaMemcpyAsync(noise_floor + off, d_nf, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return RIA_OK;
}
