// Batched Zadoff-Chu preamble detection with CFO estimation for sm_100a.
//
// Replaces sync::ZCSync::detect / correlate / computeCorrelationMag (src/sync/zc_sync.hpp:192-391,
// 485-626, 441-482) as called by MCDPSKWaveform::detectDataSync (src/waveform/mc_dpsk_waveform.cpp:
// 227-292): per root, down-mix by (carrier + known_cfo), correlate against the linearly
// interpolated 127-chip sequence (1016 samples) on a coarse lag grid (step 31) plus +-31 lags
// around the coarse maximum, first strict maximum of |corr|, prefer the earlier repetition when it
// reaches 40 % of the peak, low-SNR two-repetition combining, CFO from the phase between the two
// repetitions, start_sample = position + 2512.
//
// Bit-exact sync indices need bit-exact correlation magnitudes where they are compared, so each
// lag's 1016-term complex sum and energy are accumulated in the reference's order without FMA;
// the ZC template (chips and interpolation) is built on the host with the reference's float
// expressions.  Kernels:
//   zc_baseband_kernel   bb[i] = rx[i] * e^{-j 2 pi (fc + known_cfo) i / fs}   (HBM scratch, once per frame)
//   zc_coarse_tile_kernel  one CTA per tile of 256 coarse lags, baseband span and templates staged in shared
//                        memory, all selected roots per pass: (sum, energy) -> normalised corr + magnitude
//   zc_finish_kernel     one CTA per frame: coarse arg-max, fine lags, peak, rep-1 adjust,
//                        combining, CFO, best root

#include "ria_internal.h"
#include "rn_math.h"

#include <cmath>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

struct ZcTablesDev {
    ria_zc_config cfg{};
    float2* zc_interp = nullptr;     // [4 roots][ref_samples] interpolated template (not conjugated)
    int ref_samples = 0;
};

namespace {

constexpr int kMaxRef = 2048;

__device__ __forceinline__ float cabs_d(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}

// sum_i bb[lag+i] * conj(zc[i]) and sum_i |bb[lag+i]|^2, in order (zc_sync.hpp:536-556)
__device__ __forceinline__ void corr_at(const float2* __restrict__ bb, const float2* zc, int ref_samples, int lag,
                                        float2* sum_out, float* energy_out) {
    float2 sum = make_float2(0.f, 0.f);
    float en = 0.f;
    const float2* p = bb + lag;
#pragma unroll 4
    for (int i = 0; i < ref_samples; ++i) {
        const float2 b = p[i];
        const float2 z = zc[i];
        // b * conj(z)
        const float re = __fsub_rn(__fmul_rn(b.x, z.x), __fmul_rn(b.y, -z.y));
        const float im = __fadd_rn(__fmul_rn(b.x, -z.y), __fmul_rn(b.y, z.x));
        sum.x = __fadd_rn(sum.x, re);
        sum.y = __fadd_rn(sum.y, im);
        en = __fadd_rn(en, __fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y)));
    }
    *sum_out = sum;
    *energy_out = en;
}

__device__ __forceinline__ float norm_mag(float2 sum, float energy, float ref_energy) {
    const float denom = sqrtf(__fmul_rn(energy, ref_energy));
    return (denom > 1e-10f) ? __fdiv_rn(cabs_d(sum), denom) : 0.0f;
}

__global__ void zc_baseband_kernel(const float* __restrict__ samples, long long frame_stride, int window,
                                   const float* __restrict__ known_cfo, float carrier, float sample_rate,
                                   float2* __restrict__ bb, long long bb_stride) {
    const long long f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= window) return;
    const float dc = carrier + (known_cfo ? known_cfo[f] : 0.0f);           // downconvert_freq (:500)
    const float t = static_cast<float>(i) / sample_rate;                     // :502
    const float phase = static_cast<float>(-2.0f * M_PI * static_cast<double>(dc) * static_cast<double>(t));
    // std::cos / std::sin of a float = glibc's cosf / sinf; the mixer phase runs to several thousand radians, where glibc
    // takes its large-argument reduction (restated bit for bit in rn_math.h)
    float s, c;
    glibc_sincosf(phase, &s, &c);
    const float r = samples[f * frame_stride + i];
    bb[f * bb_stride + i] = make_float2(__fmul_rn(c, r), __fmul_rn(s, r));
}

// Blackwell packed fp32: two IEEE round-to-nearest products / sums per issue slot.  ptxas would contract a packed
// multiply that feeds a packed ADD into FFMA2 even under --fmad=false, so products only ever feed SCALAR adds.
__device__ __forceinline__ float2 zc_mul2s(float s, float2 b) {             // (s * b.x, s * b.y)
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 zc_mul2(float2 a, float2 b) {             // (a.x * b.x, a.y * b.y)
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 zc_add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

// Coarse lags (zc_sync.hpp:522-560): one CTA owns a tile of kCoarseLags consecutive coarse lags of one window and
// evaluates them for ALL selected roots at once.  The baseband span the tile touches ((lags - 1) * step + R samples)
// and the templates are staged in shared memory once; each thread then walks its lag's R terms in the reference's
// order, keeping one (sum, energy) accumulator set in registers -- the baseband sample and the energy term are shared
// by the roots.  With the odd step the threads of a half-warp read 16 different bank pairs (conflict-free), the
// template read is a broadcast.  Per term and root: two packed multiplies, two scalar adds, one packed accumulate
//   b * conj(z) = (bx zx + by zy, by zx - bx zy)   with the template stored as (zx, -zy, zy, zx)
// (the same products and the same sums as the reference's complex multiply: a - (-c) = a + c exactly).
constexpr int kCoarseLags = 256;

template <int NR>
__global__ void __launch_bounds__(kCoarseLags)
zc_coarse_tile_kernel(const float2* __restrict__ bb, long long bb_stride, int window, const float2* __restrict__ zc_g,
                      int ref_samples, int step, int n_coarse, const int* __restrict__ root_slot,
                      float2* __restrict__ coarse_corr, float* __restrict__ coarse_mag) {
    extern __shared__ __align__(16) unsigned char zc_smem[];
    float4* zt = reinterpret_cast<float4*>(zc_smem);                         // [NR][R] (zx, -zy, zy, zx)
    float2* tile = reinterpret_cast<float2*>(zt + static_cast<size_t>(NR) * ref_samples);
    const long long f = blockIdx.y;
    const int R = ref_samples;
    const int k0 = blockIdx.x * kCoarseLags;
    const int n_here = min(kCoarseLags, n_coarse - k0);
    const int span = (n_here - 1) * step + R;                                 // <= window - k0 * step by construction
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float2* zsrc = zc_g + static_cast<size_t>(root_slot[r]) * R;
        for (int i = threadIdx.x; i < R; i += kCoarseLags) { const float2 z = zsrc[i]; zt[r * R + i] = make_float4(z.x, -z.y, z.y, z.x); }
    }
    const float2* src = bb + f * bb_stride + static_cast<long long>(k0) * step;
    for (int i = threadIdx.x; i < span; i += kCoarseLags) tile[i] = src[i];
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= n_here) return;
    const float2* p = tile + t * step;
    float2 sum[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) sum[r] = make_float2(0.f, 0.f);
    float en = 0.f;
#pragma unroll 2
    for (int i = 0; i < R; ++i) {
        const float2 b = p[i];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const float4 z = zt[r * R + i];
            const float2 p1 = zc_mul2s(b.x, make_float2(z.x, z.y));          // (bx zx, bx (-zy))
            const float2 p2 = zc_mul2s(b.y, make_float2(z.z, z.w));          // (by zy, by zx)
            const float re = __fadd_rn(p1.x, p2.x), im = __fadd_rn(p1.y, p2.y);
            sum[r] = zc_add2(sum[r], make_float2(re, im));
        }
        const float2 sq = zc_mul2(b, b);
        en = __fadd_rn(en, __fadd_rn(sq.x, sq.y));
    }
    const float ref_energy = static_cast<float>(R);
    const float denom = sqrtf(__fmul_rn(en, ref_energy));
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const size_t o = (static_cast<size_t>(f) * NR + r) * n_coarse + k0 + t;
        coarse_mag[o] = (denom > 1e-10f) ? __fdiv_rn(cabs_d(sum[r]), denom) : 0.0f;
        coarse_corr[o] = (denom > 1e-10f) ? make_float2(__fdiv_rn(sum[r].x, denom), __fdiv_rn(sum[r].y, denom)) : make_float2(0.f, 0.f);
    }
}

struct FinishArgs {
    const float2* bb; long long bb_stride; int window;
    const float2* zc_g; int ref_samples, step, n_coarse, n_roots;
    int roots[4]; int root_slot[4];
    const float2* coarse_corr; const float* coarse_mag;
    float threshold, sample_rate; int preamble_len;
    ria_zc_config cfg;
    ria_sync_result* out;
};

// arg-max with the reference's "first strict maximum" semantics = (largest value, smallest index)
__device__ __forceinline__ void better(float& bv, int& bi, float v, int i) {
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}

__global__ void __launch_bounds__(128)
zc_finish_kernel(const FinishArgs a) {
    extern __shared__ float2 sm_zc[];                       // [ref_samples] template | [3 * ref_samples] baseband span
    __shared__ float red_v[128];
    __shared__ int red_i[128];
    __shared__ float2 fine_corr[160];
    __shared__ float2 lag_sum[4];                           // correlations at peak - R, peak, peak + R
    __shared__ float lag_en[4];
    float2* span = sm_zc + a.ref_samples;
    const long long f = blockIdx.x;
    const int tid = threadIdx.x;
    const float2* bb = a.bb + f * a.bb_stride;
    const int R = a.ref_samples;
    const int corr_len = a.window - R + 1;
    const float ref_energy = static_cast<float>(R);

    float best_corr = 0.0f, best_cfo = 0.0f;
    int best_root = -1, best_pos = -1;

    for (int r = 0; r < a.n_roots; ++r) {
        __syncthreads();
        const float2* zsrc = a.zc_g + static_cast<size_t>(a.root_slot[r]) * R;
        for (int i = tid; i < R; i += 128) sm_zc[i] = zsrc[i];
        const float* cmag = a.coarse_mag + (static_cast<size_t>(f) * a.n_roots + r) * a.n_coarse;
        const float2* ccorr = a.coarse_corr + (static_cast<size_t>(f) * a.n_roots + r) * a.n_coarse;
        // ---- coarse arg-max (:533-560) ----
        float bv = 0.0f; int bi = 0;
        for (int k = tid; k < a.n_coarse; k += 128) better(bv, bi, cmag[k], k * a.step);
        red_v[tid] = bv; red_i[tid] = bi;
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) {
            if (tid < s) { float v = red_v[tid]; int i = red_i[tid]; better(v, i, red_v[tid + s], red_i[tid + s]); red_v[tid] = v; red_i[tid] = i; }
            __syncthreads();
        }
        const int coarse_best = red_i[0];
        const int fine_start = max(0, coarse_best - a.step);
        const int fine_end = min(corr_len, coarse_best + a.step + 1);
        __syncthreads();
        // ---- fine lags (:568-596) ----
        // The 1016-term sums are sequential per lag; read from global memory they were bound by load latency
        // (one thread, four loads in flight).  The span the lags touch is staged in shared memory first.
        const int n_fine = fine_end - fine_start;
        for (int i = tid; i < n_fine - 1 + R; i += 128) span[i] = bb[fine_start + i];
        __syncthreads();
        if (tid < n_fine) {
            float2 sum; float en;
            corr_at(span, sm_zc, R, tid, &sum, &en);
            const float denom = sqrtf(__fmul_rn(en, ref_energy));
            fine_corr[tid] = (denom > 1e-10f) ? make_float2(__fdiv_rn(sum.x, denom), __fdiv_rn(sum.y, denom)) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        // ---- peak over the sparse correlation array (:237-248) ----
        bv = 0.0f; bi = 0;
        for (int k = tid; k < a.n_coarse; k += 128) {
            const int lag = k * a.step;
            if (lag < fine_start || lag >= fine_end) better(bv, bi, cabs_d(ccorr[k]), lag);
        }
        if (tid < n_fine) better(bv, bi, cabs_d(fine_corr[tid]), fine_start + tid);
        red_v[tid] = bv; red_i[tid] = bi;
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) {
            if (tid < s) { float v = red_v[tid]; int i = red_i[tid]; better(v, i, red_v[tid + s], red_i[tid + s]); red_v[tid] = v; red_i[tid] = i; }
            __syncthreads();
        }
        const float peak_mag = red_v[0];
        const int peak_pos = red_i[0];
        // ---- rep-1 check (:255-277) and the two repetitions (:279-370) ----
        // The reference correlates at peak - R (rep-1 check), then at timing_pos and timing_pos + R with timing_pos one of
        // {peak, peak - R}: all of them are among the lags peak - R, peak, peak + R, which three threads evaluate at once
        // from one staged span (a correlation nobody asks for has no side effect).
        const int lo = max(0, peak_pos - R), hi = min(a.window, peak_pos + 2 * R);
        for (int i = tid; i < hi - lo; i += 128) span[i] = bb[lo + i];
        __syncthreads();
        if (tid < 3) {
            const int lag = peak_pos + (tid - 1) * R;
            if (lag >= 0 && lag + R <= a.window) corr_at(span, sm_zc, R, lag - lo, &lag_sum[tid], &lag_en[tid]);
        }
        __syncthreads();
        const bool check_earlier = (peak_mag > a.threshold) && (peak_pos >= R);
        int timing_pos = peak_pos;
        if (check_earlier) {
            const float earlier_mag = norm_mag(lag_sum[0], lag_en[0], ref_energy);        // lag >= 0 and fits by construction
            if (earlier_mag > peak_mag * 0.4f) timing_pos = peak_pos - R;
        }
        const int rep2_pos = timing_pos + R;
        const bool rep2_fits = rep2_pos + R <= a.window;
        const int t1 = (timing_pos == peak_pos) ? 1 : 0;                                  // slot of timing_pos; rep2 is the next one
        const float2 rep_sum1 = lag_sum[t1], rep_sum2 = lag_sum[t1 + 1];
        const float rep_en1 = lag_en[t1], rep_en2 = lag_en[t1 + 1];
        // ---- low-SNR combining (:279-299) ----
        float combined = peak_mag;
        if (peak_mag > 0.0f && peak_mag < 0.25f && rep2_fits) {
            const float r1 = norm_mag(rep_sum1, rep_en1, ref_energy);
            const float r2 = norm_mag(rep_sum2, rep_en2, ref_energy);
            combined = sqrtf(r1 * r1 + r2 * r2) / sqrtf(2.0f);
            combined = (combined < peak_mag) ? peak_mag : combined;                      // std::max(combined, peak)
        }
        // ---- best root + CFO (:306-370) ----
        if (combined > best_corr) {
            best_corr = combined; best_root = a.roots[r]; best_pos = timing_pos;
            if (rep2_fits) {
                const float c1 = cabs_d(rep_sum1) / R, c2 = cabs_d(rep_sum2) / R;
                if (c1 > 0.1f && c2 > 0.1f) {
                    const float2 x = rep_sum2, y = rep_sum1;                              // corr2 * conj(corr1)
                    const float re = __fsub_rn(__fmul_rn(x.x, y.x), __fmul_rn(x.y, -y.y));
                    const float im = __fadd_rn(__fmul_rn(x.x, -y.y), __fmul_rn(x.y, y.x));
                    const float phase_diff = glibc_atan2f(im, re);                        // std::arg (:351) = glibc atan2f, restated
                    const float rep_duration = static_cast<float>(R) / a.sample_rate;
                    best_cfo = static_cast<float>(static_cast<double>(phase_diff) / (2.0f * M_PI * static_cast<double>(rep_duration)));
                }
            }
        }
    }
    if (tid == 0) {
        ria_sync_result res;
        res.detected = 0; res.start_sample = -1; res.correlation = best_corr; res.cfo_hz = 0.0f;
        res.snr_estimate = 0.0f; res.root = best_root; res.frame_type = 255; res.aux = 0;
        if (best_root >= 0) {
            // ZCConfig::getTypeForRoot (:101-107)
            if (best_root == a.cfg.root_ping) res.frame_type = 0;
            else if (best_root == a.cfg.root_pong) res.frame_type = 1;
            else if (best_root == a.cfg.root_data) res.frame_type = 2;
            else if (best_root == a.cfg.root_control) res.frame_type = 3;
        }
        if (best_corr > a.threshold && best_root >= 0) {
            res.detected = 1;
            res.cfo_hz = best_cfo;
            res.start_sample = best_pos + a.preamble_len;
            // correlationToSNR (:628-633)
            float snr;
            if (best_corr <= 0.01f) snr = -10.0f;
            else if (best_corr >= 0.99f) snr = 30.0f;
            else { snr = 20.0f * log10f(best_corr / (1.0f - best_corr + 0.01f)); snr = fminf(fmaxf(snr, -10.0f), 30.0f); }
            res.snr_estimate = snr;
        }
        a.out[f] = res;
    }
}

void build_zc(const ria_zc_config& cfg, std::vector<float2>& table, int& ref_samples) {
    const int length = cfg.sequence_length, up = cfg.upsample_factor;
    ref_samples = length * up;
    const int roots[4] = {cfg.root_ping, cfg.root_pong, cfg.root_data, cfg.root_control};
    table.assign(static_cast<size_t>(4) * ref_samples, make_float2(0.f, 0.f));
    for (int r = 0; r < 4; ++r) {
        std::vector<float2> zc(length);
        const bool even = (length % 2 == 0);
        for (int n = 0; n < length; ++n) {                       // generateZC (:420-436)
            float phase;
            if (even) phase = -M_PI * roots[r] * n * n / length;
            else phase = -M_PI * roots[r] * n * (n + 1) / length;
            zc[n] = make_float2(std::cos(phase), std::sin(phase));
        }
        for (int i = 0; i < ref_samples; ++i) {                  // interpolation (:536-548)
            const float chip_pos = static_cast<float>(i) / up;
            const int idx = static_cast<int>(chip_pos);
            const float frac = chip_pos - idx;
            float2 v;
            if (idx < length - 1) {
                const float a = 1.0f - frac;
                v = make_float2(zc[idx].x * a + zc[idx + 1].x * frac, zc[idx].y * a + zc[idx + 1].y * frac);
            } else v = zc[idx];
            table[static_cast<size_t>(r) * ref_samples + i] = v;
        }
    }
}

}  // namespace

void zc_tables_free(ZcTablesDev* t) {
    if (!t) return;
    if (t->zc_interp) cudaFree(t->zc_interp);
    delete t;
}

}  // namespace ria

extern "C" int ria_zc_config_default(ria_zc_config* cfg) {
    if (!cfg) return RIA_E_INVAL;
    // MCDPSKWaveform::initZCSync (src/waveform/mc_dpsk_waveform.cpp:50-64)
    *cfg = ria_zc_config{48000.0f, 127, 8, 2, 1500.0f, 10.0f, 1, 3, 5, 7};
    return RIA_OK;
}

extern "C" int ria_zc_detect_batch_dev(ria_ctx* ctx, const ria_zc_config* cfg,
                                       const float* samples_dev, int64_t frame_stride, int32_t window,
                                       const float* known_cfo_dev, float threshold, uint32_t root_mask,
                                       int64_t n_frames, ria_sync_result* out_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || window < 0 || frame_stride < window) return set_error(ctx, RIA_E_INVAL, "zc: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "zc: null buffer");
    if (cfg->sequence_length < 2 || cfg->upsample_factor < 1 || cfg->sequence_length * cfg->upsample_factor > kMaxRef ||
        cfg->num_repetitions < 1 || !(cfg->sample_rate > 0))
        return set_error(ctx, RIA_E_UNSUPPORTED, "zc: unsupported configuration");
    if (n_frames > 65535) return set_error(ctx, RIA_E_INVAL, "zc: at most 65535 windows per call");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    ZcTablesDev* t = nullptr;
    for (ZcTablesDev* x : ctx->zc_tables) if (std::memcmp(&x->cfg, cfg, sizeof *cfg) == 0) { t = x; break; }
    if (!t) {
        std::vector<float2> table; int ref = 0;
        build_zc(*cfg, table, ref);
        t = new ZcTablesDev(); t->cfg = *cfg; t->ref_samples = ref;
        ctx->zc_tables.push_back(t);
        RIA_CUDA(ctx, cudaMalloc(&t->zc_interp, table.size() * sizeof(float2)));
        RIA_CUDA(ctx, cudaMemcpy(t->zc_interp, table.data(), table.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    const int R = t->ref_samples;
    const int preamble_len = R * cfg->num_repetitions + static_cast<int>(cfg->sample_rate * cfg->gap_ms / 1000.0f);

    FinishArgs fa{};
    int n_roots = 0;
    const int cand[4] = {cfg->root_ping, cfg->root_pong, cfg->root_data, cfg->root_control};
    for (int i = 0; i < 4; ++i) if (root_mask & (1u << i)) { fa.roots[n_roots] = cand[i]; fa.root_slot[n_roots] = i; ++n_roots; }

    if (window < R || n_roots == 0) {
        // detect() returns a default result when the window is shorter than one repetition (:200-203)
        std::vector<ria_sync_result> def(static_cast<size_t>(n_frames));
        for (auto& r : def) { r = ria_sync_result{}; r.start_sample = -1; r.root = -1; r.frame_type = 255; }
        RIA_CUDA(ctx, cudaMemcpyAsync(out_dev, def.data(), def.size() * sizeof(ria_sync_result), cudaMemcpyHostToDevice, ctx->stream));
        RIA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RIA_OK;
    }
    const int corr_len = window - R + 1;
    const int step = (R / 32 > 1) ? R / 32 : 1;                     // std::max(1, ref_samples / 32) (:522)
    const int n_coarse = (corr_len + step - 1) / step;
    if (2 * step + 1 > 160) return set_error(ctx, RIA_E_UNSUPPORTED, "zc: step too large");

    const size_t bb_stride = (static_cast<size_t>(window) + 1) & ~size_t(1);
    const size_t s_bb = static_cast<size_t>(n_frames) * bb_stride * sizeof(float2);
    const size_t n_cc = static_cast<size_t>(n_frames) * n_roots * n_coarse;
    const size_t a1 = (s_bb + 255) & ~size_t(255), a2 = (n_cc * sizeof(float2) + 255) & ~size_t(255);
    int rc = ensure_scratch(ctx, a1 + a2 + n_cc * sizeof(float) + 4 * sizeof(int) + 512);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    float2* d_bb = reinterpret_cast<float2*>(base);
    float2* d_cc = reinterpret_cast<float2*>(base + a1);
    float* d_cm = reinterpret_cast<float*>(base + a1 + a2);
    int* d_slots = reinterpret_cast<int*>(base + a1 + a2 + ((n_cc * sizeof(float) + 255) & ~size_t(255)));
    RIA_CUDA(ctx, cudaMemcpyAsync(d_slots, fa.root_slot, sizeof fa.root_slot, cudaMemcpyHostToDevice, ctx->stream));

    time_begin(ctx, KK_ZC_SYNC);
    {
        dim3 grid(static_cast<unsigned>((window + 255) / 256), static_cast<unsigned>(n_frames));
        zc_baseband_kernel<<<grid, 256, 0, ctx->stream>>>(samples_dev, frame_stride, window, known_cfo_dev, cfg->carrier_freq,
                                                         cfg->sample_rate, d_bb, static_cast<long long>(bb_stride));
    }
    {
        dim3 grid(static_cast<unsigned>((n_coarse + kCoarseLags - 1) / kCoarseLags), static_cast<unsigned>(n_frames));
        const size_t smem = static_cast<size_t>(n_roots) * R * sizeof(float4) +
                            (static_cast<size_t>(kCoarseLags - 1) * step + R) * sizeof(float2);
        auto launch = [&](auto kernel) -> int {
            RIA_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kernel<<<grid, kCoarseLags, smem, ctx->stream>>>(d_bb, static_cast<long long>(bb_stride), window, t->zc_interp, R, step,
                                                            n_coarse, d_slots, d_cc, d_cm);
            return RIA_OK;
        };
        if (smem > ctx->smem_optin) return set_error(ctx, RIA_E_UNSUPPORTED, "zc: coarse tile does not fit in shared memory");
        switch (n_roots) {
            case 1: rc = launch(zc_coarse_tile_kernel<1>); break;
            case 2: rc = launch(zc_coarse_tile_kernel<2>); break;
            case 3: rc = launch(zc_coarse_tile_kernel<3>); break;
            default: rc = launch(zc_coarse_tile_kernel<4>); break;
        }
        if (rc != RIA_OK) return rc;
    }
    fa.bb = d_bb; fa.bb_stride = static_cast<long long>(bb_stride); fa.window = window;
    fa.zc_g = t->zc_interp; fa.ref_samples = R; fa.step = step; fa.n_coarse = n_coarse; fa.n_roots = n_roots;
    fa.coarse_corr = d_cc; fa.coarse_mag = d_cm; fa.threshold = threshold; fa.sample_rate = cfg->sample_rate;
    fa.preamble_len = preamble_len; fa.cfg = *cfg; fa.out = out_dev;
    const size_t finish_smem = static_cast<size_t>(4) * R * sizeof(float2);
    if (finish_smem > ctx->smem_optin) return set_error(ctx, RIA_E_UNSUPPORTED, "zc: preamble of %d samples does not fit in shared memory", R);
    RIA_CUDA(ctx, cudaFuncSetAttribute(zc_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(finish_smem)));
    zc_finish_kernel<<<static_cast<unsigned>(n_frames), 128, finish_smem, ctx->stream>>>(fa);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    return RIA_OK;
}
