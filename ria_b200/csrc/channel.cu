// On-device channel simulation (K13): AWGN and Watterson 2-path Rayleigh fading, so that a
// BER/FER sweep never moves sample buffers over PCIe.
//
// Reference: SimulatedChannel::applyChannel (tools/cli_simulator.cpp:343-366) for AWGN --
//   noise_std = sqrt(mean(s^2) / 10^(snr/10)), s += N(0,1) * noise_std --
// and sim::WattersonChannel::process / updateFading (src/sim/hf_channel.hpp:107-177, 267-284):
//   noise_std = rms(nonzero input) * 10^(-snr/20);  per sample f = (1-a) f + a * sqrt(1/a) * CN(0,1),
//   a = 1 - exp(-2 pi doppler / fs);  out = s * g1 * |f1| + s[n-D] * g2 * |f2| + noise.
// The reference draws from std::mt19937 + std::normal_distribution, which is inherently serial;
// here every (frame, sample) owns a Philox4x32-10 counter, so the parity with the reference is
// statistical (same distributions and power normalisation), as SURVEY.md 8a/a20 states.

#include "ria_internal.h"

namespace ria {
namespace {

struct Philox {
    // Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi), counter = 128 bit
    __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ static inline void gen(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo, uint32_t (&out)[4]) {
        uint32_t c[4] = {static_cast<uint32_t>(ctr_lo), static_cast<uint32_t>(ctr_lo >> 32),
                         static_cast<uint32_t>(ctr_hi), static_cast<uint32_t>(ctr_hi >> 32)};
        uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// four N(0,1) from one Philox block (Box-Muller on two uniform pairs)
__device__ inline void normal4(uint64_t seed, uint64_t stream, uint64_t idx, float (&z)[4]) {
    uint32_t r[4];
    Philox::gen(seed, stream, idx, r);
    const float k = 2.3283064365386963e-10f;                  // 2^-32
    const float u0 = (static_cast<float>(r[0]) + 0.5f) * k, u1 = static_cast<float>(r[1]) * k;
    const float u2 = (static_cast<float>(r[2]) + 0.5f) * k, u3 = static_cast<float>(r[3]) * k;
    const float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2));
    float s, c;
    sincospif(2.0f * u1, &s, &c); z[0] = m0 * c; z[1] = m0 * s;
    sincospif(2.0f * u3, &s, &c); z[2] = m1 * c; z[3] = m1 * s;
}

constexpr int kChanThreads = 256;

__device__ inline float block_sum(float v, float* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < kChanThreads / 32; ++i) t += sh[i];
    return t;
}

// AWGN: out[f] = pool[f % pool_n] + N(0, sigma_f^2), sigma_f from the frame's own mean power.
__global__ void __launch_bounds__(kChanThreads)
awgn_kernel(const float* __restrict__ pool, int pool_n, int frame_len, const float* __restrict__ snr_db,
            float snr_db_all, uint64_t seed, long long frame_id0, long long n_frames,
            float* __restrict__ out, long long out_stride) {
    __shared__ float sh[kChanThreads / 32];
    const int n4 = frame_len >> 2;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float* src = pool + static_cast<size_t>((frame_id0 + f) % pool_n) * frame_len;
        float p = 0.f;
        for (int i = threadIdx.x; i < frame_len; i += kChanThreads) { const float s = __ldg(src + i); p += s * s; }
        const float power = block_sum(p, sh) / static_cast<float>(frame_len);
        const float snr = snr_db ? snr_db[f] : snr_db_all;
        const float sigma = sqrtf(power / powf(10.0f, snr / 10.0f));
        float* dst = out + f * out_stride;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
        if (vec_ok) {
            for (int i = threadIdx.x; i < n4; i += kChanThreads) {
                float z[4];
                normal4(seed, static_cast<uint64_t>(frame_id0 + f), static_cast<uint64_t>(i), z);
                const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
                __stcs(reinterpret_cast<float4*>(dst) + i,
                       make_float4(s.x + sigma * z[0], s.y + sigma * z[1], s.z + sigma * z[2], s.w + sigma * z[3]));
            }
        } else {
            for (int i = threadIdx.x; i < n4; i += kChanThreads) {
                float z[4];
                normal4(seed, static_cast<uint64_t>(frame_id0 + f), static_cast<uint64_t>(i), z);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[4 * i + j] = __ldg(src + 4 * i + j) + sigma * z[j];
            }
        }
        for (int i = (n4 << 2) + threadIdx.x; i < frame_len; i += kChanThreads) {
            float z[4];
            normal4(seed, static_cast<uint64_t>(frame_id0 + f), static_cast<uint64_t>(n4 + i), z);
            dst[i] = __ldg(src + i) + sigma * z[0];
        }
    }
}


// Watterson two-path Rayleigh channel (src/sim/hf_channel.hpp:107-177, 267-284).  One CTA per
// frame.  The fading taps follow f[n] = (1-a) f[n-1] + a sqrt(1/a) CN(0,1), a first-order linear
// recurrence, evaluated in parallel: every thread owns a contiguous chunk, a first pass finds the
// zero-state response of each chunk, the chunk carries are chained (affine maps with the common
// factor (1-a)^S), a second pass regenerates the same Philox draws with the right initial state.
struct WattArgs {
    const float* pool; int pool_n; int frame_len;
    const float* snr_db_vec; ria_watterson_config cfg;
    unsigned long long seed; long long frame_id0, n_frames;
    float* out; long long out_stride;
};

__device__ inline void fading_draw(unsigned long long seed, unsigned long long stream, unsigned long long i, float (&z)[4]) {
    normal4(seed ^ 0x5bd1e995u, stream, (i << 1), z);
}

__global__ void __launch_bounds__(kChanThreads)
watterson_kernel(const WattArgs a) {
    __shared__ float sh[kChanThreads / 32];
    __shared__ float2 carry1[kChanThreads + 1], carry2[kChanThreads + 1];
    __shared__ float2 loc1[kChanThreads], loc2[kChanThreads];
    const int tid = threadIdx.x;
    const int L = a.frame_len;
    const int S = (L + kChanThreads - 1) / kChanThreads;
    const float fs = static_cast<float>(a.cfg.sample_rate);
    const int D = static_cast<int>(a.cfg.delay_spread_ms * fs / 1000.0f);
    const float nd = a.cfg.doppler_spread_hz / fs;
    const float alpha = 1.0f - expf(-2.0f * 3.14159265358979f * nd);
    const bool fading = a.cfg.fading_enabled && alpha > 0.0f;
    const float one_m = 1.0f - alpha;
    const float nscale = fading ? alpha * sqrtf(1.0f / alpha) : 0.0f;      // alpha * noise_scale
    for (long long f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
        const unsigned long long gid = static_cast<unsigned long long>(a.frame_id0 + f);
        const float* src = a.pool + static_cast<size_t>(gid % a.pool_n) * L;
        // rms over non-silent samples (:111-121)
        float p = 0.f, cnt = 0.f;
        for (int i = tid; i < L; i += kChanThreads) { const float v = __ldg(src + i); if (fabsf(v) > 1e-6f) { p += v * v; cnt += 1.0f; } }
        const float ps = block_sum(p, sh);
        const float cs = block_sum(cnt, sh);
        const float rms = (cs > 0.0f) ? sqrtf(ps / cs) : 0.1f;
        const float snr = a.snr_db_vec ? a.snr_db_vec[f] : a.cfg.snr_db;
        const float sigma = a.cfg.noise_enabled ? rms * powf(10.0f, -snr / 20.0f) : 0.0f;

        const int i0 = tid * S, i1 = min(L, i0 + S);
        if (fading) {
            float2 l1 = make_float2(0.f, 0.f), l2 = make_float2(0.f, 0.f);
            for (int i = i0; i < i1; ++i) {
                float z[4];
                fading_draw(a.seed, gid, static_cast<unsigned long long>(i), z);
                l1.x = one_m * l1.x + nscale * z[0]; l1.y = one_m * l1.y + nscale * z[1];
                l2.x = one_m * l2.x + nscale * z[2]; l2.y = one_m * l2.y + nscale * z[3];
            }
            loc1[tid] = l1; loc2[tid] = l2;
            __syncthreads();
            if (tid == 0) {
                float2 c1, c2;
                if (a.cfg.stationary_start) {
                    // stationary distribution of the recurrence: per component variance 1/(2-alpha)
                    float z[4];
                    normal4(a.seed ^ 0x9e3779b9u, gid, 0xFFFFFFFFull, z);
                    const float sd = sqrtf(1.0f / (2.0f - alpha));
                    c1 = make_float2(sd * z[0], sd * z[1]); c2 = make_float2(sd * z[2], sd * z[3]);
                } else {
                    c1 = make_float2(1.0f, 0.0f); c2 = make_float2(1.0f, 0.0f);      // constructor state (:68-69)
                }
                const float aS = powf(one_m, static_cast<float>(S));
                for (int t = 0; t < kChanThreads; ++t) {
                    carry1[t] = c1; carry2[t] = c2;
                    const int len = min(L, (t + 1) * S) - min(L, t * S);
                    const float at = (len == S) ? aS : powf(one_m, static_cast<float>(len));
                    c1 = make_float2(at * c1.x + loc1[t].x, at * c1.y + loc1[t].y);
                    c2 = make_float2(at * c2.x + loc2[t].x, at * c2.y + loc2[t].y);
                }
            }
            __syncthreads();
        }
        float2 f1 = fading ? carry1[tid] : make_float2(1.f, 0.f), f2 = fading ? carry2[tid] : make_float2(1.f, 0.f);
        float* dst = a.out + f * a.out_stride;
        for (int i = i0; i < i1; ++i) {
            float h1 = 1.0f, h2 = 1.0f;
            if (fading) {
                float z[4];
                fading_draw(a.seed, gid, static_cast<unsigned long long>(i), z);
                f1.x = one_m * f1.x + nscale * z[0]; f1.y = one_m * f1.y + nscale * z[1];
                f2.x = one_m * f2.x + nscale * z[2]; f2.y = one_m * f2.y + nscale * z[3];
                h1 = sqrtf(f1.x * f1.x + f1.y * f1.y); h2 = sqrtf(f2.x * f2.x + f2.y * f2.y);
            }
            const float s = __ldg(src + i);
            float o;
            if (a.cfg.multipath_enabled && D > 0) {
                const float delayed = (i >= D) ? __ldg(src + i - D) : 0.0f;
                o = s * a.cfg.path1_gain * h1 + delayed * a.cfg.path2_gain * h2;
            } else {
                o = s * h1;
            }
            if (sigma > 0.0f) {
                float z[4];
                normal4(a.seed, gid, (static_cast<unsigned long long>(i) << 1) | 1ull, z);
                o += sigma * z[0];
            }
            dst[i] = o;
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace ria

extern "C" int ria_watterson_preset(int condition, float snr_db, ria_watterson_config* cfg) {
    if (!cfg) return RIA_E_INVAL;
    // itu_r_f1487::{awgn, good, moderate, poor, flutter} (src/sim/hf_channel.hpp:411-488)
    ria_watterson_config c{snr_db, 0.0f, 0.0f, 0.707f, 0.707f, 48000, 1, 1, 1, 1};
    switch (condition) {
        case 0: c.delay_spread_ms = 0.0f; c.doppler_spread_hz = 0.0f; c.path1_gain = 1.0f; c.path2_gain = 0.0f;
                c.fading_enabled = 0; c.multipath_enabled = 0; break;
        case 1: c.delay_spread_ms = 0.5f; c.doppler_spread_hz = 0.1f; break;
        case 2: c.delay_spread_ms = 1.0f; c.doppler_spread_hz = 0.5f; break;
        case 3: c.delay_spread_ms = 2.0f; c.doppler_spread_hz = 1.0f; break;
        case 4: c.delay_spread_ms = 0.5f; c.doppler_spread_hz = 10.0f; break;
        default: return RIA_E_INVAL;
    }
    *cfg = c;
    return RIA_OK;
}

extern "C" int ria_channel_watterson_batch_dev(ria_ctx* ctx, const ria_watterson_config* cfg,
                                               const float* tx_pool_dev, int32_t pool_frames, int32_t frame_len,
                                               const float* snr_db_dev, uint64_t seed, int64_t first_frame_id,
                                               int64_t n_frames, float* out_dev, int64_t out_stride) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || pool_frames <= 0 || frame_len <= 0 || out_stride < frame_len || cfg->sample_rate == 0)
        return set_error(ctx, RIA_E_INVAL, "watterson: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!tx_pool_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "watterson: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    WattArgs a{tx_pool_dev, pool_frames, frame_len, snr_db_dev, *cfg, seed, first_frame_id, n_frames, out_dev, out_stride};
    long long grid = static_cast<long long>(ctx->sm_count) * 8;
    if (grid > n_frames) grid = n_frames;
    time_begin(ctx, KK_WATTERSON);
    watterson_kernel<<<static_cast<unsigned>(grid), kChanThreads, 0, ctx->stream>>>(a);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}

extern "C" int ria_channel_awgn_batch_dev(ria_ctx* ctx, const float* tx_pool_dev, int32_t pool_frames,
                                          int32_t frame_len, const float* snr_db_dev, float snr_db,
                                          uint64_t seed, int64_t first_frame_id, int64_t n_frames,
                                          float* out_dev, int64_t out_stride) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_frames < 0 || pool_frames <= 0 || frame_len <= 0 || out_stride < frame_len)
        return set_error(ctx, RIA_E_INVAL, "awgn: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!tx_pool_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "awgn: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    long long grid = static_cast<long long>(ctx->sm_count) * 8;
    if (grid > n_frames) grid = n_frames;
    time_begin(ctx, KK_AWGN);
    awgn_kernel<<<static_cast<unsigned>(grid), kChanThreads, 0, ctx->stream>>>(
        tx_pool_dev, pool_frames, frame_len, snr_db_dev, snr_db, seed, first_frame_id, n_frames, out_dev, out_stride);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
