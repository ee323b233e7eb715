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

}  // namespace
}  // namespace ria

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
