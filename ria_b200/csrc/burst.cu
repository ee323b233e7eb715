// Batched fec::BurstInterleaver::deinterleave (src/fec/burst_interleaver.cpp:39-78) for sm_100a:
// the byte-level block de-interleaver StreamingDecoder::finalizeBurstGroup applies to the soft
// bits of a burst group of N OFDM frames before they go to decodeFrame
// (src/gui/modem/streaming_decoder.cpp:3209-3216).
//
//   TX: physical[pf][pb] = logical[f][b]  with flat = N*b + f, pf = flat / 324, pb = flat % 324
//   RX: logical[f][8b .. 8b+7] = physical[pf][8pb .. 8pb+7]
//
// A pure permutation of 32-byte groups (the eight soft bits of one coded byte): one thread per
// group, two 128-bit loads and stores; writes are contiguous per frame, reads are whole 32-byte
// sectors.  HBM-bound: 2592 x 4 B in and out per frame.

#include "ria_internal.h"

namespace ria {
namespace {

constexpr int kBytesPerFrame = 324;      // BurstInterleaver::BYTES_PER_FRAME (burst_interleaver.hpp:30)
constexpr int kBitsPerFrame = 2592;      // BITS_PER_FRAME

__global__ void burst_deinterleave_kernel(const float* __restrict__ phys, long long in_stride, int N, long long n_groups,
                                          float* __restrict__ logi, long long out_stride) {
    const long long per_group = static_cast<long long>(N) * kBytesPerFrame;
    const long long total = n_groups * per_group;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long g = t / per_group;
        const int r = static_cast<int>(t - g * per_group);
        const int f = r / kBytesPerFrame, b = r - f * kBytesPerFrame;
        int pf = f, pb = b;
        if (N >= 2) {                                    // N < 2: returned unchanged (:43)
            const int flat = N * b + f;
            pf = flat / kBytesPerFrame; pb = flat - pf * kBytesPerFrame;
        }
        const float4* src = reinterpret_cast<const float4*>(phys + (g * N + pf) * in_stride + 8 * pb);
        float4* dst = reinterpret_cast<float4*>(logi + (g * N + f) * out_stride + 8 * b);
        const float4 lo = __ldcs(src), hi = __ldcs(src + 1);
        __stcs(dst, lo); __stcs(dst + 1, hi);
    }
}

}  // namespace
}  // namespace ria

extern "C" int ria_burst_deinterleave_batch_dev(ria_ctx* ctx, const float* physical_dev, int32_t in_stride,
                                                int32_t group_size, int64_t n_groups,
                                                float* logical_dev, int32_t out_stride) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_groups < 0 || group_size < 1) return set_error(ctx, RIA_E_INVAL, "burst: bad sizes");
    if (n_groups == 0) return RIA_OK;
    if (!physical_dev || !logical_dev) return set_error(ctx, RIA_E_INVAL, "burst: null buffer");
    if (physical_dev == logical_dev) return set_error(ctx, RIA_E_INVAL, "burst: in-place de-interleaving is not supported");
    // BurstInterleaver::deinterleave throws std::invalid_argument when a frame has < 2592 soft bits (:51-55)
    if (in_stride < kBitsPerFrame || out_stride < kBitsPerFrame)
        return set_error(ctx, RIA_E_INVAL, "burst: soft bits size mismatch (need >= 2592 per frame)");
    if ((in_stride & 3) || (out_stride & 3) || (reinterpret_cast<uintptr_t>(physical_dev) & 15) ||
        (reinterpret_cast<uintptr_t>(logical_dev) & 15))
        return set_error(ctx, RIA_E_INVAL, "burst: buffers must be 16-byte aligned with strides that are multiples of 4");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long total = n_groups * group_size * kBytesPerFrame;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(ctx->sm_count) * 32;
    if (blocks > cap) blocks = cap;
    time_begin(ctx, KK_CHASE);
    burst_deinterleave_kernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(
        physical_dev, in_stride, group_size, n_groups, logical_dev, out_stride);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
