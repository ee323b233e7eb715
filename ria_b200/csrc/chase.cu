// HARQ chase combining of cached soft bits (K10).
//
// fec::ChaseCache::store (src/fec/chase_cache.cpp:27-88) keeps, per {seq, src, dst} key and
// codeword index, the sum of the LLRs of every reception: the first reception copies, later ones
// do `existing[i] += soft_bits[i]` in reception order (:75-85).  The cache policy (key map,
// <= 4 combines per codeword, 16 entries LRU, 30 s TTL) is host logic (ria_b200/fec.py ChaseCache);
// this kernel is the arithmetic on cache slots that live in HBM, so retransmitted frames of a
// BER sweep never leave the device between demodulation and LDPC.
//   item i:  slot[i] < 0        -> ignored
//            first[i] != 0      -> acc[slot] = llr_i          (first reception)
//            otherwise          -> acc[slot] += llr_i          (fp32, one add per element)
// A slot may appear at most once per call (the host issues one call per reception round), which
// keeps the accumulation order equal to the reception order.

#include "ria_internal.h"

namespace ria {
namespace {

__global__ void chase_combine_kernel(float* __restrict__ acc, const int32_t* __restrict__ slot,
                                     const uint8_t* __restrict__ first, const float* __restrict__ llr,
                                     long long llr_stride, long long n) {
    const long long i = blockIdx.x;
    if (i >= n) return;
    const int s = slot[i];
    if (s < 0) return;
    const bool cp = first && first[i];
    float4* a = reinterpret_cast<float4*>(acc + static_cast<size_t>(s) * RIA_LDPC_N);
    const float* src = llr + i * llr_stride;
    for (int k = threadIdx.x; k < RIA_LDPC_N / 4; k += blockDim.x) {
        const float4 v = make_float4(src[4 * k], src[4 * k + 1], src[4 * k + 2], src[4 * k + 3]);
        if (cp) a[k] = v;
        else {
            const float4 e = a[k];
            a[k] = make_float4(__fadd_rn(e.x, v.x), __fadd_rn(e.y, v.y), __fadd_rn(e.z, v.z), __fadd_rn(e.w, v.w));
        }
    }
}

}  // namespace
}  // namespace ria

extern "C" int ria_chase_combine_batch_dev(ria_ctx* ctx, float* acc_dev, const int32_t* slot_dev,
                                           const uint8_t* first_dev, const float* llr_dev, int64_t llr_stride,
                                           int64_t n) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n < 0 || llr_stride < RIA_LDPC_N) return set_error(ctx, RIA_E_INVAL, "chase: bad sizes");
    if (n == 0) return RIA_OK;
    if (!acc_dev || !slot_dev || !llr_dev) return set_error(ctx, RIA_E_INVAL, "chase: null buffer");
    if ((reinterpret_cast<uintptr_t>(acc_dev) & 15) != 0) return set_error(ctx, RIA_E_INVAL, "chase: acc_dev must be 16-byte aligned");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    time_begin(ctx, KK_CHASE);
    chase_combine_kernel<<<static_cast<unsigned>(n), 192, 0, ctx->stream>>>(acc_dev, slot_dev, first_dev, llr_dev, llr_stride, n);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
