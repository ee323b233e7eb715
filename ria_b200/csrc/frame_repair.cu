// "LDPC false positive recovery" of v2::decodeFixedFrame for sm_100a
// (src/protocol/frame_v2.cpp:1558-1916).
//
// When all four codewords of a frame pass parity but the reassembled frame does not verify
// (header or frame CRC), the reference (1) searches 1- and 2-bit flips of the header codeword, or
// 1-bit flips of the payload / stored CRC through CRC deltas, then 2/3/4-bit flips among the 30
// (15) "suspect" bits with the weakest soft values, (2) re-decodes each codeword with four other
// min-sum factors and keeps a different codeword if the frame then verifies, (3) otherwise marks
// all four codewords as failed.  The searches are byte/integer logic on one frame; they live in
// frame_repair_core.h (host/device, checked on the host against the reference) and run on lane 0
// of the warp that owns the frame; the warp gathers the de-interleaved soft bits and runs the
// re-decodes.  Frames are listed by a one-thread-per-frame validity pass; the list length stays
// on the device.

#include "ldpc_core.cuh"
#include "frame_repair_core.h"

namespace ria {

namespace {

using ldpc_core::kN;
using ldpc_core::LdpcGather;

constexpr int kRepairWarps = 4;
constexpr int kCwPad = 64;

struct RepairArgs {
    const float* soft_g; LdpcGather gather;
    const int* list; const unsigned* list_len; unsigned* counter;
    const uint16_t* chk_var_g; const uint16_t* var_slot_g;
    int k, m, dv_max, max_iter, bpc;
    uint8_t* info_g; int info_stride; uint8_t* ok_g;
    uint8_t* repair_g;          // [n_frames]: 0 untouched, 1 repaired, 2 given up (all codewords marked failed)
};

__host__ __device__ inline size_t repair_union_bytes(int k, int m) {
    const size_t search = static_cast<size_t>(ria_repair::kMaxBits) * (2 + 4 + 2);
    const int kpad = (k + 3) & ~3;
    const size_t decode = (static_cast<size_t>(kN) + kpad + static_cast<size_t>(m) * 8) * 4;
    return ((search > decode ? search : decode) + 15) & ~size_t(15);
}
__host__ __device__ inline size_t repair_warp_bytes(int k, int m) {
    return 4 * kN * 4 + repair_union_bytes(k, m) + 2 * ria_repair::kMaxFrameBytes + 5 * kCwPad;
}

__global__ void __launch_bounds__(kRepairWarps * 32)
frame_repair_kernel(const RepairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.k, m = a.m, dv_max = a.dv_max, bpc = a.bpc;
    uint4* chk_var = reinterpret_cast<uint4*>(smem_raw);
    uint16_t* var_slot = reinterpret_cast<uint16_t*>(chk_var + m);
    const size_t tab_bytes = ldpc_core::ldpc_tab_bytes(k, m, dv_max);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* w = smem_raw + tab_bytes + repair_warp_bytes(k, m) * warp;
    float* soft4 = reinterpret_cast<float*>(w);                         // [4][648]
    unsigned char* un = w + 4 * kN * 4;                                 // union: search scratch / decode state
    uint16_t* deltas = reinterpret_cast<uint16_t*>(un + static_cast<size_t>(ria_repair::kMaxBits) * 4);
    float* key = reinterpret_cast<float*>(un);
    uint16_t* val = deltas + ria_repair::kMaxBits;
    float* llr = reinterpret_cast<float*>(un);
    float* tot = llr + kN;
    float4* msg = reinterpret_cast<float4*>(tot + ((k + 3) & ~3));
    uint8_t* frame = un + repair_union_bytes(k, m);
    uint8_t* trial = frame + ria_repair::kMaxFrameBytes;
    uint8_t* cwdata = trial + ria_repair::kMaxFrameBytes;               // [4][64]
    uint8_t* newcw = cwdata + 4 * kCwPad;                               // [64]
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.chk_var_g);
        for (int i = threadIdx.x; i < m; i += blockDim.x) chk_var[i] = src[i];
        const int nvs = dv_max * k;
        for (int i = threadIdx.x; i < nvs; i += blockDim.x) var_slot[i] = a.var_slot_g[i];
    }
    __syncthreads();
    const unsigned n_units = *a.list_len;
    ria_repair::Frame fr;
    for (int c = 0; c < 4; ++c) fr.cw[c] = cwdata + c * kCwPad;
    fr.bpc = bpc;
    const ria_repair::Scratch scr{deltas, key, val, frame, trial};

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_units) break;
        const long long f = a.list[t];
        for (int c = 0; c < 4; ++c) {
            ldpc_core::gather_codeword(a.soft_g, f * 4 + c, a.gather, soft4 + c * kN, lane);
            for (int b = lane; b < bpc; b += 32) cwdata[c * kCwPad + b] = a.info_g[(f * 4 + c) * a.info_stride + b];
        }
        __syncwarp();
        int recovered = 0;
        if (lane == 0) recovered = ria_repair::repair_bitflips(fr, soft4, scr) ? 1 : 0;
        recovered = __shfl_sync(0xffffffffu, recovered, 0);
        if (!recovered) {
            // fallback (:1848-1876): re-decode with other min-sum factors
            const float factors[4] = {0.75f, 0.625f, 0.5f, 0.875f};
            for (int at = 0; at < 4 && !recovered; ++at) {
                for (int c = 0; c < 4 && !recovered; ++c) {
                    __syncwarp();
                    for (int j = lane; j < kN; j += 32) llr[j] = soft4[c * kN + j];
                    bool success; int iters;
                    ldpc_core::decode_codeword(llr, tot, msg, chk_var, var_slot, k, m, dv_max, a.max_iter, factors[at], lane, success, iters);
                    if (!success) continue;
                    ldpc_core::pack_info(tot, k, newcw, kCwPad, lane);
                    __syncwarp();
                    if (lane == 0) {
                        bool differs = false;
                        for (int b = 0; b < bpc; ++b) differs |= (newcw[b] != fr.cw[c][b]);
                        if (differs) {
                            for (int b = 0; b < bpc; ++b) { const uint8_t o = fr.cw[c][b]; fr.cw[c][b] = newcw[b]; newcw[b] = o; }
                            if (ria_repair::frame_valid(fr, trial)) recovered = 1;
                            else for (int b = 0; b < bpc; ++b) fr.cw[c][b] = newcw[b];
                        }
                    }
                    recovered = __shfl_sync(0xffffffffu, recovered, 0);
                }
            }
        }
        __syncwarp();
        if (recovered) {
            for (int c = 0; c < 4; ++c)
                for (int b = lane; b < bpc; b += 32) a.info_g[(f * 4 + c) * a.info_stride + b] = cwdata[c * kCwPad + b];
        } else if (lane < 4) {
            a.ok_g[f * 4 + lane] = 0;                   // "recovery FAILED, marking as decode failure" (:1879-1884)
        }
        if (lane == 0 && a.repair_g) a.repair_g[f] = recovered ? 1 : 2;
        __syncwarp();
    }
}

// frames whose four codewords all decoded but whose reassembled frame does not verify (:1565-1578)
__global__ void frame_repair_list_bytes_kernel(const uint8_t* __restrict__ info, int info_stride, const uint8_t* __restrict__ ok,
                                         long long n_frames, int bpc, int* __restrict__ list, unsigned* __restrict__ list_len) {
    const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (f >= n_frames) return;
    const uchar4 o = reinterpret_cast<const uchar4*>(ok)[f];
    if (!(o.x && o.y && o.z && o.w)) return;
    uint8_t cw[4][ria_repair::kMaxCwBytes];
    uint8_t tmp[ria_repair::kMaxFrameBytes];
    ria_repair::Frame fr;
    for (int c = 0; c < 4; ++c) {
        fr.cw[c] = cw[c];
        for (int b = 0; b < bpc; ++b) cw[c][b] = info[(f * 4 + c) * info_stride + b];
    }
    fr.bpc = bpc;
    if (!ria_repair::frame_valid(fr, tmp)) list[atomicAdd(list_len, 1u)] = static_cast<int>(f);
}


// The same for the usual 72-byte codeword stride: a warp takes 32 consecutive frames, stages their
// 288 info bytes through padded shared rows with coalesced 32-bit loads, and every lane checks its frame
// out of shared memory.
constexpr int kListWarps = 2;
__global__ void __launch_bounds__(kListWarps * 32)
frame_repair_list_kernel(const uint8_t* __restrict__ info, const uint8_t* __restrict__ ok,
                         long long n_frames, int bpc, int* __restrict__ list, unsigned* __restrict__ list_len) {
    constexpr int kWords = 72, kRow = 73;
    __shared__ uint32_t rows[kListWarps][32 * kRow];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long f0 = (blockIdx.x * static_cast<long long>(kListWarps) + warp) * 32;
    if (f0 >= n_frames) return;
    const int nv = static_cast<int>(n_frames - f0 < 32 ? n_frames - f0 : 32);
    uint32_t* in = rows[warp];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(info + f0 * 4 * 72);
    for (int g = lane; g < nv * kWords; g += 32) in[(g / kWords) * kRow + (g % kWords)] = src[g];
    __syncwarp();
    if (lane >= nv) return;
    const long long f = f0 + lane;
    const uchar4 o = reinterpret_cast<const uchar4*>(ok)[f];
    if (!(o.x && o.y && o.z && o.w)) return;
    uint8_t tmp[ria_repair::kMaxFrameBytes];
    ria_repair::Frame fr;
    uint8_t* row = reinterpret_cast<uint8_t*>(in + lane * kRow);
    for (int c = 0; c < 4; ++c) fr.cw[c] = row + c * 72;
    fr.bpc = bpc;
    if (!ria_repair::frame_valid(fr, tmp)) list[atomicAdd(list_len, 1u)] = static_cast<int>(f);
}

}  // namespace

int frame_repair_launch(ria_ctx* ctx, int rate, int max_iter, const float* soft_dev, int64_t n_frames, int soft_stride,
                        int step, uint8_t* info_dev, int info_stride, uint8_t* ok_dev, uint8_t* repair_dev,
                        int* list_scratch) {
    if (n_frames > 0x7fffffffLL) return set_error(ctx, RIA_E_INVAL, "frame repair: too many frames");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const LdpcCodeDev* t = nullptr;
    int rc = ldpc_tables_dev(ctx, rate, &t);
    if (rc != RIA_OK) return rc;
    const int bpc = t->k / 8;
    if (bpc > ria_repair::kMaxCwBytes) return set_error(ctx, RIA_E_UNSUPPORTED, "frame repair: rate %d not supported", rate);
    const size_t smem = ldpc_core::ldpc_tab_bytes(t->k, t->m, t->dv_max) + kRepairWarps * repair_warp_bytes(t->k, t->m);
    if (smem > ctx->smem_optin) return set_error(ctx, RIA_E_UNSUPPORTED, "frame repair: kernel does not fit in shared memory");
    auto kern = frame_repair_kernel;
    int ctas_per_sm = 0;
    RIA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
    RIA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kRepairWarps * 32, smem));
    if (ctas_per_sm < 1) return set_error(ctx, RIA_E_UNSUPPORTED, "frame repair: kernel does not fit (smem %zu)", smem);
    unsigned* list_len = ctx->work_counter + 42;
    unsigned* counter = ctx->work_counter + 43;
    cudaStream_t s = ctx->stream;
    RIA_CUDA(ctx, cudaMemsetAsync(list_len, 0, 2 * sizeof(unsigned), s));
    if (repair_dev) RIA_CUDA(ctx, cudaMemsetAsync(repair_dev, 0, static_cast<size_t>(n_frames), s));
    int inv_step = 0;
    if (step > 0)
        for (int x = 1; x < kN; ++x) if ((static_cast<long long>(x) * step) % kN == 1) { inv_step = x; break; }
    const int vec_ok = (soft_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(soft_dev) & 15) == 0);
    RepairArgs a{};
    a.soft_g = soft_dev; a.gather = LdpcGather{1, soft_stride, step, inv_step, vec_ok};
    a.list = list_scratch; a.list_len = list_len; a.counter = counter;
    a.chk_var_g = t->chk_var; a.var_slot_g = t->var_slot;
    a.k = t->k; a.m = t->m; a.dv_max = t->dv_max; a.max_iter = max_iter; a.bpc = bpc;
    a.info_g = info_dev; a.info_stride = info_stride; a.ok_g = ok_dev; a.repair_g = repair_dev;
    const int threads = 128;
    const unsigned blocks = static_cast<unsigned>((n_frames + threads - 1) / threads);
    time_begin(ctx, KK_FRAME_REPAIR);
    if (info_stride == 72 && (reinterpret_cast<uintptr_t>(info_dev) & 3) == 0) {
        const unsigned lb = static_cast<unsigned>((n_frames + kListWarps * 32 - 1) / (kListWarps * 32));
        frame_repair_list_kernel<<<lb, kListWarps * 32, 0, s>>>(info_dev, ok_dev, n_frames, bpc, list_scratch, list_len);
    } else {
        frame_repair_list_bytes_kernel<<<blocks, threads, 0, s>>>(info_dev, info_stride, ok_dev, n_frames, bpc, list_scratch, list_len);
    }
    long long grid = static_cast<long long>(ctx->sm_count) * ctas_per_sm;
    const long long want = (n_frames + kRepairWarps - 1) / kRepairWarps;
    if (grid > want) grid = want;
    frame_repair_kernel<<<static_cast<unsigned>(grid), kRepairWarps * 32, smem, s>>>(a);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return RIA_OK;
}

}  // namespace ria

