// Batched flooding normalised min-sum LDPC decoder (n = 648) for sm_100a.
//
// Replaces LDPCDecoder::decodeSoft / Impl::decodeBP (src/fec/ldpc_decoder.cpp:154-260, 284-429).
// Bit-exact contract: hard bits, success flag and iteration count equal the reference's on the
// same LLRs.  That pins (a) the flooding schedule, (b) the fp32 operation order -- llr_total[j]
// is llr_in[j] plus the check messages in ascending check index (:207-214) -- (c) the sign test
// `msg < 0` (:194), (d) the +-50 clamp written as std::max(-50, std::min(50, x)) (:223) and
// (e) `sign * min_abs * factor` (:201).  All float ops below use explicit _rn intrinsics so no
// FMA contraction can change a bit.
//
// Mapping (B200): one warp per codeword, W warps (codewords) per CTA, persistent CTAs pulling
// codewords from an atomic counter.  Everything a codeword needs between iterations lives in
// shared memory; HBM is touched once for the 648 LLRs and once for the <= 68 output bytes.
//
//   per codeword smem:  llr[648] | tot[k] (info totals) | msg[2][m][4]
//   msg holds, per check i, 8 slots: 0..5 = c2v of the info edges, 6 = c2v of the identity
//   edge, 7 = running total of parity variable k+i (only its own check ever reads it).
//   The two halves [0][i][0..3] / [1][i][0..3] make the per-check accesses two conflict-free
//   128-bit shared loads/stores with lane stride 16 B.
//
//   phase A (lane = check):  v2c_e = clamp(total[var_e] - c2v_e)   (iteration 0: v2c = llr)
//                            parity of sign(total) -> early-exit vote for the previous iteration
//                            c2v_e  = (xor of other signs) * fl(min_{e'!=e}|v2c_e'| * factor)
//                            parity variable: total = llr[k+i] + c2v_identity  (degree 1)
//   phase B (lane = info variable): total[j] = llr[j] + sum_asc_check c2v   (gather via var_slot)
//
// The check update runs speculatively for the iteration after the one that converged (its
// results are discarded); this saves a separate parity pass per iteration.

#include "ldpc_core.cuh"

namespace ria {

namespace {

using ldpc_core::kN;
using ldpc_core::LdpcGather;
constexpr int kMaxWarpsPerCta = 16;

__global__ void __launch_bounds__(kMaxWarpsPerCta * 32)
ldpc_decode_kernel(const float* __restrict__ llr_g, long long n_cw, const LdpcGather gather,
                   const uint16_t* __restrict__ chk_var_g, const uint16_t* __restrict__ var_slot_g,
                   int k, int m, int dv_max, int max_iter, float factor,
                   uint8_t* __restrict__ info_g, int info_stride,
                   uint8_t* __restrict__ ok_g, int32_t* __restrict__ iters_g,
                   unsigned int* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- carve shared memory: tables (shared by all warps), then per-warp state ----
    uint4* chk_var = reinterpret_cast<uint4*>(smem_raw);                       // [m] x 8 u16
    uint16_t* var_slot = reinterpret_cast<uint16_t*>(chk_var + m);             // [dv_max][k]
    size_t tab_bytes = static_cast<size_t>(m) * 16 + static_cast<size_t>(dv_max) * k * 2;
    tab_bytes = (tab_bytes + 15) & ~size_t(15);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpad = (k + 3) & ~3;                                             // keep msg 16 B aligned
    const size_t per_warp = (static_cast<size_t>(kN) + kpad + static_cast<size_t>(m) * 8) * 4;
    float* llr = reinterpret_cast<float*>(smem_raw + tab_bytes + per_warp * warp);
    float* tot = llr + kN;
    float4* msg = reinterpret_cast<float4*>(tot + kpad);                       // [2][m]

    {   // cooperative table load (once per persistent CTA)
        const uint4* src = reinterpret_cast<const uint4*>(chk_var_g);
        for (int i = threadIdx.x; i < m; i += blockDim.x) chk_var[i] = src[i];
        const int nvs = dv_max * k;
        for (int i = threadIdx.x; i < nvs; i += blockDim.x) var_slot[i] = var_slot_g[i];
    }
    __syncthreads();

    // The clean-codeword shortcut (ldpc_core.cuh) costs one parity pass; at an operating point where the
    // channel hard decisions are rarely a codeword it is switched off after 16 tries with < 25 % hits
    // (warp-uniform, results do not depend on it).
    int sc_tries = 0, sc_hits = 0;
    for (;;) {
        long long cw;
        {
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(counter, 1u);
            cw = static_cast<long long>(__shfl_sync(0xffffffffu, t, 0));
        }
        if (cw >= n_cw) break;
        ldpc_core::gather_codeword(llr_g, cw, gather, llr, lane);
        bool success, took; int iters;
        const bool try_sc = sc_tries < 16 || 4 * sc_hits >= sc_tries;
        ldpc_core::decode_codeword(llr, tot, msg, chk_var, var_slot, k, m, dv_max, max_iter, factor, lane, success, iters,
                                   try_sc, &took);
        if (try_sc) { ++sc_tries; sc_hits += took ? 1 : 0; }
        ldpc_core::pack_info((try_sc && took) ? llr : tot, k, info_g + cw * info_stride, info_stride, lane);
        if (lane == 0) {
            ok_g[cw] = success ? 1 : 0;
            iters_g[cw] = iters;
        }
        __syncwarp();
    }
}

size_t ldpc_smem_bytes(int k, int m, int dv_max, int warps) {
    size_t tab = static_cast<size_t>(m) * 16 + static_cast<size_t>(dv_max) * k * 2;
    tab = (tab + 15) & ~size_t(15);
    const int kpad = (k + 3) & ~3;
    return tab + warps * (static_cast<size_t>(kN) + kpad + static_cast<size_t>(m) * 8) * 4;
}

}  // namespace

int ldpc_tables_dev(ria_ctx* ctx, int rate, const LdpcCodeDev** out) {
    if (!ldpc_rate_valid(rate)) return set_error(ctx, RIA_E_INVAL, "ldpc: bad rate %d", rate);
    LdpcCodeDev& d = ctx->ldpc[rate];
    if (!d.ready) {
        const LdpcCodeHost* h = nullptr;
        try { h = &ldpc_code_host(rate); }
        catch (const std::exception& e) { return set_error(ctx, RIA_E_INVAL, "ldpc: %s", e.what()); }
        RIA_CUDA(ctx, cudaMalloc(&d.chk_var, h->chk_var.size() * sizeof(uint16_t)));
        RIA_CUDA(ctx, cudaMalloc(&d.var_slot, h->var_slot.size() * sizeof(uint16_t)));
        RIA_CUDA(ctx, cudaMemcpy(d.chk_var, h->chk_var.data(), h->chk_var.size() * sizeof(uint16_t),
                                 cudaMemcpyHostToDevice));
        RIA_CUDA(ctx, cudaMemcpy(d.var_slot, h->var_slot.data(), h->var_slot.size() * sizeof(uint16_t),
                                 cudaMemcpyHostToDevice));
        d.k = h->k; d.m = h->m; d.dv_max = h->dv_max;
        d.ready = true;
    }
    *out = &d;
    return RIA_OK;
}

}  // namespace ria

namespace ria {
int ldpc_launch(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor, const float* llr_dev, int64_t n_cw,
                int frame_mode, int soft_stride, int step,
                uint8_t* info_dev, int info_stride, uint8_t* ok_dev, int32_t* iters_dev);
}

extern "C" int ria_ldpc_decode_batch_dev(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor,
                                         const float* llr_dev, int64_t n_cw,
                                         uint8_t* info_dev, int info_stride,
                                         uint8_t* ok_dev, int32_t* iters_dev) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_cw < 0 || max_iter < 0) return set_error(ctx, RIA_E_INVAL, "ldpc: negative size");
    if (n_cw == 0) return RIA_OK;
    if (!llr_dev || !info_dev || !ok_dev || !iters_dev) return set_error(ctx, RIA_E_INVAL, "ldpc: null buffer");
    if ((reinterpret_cast<uintptr_t>(llr_dev) & 15) != 0)
        return set_error(ctx, RIA_E_INVAL, "ldpc: llr_dev must be 16-byte aligned");
    return ldpc_launch(ctx, rate, max_iter, min_sum_factor, llr_dev, n_cw, 0, 0, 0, info_dev, info_stride, ok_dev, iters_dev);
}

int ria::ldpc_launch(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor, const float* llr_dev, int64_t n_cw,
                     int frame_mode, int soft_stride, int step,
                     uint8_t* info_dev, int info_stride, uint8_t* ok_dev, int32_t* iters_dev) {
    using namespace ria;
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const LdpcCodeDev* t = nullptr;
    int rc = ldpc_tables_dev(ctx, rate, &t);
    if (rc != RIA_OK) return rc;
    if (info_stride < (t->k + 7) / 8) return set_error(ctx, RIA_E_INVAL, "ldpc: info_stride too small");

    // warps (= codewords) per CTA: whatever keeps the most codewords resident per SM (settled once per rate)
    auto kern = ldpc_decode_kernel;
    LdpcCodeDev* tm = &ctx->ldpc[rate];
    if (tm->launch_warps == 0) {
        int W = 0, best_warps = 0;
        for (int w = 4; w <= kMaxWarpsPerCta; ++w) {
            const size_t need = ldpc_smem_bytes(t->k, t->m, t->dv_max, w) + 1024;      // + per-CTA reservation
            if (need > ctx->smem_optin + 1024) break;
            const int ctas = static_cast<int>(ctx->smem_per_sm / need);
            const int warps = ctas * w > 48 ? 48 : ctas * w;
            if (warps > best_warps) { best_warps = warps; W = w; }
        }
        if (W == 0) return set_error(ctx, RIA_E_UNSUPPORTED, "ldpc: kernel does not fit in shared memory");
        const size_t smem = ldpc_smem_bytes(t->k, t->m, t->dv_max, W);
        // one attribute for all rates: the largest opt-in size
        RIA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
        RIA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int ctas_per_sm = 0;
        RIA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, W * 32, smem));
        if (ctas_per_sm < 1) return set_error(ctx, RIA_E_UNSUPPORTED, "ldpc: kernel does not fit (smem %zu)", smem);
        tm->launch_warps = W; tm->launch_ctas_per_sm = ctas_per_sm; tm->launch_smem = smem;
    }
    const int W = tm->launch_warps, ctas_per_sm = tm->launch_ctas_per_sm;
    const size_t smem = tm->launch_smem;
    long long want = (n_cw + W - 1) / W;
    long long grid = static_cast<long long>(ctx->sm_count) * ctas_per_sm;
    if (grid > want) grid = want;
    RIA_CUDA(ctx, cudaMemsetAsync(ctx->work_counter, 0, sizeof(unsigned int), ctx->stream));
    int inv_step = 0;
    if (step > 0)
        for (int x = 1; x < kN; ++x) if ((static_cast<long long>(x) * step) % kN == 1) { inv_step = x; break; }
    const int vec_ok = frame_mode && (soft_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(llr_dev) & 15) == 0);
    LdpcGather gather{frame_mode, soft_stride, step, inv_step, vec_ok};
    time_begin(ctx, KK_LDPC);
    kern<<<static_cast<unsigned>(grid), W * 32, smem, ctx->stream>>>(
        llr_dev, n_cw, gather, t->chk_var, t->var_slot, t->k, t->m, t->dv_max, max_iter, min_sum_factor,
        info_dev, info_stride, ok_dev, iters_dev, ctx->work_counter);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}

extern "C" int ria_ldpc_decode_batch_host(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor,
                                          const float* llr, int64_t n_cw,
                                          uint8_t* info, int info_stride, uint8_t* ok, int32_t* iters) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_cw < 0) return set_error(ctx, RIA_E_INVAL, "ldpc: negative size");
    if (n_cw == 0) return RIA_OK;
    if (!llr || !info || !ok || !iters) return set_error(ctx, RIA_E_INVAL, "ldpc: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    // Double-buffered chunks: H2D(c+1) overlaps decode(c) overlaps D2H(c-1).
    const int64_t chunk = n_cw < 32768 ? n_cw : 32768;
    const size_t in_b = static_cast<size_t>(chunk) * kN * sizeof(float);
    const size_t out_b = static_cast<size_t>(chunk) * (info_stride + 1 + 4);
    for (int b = 0; b < 2; ++b) {
        int rc = ensure_stage(ctx, b, in_b + out_b + 64, 0);
        if (rc != RIA_OK) return rc;
    }
    cudaStream_t s = ctx->stream;
    int buf = 0;
    for (int64_t off = 0; off < n_cw; off += chunk, buf ^= 1) {
        const int64_t n = (n_cw - off < chunk) ? (n_cw - off) : chunk;
        unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[buf]);
        float* d_llr = reinterpret_cast<float*>(base);
        int32_t* d_it = reinterpret_cast<int32_t*>(base + in_b);
        uint8_t* d_info = base + in_b + static_cast<size_t>(chunk) * 4;
        uint8_t* d_ok = d_info + static_cast<size_t>(chunk) * info_stride;
        // the previous use of this buffer (two chunks ago) must have drained
        RIA_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[buf]));
        RIA_CUDA(ctx, cudaMemcpyAsync(d_llr, llr + off * kN, static_cast<size_t>(n) * kN * sizeof(float),
                                      cudaMemcpyHostToDevice, s));
        int rc = ria_ldpc_decode_batch_dev(ctx, rate, max_iter, min_sum_factor, d_llr, n,
                                           d_info, info_stride, d_ok, d_it);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(info + off * info_stride, d_info, static_cast<size_t>(n) * info_stride,
                                      cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(ok + off, d_ok, static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(iters + off, d_it, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[buf], s));
    }
    RIA_CUDA(ctx, cudaStreamSynchronize(s));
    return RIA_OK;
}
