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

#include "ria_internal.h"

#include <cfloat>

namespace ria {

namespace {

constexpr int kN = RIA_LDPC_N;
constexpr int kMaxWarpsPerCta = 16;

__device__ __forceinline__ float clamp50(float x) {
    // std::max(-50.0f, std::min(50.0f, x)) with libstdc++ semantics (NaN -> 50)
    float y = (x < 50.0f) ? x : 50.0f;
    return (-50.0f < y) ? y : -50.0f;
}

struct LdpcGather {
    int frame_mode;     // 0: llr_g is [n_cw][648];  1: llr_g is [n_frames][soft_stride], n_cw = 4 n_frames
    int soft_stride;
    int step;           // ChannelInterleaver step, 0 = no channel interleaving
    int inv_step;       // step^-1 mod 648 (0 = none: scalar gather), used by the coalesced load
    int vec_ok;         // rows are 16-byte aligned: one float4 per interleaver position
};

struct CheckIn {
    float v[7];      // v2c per slot (0..5 info, 6 identity)
    int cnt;         // number of info edges
};

// c2v for all 7 slots of one check from its v2c values.  Unused info slots (d >= cnt) are
// neutral: magnitude FLT_MAX, positive sign.
// ldpc_decoder.cpp:186-203 keeps (min1, min_idx, min2) with strict `<` updates and sends min2 to
// min_idx, min1 to everyone else.  Selecting by VALUE (|v_d| == min1 ? min2 : min1) is the same
// function: when the minimum is tied the strict updates leave min2 == min1, so every tied edge
// gets that value either way.  That turns the index bookkeeping into three min/max per edge.
__device__ __forceinline__ void check_update(const CheckIn& in, float factor, float (&out)[7]) {
    float m1 = FLT_MAX, m2 = FLT_MAX;
    float a[7];
    unsigned neg = 0;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const bool used = (d == 6) || (d < in.cnt);
        float x = fminf(fabsf(in.v[d]), FLT_MAX);   // `abs_msg < min_abs` never admits inf/NaN
        x = used ? x : FLT_MAX;
        a[d] = x;
        if (used && in.v[d] < 0.0f) neg ^= (1u << d) | 0x80u;   // bit 7 = running product
        m2 = fminf(m2, fmaxf(m1, x));
        m1 = fminf(m1, x);
    }
    const float s1 = __fmul_rn(m1, factor);
    const float s2 = __fmul_rn(m2, factor);
    const unsigned all_neg = (neg >> 7) & 1u;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const float mag = (a[d] == m1) ? s2 : s1;
        const unsigned sgn = all_neg ^ ((neg >> d) & 1u);
        out[d] = sgn ? -mag : mag;
    }
}

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

    for (;;) {
        long long cw;
        {
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(counter, 1u);
            cw = static_cast<long long>(__shfl_sync(0xffffffffu, t, 0));
        }
        if (cw >= n_cw) break;

        if (!gather.frame_mode) {
            // ---- load the codeword's 648 LLRs (162 x float4, coalesced, streaming) ----
            const float4* src = reinterpret_cast<const float4*>(llr_g + cw * kN);
            float4* dst = reinterpret_cast<float4*>(llr);
            for (int i = lane; i < kN / 4; i += 32) dst[i] = __ldcs(src + i);
        } else {
            // ---- fixed 4-codeword frame: codeword c of frame f, de-interleaved on the fly ----
            // FrameInterleaver::deinterleave (src/fec/frame_interleaver.cpp:37-47, 96-124):
            //   cw_soft[c][b] = frame_soft[4 b + (c + b) % 4]
            // ChannelInterleaver::deinterleave (src/fec/ldpc_decoder.cpp:600-625):
            //   out[p] = cw_soft[(p * step) % 648]
            const long long fr = cw >> 2;
            const int c = static_cast<int>(cw & 3);
            const float* src = llr_g + fr * gather.soft_stride;
            if (gather.vec_ok && (gather.step == 0 || gather.inv_step != 0)) {
                // Walk the frame in memory order: position b holds the four codewords' bit b as one
                // float4 (coalesced, and the four warps of a frame hit the same lines); this warp
                // keeps component (c + b) & 3 and scatters it to p = b * step^-1 mod 648.
                const float4* src4 = reinterpret_cast<const float4*>(src);
                const int inv = gather.step ? gather.inv_step : 1;
                int p = (lane * inv) % kN;
                const int dp = (32 * inv) % kN;
                for (int b = lane; b < kN; b += 32) {
                    const float4 v = __ldg(src4 + b);
                    const int sel = (c + b) & 3;
                    const float lo = (sel & 1) ? v.y : v.x, hi = (sel & 1) ? v.w : v.z;
                    llr[p] = (sel & 2) ? hi : lo;
                    p += dp;
                    if (p >= kN) p -= kN;
                }
            } else {
                for (int p = lane; p < kN; p += 32) {
                    const int b = gather.step ? (p * gather.step) % kN : p;
                    llr[p] = __ldg(src + 4 * b + ((c + b) & 3));
                }
            }
        }
        __syncwarp();
        for (int j = lane; j < k; j += 32) tot[j] = llr[j];
        for (int i = lane; i < m; i += 32) {
            msg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            msg[m + i] = make_float4(0.f, 0.f, 0.f, llr[k + i]);   // slot 7 = parity total
        }
        __syncwarp();

        int iters = max_iter;
        bool success = false;
        // it == max_iter is a parity-only pass over the totals of the last iteration.
        // The parity of iteration it-1 is seen at the start of pass `it`.  For it <= 3 -- where
        // clean frames converge -- it is a separate cheap pass, so a converged codeword does not
        // pay for a check update it will discard; later passes fold it into the update.
        for (int it = 0; it <= max_iter; ++it) {
            const bool fused_parity = it > 3;
            if (it >= 1 && !fused_parity) {
                bool bad = false;
                for (int i = lane; i < m; i += 32) {
                    const uint4 vi = chk_var[i];
                    const int cnt = vi.w >> 16;
                    const unsigned idx[6] = {vi.x & 0xFFFFu, vi.x >> 16, vi.y & 0xFFFFu,
                                             vi.y >> 16,     vi.z & 0xFFFFu, vi.z >> 16};
                    unsigned par = (reinterpret_cast<const float*>(msg + m + i)[3] < 0.0f) ? 1u : 0u;
#pragma unroll
                    for (int d = 0; d < 6; ++d)
                        if (d < cnt) par ^= (tot[idx[d]] < 0.0f) ? 1u : 0u;
                    bad |= (par != 0);
                }
                if (!__any_sync(0xffffffffu, bad)) { success = true; iters = it - 1; break; }
                if (it == max_iter) break;
            }
            // ================= phase A: checks =================
            bool bad = false;
            if (it < max_iter || fused_parity) {
                for (int i = lane; i < m; i += 32) {
                    const uint4 vi = chk_var[i];
                    const float4 c_lo = msg[i];
                    const float4 c_hi = msg[m + i];
                    const int cnt = vi.w >> 16;
                    const unsigned idx[6] = {vi.x & 0xFFFFu, vi.x >> 16, vi.y & 0xFFFFu,
                                             vi.y >> 16,     vi.z & 0xFFFFu, vi.z >> 16};
                    const float cold[7] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z};
                    CheckIn in;
                    in.cnt = cnt;
                    unsigned par = 0;
#pragma unroll
                    for (int d = 0; d < 6; ++d) {
                        float T = 0.0f;
                        if (d < cnt) T = tot[idx[d]];
                        par ^= (d < cnt && T < 0.0f) ? 1u : 0u;
                        in.v[d] = (it == 0) ? T : clamp50(__fsub_rn(T, cold[d]));
                    }
                    {
                        const float T = c_hi.w;                     // parity variable's total
                        par ^= (T < 0.0f) ? 1u : 0u;
                        in.v[6] = (it == 0) ? T : clamp50(__fsub_rn(T, cold[6]));
                    }
                    bad |= (par != 0);
                    if (it < max_iter) {
                        float out[7];
                        check_update(in, factor, out);
                        const float tp = __fadd_rn(llr[k + i], out[6]);
                        msg[i] = make_float4(out[0], out[1], out[2], out[3]);
                        msg[m + i] = make_float4(out[4], out[5], out[6], tp);
                    }
                }
            }
            if (fused_parity) {
                // totals examined in this pass belong to iteration it-1
                const bool any_bad = __any_sync(0xffffffffu, bad);
                if (!any_bad) { success = true; iters = it - 1; break; }
            }
            if (it == max_iter) break;
            __syncwarp();
            // ================= phase B: info variables =================
            for (int j = lane; j < k; j += 32) {
                float s = llr[j];
                const float* mf = reinterpret_cast<const float*>(msg);
                for (int d = 0; d < dv_max; ++d) {
                    const unsigned slot = var_slot[d * k + j];
                    if (slot == 0xFFFFu) break;
                    s = __fadd_rn(s, mf[slot]);
                }
                tot[j] = s;
            }
            __syncwarp();
        }

        // ---- outputs: pack info hard bits MSB-first (ldpc_decoder.cpp:240-257) ----
        // one ballot per 32 variables; bit-reversed it reads as 4 output bytes, MSB = lowest index
        for (int r = 0; 4 * r < info_stride; ++r) {        // bytes past ceil(k/8) come out as zero
            const int j = 32 * r + lane;
            const bool bit = (j < k) && (tot[j] < 0.0f);
            const unsigned word = __brev(__ballot_sync(0xffffffffu, bit));
            if (lane < 4 && 4 * r + lane < info_stride)
                info_g[cw * info_stride + 4 * r + lane] = static_cast<uint8_t>(word >> (24 - 8 * lane));
        }
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
    const int64_t chunk = 32768;
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
