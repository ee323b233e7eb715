// Device code shared by the LDPC kernels (ldpc.cu: first pass; ldpc_retry.cu: retry ladder):
// the frame/channel de-interleaving gather, the flooding min-sum decode of one codeword by one
// warp (src/fec/ldpc_decoder.cpp:154-260) and the MSB-first packing of the info bits.
#pragma once

#include "ria_internal.h"

#include <cfloat>

namespace ria {
namespace ldpc_core {

constexpr int kN = RIA_LDPC_N;

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


// Load one codeword's 648 LLRs into shared memory (all lanes of the warp; ends without a barrier).
__device__ __forceinline__ void gather_codeword(const float* llr_g, long long cw, const LdpcGather& gather,
                                                float* llr, int lane) {
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
}

// Per-codeword working set in shared memory and the code tables.
struct DecodeSmem {
    float* llr;             // [648] channel LLRs (input)
    float* tot;             // [k] info-variable totals
    float4* msg;            // [2][m] check messages + parity totals (see ldpc.cu header)
    const uint4* chk_var;   // [m]
    const uint16_t* var_slot;
    int k, m, dv_max;
};

// LDPCDecoder::decodeSoft for one codeword, one warp (flooding normalised min-sum, early exit on
// parity).  On return tot[] / msg slot 7 hold the totals the hard decision is taken from.
__device__ __forceinline__ void decode_codeword(float* llr, float* tot, float4* msg, const uint4* chk_var,
                                                const uint16_t* var_slot, int k, int m, int dv_max,
                                                int max_iter, float factor, int lane, bool& success, int& iters,
                                                bool try_shortcut = false, bool* took_shortcut = nullptr) {
    __syncwarp();
    // Clean codeword shortcut.  If the channel hard decisions already satisfy every check, each
    // check-to-variable message of the first iteration carries the sign of "the other bits' parity",
    // which is the variable's own sign, so no total changes sign: the reference stops after its first
    // iteration with exactly these bits and reports 0 iterations (ldpc_decoder.cpp:226-238).  The first
    // check / variable update is skipped; the result (bits, success, count) is the same.  max_iter = 0
    // runs no iteration and never succeeds (:170), so the shortcut needs max_iter >= 1.  The totals of a shortcut
    // codeword are its channel values: the caller packs the info bits from llr[] (tot[] is not written).
    if (took_shortcut) *took_shortcut = false;
    if (try_shortcut && max_iter >= 1) {
        bool bad = false;
        for (int i = lane; i < m; i += 32) {
            const uint4 vi = chk_var[i];
            const int cnt = vi.w >> 16;
            const unsigned idx[6] = {vi.x & 0xFFFFu, vi.x >> 16, vi.y & 0xFFFFu,
                                     vi.y >> 16,     vi.z & 0xFFFFu, vi.z >> 16};
            unsigned par = (llr[k + i] < 0.0f) ? 1u : 0u;
#pragma unroll
            for (int d = 0; d < 6; ++d)
                if (d < cnt) par ^= (llr[idx[d]] < 0.0f) ? 1u : 0u;
            bad |= (par != 0);
        }
        if (!__any_sync(0xffffffffu, bad)) {
            success = true; iters = 0;
            if (took_shortcut) *took_shortcut = true;
            __syncwarp();
            return;
        }
    }
    for (int j = lane; j < k; j += 32) tot[j] = llr[j];
    for (int i = lane; i < m; i += 32) {
        msg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        msg[m + i] = make_float4(0.f, 0.f, 0.f, llr[k + i]);   // slot 7 = parity total
    }
    __syncwarp();

    iters = max_iter;
    success = false;
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
}

// Pack the info hard bits MSB-first (ldpc_decoder.cpp:240-257): one ballot per 32 variables;
// bit-reversed it reads as 4 output bytes, MSB = lowest index.  Bytes past ceil(k/8) are zero.
__device__ __forceinline__ void pack_info(const float* tot, int k, uint8_t* info, int info_stride, int lane) {
    for (int r = 0; 4 * r < info_stride; ++r) {
        const int j = 32 * r + lane;
        const bool bit = (j < k) && (tot[j] < 0.0f);
        const unsigned word = __brev(__ballot_sync(0xffffffffu, bit));
        if (lane < 4 && 4 * r + lane < info_stride)
            info[4 * r + lane] = static_cast<uint8_t>(word >> (24 - 8 * lane));
    }
}

__host__ __device__ inline size_t ldpc_tab_bytes(int k, int m, int dv_max) {
    size_t tab = static_cast<size_t>(m) * 16 + static_cast<size_t>(dv_max) * k * 2;
    return (tab + 15) & ~size_t(15);
}

}  // namespace ldpc_core
}  // namespace ria
