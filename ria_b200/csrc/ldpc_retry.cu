// LDPC retry ladder of v2::decodeFixedFrame and robustDecodeSingleCW for sm_100a.
//
// Replaces the part of v2::decodeFixedFrame that runs when a codeword's first decode fails
// (src/protocol/frame_v2.cpp:1389-1546) and robustDecodeSingleCW
// (src/gui/modem/streaming_decoder.cpp:1028-1058):
//
//   phase 0   4 attempts   min-sum factors 0.875 / 0.75 / 0.625 / 0.5 on the unmodified soft bits
//   phase 1  15 attempts   soft bits + N(0, sigma) noise, factor per attempt
//   phase 2   5 attempts   clip to +-10, + noise, factor alternating 0.625 / 0.875
//   phase 3   3 attempts   0.5 x soft bits + noise
//   phase 4   3 attempts   clip to +-6, + noise
//   phase 5   5 attempts   hard decision (+-1) + noise
//   phase 6   3 attempts   0.25 x soft bits + noise
//
// Bit-exactness needs the reference's noise: std::mt19937 seeded with
// hash(first 16 soft bits) + f(attempt), fed to libstdc++'s std::normal_distribution<float>
// (Marsaglia polar on generate_canonical<float, 24>, with glibc's logf).  The generator is serial
// in the reference; here lane 0 runs the (inherently serial) seeding recurrence, the warp runs the
// twist 32 words at a time and evaluates 32 candidate pairs per step, and a ballot / prefix
// popcount assigns the accepted pairs to soft-bit positions in the reference's order.
//
// The reference reuses ONE decoder object for the four codewords of a frame and leaves its
// min-sum factor at 0.875 once a codeword has entered phase 1 (:1448, :1471), so the FIRST decode
// of the following codewords of that frame runs with 0.875 instead of 0.9375.  One warp therefore
// owns a failed frame and walks its codewords in order, carrying the factor; codewords whose
// first-pass result (factor 0.9375, ldpc.cu) is still valid are not decoded again.
//
// Mapping: the first pass (ldpc.cu) decodes every codeword; a compaction kernel lists the frames
// (or, for robustDecodeSingleCW, the codewords) that failed; persistent warps pull list entries
// from an atomic counter.  No host synchronisation: the list length stays on the device.

#include "ldpc_core.cuh"
#include "rn_math.h"

namespace ria {

namespace {

using ldpc_core::kN;
using ldpc_core::LdpcGather;

constexpr int kMaxRetryWarps = 16;
constexpr int kMtN = 624, kMtM = 397;
constexpr float kDefaultFactor = 0.9375f;

enum PerturbMode { PM_NONE = 0, PM_ADD, PM_CLIP10, PM_SCALE50, PM_CLIP6, PM_HARD, PM_SCALE25 };

struct Attempt { float factor; float sigma; int mode; unsigned seed_mul; unsigned seed_add; };

// frame_v2.cpp:1409-1542, attempt by attempt.  seed = data_hash + seed_mul * 997 + seed_add
// (phase 1: retry * 997 + retry * 31).
__constant__ Attempt kLadder[38] = {
    // phase 0 (:1409-1421)
    {0.875f, 0.f, PM_NONE, 0, 0}, {0.75f, 0.f, PM_NONE, 0, 0}, {0.625f, 0.f, PM_NONE, 0, 0}, {0.5f, 0.f, PM_NONE, 0, 0},
    // phase 1 (:1426-1449): sigmas1 / factors1, seed = hash + retry*997 + retry*31
    {0.75f, 0.3f, PM_ADD, 0, 0 * 31},   {0.625f, 0.7f, PM_ADD, 1, 1 * 31},  {0.875f, 0.3f, PM_ADD, 2, 2 * 31},
    {0.75f, 1.0f, PM_ADD, 3, 3 * 31},   {0.625f, 0.5f, PM_ADD, 4, 4 * 31},  {0.75f, 1.5f, PM_ADD, 5, 5 * 31},
    {0.5f, 0.3f, PM_ADD, 6, 6 * 31},    {0.625f, 2.0f, PM_ADD, 7, 7 * 31},  {0.875f, 0.5f, PM_ADD, 8, 8 * 31},
    {0.75f, 0.7f, PM_ADD, 9, 9 * 31},   {0.625f, 1.0f, PM_ADD, 10, 10 * 31}, {0.875f, 2.5f, PM_ADD, 11, 11 * 31},
    {0.75f, 0.3f, PM_ADD, 12, 12 * 31}, {0.5f, 1.5f, PM_ADD, 13, 13 * 31},  {0.625f, 0.5f, PM_ADD, 14, 14 * 31},
    // phase 2 (:1452-1472): seed = hash + (retry+15)*997 + 12345
    {0.625f, 0.3f, PM_CLIP10, 15, 12345}, {0.875f, 0.8f, PM_CLIP10, 16, 12345}, {0.625f, 1.5f, PM_CLIP10, 17, 12345},
    {0.875f, 2.5f, PM_CLIP10, 18, 12345}, {0.625f, 4.0f, PM_CLIP10, 19, 12345},
    // phase 3 (:1475-1492): seed = hash + (retry+20)*997 + 54321, factor stays 0.875
    {0.875f, 0.5f, PM_SCALE50, 20, 54321}, {0.875f, 1.5f, PM_SCALE50, 21, 54321}, {0.875f, 3.0f, PM_SCALE50, 22, 54321},
    // phase 4 (:1495-1513): seed = hash + (retry+23)*997 + 99999
    {0.875f, 0.5f, PM_CLIP6, 23, 99999}, {0.875f, 1.5f, PM_CLIP6, 24, 99999}, {0.875f, 3.0f, PM_CLIP6, 25, 99999},
    // phase 5 (:1516-1534): seed = hash + (retry+26)*997 + 33333
    {0.875f, 0.0f, PM_HARD, 26, 33333}, {0.875f, 0.2f, PM_HARD, 27, 33333}, {0.875f, 0.5f, PM_HARD, 28, 33333},
    {0.875f, 1.0f, PM_HARD, 29, 33333}, {0.875f, 1.5f, PM_HARD, 30, 33333},
    // phase 6 (:1537-1542): seed = hash + (retry+31)*997 + 77777
    {0.875f, 0.3f, PM_SCALE25, 31, 77777}, {0.875f, 1.0f, PM_SCALE25, 32, 77777}, {0.875f, 2.0f, PM_SCALE25, 33, 77777},
};

__device__ __forceinline__ float clampf(float x, float lim) {
    // std::max(-lim, std::min(lim, x))
    const float y = (x < lim) ? x : lim;
    return (-lim < y) ? y : -lim;
}

__device__ __forceinline__ float transform(float b, int mode) {
    switch (mode) {
        case PM_CLIP10:  return clampf(b, 10.0f);
        case PM_SCALE50: return __fmul_rn(b, 0.5f);
        case PM_CLIP6:   return clampf(b, 6.0f);
        case PM_HARD:    return (b >= 0.0f) ? 1.0f : -1.0f;
        case PM_SCALE25: return __fmul_rn(b, 0.25f);
        default:         return b;
    }
}

// std::mt19937::seed(value): the recurrence is serial (one lane)
__device__ __forceinline__ void mt_seed(unsigned* mt, unsigned seed) {
    unsigned x = seed;
    mt[0] = x;
    for (int i = 1; i < kMtN; ++i) {
        x = 1812433253u * (x ^ (x >> 30)) + static_cast<unsigned>(i);
        mt[i] = x;
    }
}

// _M_gen_rand: in-place twist, 32 words per step.  Word i needs the OLD words i, i+1 and (i < 227)
// i+397, or the NEW word i-227 (i >= 227) / NEW word 0 (i = 623); reading a step's operands before
// writing its results gives exactly the sequential semantics.
__device__ __forceinline__ void mt_twist(unsigned* mt, int lane) {
    for (int base = 0; base < kMtN; base += 32) {
        const int i = base + lane;
        unsigned v = 0;
        if (i < kMtN) {
            const int i1 = (i + 1 == kMtN) ? 0 : i + 1;
            const int im = (i + kMtM >= kMtN) ? i + kMtM - kMtN : i + kMtM;
            const unsigned y = (mt[i] & 0x80000000u) | (mt[i1] & 0x7fffffffu);
            v = mt[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (i < kMtN) mt[i] = v;
        __syncwarp();
    }
}

__device__ __forceinline__ unsigned mt_temper(unsigned y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// std::generate_canonical<float, 24>(mt19937): one draw, float(u) / 2^32, 1.0 -> nextafter(1, 0)
__device__ __forceinline__ float canonical(unsigned u) {
    const float r = __fmul_rn(__uint2float_rn(u), 2.3283064365386963e-10f);      // exact scaling by 2^-32
    return (r >= 1.0f) ? __uint_as_float(0x3f7fffffu) : r;
}

// llr[j] = transform(llr[j]) + noise_j for j < 648, noise from std::normal_distribution<float>(0, sigma)
// on std::mt19937(seed) exactly as `for (float& llr : v) llr = f(llr) + noise(rng)` consumes it
// (libstdc++ bits/random.tcc normal_distribution::operator(): polar method, second value saved).
// In place: every soft bit is read and written exactly once, by the lane that owns its accepted pair.
__device__ void perturb(float* llr, unsigned* mt, unsigned seed, float sigma, int mode, int lane) {
    if (lane == 0) mt_seed(mt, seed);
    __syncwarp();
    int n_acc = 0;                                  // accepted pairs so far (two normals each)
    while (n_acc < kN / 2) {
        mt_twist(mt, lane);                         // first draw after seeding regenerates the state
        for (int p0 = 0; p0 < kMtN / 2 && n_acc < kN / 2; p0 += 32) {
            const int p = p0 + lane;
            bool acc = false;
            float x = 0.f, y = 0.f, r2 = 1.f;
            if (p < kMtN / 2) {
                const float u1 = canonical(mt_temper(mt[2 * p]));
                const float u2 = canonical(mt_temper(mt[2 * p + 1]));
                // result_type(2.0) * aurng() - 1.0 : float product, double subtraction, back to float
                x = __double2float_rn(static_cast<double>(__fmul_rn(2.0f, u1)) - 1.0);
                y = __double2float_rn(static_cast<double>(__fmul_rn(2.0f, u2)) - 1.0);
                r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
                acc = !(r2 > 1.0f || r2 == 0.0f);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, acc);
            const int idx = n_acc + __popc(bal & ((1u << lane) - 1u));
            if (acc && idx < kN / 2) {
                const float mult = __fsqrt_rn(__fdiv_rn(__fmul_rn(-2.0f, glibc_logf(r2)), r2));
                // first call returns y * mult, the saved value x * mult serves the next call;
                // then ret * stddev + mean (mean = 0)
                const float n0 = __fadd_rn(__fmul_rn(__fmul_rn(y, mult), sigma), 0.0f);
                const float n1 = __fadd_rn(__fmul_rn(__fmul_rn(x, mult), sigma), 0.0f);
                llr[2 * idx] = __fadd_rn(transform(llr[2 * idx], mode), n0);
                llr[2 * idx + 1] = __fadd_rn(transform(llr[2 * idx + 1], mode), n1);
            }
            n_acc += __popc(bal);
        }
    }
    __syncwarp();
}

struct RetryArgs {
    const float* llr_g; LdpcGather gather;
    const int* list; const unsigned* list_len; unsigned* counter;
    const uint16_t* chk_var_g; const uint16_t* var_slot_g;
    int k, m, dv_max, max_iter;
    int per_frame;                       // 1: unit = frame of 4 codewords (decodeFixedFrame); 0: unit = codeword, phase 0 only
    uint8_t* info_g; int info_stride; uint8_t* ok_g; int32_t* iters_g;
    uint8_t* attempt_g;                  // optional [n_cw]: 0 = first decode, 1..38 = ladder attempt that succeeded, 255 = none
};

__global__ void __launch_bounds__(kMaxRetryWarps * 32)
ldpc_retry_kernel(const RetryArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.k, m = a.m, dv_max = a.dv_max;
    uint4* chk_var = reinterpret_cast<uint4*>(smem_raw);
    uint16_t* var_slot = reinterpret_cast<uint16_t*>(chk_var + m);
    const size_t tab_bytes = ldpc_core::ldpc_tab_bytes(k, m, dv_max);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpad = (k + 3) & ~3;
    // per warp: llr[648] | tot[k] | msg[2][m] -- the same footprint as the first-pass kernel, so as many
    // codewords stay resident.  The soft bits are re-gathered from global memory for every attempt (648
    // floats against a 50-80 iteration decode) instead of being kept in a second buffer, and the
    // mt19937 state lives in the message area, which the decoder re-initialises after the perturbation.
    const size_t per_warp = (static_cast<size_t>(kN) + kpad + static_cast<size_t>(m) * 8) * 4;
    float* llr = reinterpret_cast<float*>(smem_raw + tab_bytes + per_warp * warp);
    float* tot = llr + kN;
    float4* msg = reinterpret_cast<float4*>(tot + kpad);
    unsigned* mt = reinterpret_cast<unsigned*>(msg);             // 624 words <= 8 m words (m >= 108)
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.chk_var_g);
        for (int i = threadIdx.x; i < m; i += blockDim.x) chk_var[i] = src[i];
        const int nvs = dv_max * k;
        for (int i = threadIdx.x; i < nvs; i += blockDim.x) var_slot[i] = a.var_slot_g[i];
    }
    __syncthreads();
    const unsigned n_units = *a.list_len;

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_units) break;
        const long long unit = a.list[t];
        const int n_cw_unit = a.per_frame ? 4 : 1;
        float factor_state = kDefaultFactor;                 // the frame's decoder object (:1359-1361)
        for (int c = 0; c < n_cw_unit; ++c) {
            const long long cw = a.per_frame ? unit * 4 + c : unit;
            bool success = a.ok_g[cw] != 0;
            const bool first_pass_valid = (factor_state == kDefaultFactor);
            if (first_pass_valid && success) continue;       // decoded by the first pass, decoder untouched
            ldpc_core::gather_codeword(a.llr_g, cw, a.gather, llr, lane);
            __syncwarp();
            // data-dependent seed of the perturbations (:1391-1396)
            unsigned hash = 0;
            for (int j = 0; j < 16; ++j) hash ^= __float_as_uint(llr[j]) + 0x9e3779b9u + (hash << 6) + (hash >> 2);
            int iters = 0, hit = 255;
            if (!first_pass_valid) {
                // the decoder was left at 0.875 by an earlier codeword of this frame: its first decode
                // of this codeword runs with that factor (:1383-1385)
                ldpc_core::decode_codeword(llr, tot, msg, chk_var, var_slot, k, m, dv_max, a.max_iter, factor_state, lane, success, iters);
                if (lane == 0) a.iters_g[cw] = iters;
                if (success) {
                    ldpc_core::pack_info(tot, k, a.info_g + cw * a.info_stride, a.info_stride, lane);
                    hit = 0;
                }
            }
            if (!success) {
                const int n_attempts = a.per_frame ? 38 : 4;
                bool dirty = false;                              // llr no longer holds the received soft bits
                for (int at = 0; at < n_attempts && !success; ++at) {
                    const Attempt A = kLadder[at];
                    if (dirty) {
                        __syncwarp();
                        ldpc_core::gather_codeword(a.llr_g, cw, a.gather, llr, lane);
                        __syncwarp();
                        dirty = false;
                    }
                    if (A.mode != PM_NONE) {
                        perturb(llr, mt, hash + A.seed_mul * 997u + A.seed_add, A.sigma, A.mode, lane);
                        dirty = true;
                    }
                    ldpc_core::decode_codeword(llr, tot, msg, chk_var, var_slot, k, m, dv_max, a.max_iter, A.factor, lane, success, iters);
                    if (success) {
                        ldpc_core::pack_info(tot, k, a.info_g + cw * a.info_stride, a.info_stride, lane);
                        if (lane == 0) a.iters_g[cw] = iters;            // `iterations` only changes on success
                        hit = at + 1;
                    }
                    // decoder state left behind: phase 0 restores 0.9375 (:1420); every later phase
                    // ends at 0.875 (:1448, :1471) whether or not it succeeded
                    factor_state = (at < 4) ? kDefaultFactor : 0.875f;
                }
                if (!success && a.per_frame) factor_state = 0.875f;
            }
            if (lane == 0) {
                a.ok_g[cw] = success ? 1 : 0;
                if (a.attempt_g) a.attempt_g[cw] = static_cast<uint8_t>(hit);
            }
            __syncwarp();
        }
    }
}

// the soft bits ladder attempt `at` decodes, for n codewords (one warp each): observability / tests
__global__ void ladder_perturb_kernel(const float* __restrict__ llr_g, long long n_cw, int at, float* __restrict__ out_g) {
    __shared__ float llr[kN];
    __shared__ unsigned mt[kMtN];
    const int lane = threadIdx.x;
    for (long long cw = blockIdx.x; cw < n_cw; cw += gridDim.x) {
        for (int j = lane; j < kN; j += 32) llr[j] = llr_g[cw * kN + j];
        __syncwarp();
        unsigned hash = 0;
        for (int j = 0; j < 16; ++j) hash ^= __float_as_uint(llr[j]) + 0x9e3779b9u + (hash << 6) + (hash >> 2);
        const Attempt A = kLadder[at];
        if (A.mode != PM_NONE) perturb(llr, mt, hash + A.seed_mul * 997u + A.seed_add, A.sigma, A.mode, lane);
        __syncwarp();
        for (int j = lane; j < kN; j += 32) out_g[cw * kN + j] = llr[j];
        __syncwarp();
    }
}

// list the units (frames: any of 4 codewords failed; codewords: failed) for the retry kernel
__global__ void ldpc_fail_list_kernel(const uint8_t* __restrict__ ok, long long n_units, int per_frame,
                                      int* __restrict__ list, unsigned* __restrict__ list_len) {
    const long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (u >= n_units) return;
    bool fail;
    if (per_frame) {
        const uchar4 o = reinterpret_cast<const uchar4*>(ok)[u];
        fail = !(o.x && o.y && o.z && o.w);
    } else {
        fail = !ok[u];
    }
    if (fail) list[atomicAdd(list_len, 1u)] = static_cast<int>(u);
}

}  // namespace

// Retry pass over the results of ldpc_launch (same gather description).  n_units = frames
// (frame_mode, 4 codewords each) or codewords.  The fail list lives in `list_scratch` ([n_units] int).
int ldpc_retry_launch(ria_ctx* ctx, int rate, int max_iter, const float* llr_dev, int64_t n_units, int frame_mode,
                      int soft_stride, int step, uint8_t* info_dev, int info_stride, uint8_t* ok_dev,
                      int32_t* iters_dev, uint8_t* attempt_dev, int* list_scratch) {
    if (n_units > 0x7fffffffLL) return set_error(ctx, RIA_E_INVAL, "ldpc retry: too many units");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const LdpcCodeDev* t = nullptr;
    int rc = ldpc_tables_dev(ctx, rate, &t);
    if (rc != RIA_OK) return rc;
    if (frame_mode && (reinterpret_cast<uintptr_t>(ok_dev) & 3) != 0)
        return set_error(ctx, RIA_E_INVAL, "ldpc retry: ok buffer must be 4-byte aligned");
    const int kpad = (t->k + 3) & ~3;
    const size_t per_warp = (static_cast<size_t>(kN) + kpad + static_cast<size_t>(t->m) * 8) * 4;
    // warps (= frames in flight) per CTA: whatever keeps the most of them resident per SM
    const size_t tab = ldpc_core::ldpc_tab_bytes(t->k, t->m, t->dv_max);
    int W = 0, best = 0;
    for (int w = 2; w <= kMaxRetryWarps; ++w) {
        const size_t need = tab + w * per_warp + 1024;
        if (need > ctx->smem_optin + 1024) break;
        const int ctas = static_cast<int>(ctx->smem_per_sm / need);
        const int warps = ctas * w > 48 ? 48 : ctas * w;
        if (warps > best) { best = warps; W = w; }
    }
    if (W == 0) return set_error(ctx, RIA_E_UNSUPPORTED, "ldpc retry: kernel does not fit in shared memory");
    const size_t smem = tab + W * per_warp;
    auto kern = ldpc_retry_kernel;
    int ctas_per_sm = 0;
    RIA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
    RIA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RIA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, W * 32, smem));
    if (ctas_per_sm < 1) return set_error(ctx, RIA_E_UNSUPPORTED, "ldpc retry: kernel does not fit (smem %zu)", smem);
    unsigned* list_len = ctx->work_counter + 40;
    unsigned* counter = ctx->work_counter + 41;
    cudaStream_t s = ctx->stream;
    RIA_CUDA(ctx, cudaMemsetAsync(list_len, 0, 2 * sizeof(unsigned), s));
    int inv_step = 0;
    if (step > 0)
        for (int x = 1; x < kN; ++x) if ((static_cast<long long>(x) * step) % kN == 1) { inv_step = x; break; }
    const int vec_ok = frame_mode && (soft_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(llr_dev) & 15) == 0);
    RetryArgs a{};
    a.llr_g = llr_dev; a.gather = LdpcGather{frame_mode, soft_stride, step, inv_step, vec_ok};
    a.list = list_scratch; a.list_len = list_len; a.counter = counter;
    a.chk_var_g = t->chk_var; a.var_slot_g = t->var_slot;
    a.k = t->k; a.m = t->m; a.dv_max = t->dv_max; a.max_iter = max_iter;
    a.per_frame = frame_mode ? 1 : 0;
    a.info_g = info_dev; a.info_stride = info_stride; a.ok_g = ok_dev; a.iters_g = iters_dev; a.attempt_g = attempt_dev;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((n_units + threads - 1) / threads);
    time_begin(ctx, KK_LDPC_RETRY);
    if (attempt_dev) RIA_CUDA(ctx, cudaMemsetAsync(attempt_dev, 0, static_cast<size_t>(n_units) * (frame_mode ? 4 : 1), s));
    ldpc_fail_list_kernel<<<blocks, threads, 0, s>>>(ok_dev, n_units, a.per_frame, list_scratch, list_len);
    long long grid = static_cast<long long>(ctx->sm_count) * ctas_per_sm;
    const long long want = (n_units + W - 1) / W;
    if (grid > want) grid = want;
    ldpc_retry_kernel<<<static_cast<unsigned>(grid), W * 32, smem, s>>>(a);
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return RIA_OK;
}

}  // namespace ria

extern "C" int ria_ldpc_robust_decode_batch_dev(ria_ctx* ctx, int rate, const float* llr_dev, int64_t n_cw,
                                                uint8_t* info_dev, int info_stride,
                                                uint8_t* ok_dev, int32_t* iters_dev, uint8_t* attempt_dev) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_cw < 0) return set_error(ctx, RIA_E_INVAL, "ldpc: negative size");
    if (n_cw == 0) return RIA_OK;
    if (!llr_dev || !info_dev || !ok_dev || !iters_dev) return set_error(ctx, RIA_E_INVAL, "ldpc: null buffer");
    if ((reinterpret_cast<uintptr_t>(llr_dev) & 15) != 0)
        return set_error(ctx, RIA_E_INVAL, "ldpc: llr_dev must be 16-byte aligned");
    const int max_iter = recommended_ldpc_iterations(rate);
    int rc = ldpc_launch(ctx, rate, max_iter, 0.9375f, llr_dev, n_cw, 0, 0, 0, info_dev, info_stride, ok_dev, iters_dev);
    if (rc != RIA_OK) return rc;
    // fail list: context scratch (the caller's buffers are not touched beyond their documented sizes)
    rc = ensure_scratch(ctx, static_cast<size_t>(n_cw) * sizeof(int));
    if (rc != RIA_OK) return rc;
    return ldpc_retry_launch(ctx, rate, max_iter, llr_dev, n_cw, 0, 0, 0, info_dev, info_stride, ok_dev, iters_dev,
                             attempt_dev, static_cast<int*>(ctx->scratch));
}

extern "C" int ria_ldpc_ladder_perturb_dev(ria_ctx* ctx, const float* llr_dev, int64_t n_cw, int attempt, float* out_dev) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_cw < 0 || attempt < 1 || attempt > 38) return set_error(ctx, RIA_E_INVAL, "ladder: attempt must be 1..38");
    if (n_cw == 0) return RIA_OK;
    if (!llr_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "ladder: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = static_cast<unsigned>(n_cw < 4096 ? n_cw : 4096);
    ladder_perturb_kernel<<<grid, 32, 0, ctx->stream>>>(llr_dev, n_cw, attempt - 1, out_dev);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
