// On-device transmit synthesis for OFDM data frames, sample-identical to the
// reference transmitter so that a benchmark or sweep can give every frame its own payload without
// any host-side signal generation:
//
//   ria_encode_fixed_frame_batch_dev   v2::encodeFixedFrame (src/protocol/frame_v2.cpp:1285-1328):
//       pad to 4 x bytes_per_cw, LDPCEncoder::encode per codeword (systematic, parity bit i = XOR of
//       the info bits of check i: the parity part of H is the identity, ldpc_encoder.cpp:70-129,
//       193-257), ChannelInterleaver::interleave (ldpc_decoder.cpp:600-615), FrameInterleaver::interleave
//       (frame_interleaver.cpp:37-94)
//   ria_ofdm_tx_frames_dev             OFDMModulator::generateTrainingSymbols(n) + modulate(data, mod)
//       (src/ofdm/modulator.cpp:528-582, 348-477): bit mapping / differential encoding,
//       createOFDMSymbol (:217-270: carriers + pilots, FFT::inverse, cyclic prefix), complexToReal
//       (:272-283: x mixer phasor, real part, x output_scale), guard samples
//
// The inverse transform is the reference's radix-2 (src/dsp/fft.cpp:96-128: bit reversal, butterflies
// t = conj(w) * b; b = a - t; a = a + t in stage order, final scale 1/N) with the reference's twiddle
// values and no FMA contraction, so every sample has the reference's bits.  One CTA per frame; the
// transmitter is not on the receive hot path, the kernel is written for exactness, not for speed.

#include "ofdm_tables.h"
#include "rn_math.h"

namespace ria {
namespace {

constexpr int kTxThreads = 256;
constexpr int kMaxFft = 1024;

struct TxArgs {
    const uint8_t* coded; long long coded_stride; int coded_len;
    long long n_frames;
    float* out; long long out_stride;
    const float2* tw; const float2* nco; const OfdmCarrierTable* car;
    int N, logN, cp, guard, sym_len, n_train, n_data_sym, modulation, bpc;
    float scale;
    // OFDM_COX preamble (OFDMModulator::generatePreamble, modulator.cpp:479-532): pre_guard silent samples, then n_sts
    // Schmidl-Cox symbols (sync sequence on the even bins only, createSchmidlCoxSTS :298-330) before the LTS.
    // 0 / 0 for the chirp waveform.
    int pre_guard, n_sts;
};

__device__ __forceinline__ float2 cmul_ref(float2 a, float2 b) {       // std::complex<float> operator* (finite operands)
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// mapBits (modulator.cpp:76-108) and the constellation tables above it (:14-73)
__device__ float2 map_bits(unsigned bits, int mod) {
    switch (mod) {
        case RIA_BPSK: return make_float2((bits & 1) ? 1.0f : -1.0f, 0.0f);
        case RIA_QAM16: {
            const float lv[4] = {-3.f, -1.f, 3.f, 1.f};
            const float sc = 0.3162277660168379f;
            return make_float2(__fmul_rn(lv[(bits >> 2) & 3], sc), __fmul_rn(lv[bits & 3], sc));
        }
        case RIA_QAM32: {
            const float sc = 0.1961161351381840f;
            const float il[4] = {-3.f, -1.f, 1.f, 3.f};
            const int ig[4] = {0, 1, 3, 2};
            const float ql[8] = {-7.f, -5.f, -3.f, -1.f, 1.f, 3.f, 5.f, 7.f};
            const int qg[8] = {0, 1, 3, 2, 6, 7, 5, 4};
            const int qb = (bits >> 2) & 7, ib = bits & 3;
            int qi = 0, ii = 0;
            for (int i = 0; i < 4; ++i) if (ig[i] == ib) { ii = i; break; }
            for (int i = 0; i < 8; ++i) if (qg[i] == qb) { qi = i; break; }
            return make_float2(__fmul_rn(il[ii], sc), __fmul_rn(ql[qi], sc));
        }
        case RIA_QAM64: {
            const float lv[8] = {-7.f, -5.f, -1.f, -3.f, 7.f, 5.f, 1.f, 3.f};
            const float sc = 0.1543033499620919f;
            return make_float2(__fmul_rn(lv[(bits >> 3) & 7], sc), __fmul_rn(lv[bits & 7], sc));
        }
        case RIA_QAM256: {
            const float lv[16] = {-15.f, -13.f, -9.f, -11.f, -1.f, -3.f, -7.f, -5.f, 15.f, 13.f, 9.f, 11.f, 1.f, 3.f, 7.f, 5.f};
            const float sc = 0.0645497224367903f;
            return make_float2(__fmul_rn(lv[(bits >> 4) & 15], sc), __fmul_rn(lv[bits & 15], sc));
        }
        default: {      // QPSK and everything mapBits sends to QPSK_MAP
            const float s = 0.7071067811865476f;
            return make_float2((bits & 2) ? s : -s, (bits & 1) ? s : -s);
        }
    }
}

__global__ void __launch_bounds__(kTxThreads)
ofdm_tx_kernel(const TxArgs a) {
    __shared__ float2 fd[kMaxFft];
    __shared__ float2 prev[kMaxCarriers];
    __shared__ OfdmCarrierTable car;
    const int tid = threadIdx.x;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.car);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&car);
        for (int i = tid; i < static_cast<int>(sizeof(OfdmCarrierTable) / 4); i += kTxThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int N = a.N, nd = car.n_data, np = car.n_pilot;
    const int total_bits = a.coded_len * 8;
    const int n_sym = a.n_sts + a.n_train + a.n_data_sym;
    const int s_lts = a.n_sts, s_data = a.n_sts + a.n_train;

    for (long long f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
        const uint8_t* data = a.coded + f * a.coded_stride;
        float* out = a.out + f * a.out_stride;
        if (tid < nd) prev[tid] = make_float2(1.0f, 0.0f);          // generateTrainingSymbols: dbpsk_prev_symbols = 1 (:540)
        for (int i = tid; i < a.pre_guard; i += kTxThreads) out[i] = 0.0f;
        out += a.pre_guard;
        for (int s = 0; s < n_sym; ++s) {
            // mixer position of the symbol's first sample.  generatePreamble converts ONE STS and ONE LTS symbol to
            // passband and repeats the samples (modulator.cpp:513-530), so the mixer advances one symbol for the four
            // STS and one for the two LTS; generateTrainingSymbols converts every LTS symbol (:566-572).
            int nco_base = s * a.sym_len;
            if (a.n_sts > 0) nco_base = (s < s_lts) ? 0 : (s < s_data) ? a.sym_len : (2 + s - s_data) * a.sym_len;
            for (int i = tid; i < N; i += kTxThreads) fd[i] = make_float2(0.0f, 0.0f);
            __syncthreads();
            // ---- frequency domain, written at the bit-reversed position (fft.cpp:98-104) ----
            if (tid < nd) {
                float2 v = make_float2(0.0f, 0.0f);
                bool place = true;
                if (s < s_lts) {
                    v = car.tx_data[tid];                                // STS: even FFT bins only (:305-311)
                    place = (car.fft_idx[car.data_car[tid]] & 1) == 0;
                } else if (s < s_data) {
                    v = car.tx_data[tid];                                // sync_sequence[i % N] (:561-564)
                } else {
                    const int b0 = ((s - s_data) * nd + tid) * a.bpc;
                    if (b0 >= total_bits) {
                        place = true;                                    // padded with Complex(0, 0) (:455-458)
                    } else {
                        unsigned bits = 0;
                        for (int b = 0; b < a.bpc; ++b) {
                            bits <<= 1;
                            const int bi = b0 + b;
                            if (bi < total_bits) bits |= (data[bi >> 3] >> (7 - (bi & 7))) & 1u;
                        }
                        if (a.modulation == RIA_DBPSK) {
                            const float2 pc = (bits & 1) ? make_float2(-1.0f, 0.0f) : make_float2(1.0f, 0.0f);
                            v = cmul_ref(prev[tid], pc); prev[tid] = v;
                        } else if (a.modulation == RIA_DQPSK) {
                            const float2 ph[4] = {make_float2(1.f, 0.f), make_float2(0.f, 1.f), make_float2(-1.f, 0.f), make_float2(0.f, -1.f)};
                            v = cmul_ref(prev[tid], ph[bits & 3]); prev[tid] = v;
                        } else if (a.modulation == RIA_D8PSK) {
                            // 45-degree steps with a 22.5-degree offset (modulator.cpp:436-445), cosf / sinf as glibc
                            const float pi = 3.14159265358979f;
                            const float angle = __fadd_rn(__fmul_rn(static_cast<float>(bits & 7), pi / 4.0f), pi / 8.0f);
                            v = cmul_ref(prev[tid], make_float2(glibc_cosf(angle), glibc_sinf(angle))); prev[tid] = v;
                        } else {
                            v = map_bits(bits, a.modulation);
                        }
                    }
                }
                if (place) fd[__brev(static_cast<unsigned>(car.fft_idx[car.data_car[tid]])) >> (32 - a.logN)] = v;
            } else if (tid >= 64 && tid - 64 < np && s >= s_lts) {
                const int i = tid - 64;
                fd[__brev(static_cast<unsigned>(car.fft_idx[car.pilot_car[i]])) >> (32 - a.logN)] = make_float2(car.pilot_sign[i], 0.0f);
            }
            __syncthreads();
            // ---- radix-2 butterflies in the reference's stage order (fft.cpp:107-120), inverse: w = conj(W) ----
            for (int len = 2, step = N >> 1; len <= N; len <<= 1, step >>= 1) {
                const int half = len >> 1;
                for (int b = tid; b < (N >> 1); b += kTxThreads) {
                    const int k = b & (half - 1);
                    const int i0 = ((b - k) << 1) + k, i1 = i0 + half;
                    float2 w = a.tw[k * step]; w.y = -w.y;
                    const float2 x0 = fd[i0];
                    const float2 t = cmul_ref(w, fd[i1]);
                    fd[i1] = make_float2(__fsub_rn(x0.x, t.x), __fsub_rn(x0.y, t.y));
                    fd[i0] = make_float2(__fadd_rn(x0.x, t.x), __fadd_rn(x0.y, t.y));
                }
                __syncthreads();
            }
            // ---- 1/N (fft.cpp:123-128), cyclic prefix (:256-266), mixer and output scale (:272-283), guard (:469-472) ----
            const float inv_n = 1.0f / static_cast<float>(N);
            for (int i = tid; i < a.sym_len; i += kTxThreads) {
                float o = 0.0f;
                if (i < a.cp + N) {
                    const int src = (i < a.cp) ? N - a.cp + i : i - a.cp;
                    const float2 z = make_float2(__fmul_rn(fd[src].x, inv_n), __fmul_rn(fd[src].y, inv_n));
                    const float2 m = a.nco[nco_base + i];
                    o = __fmul_rn(__fsub_rn(__fmul_rn(z.x, m.x), __fmul_rn(z.y, m.y)), a.scale);
                }
                out[s * a.sym_len + i] = o;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// encodeFixedFrame
// ---------------------------------------------------------------------------------------------
struct EncArgs {
    const uint8_t* frames; long long frame_stride; int frame_len;
    long long n_frames;
    uint8_t* coded;                     // [n][324]
    const uint16_t* chk_var;            // [m][8]
    int k, m, bpc, step;
};

__global__ void __launch_bounds__(128)
encode_fixed_frame_kernel(const EncArgs a) {
    __shared__ uint8_t cw_bits[4][RIA_LDPC_N];      // coded bits of the four codewords, after the channel interleaver
    __shared__ uint8_t info[4][544];
    __shared__ uint8_t frame_bits[4 * RIA_LDPC_N];
    const int tid = threadIdx.x;
    for (long long f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
        const uint8_t* fr = a.frames + f * a.frame_stride;
        // info bits MSB-first; bytes past the frame are zero padding (:1292-1297); the k - 8*bpc tail bits are zero
        for (int i = tid; i < 4 * a.k; i += blockDim.x) {
            const int c = i / a.k, b = i - c * a.k;
            uint8_t bit = 0;
            if (b < a.bpc * 8) {
                const int byte = c * a.bpc + (b >> 3);
                if (byte < a.frame_len) bit = (fr[byte] >> (7 - (b & 7))) & 1u;
            }
            info[c][b] = bit;
        }
        __syncthreads();
        // systematic codeword: info bits, then parity bit i = XOR of the info bits of check i
        for (int i = tid; i < 4 * RIA_LDPC_N; i += blockDim.x) {
            const int c = i / RIA_LDPC_N, p = i - c * RIA_LDPC_N;
            uint8_t bit;
            if (p < a.k) bit = info[c][p];
            else {
                const uint16_t* row = a.chk_var + static_cast<size_t>(p - a.k) * 8;
                const int cnt = row[7];
                unsigned x = 0;
                for (int d = 0; d < 6; ++d) if (d < cnt) x ^= info[c][row[d]];
                bit = static_cast<uint8_t>(x & 1u);
            }
            // ChannelInterleaver::interleave: interleaved[(p * step) % 648] = bits[p]
            const int q = a.step ? (p * a.step) % RIA_LDPC_N : p;
            cw_bits[c][q] = bit;
        }
        __syncthreads();
        // FrameInterleaver::interleave: frame[4 b + (c + b) % 4] = cw[c][b]
        for (int i = tid; i < 4 * RIA_LDPC_N; i += blockDim.x) {
            const int c = i / RIA_LDPC_N, b = i - c * RIA_LDPC_N;
            frame_bits[4 * b + ((c + b) & 3)] = cw_bits[c][b];
        }
        __syncthreads();
        for (int i = tid; i < 4 * RIA_LDPC_N / 8; i += blockDim.x) {
            unsigned v = 0;
            for (int b = 0; b < 8; ++b) v = (v << 1) | frame_bits[8 * i + b];
            a.coded[f * (4 * RIA_LDPC_N / 8) + i] = static_cast<uint8_t>(v);
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace ria

extern "C" int ria_ofdm_tx_frame_samples(const ria_modem_config* cfg, int32_t coded_len) {
    using namespace ria;
    if (!cfg || coded_len < 0 || ofdm_config_error(*cfg)) return RIA_E_INVAL;
    const int bps = ria_ofdm_data_carriers(cfg) * ofdm_bits_per_carrier(cfg->modulation);
    const int n_data_sym = (coded_len * 8 + bps - 1) / bps;
    return (static_cast<int>(cfg->training_symbols) + n_data_sym) * ofdm_symbol_samples(*cfg);
}

namespace ria {
namespace {
int ofdm_tx_launch(ria_ctx* ctx, const ria_modem_config* cfg, const uint8_t* coded_dev, int64_t coded_stride, int32_t coded_len,
                   int64_t n_frames, float* samples_dev, int64_t out_stride, bool cox) {
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || coded_len <= 0 || coded_stride < coded_len) return set_error(ctx, RIA_E_INVAL, "ofdm tx: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!coded_dev || !samples_dev) return set_error(ctx, RIA_E_INVAL, "ofdm tx: null buffer");
    if (const char* err = ofdm_config_error(*cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: %s", err);
    if (cox && cfg->symbol_guard != 0) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm cox tx: symbol_guard must be 0");
    const int frame_len = cox ? ria_ofdm_cox_tx_frame_samples(cfg, coded_len) : ria_ofdm_tx_frame_samples(cfg, coded_len);
    if (out_stride < frame_len) return set_error(ctx, RIA_E_INVAL, "ofdm tx: out_stride %lld < %d samples per frame",
                                                 static_cast<long long>(out_stride), frame_len);
    const int N = static_cast<int>(cfg->fft_size);
    if (N > kMaxFft || (N & (N - 1))) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm tx: fft_size must be a power of two <= 1024");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const OfdmTablesDev* t = nullptr;
    int rc = ofdm_tables_dev(ctx, *cfg, frame_len, &t);
    if (rc != RIA_OK) return rc;
    TxArgs a{};
    a.coded = coded_dev; a.coded_stride = coded_stride; a.coded_len = coded_len; a.n_frames = n_frames;
    a.out = samples_dev; a.out_stride = out_stride;
    a.tw = t->twiddle_nat; a.nco = t->nco; a.car = t->car;
    a.N = N; a.logN = 0; while ((1 << a.logN) < N) ++a.logN;
    a.cp = t->cp; a.guard = static_cast<int>(cfg->symbol_guard); a.sym_len = t->sym_len;
    a.n_train = static_cast<int>(cfg->training_symbols);
    a.pre_guard = cox ? N + t->cp : 0;                  // one preamble symbol of silence (modulator.cpp:503-504)
    a.n_sts = cox ? 4 : 0;
    a.n_data_sym = (frame_len - a.pre_guard) / t->sym_len - a.n_train - a.n_sts;
    a.modulation = static_cast<int>(cfg->modulation); a.bpc = ofdm_bits_per_carrier(cfg->modulation);
    a.scale = 40.0f;                                    // ModemConfig::output_scale (include/ultra/types.hpp:235)
    long long grid = static_cast<long long>(ctx->sm_count) * 8;
    if (grid > n_frames) grid = n_frames;
    ofdm_tx_kernel<<<static_cast<unsigned>(grid), kTxThreads, 0, ctx->stream>>>(a);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
}  // namespace
}  // namespace ria

extern "C" int ria_ofdm_tx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                      const uint8_t* coded_dev, int64_t coded_stride, int32_t coded_len,
                                      int64_t n_frames, float* samples_dev, int64_t out_stride) {
    return ria::ofdm_tx_launch(ctx, cfg, coded_dev, coded_stride, coded_len, n_frames, samples_dev, out_stride, false);
}

extern "C" int ria_ofdm_cox_tx_frame_samples(const ria_modem_config* cfg, int32_t coded_len) {
    // guard + 4 STS + the chirp waveform's frame (2 LTS + data): modulator.cpp:479-532
    const int body = ria_ofdm_tx_frame_samples(cfg, coded_len);
    if (body < 0) return body;
    return body + 5 * (static_cast<int>(cfg->fft_size) + ria::ofdm_cyclic_prefix(*cfg));
}

extern "C" int ria_ofdm_cox_tx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                          const uint8_t* coded_dev, int64_t coded_stride, int32_t coded_len,
                                          int64_t n_frames, float* samples_dev, int64_t out_stride) {
    return ria::ofdm_tx_launch(ctx, cfg, coded_dev, coded_stride, coded_len, n_frames, samples_dev, out_stride, true);
}

extern "C" int ria_encode_fixed_frame_batch_dev(ria_ctx* ctx, int rate, int use_channel_interleave, int bits_per_symbol,
                                                const uint8_t* frames_dev, int64_t frame_stride, int32_t frame_len,
                                                int64_t n_frames, uint8_t* coded_dev) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len < 0 || frame_stride < frame_len) return set_error(ctx, RIA_E_INVAL, "encode: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!frames_dev || !coded_dev) return set_error(ctx, RIA_E_INVAL, "encode: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const LdpcCodeDev* t = nullptr;
    int rc = ldpc_tables_dev(ctx, rate, &t);
    if (rc != RIA_OK) return rc;
    if (t->k > 544) return set_error(ctx, RIA_E_UNSUPPORTED, "encode: rate %d not supported", rate);
    int step = 0;
    if (use_channel_interleave) {
        if (bits_per_symbol <= 0) return set_error(ctx, RIA_E_INVAL, "encode: bits_per_symbol must be > 0");
        step = channel_interleaver_step(bits_per_symbol, RIA_LDPC_N);
    }
    EncArgs a{};
    a.frames = frames_dev; a.frame_stride = frame_stride; a.frame_len = frame_len; a.n_frames = n_frames;
    a.coded = coded_dev; a.chk_var = t->chk_var; a.k = t->k; a.m = t->m; a.bpc = t->k / 8; a.step = step;
    long long grid = static_cast<long long>(ctx->sm_count) * 8;
    if (grid > n_frames) grid = n_frames;
    encode_fixed_frame_kernel<<<static_cast<unsigned>(grid), 128, 0, ctx->stream>>>(a);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
