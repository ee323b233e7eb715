// On-device transmit synthesis for MC-DPSK frame bodies, sample-identical to
// MultiCarrierDPSKModulator (src/psk/multi_carrier_dpsk.hpp:141-275):
//     [generateTrainingSequence][generateReferenceSymbol][modulate(data)]
// i.e. what the transmitter sends after the sync preamble and what IWaveform::process is handed.
//
// Everything that does not depend on the data is evaluated once per configuration on the HOST with
// the expressions of the reference (std::polar on float arguments = the container's libm, whose
// large-argument sinf/cosf the device does not restate): the carrier phasors polar(1, i * phase_inc),
// the differential phasors polar(1, phase_change), and the training + reference samples.  The device
// does the data-dependent part: per carrier the differential recurrence
//     current = prev * diff;  current /= abs(current)            (:245-248, a true recurrence)
// walked by one thread per carrier, then every sample of a data symbol as the ordered sum over the
// carriers of Re(current * carrier) / num_carriers (:254-259), written spreading-factor times.

#include "ria_internal.h"

#include <cmath>
#include <complex>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

struct McdpskTxTablesDev {
    ria_mcdpsk_config cfg{};
    float2* carrier = nullptr;     // [C][sps]  polar(1, i * phase_inc_c)
    float* head = nullptr;         // [(training + 1) * sps] training sequence + reference symbol
    float2 diff[4]{};              // polar(1, phase_change) per symbol value
};

void mcdpsk_tx_tables_free(McdpskTxTablesDev* t) {
    if (!t) return;
    if (t->carrier) cudaFree(t->carrier);
    if (t->head) cudaFree(t->head);
    delete t;
}

namespace {

constexpr int kMaxCar = 16;
constexpr int kTxThreads = 256;

struct TxArgs {
    const uint8_t* data; long long data_stride; int data_len;
    long long n_frames;
    float* out; long long out_stride;
    const float2* carrier; const float* head;
    float2 diff[4];
    int C, sps, bits, spread, head_len, n_ds;
};

__device__ __forceinline__ float cabs_hypotf(float2 a) {       // std::abs(complex<float>) = glibc hypotf (double inside)
    const double x = a.x, y = a.y;
    return static_cast<float>(sqrt(x * x + y * y));
}

__global__ void __launch_bounds__(kTxThreads)
mcdpsk_tx_kernel(const TxArgs a) {
    extern __shared__ float2 cur[];                 // [n_ds][C] normalised data symbols
    const int tid = threadIdx.x;
    const int C = a.C, sps = a.sps;
    const int total_bits = a.data_len * 8;
    for (long long f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
        const uint8_t* data = a.data + f * a.data_stride;
        float* out = a.out + f * a.out_stride;
        for (int i = tid; i < a.head_len; i += kTxThreads) out[i] = a.head[i];
        if (tid < C) {
            float2 prev = make_float2(1.0f, 0.0f);                       // generateReferenceSymbol (:186-187)
            for (int d = 0; d < a.n_ds; ++d) {
                int sym = 0;
                for (int b = 0; b < a.bits; ++b) {
                    const int bi = (d * C + tid) * a.bits + b;
                    const int bit = (bi < total_bits) ? ((data[bi >> 3] >> (7 - (bi & 7))) & 1) : 0;   // zero padding (:211)
                    sym = (sym << 1) | bit;
                }
                const float2 df = a.diff[sym];
                float2 c = make_float2(__fsub_rn(__fmul_rn(prev.x, df.x), __fmul_rn(prev.y, df.y)),
                                       __fadd_rn(__fmul_rn(prev.x, df.y), __fmul_rn(prev.y, df.x)));
                const float m = cabs_hypotf(c);
                c = make_float2(__fdiv_rn(c.x, m), __fdiv_rn(c.y, m));
                prev = c;
                cur[d * C + tid] = c;
            }
        }
        __syncthreads();
        const float fc = static_cast<float>(C);
        for (int p = tid; p < a.n_ds * sps; p += kTxThreads) {
            const int d = p / sps, i = p - d * sps;
            float acc = 0.0f;
            for (int c = 0; c < C; ++c) {
                const float2 s = cur[d * C + c];
                const float2 k = a.carrier[c * sps + i];
                const float re = __fsub_rn(__fmul_rn(s.x, k.x), __fmul_rn(s.y, k.y));
                acc = __fadd_rn(acc, __fdiv_rn(re, fc));
            }
            for (int rep = 0; rep < a.spread; ++rep)
                out[a.head_len + (d * a.spread + rep) * sps + i] = acc;
        }
        __syncthreads();
    }
}

int tx_tables(ria_ctx* ctx, const ria_mcdpsk_config& cfg, McdpskTxTablesDev** out) {
    for (McdpskTxTablesDev* t : ctx->mcdpsk_tx_tables)
        if (std::memcmp(&t->cfg, &cfg, sizeof cfg) == 0) { *out = t; return RIA_OK; }
    using Complex = std::complex<float>;
    const int C = static_cast<int>(cfg.num_carriers), sps = static_cast<int>(cfg.samples_per_symbol);
    const int T = static_cast<int>(cfg.training_symbols);
    // getCarrierFreqs (:68-79)
    std::vector<float> freqs(C);
    if (C == 1) freqs[0] = (cfg.freq_low + cfg.freq_high) / 2.0f;
    else { const float spacing = (cfg.freq_high - cfg.freq_low) / (C - 1); for (int i = 0; i < C; ++i) freqs[i] = cfg.freq_low + i * spacing; }
    std::vector<float2> carrier(static_cast<size_t>(C) * sps);
    for (int c = 0; c < C; ++c) {
        float phase_inc = 2.0f * M_PI * freqs[c] / cfg.sample_rate;
        for (int i = 0; i < sps; ++i) {
            float t = i * phase_inc;
            Complex k = std::polar(1.0f, t);
            carrier[static_cast<size_t>(c) * sps + i] = make_float2(k.real(), k.imag());
        }
    }
    // generateTrainingSequence (:141-176) + generateReferenceSymbol (:179-199)
    std::vector<float> head(static_cast<size_t>(T + 1) * sps, 0.0f);
    for (int sym = 0; sym < T; sym++) {
        for (int c = 0; c < C; c++) {
            float phase_offset = (c * sym) * M_PI / 2.0f;
            Complex training_sym = std::polar(1.0f, phase_offset);
            float phase_inc = 2.0f * M_PI * freqs[c] / cfg.sample_rate;
            for (int i = 0; i < sps; i++) {
                int idx = sym * sps + i;
                float t = i * phase_inc;
                Complex car = std::polar(1.0f, t);
                Complex modulated = training_sym * car;
                head[idx] += modulated.real() / C;
            }
        }
    }
    for (int c = 0; c < C; c++) {
        float phase_inc = 2.0f * M_PI * freqs[c] / cfg.sample_rate;
        Complex ref_sym(1.0f, 0.0f);
        for (int i = 0; i < sps; i++) {
            float t = i * phase_inc;
            Complex car = std::polar(1.0f, t);
            Complex modulated = ref_sym * car;
            head[static_cast<size_t>(T) * sps + i] += modulated.real() / C;
        }
    }
    McdpskTxTablesDev* t = new McdpskTxTablesDev();
    t->cfg = cfg;
    // modulate (:228-242): DQPSK 00=+45, 01=+135, 11=-135, 10=-45 degrees; DBPSK 0 / pi
    if (cfg.bits_per_symbol == 2) {
        static const float dqpsk_phases[] = {static_cast<float>(M_PI / 4), static_cast<float>(3 * M_PI / 4),
                                             static_cast<float>(-3 * M_PI / 4), static_cast<float>(-M_PI / 4)};
        for (int v = 0; v < 4; ++v) { Complex d = std::polar(1.0f, dqpsk_phases[v]); t->diff[v] = make_float2(d.real(), d.imag()); }
    } else {
        for (int v = 0; v < 2; ++v) {
            float phase_change = v ? M_PI : 0.0f;
            Complex d = std::polar(1.0f, phase_change);
            t->diff[v] = make_float2(d.real(), d.imag());
        }
    }
    ctx->mcdpsk_tx_tables.push_back(t);
    RIA_CUDA(ctx, cudaMalloc(&t->carrier, carrier.size() * sizeof(float2)));
    RIA_CUDA(ctx, cudaMalloc(&t->head, head.size() * sizeof(float)));
    RIA_CUDA(ctx, cudaMemcpy(t->carrier, carrier.data(), carrier.size() * sizeof(float2), cudaMemcpyHostToDevice));
    RIA_CUDA(ctx, cudaMemcpy(t->head, head.data(), head.size() * sizeof(float), cudaMemcpyHostToDevice));
    *out = t;
    return RIA_OK;
}

const char* tx_config_error(const ria_mcdpsk_config& c) {
    if (c.samples_per_symbol < 1 || c.samples_per_symbol > 4096) return "samples_per_symbol out of range";
    if (c.num_carriers < 1 || c.num_carriers > kMaxCar) return "num_carriers must be in [1, 16]";
    if (c.bits_per_symbol != 1 && c.bits_per_symbol != 2) return "bits_per_symbol must be 1 (DBPSK) or 2 (DQPSK)";
    if (c.spreading != 1 && c.spreading != 2 && c.spreading != 4) return "spreading must be 1, 2 or 4";
    if (c.training_symbols > 64) return "bad training_symbols";
    if (!(c.sample_rate > 0.0f)) return "sample_rate must be > 0";
    return nullptr;
}

}  // namespace
}  // namespace ria

extern "C" int ria_mcdpsk_tx_frame_samples(const ria_mcdpsk_config* cfg, int32_t data_len) {
    using namespace ria;
    if (!cfg || data_len < 0 || tx_config_error(*cfg)) return RIA_E_INVAL;
    const int bits_per_sym = static_cast<int>(cfg->num_carriers * cfg->bits_per_symbol);
    const int n_ds = (data_len * 8 + bits_per_sym - 1) / bits_per_sym;
    return static_cast<int>((cfg->training_symbols + 1 + n_ds * cfg->spreading) * cfg->samples_per_symbol);
}

extern "C" int ria_mcdpsk_tx_frames_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg,
                                        const uint8_t* data_dev, int64_t data_stride, int32_t data_len,
                                        int64_t n_frames, float* samples_dev, int64_t out_stride) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || data_len <= 0 || data_stride < data_len) return set_error(ctx, RIA_E_INVAL, "mcdpsk tx: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!data_dev || !samples_dev) return set_error(ctx, RIA_E_INVAL, "mcdpsk tx: null buffer");
    if (const char* err = tx_config_error(*cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "mcdpsk tx: %s", err);
    const int frame_len = ria_mcdpsk_tx_frame_samples(cfg, data_len);
    if (out_stride < frame_len) return set_error(ctx, RIA_E_INVAL, "mcdpsk tx: out_stride too small (%d samples per frame)", frame_len);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    McdpskTxTablesDev* t = nullptr;
    int rc = tx_tables(ctx, *cfg, &t);
    if (rc != RIA_OK) return rc;
    TxArgs a{};
    a.data = data_dev; a.data_stride = data_stride; a.data_len = data_len; a.n_frames = n_frames;
    a.out = samples_dev; a.out_stride = out_stride; a.carrier = t->carrier; a.head = t->head;
    for (int v = 0; v < 4; ++v) a.diff[v] = t->diff[v];
    a.C = static_cast<int>(cfg->num_carriers); a.sps = static_cast<int>(cfg->samples_per_symbol);
    a.bits = static_cast<int>(cfg->bits_per_symbol); a.spread = static_cast<int>(cfg->spreading);
    a.head_len = static_cast<int>((cfg->training_symbols + 1) * cfg->samples_per_symbol);
    const int bits_per_sym = a.C * a.bits;
    a.n_ds = (data_len * 8 + bits_per_sym - 1) / bits_per_sym;
    const size_t smem = static_cast<size_t>(a.n_ds) * a.C * sizeof(float2);
    if (smem > ctx->smem_optin) return set_error(ctx, RIA_E_UNSUPPORTED, "mcdpsk tx: frame too long for one CTA");
    RIA_CUDA(ctx, cudaFuncSetAttribute(mcdpsk_tx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
    long long grid = static_cast<long long>(ctx->sm_count) * 4;
    if (grid > n_frames) grid = n_frames;
    mcdpsk_tx_kernel<<<static_cast<unsigned>(grid), kTxThreads, smem, ctx->stream>>>(a);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
