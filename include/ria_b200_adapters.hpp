// ria_b200_adapters.hpp -- C++ host adapters over the C ABI (ria_b200.h), header-only.
//
// They mirror the reference classes on the hot path with the same method names, argument meaning,
// ownership and error behaviour, so that a maintainer can swap them in behind the reference's own
// plugin seams:
//
//   ria::LDPCDecoder        <- ultra::LDPCDecoder              include/ultra/fec.hpp:48-81
//   ria::LDPCCodec          <- ultra::fec::LDPCCodec / ICodec  src/fec/ldpc_codec.hpp:38-105
//   ria::OFDMChirpRx        <- RX half of ultra::OFDMChirpWaveform (stand-alone, no reference headers needed)
//                                                              src/waveform/ofdm_chirp_waveform.cpp:79-105, 391-485
//   ria::OFDMChirpWaveform  <- ultra::OFDMChirpWaveform   } RIA_WITH_ULTRA only: real ultra::IWaveform objects
//   ria::MCDPSKWaveform     <- ultra::MCDPSKWaveform      } (src/waveform/waveform_interface.hpp:47-220) whose receive
//   ria::createWaveform     <- WaveformFactory::create    } half (detectSync / detectDataSync / process / getSoftBits /
//                                                           status getters) runs on the B200; the transmit half, the sizing
//                                                           getters and the capabilities are the reference's, inherited
//   ria::decodeFixedFrame   <- ultra::protocol::v2::decodeFixedFrame (soft bits in: the complete function with
//                              retry ladder and false-positive repair; samples in: demod + first pass)
//                                                              src/protocol/frame_v2.cpp:1335-1920
//
// Compile inside the reference tree with -DRIA_WITH_ULTRA to make LDPCCodec derive from
// ultra::fec::ICodec (then it can be returned by CodecFactory::create).  Without that macro the
// header depends on nothing but the C ABI and the standard library.
//
// Like the reference objects, an adapter instance is driven by one thread at a time.  Batch = 1
// calls go through the *_host entry points (H2D, kernel, D2H inside the call); the batched entry
// points of the C ABI are what a many-channel receiver should call directly.
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ria_b200.h"

#ifdef RIA_WITH_ULTRA
#include "fec/codec_interface.hpp"
#include "ultra/fec.hpp"
#include "waveform/mc_dpsk_waveform.hpp"
#include "waveform/ofdm_chirp_waveform.hpp"
#include "waveform/ofdm_cox_waveform.hpp"
#endif

namespace ria {

using Bytes = std::vector<uint8_t>;

// One shared context per process and device (created on first use).  ria_ctx_create fails when
// no B200 is usable; the adapters then throw std::runtime_error -- there is no CPU fallback.
class Context {
public:
    static Context& instance(int device = 0) {
        static Context ctx(device);
        return ctx;
    }
    ria_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc != RIA_OK) throw std::runtime_error(std::string("ria_b200: ") + ria_last_error(ctx_));
    }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ~Context() { ria_ctx_destroy(ctx_); }

private:
    explicit Context(int device) {
        if (ria_ctx_create(device, &ctx_) != RIA_OK || !ctx_)
            throw std::runtime_error("ria_b200: no usable B200 (ria_ctx_create failed); no CPU fallback");
    }
    ria_ctx* ctx_ = nullptr;
};

// ------------------------------------------------------------------------------------------------
// ultra::LDPCDecoder
// ------------------------------------------------------------------------------------------------
class LDPCDecoder {
public:
    explicit LDPCDecoder(int rate) { setRate(rate); }

    // ldpc_decoder.cpp:268-282
    Bytes decode(std::span<const uint8_t> coded_data) {
        std::vector<float> llrs;
        llrs.reserve(coded_data.size() * 8);
        for (uint8_t byte : coded_data)
            for (int b = 7; b >= 0; --b) llrs.push_back(((byte >> b) & 1) ? -6.0f : 6.0f);
        return decodeSoft(llrs);
    }

    // ldpc_decoder.cpp:284-429: single block (<= 648 LLRs, zero padded) or bit-level
    // concatenation of the k info bits of every 648-LLR block (+ a zero-padded partial block).
    Bytes decodeSoft(std::span<const float> llrs) {
        if (llrs.empty()) { last_success_ = false; return {}; }
        const size_t n = RIA_LDPC_N;
        const bool single = llrs.size() <= n;
        const size_t full = llrs.size() / n, rem = llrs.size() - full * n;
        const size_t blocks = single ? 1 : full + (rem ? 1 : 0);
        std::vector<float> padded(blocks * n, 0.0f);
        std::memcpy(padded.data(), llrs.data(), llrs.size() * sizeof(float));
        const int stride = (k_ + 7) / 8;
        std::vector<uint8_t> info(blocks * stride), ok(blocks);
        std::vector<int32_t> iters(blocks);
        Context& c = Context::instance();
        c.check(ria_ldpc_decode_batch_host(c.get(), rate_, max_iter_, factor_, padded.data(),
                                           static_cast<int64_t>(blocks), info.data(), stride, ok.data(), iters.data()));
        last_iters_ = iters.back();
        if (single) { last_success_ = ok[0] != 0; return info; }
        // a trailing partial block goes through decodeBP, which overwrites the flag (:401)
        if (rem) last_success_ = ok.back() != 0;
        else { last_success_ = true; for (size_t b = 0; b < full; ++b) last_success_ = last_success_ && ok[b]; }
        Bytes out;
        uint8_t byte = 0;
        int cnt = 0;
        for (size_t b = 0; b < blocks; ++b)
            for (int j = 0; j < k_; ++j) {
                const uint8_t bit = (info[b * stride + (j >> 3)] >> (7 - (j & 7))) & 1;
                byte = static_cast<uint8_t>((byte << 1) | bit);
                if (++cnt == 8) { out.push_back(byte); byte = 0; cnt = 0; }
            }
        if (cnt > 0) out.push_back(static_cast<uint8_t>(byte << (8 - cnt)));
        return out;
    }

    bool lastDecodeSuccess() const { return last_success_; }
    int lastIterations() const { return last_iters_; }
    void setRate(int rate) {
        int k = 0;
        if (ria_ldpc_params(rate, &k, nullptr, nullptr) != RIA_OK) throw std::invalid_argument("bad code rate");
        rate_ = rate; k_ = k;
    }
    int getRate() const { return rate_; }
    void setMaxIterations(int it) { max_iter_ = it; }
    void setMinSumFactor(float f) { factor_ = f; }

private:
    int rate_ = RIA_R1_2, k_ = 324;
    int max_iter_ = 50;        // Impl::max_iterations, ldpc_decoder.cpp:43
    float factor_ = 0.75f;     // Impl::min_sum_factor, ldpc_decoder.cpp:44
    bool last_success_ = false;
    int last_iters_ = 0;
};

// ------------------------------------------------------------------------------------------------
// ultra::fec::LDPCCodec (decode side of ICodec)
// ------------------------------------------------------------------------------------------------
#ifdef RIA_WITH_ULTRA
using DecodeResult = ultra::fec::DecodeResult;
#else
struct DecodeResult {
    bool success = false;
    Bytes data;
    int iterations = 0;
    float ber_estimate = 0.0f;
};
#endif

class LDPCCodec
#ifdef RIA_WITH_ULTRA
    : public ultra::fec::ICodec
#endif
{
public:
    static constexpr size_t CODEWORD_BITS = 648;
    static constexpr size_t CODEWORD_BYTES = 81;

    // ldpc_codec.hpp:86-96
    static int getRecommendedIterations(int rate) {
        switch (rate) {
            case RIA_R3_4: return 60;
            case RIA_R2_3: return 70;
            case RIA_R1_2: return 80;
            case RIA_R1_3: return 60;
            case RIA_R1_4: return 50;
            default: return 50;
        }
    }

    explicit LDPCCodec(int rate = RIA_R1_2)
        : rate_(rate), max_iterations_(getRecommendedIterations(rate)), decoder_(rate) {
        decoder_.setMaxIterations(max_iterations_);
    }

#ifdef RIA_WITH_ULTRA
#define RIA_OVERRIDE override
    explicit LDPCCodec(ultra::CodeRate rate) : LDPCCodec(static_cast<int>(rate)) {}
    void setRate(ultra::CodeRate rate) override { setRate(static_cast<int>(rate)); }
    ultra::CodeRate getRate() const override { return static_cast<ultra::CodeRate>(rate_); }
    // TX stays on the CPU: the reference encoder is used as is
    Bytes encode(const Bytes& data) override {
        ultra::LDPCEncoder enc(static_cast<ultra::CodeRate>(rate_));
        return enc.encode(ultra::ByteSpan(data.data(), data.size()));
    }
#else
#define RIA_OVERRIDE
#endif
    std::string getName() const RIA_OVERRIDE { return "802.11n LDPC"; }
    void setRate(int rate) {
        rate_ = rate;
        decoder_.setRate(rate);
        const int rec = getRecommendedIterations(rate);
        if (max_iterations_ != rec) { max_iterations_ = rec; decoder_.setMaxIterations(rec); }
    }
    int getRateValue() const { return rate_; }
    void setMaxIterations(int it) RIA_OVERRIDE { max_iterations_ = it; decoder_.setMaxIterations(it); }
    int getMaxIterations() const RIA_OVERRIDE { return max_iterations_; }

    std::pair<bool, Bytes> decode(const std::vector<float>& soft_bits) RIA_OVERRIDE {
        Bytes d = decoder_.decodeSoft(soft_bits);
        return {decoder_.lastDecodeSuccess(), std::move(d)};
    }
    DecodeResult decodeExtended(const std::vector<float>& soft_bits) RIA_OVERRIDE {
        DecodeResult r;
        r.data = decoder_.decodeSoft(soft_bits);
        r.success = decoder_.lastDecodeSuccess();
        r.iterations = decoder_.lastIterations();
        r.ber_estimate = r.success ? static_cast<float>(r.iterations) / (max_iterations_ * 10.0f) : 0.5f;
        return r;
    }
    size_t getCodewordBits() const RIA_OVERRIDE { return CODEWORD_BITS; }
    size_t getInfoBits() const RIA_OVERRIDE { int k = 0; ria_ldpc_params(rate_, &k, nullptr, nullptr); return static_cast<size_t>(k); }
    size_t getParityBits() const RIA_OVERRIDE { return CODEWORD_BITS - getInfoBits(); }
    size_t getCodewordBytes() const RIA_OVERRIDE { return CODEWORD_BYTES; }
    size_t getDataBytes() const RIA_OVERRIDE { return getInfoBits() / 8; }
    float getEffectiveRate() const RIA_OVERRIDE { return static_cast<float>(getInfoBits()) / static_cast<float>(CODEWORD_BITS); }
#undef RIA_OVERRIDE

private:
    int rate_;
    int max_iterations_;
    LDPCDecoder decoder_;
};

// ------------------------------------------------------------------------------------------------
// v2::decodeFixedFrame (first pass) for one frame
// ------------------------------------------------------------------------------------------------
struct CodewordStatus {
    std::vector<bool> decoded;       // 4 entries
    std::vector<Bytes> data;         // bytes_per_cw each (empty when the codeword failed)
    ria_frame_status status{};       // header fields / CRC flags
    bool allSuccess() const { for (bool b : decoded) if (!b) return false; return !decoded.empty(); }
};

// v2::decodeFixedFrame(interleaved_soft, rate, use_channel_deinterleave, bits_per_symbol)
// (frame_v2.cpp:1335-1920), same arguments and result: first pass, retry ladder (:1389-1546) and
// false-positive repair (:1558-1916).  Fewer than 2592 soft bits -> all four codewords failed (:1343-1345).
inline CodewordStatus decodeFixedFrame(const std::vector<float>& interleaved_soft, int rate,
                                       bool use_channel_deinterleave = false, size_t bits_per_symbol = 0) {
    int k = 0;
    if (ria_ldpc_params(rate, &k, nullptr, nullptr) != RIA_OK) throw std::invalid_argument("bad code rate");
    const int bpc = k / 8;
    CodewordStatus st;
    st.decoded.assign(4, false);
    st.data.resize(4);
    if (interleaved_soft.size() < 2592) return st;
    Bytes data(4 * bpc);
    Context& c = Context::instance();
    const int saved = ria_ctx_get_decode_flags(c.get());
    c.check(ria_ctx_set_decode_flags(c.get(), RIA_DECODE_FULL));
    const int rc = ria_frame_decode_batch_host(c.get(), rate, use_channel_deinterleave ? 1 : 0,
                                               static_cast<int>(bits_per_symbol), interleaved_soft.data(),
                                               static_cast<int32_t>(interleaved_soft.size()), 1, data.data(), &st.status);
    ria_ctx_set_decode_flags(c.get(), saved);
    c.check(rc);
    for (int i = 0; i < 4; ++i) {
        st.decoded[i] = st.status.cw_ok[i] != 0;
        if (st.decoded[i]) st.data[i].assign(data.begin() + i * bpc, data.begin() + (i + 1) * bpc);
    }
    return st;
}

// Batch = 1 convenience over a device round trip; a real receiver batches frames and calls
// ria_frame_decode_batch_dev / ria_ofdm_rx_frames_host directly.
inline CodewordStatus decodeFixedFrame(const ria_modem_config& cfg, int rate, bool use_channel_interleave,
                                       std::span<const float> frame_samples, float cfo_hz = 0.0f,
                                       float phase = 0.0f, float* snr_db = nullptr) {
    int k = 0;
    if (ria_ldpc_params(rate, &k, nullptr, nullptr) != RIA_OK) throw std::invalid_argument("bad code rate");
    const int bpc = k / 8;
    Bytes data(4 * bpc);
    CodewordStatus st;
    float snr = 0.f;
    Context& c = Context::instance();
    c.check(ria_ofdm_rx_frames_host(c.get(), &cfg, rate, use_channel_interleave ? 1 : 0, frame_samples.data(),
                                    static_cast<int64_t>(frame_samples.size()), static_cast<int32_t>(frame_samples.size()),
                                    &cfo_hz, &phase, 1, data.data(), &st.status, &snr));
    if (snr_db) *snr_db = snr;
    st.decoded.resize(4);
    st.data.resize(4);
    for (int i = 0; i < 4; ++i) {
        st.decoded[i] = st.status.cw_ok[i] != 0;
        if (st.decoded[i]) st.data[i].assign(data.begin() + i * bpc, data.begin() + (i + 1) * bpc);
    }
    return st;
}

// ------------------------------------------------------------------------------------------------
// RX half of ultra::OFDMChirpWaveform (IWaveform): configure / setFrequencyOffset /
// setAbsoluteTrainingPosition / process / getSoftBits / estimatedSNR / estimatedCFO /
// getFadingIndex / reset with the reference's semantics (ofdm_chirp_waveform.cpp:79-105, 391-485)
// ------------------------------------------------------------------------------------------------
class OFDMChirpRx {
public:
    OFDMChirpRx() { ria_modem_config_for(RIA_DQPSK, RIA_R1_2, &config_); }
    explicit OFDMChirpRx(const ria_modem_config& cfg) : config_(cfg) {}

    void configure(int mod, int rate) {
        switch (mod) {
            case RIA_DBPSK: case RIA_DQPSK: case RIA_D8PSK: case RIA_QPSK: case RIA_BPSK:
            case RIA_QAM16: case RIA_QAM32: case RIA_QAM64: break;
            default: mod = RIA_DQPSK;                        // :81-87
        }
        ria_modem_config_for(mod, rate, &config_);
    }
    void setFrequencyOffset(float cfo_hz) { cfo_hz_ = cfo_hz; }
    void setAbsoluteTrainingPosition(size_t pos) { abs_pos_ = pos; has_abs_pos_ = true; }
    int getSamplesPerSymbol() const { return ria_ofdm_symbol_samples(&config_); }

    bool process(std::span<const float> samples) {
        if (static_cast<int>(samples.size()) < getSamplesPerSymbol()) return false;
        const size_t ref = has_abs_pos_ ? abs_pos_ : training_start_;
        // :404-413, evaluated in double like the reference expression
        float ph = static_cast<float>(-2.0f * 3.14159265358979323846 * cfo_hz_ * ref / config_.sample_rate);
        while (ph > 3.14159265358979323846) ph -= 2.0f * 3.14159265358979323846;
        while (ph < -3.14159265358979323846) ph += 2.0f * 3.14159265358979323846;
        const int n_sym = static_cast<int>(samples.size()) / getSamplesPerSymbol();
        const int bits = ria_ofdm_data_carriers(&config_) * bitsPerCarrier();
        const int stride = ((n_sym > 2 ? n_sym - 2 : 0) * bits + 3) & ~3;
        std::vector<float> llr(static_cast<size_t>(stride > 4 ? stride : 4));
        int32_t n_llr = 0;
        float snr = 0, cfo = 0, fad = 0;
        Context& c = Context::instance();
        c.check(ria_ofdm_presynced_batch_host(c.get(), &config_, samples.data(), static_cast<int64_t>(samples.size()),
                                              static_cast<int32_t>(samples.size()), &cfo_hz_, &ph, 1, llr.data(),
                                              static_cast<int32_t>(llr.size()), &n_llr, &snr, &cfo, &fad));
        fading_ = fad;
        const bool ready = n_llr >= RIA_LDPC_N;
        if (ready) {
            llr.resize(static_cast<size_t>(n_llr));
            soft_bits_ = std::move(llr);
            last_snr_ = snr;
            cfo_hz_ = last_cfo_ = cfo;                       // CFO feedback (:447-455)
        }
        return ready;
    }
    std::vector<float> getSoftBits() { return std::move(soft_bits_); }   // moves out (:470-472)
    float estimatedSNR() const { return last_snr_; }
    float estimatedCFO() const { return std::fabs(last_cfo_) > 0.1f ? last_cfo_ : cfo_hz_; }
    float getFadingIndex() const { return fading_; }
    void reset() { soft_bits_.clear(); has_abs_pos_ = false; abs_pos_ = 0; }   // CFO is kept (:474-485)
    const ria_modem_config& config() const { return config_; }

private:
    int bitsPerCarrier() const {
        switch (config_.modulation) {
            case RIA_DQPSK: case RIA_QPSK: return 2;
            case RIA_D8PSK: case RIA_QAM8: return 3;
            case RIA_QAM16: return 4;
            case RIA_QAM32: return 5;
            case RIA_QAM64: return 6;
            case RIA_QAM256: return 8;
            default: return 1;
        }
    }
    ria_modem_config config_{};
    float cfo_hz_ = 0.0f, last_cfo_ = 0.0f, last_snr_ = 0.0f, fading_ = 0.0f;
    size_t training_start_ = 0, abs_pos_ = 0;
    bool has_abs_pos_ = false;
    std::vector<float> soft_bits_;
};


#ifdef RIA_WITH_ULTRA
// ------------------------------------------------------------------------------------------------
// ultra::IWaveform drop-ins.  They ARE the reference's waveform classes for everything that is not on the
// receive hot path (preamble / modulate, sizing, capabilities: inherited), and override every virtual that
// touches receive state so that the reference's own CPU demodulator / synchronisers are never run:
//   detectSync      -> ria_chirp_detect_dual_batch_host      (ofdm_chirp_waveform.cpp:163-205, mc_dpsk_waveform.cpp:177-224)
//   detectDataSync  -> ria_ofdm_data_sync_batch_host / ria_zc_detect_batch_host   (:207-384 / :227-292)
//   process         -> ria_ofdm_presynced_batch_host / ria_mcdpsk_process_batch_host   (:391-468 / :294-338)
// One instance is driven by one thread at a time, like the reference objects (SURVEY.md 8b "Threading").
// ------------------------------------------------------------------------------------------------
inline ria_modem_config toRiaConfig(const ultra::ModemConfig& m) {
    ria_modem_config c{};
    c.sample_rate = static_cast<uint32_t>(m.sample_rate);
    c.center_freq = static_cast<uint32_t>(m.center_freq);
    c.fft_size = static_cast<uint32_t>(m.fft_size);
    c.num_carriers = static_cast<uint32_t>(m.num_carriers);
    c.cp_mode = static_cast<uint32_t>(m.cp_mode);
    c.symbol_guard = static_cast<uint32_t>(m.symbol_guard);
    c.use_pilots = m.use_pilots ? 1u : 0u;
    c.pilot_spacing = static_cast<uint32_t>(m.pilot_spacing);
    c.modulation = static_cast<uint32_t>(m.modulation);
    c.training_symbols = 2;
    return c;
}

class OFDMChirpWaveform : public ultra::OFDMChirpWaveform {
public:
    OFDMChirpWaveform() : ultra::OFDMChirpWaveform() { cfg_ = toRiaConfig(getConfig()); ria_chirp_config_default(&chirp_); }
    explicit OFDMChirpWaveform(const ultra::ModemConfig& config) : ultra::OFDMChirpWaveform(config) {
        cfg_ = toRiaConfig(getConfig());
        ria_chirp_config_default(&chirp_);
    }

    void configure(ultra::Modulation mod, ultra::CodeRate rate) override {
        ultra::OFDMChirpWaveform::configure(mod, rate);          // TX side + the pilot layout rule (:79-105)
        cfg_ = toRiaConfig(getConfig());
    }
    void setFrequencyOffset(float cfo_hz) override { cfo_hz_ = cfo_hz; }
    float getFrequencyOffset() const override { return cfo_hz_; }

    bool detectSync(ultra::SampleSpan samples, ultra::SyncResult& result, float threshold = 0.15f) override {
        ria_sync_result r{};
        Context& c = Context::instance();
        c.check(ria_chirp_detect_dual_batch_host(c.get(), &chirp_, samples.data(), static_cast<int64_t>(samples.size()),
                                                 static_cast<int32_t>(samples.size()), threshold, 1, &r));
        result.detected = r.detected != 0;
        result.correlation = std::max(r.correlation, r.snr_estimate);           // max(up, down) (:173)
        result.cfo_hz = r.cfo_hz;
        result.has_training = true;
        if (result.detected) {
            synced_ = true;
            last_cfo_ = r.cfo_hz;
            const int chirp_samples = static_cast<int>(static_cast<size_t>(chirp_.sample_rate * chirp_.duration_ms / 1000.0f));
            const int gap_samples = static_cast<int>(static_cast<size_t>(cfg_.sample_rate * 100.0f / 1000.0f));
            result.start_sample = r.aux + chirp_samples + gap_samples;         // training after the down chirp (:194-196)
            training_start_ = static_cast<size_t>(result.start_sample);
        }
        return result.detected;
    }

    bool detectDataSync(ultra::SampleSpan samples, ultra::SyncResult& result, float known_cfo_hz = 0.0f,
                        float threshold = 0.3f) override {
        result.detected = false;
        result.correlation = 0.0f;
        result.cfo_hz = known_cfo_hz;
        result.has_training = true;
        if (samples.size() < static_cast<size_t>(getSamplesPerSymbol()) * 3) return false;
        ria_sync_result r{};
        Context& c = Context::instance();
        c.check(ria_ofdm_data_sync_batch_host(c.get(), &cfg_, samples.data(), static_cast<int64_t>(samples.size()),
                                              static_cast<int32_t>(samples.size()), &known_cfo_hz, threshold, 1, &r));
        result.correlation = r.correlation;
        burst_latched_ = burst_pending_ = false;                               // reset at every attempt (:356-358)
        if (r.detected) {
            result.detected = true;
            result.start_sample = r.start_sample;
            training_start_ = static_cast<size_t>(r.start_sample);
            synced_ = true;
            last_cfo_ = known_cfo_hz;
            burst_pending_ = burst_latched_ = r.aux != 0;                      // negated first LTS = burst marker (:366-375)
        }
        return result.detected;
    }

    void setAbsoluteTrainingPosition(size_t pos) override { abs_pos_ = pos; has_abs_pos_ = true; }

    bool process(ultra::SampleSpan samples) override {
        if (static_cast<int>(samples.size()) < getSamplesPerSymbol()) return false;
        const size_t ref = has_abs_pos_ ? abs_pos_ : training_start_;
        float ph = static_cast<float>(-2.0f * 3.14159265358979323846 * cfo_hz_ * ref / cfg_.sample_rate);   // (:404)
        while (ph > 3.14159265358979323846) ph -= 2.0f * 3.14159265358979323846;
        while (ph < -3.14159265358979323846) ph += 2.0f * 3.14159265358979323846;
        const float* in = samples.data();
        std::vector<float> undone;
        if (burst_pending_) {                                                   // one-shot: restore the marked LTS (:423-436)
            burst_pending_ = false;
            undone.assign(samples.begin(), samples.end());
            const size_t L = static_cast<size_t>(getSamplesPerSymbol());
            for (size_t i = 0; i < L && i < undone.size(); ++i) undone[i] = -undone[i];
            in = undone.data();
        }
        const int n_sym = static_cast<int>(samples.size()) / getSamplesPerSymbol();
        const int bits = ria_ofdm_data_carriers(&cfg_) * bitsPerCarrier(cfg_.modulation);
        const int stride = std::max(4, ((n_sym > 2 ? n_sym - 2 : 0) * bits + 3) & ~3);
        std::vector<float> llr(static_cast<size_t>(stride));
        int32_t n_llr = 0;
        float snr = 0, cfo = 0, fad = 0;
        Context& c = Context::instance();
        c.check(ria_ofdm_presynced_batch_host(c.get(), &cfg_, in, static_cast<int64_t>(samples.size()),
                                              static_cast<int32_t>(samples.size()), &cfo_hz_, &ph, 1, llr.data(),
                                              static_cast<int32_t>(llr.size()), &n_llr, &snr, &cfo, &fad));
        fading_ = fad;
        const bool ready = n_llr >= RIA_LDPC_N;
        if (ready) {
            llr.resize(static_cast<size_t>(n_llr));
            soft_bits_ = std::move(llr);
            last_snr_ = snr;
            cfo_hz_ = last_cfo_ = cfo;                                          // CFO feedback (:447-455)
        }
        return ready;
    }
    std::vector<float> getSoftBits() override { return std::move(soft_bits_); }
    void reset() override { soft_bits_.clear(); synced_ = false; has_abs_pos_ = false; abs_pos_ = 0; }   // CFO kept (:474-485)
    bool isSynced() const override { return synced_; }
    bool hasData() const override { return !soft_bits_.empty(); }
    float estimatedSNR() const override { return last_snr_; }
    float estimatedCFO() const override { return std::fabs(last_cfo_) > 0.1f ? last_cfo_ : cfo_hz_; }
    float getFadingIndex() const override { return fading_; }
    bool wasBurstInterleaved() const override { return burst_latched_; }
    std::vector<std::complex<float>> getConstellationSymbols() const override { return {}; }

    static int bitsPerCarrier(uint32_t modulation) {
        switch (modulation) {
            case RIA_DQPSK: case RIA_QPSK: return 2;
            case RIA_D8PSK: case RIA_QAM8: return 3;
            case RIA_QAM16: return 4;
            case RIA_QAM32: return 5;
            case RIA_QAM64: return 6;
            case RIA_QAM256: return 8;
            default: return 1;
        }
    }

private:
    ria_modem_config cfg_{};
    ria_chirp_config chirp_{};
    float cfo_hz_ = 0.0f, last_cfo_ = 0.0f, last_snr_ = 0.0f, fading_ = 0.0f;
    size_t training_start_ = 0, abs_pos_ = 0;
    bool has_abs_pos_ = false, synced_ = false, burst_pending_ = false, burst_latched_ = false;
    std::vector<float> soft_bits_;
};

class MCDPSKWaveform : public ultra::MCDPSKWaveform {
public:
    MCDPSKWaveform() : ultra::MCDPSKWaveform() { sync_configs(); }
    explicit MCDPSKWaveform(int num_carriers) : ultra::MCDPSKWaveform(num_carriers) { sync_configs(); }
    explicit MCDPSKWaveform(const ultra::MultiCarrierDPSKConfig& config) : ultra::MCDPSKWaveform(config) { sync_configs(); }

    void configure(ultra::Modulation mod, ultra::CodeRate rate) override {
        ultra::MCDPSKWaveform::configure(mod, rate);
        sync_configs();
    }
    // setCarrierCount / setSpreadingMode are not virtual in the reference: call these instead of the base's
    void setCarrierCount(int carriers) { ultra::MCDPSKWaveform::setCarrierCount(carriers); sync_configs(); }
    void setSpreadingMode(ultra::SpreadingMode mode) { ultra::MCDPSKWaveform::setSpreadingMode(mode); sync_configs(); }
    void setFrequencyOffset(float cfo_hz) override { cfo_hz_ = cfo_hz; }
    float getFrequencyOffset() const override { return cfo_hz_; }

    bool detectSync(ultra::SampleSpan samples, ultra::SyncResult& result, float threshold = 0.15f) override {
        ria_sync_result r{};
        Context& c = Context::instance();
        c.check(ria_chirp_detect_dual_batch_host(c.get(), &chirp_, samples.data(), static_cast<int64_t>(samples.size()),
                                                 static_cast<int32_t>(samples.size()), threshold, 1, &r));
        result.detected = r.detected != 0;
        result.start_sample = r.start_sample;                                   // up_chirp_start (:187)
        result.correlation = std::max(r.correlation, r.snr_estimate);
        result.cfo_hz = r.cfo_hz;
        result.has_training = true;
        if (result.detected) {
            synced_ = true;
            last_cfo_ = r.cfo_hz;
            const int chirp_samples = static_cast<int>(static_cast<size_t>(chirp_.sample_rate * chirp_.duration_ms / 1000.0f));
            const int gap_samples = static_cast<int>(static_cast<size_t>(cfg_.sample_rate * chirp_.gap_ms / 1000.0f));
            result.start_sample = r.aux + chirp_samples + gap_samples;         // training after the down chirp (:209-212)
        }
        return result.detected;
    }

    bool detectDataSync(ultra::SampleSpan samples, ultra::SyncResult& result, float known_cfo_hz = 0.0f,
                        float threshold = 0.2f) override {
        if (std::fabs(known_cfo_hz) > 0.1f) setFrequencyOffset(known_cfo_hz);   // (:247-249)
        ria_sync_result r{};
        Context& c = Context::instance();
        const uint32_t roots = 4u | 8u;                                         // DATA | CONTROL (:253)
        c.check(ria_zc_detect_batch_host(c.get(), &zc_, samples.data(), static_cast<int64_t>(samples.size()),
                                         static_cast<int32_t>(samples.size()), &known_cfo_hz, threshold, roots, 1, &r));
        result.detected = r.detected != 0;
        result.correlation = r.correlation;
        result.cfo_hz = r.cfo_hz;
        result.has_training = true;
        if (result.detected) {
            synced_ = true;
            result.start_sample = r.start_sample;
            last_cfo_ = std::fabs(known_cfo_hz) > 0.1f ? known_cfo_hz + r.cfo_hz : r.cfo_hz;      // (:275-281)
            cfo_hz_ = last_cfo_;
        }
        return result.detected;
    }

    bool process(ultra::SampleSpan samples) override {
        const int n_soft = ria_mcdpsk_soft_bits_per_frame(&cfg_, static_cast<int32_t>(samples.size()));
        if (n_soft <= 0) return false;
        std::vector<float> llr(static_cast<size_t>((n_soft + 3) & ~3));
        int32_t n_llr = 0;
        float fad = 0, cfo = 0;
        Context& c = Context::instance();
        c.check(ria_mcdpsk_process_batch_host(c.get(), &cfg_, samples.data(), static_cast<int64_t>(samples.size()),
                                              static_cast<int32_t>(samples.size()), &cfo_hz_, nullptr, 1, llr.data(),
                                              static_cast<int32_t>(llr.size()), &n_llr, &fad, &cfo));
        fading_ = fad;
        const bool ready = n_llr > 0;
        if (ready) {
            llr.resize(static_cast<size_t>(n_llr));
            soft_bits_ = std::move(llr);
            synced_ = true;
            last_cfo_ = cfo;
        }
        return ready;
    }
    std::vector<float> getSoftBits() override { return std::move(soft_bits_); }
    void reset() override { soft_bits_.clear(); synced_ = false; }
    bool isSynced() const override { return synced_; }
    bool hasData() const override { return !soft_bits_.empty(); }
    float estimatedSNR() const override { return last_snr_; }
    float estimatedCFO() const override { return last_cfo_; }
    float getFadingIndex() const override { return fading_; }
    bool isFading() const override { return fading_ > 0.65f; }
    std::vector<std::complex<float>> getConstellationSymbols() const override { return {}; }

private:
    void sync_configs() {
        const ultra::MultiCarrierDPSKConfig& m = getConfig();
        cfg_.sample_rate = m.sample_rate;
        cfg_.num_carriers = static_cast<uint32_t>(m.num_carriers);
        cfg_.freq_low = m.freq_low;
        cfg_.freq_high = m.freq_high;
        cfg_.samples_per_symbol = static_cast<uint32_t>(m.samples_per_symbol);
        cfg_.bits_per_symbol = static_cast<uint32_t>(m.bits_per_symbol);
        cfg_.spreading = static_cast<uint32_t>(m.getSpreadingFactor());
        cfg_.training_symbols = static_cast<uint32_t>(m.training_symbols);
        ria_chirp_config_default(&chirp_);
        const auto cc = m.getChirpConfig();
        chirp_.sample_rate = cc.sample_rate; chirp_.f_start = cc.f_start; chirp_.f_end = cc.f_end;
        chirp_.duration_ms = cc.duration_ms; chirp_.gap_ms = cc.gap_ms;
        ria_zc_config_default(&zc_);
    }
    ria_mcdpsk_config cfg_{};
    ria_chirp_config chirp_{};
    ria_zc_config zc_{};
    float cfo_hz_ = 0.0f, last_cfo_ = 0.0f, last_snr_ = 0.0f, fading_ = 0.0f;
    bool synced_ = false;
    std::vector<float> soft_bits_;
};

// OFDM_COX (ultra::OFDMNvisWaveform, src/waveform/ofdm_cox_waveform.cpp): Schmidl-Cox acquisition instead of the chirp,
// then the same presynced demodulator.
//   detectSync -> ria_ofdm_cox_search_sync_batch_host   (OFDMDemodulator::searchForSync, :121-153)
//   process    -> ria_ofdm_presynced_batch_host         (:160-218)
// The library covers the modem's 1024-point configuration (ModemConfig defaults / createNvisMode()); the
// default-constructed 512-point variant is reported as RIA_E_UNSUPPORTED by the calls below.
class OFDMNvisWaveform : public ultra::OFDMNvisWaveform {
public:
    OFDMNvisWaveform() : ultra::OFDMNvisWaveform() { cfg_ = toRiaConfig(getConfig()); }
    explicit OFDMNvisWaveform(const ultra::ModemConfig& config) : ultra::OFDMNvisWaveform(config) { cfg_ = toRiaConfig(getConfig()); }

    // configure / setTxFrequencyOffset / setUsePilots rebuild the reference's demodulator (initComponents), which also
    // forgets its noise-floor tracker
    void configure(ultra::Modulation mod, ultra::CodeRate rate) override {
        ultra::OFDMNvisWaveform::configure(mod, rate);
        cfg_ = toRiaConfig(getConfig());
        noise_floor_ = 0.0f; est_cfo_ = 0.0f; last_snr_ = 0.0f; demod_synced_ = false;
    }
    void setTxFrequencyOffset(float cfo_hz) override {
        ultra::OFDMNvisWaveform::setTxFrequencyOffset(cfo_hz);
        noise_floor_ = 0.0f; est_cfo_ = 0.0f; last_snr_ = 0.0f; demod_synced_ = false;
    }
    void setUsePilots(bool use_pilots) {                                        // not virtual in the reference
        ultra::OFDMNvisWaveform::setUsePilots(use_pilots);
        cfg_ = toRiaConfig(getConfig());
        noise_floor_ = 0.0f; est_cfo_ = 0.0f; last_snr_ = 0.0f; demod_synced_ = false;
    }
    void setFrequencyOffset(float cfo_hz) override { cfo_hz_ = est_cfo_ = cfo_hz; }
    float getFrequencyOffset() const override { return cfo_hz_; }

    bool detectSync(ultra::SampleSpan samples, ultra::SyncResult& result, float threshold = 0.8f) override {
        ria_sync_result r{};
        Context& c = Context::instance();
        c.check(ria_ofdm_cox_search_sync_batch_host(c.get(), &cfg_, samples.data(), static_cast<int64_t>(samples.size()),
                                                    static_cast<int32_t>(samples.size()), threshold, &noise_floor_, 1, &r));
        if (!r.detected) return false;                                          // result untouched (:151-152)
        result.detected = true;
        result.start_sample = r.start_sample;                                   // first LTS sample
        result.cfo_hz = r.cfo_hz;
        result.snr_estimate = 0.0f;
        result.has_training = true;
        result.correlation = 0.9f;
        cfo_hz_ = r.cfo_hz;
        synced_ = true;
        training_start_ = static_cast<size_t>(r.start_sample);
        return true;
    }

    void setAbsoluteTrainingPosition(size_t pos) override { abs_pos_ = pos; has_abs_pos_ = true; }

    bool process(ultra::SampleSpan samples) override {
        if (!soft_bits_.empty()) return true;                                   // (:165-167)
        const size_t ref = has_abs_pos_ ? abs_pos_ : training_start_;
        float ph = static_cast<float>(-2.0f * 3.14159265358979323846 * cfo_hz_ * ref / cfg_.sample_rate);   // (:176)
        while (ph > 3.14159265358979323846) ph -= 2.0f * 3.14159265358979323846;
        while (ph < -3.14159265358979323846) ph += 2.0f * 3.14159265358979323846;
        est_cfo_ = cfo_hz_;                                                     // setFrequencyOffsetWithPhase
        if (static_cast<int>(samples.size()) < getSamplesPerSymbol()) return false;
        demod_synced_ = true;
        const int n_sym = static_cast<int>(samples.size()) / getSamplesPerSymbol();
        const int bits = ria_ofdm_data_carriers(&cfg_) * OFDMChirpWaveform::bitsPerCarrier(cfg_.modulation);
        const int stride = std::max(4, ((n_sym > 2 ? n_sym - 2 : 0) * bits + 3) & ~3);
        std::vector<float> llr(static_cast<size_t>(stride));
        int32_t n_llr = 0;
        float snr = 0, cfo = 0, fad = 0;
        Context& c = Context::instance();
        c.check(ria_ofdm_presynced_batch_host(c.get(), &cfg_, samples.data(), static_cast<int64_t>(samples.size()),
                                              static_cast<int32_t>(samples.size()), &cfo_hz_, &ph, 1, llr.data(),
                                              static_cast<int32_t>(llr.size()), &n_llr, &snr, &cfo, &fad));
        est_cfo_ = cfo;
        last_snr_ = snr;
        const bool ready = n_llr >= RIA_LDPC_N;
        if (ready) {
            llr.resize(static_cast<size_t>(n_llr));
            soft_bits_ = std::move(llr);
            cfo_hz_ = cfo;                                                      // CFO feedback (:205-210)
        }
        return ready;
    }
    std::vector<float> getSoftBits() override { return std::move(soft_bits_); }
    void reset() override {                                                     // CFO kept (:224-235); the demodulator's is not
        soft_bits_.clear();
        synced_ = demod_synced_ = false;
        training_start_ = 0; has_abs_pos_ = false; abs_pos_ = 0;
        est_cfo_ = 0.0f;
    }
    bool isSynced() const override { return synced_ || demod_synced_; }
    bool hasData() const override { return !soft_bits_.empty(); }
    float estimatedSNR() const override { return last_snr_; }
    float estimatedCFO() const override { return est_cfo_; }
    std::vector<std::complex<float>> getConstellationSymbols() const override { return {}; }

private:
    ria_modem_config cfg_{};
    float cfo_hz_ = 0.0f, est_cfo_ = 0.0f, last_snr_ = 0.0f, noise_floor_ = 0.0f;
    size_t training_start_ = 0, abs_pos_ = 0;
    bool has_abs_pos_ = false, synced_ = false, demod_synced_ = false;
    std::vector<float> soft_bits_;
};

// What WaveformFactory::create(mode) / create(mode, config) (src/waveform/waveform_factory.cpp:13-61) would
// return for the three waveforms of the hot path; nullptr for every other mode, like the factory's default branch.
inline ultra::WaveformPtr createWaveform(ultra::protocol::WaveformMode mode) {
    switch (mode) {
        case ultra::protocol::WaveformMode::OFDM_CHIRP: return std::make_unique<OFDMChirpWaveform>();
        case ultra::protocol::WaveformMode::MC_DPSK: return std::make_unique<MCDPSKWaveform>();
        case ultra::protocol::WaveformMode::OFDM_COX: return std::make_unique<OFDMNvisWaveform>();
        default: return nullptr;
    }
}
inline ultra::WaveformPtr createWaveform(ultra::protocol::WaveformMode mode, const ultra::ModemConfig& config) {
    if (mode == ultra::protocol::WaveformMode::OFDM_CHIRP) return std::make_unique<OFDMChirpWaveform>(config);
    if (mode == ultra::protocol::WaveformMode::OFDM_COX) return std::make_unique<OFDMNvisWaveform>(config);
    return createWaveform(mode);
}
#endif  // RIA_WITH_ULTRA

}  // namespace ria
