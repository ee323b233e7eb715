// TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" shim over the UNMODIFIED reference classes, compiled together with the
// reference sources where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/libria_ref.so.  It exists so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can call the reference's own implementation of the
// hot path through ctypes.  Nothing in ria_b200/ may link or load it.
//
// Every function here only marshals plain pointers into the reference's public API; no
// algorithm is restated in this file.

#include "ultra/types.hpp"
#include "ultra/fec.hpp"
#include "ultra/ofdm.hpp"
#include "ultra/dsp.hpp"
#include "ultra/logging.hpp"
#include "ofdm/demodulator_impl.hpp"      // private Impl, reached with -fno-access-control (taps)
#include "fec/frame_interleaver.hpp"
#include "fec/ldpc_codec.hpp"
#include "protocol/frame_v2.hpp"
#include "psk/multi_carrier_dpsk.hpp"
#include "fec/chase_cache.hpp"
#include "sync/zc_sync.hpp"
#include "sync/chirp_sync.hpp"
#include "sim/hf_channel.hpp"
#include "waveform/ofdm_chirp_waveform.hpp"
#include "protocol/waveform_selection.hpp"
#include "gui/modem/streaming_decoder.hpp"
#include "gui/modem/streaming_encoder.hpp"   // test transmissions exactly as a station sends them   // frame-level decode semantics (private members reached with -fno-access-control)

#include "ria_b200.h"                     // POD config / status structs shared with the product ABI

#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include "fec/burst_interleaver.hpp"
#include <vector>

using namespace ultra;

extern "C" {

// Silence the reference's INFO logging (hot functions log under a global
// mutex; any CPU timing must run at ERROR level).
void ref_quiet(void) { setLogLevel(LogLevel::ERROR); }

// ---------------------------------------------------------------------------------------------
// LDPC  (include/ultra/fec.hpp:21-81)
// ---------------------------------------------------------------------------------------------

// LDPCEncoder::encode (src/fec/ldpc_encoder.cpp:193-257). Returns number of coded bytes.
int ref_ldpc_encode(int rate, const uint8_t* data, int len, uint8_t* out, int out_cap) {
    LDPCEncoder enc(static_cast<CodeRate>(rate));
    Bytes coded = enc.encode(ByteSpan(data, static_cast<size_t>(len)));
    int n = static_cast<int>(coded.size());
    if (n > out_cap) return -n;
    std::memcpy(out, coded.data(), coded.size());
    return n;
}

struct RefLdpcDecoder {
    LDPCDecoder dec;
    explicit RefLdpcDecoder(CodeRate r) : dec(r) {}
};

void* ref_ldpc_decoder_new(int rate, int max_iter, float factor) {
    auto* d = new RefLdpcDecoder(static_cast<CodeRate>(rate));
    d->dec.setMaxIterations(max_iter);
    d->dec.setMinSumFactor(factor);
    return d;
}
void ref_ldpc_decoder_free(void* h) { delete static_cast<RefLdpcDecoder*>(h); }

// LDPCDecoder::decodeSoft (src/fec/ldpc_decoder.cpp:284-429) on one span of n_llr floats.
// Returns number of output bytes; *ok = lastDecodeSuccess(), *iters = lastIterations().
int ref_ldpc_decode_soft(void* h, const float* llr, int n_llr, uint8_t* out, int out_cap,
                         int* ok, int* iters) {
    auto* d = static_cast<RefLdpcDecoder*>(h);
    Bytes b = d->dec.decodeSoft(std::span<const float>(llr, static_cast<size_t>(n_llr)));
    *ok = d->dec.lastDecodeSuccess() ? 1 : 0;
    *iters = d->dec.lastIterations();
    int n = static_cast<int>(b.size());
    if (n > out_cap) return -n;
    if (n) std::memcpy(out, b.data(), b.size());
    return n;
}

// Batch convenience for the CPU baseline: n_cw independent codewords of 648 LLRs each,
// decoded one by one with the same decoder object (exactly what decodeFixedFrame does,
// src/protocol/frame_v2.cpp:1359-1385).  out is [n_cw][out_stride].
void ref_ldpc_decode_batch(void* h, const float* llr, int n_cw, uint8_t* out, int out_stride,
                           uint8_t* ok, int32_t* iters) {
    auto* d = static_cast<RefLdpcDecoder*>(h);
    for (int c = 0; c < n_cw; ++c) {
        Bytes b = d->dec.decodeSoft(std::span<const float>(llr + static_cast<size_t>(c) * 648, 648));
        ok[c] = d->dec.lastDecodeSuccess() ? 1 : 0;
        iters[c] = d->dec.lastIterations();
        size_t n = std::min(b.size(), static_cast<size_t>(out_stride));
        std::memset(out + static_cast<size_t>(c) * out_stride, 0, out_stride);
        std::memcpy(out + static_cast<size_t>(c) * out_stride, b.data(), n);
    }
}


// ---------------------------------------------------------------------------------------------
// OFDM  (include/ultra/ofdm.hpp)
// ---------------------------------------------------------------------------------------------
static ModemConfig to_cfg(const ria_modem_config* c) {
    ModemConfig m;
    m.sample_rate = c->sample_rate;
    m.center_freq = c->center_freq;
    m.fft_size = c->fft_size;
    m.num_carriers = c->num_carriers;
    m.cp_mode = static_cast<CyclicPrefixMode>(c->cp_mode);
    m.symbol_guard = c->symbol_guard;
    m.use_pilots = c->use_pilots != 0;
    m.pilot_spacing = c->pilot_spacing;
    m.modulation = static_cast<Modulation>(c->modulation);
    return m;
}

// TX of one frame exactly as OFDMChirpWaveform does after the chirp:
// generateTrainingSymbols(2) (modulator.cpp:528-582) + modulate(data, mod) (modulator.cpp:348-477).
int ref_ofdm_tx_frame(const ria_modem_config* c, const uint8_t* data, int len, float* out, int cap) {
    ModemConfig m = to_cfg(c);
    OFDMModulator mod(m);
    Samples tr = mod.generateTrainingSymbols(static_cast<int>(c->training_symbols));
    Samples d = mod.modulate(ByteSpan(data, static_cast<size_t>(len)), m.modulation);
    int n = static_cast<int>(tr.size() + d.size());
    if (n > cap) return -n;
    std::memcpy(out, tr.data(), tr.size() * sizeof(float));
    std::memcpy(out + tr.size(), d.data(), d.size() * sizeof(float));
    return n;
}

struct RefOfdmDemod {
    ModemConfig cfg;
    OFDMDemodulator dem;
    explicit RefOfdmDemod(const ModemConfig& m) : cfg(m), dem(m) {}
};

void* ref_ofdm_demod_new(const ria_modem_config* c) { return new RefOfdmDemod(to_cfg(c)); }
void ref_ofdm_demod_free(void* h) { delete static_cast<RefOfdmDemod*>(h); }

// What OFDMChirpWaveform::process does with one frame (ofdm_chirp_waveform.cpp:391-468):
// setFrequencyOffsetWithPhase(cfo, phase); processPresynced(samples, training); drain soft bits.
// Returns processPresynced's bool.  h_lts/bins are optional taps read from the private Impl.
int ref_ofdm_process_presynced(void* h, const float* samples, int n, float cfo_hz, float phase,
                               float* soft, int soft_cap, int* n_soft, float* snr_db,
                               float* cfo_out, float* fading, float* h_final /*[num_carriers][2]*/) {
    auto* d = static_cast<RefOfdmDemod*>(h);
    d->dem.reset();
    d->dem.setFrequencyOffsetWithPhase(cfo_hz, phase);
    bool ready = d->dem.processPresynced(SampleSpan(samples, static_cast<size_t>(n)), 2);
    if (snr_db) *snr_db = d->dem.getEstimatedSNR();
    if (cfo_out) *cfo_out = d->dem.getFrequencyOffset();
    if (fading) *fading = d->dem.getFadingIndex();
    if (h_final) {
        auto& im = *d->dem.impl_;
        for (size_t i = 0; i < im.all_carrier_fft_indices.size(); ++i) {
            Complex v = im.channel_estimate[im.all_carrier_fft_indices[i]];
            h_final[2 * i] = v.real();
            h_final[2 * i + 1] = v.imag();
        }
    }
    int total = 0;
    for (;;) {
        std::vector<float> chunk = d->dem.getSoftBits();
        if (chunk.empty()) break;
        if (total + static_cast<int>(chunk.size()) <= soft_cap)
            std::memcpy(soft + total, chunk.data(), chunk.size() * sizeof(float));
        total += static_cast<int>(chunk.size());
    }
    *n_soft = total;
    return ready ? 1 : 0;
}

// ultra::FFT::forward (src/dsp/fft.cpp:130-146), interleaved re/im
void ref_fft_forward(int size, const float* in, float* out) {
    FFT fft(static_cast<size_t>(size));
    fft.forward(reinterpret_cast<const Complex*>(in), reinterpret_cast<Complex*>(out));
}

// Per-symbol taps: toBaseband + extractSymbol (channel_equalizer.cpp:99-187) of the first
// n_sym symbols of a frame with the mixer running from 0, CFO correction as given.
// bins: [n_sym][num_carriers][2] in logical carrier order.
void ref_ofdm_symbol_bins(void* h, const float* samples, int n_sym, float cfo_hz, float phase, float* bins) {
    auto* d = static_cast<RefOfdmDemod*>(h);
    d->dem.reset();
    d->dem.setFrequencyOffsetWithPhase(cfo_hz, phase);
    auto& im = *d->dem.impl_;
    im.mixer.reset();
    const size_t L = im.symbol_samples;
    const size_t nc = im.all_carrier_fft_indices.size();
    for (int s = 0; s < n_sym; ++s) {
        auto bb = im.toBaseband(SampleSpan(samples + s * L, L));
        auto fd = im.extractSymbol(bb, 0);
        for (size_t i = 0; i < nc; ++i) {
            Complex v = fd[im.all_carrier_fft_indices[i]];
            bins[(s * nc + i) * 2] = v.real();
            bins[(s * nc + i) * 2 + 1] = v.imag();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fixed 4-codeword frame  (src/protocol/frame_v2.cpp:1285-1385)
// ---------------------------------------------------------------------------------------------
int ref_encode_fixed_frame(const uint8_t* data, int len, int rate, int use_ci, int bps, uint8_t* out, int cap) {
    Bytes in(data, data + len);
    Bytes coded = protocol::v2::encodeFixedFrame(in, static_cast<CodeRate>(rate), use_ci != 0, static_cast<size_t>(bps));
    int n = static_cast<int>(coded.size());
    if (n > cap) return -n;
    std::memcpy(out, coded.data(), coded.size());
    return n;
}

// First pass of decodeFixedFrame composed from the reference's own pieces, no retry ladder:
// FrameInterleaver::deinterleave -> ChannelInterleaver::deinterleave -> decodeSoft x4.
void ref_frame_decode_first_pass(const float* soft, int rate, int use_ci, int bps,
                                 uint8_t* data /*[4*bytes_per_cw]*/, uint8_t* ok, int32_t* iters) {
    using namespace fec;
    std::vector<float> v(soft, soft + FrameInterleaver::TOTAL_FRAME_BITS);
    auto cws = FrameInterleaver::deinterleave(v);
    CodeRate r = static_cast<CodeRate>(rate);
    LDPCDecoder dec(r);
    dec.setMaxIterations(LDPCCodec::getRecommendedIterations(r));
    dec.setMinSumFactor(0.9375f);
    size_t bpc = protocol::v2::getBytesPerCodeword(r);
    std::unique_ptr<ChannelInterleaver> il;
    if (use_ci) il = std::make_unique<ChannelInterleaver>(static_cast<size_t>(bps), 648);
    for (int c = 0; c < 4; ++c) {
        std::vector<float> bits = cws[c];
        if (il) bits = il->deinterleave(bits);
        Bytes b = dec.decodeSoft(bits);
        ok[c] = dec.lastDecodeSuccess() ? 1 : 0;
        iters[c] = dec.lastIterations();
        std::memset(data + c * bpc, 0, bpc);
        if (ok[c] && b.size() >= bpc) std::memcpy(data + c * bpc, b.data(), bpc);
    }
}

// The reference's complete decodeFixedFrame (with its retry ladder and repair passes).
void ref_decode_fixed_frame_full(const float* soft, int n_soft, int rate, int use_ci, int bps,
                                 uint8_t* data, uint8_t* ok) {
    std::vector<float> v(soft, soft + n_soft);
    CodeRate r = static_cast<CodeRate>(rate);
    auto st = protocol::v2::decodeFixedFrame(v, r, use_ci != 0, static_cast<size_t>(bps));
    size_t bpc = protocol::v2::getBytesPerCodeword(r);
    for (int c = 0; c < 4; ++c) {
        ok[c] = st.decoded[c] ? 1 : 0;
        std::memset(data + c * bpc, 0, bpc);
        if (st.decoded[c] && st.data[c].size() >= bpc) std::memcpy(data + c * bpc, st.data[c].data(), bpc);
    }
}

// The soft-bit perturbation of one retry-ladder attempt, stated with the SAME library objects the
// reference uses (std::mt19937 + std::normal_distribution<float> of this libstdc++; the ladder itself
// is inline in decodeFixedFrame, frame_v2.cpp:1389-1546, and cannot be called separately).  Test
// helper for ria_ldpc_ladder_perturb_dev; the end-to-end check is ref_decode_fixed_frame_full.
//   kind: 1 add, 2 clip 10, 3 scale 0.5, 4 clip 6, 5 hard, 6 scale 0.25
void ref_ladder_perturb(const float* in, int n, unsigned seed, float sigma, int kind, float* out) {
    std::mt19937 rng(seed);
    std::normal_distribution<float> noise(0.0f, sigma);
    for (int i = 0; i < n; ++i) {
        float llr = in[i];
        switch (kind) {
            case 2: llr = std::max(-10.0f, std::min(10.0f, llr)); llr += noise(rng); break;
            case 3: llr = llr * 0.5f + noise(rng); break;
            case 4: llr = std::max(-6.0f, std::min(6.0f, llr)); llr += noise(rng); break;
            case 5: llr = (llr >= 0) ? 1.0f : -1.0f; llr += noise(rng); break;
            case 6: llr = llr * 0.25f + noise(rng); break;
            default: llr += noise(rng); break;
        }
        out[i] = llr;
    }
}

// fec::BurstInterleaver::deinterleave (src/fec/burst_interleaver.cpp:39-78): phys / out are [n][2592]
void ref_burst_deinterleave(const float* phys, int n, float* out) {
    std::vector<std::vector<float>> p(n);
    for (int i = 0; i < n; ++i) p[i].assign(phys + i * 2592, phys + (i + 1) * 2592);
    auto l = fec::BurstInterleaver::deinterleave(p);
    for (int i = 0; i < n; ++i) std::memcpy(out + i * 2592, l[i].data(), 2592 * sizeof(float));
}

// v2::parseHeader (frame_v2.cpp:1195-1253) + DataFrame::deserialize frame CRC (:555-600)
void ref_parse_header(const uint8_t* data, int len, ria_frame_status* st) {
    Bytes b(data, data + len);
    auto hi = protocol::v2::parseHeader(b);
    st->header_valid = hi.valid ? 1 : 0;
    st->type = static_cast<uint8_t>(hi.type);
    st->seq = hi.seq;
    st->src_hash = hi.src_hash;
    st->dst_hash = hi.dst_hash;
    st->total_cw = hi.total_cw;
    st->payload_len = hi.payload_len;
    st->frame_crc_ok = 0;
    if (hi.valid && !hi.is_control) {
        auto f = protocol::v2::DataFrame::deserialize(ByteSpan(data, static_cast<size_t>(len)));
        st->frame_crc_ok = f.has_value() ? 1 : 0;
    }
}

// What decodeFixedFrame's caller sees for a frame whose codeword results are (ok[c], data[c]): the header of
// codeword 0 (parseHeader) and whether CodewordStatus::reassemble() (frame_v2.cpp:1030-1066, incl. the
// DATA_CW_MARKER rule of reassembleCodewords :960-985) yields a frame that DataFrame::deserialize accepts.
void ref_frame_status_reassembled(const uint8_t* data, const uint8_t* ok, int bpc, ria_frame_status* st) {
    protocol::v2::CodewordStatus cs;
    for (int c = 0; c < 4; ++c) {
        cs.decoded.push_back(ok[c] != 0);
        cs.data.emplace_back(data + c * bpc, data + (c + 1) * bpc);
    }
    std::memset(st, 0, sizeof *st);
    if (!ok[0]) return;
    auto hi = protocol::v2::parseHeader(cs.data[0]);
    st->header_valid = hi.valid ? 1 : 0;
    st->type = static_cast<uint8_t>(hi.type);
    st->seq = hi.seq;
    st->src_hash = hi.src_hash;
    st->dst_hash = hi.dst_hash;
    st->total_cw = hi.total_cw;
    st->payload_len = hi.payload_len;
    if (hi.valid && !hi.is_control) {
        Bytes fr = cs.reassemble();
        auto f = protocol::v2::DataFrame::deserialize(ByteSpan(fr.data(), fr.size()));
        st->frame_crc_ok = f.has_value() ? 1 : 0;
    }
}

uint16_t ref_crc16(const uint8_t* data, int len) {
    return protocol::v2::ControlFrame::calculateCRC(data, static_cast<size_t>(len));
}

// Build a serialized v2 data frame (DataFrame::makeData + serialize) for test inputs.
int ref_make_data_frame(const char* src, const char* dst, int seq, const uint8_t* payload, int len,
                        uint8_t* out, int cap) {
    Bytes p(payload, payload + len);
    auto f = protocol::v2::DataFrame::makeData(src, dst, static_cast<uint16_t>(seq), p);
    Bytes s = f.serialize();
    int n = static_cast<int>(s.size());
    if (n > cap) return -n;
    std::memcpy(out, s.data(), s.size());
    return n;
}


// ---------------------------------------------------------------------------------------------
// StreamingDecoder frame-level decode (src/gui/modem/streaming_decoder.cpp)
// ---------------------------------------------------------------------------------------------
// One StreamingDecoder object = one receiver: its HARQ chase cache persists across calls, so feeding it the
// receptions of a retransmitted frame one after the other exercises the combining path (:2762-2789).
void* ref_stream_decoder_new(void) { return new gui::StreamingDecoder(); }
void ref_stream_decoder_free(void* h) { delete static_cast<gui::StreamingDecoder*>(h); }

struct ref_decode_result {
    int32_t success, frame_type, codewords_ok, codewords_failed, is_ping, n_bytes;
};

// StreamingDecoder::decodeMCDPSKFrame(soft_bits, rate, bytes_per_cw, snr, cfo) (:2580-2822)
void ref_stream_decode_mcdpsk_frame(void* h, const float* soft, int n_soft, int rate, ref_decode_result* out,
                                    uint8_t* bytes, int cap) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    std::vector<float> v(soft, soft + n_soft);
    CodeRate r = static_cast<CodeRate>(rate);
    auto res = d->decodeMCDPSKFrame(v, r, protocol::v2::getBytesPerCodeword(r), 0.0f, 0.0f);
    out->success = res.success ? 1 : 0;
    out->frame_type = static_cast<int32_t>(res.frame_type);
    out->codewords_ok = res.codewords_ok;
    out->codewords_failed = res.codewords_failed;
    out->is_ping = res.is_ping ? 1 : 0;
    out->n_bytes = static_cast<int32_t>(res.frame_data.size());
    std::memcpy(bytes, res.frame_data.data(), std::min<size_t>(res.frame_data.size(), static_cast<size_t>(cap)));
}

// StreamingDecoder::decodeFrame(soft_bits, snr, cfo) (:2820-3056) for an OFDM receiver in the given state.  The state
// a StreamingDecoder reaches through setMode / setConnectedOFDMMode / setDataMode is written directly (private
// members, -fno-access-control) so that no audio has to be streamed through it.
void ref_stream_decode_ofdm_frame(void* h, const float* soft, int n_soft, int connected, int modulation, int rate,
                                  int data_carriers, int use_channel_interleave, ref_decode_result* out,
                                  uint8_t* bytes, int cap) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    d->mode_ = protocol::WaveformMode::OFDM_CHIRP;
    d->connected_ = connected != 0;
    d->current_modulation_ = static_cast<Modulation>(modulation);
    d->code_rate_ = static_cast<CodeRate>(rate);
    d->ofdm_data_carriers_ = data_carriers;
    d->use_channel_interleave_ = use_channel_interleave != 0;
    d->interleaver_ = std::make_unique<ChannelInterleaver>(
        static_cast<size_t>(data_carriers) * getBitsPerSymbol(static_cast<Modulation>(modulation)), protocol::v2::LDPC_CODEWORD_BITS);
    std::vector<float> v(soft, soft + n_soft);
    auto res = d->decodeFrame(v, 0.0f, 0.0f);
    out->success = res.success ? 1 : 0;
    out->frame_type = static_cast<int32_t>(res.frame_type);
    out->codewords_ok = res.codewords_ok;
    out->codewords_failed = res.codewords_failed;
    out->is_ping = res.is_ping ? 1 : 0;
    out->n_bytes = static_cast<int32_t>(res.frame_data.size());
    std::memcpy(bytes, res.frame_data.data(), std::min<size_t>(res.frame_data.size(), static_cast<size_t>(cap)));
}

// One step of the receive state machine: StreamingDecoder::decodeCurrentFrame (:1060-1760) on `samples` placed at the
// start of the ring buffer as if sync had been found at their first sample (detectSync / detectDataSync having left
// the waveform with `training_start` / the burst marker is not reproduced: the samples start AT the training
// symbols at samples[sync_pos]).  The decoder object keeps its mode between calls (ref_stream_setup_*).
struct ref_stream_step_out {
    int32_t state;              // DecoderState after the step (0 SEARCHING, 1 SYNC_FOUND, 2 DECODING, 3 BURST_ACCUMULATING)
    int32_t pending_total_cw;   // escalation request (0 = none)
    int32_t has_frame;
    ref_decode_result frame;
    float last_cfo;
    int32_t sync_pos;           // sync_position_ after the step (moved by the multi-candidate recovery)
};

void ref_stream_setup_ofdm(void* h, int connected, int modulation, int rate) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    ModemConfig cfg;                                            // the modem's configuration: 1024-point FFT, 59 carriers
    cfg.use_pilots = true;
    cfg.pilot_spacing = 10;
    if (connected) d->setConnectedOFDMMode(protocol::WaveformMode::OFDM_CHIRP, cfg, static_cast<Modulation>(modulation),
                                           static_cast<CodeRate>(rate));
    else { d->setMode(protocol::WaveformMode::OFDM_CHIRP, false); d->setOFDMConfig(cfg);
           d->setDataMode(static_cast<Modulation>(modulation), static_cast<CodeRate>(rate)); }
}

void ref_stream_setup_mcdpsk(void* h, int connected, int carriers, int modulation, int rate, int spreading) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    d->setMCDPSKCarriers(carriers);
    d->setSpreadingMode(static_cast<SpreadingMode>(spreading));
    d->mode_ = protocol::WaveformMode::OFDM_CHIRP;              // force setMode to rebuild the waveform
    d->setMode(protocol::WaveformMode::MC_DPSK, connected != 0);
    if (connected) d->setDataMode(static_cast<Modulation>(modulation), static_cast<CodeRate>(rate));
}

int ref_stream_min_control_samples(void* h) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    return d->waveform_ ? d->waveform_->getMinSamplesForControlFrame() : 0;
}

void ref_stream_step(void* h, const float* samples, int n, int sync_pos, float sync_cfo, float sync_snr, int pending_total_cw,
                     float last_cfo, ref_stream_step_out* out, uint8_t* bytes, int cap) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    {
        std::lock_guard<std::mutex> lock(d->buffer_mutex_);
        if (d->buffer_.size() < gui::StreamingDecoder::MAX_BUFFER_SAMPLES) d->buffer_.assign(gui::StreamingDecoder::MAX_BUFFER_SAMPLES, 0.0f);
        std::memcpy(d->buffer_.data(), samples, static_cast<size_t>(n) * sizeof(float));
        d->write_pos_ = static_cast<size_t>(n);
        d->total_fed_ = static_cast<size_t>(n);
        d->sync_position_ = static_cast<size_t>(sync_pos);
        d->correlation_pos_ = static_cast<size_t>(sync_pos);
    }
    while (!d->frame_queue_.empty()) d->frame_queue_.pop();
    d->sync_cfo_ = sync_cfo;
    d->sync_snr_ = sync_snr;
    d->last_cfo_.store(last_cfo);
    d->pending_total_cw_ = pending_total_cw;
    d->state_ = gui::DecoderState::DECODING;
    if (d->waveform_) {                                         // what searchForSync leaves behind (:880-897)
        d->waveform_->reset();
        d->waveform_->setAbsoluteTrainingPosition(static_cast<size_t>(sync_pos));
    }
    d->decodeCurrentFrame();
    out->sync_pos = static_cast<int32_t>(d->sync_position_);
    out->state = static_cast<int32_t>(d->state_);
    out->pending_total_cw = d->pending_total_cw_;
    out->last_cfo = d->last_cfo_.load();
    out->has_frame = d->frame_queue_.empty() ? 0 : 1;
    std::memset(&out->frame, 0, sizeof out->frame);
    if (out->has_frame) {
        auto res = d->frame_queue_.front();
        out->frame.success = res.success ? 1 : 0;
        out->frame.frame_type = static_cast<int32_t>(res.frame_type);
        out->frame.codewords_ok = res.codewords_ok;
        out->frame.codewords_failed = res.codewords_failed;
        out->frame.is_ping = res.is_ping ? 1 : 0;
        out->frame.n_bytes = static_cast<int32_t>(res.frame_data.size());
        std::memcpy(bytes, res.frame_data.data(), std::min<size_t>(res.frame_data.size(), static_cast<size_t>(cap)));
    }
}

// StreamingEncoder::encodeBurstLight (streaming_encoder.cpp:302-390) of `n_frames` equally long frames with the burst
// interleaver on: every group of `group` frames is interleaved and its first LTS symbol negated as the marker.
int ref_stream_encode_burst(int modulation, int rate, int group, const uint8_t* frames, int frame_len, int n_frames,
                            float* out, int cap) {
    gui::StreamingEncoder enc;
    enc.setMode(protocol::WaveformMode::OFDM_CHIRP);
    ModemConfig cfg;
    cfg.use_pilots = true;
    cfg.pilot_spacing = 10;
    enc.setOFDMConfig(cfg);
    enc.setDataMode(Modulation::QPSK, CodeRate::R1_2);
    enc.setDataMode(static_cast<Modulation>(modulation), static_cast<CodeRate>(rate));
    enc.setBurstInterleave(true);
    enc.setBurstInterleaveGroupSize(group);
    std::vector<Bytes> list;
    for (int i = 0; i < n_frames; ++i) list.emplace_back(frames + i * frame_len, frames + (i + 1) * frame_len);
    std::vector<float> s = enc.encodeBurstLight(list);
    const int n = static_cast<int>(s.size());
    if (n > cap) return -n;
    std::memcpy(out, s.data(), s.size() * sizeof(float));
    return n;
}

// A burst group through the reference state machine: decodeCurrentFrame with the burst marker latched (what
// detectDataSync leaves behind when it sees the negated LTS, ofdm_chirp_waveform.cpp:366-375), then
// accumulateBurstFrames until the group is finalised or aborted (:3065-3240).  Results: the queued DecodeResults in
// order (frame bytes concatenated with stride `stride`); returns their count, or -1 - state when the decoder did not
// return to SEARCHING.
int ref_stream_burst_group(void* h, const float* samples, int n, int sync_pos, float sync_cfo, float last_cfo, int group,
                           ref_decode_result* results, uint8_t* bytes, int stride, int cap, float* last_cfo_out) {
    auto* d = static_cast<gui::StreamingDecoder*>(h);
    {
        std::lock_guard<std::mutex> lock(d->buffer_mutex_);
        if (d->buffer_.size() < gui::StreamingDecoder::MAX_BUFFER_SAMPLES) d->buffer_.assign(gui::StreamingDecoder::MAX_BUFFER_SAMPLES, 0.0f);
        std::memcpy(d->buffer_.data(), samples, static_cast<size_t>(n) * sizeof(float));
        d->write_pos_ = static_cast<size_t>(n);
        d->total_fed_ = static_cast<size_t>(n);
        d->sync_position_ = static_cast<size_t>(sync_pos);
        d->correlation_pos_ = static_cast<size_t>(sync_pos);
    }
    while (!d->frame_queue_.empty()) d->frame_queue_.pop();
    d->sync_cfo_ = sync_cfo;
    d->sync_snr_ = 15.0f;
    d->last_cfo_.store(last_cfo);
    d->pending_total_cw_ = 0;
    d->use_burst_interleave_ = true;
    d->burst_group_size_ = group;
    d->state_ = gui::DecoderState::DECODING;
    auto* wf = dynamic_cast<OFDMChirpWaveform*>(d->waveform_.get());
    if (!wf) return -100;
    wf->reset();
    wf->setAbsoluteTrainingPosition(static_cast<size_t>(sync_pos));
    wf->burst_interleaved_detected_ = true;
    wf->burst_interleave_latched_ = true;
    d->decodeCurrentFrame();
    for (int it = 0; it < 4 * group && d->state_ == gui::DecoderState::BURST_ACCUMULATING; ++it) d->accumulateBurstFrames();
    wf->burst_interleave_latched_ = false;
    *last_cfo_out = d->last_cfo_.load();
    if (d->state_ != gui::DecoderState::SEARCHING) return -1 - static_cast<int>(d->state_);
    int count = 0;
    while (!d->frame_queue_.empty() && count < cap) {
        auto res = d->frame_queue_.front();
        d->frame_queue_.pop();
        ref_decode_result& o = results[count];
        o.success = res.success ? 1 : 0;
        o.frame_type = static_cast<int32_t>(res.frame_type);
        o.codewords_ok = res.codewords_ok;
        o.codewords_failed = res.codewords_failed;
        o.is_ping = res.is_ping ? 1 : 0;
        o.n_bytes = static_cast<int32_t>(res.frame_data.size());
        std::memcpy(bytes + static_cast<size_t>(count) * stride, res.frame_data.data(),
                    std::min<size_t>(res.frame_data.size(), static_cast<size_t>(stride)));
        ++count;
    }
    return count;
}

// v2::ControlFrame::makeAck / makeNack(...).serialize(): 20-byte control frames (frame_v2.cpp)
int ref_make_ack_frame(const char* src, const char* dst, int seq, int nack, uint8_t* out, int cap) {
    auto f = nack ? protocol::v2::ControlFrame::makeNack(src, dst, static_cast<uint16_t>(seq), 0)
                  : protocol::v2::ControlFrame::makeAck(src, dst, static_cast<uint16_t>(seq));
    Bytes b = f.serialize();
    if (static_cast<int>(b.size()) > cap) return -static_cast<int>(b.size());
    std::memcpy(out, b.data(), b.size());
    return static_cast<int>(b.size());
}

// StreamingEncoder::encodeFrame / encodeFrameLight / encodePing (src/gui/modem/streaming_encoder.cpp:209-300, 392-430):
// what a station transmits for one frame.  waveform 1 = OFDM_CHIRP (modem configuration), 2 = MC_DPSK.
// kind 0 = full preamble, 1 = light (connected-mode) preamble, 2 = PING (frame ignored).
int ref_stream_encode(int waveform, int carriers, int spreading, int modulation, int rate, int kind,
                      const uint8_t* frame, int len, float* out, int cap) {
    gui::StreamingEncoder enc;
    if (waveform == 1) {
        enc.setMode(protocol::WaveformMode::OFDM_CHIRP);
        ModemConfig cfg;                                        // the modem's configuration; pilots on, as the engine sets it
        cfg.use_pilots = true;
        cfg.pilot_spacing = 10;
        enc.setOFDMConfig(cfg);
        enc.setDataMode(Modulation::QPSK, CodeRate::R1_2);      // setDataMode returns early when nothing changes: move off the target first
    } else {
        enc.setMCDPSKCarriers(carriers);
        enc.setSpreadingMode(static_cast<SpreadingMode>(spreading));
        enc.setMode(protocol::WaveformMode::MC_DPSK);
    }
    enc.setDataMode(static_cast<Modulation>(modulation), static_cast<CodeRate>(rate));
    Bytes f(frame, frame + len);
    std::vector<float> s = kind == 2 ? enc.encodePing() : kind == 1 ? enc.encodeFrameLight(f) : enc.encodeFrame(f);
    const int n = static_cast<int>(s.size());
    if (n > cap) return -n;
    std::memcpy(out, s.data(), s.size() * sizeof(float));
    return n;
}

// v2::encodeFrameWithLDPC(frame_data, rate): CW0 = first bytes of the frame, CW1+ carry [0xD5][index][payload]
// (frame_v2.cpp:925-957) -- the multi-codeword transmit format of MC-DPSK frames.  Returns the codeword count.
int ref_encode_frame_with_ldpc(const uint8_t* frame, int len, int rate, uint8_t* out /*[n_cw][81]*/, int cap_cw) {
    Bytes f(frame, frame + len);
    auto cws = protocol::v2::encodeFrameWithLDPC(f, static_cast<CodeRate>(rate));
    const int n = static_cast<int>(cws.size());
    if (n > cap_cw) return -n;
    for (int i = 0; i < n; ++i) {
        std::memset(out + i * 81, 0, 81);
        std::memcpy(out + i * 81, cws[i].data(), std::min<size_t>(81, cws[i].size()));
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// MC-DPSK  (src/psk/multi_carrier_dpsk.hpp)
// ---------------------------------------------------------------------------------------------
static MultiCarrierDPSKConfig to_mc(const ria_mcdpsk_config* c) {
    MultiCarrierDPSKConfig m;
    m.sample_rate = c->sample_rate;
    m.num_carriers = static_cast<int>(c->num_carriers);
    m.freq_low = c->freq_low;
    m.freq_high = c->freq_high;
    m.samples_per_symbol = static_cast<int>(c->samples_per_symbol);
    m.bits_per_symbol = static_cast<int>(c->bits_per_symbol);
    m.spreading_mode = c->spreading == 4 ? SpreadingMode::TIME_4X
                     : c->spreading == 2 ? SpreadingMode::TIME_2X : SpreadingMode::NONE;
    m.training_symbols = static_cast<int>(c->training_symbols);
    return m;
}

// TX after the sync preamble: [training][reference][modulate(data)]
// (MultiCarrierDPSKModulator::generateTrainingSequence / generateReferenceSymbol / modulate)
int ref_mcdpsk_tx_frame(const ria_mcdpsk_config* c, const uint8_t* data, int len, float* out, int cap) {
    MultiCarrierDPSKModulator mod(to_mc(c));
    Samples tr = mod.generateTrainingSequence();
    Samples rf = mod.generateReferenceSymbol();
    Samples d = mod.modulate(Bytes(data, data + len));
    int n = static_cast<int>(tr.size() + rf.size() + d.size());
    if (n > cap) return -n;
    std::memcpy(out, tr.data(), tr.size() * 4);
    std::memcpy(out + tr.size(), rf.data(), rf.size() * 4);
    std::memcpy(out + tr.size() + rf.size(), d.data(), d.size() * 4);
    return n;
}

void* ref_mcdpsk_demod_new(const ria_mcdpsk_config* c) { return new MultiCarrierDPSKDemodulator(to_mc(c)); }
void ref_mcdpsk_demod_free(void* h) { delete static_cast<MultiCarrierDPSKDemodulator*>(h); }

// What MCDPSKWaveform::process does (mc_dpsk_waveform.cpp:294-338): setChirpDetected(cfo);
// process(samples); getSoftBits().  `phase` mirrors setCFOWithPhase's initial phase.
int ref_mcdpsk_process(void* h, const float* samples, int n, float cfo_hz, float phase,
                       float* soft, int cap, int* n_soft, float* fading, float* cfo_out) {
    auto* d = static_cast<MultiCarrierDPSKDemodulator*>(h);
    d->reset();
    d->setCFOWithPhase(cfo_hz, phase);
    d->setChirpDetected(cfo_hz);
    bool ready = d->process(SampleSpan(samples, static_cast<size_t>(n)));
    *n_soft = 0;
    if (ready) {
        auto sb = d->getSoftBits();
        *n_soft = static_cast<int>(sb.size());
        if (*n_soft <= cap) std::memcpy(soft, sb.data(), sb.size() * 4);
    }
    if (fading) *fading = d->getFadingIndex();
    if (cfo_out) *cfo_out = d->getEstimatedCFO();
    return ready ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------
// Sync  (src/sync/zc_sync.hpp, src/sync/chirp_sync.hpp)
// ---------------------------------------------------------------------------------------------
static sync::ZCConfig to_zc(const ria_zc_config* c) {
    sync::ZCConfig z;
    z.sample_rate = c->sample_rate; z.sequence_length = c->sequence_length; z.upsample_factor = c->upsample_factor;
    z.num_repetitions = c->num_repetitions; z.carrier_freq = c->carrier_freq; z.gap_ms = c->gap_ms;
    z.root_ping = c->root_ping; z.root_pong = c->root_pong; z.root_data = c->root_data; z.root_control = c->root_control;
    return z;
}

int ref_zc_preamble(const ria_zc_config* c, int frame_type, float* out, int cap) {
    sync::ZCSync zc(to_zc(c));
    Samples p = zc.generatePreamble(static_cast<sync::ZCFrameType>(frame_type));
    int n = static_cast<int>(p.size());
    if (n > cap) return -n;
    std::memcpy(out, p.data(), p.size() * 4);
    return n;
}

void ref_zc_detect(const ria_zc_config* c, const float* samples, int n, float threshold, unsigned root_mask,
                   float known_cfo, ria_sync_result* out) {
    sync::ZCSync zc(to_zc(c));
    auto r = zc.detect(SampleSpan(samples, static_cast<size_t>(n)), threshold, false,
                       static_cast<uint8_t>(root_mask), known_cfo);
    out->detected = r.detected ? 1 : 0;
    out->start_sample = r.start_sample;
    out->correlation = r.correlation;
    out->cfo_hz = r.cfo_hz;
    out->snr_estimate = r.snr_estimate;
    out->root = r.root_detected;
    out->frame_type = static_cast<int>(r.frame_type);
    out->aux = 0;
}

static sync::ChirpSync& chirp_instance() {
    static sync::ChirpSync cs{sync::ChirpConfig{}};    // the defaults both waveforms use
    return cs;
}

int ref_chirp_generate(float* out, int cap) {
    Samples p = chirp_instance().generate();
    int n = static_cast<int>(p.size());
    if (n > cap) return -n;
    std::memcpy(out, p.data(), p.size() * 4);
    return n;
}

void ref_chirp_detect_dual(const float* samples, int n, float threshold, ria_sync_result* out) {
    auto r = chirp_instance().detectDualChirp(SampleSpan(samples, static_cast<size_t>(n)), threshold);
    out->detected = r.success ? 1 : 0;
    out->start_sample = r.up_chirp_start;
    out->correlation = r.up_correlation;
    out->cfo_hz = r.cfo_hz;
    out->snr_estimate = r.down_correlation;
    out->root = 0; out->frame_type = 0;
    out->aux = r.down_chirp_start;
}


// ---------------------------------------------------------------------------------------------
// Channel, chase cache, waveform selection
// ---------------------------------------------------------------------------------------------
void ref_watterson_process(const ria_watterson_config* c, unsigned seed, const float* in, int n, float* out) {
    sim::WattersonChannel::Config cfg;
    cfg.snr_db = c->snr_db; cfg.delay_spread_ms = c->delay_spread_ms; cfg.doppler_spread_hz = c->doppler_spread_hz;
    cfg.path1_gain = c->path1_gain; cfg.path2_gain = c->path2_gain; cfg.sample_rate = c->sample_rate;
    cfg.fading_enabled = c->fading_enabled != 0; cfg.multipath_enabled = c->multipath_enabled != 0;
    cfg.noise_enabled = c->noise_enabled != 0; cfg.cfo_enabled = false;
    sim::WattersonChannel ch(cfg, seed);
    Samples o = ch.process(SampleSpan(in, static_cast<size_t>(n)));
    std::memcpy(out, o.data(), o.size() * 4);
}

void ref_recommend_waveform(float snr_db, float fading, ria_waveform_recommendation* out) {
    auto r = protocol::recommendWaveformAndRate(snr_db, fading);
    out->waveform = static_cast<int>(r.waveform); out->modulation = static_cast<int>(r.modulation);
    out->rate = static_cast<int>(r.rate); out->estimated_throughput_bps = r.estimated_throughput_bps;
    out->num_carriers = r.num_carriers;
    out->spreading = r.spreading == SpreadingMode::TIME_4X ? 4 : r.spreading == SpreadingMode::TIME_2X ? 2 : 1;
}

void ref_recommend_data_mode(float snr_db, int waveform, float fading, ria_waveform_recommendation* out) {
    Modulation mod = Modulation::DQPSK; CodeRate rate = CodeRate::R1_4; int nc = 10; SpreadingMode sp = SpreadingMode::NONE;
    protocol::recommendDataMode(snr_db, static_cast<protocol::WaveformMode>(waveform), mod, rate, fading, &nc, &sp);
    out->waveform = waveform; out->modulation = static_cast<int>(mod); out->rate = static_cast<int>(rate);
    out->estimated_throughput_bps = 0.0f; out->num_carriers = nc;
    out->spreading = sp == SpreadingMode::TIME_4X ? 4 : sp == SpreadingMode::TIME_2X ? 2 : 1;
}

// fec::ChaseCache: store the receptions in order, return the combined soft bits of one codeword
int ref_chase_combine(const float* soft, int n_receptions, int cw_index, int total_cw, float* out, int* count) {
    fec::ChaseCache cache;
    fec::ChaseCacheKey key{1, 0x123456, 0x654321};
    int stored = 0;
    for (int r = 0; r < n_receptions; ++r) {
        std::vector<float> v(soft + r * 648, soft + (r + 1) * 648);
        stored += cache.store(key, cw_index, v, total_cw, protocol::v2::FrameType::DATA) ? 1 : 0;
    }
    auto c = cache.getCombined(key, cw_index);
    *count = cache.getCombineCount(key, cw_index);
    if (c) std::memcpy(out, c->data(), 648 * 4);
    return stored;
}


// OFDM_COX transmit frame: OFDMNvisWaveform::generatePreamble() + modulate() (src/waveform/ofdm_cox_waveform.cpp:106-119),
// i.e. OFDMModulator::generatePreamble (src/ofdm/modulator.cpp:479-532: guard, 4 x STS, 2 x LTS) followed by
// OFDMModulator::modulate with the mixer running on.
int ref_ofdm_cox_tx_frame(const ria_modem_config* c, const uint8_t* data, int len, float* out, int cap) {
    ModemConfig m = to_cfg(c);
    OFDMModulator mod(m);
    Samples pre = mod.generatePreamble();
    Samples d = mod.modulate(ByteSpan(data, static_cast<size_t>(len)), m.modulation);
    int n = static_cast<int>(pre.size() + d.size());
    if (n > cap) return -n;
    std::memcpy(out, pre.data(), pre.size() * sizeof(float));
    std::memcpy(out + pre.size(), d.data(), d.size() * sizeof(float));
    return n;
}

// OFDMDemodulator::searchForSync (src/ofdm/demodulator.cpp:1450-1542) = what OFDMNvisWaveform::detectSync runs.
// noise_floor (optional) is Impl::noise_floor_energy before / after the call: the only state the search keeps.
int ref_ofdm_cox_search_sync(void* h, const float* samples, int n, float threshold, float* noise_floor,
                             long long* position, float* cfo_hz) {
    auto* d = static_cast<RefOfdmDemod*>(h);
    if (noise_floor) d->dem.impl_->noise_floor_energy = *noise_floor;
    size_t pos = 0; float cfo = 0.0f;
    bool found = d->dem.searchForSync(SampleSpan(samples, static_cast<size_t>(n)), pos, cfo, threshold);
    if (noise_floor) *noise_floor = d->dem.impl_->noise_floor_energy;
    *position = found ? static_cast<long long>(pos) : -1;
    *cfo_hz = found ? cfo : 0.0f;
    return found ? 1 : 0;
}

// taps for the acquisition stages: Impl::measureCorrelation / refineLTSTiming / estimateCoarseCFO on a buffer
float ref_ofdm_cox_correlation(void* h, const float* samples, int n, int offset) {
    auto* d = static_cast<RefOfdmDemod*>(h);
    auto& im = *d->dem.impl_;
    std::vector<float> saved = std::move(im.rx_buffer);
    im.rx_buffer.assign(samples, samples + n);
    float c = im.measureCorrelation(static_cast<size_t>(offset));
    im.rx_buffer = std::move(saved);
    return c;
}

long long ref_ofdm_cox_refine_lts(void* h, const float* samples, int n, int coarse_sts, float* cfo_hz) {
    auto* d = static_cast<RefOfdmDemod*>(h);
    auto& im = *d->dem.impl_;
    std::vector<float> saved = std::move(im.rx_buffer);
    im.rx_buffer.assign(samples, samples + n);
    size_t r = im.refineLTSTiming(static_cast<size_t>(coarse_sts));
    if (cfo_hz) *cfo_hz = im.estimateCoarseCFO(static_cast<size_t>(coarse_sts));
    im.rx_buffer = std::move(saved);
    return r == SIZE_MAX ? -1 : static_cast<long long>(r);
}

// OFDMChirpWaveform::detectDataSync (src/waveform/ofdm_chirp_waveform.cpp:207-384)
void ref_ofdm_data_sync(const ria_modem_config* c, const float* samples, int n, float known_cfo, float threshold,
                        ria_sync_result* out) {
    OFDMChirpWaveform wf(to_cfg(c));
    SyncResult r;
    bool det = wf.detectDataSync(SampleSpan(samples, static_cast<size_t>(n)), r, known_cfo, threshold);
    out->detected = det ? 1 : 0;
    out->start_sample = r.start_sample;
    out->correlation = r.correlation;
    out->cfo_hz = r.cfo_hz;
    out->snr_estimate = 0.0f; out->root = 0; out->frame_type = 0;
    out->aux = wf.wasBurstInterleaved() ? 1 : 0;
}

}  // extern "C"
