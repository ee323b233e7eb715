// Waveform / rate selection ladder (host logic, drives the adaptive SNR sweep of configs[4]).
// Mirrors protocol::recommendWaveformAndRate and recommendDataMode
// (src/protocol/waveform_selection.hpp:112-222, 250-314) threshold for threshold.

#include "ria_internal.h"

namespace {

// selectOFDMCodeRate (:36-41)
int select_ofdm_rate(float snr_db, float fading) {
    if (fading < 0.15f && snr_db >= 20.0f) return RIA_R3_4;
    if (fading < 0.65f && snr_db >= 20.0f) return RIA_R2_3;
    if (fading < 1.10f && snr_db >= 15.0f) return RIA_R1_2;
    return RIA_R1_4;
}

float dqpsk_throughput(int rate) {
    return rate == RIA_R3_4 ? 3900.0f : rate == RIA_R2_3 ? 3200.0f : rate == RIA_R1_2 ? 2300.0f : 1150.0f;
}

}  // namespace

extern "C" int ria_recommend_waveform(float snr_db, float fading_index, ria_waveform_recommendation* rec) {
    if (!rec) return RIA_E_INVAL;
    ria_waveform_recommendation r{};
    r.num_carriers = 10;
    r.spreading = 1;
    auto mc = [&](int mod, int spreading, float bps) {
        r.waveform = RIA_WAVEFORM_MC_DPSK; r.modulation = mod; r.rate = RIA_R1_4; r.spreading = spreading;
        r.estimated_throughput_bps = bps; r.num_carriers = 10;
    };
    if (snr_db < -7.0f) mc(RIA_DBPSK, 4, 117.0f);
    else if (snr_db < -3.0f) mc(RIA_DBPSK, 2, 235.0f);
    else if (snr_db < 5.0f) mc(RIA_DBPSK, 1, 469.0f);
    else if (snr_db < 10.0f) mc(RIA_DQPSK, 1, 938.0f);
    else if (fading_index < 0.15f) {
        r.waveform = RIA_WAVEFORM_OFDM_CHIRP;
        if (snr_db >= 25.0f) { r.modulation = RIA_QAM64; r.rate = RIA_R3_4; r.estimated_throughput_bps = 7200.0f; }
        else if (snr_db >= 22.0f) { r.modulation = RIA_QAM32; r.rate = RIA_R3_4; r.estimated_throughput_bps = 6000.0f; }
        else if (snr_db >= 18.0f) {
            r.modulation = RIA_QAM16; r.rate = select_ofdm_rate(snr_db, fading_index);
            r.estimated_throughput_bps = r.rate == RIA_R3_4 ? 4800.0f : r.rate == RIA_R2_3 ? 4000.0f
                                       : r.rate == RIA_R1_2 ? 3000.0f : 1500.0f;
        } else {
            r.modulation = RIA_DQPSK; r.rate = select_ofdm_rate(snr_db, fading_index);
            r.estimated_throughput_bps = dqpsk_throughput(r.rate);
        }
    } else if (fading_index < 0.65f && snr_db >= 10.0f) {
        r.waveform = RIA_WAVEFORM_OFDM_CHIRP;
        if (snr_db >= 22.0f) { r.modulation = RIA_QAM16; r.rate = RIA_R2_3; r.estimated_throughput_bps = 4000.0f; }
        else { r.modulation = RIA_DQPSK; r.rate = select_ofdm_rate(snr_db, fading_index); r.estimated_throughput_bps = dqpsk_throughput(r.rate); }
    } else if (fading_index < 1.10f && snr_db >= 10.0f) {
        r.waveform = RIA_WAVEFORM_OFDM_CHIRP; r.modulation = RIA_DQPSK;
        r.rate = select_ofdm_rate(snr_db, fading_index); r.estimated_throughput_bps = dqpsk_throughput(r.rate);
    } else if (snr_db >= 10.0f) {
        r.waveform = RIA_WAVEFORM_OFDM_CHIRP; r.modulation = RIA_DQPSK; r.rate = RIA_R1_4; r.estimated_throughput_bps = 1150.0f;
    } else {
        mc(RIA_DQPSK, 1, 938.0f);
    }
    *rec = r;
    return RIA_OK;
}

extern "C" int ria_recommend_data_mode(float snr_db, int waveform, float fading_index, ria_waveform_recommendation* rec) {
    if (!rec) return RIA_E_INVAL;
    ria_waveform_recommendation r{};
    r.waveform = waveform; r.num_carriers = 10; r.spreading = 1;
    if (waveform == RIA_WAVEFORM_MC_DPSK) {
        r.rate = RIA_R1_4;
        if (snr_db < -7.0f) { r.modulation = RIA_DBPSK; r.spreading = 4; }
        else if (snr_db < -3.0f) { r.modulation = RIA_DBPSK; r.spreading = 2; }
        else if (snr_db < 5.0f) { r.modulation = RIA_DBPSK; }
        else { r.modulation = RIA_DQPSK; }
        *rec = r;
        return RIA_OK;
    }
    bool done = false;
    if (fading_index < 0.15f) {
        if (snr_db >= 25.0f) { r.modulation = RIA_QAM64; r.rate = RIA_R3_4; done = true; }
        else if (snr_db >= 22.0f) { r.modulation = RIA_QAM32; r.rate = RIA_R3_4; done = true; }
        else if (snr_db >= 18.0f) { r.modulation = RIA_QAM16; r.rate = select_ofdm_rate(snr_db, fading_index); done = true; }
    } else if (fading_index < 0.65f) {
        if (snr_db >= 22.0f) { r.modulation = RIA_QAM16; r.rate = RIA_R2_3; done = true; }
    }
    if (!done) { r.modulation = RIA_DQPSK; r.rate = select_ofdm_rate(snr_db, fading_index); }
    *rec = r;
    return RIA_OK;
}
