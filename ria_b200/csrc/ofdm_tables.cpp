// Host-side constant tables of the OFDM receive chain (per ModemConfig), uploaded once per
// context.  Everything here is frame-independent state the reference computes in constructors:
//   - FFT twiddles                      FFT::FFT, src/dsp/fft.cpp:83-91
//   - mixer (NCO) phasor sequence       NCO::NCO / NCO::next, src/dsp/filters.cpp:228-238
//     (the mixer restarts at 0 for every frame, demodulator.cpp:1270, and accumulates its phase
//      in fp32 with a double-promoted wrap, so the sequence is a fixed table, not a closed form)
//   - carrier maps, pilot sequence      Impl::setupCarriers / generateSequences,
//     sync (Zadoff-Chu) sequence        src/ofdm/demodulator.cpp:45-117
//   - pilot interpolation table         Impl::buildInterpTable, src/ofdm/demodulator.cpp:146-202
// The transcendental calls use the host libm in single precision exactly where the reference
// does (std::cos(float) etc.), so the tables are bit-identical to the reference's.

#include "ofdm_tables.h"

#include <cmath>
#include <random>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ria {

int ofdm_cyclic_prefix(const ria_modem_config& c) {
    // ModemConfig::getCyclicPrefix, include/ultra/types.hpp:262-273
    uint32_t base;
    switch (c.cp_mode) {
        case 0: base = 32; break;
        case 1: base = 48; break;
        case 2: base = 64; break;
        default: base = 48;
    }
    return static_cast<int>(base * (c.fft_size / 512));
}

int ofdm_symbol_samples(const ria_modem_config& c) {
    return static_cast<int>(c.fft_size) + ofdm_cyclic_prefix(c) + static_cast<int>(c.symbol_guard);
}

bool ofdm_is_differential(uint32_t mod) {
    return mod == RIA_DBPSK || mod == RIA_DQPSK || mod == RIA_D8PSK;
}

int ofdm_bits_per_carrier(uint32_t mod) {
    // getBitsPerSymbol, include/ultra/types.hpp:42-56
    switch (mod) {
        case RIA_DBPSK: case RIA_BPSK: return 1;
        case RIA_DQPSK: case RIA_QPSK: return 2;
        case RIA_D8PSK: case RIA_QAM8: return 3;
        case RIA_QAM16: return 4;
        case RIA_QAM32: return 5;
        case RIA_QAM64: return 6;
        case RIA_QAM256: return 8;
        default: return 1;
    }
}

const char* ofdm_config_error(const ria_modem_config& c) {
    if (c.fft_size != 1024) return "fft_size must be 1024";
    if (c.num_carriers < 2 || c.num_carriers > 64) return "num_carriers must be in [2, 64]";
    if (c.sample_rate == 0) return "sample_rate must be > 0";
    if (c.training_symbols != 2) return "training_symbols must be 2 (the reference waveform always sends 2 LTS)";
    if (c.use_pilots && c.pilot_spacing == 0) return "pilot_spacing must be > 0";
    if (c.use_pilots && (c.num_carriers + c.pilot_spacing - 1) / c.pilot_spacing > 32)
        return "more than 32 pilot carriers (pilot_spacing must be >= 2)";
    switch (c.modulation) {
        case RIA_DBPSK: case RIA_BPSK: case RIA_DQPSK: case RIA_QPSK: case RIA_D8PSK:
        case RIA_QAM16: case RIA_QAM32: case RIA_QAM64: case RIA_QAM256: break;
        // QAM8 has no demapper in the reference either (demodulateSymbol's default branch emits two QPSK soft bits
        // for a three-bit symbol, demodulator.cpp:404-409)
        default: return "modulation not supported (QAM8 has no demapper in the reference)";
    }
    if (c.cp_mode > 2) return "bad cp_mode";
    return nullptr;
}

void ofdm_build_tables(const ria_modem_config& cfg, int nco_len, OfdmTablesHost& t) {
    const int N = static_cast<int>(cfg.fft_size);
    t.cfg = cfg;
    t.cp = ofdm_cyclic_prefix(cfg);
    t.sym_len = ofdm_symbol_samples(cfg);

    // ---- twiddles (fft.cpp:85-90): angle evaluated in double, rounded to float, then cosf/sinf
    t.twiddle.resize(N / 2);
    for (int k = 0; k < N / 2; ++k) {
        float angle = -2.0f * M_PI * k / static_cast<size_t>(N);
        t.twiddle[k] = make_float2(std::cos(angle), std::sin(angle));
    }

    // ---- mixer phasors (filters.cpp:228-238)
    {
        float phase_inc = 2.0f * M_PI * static_cast<float>(cfg.center_freq) / static_cast<float>(cfg.sample_rate);
        float phase = 0;
        t.nco.resize(nco_len);
        for (int i = 0; i < nco_len; ++i) {
            t.nco[i] = make_float2(std::cos(phase), std::sin(phase));
            phase += phase_inc;
            if (phase > 2.0f * M_PI) phase -= 2.0f * M_PI;
            if (phase < 0) phase += 2.0f * M_PI;
        }
    }

    // ---- carriers (demodulator.cpp:45-74)
    const int nc = static_cast<int>(cfg.num_carriers);
    const int neg_limit = nc / 2, pos_limit = (nc + 1) / 2;
    t.car = OfdmCarrierTable{};
    OfdmCarrierTable& c = t.car;
    c.num_carriers = nc;
    int logical = 0, nd = 0, np = 0;
    for (int i = -neg_limit; i <= pos_limit; ++i) {
        if (i == 0) continue;
        const int fft_idx = (i + N) % N;
        const bool is_pilot = cfg.use_pilots && (logical % static_cast<int>(cfg.pilot_spacing) == 0);
        c.fft_idx[logical] = static_cast<int16_t>(fft_idx);
        c.is_pilot[logical] = is_pilot ? 1 : 0;
        // signed carrier number used for the phase slope (channel_equalizer.cpp:897, 935)
        c.car_k[logical] = static_cast<int16_t>((fft_idx <= N / 2) ? fft_idx : fft_idx - N);
        if (is_pilot) { c.pilot_car[np] = static_cast<int16_t>(logical); c.sub_idx[logical] = static_cast<int16_t>(np); ++np; }
        else          { c.data_car[nd] = static_cast<int16_t>(logical);  c.sub_idx[logical] = static_cast<int16_t>(nd); ++nd; }
        ++logical;
    }
    c.n_data = nd;
    c.n_pilot = np;

    // ---- sequences (demodulator.cpp:78-94): ZC root 1 over num_carriers, +-1 pilots
    {
        const size_t Nz = static_cast<size_t>(nc), u = 1;
        for (int i = 0; i < nd; ++i) {
            const size_t n = static_cast<size_t>(i) % Nz;
            float phase = -M_PI * u * n * (n + 1) / Nz;
            c.tx_data[i] = make_float2(std::cos(phase), std::sin(phase));
        }
        std::mt19937 rng(0x50494C54u);   // PILOT_RNG_SEED, demodulator_constants.hpp:36
        for (int i = 0; i < np; ++i) c.pilot_sign[i] = (rng() & 1) ? 1.0f : -1.0f;
    }

    // ---- interpolation table (demodulator.cpp:146-202); neighbours as indices into the pilot list
    {
        std::vector<int> is_p(nc);
        for (int ci = 0; ci < nc; ++ci) is_p[ci] = (ci % static_cast<int>(cfg.pilot_spacing ? cfg.pilot_spacing : 1) == 0);
        for (int ci = 0; ci < nc; ++ci) {
            if (c.is_pilot[ci] || !cfg.use_pilots) continue;
            const int di = c.sub_idx[ci];
            int lower = -1, upper = -1;
            for (int j = ci - 1; j >= 0; --j) if (is_p[j]) { lower = j; break; }
            for (int j = ci + 1; j < nc; ++j) if (is_p[j]) { upper = j; break; }
            float alpha = 0.5f;
            if (lower >= 0 && upper >= 0) {
                float total = static_cast<float>(upper - lower);
                alpha = (total > 0) ? static_cast<float>(ci - lower) / total : 0.5f;
            }
            c.interp_lo[di] = static_cast<int16_t>(lower >= 0 ? c.sub_idx[lower] : -1);
            c.interp_hi[di] = static_cast<int16_t>(upper >= 0 ? c.sub_idx[upper] : -1);
            c.interp_alpha[di] = alpha;
        }
    }
}

int channel_interleaver_step(int bits_per_symbol, int total_bits) {
    // findCoprimeStep, src/fec/ldpc_decoder.cpp:552-577
    auto gcd = [](size_t a, size_t b) { while (b != 0) { size_t r = a % b; a = b; b = r; } return a; };
    const size_t n = static_cast<size_t>(bits_per_symbol), total = static_cast<size_t>(total_bits);
    size_t target = n * 3;
    if (target >= total) target = total / 2;
    for (size_t s = target; s < total; ++s) if (gcd(s, total) == 1) return static_cast<int>(s);
    for (size_t s = n + 1; s < total; ++s) if (gcd(s, total) == 1) return static_cast<int>(s);
    return static_cast<int>(n + 1);
}

}  // namespace ria

extern "C" int ria_modem_config_for(int modulation, int rate, ria_modem_config* cfg) {
    if (!cfg) return RIA_E_INVAL;
    // recommendedPilotSpacing, include/ultra/ofdm_link_adaptation.hpp:26-64
    const bool coherent = modulation == RIA_BPSK || modulation == RIA_QPSK || modulation == RIA_QAM8 ||
                          modulation == RIA_QAM16 || modulation == RIA_QAM32 || modulation == RIA_QAM64 ||
                          modulation == RIA_QAM256;
    int spacing;
    if (coherent) spacing = (rate == RIA_R5_6 || rate == RIA_R7_8) ? 6 : (rate == RIA_R3_4 ? 8 : 5);
    else if (modulation == RIA_D8PSK) spacing = (rate == RIA_R3_4 || rate == RIA_R2_3 || rate == RIA_R1_2) ? 8 : 10;
    else spacing = (rate == RIA_R3_4) ? 15 : 10;
    *cfg = ria_modem_config{48000, 1500, 1024, 59, 1, 0, 1, static_cast<uint32_t>(spacing),
                            static_cast<uint32_t>(modulation), 2};
    return RIA_OK;
}

extern "C" int ria_ofdm_symbol_samples(const ria_modem_config* cfg) {
    return cfg ? ria::ofdm_symbol_samples(*cfg) : RIA_E_INVAL;
}

extern "C" int ria_ofdm_pilot_carriers(const ria_modem_config* cfg) {
    if (!cfg) return RIA_E_INVAL;
    if (!cfg->use_pilots || cfg->pilot_spacing == 0) return 0;
    return static_cast<int>((cfg->num_carriers + cfg->pilot_spacing - 1) / cfg->pilot_spacing);
}

extern "C" int ria_ofdm_data_carriers(const ria_modem_config* cfg) {
    if (!cfg) return RIA_E_INVAL;
    return static_cast<int>(cfg->num_carriers) - ria_ofdm_pilot_carriers(cfg);
}

extern "C" int ria_channel_interleaver_step(int bits_per_symbol, int total_bits) {
    if (bits_per_symbol <= 0 || total_bits <= 1) return RIA_E_INVAL;
    return ria::channel_interleaver_step(bits_per_symbol, total_bits);
}

extern "C" uint16_t ria_crc16(const uint8_t* data, size_t len) {
    // ControlFrame::calculateCRC, src/protocol/frame_v2.cpp:115-128
    uint16_t crc = 0xFFFF;
    for (size_t i = 0; i < len; ++i) {
        crc ^= static_cast<uint16_t>(static_cast<uint16_t>(data[i]) << 8);
        for (int b = 0; b < 8; ++b)
            crc = (crc & 0x8000) ? static_cast<uint16_t>((crc << 1) ^ 0x1021) : static_cast<uint16_t>(crc << 1);
    }
    return crc;
}
