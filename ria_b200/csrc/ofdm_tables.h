// Constant tables of the OFDM receive chain (host build + device layout).
#pragma once

#include "ria_internal.h"

namespace ria {

constexpr int kMaxCarriers = 64;

// Copied to the device once per (context, config) and from there into shared memory by each CTA.
struct OfdmCarrierTable {
    int32_t num_carriers;
    int32_t n_data;
    int32_t n_pilot;
    int32_t pad0;
    int16_t fft_idx[kMaxCarriers];     // logical carrier -> FFT bin
    int16_t car_k[kMaxCarriers];       // logical carrier -> signed carrier number
    int16_t sub_idx[kMaxCarriers];     // logical carrier -> index in the data (or pilot) list
    uint8_t is_pilot[kMaxCarriers];
    int16_t data_car[kMaxCarriers];    // data index -> logical carrier
    int16_t pilot_car[kMaxCarriers];   // pilot index -> logical carrier
    int16_t interp_lo[kMaxCarriers];   // data index -> lower pilot (index in pilot list) or -1
    int16_t interp_hi[kMaxCarriers];   // data index -> upper pilot or -1
    float   interp_alpha[kMaxCarriers];
    float2  tx_data[kMaxCarriers];     // LTS value on data carrier i (sync_sequence[i])
    float   pilot_sign[kMaxCarriers];  // pilot_sequence[i] = (+-1, 0)
};

struct OfdmTablesHost {
    ria_modem_config cfg{};
    int cp = 0, sym_len = 0;
    std::vector<float2> twiddle;   // fft_size/2
    std::vector<float2> nco;       // (cos, sin) of the mixer phase at frame sample n
    OfdmCarrierTable car{};
};

int  ofdm_cyclic_prefix(const ria_modem_config& c);
int  ofdm_symbol_samples(const ria_modem_config& c);
bool ofdm_is_differential(uint32_t mod);
int  ofdm_bits_per_carrier(uint32_t mod);
const char* ofdm_config_error(const ria_modem_config& c);   // nullptr = supported
void ofdm_build_tables(const ria_modem_config& cfg, int nco_len, OfdmTablesHost& t);
int  channel_interleaver_step(int bits_per_symbol, int total_bits);

struct OfdmTablesDev {
    bool ready = false;
    ria_modem_config cfg{};
    int cp = 0, sym_len = 0, nco_len = 0;
    float2* twiddle = nullptr;       // stage-major (receive FFT kernel)
    float2* twiddle_nat = nullptr;   // natural order W[k], k < fft_size/2 (transmit IFFT)
    float2* nco = nullptr;
    OfdmCarrierTable* car = nullptr;
    OfdmCarrierTable car_host{};
};

}  // namespace ria
