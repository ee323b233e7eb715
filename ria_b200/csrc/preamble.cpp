// Sync preambles of the reference transmitter, evaluated on the host.
//
// sync::ZCSync::generatePreambleForRoot (src/sync/zc_sync.hpp:133-190, generateZC :420-436) and
// sync::ChirpSync::generate (src/sync/chirp_sync.hpp:61-108) do not depend on the payload: a batch
// needs each of them once.  They call sinf/cosf on phases of several thousand radians, where the
// container's libm uses its large-argument reduction; evaluating them here with the reference's own
// float expressions and the same libm gives the reference's samples bit for bit
// (tests/test_txsynth_cpu.py), and the batched channel kernels take the result as their TX pool.

#include "ria_internal.h"

#include <cmath>
#include <complex>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

extern "C" int ria_zc_preamble_host(const ria_zc_config* cfg, int root, float* out, int cap) {
    if (!cfg || cfg->sequence_length < 1 || cfg->upsample_factor < 1 || cfg->num_repetitions < 1) return RIA_E_INVAL;
    using Complex = std::complex<float>;
    const int length = cfg->sequence_length, upsample = cfg->upsample_factor;
    const int single_rep_len = length * upsample;
    const int gap = static_cast<int>(cfg->sample_rate * cfg->gap_ms / 1000.0f);
    const int total = single_rep_len * cfg->num_repetitions + gap;
    if (!out || cap < total) return -total;
    // generateZC (:420-436)
    std::vector<Complex> zc(length);
    const bool even = (length % 2 == 0);
    for (int n = 0; n < length; ++n) {
        float phase;
        if (even) phase = -M_PI * root * n * n / length;
        else phase = -M_PI * root * n * (n + 1) / length;
        zc[n] = Complex(std::cos(phase), std::sin(phase));
    }
    int w = 0;
    for (int rep = 0; rep < cfg->num_repetitions; rep++) {
        for (int i = 0; i < single_rep_len; ++i) {
            float chip_pos = static_cast<float>(i) / upsample;
            int chip_idx = static_cast<int>(chip_pos);
            float frac = chip_pos - chip_idx;
            Complex interp;
            if (chip_idx < length - 1) interp = zc[chip_idx] * (1.0f - frac) + zc[chip_idx + 1] * frac;
            else interp = zc[chip_idx];
            int global_i = rep * single_rep_len + i;
            float t = static_cast<float>(global_i) / cfg->sample_rate;
            float carrier_phase = 2.0f * M_PI * cfg->carrier_freq * t;
            float sample = interp.real() * std::cos(carrier_phase) - interp.imag() * std::sin(carrier_phase);
            out[w++] = sample;
        }
    }
    float max_amp = 0.0f;
    for (int i = 0; i < w; ++i) max_amp = std::max(max_amp, std::abs(out[i]));
    if (max_amp > 0.0f) {
        float scale = 0.8f / max_amp;                       // ZC_AMPLITUDE_SCALE (:51)
        for (int i = 0; i < w; ++i) out[i] *= scale;
    }
    for (int i = 0; i < gap; i++) out[w++] = 0.0f;
    return w;
}

extern "C" int ria_chirp_generate_host(const ria_chirp_config* cfg, float* out, int cap) {
    if (!cfg) return RIA_E_INVAL;
    // ChirpSync::generate (:61-108), dual chirp, amplitude 0.5, no TX frequency offset
    size_t chirp_samples = static_cast<size_t>(cfg->sample_rate * cfg->duration_ms / 1000.0f);
    size_t gap_samples = static_cast<size_t>(cfg->sample_rate * cfg->gap_ms / 1000.0f);
    const size_t total = 2 * chirp_samples + 2 * gap_samples;
    if (!out || static_cast<size_t>(cap) < total) return -static_cast<int>(total);
    for (size_t i = 0; i < total; ++i) out[i] = 0.0f;
    const float amplitude = 0.5f;
    float T = cfg->duration_ms / 1000.0f;
    float k = (cfg->f_end - cfg->f_start) / T;
    float cfo = 0.0f;
    float f_start_up = cfg->f_start + cfo;
    for (size_t i = 0; i < chirp_samples; i++) {
        float t = static_cast<float>(i) / cfg->sample_rate;
        float phase = 2.0f * M_PI * (f_start_up * t + 0.5f * k * t * t);
        out[i] = amplitude * std::sin(phase);
    }
    size_t down_start = chirp_samples + gap_samples;
    float f_start_down = cfg->f_end + cfo;
    for (size_t i = 0; i < chirp_samples; i++) {
        float t = static_cast<float>(i) / cfg->sample_rate;
        float phase = 2.0f * M_PI * (f_start_down * t - 0.5f * k * t * t);
        out[down_start + i] = amplitude * std::sin(phase);
    }
    return static_cast<int>(total);
}
