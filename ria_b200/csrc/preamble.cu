// Sync preambles of the transmitter: the Zadoff-Chu preamble of a root
// (sync::ZCSync::generatePreambleForRoot, src/sync/zc_sync.hpp:133-190 with generateZC :420-436)
// and the dual up/down chirp (sync::ChirpSync::generate, src/sync/chirp_sync.hpp:61-108).
//
// Both are closed-form per sample, so every output sample is an independent evaluation here: one
// thread per sample on the device (`*_dev`, used by the batched channel kernels as their TX pool),
// and the same __host__ __device__ sample functions in a plain loop for the `*_host` entry points
// (CPU tests; no GPU needed).  What has to agree with the reference is the value of each float
// expression: where the reference promotes through double (M_PI is a double constant) this does
// too, products and sums stay un-contracted, and sinf / cosf are glibc's algorithms from
// rn_math.h -- including reduce_large, because the carrier and chirp phases reach several thousand
// radians.  The ZC preamble's 0.8 / max|x| normalisation needs one reduction: an atomic max over
// the float bit patterns (non-negative floats order like unsigned integers), then a scaling pass.

#include "ria_internal.h"
#include "rn_math.h"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr float kZcAmplitude = 0.8f;                 // ZC_AMPLITUDE_SCALE (zc_sync.hpp:51)
constexpr float kChirpAmplitude = 0.5f;              // ChirpConfig::amplitude (chirp_sync.hpp:36)

struct ZcGeometry {
    int chips, upsample, reps, rep_len, gap, body, total;
};

__host__ __device__ inline ZcGeometry zc_geometry(const ria_zc_config& c) {
    ZcGeometry g;
    g.chips = c.sequence_length; g.upsample = c.upsample_factor; g.reps = c.num_repetitions;
    g.rep_len = g.chips * g.upsample;
    g.gap = static_cast<int>(c.sample_rate * c.gap_ms / 1000.0f);
    g.body = g.rep_len * g.reps;
    g.total = g.body + g.gap;
    return g;
}

// chip n of the root-u Zadoff-Chu sequence as a unit phasor (cos, sin); the phase is formed in
// double (the reference's -M_PI * u * n * n / N) and rounded to float before cosf / sinf
__host__ __device__ inline void zc_chip(int root, int n, int chips, float* re, float* im) {
    const double q = (chips % 2 == 0) ? static_cast<double>(n) : static_cast<double>(n + 1);
    const float phase = static_cast<float>(-kPi * root * n * q / chips);
    *re = glibc_cosf(phase);
    *im = glibc_sinf(phase);
}

// sample i of the preamble body before the peak normalisation: linear interpolation between the two
// neighbouring chips, then up-conversion to the carrier
__host__ __device__ inline float zc_body_sample(const ria_zc_config& c, const ZcGeometry& g, int root, int i) {
    const int within = i % g.rep_len;
    const float chip_pos = rn_div(static_cast<float>(within), static_cast<float>(g.upsample));
    const int chip = static_cast<int>(chip_pos);
    const float frac = rn_add(chip_pos, -static_cast<float>(chip));
    float re, im;
    zc_chip(root, chip, g.chips, &re, &im);
    if (chip < g.chips - 1) {
        float re1, im1;
        zc_chip(root, chip + 1, g.chips, &re1, &im1);
        const float keep = rn_add(1.0f, -frac);
        re = rn_add(rn_mul(re, keep), rn_mul(re1, frac));
        im = rn_add(rn_mul(im, keep), rn_mul(im1, frac));
    }
    const float t = rn_div(static_cast<float>(i), c.sample_rate);
    const float carrier = static_cast<float>(static_cast<double>(2.0f) * kPi * c.carrier_freq * t);
    float sn, cs;
    glibc_sincosf(carrier, &sn, &cs);
    return rn_add(rn_mul(re, cs), -rn_mul(im, sn));
}

struct ChirpGeometry {
    int chirp, gap, total;
    float rate;                                       // (f_end - f_start) / T, Hz per second
};

__host__ __device__ inline ChirpGeometry chirp_geometry(const ria_chirp_config& c) {
    ChirpGeometry g;
    g.chirp = static_cast<int>(static_cast<size_t>(c.sample_rate * c.duration_ms / 1000.0f));
    g.gap = static_cast<int>(static_cast<size_t>(c.sample_rate * c.gap_ms / 1000.0f));
    g.total = 2 * g.chirp + 2 * g.gap;
    g.rate = rn_div(rn_add(c.f_end, -c.f_start), rn_div(c.duration_ms, 1000.0f));
    return g;
}

// sample i of [up chirp][gap][down chirp][gap]: amplitude * sin(2 pi (f0 t +- rate t^2 / 2)), the
// instantaneous-phase polynomial in float, the 2 pi factor in double
__host__ __device__ inline float chirp_sample(const ria_chirp_config& c, const ChirpGeometry& g, int i) {
    const int second = g.chirp + g.gap;
    int local; float f0, sgn;
    if (i < g.chirp) { local = i; f0 = c.f_start; sgn = 1.0f; }
    else if (i >= second && i < second + g.chirp) { local = i - second; f0 = c.f_end; sgn = -1.0f; }
    else return 0.0f;
    const float t = rn_div(static_cast<float>(local), c.sample_rate);
    const float quad = rn_mul(rn_mul(rn_mul(0.5f, g.rate), t), t);
    const float cycles = rn_add(rn_mul(f0, t), sgn * quad);
    const float phase = static_cast<float>(static_cast<double>(2.0f) * kPi * cycles);
    return rn_mul(kChirpAmplitude, glibc_sinf(phase));
}

__global__ void zc_body_kernel(ria_zc_config c, int root, float* out, unsigned int* peak_bits) {
    const ZcGeometry g = zc_geometry(c);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float mag = 0.0f;
    if (i < g.body) {
        const float v = zc_body_sample(c, g, root, i);
        out[i] = v;
        mag = fabsf(v);
    } else if (i < g.total) {
        out[i] = 0.0f;
    }
    // warp max first, one atomic per warp
    unsigned int bits = __float_as_uint(mag);
    for (int d = 16; d > 0; d >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, d));
    if ((threadIdx.x & 31) == 0 && bits) atomicMax(peak_bits, bits);
}

__global__ void zc_scale_kernel(float* out, int n, const unsigned int* peak_bits) {
    const float peak = __uint_as_float(*peak_bits);
    if (!(peak > 0.0f)) return;
    const float scale = __fdiv_rn(kZcAmplitude, peak);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fmul_rn(out[i], scale);
}

__global__ void chirp_generate_kernel(ria_chirp_config c, float* out) {
    const ChirpGeometry g = chirp_geometry(c);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.total) out[i] = chirp_sample(c, g, i);
}

bool zc_config_ok(const ria_zc_config* c) {
    return c && c->sequence_length >= 1 && c->upsample_factor >= 1 && c->num_repetitions >= 1 && c->sample_rate > 0.0f;
}

}  // namespace

extern "C" int ria_zc_preamble_samples(const ria_zc_config* cfg) {
    if (!zc_config_ok(cfg)) return RIA_E_INVAL;
    return zc_geometry(*cfg).total;
}

extern "C" int ria_chirp_generate_samples(const ria_chirp_config* cfg) {
    if (!cfg || !(cfg->sample_rate > 0.0f)) return RIA_E_INVAL;
    return chirp_geometry(*cfg).total;
}

extern "C" int ria_zc_preamble_host(const ria_zc_config* cfg, int root, float* out, int cap) {
    if (!zc_config_ok(cfg)) return RIA_E_INVAL;
    const ZcGeometry g = zc_geometry(*cfg);
    if (!out || cap < g.total) return -g.total;
    float peak = 0.0f;
    for (int i = 0; i < g.body; ++i) {
        out[i] = zc_body_sample(*cfg, g, root, i);
        peak = fmaxf(peak, fabsf(out[i]));
    }
    if (peak > 0.0f) {
        const float scale = kZcAmplitude / peak;
        for (int i = 0; i < g.body; ++i) out[i] *= scale;
    }
    for (int i = g.body; i < g.total; ++i) out[i] = 0.0f;
    return g.total;
}

extern "C" int ria_chirp_generate_host(const ria_chirp_config* cfg, float* out, int cap) {
    if (!cfg || !(cfg->sample_rate > 0.0f)) return RIA_E_INVAL;
    const ChirpGeometry g = chirp_geometry(*cfg);
    if (!out || cap < g.total) return -g.total;
    for (int i = 0; i < g.total; ++i) out[i] = chirp_sample(*cfg, g, i);
    return g.total;
}

extern "C" int ria_zc_preamble_dev(ria_ctx* ctx, const ria_zc_config* cfg, int root, float* out_dev, int cap) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (!zc_config_ok(cfg)) return set_error(ctx, RIA_E_INVAL, "zc preamble: bad config");
    const ZcGeometry g = zc_geometry(*cfg);
    if (!out_dev || cap < g.total) return set_error(ctx, RIA_E_INVAL, "zc preamble: need %d samples", g.total);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    unsigned int* peak = ctx->work_counter + 15;      // slot 15 of the per-context scratch words
    RIA_CUDA(ctx, cudaMemsetAsync(peak, 0, sizeof(unsigned int), s));
    const int threads = 256, blocks = (g.total + threads - 1) / threads;
    zc_body_kernel<<<blocks, threads, 0, s>>>(*cfg, root, out_dev, peak);
    zc_scale_kernel<<<(g.body + threads - 1) / threads, threads, 0, s>>>(out_dev, g.body, peak);
    ctx->launches += 2;
    RIA_CUDA(ctx, cudaGetLastError());
    return g.total;
}

extern "C" int ria_chirp_generate_dev(ria_ctx* ctx, const ria_chirp_config* cfg, float* out_dev, int cap) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (!cfg || !(cfg->sample_rate > 0.0f)) return set_error(ctx, RIA_E_INVAL, "chirp generate: bad config");
    const ChirpGeometry g = chirp_geometry(*cfg);
    if (!out_dev || cap < g.total) return set_error(ctx, RIA_E_INVAL, "chirp generate: need %d samples", g.total);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = 256;
    chirp_generate_kernel<<<(g.total + threads - 1) / threads, threads, 0, ctx->stream>>>(*cfg, out_dev);
    ctx->launches += 1;
    RIA_CUDA(ctx, cudaGetLastError());
    return g.total;
}
