/* TEST INFRASTRUCTURE ONLY -- see ria_oracle.h.
 *
 * Plain-C restatement of the Schmidl-Cox timing metric of the OFDM_COX acquisition:
 *   FFT (fallback Cooley-Tukey build, no FFTW)      src/dsp/fft.cpp:83-128
 *   OFDMDemodulator::Impl::toAnalytic               src/ofdm/ofdm_sync.cpp:56-84
 *   Impl::measureSchmidlCoxCorrelation              src/ofdm/ofdm_sync.cpp:118-163
 *   Impl::estimateCoarseCFO                         src/ofdm/ofdm_sync.cpp:230-261
 * Float expressions keep the reference's types and order (compiled with -ffp-contract=off, the same
 * glibc cosf / sinf / hypotf / sqrtf the reference calls), so the value is the reference's bit for bit;
 * tests/test_oracle_cox_cpu.py pins it against oracle/_ref and tests/golden/cox_golden.npz. */
#include <math.h>
#include <stdlib.h>

#include "ria_oracle.h"

#define ORC_PI 3.14159265358979323846        /* M_PI of <cmath>; -std=c11 hides the macro */

typedef struct { float re, im; } cf;

/* std::complex<float> operator* for finite operands (libstdc++ / libgcc __mulsc3 main path) */
static cf cmul(cf a, cf b) {
    cf r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

/* fft.cpp:83-88: twiddle k of a size-n transform; the angle is formed in double and rounded to float */
static cf twiddle(int k, int n) {
    const float angle = (float)(-2.0f * ORC_PI * (double)k / (double)n);
    cf w = { cosf(angle), sinf(angle) };
    return w;
}

/* fft.cpp:96-128: in-place radix-2 decimation in time, natural-order input */
static void fft_inplace(cf* x, int n, const cf* tw, int inverse) {
    for (int i = 0, j = 0; i < n - 1; ++i) {                 /* bit-reversal permutation (:98-105) */
        if (i < j) { const cf t = x[i]; x[i] = x[j]; x[j] = t; }
        int k = n / 2;
        while (k <= j) { j -= k; k /= 2; }
        j += k;
    }
    for (int len = 2; len <= n; len *= 2) {                   /* butterflies (:107-120) */
        const int half = len / 2, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; ++k) {
                cf w = tw[k * step];
                if (inverse) w.im = -w.im;
                const cf t = cmul(w, x[i + k + half]);
                const cf u = x[i + k];
                x[i + k + half].re = u.re - t.re; x[i + k + half].im = u.im - t.im;
                x[i + k].re = u.re + t.re;        x[i + k].im = u.im + t.im;
            }
    }
    if (inverse) {                                            /* :122-127 */
        const float scale = 1.0f / (float)n;
        for (int i = 0; i < n; ++i) { x[i].re *= scale; x[i].im *= scale; }
    }
}

/* the FFT window at s -> analytic signal in x (toAnalytic, ofdm_sync.cpp:56-84), optionally after the DC removal of
 * measureSchmidlCoxCorrelation (:130-140).  Returns 0 on allocation failure. */
static cf* analytic_window(const float* s, int fft_len, int remove_dc) {
    cf* x = (cf*)malloc(sizeof(cf) * (size_t)fft_len);
    cf* tw = (cf*)malloc(sizeof(cf) * (size_t)(fft_len / 2));
    if (!x || !tw) { free(x); free(tw); return 0; }
    for (int k = 0; k < fft_len / 2; ++k) tw[k] = twiddle(k, fft_len);
    float dc = 0.0f;
    if (remove_dc) {
        float dc_sum = 0.0f;                                  /* a sequential fp32 sum */
        for (int i = 0; i < fft_len; ++i) dc_sum += s[i];
        dc = dc_sum / (float)fft_len;
    }
    for (int i = 0; i < fft_len; ++i) { x[i].re = remove_dc ? s[i] - dc : s[i]; x[i].im = 0.0f; }
    fft_inplace(x, fft_len, tw, 0);
    for (int i = 1; i < fft_len / 2; ++i) { x[i].re *= 2.0f; x[i].im *= 2.0f; }     /* :72-74 */
    for (int i = fft_len / 2 + 1; i < fft_len; ++i) { x[i].re = 0.0f; x[i].im = 0.0f; }   /* :75-77 */
    fft_inplace(x, fft_len, tw, 1);
    free(tw);
    return x;
}

int orc_cox_correlation(const float* samples, int n_samples, int offset, int cp_len, int fft_len,
                        float* metric, float* p_re, float* p_im, float* r1, float* r2) {
    *metric = 0.0f; *p_re = 0.0f; *p_im = 0.0f; *r1 = 0.0f; *r2 = 0.0f;
    if (fft_len < 2 || (fft_len & (fft_len - 1)) != 0) return -1;
    if (offset < 0 || (long long)offset + cp_len + fft_len > n_samples) return 0;      /* :123-126 */
    cf* x = analytic_window(samples + offset + cp_len, fft_len, 1);
    if (!x) return -1;
    const int half = fft_len / 2;
    float pr = 0.0f, pi = 0.0f, e1 = 0.0f, e2 = 0.0f;         /* :145-153 */
    for (int i = 0; i < half; ++i) {
        const cf a = x[i], b = x[i + half];
        cf ca = { a.re, -a.im };
        const cf t = cmul(ca, b);
        pr += t.re; pi += t.im;
        e1 += a.re * a.re + a.im * a.im;                      /* std::norm */
        e2 += b.re * b.re + b.im * b.im;
    }
    free(x);
    *p_re = pr; *p_im = pi; *r1 = e1; *r2 = e2;
    const float normalization = sqrtf(e1 * e2);               /* :158-163 */
    if (normalization < 1e-10f) return 1;
    *metric = hypotf(pr, pi) / normalization;                 /* std::abs(complex<float>) = cabsf = hypotf */
    return 1;
}

/* Impl::estimateCoarseCFO (ofdm_sync.cpp:230-261): no DC removal here, P over the two halves, CFO from its angle */
float orc_cox_coarse_cfo(const float* samples, int n_samples, int sync_offset, int cp_len, int fft_len,
                         unsigned sample_rate) {
    if (fft_len < 2 || (fft_len & (fft_len - 1)) != 0) return 0.0f;
    if (sync_offset < 0 || (long long)sync_offset + cp_len + fft_len > n_samples) return 0.0f;   /* :237-239 */
    cf* x = analytic_window(samples + sync_offset + cp_len, fft_len, 0);
    if (!x) return 0.0f;
    const int half = fft_len / 2;
    float pr = 0.0f, pi = 0.0f;
    for (int i = 0; i < half; ++i) {
        cf ca = { x[i].re, -x[i].im };
        const cf t = cmul(ca, x[i + half]);
        pr += t.re; pi += t.im;
    }
    free(x);
    const float phase = atan2f(pi, pr);                                             /* :248 */
    float cfo_hz = (float)((double)(phase * (float)sample_rate) / (ORC_PI * (double)fft_len));   /* :251 */
    const float max_cfo = (float)(sample_rate / (unsigned)fft_len);                 /* :254, an integer division */
    if (cfo_hz > max_cfo) cfo_hz = max_cfo;                                         /* std::min(max_cfo, cfo) */
    if (cfo_hz < -max_cfo) cfo_hz = -max_cfo;                                       /* std::max(-max_cfo, ...) */
    return cfo_hz;
}
