/* ria_b200.h -- C ABI of the B200-native batched RIA receive chain (libria_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry
 * point names the reference interface (file:line under the reference tree) it replaces.
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md; the C++
 * adapters mirroring the reference classes live in include/ria_b200_adapters.hpp.
 *
 * Conventions
 *   - return value: 0 = ok, negative = error (RIA_E_*); ria_last_error() gives the text.
 *     No exception ever crosses this boundary.
 *   - one ria_ctx per GPU, bound to a CUDA stream; thread-compatible, not thread-safe
 *     (matches the reference: one waveform/decoder instance is driven by one thread,
 *     src/gui/modem/streaming_decoder.cpp:718-723).
 *   - "_dev" pointers are device pointers on the context's GPU and the call is asynchronous on
 *     the context stream; "_host" entry points take host pointers, stage through pinned
 *     buffers owned by the context and return after the results are in the host buffers.
 *   - there is NO CPU fallback: without a usable GPU ria_ctx_create fails.
 */
#ifndef RIA_B200_H
#define RIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIA_OK            0
#define RIA_E_INVAL      (-1)   /* bad argument                                  */
#define RIA_E_CUDA       (-2)   /* CUDA runtime error (see ria_last_error)        */
#define RIA_E_NOMEM      (-3)
#define RIA_E_UNSUPPORTED (-4)

#define RIA_LDPC_N        648   /* codeword bits, src/fec/ldpc_codec.hpp:83       */

/* ultra::CodeRate, include/ultra/types.hpp:91-100 (same numeric values) */
typedef enum {
    RIA_R1_4 = 0, RIA_R1_3 = 1, RIA_R1_2 = 2, RIA_R2_3 = 3, RIA_R3_4 = 4, RIA_R5_6 = 5, RIA_R7_8 = 6
} ria_code_rate;

/* ultra::Modulation, include/ultra/types.hpp:27-39 (same numeric values) */
typedef enum {
    RIA_DBPSK = 0, RIA_BPSK = 1, RIA_DQPSK = 2, RIA_QPSK = 3, RIA_D8PSK = 4, RIA_QAM8 = 5,
    RIA_QAM16 = 6, RIA_QAM32 = 7, RIA_QAM64 = 8, RIA_QAM256 = 10
} ria_modulation;

typedef struct ria_ctx ria_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int  ria_ctx_create(int device, ria_ctx** out);
int  ria_ctx_destroy(ria_ctx* ctx);
/* Bind the context to an existing cudaStream_t (e.g. torch's current stream).  NULL is the legacy
 * default stream (what torch uses unless told otherwise).  Until this is called the context issues
 * its work on a private non-blocking stream. */
int  ria_ctx_set_stream(ria_ctx* ctx, void* cuda_stream);
int  ria_ctx_synchronize(ria_ctx* ctx);
const char* ria_last_error(const ria_ctx* ctx);
const char* ria_version(void);
/* Per-kernel timing for bench.py: when enabled, every kernel launch is bracketed by CUDA events on
 * the context stream; get_timing sums the elapsed time of all launches of one kind since enable.
 * kinds: 0 LDPC, 1 OFDM demod, 2 frame status/CRC, 3 AWGN channel, 4 MC-DPSK demod, 5 ZC sync,
 * 6 chirp sync, 7 chase combine, 8 Watterson channel, 9 MC-DPSK CFO correction, 10 OFDM data sync,
 * 11 OFDM FFT stage, 12 OFDM carrier stage, 13 OFDM CFO phase scan, 14 LDPC retry ladder,
 * 15 false-positive frame repair. */
int  ria_ctx_set_timing(ria_ctx* ctx, int enable);
/* Decode options of the frame entry points (ria_frame_decode_batch_dev, ria_ofdm_rx_frames_dev/_host):
 *   RIA_DECODE_RETRY_LADDER  run the LDPC retry ladder of v2::decodeFixedFrame
 *                            (src/protocol/frame_v2.cpp:1389-1546: 4 min-sum factors, then 34 seeded
 *                            soft-bit perturbations) on every codeword whose first decode failed,
 *                            including the reference's carry-over of the decoder's min-sum factor to
 *                            the following codewords of the frame.
 *   RIA_DECODE_FP_REPAIR     run the "LDPC false positive recovery" (:1558-1916) on every frame whose
 *                            four codewords pass parity while the reassembled frame does not verify:
 *                            CRC-guided 1..4-bit flips, re-decode with other min-sum factors, else all
 *                            four codewords are marked failed.
 * Default 0 = first pass only (:1335-1385).  RIA_DECODE_FULL = the complete v2::decodeFixedFrame. */
#define RIA_DECODE_RETRY_LADDER 1
#define RIA_DECODE_FP_REPAIR    2
#define RIA_DECODE_FULL         3
int  ria_ctx_set_decode_flags(ria_ctx* ctx, int flags);
int  ria_ctx_get_decode_flags(const ria_ctx* ctx);
int  ria_ctx_get_timing(ria_ctx* ctx, int kind, double* total_ms, int64_t* launches);
/* number of kernels this library has launched on the context since creation */
int64_t ria_ctx_launch_count(const ria_ctx* ctx);

/* ---- LDPC --------------------------------------------------------------------------------- */
/* Code dimensions for a rate: getCodeParams, src/fec/ldpc_decoder.cpp:21-36. */
int ria_ldpc_params(int rate, int* k_info, int* m_parity, int* n_edges);

/* Copy out the parity-check structure the library generated for `rate` (row_ptr has m+1
 * entries, edge_var n_edges entries, H_rows order) -- buildMatrix, src/fec/ldpc_decoder.cpp:65-138.
 * Used by the CPU tests to pin the table generation without a GPU. */
int ria_ldpc_get_matrix(int rate, int32_t* row_ptr, int32_t* edge_var);

/* Batched replacement for LDPCDecoder::decodeSoft (src/fec/ldpc_decoder.cpp:284-429 / decodeBP
 * :154-260) as used per codeword by v2::decodeFixedFrame (src/protocol/frame_v2.cpp:1359-1385)
 * and robustDecodeSingleCW (src/gui/modem/streaming_decoder.cpp:1028-1058).
 *   llr_dev   [n_cw][648] fp32, one codeword per row
 *   info_dev  [n_cw][info_stride] bytes; first ceil(k/8) bytes = info bits MSB-first, last byte
 *             left-aligned (ldpc_decoder.cpp:240-257); info_stride >= ceil(k/8)
 *   ok_dev    [n_cw] 1 = parity satisfied (lastDecodeSuccess)
 *   iters_dev [n_cw] lastIterations(): zero-based index of the succeeding iteration, or
 *             max_iter on failure
 * max_iter / min_sum_factor: setMaxIterations / setMinSumFactor (ldpc_decoder.cpp:449-455). */
int ria_ldpc_decode_batch_dev(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor,
                              const float* llr_dev, int64_t n_cw,
                              uint8_t* info_dev, int info_stride,
                              uint8_t* ok_dev, int32_t* iters_dev);

/* Same, host buffers (pageable or pinned): chunked H2D -> decode -> D2H inside the call. */
int ria_ldpc_decode_batch_host(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor,
                               const float* llr, int64_t n_cw,
                               uint8_t* info, int info_stride, uint8_t* ok, int32_t* iters);

/* Batched robustDecodeSingleCW (src/gui/modem/streaming_decoder.cpp:1028-1058): decodeSoft with
 * factor 0.9375 and LDPCCodec::getRecommendedIterations(rate); codewords that fail are retried with
 * the min-sum factors 0.875, 0.75, 0.625, 0.5 in that order (first success wins).
 *   attempt_dev  optional [n_cw]: 0 = first decode succeeded, 1..4 = retry that succeeded, 255 = all failed
 *   iters_dev    iterations of the decode that produced the result (first decode when all failed) */
int ria_ldpc_robust_decode_batch_dev(ria_ctx* ctx, int rate, const float* llr_dev, int64_t n_cw,
                                     uint8_t* info_dev, int info_stride,
                                     uint8_t* ok_dev, int32_t* iters_dev, uint8_t* attempt_dev);

/* The soft bits that attempt `attempt` (1..38) of the decodeFixedFrame retry ladder hands to the
 * decoder (src/protocol/frame_v2.cpp:1409-1542): attempts 1-4 the soft bits themselves, 5-38 the
 * clipped / scaled / hard-limited soft bits plus std::normal_distribution<float> noise drawn from
 * std::mt19937(hash(first 16 soft bits) + f(attempt)).  llr_dev / out_dev: [n_cw][648]. */
int ria_ldpc_ladder_perturb_dev(ria_ctx* ctx, const float* llr_dev, int64_t n_cw, int attempt, float* out_dev);


/* ---- OFDM presynced receive chain ---------------------------------------------------------- */
/* POD mirror of the RX-relevant fields of ultra::ModemConfig (include/ultra/types.hpp:193-289).
 * Defaults of the reference: sample_rate 48000, center_freq 1500, fft_size 1024,
 * num_carriers 59, cp_mode 1 (MEDIUM -> 48*(fft/512) = 96 samples), symbol_guard 0. */
typedef struct {
    uint32_t sample_rate;
    uint32_t center_freq;
    uint32_t fft_size;        /* 1024 (the only size the kernels are built for)          */
    uint32_t num_carriers;    /* <= 64                                                   */
    uint32_t cp_mode;         /* CyclicPrefixMode: 0 SHORT, 1 MEDIUM, 2 LONG             */
    uint32_t symbol_guard;
    uint32_t use_pilots;      /* ModemConfig::use_pilots                                 */
    uint32_t pilot_spacing;   /* ModemConfig::pilot_spacing (pilot every N-th carrier)   */
    uint32_t modulation;      /* ria_modulation                                          */
    uint32_t training_symbols;/* LTS symbols in front of the data (reference always 2)   */
} ria_modem_config;

/* Fill `cfg` with the reference defaults for a (modulation, code rate) pair the way
 * OFDMChirpWaveform::configure does (src/waveform/ofdm_chirp_waveform.cpp:79-105,
 * ofdm_link_adaptation.hpp:26-64): use_pilots = 1, pilot_spacing = recommendedPilotSpacing. */
int ria_modem_config_for(int modulation, int rate, ria_modem_config* cfg);
/* samples per OFDM symbol (ModemConfig::getSymbolDuration) and data carriers of a config */
int ria_ofdm_symbol_samples(const ria_modem_config* cfg);
int ria_ofdm_data_carriers(const ria_modem_config* cfg);
int ria_ofdm_pilot_carriers(const ria_modem_config* cfg);

/* Batched replacement for OFDMChirpWaveform::process (src/waveform/ofdm_chirp_waveform.cpp:
 * 391-468) = setFrequencyOffsetWithPhase + OFDMDemodulator::processPresynced
 * (src/ofdm/demodulator.cpp:1250-1414; toBaseband/extractSymbol/estimateChannelFromLTS/
 * updateChannelEstimate/equalize in src/ofdm/channel_equalizer.cpp; demodulateSymbol and
 * soft_demap in src/ofdm/demodulator.cpp:208-508, src/ofdm/soft_demap.hpp).
 *
 *   samples_dev   fp32, frame f starts at samples_dev + f*frame_stride and holds frame_len
 *                 samples beginning at the first LTS symbol (what process() is handed)
 *   cfo_hz_dev    [n] per-frame CFO passed to setFrequencyOffsetWithPhase (may be NULL = 0)
 *   phase_dev     [n] per-frame initial correction phase in radians (may be NULL = 0)
 *   llr_dev       [n][llr_stride] soft bits in the reference's order (symbol-major, carrier,
 *                 bit); positions >= n_llr[f] are zero
 *   n_llr_dev     [n] soft bits produced = data_symbols * data_carriers * bits_per_carrier
 *   snr_db_dev    [n] OFDMDemodulator::getEstimatedSNR()   (may be NULL)
 *   cfo_out_dev   [n] OFDMDemodulator::getFrequencyOffset() after LTS refinement (may be NULL)
 *   fading_dev    [n] OFDMDemodulator::getFadingIndex()    (may be NULL)
 * Frames shorter than one symbol yield n_llr = 0 (processPresynced returns false). */
int ria_ofdm_presynced_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                 const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                 const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                 float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                 float* snr_db_dev, float* cfo_out_dev, float* fading_dev);

/* Same with HOST buffers (H2D -> kernel -> D2H inside the call). */
int ria_ofdm_presynced_batch_host(ria_ctx* ctx, const ria_modem_config* cfg,
                                  const float* samples, int64_t frame_stride, int32_t frame_len,
                                  const float* cfo_hz, const float* phase, int64_t n_frames,
                                  float* llr, int32_t llr_stride, int32_t* n_llr,
                                  float* snr_db, float* cfo_out, float* fading);

/* Debug/parity taps of the same kernel: frequency-domain bins of every symbol for the used
 * carriers ([n][n_symbols][num_carriers] interleaved re,im; logical carrier order) and the
 * channel estimate after the LTS ([n][num_carriers] re,im).  Either may be NULL. */
int ria_ofdm_presynced_batch_taps_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                      const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                      const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                      float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                      float* snr_db_dev, float* cfo_out_dev, float* fading_dev,
                                      float* bins_dev, float* h_lts_dev);

/* ---- fixed 4-codeword frame: de-interleave + LDPC + header/CRC ---------------------------- */
typedef struct {
    uint8_t  cw_ok[4];        /* per codeword parity success (CodewordStatus::decoded)      */
    int32_t  cw_iters[4];     /* per codeword LDPCDecoder::lastIterations()                 */
    uint8_t  all_ok;          /* CodewordStatus::allSuccess()                               */
    uint8_t  header_valid;    /* v2::parseHeader(...).valid on the reassembled bytes        */
    uint8_t  frame_crc_ok;    /* data frame: frame CRC over header+payload (deserialize)    */
    uint8_t  type;            /* HeaderInfo::type                                           */
    uint16_t seq;
    uint16_t payload_len;
    uint32_t src_hash;
    uint32_t dst_hash;
    uint8_t  total_cw;
    uint8_t  ladder_cw_mask;     /* bit c: codeword c failed its first decode and the retry ladder recovered it */
    uint8_t  ladder_max_attempt; /* highest ladder attempt (1..38, frame_v2.cpp:1409-1542) a recovered codeword needed */
    uint8_t  fp_repair;          /* false-positive repair (frame_v2.cpp:1558-1916): 0 not needed / not run,
                                    1 frame repaired, 2 given up (all four codewords marked failed) */
} ria_frame_status;

/* Batched first pass of v2::decodeFixedFrame (src/protocol/frame_v2.cpp:1335-1385, 1548-1556):
 * FrameInterleaver::deinterleave (src/fec/frame_interleaver.cpp:96-124), optional
 * ChannelInterleaver::deinterleave (src/fec/ldpc_decoder.cpp:552-625) with
 * bits_per_symbol, 4 x LDPCDecoder::decodeSoft(factor 0.9375, getRecommendedIterations(rate)),
 * take bytes_per_cw bytes per codeword, then v2::parseHeader + frame CRC
 * (src/protocol/frame_v2.cpp:115-128, 1195-1253, 555-600).  The retry ladder (:1389-1546) runs when
 * the context has RIA_DECODE_RETRY_LADDER set, the false-positive repair (:1558-1916) with
 * RIA_DECODE_FP_REPAIR (ria_ctx_set_decode_flags; RIA_DECODE_FULL = both = the whole function).
 *   soft_dev   [n][soft_stride] >= 2592 soft bits per frame (soft_stride >= 2592)
 *   data_dev   [n][4*bytes_per_cw] reassembled info bytes (codewords that failed are zeros)
 *   status_dev [n] */
int ria_frame_decode_batch_dev(ria_ctx* ctx, int rate, int use_channel_interleave,
                               int bits_per_symbol, const float* soft_dev, int32_t soft_stride,
                               int64_t n_frames, uint8_t* data_dev, ria_frame_status* status_dev);

/* Same with HOST buffers (soft [n][soft_stride], data [n][4*bytes_per_cw], status [n]): H2D, kernels and
 * D2H inside the call.  With RIA_DECODE_FULL this is v2::decodeFixedFrame(interleaved_soft, rate,
 * use_channel_deinterleave, bits_per_symbol) for n frames. */
int ria_frame_decode_batch_host(ria_ctx* ctx, int rate, int use_channel_interleave,
                                int bits_per_symbol, const float* soft, int32_t soft_stride,
                                int64_t n_frames, uint8_t* data, ria_frame_status* status);

/* Whole receive chain for OFDM data frames: presynced demod -> frame decode, one call.
 * Device buffers; llr scratch is owned by the context. */
int ria_ofdm_rx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg, int rate,
                           int use_channel_interleave,
                           const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                           const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                           uint8_t* data_dev, ria_frame_status* status_dev, float* snr_db_dev);

/* Same with HOST buffers: chunked H2D -> demod -> decode -> D2H inside the call. */
int ria_ofdm_rx_frames_host(ria_ctx* ctx, const ria_modem_config* cfg, int rate,
                            int use_channel_interleave,
                            const float* samples, int64_t frame_stride, int32_t frame_len,
                            const float* cfo_hz, const float* phase, int64_t n_frames,
                            uint8_t* data, ria_frame_status* status, float* snr_db);

/* ---- transmit synthesis on the device ---------------------------------- */
/* Batched v2::encodeFixedFrame (src/protocol/frame_v2.cpp:1285-1328): frames_dev [n][frame_stride] bytes
 * (frame_len valid, zero padded / truncated to 4 x bytes_per_cw) -> coded_dev [n][324] bytes: LDPC encode
 * (src/fec/ldpc_encoder.cpp:193-257), ChannelInterleaver::interleave when use_channel_interleave
 * (src/fec/ldpc_decoder.cpp:600-615, step from bits_per_symbol), FrameInterleaver::interleave. */
int ria_encode_fixed_frame_batch_dev(ria_ctx* ctx, int rate, int use_channel_interleave, int bits_per_symbol,
                                     const uint8_t* frames_dev, int64_t frame_stride, int32_t frame_len,
                                     int64_t n_frames, uint8_t* coded_dev);

/* Samples of one transmitted frame: (training_symbols + ceil(8 coded_len / bits per OFDM symbol)) symbols. */
int ria_ofdm_tx_frame_samples(const ria_modem_config* cfg, int32_t coded_len);

/* OFDM_COX: samples of one transmitted frame = one symbol of silence + 4 STS + the frame above. */
int ria_ofdm_cox_tx_frame_samples(const ria_modem_config* cfg, int32_t coded_len);

/* Batched OFDMModulator::generateTrainingSymbols(cfg->training_symbols) + modulate(coded, cfg->modulation)
 * (src/ofdm/modulator.cpp:528-582, 348-477): what OFDMChirpWaveform transmits after the chirp and what
 * IWaveform::process is handed on the receive side.  Sample-identical to the reference (same radix-2
 * inverse transform, mixer phasors and operation order).  coded_dev [n][coded_stride] bytes,
 * samples_dev [n][out_stride] fp32 with out_stride >= ria_ofdm_tx_frame_samples(). */
int ria_ofdm_tx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                           const uint8_t* coded_dev, int64_t coded_stride, int32_t coded_len,
                           int64_t n_frames, float* samples_dev, int64_t out_stride);

/* Batched OFDMModulator::generatePreamble() + modulate(coded, cfg->modulation) (src/ofdm/modulator.cpp:479-532,
 * 348-477): what OFDMNvisWaveform (OFDM_COX) transmits (src/waveform/ofdm_cox_waveform.cpp:106-119) -- guard,
 * 4 Schmidl-Cox STS (one symbol repeated), 2 LTS (one symbol repeated), data.  Sample-identical to the
 * reference.  out_stride >= ria_ofdm_cox_tx_frame_samples(); cfg->symbol_guard must be 0. */
int ria_ofdm_cox_tx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                               const uint8_t* coded_dev, int64_t coded_stride, int32_t coded_len,
                               int64_t n_frames, float* samples_dev, int64_t out_stride);

/* Batched fec::BurstInterleaver::deinterleave (src/fec/burst_interleaver.cpp:39-78), applied by
 * StreamingDecoder::finalizeBurstGroup (src/gui/modem/streaming_decoder.cpp:3209-3216) to the soft bits of
 * a burst group before decodeFrame.  physical_dev [n_groups][group_size][in_stride] -> logical_dev
 * [n_groups][group_size][out_stride], 2592 soft bits per frame; group_size < 2 copies (:43).  Soft bits past
 * 2592 in a row are not touched.  Strides multiples of 4, buffers 16-byte aligned, not in place. */
int ria_burst_deinterleave_batch_dev(ria_ctx* ctx, const float* physical_dev, int32_t in_stride,
                                     int32_t group_size, int64_t n_groups,
                                     float* logical_dev, int32_t out_stride);

/* ---- synchronisation ------------------------------------------------------------------------- */
/* sync::ZCConfig (src/sync/zc_sync.hpp:61-108) */
typedef struct {
    float   sample_rate;
    int32_t sequence_length;     /* 127 */
    int32_t upsample_factor;     /* 8   */
    int32_t num_repetitions;     /* 2   */
    float   carrier_freq;        /* 1500 */
    float   gap_ms;              /* 10  */
    int32_t root_ping, root_pong, root_data, root_control;   /* 1, 3, 5, 7 */
} ria_zc_config;

/* sync::ZCSyncResult (zc_sync.hpp:111-119) / sync::ChirpSync::DualChirpResult (chirp_sync.hpp:343-350) */
typedef struct {
    int32_t detected;
    int32_t start_sample;        /* ZC: position + preamble length; chirp: CFO-corrected up-chirp start */
    float   correlation;         /* ZC: best (combined) correlation; chirp: up-chirp correlation        */
    float   cfo_hz;
    float   snr_estimate;        /* ZC only (correlationToSNR); chirp: down-chirp correlation           */
    int32_t root;                /* ZC: detected root, -1 if none; chirp: raw up-chirp position         */
    int32_t frame_type;          /* ZC: ZCFrameType, 255 = unknown; chirp: raw down-chirp position      */
    int32_t aux;                 /* chirp: CFO-corrected down-chirp start                               */
} ria_sync_result;

/* the configuration MCDPSKWaveform::initZCSync uses (src/waveform/mc_dpsk_waveform.cpp:50-64) */
int ria_zc_config_default(ria_zc_config* cfg);

/* Batched replacement for sync::ZCSync::detect(samples, threshold, false, root_mask, known_cfo_hz)
 * (src/sync/zc_sync.hpp:192-391) as called by MCDPSKWaveform::detectDataSync
 * (src/waveform/mc_dpsk_waveform.cpp:227-292).
 *   samples_dev    fp32 search windows, window f at samples_dev + f*frame_stride, `window` samples
 *   known_cfo_dev  [n] known CFO in Hz added to the down-conversion frequency (NULL = 0)
 *   root_mask      bit 0 PING, 1 PONG, 2 DATA, 3 CONTROL (ZC_ROOT_MASK_*, zc_sync.hpp:34-41)
 *   out_dev        [n] results; at most 65535 windows per call */
int ria_zc_detect_batch_dev(ria_ctx* ctx, const ria_zc_config* cfg,
                            const float* samples_dev, int64_t frame_stride, int32_t window,
                            const float* known_cfo_dev, float threshold, uint32_t root_mask,
                            int64_t n_frames, ria_sync_result* out_dev);

/* sync::ChirpConfig (src/sync/chirp_sync.hpp:30-39); dual chirp always on */
typedef struct {
    float sample_rate;   /* 48000 */
    float f_start;       /* 300   */
    float f_end;         /* 2700  */
    float duration_ms;   /* 500   */
    float gap_ms;        /* 100   */
} ria_chirp_config;

int ria_chirp_config_default(ria_chirp_config* cfg);

/* Batched replacement for sync::ChirpSync::detectDualChirp(samples, threshold)
 * (src/sync/chirp_sync.hpp:352-512) -- what IWaveform::detectSync runs for both waveforms.
 * Result fields: detected = success, start_sample = up_chirp_start (CFO-corrected),
 * aux = down_chirp_start (CFO-corrected), correlation = up_correlation,
 * snr_estimate = down_correlation, cfo_hz, root / frame_type = raw up / down peak positions.
 * window <= 131072 samples, at most 65535 windows per call. */
int ria_chirp_detect_dual_batch_dev(ria_ctx* ctx, const ria_chirp_config* cfg,
                                    const float* samples_dev, int64_t frame_stride, int32_t window,
                                    float threshold, int64_t n_frames, ria_sync_result* out_dev);

/* Batched replacement for OFDMChirpWaveform::detectDataSync(samples, result, known_cfo_hz, threshold)
 * (src/waveform/ofdm_chirp_waveform.cpp:207-384): light (training-only) preamble detection for
 * connected-mode frames.  Result fields: detected, start_sample = first LTS sample, correlation,
 * cfo_hz = known CFO, aux = 1 when the burst-interleave marker (negated first LTS) is seen. */
int ria_ofdm_data_sync_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                 const float* samples_dev, int64_t frame_stride, int32_t window,
                                 const float* known_cfo_dev, float threshold, int64_t n_frames,
                                 ria_sync_result* out_dev);

/* Batched replacement for OFDMDemodulator::searchForSync(samples, out_position, out_cfo_hz, threshold)
 * (src/ofdm/demodulator.cpp:1450-1542) -- what OFDMNvisWaveform::detectSync (OFDM_COX) runs
 * (src/waveform/ofdm_cox_waveform.cpp:121-153): energy-gated Schmidl-Cox search in steps of 64 samples, plateau
 * check, LTS fine timing, coarse CFO.  Result fields: detected, start_sample = first LTS sample (what
 * processPresynced is then handed), cfo_hz, correlation = 0.9 (the waveform's constant), aux = the Schmidl-Cox
 * peak position.  noise_floor_dev (nullable) is OFDMDemodulator::Impl::noise_floor_energy per window before /
 * after the call, the only state the reference search carries between calls (0 = fresh demodulator).
 * window <= 65536 samples. */
int ria_ofdm_cox_search_sync_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                       const float* samples_dev, int64_t window_stride, int32_t window,
                                       float threshold, float* noise_floor_dev, int64_t n_windows,
                                       ria_sync_result* out_dev);

/* Whole receive chain for OFDM_COX frames, one call: OFDMNvisWaveform::detectSync (the batched searchForSync of
 * ria_ofdm_cox_search_sync_batch_dev) -> OFDMNvisWaveform::process at the LTS position found, with the CFO found and the
 * initial mixer phase -2 pi cfo pos / fs (src/waveform/ofdm_cox_waveform.cpp:121-218) -> the frame decode of
 * ria_ofdm_rx_frames_dev.  samples [n][window_stride] windows of `window` samples, frame_len = samples handed to
 * process() from the LTS on (2 training symbols + data).  A window in which nothing is found (or whose frame would
 * run past the window) yields a status with nothing valid.  noise_floor (nullable, in/out [n]) as in the search;
 * sync (nullable, [n]) receives the search results. */
int ria_ofdm_cox_rx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg, int rate, int use_channel_interleave,
                               const float* samples_dev, int64_t window_stride, int32_t window, int32_t frame_len,
                               float threshold, float* noise_floor_dev, int64_t n_windows,
                               uint8_t* data_dev, ria_frame_status* status_dev, float* snr_db_dev,
                               ria_sync_result* sync_dev);
int ria_ofdm_cox_rx_frames_host(ria_ctx* ctx, const ria_modem_config* cfg, int rate, int use_channel_interleave,
                                const float* samples, int64_t window_stride, int32_t window, int32_t frame_len,
                                float threshold, float* noise_floor, int64_t n_windows,
                                uint8_t* data, ria_frame_status* status, float* snr_db, ria_sync_result* sync);

/* Tap of the search above: OFDMDemodulator::Impl::measureCorrelation(offset) (src/ofdm/ofdm_sync.cpp:118-190), the
 * Schmidl-Cox metric of the FFT window that follows offset + cyclic prefix, one offset per window. */
int ria_ofdm_cox_correlation_batch_dev(ria_ctx* ctx, const ria_modem_config* cfg,
                                       const float* samples_dev, int64_t window_stride, int32_t window,
                                       const int32_t* offsets_dev, int64_t n_windows, float* corr_dev);

/* ---- MC-DPSK receive path -------------------------------------------------------------------- */
/* POD mirror of the RX-relevant fields of ultra::MultiCarrierDPSKConfig
 * (src/psk/multi_carrier_dpsk.hpp:27-100).  Reference defaults: 48000, 8 carriers (the tools use
 * 10), 500..2500 Hz, 512 samples per symbol, 8 training symbols. */
typedef struct {
    float    sample_rate;
    uint32_t num_carriers;        /* 1..16                                          */
    float    freq_low, freq_high;
    uint32_t samples_per_symbol;  /* 512                                            */
    uint32_t bits_per_symbol;     /* 1 = DBPSK, 2 = DQPSK                           */
    uint32_t spreading;           /* SpreadingMode as a factor: 1, 2 or 4           */
    uint32_t training_symbols;
} ria_mcdpsk_config;

/* soft bits one frame of frame_len samples yields (0 if it is shorter than training + ref) */
int ria_mcdpsk_soft_bits_per_frame(const ria_mcdpsk_config* cfg, int32_t frame_len);

/* Batched replacement for MCDPSKWaveform::process (src/waveform/mc_dpsk_waveform.cpp:294-338) =
 * MultiCarrierDPSKDemodulator::setChirpDetected(cfo) + process (src/psk/multi_carrier_dpsk.hpp:
 * 797-895): CFO correction by Hilbert transform + rotation when |cfo| > 0.1 Hz (:901-926),
 * setReference (:507-518), demodulateSoft with 2x/4x coherent despreading (:520-736).
 *   samples_dev  fp32, frame f at samples_dev + f*frame_stride, frame_len samples laid out
 *                [training_symbols x 512][reference 512][data ...] (what process() is handed)
 *   cfo_hz_dev   [n] CFO handed to setChirpDetected (NULL = 0: no correction pass at all)
 *   phase_dev    [n] cfo_initial_phase_ (setCFOWithPhase), NULL = 0
 *   llr_dev      [n][llr_stride] soft bits, symbol-major / carrier-minor (:683-698)
 *   fading_dev   [n] getFadingIndex() (frequency CV + temporal CV), may be NULL
 *   cfo_out_dev  [n] getEstimatedCFO() after the call, may be NULL */
int ria_mcdpsk_process_batch_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg,
                                 const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                 const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                 float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                 float* fading_dev, float* cfo_out_dev);

/* Same, for frames that sit somewhere inside their row: frame f starts at
 * samples_dev + f*frame_stride + start_dev[f] (what the streaming decoder does with
 * SyncResult::start_sample, src/gui/modem/streaming_decoder.cpp:1700-1760).  start_dev == NULL
 * means 0.  A frame whose [start, start + frame_len) leaves the row reports 0 soft bits. */
int ria_mcdpsk_process_batch_at_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg,
                                    const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                    const int32_t* start_dev,
                                    const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                    float* llr_dev, int32_t llr_stride, int32_t* n_llr_dev,
                                    float* fading_dev, float* cfo_out_dev);

/* Whole receive chain for chirp-acquired MC-DPSK frames (one LDPC codeword per frame), one call:
 * MCDPSKWaveform::detectSync (dual chirp, src/waveform/mc_dpsk_waveform.cpp:176-224) on the first
 * sync_window samples of every row -> MCDPSKWaveform::process at the detected training start with
 * the detected CFO (:294-338) -> fec::ChaseCache::store arithmetic into acc_dev[f][648]
 * (first_reception != 0 overwrites, otherwise adds; src/fec/chase_cache.cpp:75-85) ->
 * LDPCDecoder::decodeSoft on the combined soft bits.
 *   samples_dev  fp32 rows [n][row_stride]; a row holds [noise][chirp preamble][training][ref][data]
 *   frame_len    samples handed to process() from the training start (training + reference + data)
 *   sync_dev     [n] what detectDualChirp returned (aux = down-chirp start)
 * An undetected frame contributes all-zero soft bits and fails to decode. */
int ria_mcdpsk_rx_frames_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const ria_chirp_config* chirp,
                             const float* samples_dev, int64_t row_stride, int32_t sync_window,
                             int32_t frame_len, float threshold, int64_t n_frames,
                             int rate, int max_iter, float min_sum_factor,
                             float* acc_dev, int first_reception,
                             uint8_t* info_dev, int32_t info_stride, uint8_t* ok_dev, int32_t* iters_dev,
                             ria_sync_result* sync_dev);

/* The same chain for connected-mode receptions, acquired on the Zadoff-Chu data preamble:
 * MCDPSKWaveform::detectDataSync (src/waveform/mc_dpsk_waveform.cpp:227-292: ZCSync::detect with root_mask, e.g.
 * DATA | CONTROL, in a sync_window of 31 120 samples, streaming_decoder.cpp:423-435) -> process at the reported training
 * start with known + residual CFO (:275-281; known_cfo_dev nullable) -> chase combining -> LDPC.  A retransmission of a
 * chirp-acquired frame (first_reception = 0, same acc_dev rows) is combined with it. */
int ria_mcdpsk_zc_rx_frames_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const ria_zc_config* zc,
                                const float* samples_dev, int64_t row_stride, int32_t sync_window,
                                int32_t frame_len, const float* known_cfo_dev, float threshold, uint32_t root_mask,
                                int64_t n_frames, int rate, int max_iter, float min_sum_factor,
                                float* acc_dev, int first_reception,
                                uint8_t* info_dev, int32_t info_stride, uint8_t* ok_dev, int32_t* iters_dev,
                                ria_sync_result* sync_dev);

/* Same with HOST buffers (single reception, no cache): chunked H2D -> chain -> D2H inside the call. */
int ria_mcdpsk_rx_frames_host(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const ria_chirp_config* chirp,
                              const float* samples, int64_t row_stride, int32_t sync_window,
                              int32_t frame_len, float threshold, int64_t n_frames,
                              int rate, int max_iter, float min_sum_factor,
                              uint8_t* info, int32_t info_stride, uint8_t* ok, int32_t* iters,
                              ria_sync_result* sync);

/* ---- HARQ chase combining ---------------------------------------------------------------------- */
/* Sync preambles of the transmitter (payload-independent; a batch needs each once).  Every sample is a
 * closed-form evaluation, bit-identical to the reference (glibc's sinf/cosf incl. the large-argument
 * reduction are restated in csrc/rn_math.h):
 *   ria_zc_preamble_*     sync::ZCSync::generatePreambleForRoot (src/sync/zc_sync.hpp:133-190)
 *   ria_chirp_generate_*  sync::ChirpSync::generate, dual chirp (src/sync/chirp_sync.hpp:61-108)
 * `_samples` return the preamble length; `_dev` run one thread per sample on the context stream and
 * write out_dev[cap]; `_host` evaluate the same sample functions on the host (no GPU needed).  `_dev` and
 * `_host` return the number of samples written; `_host` returns minus the required capacity when `out`
 * is NULL or too small. */
int ria_zc_preamble_samples(const ria_zc_config* cfg);
int ria_chirp_generate_samples(const ria_chirp_config* cfg);
int ria_zc_preamble_dev(ria_ctx* ctx, const ria_zc_config* cfg, int root, float* out_dev, int cap);
int ria_chirp_generate_dev(ria_ctx* ctx, const ria_chirp_config* cfg, float* out_dev, int cap);
int ria_zc_preamble_host(const ria_zc_config* cfg, int root, float* out, int cap);
int ria_chirp_generate_host(const ria_chirp_config* cfg, float* out, int cap);

/* ---- MC-DPSK transmit synthesis on the device ---------------------------- */
/* Samples of one MC-DPSK frame body: (training_symbols + 1 + data symbols x spreading) x samples_per_symbol. */
int ria_mcdpsk_tx_frame_samples(const ria_mcdpsk_config* cfg, int32_t data_len);

/* Batched MultiCarrierDPSKModulator::generateTrainingSequence + generateReferenceSymbol + modulate(data)
 * (src/psk/multi_carrier_dpsk.hpp:141-275): the frame body a transmitter sends after the sync preamble and
 * IWaveform::process is handed.  Sample-identical to the reference.  data_dev [n][data_stride] bytes (one
 * LDPC codeword = 81 bytes in the reference's use), samples_dev [n][out_stride] fp32. */
int ria_mcdpsk_tx_frames_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg,
                             const uint8_t* data_dev, int64_t data_stride, int32_t data_len,
                             int64_t n_frames, float* samples_dev, int64_t out_stride);

/* Arithmetic of fec::ChaseCache::store (src/fec/chase_cache.cpp:27-88) on cache slots resident in
 * HBM: item i with slot_dev[i] >= 0 either overwrites (first_dev[i] != 0: first reception) or
 * accumulates (`existing[j] += soft[j]`, :81) its 648 LLRs into acc_dev[slot][648].  A slot may
 * appear at most once per call, so the accumulation order is the reception order.  The cache
 * policy (keys, <= 4 combines, LRU, TTL) is host logic (ria_b200.fec.ChaseCache). */
int ria_chase_combine_batch_dev(ria_ctx* ctx, float* acc_dev, const int32_t* slot_dev,
                                const uint8_t* first_dev, const float* llr_dev, int64_t llr_stride,
                                int64_t n);

/* ---- waveform / rate selection (host logic) ---------------------------------------------------- */
#define RIA_WAVEFORM_OFDM_COX   0   /* protocol::WaveformMode, src/protocol/frame_v2.hpp */
#define RIA_WAVEFORM_OTFS_EQ    1
#define RIA_WAVEFORM_OTFS_RAW   2
#define RIA_WAVEFORM_MFSK       3
#define RIA_WAVEFORM_MC_DPSK    4
#define RIA_WAVEFORM_OFDM_CHIRP 5
typedef struct {
    int32_t waveform;                 /* RIA_WAVEFORM_*                      */
    int32_t modulation;               /* ria_modulation                      */
    int32_t rate;                     /* ria_code_rate                       */
    float   estimated_throughput_bps;
    int32_t num_carriers;
    int32_t spreading;                /* 1, 2 or 4                           */
} ria_waveform_recommendation;
/* protocol::recommendWaveformAndRate (src/protocol/waveform_selection.hpp:112-222) */
int ria_recommend_waveform(float snr_db, float fading_index, ria_waveform_recommendation* rec);
/* protocol::recommendDataMode (waveform_selection.hpp:250-314); estimated_throughput_bps unset */
int ria_recommend_data_mode(float snr_db, int waveform, float fading_index, ria_waveform_recommendation* rec);

/* ---- channel simulation on the device ------------------------------------------------------ */
/* AWGN as SimulatedChannel::applyChannel (tools/cli_simulator.cpp:343-366): frame f of the batch is
 * tx_pool[(first_frame_id + f) % pool_frames] plus white Gaussian noise whose standard deviation
 * is sqrt(mean(s^2) / 10^(snr/10)) measured on that frame.  snr_db_dev ([n], may be NULL) overrides
 * the scalar snr_db per frame.  Noise comes from a Philox4x32-10 counter keyed by
 * (seed, first_frame_id + f, sample), so results are reproducible and independent of the sharding;
 * parity with the reference's std::mt19937 stream is statistical only. */
int ria_channel_awgn_batch_dev(ria_ctx* ctx, const float* tx_pool_dev, int32_t pool_frames,
                               int32_t frame_len, const float* snr_db_dev, float snr_db,
                               uint64_t seed, int64_t first_frame_id, int64_t n_frames,
                               float* out_dev, int64_t out_stride);

/* sim::WattersonChannel::Config (src/sim/hf_channel.hpp:33-47), fields that shape the output */
typedef struct {
    float    snr_db;
    float    delay_spread_ms;
    float    doppler_spread_hz;
    float    path1_gain, path2_gain;     /* 0.707 each */
    uint32_t sample_rate;
    uint32_t fading_enabled, multipath_enabled, noise_enabled;
    uint32_t stationary_start;           /* 1: draw the initial tap state from the stationary
                                            distribution (frames are independent draws of a
                                            long-running channel); 0: start at (1,0) like a freshly
                                            constructed WattersonChannel (:68-69)                 */
} ria_watterson_config;
/* itu_r_f1487 presets (hf_channel.hpp:411-488): 0 AWGN, 1 Good, 2 Moderate, 3 Poor, 4 Flutter */
int ria_watterson_preset(int condition, float snr_db, ria_watterson_config* cfg);
/* WattersonChannel::process for a batch (hf_channel.hpp:107-177, 267-284): frame f of the batch is
 * tx_pool[(first_frame_id + f) % pool_frames] through two Rayleigh taps (first-order IIR-shaped
 * complex Gaussians), a delay of delay_spread_ms on the second path and AWGN scaled to the rms of
 * the non-silent input samples.  The CFO stage of the class (applyCFO) is not part of this call.
 * Philox-keyed like ria_channel_awgn_batch_dev; parity with the reference is statistical. */
int ria_channel_watterson_batch_dev(ria_ctx* ctx, const ria_watterson_config* cfg,
                                    const float* tx_pool_dev, int32_t pool_frames, int32_t frame_len,
                                    const float* snr_db_dev, uint64_t seed, int64_t first_frame_id,
                                    int64_t n_frames, float* out_dev, int64_t out_stride);

/* The PING energy test of StreamingDecoder::decodeCurrentFrame (src/gui/modem/streaming_decoder.cpp:1127-1160,
 * 1219-1229) for a batch of receptions that start at the sync position: RMS of the first training_skip samples
 * (4608 for MC-DPSK) against the RMS of the <= 5000 samples behind them, sums in sample order.
 * out_dev [n][4] fp32 = {training_rms, data_rms, ratio (0 when training_rms <= 0.001), is_ping (ratio < 0.6)}. */
int ria_ping_energy_batch_dev(ria_ctx* ctx, const float* frames_dev, int64_t frame_stride, int32_t frame_len,
                              int32_t training_skip, int64_t n_frames, float* out_dev);

/* ---- HOST-buffer variants of the synchronisers and the MC-DPSK demodulator --------------------- */
/* Same arguments as the `_dev` entry points with host pointers (H2D -> `_dev` -> D2H inside the call, which
 * returns when the results are in the caller's buffers).  These are what the batch = 1 IWaveform adapters of
 * include/ria_b200_adapters.hpp call behind detectSync / detectDataSync / process:
 *   ria_chirp_detect_dual_batch_host  IWaveform::detectSync          (ofdm_chirp_waveform.cpp:163-205, mc_dpsk_waveform.cpp:177-224)
 *   ria_zc_detect_batch_host          MCDPSKWaveform::detectDataSync (mc_dpsk_waveform.cpp:227-292)
 *   ria_ofdm_data_sync_batch_host     OFDMChirpWaveform::detectDataSync (ofdm_chirp_waveform.cpp:207-384)
 *   ria_mcdpsk_process_batch_host     MCDPSKWaveform::process        (mc_dpsk_waveform.cpp:294-338)
 *   ria_ofdm_cox_search_sync_batch_host  OFDMNvisWaveform::detectSync (ofdm_cox_waveform.cpp:121-153); noise_floor is
 *                                     a host array here too */
int ria_chirp_detect_dual_batch_host(ria_ctx* ctx, const ria_chirp_config* cfg, const float* samples,
                                     int64_t frame_stride, int32_t window, float threshold,
                                     int64_t n_frames, ria_sync_result* out);
int ria_zc_detect_batch_host(ria_ctx* ctx, const ria_zc_config* cfg, const float* samples, int64_t frame_stride,
                             int32_t window, const float* known_cfo, float threshold, uint32_t root_mask,
                             int64_t n_frames, ria_sync_result* out);
int ria_ofdm_data_sync_batch_host(ria_ctx* ctx, const ria_modem_config* cfg, const float* samples,
                                  int64_t frame_stride, int32_t window, const float* known_cfo, float threshold,
                                  int64_t n_frames, ria_sync_result* out);
int ria_ofdm_cox_search_sync_batch_host(ria_ctx* ctx, const ria_modem_config* cfg, const float* samples,
                                        int64_t window_stride, int32_t window, float threshold, float* noise_floor,
                                        int64_t n_windows, ria_sync_result* out);
int ria_mcdpsk_process_batch_host(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const float* samples,
                                  int64_t frame_stride, int32_t frame_len, const float* cfo_hz, const float* phase,
                                  int64_t n_frames, float* llr, int32_t llr_stride, int32_t* n_llr,
                                  float* fading, float* cfo_out);

/* ---- error counters and their reduction over the GPUs of the box ------------ */
/* counters_dev[0..7] += {frames, frames_ok (4/4 codewords + header + frame CRC), codewords, codewords failed,
 * 0, 0, frames without a valid header, frames whose codewords all decoded but whose CRC failed} of a
 * ria_frame_status array, produced by a kernel on the context stream (what cli_simulator's per-station
 * statistics count, tools/cli_simulator.cpp:2226-2290).  The caller zeroes the counters. */
int ria_frame_counters_dev(ria_ctx* ctx, const ria_frame_status* status_dev, int64_t n_frames, int64_t* counters_dev);
/* The path's only collective: ncclAllReduce(sum, int64) of a counter vector, in place, on the context
 * stream.  `nccl_comm` is an ncclComm_t the host created (ncclCommInitRank), or one made with the helpers
 * below: rank 0 calls ria_nccl_get_unique_id, hands the 128 bytes to the other ranks by any means, every
 * rank calls ria_nccl_comm_create.  NCCL is resolved from the process image at run time (libria_b200.so does
 * not link it); RIA_E_UNSUPPORTED when it cannot be found. */
typedef struct { char internal[128]; } ria_nccl_unique_id;      /* layout of ncclUniqueId */
int ria_nccl_get_unique_id(ria_nccl_unique_id* id);
int ria_nccl_comm_create(ria_ctx* ctx, const ria_nccl_unique_id* id, int rank, int world, void** nccl_comm);
int ria_nccl_comm_destroy(void* nccl_comm);
int ria_counters_allreduce(ria_ctx* ctx, void* nccl_comm, int64_t* counters_dev, int32_t n);

/* CRC-16/CCITT-FALSE as ControlFrame::calculateCRC (src/protocol/frame_v2.cpp:115-128); host. */
uint16_t ria_crc16(const uint8_t* data, size_t len);
/* ChannelInterleaver step: findCoprimeStep (src/fec/ldpc_decoder.cpp:552-577); host. */
int ria_channel_interleaver_step(int bits_per_symbol, int total_bits);

#ifdef __cplusplus
}
#endif
#endif /* RIA_B200_H */
