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
/* Bind the context to an existing cudaStream_t (e.g. torch's current stream); NULL = own stream. */
int  ria_ctx_set_stream(ria_ctx* ctx, void* cuda_stream);
int  ria_ctx_synchronize(ria_ctx* ctx);
const char* ria_last_error(const ria_ctx* ctx);
const char* ria_version(void);
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

#ifdef __cplusplus
}
#endif
#endif /* RIA_B200_H */
