/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the RIA receive hot path.
 *
 * Plain C restatement of the reference algorithms listed in SURVEY.md section 8(a), each function
 * citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load libria_oracle.so; the product
 * (ria_b200/) never does and has no CPU fallback.
 *
 * Pinning: the reference ships no golden vectors, so every function here
 * is pinned against outputs of the unmodified reference compiled into oracle/_ref/libria_ref.so
 * (tests/test_oracle_ldpc.py, tests/test_oracle_cox_cpu.py) and against fixtures generated from it and
 * committed under tests/golden/ (tests/golden/make_golden.py).
 */
#ifndef RIA_ORACLE_H
#define RIA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_LDPC_N 648
#define ORC_LDPC_MAX_EDGES 4096

/* CodeRate numeric values follow include/ultra/types.hpp:91-100. */
enum { ORC_R1_4 = 0, ORC_R1_3 = 1, ORC_R1_2 = 2, ORC_R2_3 = 3, ORC_R3_4 = 4, ORC_R5_6 = 5, ORC_R7_8 = 6 };

typedef struct {
    int rate;
    int k, m, n;
    int n_edges;
    int row_ptr[ORC_LDPC_N + 1];          /* check i owns edges row_ptr[i]..row_ptr[i+1]-1      */
    int edge_var[ORC_LDPC_MAX_EDGES];     /* variable index of each edge, H_rows order          */
} orc_ldpc_code;

/* src/fec/ldpc_decoder.cpp:21-36, 65-138 */
void orc_ldpc_build(int rate, orc_ldpc_code* code);

/* src/fec/ldpc_encoder.cpp:193-257 : data -> coded bytes; returns number of coded bytes */
int orc_ldpc_encode(const orc_ldpc_code* code, const uint8_t* data, int len, uint8_t* out, int out_cap);

/* src/fec/ldpc_decoder.cpp:154-260 (decodeBP) on one codeword of n_llr <= 648 LLRs.
 * out gets ceil(k/8) bytes; returns 1 on parity success; *iters = last_iters. */
int orc_ldpc_decode(const orc_ldpc_code* code, const float* llr, int n_llr, int max_iter,
                    float factor, uint8_t* out, int* iters);

/* batch of independent codewords [n_cw][648] -> out [n_cw][out_stride] */
void orc_ldpc_decode_batch(const orc_ldpc_code* code, const float* llr, int n_cw, int max_iter,
                           float factor, uint8_t* out, int out_stride, uint8_t* ok, int32_t* iters);

/* std::mt19937 (32-bit Mersenne Twister as specified by the C++ standard) */
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;
void orc_mt_seed(orc_mt19937* g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937* g);

/* src/protocol/frame_v2.cpp:115-128 : CRC-16/CCITT-FALSE */
uint16_t orc_crc16(const uint8_t* data, int len);

/* Schmidl-Cox timing metric of the OFDM_COX acquisition for the FFT window that follows offset + cp_len:
 * DC removal, analytic signal through the reference's radix-2 FFT pair, P / R1 / R2 over the two halves
 * (src/ofdm/ofdm_sync.cpp:56-84, 118-163; src/dsp/fft.cpp:83-128).  Returns 1 (0 when the window does not
 * fit: the metric is then 0, :123-126; -1 on a bad size). */
int orc_cox_correlation(const float* samples, int n_samples, int offset, int cp_len, int fft_len,
                        float* metric, float* p_re, float* p_im, float* r1, float* r2);

/* Impl::estimateCoarseCFO (src/ofdm/ofdm_sync.cpp:230-261) at the Schmidl-Cox peak */
float orc_cox_coarse_cfo(const float* samples, int n_samples, int sync_offset, int cp_len, int fft_len,
                         unsigned sample_rate);

#ifdef __cplusplus
}
#endif
#endif
