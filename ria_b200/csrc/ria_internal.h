// Internal declarations shared by the translation units of libria_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "ria_b200.h"

namespace ria {

// ---------------------------------------------------------------------------------------------
// LDPC code tables (host side: ldpc_code.cpp)
// ---------------------------------------------------------------------------------------------
struct LdpcCodeHost {
    int rate = -1;
    int k = 0, m = 0, n = 0, n_edges = 0;
    int dv_max = 0;                         // max degree over the k info variables
    std::vector<int32_t> row_ptr;           // m+1
    std::vector<int32_t> edge_var;          // n_edges, H_rows order (identity edge last in each row)
    // device-layout tables
    std::vector<uint16_t> chk_var;          // [m][8]: info-variable index per slot, 0xFFFF = unused;
                                            //         slot 7 holds the number of INFO edges of the check
    std::vector<uint16_t> var_slot;         // [dv_max][k]: c2v slot ids of variable j, ascending
                                            //         check index; 0xFFFF = unused
};

const LdpcCodeHost& ldpc_code_host(int rate);   // cached, thread-safe; throws on bad rate
bool ldpc_rate_valid(int rate);

struct OfdmTablesDev;   // ofdm_tables.h
struct McdpskTablesDev; // mcdpsk.cu
struct ZcTablesDev;     // zc_sync.cu
struct ChirpTablesDev;  // chirp_sync.cu
struct McdpskTxTablesDev;  // mcdpsk_tx.cu
struct CoxTablesDev;    // ofdm_cox.cu

struct LdpcCodeDev {
    bool ready = false;
    int k = 0, m = 0, dv_max = 0;
    uint16_t* chk_var = nullptr;
    uint16_t* var_slot = nullptr;
    int launch_warps = 0, launch_ctas_per_sm = 0;     // launch shape, settled on first use
    size_t launch_smem = 0;
};

}  // namespace ria

// ---------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------
struct ria_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_per_sm = 0;                 // shared memory per SM / opt-in maximum per CTA
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;      // stream all work is issued on
    cudaStream_t own_stream = nullptr;  // created by us (destroyed with the ctx)
    cudaStream_t copy_stream = nullptr; // H2D/D2H staging for *_host entry points
    std::string last_error;
    int64_t launches = 0;
    int decode_flags = 0;                   // RIA_DECODE_* (ria_ctx_set_decode_flags)
    ria::LdpcCodeDev ldpc[8];
    unsigned int* work_counter = nullptr;   // device, dynamic tile schedulers (one slot per kernel)
    std::vector<ria::OfdmTablesDev*> ofdm_tables;
    std::vector<ria::McdpskTablesDev*> mcdpsk_tables;
    std::vector<ria::ZcTablesDev*> zc_tables;
    std::vector<ria::ChirpTablesDev*> chirp_tables;
    std::vector<ria::McdpskTxTablesDev*> mcdpsk_tx_tables;
    std::vector<ria::CoxTablesDev*> cox_tables;
    float* hilbert65 = nullptr;             // 65-tap Hilbert FIR (OFDM data sync)
    // scratch owned by the context for the fused chain entry points
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    std::unordered_map<const void*, int> occ_cache;   // kernel -> resident CTAs per SM (attributes set once)
    void* chain_scratch = nullptr;          // buffers between the stages of the MC-DPSK chain
    size_t chain_scratch_bytes = 0;
    void* ofdm_scratch = nullptr;           // carrier bins / CFO phases between the OFDM stages
    size_t ofdm_scratch_bytes = 0;
    // optional per-kernel timing (CUDA events on the launching stream), see ria_ctx_set_timing
    bool timing = false;
    struct TimedLaunch { int kind; cudaEvent_t start, stop; };
    std::vector<TimedLaunch> timed;
    double timed_ms[32] = {};               // folded totals per kind (the event list is bounded, see time_begin)
    int64_t timed_n[32] = {};
    // staging for *_host entry points (grown on demand)
    void* stage_dev[2] = {nullptr, nullptr};
    size_t stage_dev_bytes[2] = {0, 0};
    void* stage_pin[2] = {nullptr, nullptr};
    size_t stage_pin_bytes[2] = {0, 0};
    cudaEvent_t stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace ria {

int set_error(ria_ctx* ctx, int code, const char* fmt, ...);
int ensure_stage(ria_ctx* ctx, int which, size_t dev_bytes, size_t pin_bytes);
int ldpc_tables_dev(ria_ctx* ctx, int rate, const LdpcCodeDev** out);
int ofdm_tables_dev(ria_ctx* ctx, const ria_modem_config& cfg, int need_nco, const OfdmTablesDev** out);
void ofdm_tables_free(OfdmTablesDev* t);
void mcdpsk_tables_free(McdpskTablesDev* t);
void zc_tables_free(ZcTablesDev* t);
void chirp_tables_free(ChirpTablesDev* t);
void mcdpsk_tx_tables_free(McdpskTxTablesDev* t);
void cox_tables_free(CoxTablesDev* t);
int ensure_scratch(ria_ctx* ctx, size_t bytes);
int recommended_ldpc_iterations(int rate);          // LDPCCodec::getRecommendedIterations (frame.cu)
// ldpc.cu / ldpc_retry.cu
int ldpc_launch(ria_ctx* ctx, int rate, int max_iter, float min_sum_factor, const float* llr_dev, int64_t n_cw,
                int frame_mode, int soft_stride, int step,
                uint8_t* info_dev, int info_stride, uint8_t* ok_dev, int32_t* iters_dev);
int ldpc_retry_launch(ria_ctx* ctx, int rate, int max_iter, const float* llr_dev, int64_t n_units, int frame_mode,
                      int soft_stride, int step, uint8_t* info_dev, int info_stride, uint8_t* ok_dev,
                      int32_t* iters_dev, uint8_t* attempt_dev, int* list_scratch);
int frame_repair_launch(ria_ctx* ctx, int rate, int max_iter, const float* soft_dev, int64_t n_frames, int soft_stride,
                        int step, uint8_t* info_dev, int info_stride, uint8_t* ok_dev, uint8_t* repair_dev,
                        int* list_scratch);

// kernel kinds for the timing hook / launch accounting
enum KernelKind { KK_LDPC = 0, KK_OFDM_DEMOD = 1, KK_FRAME_STATUS = 2, KK_AWGN = 3, KK_MCDPSK = 4,
                  KK_ZC_SYNC = 5, KK_CHIRP_SYNC = 6, KK_CHASE = 7, KK_WATTERSON = 8, KK_MCDPSK_CFO = 9, KK_OFDM_SYNC = 10,
                  KK_OFDM_FFT = 11, KK_OFDM_CARRIER = 12, KK_OFDM_PHASE = 13, KK_LDPC_RETRY = 14, KK_FRAME_REPAIR = 15, KK_COUNT = 16 };
void time_begin(ria_ctx* ctx, int kind);
void time_end(ria_ctx* ctx);

}  // namespace ria

#define RIA_CUDA(ctx, expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return ria::set_error((ctx), RIA_E_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)
