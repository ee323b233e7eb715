// Whole receive chain for chirp-acquired MC-DPSK frames (BASELINE configs[2]), one call:
//
//   IWaveform::detectSync   MCDPSKWaveform::detectSync = ChirpSync::detectDualChirp, training
//                           start = down_chirp_start + chirp + gap, CFO from the peak gap
//                           (src/waveform/mc_dpsk_waveform.cpp:176-224)
//   IWaveform::process      MCDPSKWaveform::process at that offset with that CFO (:294-338)
//   fec::ChaseCache::store  first reception copies, later receptions add (src/fec/chase_cache.cpp:75-85)
//   LDPCDecoder::decodeSoft one codeword per frame on the combined soft bits
//
// Every stage is the batched kernel of its own translation unit; this file only strings them
// together on the context stream and owns the buffers that live between the stages (sync
// results, per-frame start / CFO, soft bits).  The _host variant moves row chunks over PCIe with
// two staging buffers so that the upload of chunk c+1 overlaps the kernels of chunk c.

#include "ria_internal.h"

namespace ria {

namespace {

// start[f] = training start inside the row (-1 = not detected: the demodulator reports no soft
// bits), cfo[f] = chirp CFO estimate, slot[f] = frame's accumulator row, first[f] = overwrite flag
__global__ void chain_prepare_kernel(const ria_sync_result* __restrict__ sync, long long n, int after_down,
                                     int first_flag, int* __restrict__ start, float* __restrict__ cfo,
                                     int* __restrict__ slot, unsigned char* __restrict__ first) {
    const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (f >= n) return;
    const ria_sync_result r = sync[f];
    start[f] = r.detected ? r.aux + after_down : -1;
    cfo[f] = r.detected ? r.cfo_hz : 0.0f;
    slot[f] = static_cast<int>(f);       // an undetected reception contributes all-zero soft bits
    first[f] = static_cast<unsigned char>(first_flag);
}

// the same for a ZC-acquired reception: start = training start reported by the detector, CFO = known + residual when
// a CFO is known (mc_dpsk_waveform.cpp:275-281)
__global__ void chain_prepare_zc_kernel(const ria_sync_result* __restrict__ sync, const float* __restrict__ known_cfo,
                                        long long n, int first_flag, int* __restrict__ start, float* __restrict__ cfo,
                                        int* __restrict__ slot, unsigned char* __restrict__ first) {
    const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (f >= n) return;
    const ria_sync_result r = sync[f];
    const float k = known_cfo ? known_cfo[f] : 0.0f;
    start[f] = r.detected ? r.start_sample : -1;
    cfo[f] = r.detected ? ((fabsf(k) > 0.1f) ? __fadd_rn(k, r.cfo_hz) : r.cfo_hz) : 0.0f;
    slot[f] = static_cast<int>(f);
    first[f] = static_cast<unsigned char>(first_flag);
}

int ensure_chain_scratch(ria_ctx* ctx, size_t bytes) {
    if (bytes > ctx->chain_scratch_bytes) {
        if (ctx->chain_scratch) RIA_CUDA(ctx, cudaFree(ctx->chain_scratch));
        ctx->chain_scratch = nullptr; ctx->chain_scratch_bytes = 0;
        RIA_CUDA(ctx, cudaMalloc(&ctx->chain_scratch, bytes));
        ctx->chain_scratch_bytes = bytes;
    }
    return RIA_OK;
}

}  // namespace

}  // namespace ria

extern "C" int ria_mcdpsk_rx_frames_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const ria_chirp_config* chirp,
                                        const float* samples_dev, int64_t row_stride, int32_t sync_window,
                                        int32_t frame_len, float threshold, int64_t n_frames,
                                        int rate, int max_iter, float min_sum_factor,
                                        float* acc_dev, int first_reception,
                                        uint8_t* info_dev, int32_t info_stride, uint8_t* ok_dev, int32_t* iters_dev,
                                        ria_sync_result* sync_dev) {
    using namespace ria;
    if (!ctx || !cfg || !chirp) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len <= 0 || sync_window <= 0 || row_stride < sync_window)
        return set_error(ctx, RIA_E_INVAL, "mcdpsk rx: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !acc_dev || !info_dev || !ok_dev || !iters_dev || !sync_dev)
        return set_error(ctx, RIA_E_INVAL, "mcdpsk rx: null buffer");
    const int n_soft = ria_mcdpsk_soft_bits_per_frame(cfg, frame_len);
    if (n_soft < RIA_LDPC_N) return set_error(ctx, RIA_E_INVAL, "mcdpsk rx: frame_len %d holds %d soft bits, one codeword needs %d",
                                             frame_len, n_soft, RIA_LDPC_N);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int llr_stride = (n_soft + 3) & ~3;
    const size_t b_llr = static_cast<size_t>(n_frames) * llr_stride * sizeof(float);
    const size_t b_i = (static_cast<size_t>(n_frames) * 4 + 255) & ~size_t(255);
    int rc = ensure_chain_scratch(ctx, b_llr + 4 * b_i + 256);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->chain_scratch);
    float* d_llr = reinterpret_cast<float*>(base);
    int* d_start = reinterpret_cast<int*>(base + b_llr);
    float* d_cfo = reinterpret_cast<float*>(base + b_llr + b_i);
    int* d_slot = reinterpret_cast<int*>(base + b_llr + 2 * b_i);
    unsigned char* d_first = base + b_llr + 3 * b_i;

    // ---- acquisition: dual chirp, slices bounded by the 3 MiB of spectra each window needs ----
    const int64_t slice = 2048;      // one wave of the peak kernel (14 CTAs x 148 SMs); 7 GB of spectra
    for (int64_t off = 0; off < n_frames; off += slice) {
        const int64_t n = (n_frames - off < slice) ? (n_frames - off) : slice;
        rc = ria_chirp_detect_dual_batch_dev(ctx, chirp, samples_dev + off * row_stride, row_stride, sync_window,
                                             threshold, n, sync_dev + off);
        if (rc != RIA_OK) return rc;
    }
    const int chirp_samples = static_cast<int>(chirp->sample_rate * chirp->duration_ms / 1000.0f);
    const int gap_samples = static_cast<int>(chirp->sample_rate * chirp->gap_ms / 1000.0f);
    chain_prepare_kernel<<<static_cast<unsigned>((n_frames + 255) / 256), 256, 0, st>>>(
        sync_dev, n_frames, chirp_samples + gap_samples, first_reception ? 1 : 0, d_start, d_cfo, d_slot, d_first);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;

    // ---- demodulate at the detected offsets with the detected CFO ----
    rc = ria_mcdpsk_process_batch_at_dev(ctx, cfg, samples_dev, row_stride, frame_len, d_start, d_cfo, nullptr, n_frames,
                                         d_llr, llr_stride, iters_dev /* n_llr, overwritten by the decoder */, nullptr, nullptr);
    if (rc != RIA_OK) return rc;
    // ---- chase combining into the caller's accumulators, then LDPC on the sums ----
    rc = ria_chase_combine_batch_dev(ctx, acc_dev, d_slot, d_first, d_llr, llr_stride, n_frames);
    if (rc != RIA_OK) return rc;
    return ria_ldpc_decode_batch_dev(ctx, rate, max_iter, min_sum_factor, acc_dev, n_frames, info_dev, info_stride,
                                     ok_dev, iters_dev);
}

// Connected-mode receptions: the Zadoff-Chu data preamble instead of the dual chirp (SURVEY.md 8d configs[2] variant ii;
// MCDPSKWaveform::detectDataSync, src/waveform/mc_dpsk_waveform.cpp:227-292), then the same process -> chase -> LDPC.
extern "C" int ria_mcdpsk_zc_rx_frames_dev(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const ria_zc_config* zc,
                                           const float* samples_dev, int64_t row_stride, int32_t sync_window,
                                           int32_t frame_len, const float* known_cfo_dev, float threshold, uint32_t root_mask,
                                           int64_t n_frames, int rate, int max_iter, float min_sum_factor,
                                           float* acc_dev, int first_reception,
                                           uint8_t* info_dev, int32_t info_stride, uint8_t* ok_dev, int32_t* iters_dev,
                                           ria_sync_result* sync_dev) {
    using namespace ria;
    if (!ctx || !cfg || !zc) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len <= 0 || sync_window <= 0 || row_stride < sync_window)
        return set_error(ctx, RIA_E_INVAL, "mcdpsk zc rx: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !acc_dev || !info_dev || !ok_dev || !iters_dev || !sync_dev)
        return set_error(ctx, RIA_E_INVAL, "mcdpsk zc rx: null buffer");
    const int n_soft = ria_mcdpsk_soft_bits_per_frame(cfg, frame_len);
    if (n_soft < RIA_LDPC_N) return set_error(ctx, RIA_E_INVAL, "mcdpsk zc rx: frame_len %d holds %d soft bits, one codeword needs %d",
                                             frame_len, n_soft, RIA_LDPC_N);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int llr_stride = (n_soft + 3) & ~3;
    const size_t b_llr = static_cast<size_t>(n_frames) * llr_stride * sizeof(float);
    const size_t b_i = (static_cast<size_t>(n_frames) * 4 + 255) & ~size_t(255);
    int rc = ensure_chain_scratch(ctx, b_llr + 4 * b_i + 256);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->chain_scratch);
    float* d_llr = reinterpret_cast<float*>(base);
    int* d_start = reinterpret_cast<int*>(base + b_llr);
    float* d_cfo = reinterpret_cast<float*>(base + b_llr + b_i);
    int* d_slot = reinterpret_cast<int*>(base + b_llr + 2 * b_i);
    unsigned char* d_first = base + b_llr + 3 * b_i;
    rc = ria_zc_detect_batch_dev(ctx, zc, samples_dev, row_stride, sync_window, known_cfo_dev, threshold, root_mask, n_frames, sync_dev);
    if (rc != RIA_OK) return rc;
    chain_prepare_zc_kernel<<<static_cast<unsigned>((n_frames + 255) / 256), 256, 0, st>>>(
        sync_dev, known_cfo_dev, n_frames, first_reception ? 1 : 0, d_start, d_cfo, d_slot, d_first);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    rc = ria_mcdpsk_process_batch_at_dev(ctx, cfg, samples_dev, row_stride, frame_len, d_start, d_cfo, nullptr, n_frames,
                                         d_llr, llr_stride, iters_dev, nullptr, nullptr);
    if (rc != RIA_OK) return rc;
    rc = ria_chase_combine_batch_dev(ctx, acc_dev, d_slot, d_first, d_llr, llr_stride, n_frames);
    if (rc != RIA_OK) return rc;
    return ria_ldpc_decode_batch_dev(ctx, rate, max_iter, min_sum_factor, acc_dev, n_frames, info_dev, info_stride,
                                     ok_dev, iters_dev);
}

extern "C" int ria_mcdpsk_rx_frames_host(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const ria_chirp_config* chirp,
                                         const float* samples, int64_t row_stride, int32_t sync_window,
                                         int32_t frame_len, float threshold, int64_t n_frames,
                                         int rate, int max_iter, float min_sum_factor,
                                         uint8_t* info, int32_t info_stride, uint8_t* ok, int32_t* iters,
                                         ria_sync_result* sync) {
    using namespace ria;
    if (!ctx || !cfg || !chirp) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len <= 0 || sync_window <= 0 || row_stride < sync_window || info_stride <= 0)
        return set_error(ctx, RIA_E_INVAL, "mcdpsk rx: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !info || !ok || !iters || !sync) return set_error(ctx, RIA_E_INVAL, "mcdpsk rx: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = n_frames < 512 ? n_frames : 512;                 // <= ~400 MB of samples per staging buffer
    const size_t in_b = static_cast<size_t>(chunk) * row_stride * sizeof(float);
    const size_t acc_b = static_cast<size_t>(chunk) * RIA_LDPC_N * sizeof(float);
    const size_t out_b = static_cast<size_t>(chunk) * (info_stride + 1 + 4 + sizeof(ria_sync_result)) + 1024;
    for (int b = 0; b < 2; ++b) {
        int rc = ensure_stage(ctx, b, in_b + acc_b + out_b + 1024, 0);
        if (rc != RIA_OK) return rc;
    }
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    int buf = 0;
    for (int64_t off = 0; off < n_frames; off += chunk, buf ^= 1) {
        const int64_t n = (n_frames - off < chunk) ? (n_frames - off) : chunk;
        unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[buf]);
        float* d_samp = reinterpret_cast<float*>(base);
        float* d_acc = reinterpret_cast<float*>(base + in_b);
        ria_sync_result* d_sync = reinterpret_cast<ria_sync_result*>(base + in_b + acc_b);
        int32_t* d_it = reinterpret_cast<int32_t*>(d_sync + chunk);
        uint8_t* d_ok = reinterpret_cast<uint8_t*>(d_it + chunk);
        uint8_t* d_info = d_ok + ((chunk + 15) & ~int64_t(15));
        RIA_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->stage_ev[buf], 0));     // buffer free again?
        RIA_CUDA(ctx, cudaMemcpyAsync(d_samp, samples + off * row_stride, static_cast<size_t>(n) * row_stride * sizeof(float),
                                      cudaMemcpyHostToDevice, cs));
        RIA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[2 + buf], cs));
        RIA_CUDA(ctx, cudaStreamWaitEvent(s, ctx->stage_ev[2 + buf], 0));
        int rc = ria_mcdpsk_rx_frames_dev(ctx, cfg, chirp, d_samp, row_stride, sync_window, frame_len, threshold, n,
                                          rate, max_iter, min_sum_factor, d_acc, 1, d_info, info_stride, d_ok, d_it, d_sync);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(info + off * info_stride, d_info, static_cast<size_t>(n) * info_stride, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(ok + off, d_ok, static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(iters + off, d_it, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(sync + off, d_sync, static_cast<size_t>(n) * sizeof(ria_sync_result), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[buf], s));
    }
    RIA_CUDA(ctx, cudaStreamSynchronize(s));
    return RIA_OK;
}
