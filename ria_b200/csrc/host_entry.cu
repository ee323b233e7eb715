// HOST-buffer entry points for the stages that only had device-pointer variants (the three synchronisers and
// the MC-DPSK demodulator), the device-side frame counters and the NCCL all-reduce of those counters.
//
// The `_host` variants are what the batch = 1 C++ adapters (include/ria_b200_adapters.hpp: ria::OFDMChirpWaveform /
// ria::MCDPSKWaveform, the IWaveform drop-ins) call: rows go up with one 2-D copy per chunk, the `_dev` entry
// point of the stage runs on the context stream, the results come back, and the call returns when they are in
// the caller's buffers.  They add no arithmetic of their own.
//
// ria_counters_allreduce is SURVEY.md 8(b)/(e)'s one collective: the int64 error counters of a batch, summed
// over the GPUs of the box with ncclAllReduce on the context stream.  libria_b200.so does not link NCCL: the
// four NCCL entry points it needs are resolved at run time from the process image (torch has loaded
// libnccl.so.2) or from the system library, so a host without NCCL can still load the receive chain.

#include <dlfcn.h>

#include <algorithm>

#include "ria_internal.h"

namespace ria {

namespace {

// rows of `width` floats, `n` of them, host stride -> packed device rows
int upload_rows(ria_ctx* ctx, float* dst, const float* src, int64_t stride, int32_t width, int64_t n) {
    RIA_CUDA(ctx, cudaMemcpy2DAsync(dst, static_cast<size_t>(width) * sizeof(float), src,
                                    static_cast<size_t>(stride) * sizeof(float), static_cast<size_t>(width) * sizeof(float),
                                    static_cast<size_t>(n), cudaMemcpyHostToDevice, ctx->stream));
    return RIA_OK;
}

size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

}  // namespace

}  // namespace ria

using namespace ria;

extern "C" int ria_chirp_detect_dual_batch_host(ria_ctx* ctx, const ria_chirp_config* cfg, const float* samples,
                                                int64_t frame_stride, int32_t window, float threshold,
                                                int64_t n_frames, ria_sync_result* out) {
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || window < 0 || frame_stride < window) return set_error(ctx, RIA_E_INVAL, "chirp host: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !out) return set_error(ctx, RIA_E_INVAL, "chirp host: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = std::min<int64_t>(n_frames, 1024);
    const size_t b_in = align256(static_cast<size_t>(chunk) * window * sizeof(float));
    int rc = ensure_stage(ctx, 0, b_in + chunk * sizeof(ria_sync_result) + 256, 0);
    if (rc != RIA_OK) return rc;
    float* d_in = static_cast<float*>(ctx->stage_dev[0]);
    ria_sync_result* d_out = reinterpret_cast<ria_sync_result*>(static_cast<unsigned char*>(ctx->stage_dev[0]) + b_in);
    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = std::min(chunk, n_frames - off);
        if ((rc = upload_rows(ctx, d_in, samples + off * frame_stride, frame_stride, window, n)) != RIA_OK) return rc;
        if ((rc = ria_chirp_detect_dual_batch_dev(ctx, cfg, d_in, window, window, threshold, n, d_out)) != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, n * sizeof(ria_sync_result), cudaMemcpyDeviceToHost, ctx->stream));
        RIA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RIA_OK;
}

extern "C" int ria_zc_detect_batch_host(ria_ctx* ctx, const ria_zc_config* cfg, const float* samples, int64_t frame_stride,
                                        int32_t window, const float* known_cfo, float threshold, uint32_t root_mask,
                                        int64_t n_frames, ria_sync_result* out) {
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || window < 0 || frame_stride < window) return set_error(ctx, RIA_E_INVAL, "zc host: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !out) return set_error(ctx, RIA_E_INVAL, "zc host: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = std::min<int64_t>(n_frames, 8192);
    const size_t b_in = align256(static_cast<size_t>(chunk) * window * sizeof(float));
    const size_t b_cfo = align256(static_cast<size_t>(chunk) * sizeof(float));
    int rc = ensure_stage(ctx, 0, b_in + b_cfo + chunk * sizeof(ria_sync_result) + 256, 0);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_cfo = reinterpret_cast<float*>(base + b_in);
    ria_sync_result* d_out = reinterpret_cast<ria_sync_result*>(base + b_in + b_cfo);
    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = std::min(chunk, n_frames - off);
        if ((rc = upload_rows(ctx, d_in, samples + off * frame_stride, frame_stride, window, n)) != RIA_OK) return rc;
        if (known_cfo) RIA_CUDA(ctx, cudaMemcpyAsync(d_cfo, known_cfo + off, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = ria_zc_detect_batch_dev(ctx, cfg, d_in, window, window, known_cfo ? d_cfo : nullptr, threshold, root_mask, n, d_out)) != RIA_OK)
            return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, n * sizeof(ria_sync_result), cudaMemcpyDeviceToHost, ctx->stream));
        RIA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RIA_OK;
}

extern "C" int ria_ofdm_data_sync_batch_host(ria_ctx* ctx, const ria_modem_config* cfg, const float* samples,
                                             int64_t frame_stride, int32_t window, const float* known_cfo, float threshold,
                                             int64_t n_frames, ria_sync_result* out) {
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || window < 0 || frame_stride < window) return set_error(ctx, RIA_E_INVAL, "data sync host: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !out) return set_error(ctx, RIA_E_INVAL, "data sync host: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = std::min<int64_t>(n_frames, 8192);
    const size_t b_in = align256(static_cast<size_t>(chunk) * window * sizeof(float));
    const size_t b_cfo = align256(static_cast<size_t>(chunk) * sizeof(float));
    int rc = ensure_stage(ctx, 0, b_in + b_cfo + chunk * sizeof(ria_sync_result) + 256, 0);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_cfo = reinterpret_cast<float*>(base + b_in);
    ria_sync_result* d_out = reinterpret_cast<ria_sync_result*>(base + b_in + b_cfo);
    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = std::min(chunk, n_frames - off);
        if ((rc = upload_rows(ctx, d_in, samples + off * frame_stride, frame_stride, window, n)) != RIA_OK) return rc;
        if (known_cfo) RIA_CUDA(ctx, cudaMemcpyAsync(d_cfo, known_cfo + off, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = ria_ofdm_data_sync_batch_dev(ctx, cfg, d_in, window, window, known_cfo ? d_cfo : nullptr, threshold, n, d_out)) != RIA_OK)
            return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, n * sizeof(ria_sync_result), cudaMemcpyDeviceToHost, ctx->stream));
        RIA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RIA_OK;
}

extern "C" int ria_ofdm_cox_search_sync_batch_host(ria_ctx* ctx, const ria_modem_config* cfg, const float* samples,
                                                   int64_t window_stride, int32_t window, float threshold,
                                                   float* noise_floor, int64_t n_windows, ria_sync_result* out) {
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_windows < 0 || window < 0 || window_stride < window) return set_error(ctx, RIA_E_INVAL, "ofdm cox host: bad sizes");
    if (n_windows == 0) return RIA_OK;
    if (!samples || !out) return set_error(ctx, RIA_E_INVAL, "ofdm cox host: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = std::min<int64_t>(n_windows, 4096);
    const int32_t d_stride = (window + 3) & ~3;
    const size_t b_in = align256(static_cast<size_t>(chunk) * d_stride * sizeof(float));
    const size_t b_nf = align256(static_cast<size_t>(chunk) * sizeof(float));
    int rc = ensure_stage(ctx, 0, b_in + b_nf + chunk * sizeof(ria_sync_result) + 256, 0);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_nf = reinterpret_cast<float*>(base + b_in);
    ria_sync_result* d_out = reinterpret_cast<ria_sync_result*>(base + b_in + b_nf);
    for (int64_t off = 0; off < n_windows; off += chunk) {
        const int64_t n = std::min(chunk, n_windows - off);
        RIA_CUDA(ctx, cudaMemcpy2DAsync(d_in, d_stride * sizeof(float), samples + off * window_stride, window_stride * sizeof(float),
                                        window * sizeof(float), n, cudaMemcpyHostToDevice, ctx->stream));
        if (noise_floor) RIA_CUDA(ctx, cudaMemcpyAsync(d_nf, noise_floor + off, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = ria_ofdm_cox_search_sync_batch_dev(ctx, cfg, d_in, d_stride, window, threshold, noise_floor ? d_nf : nullptr, n, d_out)) != RIA_OK)
            return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, n * sizeof(ria_sync_result), cudaMemcpyDeviceToHost, ctx->stream));
        if (noise_floor) RIA_CUDA(ctx, cudaMemcpyAsync(noise_floor + off, d_nf, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        RIA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RIA_OK;
}

extern "C" int ria_mcdpsk_process_batch_host(ria_ctx* ctx, const ria_mcdpsk_config* cfg, const float* samples,
                                             int64_t frame_stride, int32_t frame_len, const float* cfo_hz, const float* phase,
                                             int64_t n_frames, float* llr, int32_t llr_stride, int32_t* n_llr,
                                             float* fading, float* cfo_out) {
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len < 0 || frame_stride < frame_len || llr_stride < 0)
        return set_error(ctx, RIA_E_INVAL, "mcdpsk host: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !llr || !n_llr) return set_error(ctx, RIA_E_INVAL, "mcdpsk host: null buffer");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = std::min<int64_t>(n_frames, 1024);
    const int32_t d_stride = (llr_stride + 3) & ~3;
    const size_t b_in = align256(static_cast<size_t>(chunk) * frame_len * sizeof(float));
    const size_t b_llr = align256(static_cast<size_t>(chunk) * d_stride * sizeof(float));
    const size_t b_f = align256(static_cast<size_t>(chunk) * sizeof(float));
    int rc = ensure_stage(ctx, 0, b_in + b_llr + 5 * b_f + 256, 0);
    if (rc != RIA_OK) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_llr = reinterpret_cast<float*>(base + b_in);
    float* d_cfo = reinterpret_cast<float*>(base + b_in + b_llr);
    float* d_ph = reinterpret_cast<float*>(base + b_in + b_llr + b_f);
    int32_t* d_n = reinterpret_cast<int32_t*>(base + b_in + b_llr + 2 * b_f);
    float* d_fad = reinterpret_cast<float*>(base + b_in + b_llr + 3 * b_f);
    float* d_co = reinterpret_cast<float*>(base + b_in + b_llr + 4 * b_f);
    cudaStream_t s = ctx->stream;
    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = std::min(chunk, n_frames - off);
        if ((rc = upload_rows(ctx, d_in, samples + off * frame_stride, frame_stride, frame_len, n)) != RIA_OK) return rc;
        if (cfo_hz) RIA_CUDA(ctx, cudaMemcpyAsync(d_cfo, cfo_hz + off, n * sizeof(float), cudaMemcpyHostToDevice, s));
        if (phase) RIA_CUDA(ctx, cudaMemcpyAsync(d_ph, phase + off, n * sizeof(float), cudaMemcpyHostToDevice, s));
        rc = ria_mcdpsk_process_batch_dev(ctx, cfg, d_in, frame_len, frame_len, cfo_hz ? d_cfo : nullptr, phase ? d_ph : nullptr, n,
                                          d_llr, d_stride, d_n, d_fad, d_co);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpy2DAsync(llr + off * llr_stride, static_cast<size_t>(llr_stride) * sizeof(float), d_llr,
                                        static_cast<size_t>(d_stride) * sizeof(float), static_cast<size_t>(llr_stride) * sizeof(float),
                                        static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(n_llr + off, d_n, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (fading) RIA_CUDA(ctx, cudaMemcpyAsync(fading + off, d_fad, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (cfo_out) RIA_CUDA(ctx, cudaMemcpyAsync(cfo_out + off, d_co, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return RIA_OK;
}

// ---------------------------------------------------------------------------------------------
// error counters of a batch, produced on the device
// ---------------------------------------------------------------------------------------------
namespace ria {
namespace {

// counters[0..7] += {frames, frames_ok, codewords, codewords failed, 0, 0, frames without a valid header,
// frames whose four codewords decoded but whose header / frame CRC failed}; warp-reduced, one atomic per warp
__global__ void frame_counters_kernel(const ria_frame_status* __restrict__ st, long long n, unsigned long long* __restrict__ out) {
    unsigned long long c[4] = {0, 0, 0, 0};      // frames_ok, cw_fail, no_header, crc_fail
    for (long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; f < n;
         f += static_cast<long long>(gridDim.x) * blockDim.x) {
        const ria_frame_status s = st[f];
        const bool ok = s.all_ok == 1 && s.header_valid == 1 && s.frame_crc_ok == 1;
        c[0] += ok ? 1 : 0;
        c[1] += (s.cw_ok[0] == 0) + (s.cw_ok[1] == 0) + (s.cw_ok[2] == 0) + (s.cw_ok[3] == 0);
        c[2] += s.header_valid == 1 ? 0 : 1;
        c[3] += (s.all_ok == 1 && !ok) ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        for (int d = 16; d > 0; d >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], d);
    if ((threadIdx.x & 31) == 0) {
        if (c[0]) atomicAdd(out + 1, c[0]);
        if (c[1]) atomicAdd(out + 3, c[1]);
        if (c[2]) atomicAdd(out + 6, c[2]);
        if (c[3]) atomicAdd(out + 7, c[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(out + 0, static_cast<unsigned long long>(n));
        atomicAdd(out + 2, static_cast<unsigned long long>(4 * n));
    }
}

// NCCL, resolved at run time
struct NcclApi {
    bool tried = false, ok = false;
    int (*get_unique_id)(void*) = nullptr;
    int (*comm_init_rank)(void**, int, ria_nccl_unique_id, int) = nullptr;
    int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*comm_destroy)(void*) = nullptr;
    const char* (*get_error_string)(int) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.tried) return api;
    api.tried = true;
    void* h = RTLD_DEFAULT;
    if (!dlsym(h, "ncclAllReduce")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return api;
    }
    api.get_unique_id = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<int (*)(void**, int, ria_nccl_unique_id, int)>(dlsym(h, "ncclCommInitRank"));
    api.all_reduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
    api.comm_destroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    api.get_error_string = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy;
    return api;
}

int nccl_fail(ria_ctx* ctx, const char* what, int rc) {
    NcclApi& a = nccl();
    return set_error(ctx, RIA_E_CUDA, "%s failed: %s", what, a.get_error_string ? a.get_error_string(rc) : "NCCL error");
}

}  // namespace
}  // namespace ria

extern "C" int ria_frame_counters_dev(ria_ctx* ctx, const ria_frame_status* status_dev, int64_t n_frames, int64_t* counters_dev) {
    if (!ctx) return RIA_E_INVAL;
    if (n_frames < 0 || !counters_dev || (n_frames > 0 && !status_dev)) return set_error(ctx, RIA_E_INVAL, "frame counters: bad arguments");
    if (n_frames == 0) return RIA_OK;
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = 256;
    const int blocks = static_cast<int>(std::min<int64_t>((n_frames + threads - 1) / threads, 4 * static_cast<int64_t>(ctx->sm_count)));
    frame_counters_kernel<<<blocks, threads, 0, ctx->stream>>>(status_dev, n_frames, reinterpret_cast<unsigned long long*>(counters_dev));
    ctx->launches += 1;
    RIA_CUDA(ctx, cudaGetLastError());
    return RIA_OK;
}

extern "C" int ria_nccl_get_unique_id(ria_nccl_unique_id* id) {
    if (!id) return RIA_E_INVAL;
    NcclApi& a = nccl();
    if (!a.ok) return RIA_E_UNSUPPORTED;
    return a.get_unique_id(id) == 0 ? RIA_OK : RIA_E_CUDA;
}

extern "C" int ria_nccl_comm_create(ria_ctx* ctx, const ria_nccl_unique_id* id, int rank, int world, void** comm) {
    if (!ctx || !id || !comm) return RIA_E_INVAL;
    NcclApi& a = nccl();
    if (!a.ok) return set_error(ctx, RIA_E_UNSUPPORTED, "NCCL is not loadable in this process");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = a.comm_init_rank(comm, world, *id, rank);
    return rc == 0 ? RIA_OK : nccl_fail(ctx, "ncclCommInitRank", rc);
}

extern "C" int ria_nccl_comm_destroy(void* comm) {
    NcclApi& a = nccl();
    if (!a.ok || !comm) return RIA_E_INVAL;
    return a.comm_destroy(comm) == 0 ? RIA_OK : RIA_E_CUDA;
}

extern "C" int ria_counters_allreduce(ria_ctx* ctx, void* nccl_comm, int64_t* counters_dev, int32_t n) {
    if (!ctx) return RIA_E_INVAL;
    if (!nccl_comm || !counters_dev || n <= 0) return set_error(ctx, RIA_E_INVAL, "counters allreduce: bad arguments");
    NcclApi& a = nccl();
    if (!a.ok) return set_error(ctx, RIA_E_UNSUPPORTED, "NCCL is not loadable in this process");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    // ncclInt64 = 4, ncclSum = 0 (nccl.h); in place, on the stream the counters were produced on
    const int rc = a.all_reduce(counters_dev, counters_dev, static_cast<size_t>(n), 4, 0, nccl_comm, ctx->stream);
    return rc == 0 ? RIA_OK : nccl_fail(ctx, "ncclAllReduce", rc);
}
