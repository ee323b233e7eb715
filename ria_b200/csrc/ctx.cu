// Context management for libria_b200.so: one ria_ctx per GPU, bound to a stream.

#include "ria_internal.h"

#include <cstdarg>
#include <new>

namespace ria {

int set_error(ria_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    else fprintf(stderr, "ria_b200: %s\n", buf);
    return code;
}

int ensure_stage(ria_ctx* ctx, int which, size_t dev_bytes, size_t pin_bytes) {
    if (dev_bytes > ctx->stage_dev_bytes[which]) {
        if (ctx->stage_dev[which]) RIA_CUDA(ctx, cudaFree(ctx->stage_dev[which]));
        ctx->stage_dev[which] = nullptr; ctx->stage_dev_bytes[which] = 0;
        RIA_CUDA(ctx, cudaMalloc(&ctx->stage_dev[which], dev_bytes));
        ctx->stage_dev_bytes[which] = dev_bytes;
    }
    if (pin_bytes > ctx->stage_pin_bytes[which]) {
        if (ctx->stage_pin[which]) RIA_CUDA(ctx, cudaFreeHost(ctx->stage_pin[which]));
        ctx->stage_pin[which] = nullptr; ctx->stage_pin_bytes[which] = 0;
        RIA_CUDA(ctx, cudaMallocHost(&ctx->stage_pin[which], pin_bytes));
        ctx->stage_pin_bytes[which] = pin_bytes;
    }
    return RIA_OK;
}

// Fold the recorded launches into the per-kind totals and release their events (waits for them).
static void fold_timed(ria_ctx* ctx) {
    for (auto& t : ctx->timed) {
        float e = 0.f;
        if (cudaEventSynchronize(t.stop) == cudaSuccess && cudaEventElapsedTime(&e, t.start, t.stop) == cudaSuccess &&
            t.kind >= 0 && t.kind < 32) {
            ctx->timed_ms[t.kind] += e;
            ctx->timed_n[t.kind] += 1;
        }
        cudaEventDestroy(t.start);
        cudaEventDestroy(t.stop);
    }
    ctx->timed.clear();
}

void time_begin(ria_ctx* ctx, int kind) {
    if (!ctx->timing) return;
    if (ctx->timed.size() >= 2048) fold_timed(ctx);      // bounded event list however long timing stays on
    ria_ctx::TimedLaunch t{kind, nullptr, nullptr};
    if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) return;
    cudaEventRecord(t.start, ctx->stream);
    ctx->timed.push_back(t);
}

void time_end(ria_ctx* ctx) {
    if (!ctx->timing || ctx->timed.empty()) return;
    cudaEventRecord(ctx->timed.back().stop, ctx->stream);
}

int ensure_scratch(ria_ctx* ctx, size_t bytes) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) RIA_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr; ctx->scratch_bytes = 0;
        RIA_CUDA(ctx, cudaMalloc(&ctx->scratch, bytes));
        ctx->scratch_bytes = bytes;
    }
    return RIA_OK;
}

}  // namespace ria

extern "C" const char* ria_version(void) { return "ria_b200 0.1 (sm_100a)"; }

extern "C" int ria_ctx_create(int device, ria_ctx** out) {
    if (!out) return RIA_E_INVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return ria::set_error(nullptr, RIA_E_CUDA, "no usable CUDA device (%s); there is no CPU fallback",
                              cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return ria::set_error(nullptr, RIA_E_INVAL, "device %d out of range (have %d)", device, count);
    ria_ctx* ctx = new (std::nothrow) ria_ctx();
    if (!ctx) return RIA_E_NOMEM;
    ctx->device = device;
    auto fail = [&](cudaError_t err, const char* what) {
        fprintf(stderr, "ria_b200: %s: %s\n", what, cudaGetErrorString(err));
        for (auto& ev : ctx->stage_ev) if (ev) cudaEventDestroy(ev);
        if (ctx->work_counter) cudaFree(ctx->work_counter);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
        delete ctx;
        return RIA_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) {
        fprintf(stderr, "ria_b200: device %d is sm_%d%d; this library is built for sm_100a only\n",
                device, prop.major, prop.minor);
        delete ctx;
        return RIA_E_UNSUPPORTED;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail(e, "cudaStreamCreate");
    ctx->stream = ctx->own_stream;
    if ((e = cudaMalloc(&ctx->work_counter, 64 * sizeof(unsigned int))) != cudaSuccess)
        return fail(e, "cudaMalloc");
    for (auto& ev : ctx->stage_ev)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess)
            return fail(e, "cudaEventCreate");
    *out = ctx;
    return RIA_OK;
}

extern "C" int ria_ctx_destroy(ria_ctx* ctx) {
    if (!ctx) return RIA_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& t : ctx->ldpc) {
        if (t.chk_var) cudaFree(t.chk_var);
        if (t.var_slot) cudaFree(t.var_slot);
    }
    for (auto& t : ctx->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
    for (auto* t : ctx->ofdm_tables) ria::ofdm_tables_free(t);
    for (auto* t : ctx->mcdpsk_tables) ria::mcdpsk_tables_free(t);
    for (auto* t : ctx->zc_tables) ria::zc_tables_free(t);
    for (auto* t : ctx->chirp_tables) ria::chirp_tables_free(t);
    for (auto* t : ctx->mcdpsk_tx_tables) ria::mcdpsk_tx_tables_free(t);
    for (auto* t : ctx->cox_tables) ria::cox_tables_free(t);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->ofdm_scratch) cudaFree(ctx->ofdm_scratch);
    if (ctx->chain_scratch) cudaFree(ctx->chain_scratch);
    if (ctx->hilbert65) cudaFree(ctx->hilbert65);
    if (ctx->work_counter) cudaFree(ctx->work_counter);
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage_dev[i]) cudaFree(ctx->stage_dev[i]);
        if (ctx->stage_pin[i]) cudaFreeHost(ctx->stage_pin[i]);
    }
    for (auto& ev : ctx->stage_ev) if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
    return RIA_OK;
}

extern "C" int ria_ctx_set_stream(ria_ctx* ctx, void* cuda_stream) {
    if (!ctx) return RIA_E_INVAL;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);     // NULL = the legacy default stream, like torch's default
    return RIA_OK;
}

extern "C" int ria_ctx_synchronize(ria_ctx* ctx) {
    if (!ctx) return RIA_E_INVAL;
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    RIA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RIA_OK;
}

extern "C" int ria_ctx_set_decode_flags(ria_ctx* ctx, int flags) {
    if (!ctx) return RIA_E_INVAL;
    if (flags & ~RIA_DECODE_FULL) return ria::set_error(ctx, RIA_E_INVAL, "ctx: unknown decode flags 0x%x", flags);
    ctx->decode_flags = flags;
    return RIA_OK;
}

extern "C" int ria_ctx_get_decode_flags(const ria_ctx* ctx) { return ctx ? ctx->decode_flags : 0; }

extern "C" int ria_ctx_set_timing(ria_ctx* ctx, int enable) {
    if (!ctx) return RIA_E_INVAL;
    for (auto& t : ctx->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
    ctx->timed.clear();
    for (int k = 0; k < 32; ++k) { ctx->timed_ms[k] = 0.0; ctx->timed_n[k] = 0; }
    ctx->timing = enable != 0;
    return RIA_OK;
}

extern "C" int ria_ctx_get_timing(ria_ctx* ctx, int kind, double* total_ms, int64_t* launches) {
    if (!ctx || !total_ms || !launches) return RIA_E_INVAL;
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    ria::fold_timed(ctx);
    const bool known = kind >= 0 && kind < 32;
    *total_ms = known ? ctx->timed_ms[kind] : 0.0;
    *launches = known ? ctx->timed_n[kind] : 0;
    return RIA_OK;
}

extern "C" const char* ria_last_error(const ria_ctx* ctx) {
    return ctx ? ctx->last_error.c_str() : "null context";
}

extern "C" int64_t ria_ctx_launch_count(const ria_ctx* ctx) { return ctx ? ctx->launches : 0; }
