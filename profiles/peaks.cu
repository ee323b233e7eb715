// Measured SM-side roofline denominators for the kernels SURVEY.md 8(d) bounds by SM throughput rather
// than HBM: fp32 FMA rate, fp32 non-fused (separate multiply / add) instruction rate -- the parity contract
// forbids FMA contraction in most of the chain, so that is the rate those kernels can reach -- packed
// FFMA2 / FADD2 rate, and shared-memory load bandwidth (128-bit conflict-free LDS).
// MEASURED_PEAKS.json (driver-written) only has the HBM copy bandwidth and the bf16 GEMM rate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/peaks_bench profiles/peaks.cu
//   gpurun -- 'profiles/peaks_bench > gpurun_out/sm_peaks.json'
//
// Each figure is the best of 5 timed launches (CUDA events) after a warm-up launch; grid = SMs x resident
// CTAs so that every SM sub-partition has 8+ warps of independent work.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kIters = 4096;
constexpr int kAcc = 8;

__global__ void __launch_bounds__(256) fma_kernel(float* out, float a, float b) {
    float acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kAcc; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

// separate multiply and add (what --fmad=false code issues): 2 instructions per "mad"
__global__ void __launch_bounds__(256) mul_add_kernel(float* out, float a, float b) {
    float acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kAcc; ++i) acc[i] = __fadd_rn(__fmul_rn(acc[i], a), b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

// packed fp32: fma.rn.f32x2 (FFMA2), two flops x two lanes per instruction
__global__ void __launch_bounds__(256) fma2_kernel(float* out, float a, float b) {
    unsigned long long acc[kAcc];
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < kAcc; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v)); }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kAcc; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    if (s == 12345.678f) out[0] = s;
}

// conflict-free 128-bit shared loads; every warp streams over a 16 KB tile
__global__ void __launch_bounds__(256) lds_kernel(float* out) {
    __shared__ float4 tile[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int idx = threadIdx.x;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 v = tile[(idx + u * 256) & 1023];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        idx = (idx + 32) & 1023;
    }
    if (s.x + s.y + s.z + s.w == 12345.678f) out[0] = s.x;
}

template <typename F>
static double best_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, static_cast<double>(ms));
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float* out;
    CK(cudaMalloc(&out, 64));
    const int ctas = sms * 8, threads = 256;
    const double n_thr = static_cast<double>(ctas) * threads;
    const double fma_ms = best_ms([&] { fma_kernel<<<ctas, threads>>>(out, 1.0001f, 0.5f); });
    const double ma_ms = best_ms([&] { mul_add_kernel<<<ctas, threads>>>(out, 1.0001f, 0.5f); });
    const double fma2_ms = best_ms([&] { fma2_kernel<<<ctas, threads>>>(out, 1.0001f, 0.5f); });
    const double lds_ms = best_ms([&] { lds_kernel<<<ctas, threads>>>(out); });
    const double n_op = n_thr * kIters * kAcc;
    const double fma_tflops = 2.0 * n_op / (fma_ms * 1e-3) / 1e12;
    const double ma_tinstr = 2.0 * n_op / (ma_ms * 1e-3) / 1e12;            // instructions = flops here
    const double fma2_tflops = 4.0 * n_op / (fma2_ms * 1e-3) / 1e12;
    const double lds_tbs = n_thr * kIters * 4.0 * 16.0 / (lds_ms * 1e-3) / 1e12;
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_attr\": %.0f, "
           "\"fp32_fma_tflops\": %.2f, \"fp32_unfused_tinstr_per_s\": %.2f, \"fp32_unfused_tflops\": %.2f, "
           "\"fp32_packed_fma2_tflops\": %.2f, \"smem_lds128_tb_per_s\": %.2f, "
           "\"how\": \"profiles/peaks.cu: %d CTAs x 256 threads, %d x %d independent ops per thread, best of 5 (CUDA events)\"}\n",
           prop.name, sms, clk / 1000.0, fma_tflops, ma_tinstr, ma_tinstr, fma2_tflops, lds_tbs, ctas, kIters, kAcc);
    return 0;
}
