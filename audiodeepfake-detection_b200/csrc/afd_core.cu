// Error plumbing, version, and the FP32-FMA peak probe of libafd_b200.
#include <stdarg.h>
#include <string.h>

#include "afd_common.cuh"

namespace afd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return static_cast<int>(e);
}

// 8 independent FFMA chains per thread, tap operands from the constant bank like the filter-bank kernels.
__global__ void __launch_bounds__(256, 4) fma_probe_kernel(float* sink, float a, float b, int inner) {
    float v0 = threadIdx.x, v1 = v0 + 1.f, v2 = v0 + 2.f, v3 = v0 + 3.f, v4 = v0 + 4.f, v5 = v0 + 5.f, v6 = v0 + 6.f,
          v7 = v0 + 7.f;
    for (int i = 0; i < inner; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            v0 = fmaf(v0, a, b); v1 = fmaf(v1, a, b); v2 = fmaf(v2, a, b); v3 = fmaf(v3, a, b);
            v4 = fmaf(v4, a, b); v5 = fmaf(v5, a, b); v6 = fmaf(v6, a, b); v7 = fmaf(v7, a, b);
        }
    }
    if (v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7 == 12345.678f) sink[0] = v0;
}

}  // namespace afd

using namespace afd;

extern "C" int afd_version(void) { return 200; }

// Build provenance: sha256 of csrc/* + include/afd_b200.h as computed by build.py (source_hash()) when this unit was
// compiled; build.py reads the marker back from the file to decide whether the binary matches the sources on disk.
#ifndef AFD_SOURCE_HASH
#define AFD_SOURCE_HASH "unknown"
#endif
static const char kSourceHash[] = "AFD_SOURCE_HASH=" AFD_SOURCE_HASH;
extern "C" const char* afd_source_hash(void) { return kSourceHash + 16; }

extern "C" const char* afd_last_error(void) { return g_err; }

extern "C" int afd_measure_fp32_fma_tflops(int iters, double* tflops, void* stream) {
    if (!tflops || iters < 1) return fail(AFD_ERR_INVALID_ARG, "afd_measure_fp32_fma_tflops: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    AFD_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int inner = 4096;
    const int grid = sms * 8;
    float ms = 0.f;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        fma_probe_kernel<<<grid, 256, 0, s>>>(sink, 0.999f, 0.001f, inner);  // warm-up
        e = cudaEventRecord(e0, s);
    }
    if (e == cudaSuccess) {
        for (int i = 0; i < iters; ++i) fma_probe_kernel<<<grid, 256, 0, s>>>(sink, 0.999f, 0.001f, inner);
        e = cudaEventRecord(e1, s);
    }
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);          // released on every path
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (e != cudaSuccess) return cuda_fail(e, "afd_measure_fp32_fma_tflops");
    const double flops = 2.0 * 8 * 16 * double(inner) * 256.0 * grid * iters;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return AFD_OK;
}
