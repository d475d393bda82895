// Shared helpers for libafd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/afd_b200.h"

namespace afd {

constexpr int kMaxSmemPerCta = 232448;  // 227 KB opt-in limit on sm_100
constexpr int kNumSmsFallback = 148;

constexpr int kMaxLatticeStages = 32;

// Paraunitary lattice of an orthogonal analysis filter pair (afd_lattice.cu).
struct LatticeInfo {
    int stages;                            // J = F / 2
    int reflect0;                          // stage 0 is a reflection (det = -1): kernels fall back to the direct form
    int usable;                            // residual, sign and dynamic range are fit for the fp32 kernels
    double tan_theta[kMaxLatticeStages];   // stage tangents, stage 0 first
    double scale;                          // product of the stage cosines: true = scale * unscaled lattice output
    double residual;                       // max |tap error| of the re-synthesised filter pair
    double max_abs_tan;
};
int lattice_factor(const double* dec_lo, int F, LatticeInfo* info);

// Extended STFT epilogue (afd_stft_power_ex): fused Normalize and feature moments.
struct StftExtras {
    int normalize;
    float nmean, nrstd;
    double* moments;
};

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define AFD_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::afd::cuda_fail(_e, #expr); \
    } while (0)

__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming (evict-first) global stores for write-once feature tensors
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs4(float4* p, float4 v) { __stcs(p, v); }

// log(|c|^power + offset) -- the reference's epilogue (wavelet_math.py:209, :66 for the STFT).
// power == 2 is the only value the reference's experiments use; pow(x, 2.0) is x*x exactly.
// lg2.approx has <= 2 ulp error on the normal range; after the ln2 scale the result is within 1e-6 absolute
// of logf on the feature range [-27.7, 10], far inside the 1e-4 parity budget.
__device__ __forceinline__ float ln_approx(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r * 0.69314718055994530942f;
}
__device__ __forceinline__ float log_power(float c, float power, float offset, bool square) {
    const float pw = square ? fmaf(c, c, offset) : __powf(fabsf(c), power) + offset;
    return ln_approx(pw);
}

}  // namespace afd
