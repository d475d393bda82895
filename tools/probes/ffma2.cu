// Micro-probe: packed fp32x2 FMA (fma.rn.f32x2 -> FFMA2) vs scalar FFMA throughput on sm_100a, lattice-shaped stream.
#include <cuda_runtime.h>
#include <stdio.h>

struct Taps { float t[16]; };

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// scalar: 16 (u, v) pairs rotated by 16 taps per iteration: 512 FFMA
__global__ void __launch_bounds__(256, 2) probe_scalar(float* sink, const __grid_constant__ Taps taps, int inner) {
    float u[16], v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { u[i] = threadIdx.x + i; v[i] = threadIdx.x - i; }
    for (int it = 0; it < inner; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float t = taps.t[k];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float un = fmaf(t, v[i], u[i]);
                v[i] = fmaf(-t, u[i], v[i]);
                u[i] = un;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += u[i] + v[i];
    if (s == 12345.678f) sink[0] = s;
}

// packed: the same 16 pairs as 8 packed (u_i, u_{i+8}) / (v_i, v_{i+8}) operands: 256 FFMA2
__global__ void __launch_bounds__(256, 2) probe_packed(float* sink, const __grid_constant__ Taps taps, int inner) {
    unsigned long long u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = pack(threadIdx.x + i, threadIdx.x + i + 8); v[i] = pack(threadIdx.x - i, threadIdx.x - i - 8); }
    for (int it = 0; it < inner; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float t = taps.t[k];
            const unsigned long long tp = pack(t, t), tn = pack(-t, -t);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned long long un = fma2(tp, v[i], u[i]);
                v[i] = fma2(tn, u[i], v[i]);
                u[i] = un;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b, c, d; unpack(u[i], a, b); unpack(v[i], c, d); s += a + b + c + d; }
    if (s == 12345.678f) sink[0] = s;
}

template <typename K>
double run(K kern, float* sink, const Taps& taps, int sms) {
    const int inner = 512, grid = sms * 2 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, 256>>>(sink, taps, inner);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<grid, 256>>>(sink, taps, inner);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 2 * 16 * 16 * double(inner) * 256.0 * grid * 10;
    return flops / (ms * 1e-3) / 1e12;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* sink; cudaMalloc(&sink, 4);
    Taps taps;
    for (int k = 0; k < 16; ++k) taps.t[k] = 1e-3f * (k + 1);
    for (int rep = 0; rep < 3; ++rep)
        printf("fp32 TFLOP/s  scalar FFMA %.2f   packed FFMA2 %.2f\n", run(probe_scalar, sink, taps, sms), run(probe_packed, sink, taps, sms));
    return 0;
}
