// Micro-probe: FFMA throughput by operand form on sm_100a (3-register vs uniform-register multiplier).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_forms ffma_forms.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdio.h>

struct Taps { float t[16]; };

// lattice-like: pairs (u, v) rotated by uniform taps: u' = u + t v ; v' = v - t u   (two dependent-free FFMAs per pair)
template <bool UR>
__global__ void __launch_bounds__(256, 2) probe(float* sink, const __grid_constant__ Taps taps, const float* gt, int inner) {
    float u[16], v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { u[i] = threadIdx.x + i; v[i] = threadIdx.x - i; }
    float tr[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) tr[k] = gt[k];          // register copies (loaded from global: not uniform-provable)
    for (int it = 0; it < inner; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float t = UR ? taps.t[k] : tr[k];
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

template <bool UR>
double run(float* sink, const Taps& taps, const float* gt, int sms) {
    const int inner = 512, grid = sms * 2 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<UR><<<grid, 256>>>(sink, taps, gt, inner);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) probe<UR><<<grid, 256>>>(sink, taps, gt, inner);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 2 * 16 * 16 * double(inner) * 256.0 * grid * 10;
    return flops / (ms * 1e-3) / 1e12;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *sink, *gt; cudaMalloc(&sink, 4); cudaMalloc(&gt, 64);
    Taps taps; float h[16];
    for (int k = 0; k < 16; ++k) { taps.t[k] = 1e-3f * (k + 1); h[k] = taps.t[k]; }
    cudaMemcpy(gt, h, 64, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep)
        printf("FFMA TFLOP/s  register-multiplier %.2f   uniform-register-multiplier %.2f\n",
               run<false>(sink, taps, gt, sms), run<true>(sink, taps, gt, sms));
    return 0;
}
