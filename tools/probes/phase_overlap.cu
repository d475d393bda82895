// Micro-probe: does phase alignment (CTA-wide barrier per pass) cost throughput for the packet kernel's work items?
// Every thread runs the real mid-level item (13 x LDS.128 window -> sym5 lattice R = 22 -> 22 x STS.64 x 2) `iters` times,
// (a) with a __syncthreads() per item like the tree kernel, (b) with __syncwarp() only, (c) with nothing in between.
// Same residency as the product kernel: 256 threads, 2 CTAs per SM, ~100 KB dynamic shared memory per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audiodeepfake-detection_b200/csrc -o tools/probes/phase_overlap tools/probes/phase_overlap.cu
#include <stdio.h>

#include "afd_wpt_kernel.cuh"

namespace afd {
void set_error(const char*, ...) {}
int fail(int code, const char*, ...) { return code; }
int cuda_fail(cudaError_t e, const char*) { return static_cast<int>(e); }
int lattice_factor(const double*, int, LatticeInfo*) { return -1; }
}  // namespace afd

using namespace afd;

template <int F, int R, int MODE>
__global__ void __launch_bounds__(256, 2) probe(float* sink, const __grid_constant__ Coefs<F> cf, int iters) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    float* in = smem;                       // 256 windows, stride 2R floats
    float* out = smem + 12288;
    for (int i = tid; i < 24576; i += 256) smem[i] = 0.001f * (i & 1023);      // 2 x 12288 floats = 96 KB
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float lo[R], hi[R];
        const float* src = (it & 1) ? out : in;
        float* dst = (it & 1) ? in : out;
        filter_pair<F, R, true, true>(src + tid * 2 * R, cf, lo, hi, false, 1 << 20);
        vec_store<R>(dst + tid * R, lo);             // like the tree: consecutive items write consecutive R-chunks of a child
        vec_store<R>(dst + 6144 + tid * R, hi);
        acc += lo[0];
        if (MODE == 0) __syncthreads();
        else if (MODE == 1) __syncwarp();
    }
    if (acc == 12345.678f) sink[0] = acc;
}

// The real stored-level routine on the real plan: pass `pi` of the sym5 / coif4 level-8 plan, repeated.
template <int F, int RA, int RB>
__global__ void __launch_bounds__(256, 2) probe_mid(float* sink, const __grid_constant__ WptPlan plan,
                                                    const __grid_constant__ Coefs<F> cf, int pi, int iters) {
    extern __shared__ __align__(16) float smem[];
    for (int i = threadIdx.x; i < plan.smem_floats; i += 256) smem[i] = 0.001f * (i & 1023);
    __syncthreads();
    const Pass& ps = plan.pass[pi];
    for (int it = 0; it < iters; ++it) {
        if (ps.rsel == 0) mid_level<F, RA, true, true>(smem + ps.in_off, smem + ps.out_off, ps, cf);
        else mid_level<F, RB, true, true>(smem + ps.in_off, smem + ps.out_off, ps, cf);
        __syncthreads();
    }
    if (smem[threadIdx.x] == 12345.678f) sink[0] = 1.f;
}

template <int F>
static void run_mid(const Coefs<F>& cf, float* sink) {
    WptPlan plan;
    Tuning tu{14, 22, 26, 14, 24, true, (F / 2 - 1) * 0.5};
    if (make_plan(22050, F, 8, tu, 2, 0.9, &plan) != AFD_OK) { printf("plan failed\n"); return; }
    auto k = probe_mid<F, 22, 26>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * plan.smem_floats);
    for (int pi = 0; pi < plan.npass - 1; ++pi) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 1000;
        k<<<296, 256, 4 * plan.smem_floats>>>(sink, plan, cf, pi, 10);
        cudaEventRecord(e0);
        k<<<296, 256, 4 * plan.smem_floats>>>(sink, plan, cf, pi, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const Pass& ps = plan.pass[pi];
        printf("F=%d mid_level pass %d (level %d, %d parents, n_out %d, R %d, items %d, NI %d NE %d): %.0f cycles per CTA-pass (%s)\n", F, pi,
               pi + 2, ps.parents, ps.n_out, ps.rsel ? 26 : 22, ps.parents * ps.C, ps.NI, ps.NE, ms * 1e-3 * 1.965e9 / iters,
               cudaGetErrorString(cudaGetLastError()));
    }
}

template <int F, int R, int MODE>
static void run(const char* what, const Coefs<F>& cf, float* sink) {
    auto k = probe<F, R, MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    k<<<296, 256, 100 * 1024>>>(sink, cf, 10);
    cudaEventRecord(e0);
    k<<<296, 256, 100 * 1024>>>(sink, cf, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 2 CTAs x iters passes
    printf("F=%d R=%d %-14s %.1f us  -> %.0f cycles per CTA-pass at 1.965 GHz (%s)\n", F, R, what, ms * 1e3,
           ms * 1e-3 * 1.965e9 / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* sink;
    cudaMalloc(&sink, 4);
    Coefs<10> c10{};
    for (int i = 0; i < 5; ++i) c10.t[i] = 0.3f + 0.1f * i;
    Coefs<24> c24{};
    for (int i = 0; i < 12; ++i) c24.t[i] = 0.2f + 0.05f * i;
    run<10, 22, 0>("syncthreads", c10, sink);
    run<10, 22, 1>("syncwarp", c10, sink);
    run<10, 22, 2>("free", c10, sink);
    run<10, 14, 0>("syncthreads", c10, sink);
    run<10, 14, 2>("free", c10, sink);
    run_mid<10>(c10, sink);
    run_mid<24>(c24, sink);
    return 0;
}
