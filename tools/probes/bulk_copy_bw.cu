// Micro-probe: what does ONE SM get out of cp.async.bulk (global -> shared, mbarrier completion) when all 148 SMs stream
// disjoint rows, as a function of the copy size and of the number of copies in flight?  Compared with LDGSTS (cp.async.16)
// issued by `lw` warps.  One persistent CTA per SM, no compute: thread 0 (bulk) / the loader warps (cp.async) keep `depth`
// slots of `chunk` bytes in flight, every slot is re-armed as soon as it has landed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probes/bulk_copy_bw tools/probes/bulk_copy_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// every CTA streams `per_cta` bytes starting at x + blockIdx.x * per_cta
__global__ void __launch_bounds__(512, 1) bulk_probe(const char* __restrict__ x, long long per_cta, int chunk, int depth,
                                                      int src_off = 0, int dst_off = 0, int noise = 0, float* sink = nullptr) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) unsigned long long bars[32];
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bars + i))));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 32) {
        // noise: the other warps hammer a separate 16 KB of shared memory with 128-bit loads and stores until the copies are done
        if (!noise || threadIdx.x >= 32 + 32 * noise) return;
        float4* area = reinterpret_cast<float4*>(smem + 200 * 1024 - 16384);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int i = threadIdx.x & 1023;
        while (!stop) {
#pragma unroll 8
            for (int k = 0; k < 8; ++k) {
                const float4 v = area[i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                area[i ^ 512] = acc;
                i = (i + 37) & 1023;
            }
        }
        if (acc.x == 12345.f && sink) sink[0] = acc.y;
        return;
    }
    if (threadIdx.x != 0) return;
    const char* src = x + blockIdx.x * per_cta + src_off;
    const long long n = per_cta / chunk - 1;
    const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + dst_off;
    for (long long i = 0; i < n + depth; ++i) {
        const int slot = static_cast<int>(i % depth);
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(bars + slot));
        if (i >= depth) mbar_wait(bar, static_cast<uint32_t>(((i / depth) - 1) & 1));
        if (i < n) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s0 + slot * chunk), "l"(src + i * chunk), "r"(chunk), "r"(bar) : "memory");
        }
    }
    stop = 1;
}

// `lw` warps; each round the CTA stages `chunk` bytes with cp.async.16 and keeps `depth` committed groups in flight
__global__ void __launch_bounds__(512, 1) ldgsts_probe(const char* __restrict__ x, long long per_cta, int chunk, int depth) {
    extern __shared__ __align__(128) char smem[];
    const char* src = x + blockIdx.x * per_cta;
    const long long n = per_cta / chunk;
    const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    for (long long i = 0; i < n; ++i) {
        const int slot = static_cast<int>(i % depth);
        for (int o = 16 * threadIdx.x; o < chunk; o += 16 * blockDim.x)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s0 + slot * chunk + o), "l"(src + i * chunk + o));
        asm volatile("cp.async.commit_group;\n" ::);
        if (depth == 1) asm volatile("cp.async.wait_group 0;\n" ::);
        else if (depth == 2) asm volatile("cp.async.wait_group 1;\n" ::);
        else asm volatile("cp.async.wait_group 3;\n" ::);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long per_cta = 64LL << 20;                 // 64 MiB per SM: 9.5 GB in all, far beyond L2
    char* x = nullptr;
    if (cudaMalloc(&x, per_cta * sms) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(x, 1, per_cta * sms);
    cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ldgsts_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* what, int chunk, int depth, int lw, float ms) {
        const double gbs = double(per_cta) * sms / (ms * 1e-3) / 1e9;
        printf("%-8s chunk %6d B  depth %2d  warps %2d : %8.1f GB/s  (%5.1f GB/s per SM)\n", what, chunk, depth, lw, gbs, gbs / sms);
    };
    const int chunks[] = {2048, 8192, 16384, 45056, 90112};
    for (int chunk : chunks)
        for (int depth : {1, 2, 4, 8}) {
            if (static_cast<long long>(chunk) * depth > 192 * 1024) continue;
            float ms = 0.f;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                bulk_probe<<<sms, 32, chunk * depth + 256>>>(x, per_cta / chunk * chunk, chunk, depth);
                cudaEventRecord(e1);
                if (cudaEventSynchronize(e1) != cudaSuccess) { printf("bulk launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                cudaEventElapsedTime(&ms, e0, e1);
            }
            report("bulk", chunk, depth, 1, ms);
        }
    // alignment of the two ends and shared-memory traffic from the other warps (the packet / Haar kernels' situation)
    for (int chunk : {45056, 90112})
        for (int so : {0, 16, 64})
            for (int dof : {0, 16, 64})
                for (int noise : {0, 4, 15}) {
                    if (so != dof && noise) continue;
                    float ms = 0.f;
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        bulk_probe<<<sms, 512, 200 * 1024>>>(x, per_cta / chunk * chunk, chunk, 1, so, dof, noise, nullptr);
                        cudaEventRecord(e1);
                        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("bulk launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                        cudaEventElapsedTime(&ms, e0, e1);
                    }
                    const double gbs = double(per_cta) * sms / (ms * 1e-3) / 1e9;
                    printf("bulk     chunk %6d B  depth 1  src+%2d dst+%2d  noise warps %2d : %8.1f GB/s  (%5.1f GB/s per SM)\n", chunk, so, dof, noise, gbs, gbs / sms);
                }
    for (int lw : {8})
        for (int chunk : {16384, 45056})
            for (int depth : {1, 2, 4}) {
                float ms = 0.f;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    ldgsts_probe<<<sms, 32 * lw, chunk * depth>>>(x, per_cta / chunk * chunk, chunk, depth);
                    cudaEventRecord(e1);
                    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("ldgsts launch failed\n"); return 1; }
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                report("cp.async", chunk, depth, lw, ms);
            }
    cudaFree(x);
    return 0;
}
