// Haar wavelet-packet "fingerprint": sum over clips and positions of |c| for every level-L packet.
//
// Replaces pywt.WaveletPacket(clips, "haar", mode="reflect").get_level(14, order="freq") + np.stack + np.abs +
// the summation inside np.mean  (reference scripts/freq_visual/fingerprints.py:101-115).
//
// For the 2-tap Haar filter the analysis step on a node x of length m is
//     lo[k] = s (x[2k] + x[2k+1]),  hi[k] = s (x[2k] - x[2k+1]),  s = 1/sqrt(2),  k < ceil(m/2)
// with the reflect extension x[m] := x[m-2] supplying the partner of the last sample when m is odd (pywt /
// ptwt pad nothing on the left for F = 2).  The tree is computed IN PLACE in shared memory with the
// Walsh-Hadamard addressing: after l levels, element i of node o (o = path read LSB-first: bit j is the
// filter chosen at level j+1) sits at position o + i * 2^l.  A node of odd length needs one extra slot
// (position o + m * 2^l) which is free by construction, so the whole level-l tree occupies 2^l * L_l floats.
// The last level is never stored: each thread adds |lo|, |hi| to register accumulators that persist across
// all clips the CTA processes (thread t always meets the same nodes because the thread count divides 2^(L-1));
// one double-precision atomicAdd per packet and CTA publishes them at the end.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "afd_common.cuh"

namespace afd {

constexpr int kHaarThreads = 512;
constexpr int kHaarMaxAcc = 32;   // 2 * 2^(L-1) / threads accumulators per thread  -> L <= 14
constexpr int kHaarMaxLevel = 14;

struct HaarPlan {
    int N, L;
    int n[kHaarMaxLevel + 1];
    int buf_floats;
};

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

__global__ void __launch_bounds__(kHaarThreads, 2)
haar_fingerprint_kernel(const float* __restrict__ x, long long x_row_stride, long long B,
                        double* __restrict__ sums, const __grid_constant__ HaarPlan plan) {
    extern __shared__ __align__(16) float buf[];
    const int tid = threadIdx.x;
    const int L = plan.L;
    const float s = 0.70710678118654752440f;
    float acc[kHaarMaxAcc];
#pragma unroll
    for (int i = 0; i < kHaarMaxAcc; ++i) acc[i] = 0.f;

    const int Sl = 1 << (L - 1);                  // stride (= node count) entering the last level
    const int m_last = plan.n[L - 1];             // node length entering the last level
    const int K_last = m_last >> 1;
    const bool odd_last = m_last & 1;

    for (long long clip = blockIdx.x; clip < B; clip += gridDim.x) {
        const float* xg = x + clip * x_row_stride;
        // ---- load the clip
        if ((reinterpret_cast<uintptr_t>(xg) & 7) == 0) {
            const int pairs = plan.N >> 1;
            for (int i = tid; i < pairs; i += kHaarThreads) cp_async_8(buf + 2 * i, xg + 2 * i);
            if ((plan.N & 1) && tid == 0) cp_async_4(buf + plan.N - 1, xg + plan.N - 1);
        } else {
            for (int i = tid; i < plan.N; i += kHaarThreads) cp_async_4(buf + i, xg + i);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // ---- levels 1 .. L-1 in place
        for (int l = 0; l < L - 1; ++l) {
            const int S = 1 << l;
            const int m = plan.n[l];
            const int K = m >> 1;              // full pairs per node
            const bool odd = m & 1;
            const int total = K << l;          // S * K
            for (int u = tid; u < total; u += kHaarThreads) {
                const int o = u & (S - 1);
                const int k = u >> l;
                const int i = o + ((2 * k) << l);
                const float a = buf[i], b = buf[i + S];
                buf[i] = s * (a + b);
                buf[i + S] = s * (a - b);
                if (odd && k == K - 1) {       // tail sample x[m-1] pairs with the reflected x[m-2] = b
                    const float c = buf[i + 2 * S];
                    buf[i + 2 * S] = s * (c + b);
                    buf[i + 3 * S] = s * (c - b);
                }
            }
            __syncthreads();
        }
        // ---- last level: accumulate |lo| (node o) and |hi| (node o + Sl) instead of storing
        if (Sl >= kHaarThreads) {
#pragma unroll
            for (int j = 0; j < kHaarMaxAcc / 2; ++j) {
                const int o = tid + j * kHaarThreads;
                if (o < Sl) {
                    float alo = 0.f, ahi = 0.f, b = 0.f;
                    for (int k = 0; k < K_last; ++k) {
                        const float a = buf[o + ((2 * k) << (L - 1))];
                        b = buf[o + ((2 * k + 1) << (L - 1))];
                        alo += fabsf(s * (a + b));
                        ahi += fabsf(s * (a - b));
                    }
                    if (odd_last) {
                        const float c = buf[o + ((m_last - 1) << (L - 1))];
                        alo += fabsf(s * (c + b));
                        ahi += fabsf(s * (c - b));
                    }
                    acc[2 * j] += alo;
                    acc[2 * j + 1] += ahi;
                }
            }
        } else if (tid < Sl) {                // small trees: one node pair per thread
            float alo = 0.f, ahi = 0.f, b = 0.f;
            for (int k = 0; k < K_last; ++k) {
                const float a = buf[tid + ((2 * k) << (L - 1))];
                b = buf[tid + ((2 * k + 1) << (L - 1))];
                alo += fabsf(s * (a + b));
                ahi += fabsf(s * (a - b));
            }
            if (odd_last) {
                const float c = buf[tid + ((m_last - 1) << (L - 1))];
                alo += fabsf(s * (c + b));
                ahi += fabsf(s * (c - b));
            }
            acc[0] += alo;
            acc[1] += ahi;
        }
        __syncthreads();   // buf is overwritten by the next clip
    }
    // ---- publish: node id (LSB-first path) -> natural index (MSB-first) -> frequency position (Gray decode)
#pragma unroll
    for (int j = 0; j < kHaarMaxAcc / 2; ++j) {
        const int o = tid + j * kHaarThreads;
        const bool live = (Sl >= kHaarThreads) ? (o < Sl) : (j == 0 && tid < Sl);
        if (live) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned node = static_cast<unsigned>(o) + (h ? Sl : 0);
                unsigned nat = bitrev(node, L);
                unsigned p = nat;                       // Gray decode: p = nat ^ (nat>>1) ^ (nat>>2) ...
                for (int sft = 1; sft < L; sft <<= 1) p ^= p >> sft;
                atomicAdd(sums + p, static_cast<double>(acc[2 * j + h]));
            }
        }
    }
}


// ================================================================================================
// Fast path (levels 11 .. 14, the reference's level 14 included): three register-blocked passes.
//
// The tree is still in place in shared memory with the Walsh-Hadamard addressing (node o, element i of level l at
// logical position o + i * 2^l), but a thread now carries 32 elements through FIVE levels in registers, so the
// whole clip makes three round trips through shared memory instead of fourteen:
//     pass A  levels 1-5    thread <-> 32 consecutive samples (one padded row, 8 x LDS.128 / STS.128)
//     pass B  levels 6-10   lanes <-> the 32 level-5 nodes, warp <-> block of 32 elements of each node
//     pass C  levels 11-L   lanes <-> the 1024 level-10 nodes; |c| goes straight into register accumulators
// Odd node lengths (reflect: x~[m] = x[m-2]) are removed up front: whenever a level has an odd number of elements
// the signal is extended by a copy of its second-to-last element's samples, so every pass sees whole blocks.  The
// extension of the clip (<= 31 samples) is fetched from global memory with the clip, the one of the level-5 nodes
// is written by the pass-A threads that produce the source elements, the one of the level-10 nodes is a redirected
// load in pass C (tables built on the host).  The 1/sqrt(2) per level is applied once at the end (2^-7 for L = 14).
// Physical layout: 4 floats of padding after every 32 (row stride 36 floats = odd number of 16-byte units), which
// makes all three access patterns bank-conflict free.
// ================================================================================================
constexpr int kFastThreads = 256;
#ifndef AFD_HAAR_FLUSH_EVERY
#define AFD_HAAR_FLUSH_EVERY 64      // 32 -> 128 is worth +3 % at 32768-clip launches; 64 keeps the fp32 partial sums (128 terms) well inside 1e-5
#endif
constexpr int kFlushEvery = AFD_HAAR_FLUSH_EVERY;     // clips between flushes of the fp32 register accumulators into the fp64 sums

struct HaarFastPlan {
    int N, L;
    int n5, n10, nL;            // node lengths after 5, 10 and L levels
    int extA, extB, extC;       // elements appended in front of pass A / B / C
    int tabA[32], tabB[32], tabC[32];   // source element of every appended element
    int buf_floats;             // physical shared-memory floats
    float final_scale;          // (1/sqrt 2)^L
};

// Butterfly stages.  From span 2 on, two adjacent elements ride in one 64-bit operand and the sum / difference are
// packed fp32x2 adds (add.rn.f32x2 / sub.rn.f32x2 -> FADD2 on sm_100): half the issue slots of the scalar form.
// AFD_HAAR_FADD2=0 builds the scalar butterflies (A/B measurements).
#ifndef AFD_HAAR_FADD2
#define AFD_HAAR_FADD2 1
#endif
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int SPAN, int LEN>
__device__ __forceinline__ void haar_stage_n(float (&v)[LEN]) {
    if constexpr (SPAN >= 2 && AFD_HAAR_FADD2 != 0) {
#pragma unroll
        for (int p = 0; p < LEN; p += 2)
            if ((p & SPAN) == 0) {
                const u64 a = pk2(v[p], v[p + 1]), b = pk2(v[p + SPAN], v[p + SPAN + 1]);
                upk2(add2(a, b), v[p], v[p + 1]);
                upk2(sub2(a, b), v[p + SPAN], v[p + SPAN + 1]);
            }
    } else {
#pragma unroll
        for (int p = 0; p < LEN; ++p)
            if ((p & SPAN) == 0) {
                const float a = v[p], b = v[p + SPAN];
                v[p] = a + b;
                v[p + SPAN] = a - b;
            }
    }
}
template <int SPAN>
__device__ __forceinline__ void haar_stage(float (&v)[32]) { haar_stage_n<SPAN, 32>(v); }
template <int SPAN>
__device__ __forceinline__ void haar_stage16(float (&v)[16]) { haar_stage_n<SPAN, 16>(v); }

template <int K>   // K = L - 10 levels in the last pass
__global__ void __launch_bounds__(kFastThreads, 2)
haar_fast_kernel(const float* __restrict__ x, long long x_row_stride, long long B, double* __restrict__ sums,
                 const __grid_constant__ HaarFastPlan plan) {
    extern __shared__ __align__(16) float buf[];
    constexpr int BLK = 1 << K;                 // elements per pass-C block
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    float acc[4][BLK];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < BLK; ++c) acc[q][c] = 0.f;

    auto flush = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned o = static_cast<unsigned>(tid + kFastThreads * q);       // level-10 node (LSB-first path)
#pragma unroll
            for (int c = 0; c < BLK; ++c) {
                const unsigned node = o + (static_cast<unsigned>(c) << 10);
                const unsigned nat = bitrev(node, plan.L);
                unsigned p = nat;
                for (int sft = 1; sft < plan.L; sft <<= 1) p ^= p >> sft;
                atomicAdd(sums + p, static_cast<double>(acc[q][c]) * static_cast<double>(plan.final_scale));
                acc[q][c] = 0.f;
            }
        }
    };

    int since_flush = 0;
    for (long long clip = blockIdx.x; clip < B; clip += gridDim.x) {
        const float* xg = x + clip * x_row_stride;
        // ---- load the clip into the padded layout (sample s -> s + 4 * (s / 32)) plus its <= 31 appended samples
        {
            const int N = plan.N;
            const unsigned mis = static_cast<unsigned>(reinterpret_cast<uintptr_t>(xg));
            if ((mis & 15) == 0) {
                const int units = N >> 2;
                for (int i = tid; i < units; i += kFastThreads) cp_async_16(buf + 4 * i + ((i >> 3) << 2), xg + 4 * i);
                if (tid < (N & 3)) { const int s = 4 * units + tid; cp_async_4(buf + s + ((s >> 5) << 2), xg + s); }
            } else if ((mis & 7) == 0) {
                const int units = N >> 1;
                for (int i = tid; i < units; i += kFastThreads) cp_async_8(buf + 2 * i + ((i >> 4) << 2), xg + 2 * i);
                if ((N & 1) && tid == 0) { const int s = N - 1; cp_async_4(buf + s + ((s >> 5) << 2), xg + s); }
            } else {
                for (int s = tid; s < N; s += kFastThreads) cp_async_4(buf + s + ((s >> 5) << 2), xg + s);
            }
            if (tid < plan.extA) { const int s = N + tid; cp_async_4(buf + s + ((s >> 5) << 2), xg + plan.tabA[tid]); }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        // ---- pass A: levels 1-5, one padded row per item
        for (int b = tid; b < plan.n5; b += kFastThreads) {
            float v[32];
            float4* row = reinterpret_cast<float4*>(buf + 36 * b);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 q4 = row[u];
                v[4 * u] = q4.x; v[4 * u + 1] = q4.y; v[4 * u + 2] = q4.z; v[4 * u + 3] = q4.w;
            }
            haar_stage<1>(v); haar_stage<2>(v); haar_stage<4>(v); haar_stage<8>(v); haar_stage<16>(v);
#pragma unroll
            for (int u = 0; u < 8; ++u) row[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            if (b + 32 >= plan.n5) {            // this element may be the source of an appended level-5 element
                for (int e = 0; e < plan.extB; ++e)
                    if (plan.tabB[e] == b) {
                        float4* dst = reinterpret_cast<float4*>(buf + 36 * (plan.n5 + e));
#pragma unroll
                        for (int u = 0; u < 8; ++u) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    }
            }
        }
        __syncthreads();
        // ---- pass B: levels 6-10, lane = level-5 node, warp item = block of 32 elements
        for (int b = warp; b < plan.n10; b += kFastThreads / 32) {
            float v[32];
            float* base = buf + lane + 1152 * b;          // element 32b + j of node `lane` sits at base + 36 j
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = base[36 * j];
            haar_stage<1>(v); haar_stage<2>(v); haar_stage<4>(v); haar_stage<8>(v); haar_stage<16>(v);
#pragma unroll
            for (int j = 0; j < 32; ++j) base[36 * j] = v[j];
        }
        __syncthreads();
        // ---- pass C: levels 11-L, thread item = (level-10 node, block); |c| accumulates in registers
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int o = tid + kFastThreads * q;
            const float* base = buf + o + ((o >> 5) << 2);            // element i of node o sits at base + 1152 i
            for (int b = 0; b < plan.nL; ++b) {
                float v[16];
                const bool redirect = (b + 1) * BLK > plan.n10;
#pragma unroll
                for (int j = 0; j < BLK; ++j) {
                    int i = b * BLK + j;
                    if (redirect && i >= plan.n10) i = plan.tabC[i - plan.n10];
                    v[j] = base[1152 * i];
                }
                if (K >= 1) haar_stage16<1>(v);
                if (K >= 2) haar_stage16<2>(v);
                if (K >= 3) haar_stage16<4>(v);
                if (K >= 4) haar_stage16<8>(v);
#pragma unroll
                for (int c = 0; c < BLK; ++c) acc[q][c] += fabsf(v[c]);
            }
        }
        if (++since_flush == kFlushEvery) { flush(); since_flush = 0; }
        __syncthreads();   // buf is overwritten by the next clip
    }
    if (since_flush) flush();
}

// ================================================================================================
// Streaming variant of the fast path (r2): ONE persistent CTA of 512 threads per SM, TWO clip buffers, ONE CTA-wide barrier
// per clip.  The fast kernel above runs load -> pass A -> pass B -> pass C in lock step (three barriers per clip, every warp
// in the same phase: the shared-memory pipe and the adders take turns; ncu: 42 % issue utilisation, 18 % barrier stalls, 23 %
// of the samples in the load region).  Here
//   * passes A and B of a 32-row block need only that block: the warp that owns block b stages its 4 KB itself (its own
//     cp.async group), runs pass A on its rows and pass B on the same 32 x 32 tile with __syncwarp() in between;
//   * pass C of clip n (reads the whole buffer) and the staging + passes A/B of clip n+1 (other buffer) happen in the SAME
//     barrier interval, every warp at its own pace: loads are in flight during pass C, and one warp's butterflies overlap
//     another warp's shared-memory traffic;
//   * the copies are issued by a few LOADER warps that do nothing else: 88 KB of cp.async per clip exceed what an SM keeps in
//     flight, so the issuing warp blocks in the LSU queue for most of the transfer (phase timing of a first version in
//     which every warp staged its own blocks: 5.2 k of 10.7 k cycles per clip spent issuing).  Completion is signalled per
//     block through an mbarrier (cp.async.mbarrier.arrive.noinc) that the block's worker warp waits on;
//   * blocks (2.4 cost units) and pass-C node groups (1 unit) are dealt to the worker warps by a host-side greedy schedule.
// Same layout, same butterflies and the same per-thread accumulation order over the clips of a CTA as the fast kernel.
// ================================================================================================
#ifndef AFD_HAAR_LOADER_WARPS
#define AFD_HAAR_LOADER_WARPS 4
#endif
#ifndef AFD_HAAR_PHASE_TIMING
#define AFD_HAAR_PHASE_TIMING 0
#endif
#if AFD_HAAR_PHASE_TIMING
static __device__ unsigned long long g_haar_phase[8];
static __device__ unsigned long long g_haar_warp[32];
static __device__ unsigned long long g_haar_wphase[32][4];      // linear kernel, per warp: pass C, chunk waits, passes A/B, barrier
#define AFD_HAAR_MARK(slot)                                                                          \
    do {                                                                                             \
        if (lane == 0) {                                                                             \
            const long long now_ = clock64();                                                        \
            atomicAdd(&g_haar_wphase[warp & 31][slot], static_cast<unsigned long long>(now_ - t0_)); \
            if (warp == 0 || warp == 15) atomicAdd(&g_haar_phase[(slot) + (warp ? 4 : 0)], static_cast<unsigned long long>(now_ - t0_)); \
            t0_ = now_;                                                                              \
        }                                                                                            \
    } while (0)
#else
#define AFD_HAAR_MARK(slot) do { } while (0)
#endif
constexpr int kStreamThreads = 512;
constexpr int kStreamWarps = kStreamThreads / 32;
constexpr int kLoaderWarps = AFD_HAAR_LOADER_WARPS;      // warps that only issue the cp.async of the next clip
constexpr int kWorkerWarps = kStreamWarps - kLoaderWarps;
constexpr int kMaxBlocksPerWarp = 4;      // n10 <= 64 blocks
constexpr int kMaxGroupsPerWarp = 3;      // 32 node groups of 32 level-10 nodes over the worker warps (>= 11)

struct HaarStreamPlan {
    HaarFastPlan fast;
    int ext_block;                                    // block that holds the appended samples / level-5 elements
    signed char blocks[kStreamWarps][kMaxBlocksPerWarp];   // -1: none
    signed char groups[kStreamWarps][kMaxGroupsPerWarp];
};

template <int K>
__global__ void __launch_bounds__(kStreamThreads, 1)
haar_stream_kernel(const float* __restrict__ x, long long x_row_stride, long long B, double* __restrict__ sums,
                   const __grid_constant__ HaarStreamPlan sp) {
    extern __shared__ __align__(16) float smem_h[];
    constexpr int BLK = 1 << K;
    const HaarFastPlan& plan = sp.fast;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n5 = plan.n5, n10 = plan.n10, N = plan.N;
    float acc[kMaxGroupsPerWarp][BLK];
#pragma unroll
    for (int q = 0; q < kMaxGroupsPerWarp; ++q)
#pragma unroll
        for (int c = 0; c < BLK; ++c) acc[q][c] = 0.f;

    auto flush = [&]() {
#pragma unroll
        for (int q = 0; q < kMaxGroupsPerWarp; ++q) {
            const int g = sp.groups[warp][q];
            if (g < 0) continue;
            const unsigned o = static_cast<unsigned>(32 * g + lane);       // level-10 node (LSB-first path)
#pragma unroll
            for (int c = 0; c < BLK; ++c) {
                const unsigned node = o + (static_cast<unsigned>(c) << 10);
                const unsigned nat = bitrev(node, plan.L);
                unsigned p = nat;
                for (int sft = 1; sft < plan.L; sft <<= 1) p ^= p >> sft;
                atomicAdd(sums + p, static_cast<double>(acc[q][c]) * static_cast<double>(plan.final_scale));
                acc[q][c] = 0.f;
            }
        }
    };

    // Staging of block b (samples [1024 b, 1024 b + 1024), rows 32 b .. 32 b + 31 of the padded layout) by a loader warp:
    // cp.async of the widest size the clip's alignment allows (a clip is 88,200 B: odd clips are only 8-byte aligned), then
    // every lane arrives on the block's mbarrier when its copies have landed.  (Measured alternatives: one TMA bulk copy per
    // 128-byte row -- the rows are 144 bytes apart in the padded layout -- costs ~95 cycles per copy and is 2.9x slower.)
    auto stage_block = [&](float* buf, uint32_t bar, const float* xg, int b) {
        const int s0 = 1024 * b;
        const unsigned mis = static_cast<unsigned>(reinterpret_cast<uintptr_t>(xg));
        if ((mis & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = 256 * b + 32 * j + lane;                     // float4 index in the clip
                if (4 * i + 3 < N) cp_async_16(buf + 4 * i + ((i >> 3) << 2), xg + 4 * i);
            }
            const int tail0 = N & ~3;                                      // the clip's last 1 .. 3 samples
            if (tail0 >= s0 && tail0 < s0 + 1024 && lane < (N & 3)) { const int s = tail0 + lane; cp_async_4(buf + s + ((s >> 5) << 2), xg + s); }
        } else if ((mis & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = 512 * b + 32 * j + lane;                     // float2 index
                if (2 * i + 1 < N) cp_async_8(buf + 2 * i + ((i >> 4) << 2), xg + 2 * i);
            }
            if ((N & 1) && N - 1 >= s0 && N - 1 < s0 + 1024 && lane == 0) { const int s = N - 1; cp_async_4(buf + s + ((s >> 5) << 2), xg + s); }
        } else {
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const int s = s0 + 32 * j + lane;
                if (s < N) cp_async_4(buf + s + ((s >> 5) << 2), xg + s);
            }
        }
        if (b == sp.ext_block && lane < plan.extA) { const int s = N + lane; cp_async_4(buf + s + ((s >> 5) << 2), xg + plan.tabA[lane]); }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
    };

    // passes A and B of block b, in place, by the warp that staged it
    auto pass_ab = [&](float* buf, int b, int) {
        const int r = 32 * b + lane;                        // pass A: lane = row
        if (r < n5) {
            float v[32];
            {
                const float4* row = reinterpret_cast<const float4*>(buf + 36 * r);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 q4 = row[u];
                    v[4 * u] = q4.x; v[4 * u + 1] = q4.y; v[4 * u + 2] = q4.z; v[4 * u + 3] = q4.w;
                }
            }
            haar_stage<1>(v); haar_stage<2>(v); haar_stage<4>(v); haar_stage<8>(v); haar_stage<16>(v);
            float4* row = reinterpret_cast<float4*>(buf + 36 * r);
#pragma unroll
            for (int u = 0; u < 8; ++u) row[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            if (b == sp.ext_block) {                        // this element may be the source of an appended level-5 element
                for (int e = 0; e < plan.extB; ++e)
                    if (plan.tabB[e] == r) {
                        float4* dst = reinterpret_cast<float4*>(buf + 36 * (n5 + e));
#pragma unroll
                        for (int u = 0; u < 8; ++u) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    }
            }
        }
        __syncwarp();
        {                                                   // pass B: lane = level-5 node, the block's 32 elements
            float v[32];
            float* base = buf + lane + 1152 * b;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = base[36 * j];
            haar_stage<1>(v); haar_stage<2>(v); haar_stage<4>(v); haar_stage<8>(v); haar_stage<16>(v);
#pragma unroll
            for (int j = 0; j < 32; ++j) base[36 * j] = v[j];
        }
    };

    float* const buf0 = smem_h;
    float* const buf1 = smem_h + plan.buf_floats;
    // per buffer and block: an mbarrier that completes when the loader warp's copies of that block have landed
    const uint32_t mbar0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem_h + 2 * plan.buf_floats));
    auto mbar = [&](int bufi, int b) { return mbar0 + 8u * static_cast<uint32_t>(bufi * (kStreamWarps * kMaxBlocksPerWarp) + b); };
    if (tid < 2 * n10) {
        const uint32_t a = mbar(tid / n10, tid % n10);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const bool loader = warp >= kWorkerWarps;
    // loader warp l stages blocks l, l + 2, ... of a clip
    auto load_clip = [&](float* buf, int bufi, const float* xg) {
        for (int b = warp - kWorkerWarps; b < n10; b += kLoaderWarps) stage_block(buf, mbar(bufi, b), xg, b);
    };
    auto shift_of = [&](const float* xg) { return (static_cast<unsigned>(reinterpret_cast<uintptr_t>(xg)) & 15u) ? 2 : 0; };
    auto wait_block = [&](int bufi, int b, uint32_t parity) {
        const uint32_t a = mbar(bufi, b);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(a), "r"(parity) : "memory");
        }
    };
    int cur = 0;
    int since_flush = 0;
    long long it = 0;                                       // clips this CTA has started: buffer it & 1, use it >> 1
    if (blockIdx.x < B) {
        const float* xg = x + blockIdx.x * x_row_stride;
        if (loader) load_clip(buf0, 0, xg);
        else {
#pragma unroll
            for (int t = 0; t < kMaxBlocksPerWarp; ++t)
                if (sp.blocks[warp][t] >= 0) { wait_block(0, sp.blocks[warp][t], 0); pass_ab(buf0, sp.blocks[warp][t], shift_of(xg)); }
        }
    }
    __syncthreads();
#if AFD_HAAR_PHASE_TIMING
    long long t0_ = clock64();
#endif
    for (long long clip = blockIdx.x; clip < B; clip += gridDim.x, cur ^= 1, ++it) {
        float* const buf = cur ? buf1 : buf0;
        float* const nbuf = cur ? buf0 : buf1;
        const bool more = clip + gridDim.x < B;
        if (loader) {
            // the other buffer's pass C ended before the last barrier
            if (more) load_clip(nbuf, cur ^ 1, x + (clip + gridDim.x) * x_row_stride);
            AFD_HAAR_MARK(0);
        } else {
            // ---- pass C of this clip: levels 11-L, lane = level-10 node of one of the warp's groups; |c| accumulates in registers
#pragma unroll
            for (int q = 0; q < kMaxGroupsPerWarp; ++q) {
                const int g = sp.groups[warp][q];
                if (g < 0) continue;
                const int o = 32 * g + lane;
                const float* base = buf + o + ((o >> 5) << 2);            // element i of node o sits at base + 1152 i
                for (int b = 0; b < plan.nL; ++b) {
                    float v[16];
                    const bool redirect = (b + 1) * BLK > n10;
#pragma unroll
                    for (int j = 0; j < BLK; ++j) {
                        int i = b * BLK + j;
                        if (redirect && i >= n10) i = plan.tabC[i - n10];
                        v[j] = base[1152 * i];
                    }
                    if (K >= 1) haar_stage16<1>(v);
                    if (K >= 2) haar_stage16<2>(v);
                    if (K >= 3) haar_stage16<4>(v);
                    if (K >= 4) haar_stage16<8>(v);
#pragma unroll
                    for (int c = 0; c < BLK; ++c) acc[q][c] += fabsf(v[c]);
                }
            }
            if (++since_flush == kFlushEvery) { flush(); since_flush = 0; }
            AFD_HAAR_MARK(0);
            // ---- passes A and B of the next clip's blocks, each as soon as its copies have landed
            if (more) {
                const uint32_t parity = static_cast<uint32_t>(((it + 1) >> 1) & 1);
                const int nsh = shift_of(x + (clip + gridDim.x) * x_row_stride);
#pragma unroll
                for (int t = 0; t < kMaxBlocksPerWarp; ++t)
                    if (sp.blocks[warp][t] >= 0) { wait_block(cur ^ 1, sp.blocks[warp][t], parity); pass_ab(nbuf, sp.blocks[warp][t], nsh); }
            }
        }
        AFD_HAAR_MARK(2);
        __syncthreads();
        AFD_HAAR_MARK(3);
    }
    if (since_flush) flush();
}

// Shared-memory accesses by 32-bit shared address + compile-time byte offset (volatile: program order is kept).  The linear
// kernel forms its XOR-swizzled addresses as (aligned base ^ small constant), one LOP3 per access instead of XOR + scale + add.
template <int IMM>
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(IMM), "f"(x), "f"(y) : "memory");
}
template <int IMM>
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ void sts32(uint32_t a, float x) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(IMM), "f"(x) : "memory");
}

template <int SB, int... J>
__device__ __forceinline__ void linear_b_load(uint32_t tl, float (&v)[32], std::integer_sequence<int, J...>) {
    ((v[J] = lds32<SB + 128 * J>(tl ^ (8u * J)), v[J + 16] = lds32<SB + 128 * (J + 16)>(tl ^ (8u * J))), ...);
}
template <int SB, int... M>
__device__ __forceinline__ void linear_b_store(uint32_t ts, const float (&v)[32], std::integer_sequence<int, M...>) {
    ((sts32<SB>(ts ^ (128u * M), v[M]), sts32<SB + 2048>(ts ^ (128u * M), v[M + 16])), ...);
}

// Passes A and B of one 32-row block of the linear layout (see haar_linear_kernel), SH = the clip's shift in floats (0 / 2).
// `blk` = shared address of the block's first float WITHOUT the shift (4096-byte aligned), rows_valid = rows to transform.
template <int SH>
__device__ __forceinline__ void haar_linear_block(uint32_t blk, int lane, int rows_valid) {
    constexpr int SB = 4 * SH;                              // shift in bytes: an immediate of every access
    const int k = lane & 15;
    if (lane < rows_valid) {                                // pass A: lane = row, float2 column u ^ k in step u
        float v[32];
        const uint32_t t = (blk + 128u * lane) ^ (8u * k);  // row base is 128-byte aligned: base + 8 (u ^ k) = (base ^ 8k) ^ 8u
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const float2 q2 = lds64<SB>(t ^ (8u * u));
            v[2 * u] = q2.x; v[2 * u + 1] = q2.y;
        }
        haar_stage<1>(v); haar_stage<2>(v); haar_stage<4>(v); haar_stage<8>(v); haar_stage<16>(v);
#pragma unroll
        for (int u = 0; u < 16; ++u) sts64<SB>(t ^ (8u * u), v[2 * u], v[2 * u + 1]);
    }
    __syncwarp();
    {                                                       // pass B: lane = level-5 node c, the block's 32 elements (rows)
        float v[32];
        const uint32_t tl = blk + 4u * lane;                // float (lane ^ 2 (j & 15)) of row j: (blk + 4 lane) ^ 8 (j & 15) + 128 j
        linear_b_load<SB>(tl, v, std::make_integer_sequence<int, 16>{});
        haar_stage<1>(v); haar_stage<2>(v); haar_stage<4>(v); haar_stage<8>(v); haar_stage<16>(v);
        __syncwarp();                                       // every lane has read the tile before anybody overwrites it
        const uint32_t ts = tl ^ (128u * (lane >> 1));      // row (m & 16) | ((m ^ hl) & 15), column lane (blk is 4096-byte aligned)
        linear_b_store<SB>(ts, v, std::make_integer_sequence<int, 16>{});
    }
}

// ================================================================================================
// Linear-layout variant (r2): the streaming kernel above still spends a quarter of its warps on issuing cp.async (each
// LDGSTS occupies its warp for ~90 cycles, and the copies crowd the same LSU queue the workers' LDS / STS go through).
// A whole clip can be staged by ONE bulk copy (TMA engine, no thread involved, nothing in the LSU queue) -- if it may land
// contiguously.  The 36-float row padding of the layout above exists only for pass A (a lane owns a 128-byte row; 32 lanes
// reading the same float4 column of 32 rows is an 8-way bank conflict).  Here the clip stays LINEAR and pass A reads
// float2 column (u ^ k) of its row in step u, k = row & 15: the 16 lanes of a half warp hit 16 distinct float2 columns.
// The Haar stages are Walsh-Hadamard butterflies, and a WHT of an XOR-permuted input is the WHT times a Walsh sign:
//     in'[q] = in[q ^ k]  =>  out'[h] = (-1)^<k, h> out[h]
// so pass A runs the unchanged butterflies on the permuted registers and writes them back in place: row r now holds
// output pair hq at float2 column hq ^ k with sign (-1)^<k, hq>.  Pass B (lane = column c, WHT along the 32 rows j of a
// block) reads row j at float (c ^ 2 (j & 15)), i.e. its input carries the sign (-1)^<j & 15, c >> 1> -- a Walsh function of
// j -- and a WHT of a Walsh-modulated input is the WHT XOR-permuted: register m holds output m ^ (c >> 1).  Pass B stores
// register m to row m ^ (c >> 1), plain column c: signs and permutations cancel with no extra arithmetic, and pass C reads a
// plain [element][node] array.  Odd clips (88,200 B apart: 8-byte aligned) are copied from 8 bytes earlier; all indices are
// then shifted by sh = 2 floats, which the float2 accesses of pass A and the scalar accesses of passes B / C absorb.
// All 18 warps work; blocks and node groups are dealt by the same greedy schedule (measured costs: 1.35 : 1).
// ================================================================================================
#ifndef AFD_HAAR_LIN_CHUNKS
#define AFD_HAAR_LIN_CHUNKS 4
#endif
constexpr int kLinChunks = AFD_HAAR_LIN_CHUNKS;             // bulk copies (and mbarriers) per clip
#ifndef AFD_HAAR_EARLY_PAIRS
#define AFD_HAAR_EARLY_PAIRS 1
#endif
#ifndef AFD_HAAR_BLOCK_COST
#define AFD_HAAR_BLOCK_COST 1.35
#endif
#ifndef AFD_HAAR_LIN_THREADS
#define AFD_HAAR_LIN_THREADS 576     // 18 warps: 22 blocks + 32 node groups split 4 x (2 blocks + 1 group) / 14 x (1 block + 2 groups); 512: +5 % time, 640 / 704: spills
#endif
constexpr int kLinThreads = AFD_HAAR_LIN_THREADS;
constexpr int kLinWarps = kLinThreads / 32;
constexpr int kLinMaxWarps = 32;
#ifndef AFD_HAAR_LIN_GROUPS
#define AFD_HAAR_LIN_GROUPS (AFD_HAAR_LIN_THREADS / 32 >= 18 ? 2 : 3)
#endif
constexpr int kLinGroups = AFD_HAAR_LIN_GROUPS;   // node groups a warp may own (acc registers)

struct HaarLinearPlan {
    HaarFastPlan fast;
    int ext_block;
    int buf_floats;                                   // floats per clip buffer (linear clip + shift + appended samples)
    int chunk_first[kLinChunks + 1];                  // chunk c covers blocks [chunk_first[c], chunk_first[c + 1])
    signed char blocks[kLinMaxWarps][kMaxBlocksPerWarp];
    signed char groups[kLinMaxWarps][kMaxGroupsPerWarp];
};

// HEAD: the reference's shape (22050-sample clips, level 14: 22 level-10 elements per node, appended elements 22, 23 = copies of
// 18, 19 and 24 .. 31 = copies of 8 .. 15): pass C's element offsets are compile-time immediates.
__host__ __device__ constexpr int head_elem(int i) { return i < 22 ? i : (i < 24 ? i - 4 : i - 16); }

template <int K, bool HEAD>
__global__ void __launch_bounds__(kLinThreads, 1)
haar_linear_kernel(const float* __restrict__ x, long long x_row_stride, long long B, double* __restrict__ sums,
                   const __grid_constant__ HaarLinearPlan sp) {
    extern __shared__ __align__(128) float smem_h[];
    constexpr int BLK = 1 << K;
    const HaarFastPlan& plan = sp.fast;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n5 = plan.n5, n10 = plan.n10, N = plan.N;
    float acc[kLinGroups][BLK];
#pragma unroll
    for (int q = 0; q < kLinGroups; ++q)
#pragma unroll
        for (int c = 0; c < BLK; ++c) acc[q][c] = 0.f;

    // (Measured alternative: a private [grid][2^L] slab with plain fp64 read-modify-write + a reduce kernel instead of the
    // 148 x 16384 atomics per flush is 9 % SLOWER -- the RED.ADD.F64 are fire-and-forget, the slab's loads are not.)
    auto flush = [&]() {
#pragma unroll
        for (int q = 0; q < kLinGroups; ++q) {
            const int g = sp.groups[warp][q];
            if (g < 0) continue;
            const unsigned o = static_cast<unsigned>(32 * g + lane);       // level-10 node (LSB-first path)
#pragma unroll
            for (int c = 0; c < BLK; ++c) {
                const unsigned node = o + (static_cast<unsigned>(c) << 10);
                const unsigned nat = bitrev(node, plan.L);
                unsigned p = nat;
                for (int sft = 1; sft < plan.L; sft <<= 1) p ^= p >> sft;
                atomicAdd(sums + p, static_cast<double>(acc[q][c]) * static_cast<double>(plan.final_scale));
                acc[q][c] = 0.f;
            }
        }
    };

    // the clip buffers are 4096-byte aligned (the XOR addressing of haar_linear_block relies on it)
    float* const buf0 = smem_h + ((4096u - (static_cast<uint32_t>(__cvta_generic_to_shared(smem_h)) & 4095u)) & 4095u) / 4;
    float* const buf1 = buf0 + sp.buf_floats;
    const uint32_t mbar0 = static_cast<uint32_t>(__cvta_generic_to_shared(buf0 + 2 * sp.buf_floats));
    auto mbar = [&](int bufi, int c) { return mbar0 + 8u * static_cast<uint32_t>(bufi * kLinChunks + c); };
    if (tid < 2 * kLinChunks) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar0 + 8u * tid), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    auto shift_of = [&](const float* xg) { return (static_cast<unsigned>(reinterpret_cast<uintptr_t>(xg)) & 15u) ? 2 : 0; };

    // Stage a clip (warp 0): chunk c = one bulk copy of the floats [4 f0, 4 f1) of the 16-byte aligned array that starts
    // sh floats before the clip; what does not fill a 16-byte unit at the clip's end is moved by lanes (never reads past x).
    // The last unit of every clip but the batch's last one is copied whole (it runs up to 3 floats into the row padding / the
    // next clip, x_row_stride >= N; the surplus lands where pass A's appended samples are written afterwards); only the
    // last clip moves its <= 3 tail samples by lanes -- a load -> store pair in front of the bulk copies would hold them
    // back by a DRAM round trip.
    auto load_clip = [&](float* buf, int bufi, const float* xg, bool last_clip) {
        const int sh = shift_of(xg);
        const float* src = xg - sh;                                        // 16-byte aligned
        const int whole4 = (N + sh) >> 2;                                  // whole float4 units inside the clip
        const int total4 = last_clip ? whole4 : (N + sh + 3) >> 2;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic accesses of the free buffer before the async writes
        __syncwarp();
        if (lane < kLinChunks) {
            const int c = lane;
            int f0 = (1024 * sp.chunk_first[c]) >> 2, f1 = (1024 * sp.chunk_first[c + 1]) >> 2;
            if (c == kLinChunks - 1 || f1 > total4) f1 = total4;
            if (f0 > total4) f0 = total4;
            // shifted clips: float index = sample + 2, so chunk boundaries (sample multiples of 1024) sit 2 floats into a
            // unit; a unit that straddles two chunks is counted with the LATER chunk's barrier only through f0's rounding:
            // every unit belongs to exactly one copy because consecutive chunks share f1 == next f0.
            const uint32_t bytes = static_cast<uint32_t>(f1 - f0) * 16u;
            const uint32_t bar = mbar(bufi, c);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            if (bytes) {
                const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(buf + 4 * f0));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src + 4 * f0), "r"(bytes), "r"(bar) : "memory");
            }
        }
        if (last_clip)
            for (int s = 4 * whole4 - sh + lane; s < N; s += 32) buf[sh + s] = __ldg(xg + s);     // <= 3 tail samples
    };
    auto wait_chunk = [&](int bufi, int c, uint32_t parity) {
        const uint32_t a = mbar(bufi, c);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(a), "r"(parity) : "memory");
        }
    };
    // a block's samples [1024 b - ?, ...): with the shift its first two floats belong to the previous chunk's last unit, so a
    // block waits for its own chunk and (shifted clips, first block of a chunk) the previous one
    auto wait_block = [&](int bufi, int b, int sh, uint32_t parity) {
        int c = 0;
#pragma unroll
        for (int t = 1; t < kLinChunks; ++t) c += b >= sp.chunk_first[t] ? 1 : 0;
        wait_chunk(bufi, c, parity);
        if (sh != 0 && c + 1 < kLinChunks && b + 1 == sp.chunk_first[c + 1]) wait_chunk(bufi, c + 1, parity);   // its last 2 samples
        if (b == sp.ext_block) wait_chunk(bufi, kLinChunks - 1, parity);                                          // tail samples
    };

    // passes A and B of block b, in place, by one warp
    auto pass_ab = [&](float* buf, int b, int sh) {
        float* const lin = buf + sh;                        // sample s of the clip at lin[s]
        if (b == sp.ext_block) {
            // Appended samples (copies of samples of this block, host-checked), then the appended level-5 elements: instead
            // of copying a source row's pass-A OUTPUT (which would need the target row's sign / column convention), the
            // source row's raw samples are duplicated and the otherwise idle lanes of this block transform the copies.
            if (lane < plan.extA) lin[N + lane] = lin[plan.tabA[lane]];
            __syncwarp();
            for (int e = 0; e < plan.extB; ++e) lin[32 * (n5 + e) + lane] = lin[32 * plan.tabB[e] + lane];
            __syncwarp();
        }
        const uint32_t blk = static_cast<uint32_t>(__cvta_generic_to_shared(buf)) + 4096u * b;
        const int rows_valid = n5 + plan.extB - 32 * b;     // >= 32 for every block but the last
        if (sh) haar_linear_block<2>(blk, lane, rows_valid);
        else haar_linear_block<0>(blk, lane, rows_valid);
    };

    int cur = 0;
    int since_flush = 0;
    long long it = 0;
    if (blockIdx.x < B) {
        const float* xg = x + blockIdx.x * x_row_stride;
        if (warp == 0) {
            load_clip(buf0, 0, xg, blockIdx.x == B - 1);
        }
        if (blockIdx.x == B - 1) __syncthreads();      // CTA-uniform: the last clip's tail samples are stored by warp 0's lanes
        const int sh = shift_of(xg);
#pragma unroll
        for (int t = 0; t < kMaxBlocksPerWarp; ++t)
            if (sp.blocks[warp][t] >= 0) { wait_block(0, sp.blocks[warp][t], sh, 0); pass_ab(buf0, sp.blocks[warp][t], sh); }
    }
    __syncthreads();
#if AFD_HAAR_PHASE_TIMING
    long long t0_ = clock64();
#endif
    for (long long clip = blockIdx.x; clip < B; clip += gridDim.x, cur ^= 1, ++it) {
        float* const buf = cur ? buf1 : buf0;
        float* const nbuf = cur ? buf0 : buf1;
        const bool more = clip + gridDim.x < B;
        const float* xn = x + (clip + gridDim.x) * x_row_stride;
#if AFD_HAAR_PHASE_TIMING
        const long long it0_ = clock64();
#endif
        if (more && warp == 0) load_clip(nbuf, cur ^ 1, xn, clip + gridDim.x == B - 1);   // the other buffer's pass C ended before the last barrier
        // ---- pass C of this clip: levels 11-L, lane = level-10 node of one of the warp's groups; |c| accumulates in registers
        const float* lin = buf + shift_of(x + clip * x_row_stride);
#pragma unroll
        for (int q = 0; q < kLinGroups; ++q) {
            const int g = sp.groups[warp][q];
            if (g < 0) continue;
            const float* base = lin + 32 * g + lane;                      // element i of node o = 32 g + lane at base[1024 i]
            if constexpr (HEAD) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = base[1024 * head_elem(16 * b + j)];
                    haar_stage16<1>(v); haar_stage16<2>(v); haar_stage16<4>(v); haar_stage16<8>(v);
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[q][c] += fabsf(v[c]);
                }
            } else {
                for (int b = 0; b < plan.nL; ++b) {
                    float v[16];
                    const bool redirect = (b + 1) * BLK > n10;
#pragma unroll
                    for (int j = 0; j < BLK; ++j) {
                        int i = b * BLK + j;
                        if (redirect && i >= n10) i = plan.tabC[i - n10];
                        v[j] = base[1024 * i];
                    }
                    if (K >= 1) haar_stage16<1>(v);
                    if (K >= 2) haar_stage16<2>(v);
                    if (K >= 3) haar_stage16<4>(v);
                    if (K >= 4) haar_stage16<8>(v);
#pragma unroll
                    for (int c = 0; c < BLK; ++c) acc[q][c] += fabsf(v[c]);
                }
            }
        }
        if (++since_flush == kFlushEvery) { flush(); since_flush = 0; }
        AFD_HAAR_MARK(0);
        // ---- passes A and B of the next clip's blocks, each as soon as its chunk has landed
        if (more) {
            const uint32_t parity = static_cast<uint32_t>(((it + 1) >> 1) & 1);
            const int nsh = shift_of(xn);
#pragma unroll
            for (int t = 0; t < kMaxBlocksPerWarp; ++t)
                if (sp.blocks[warp][t] >= 0) {
                    wait_block(cur ^ 1, sp.blocks[warp][t], nsh, parity);
                    AFD_HAAR_MARK(1);
                    pass_ab(nbuf, sp.blocks[warp][t], nsh);
                    AFD_HAAR_MARK(2);
                }
        }
#if AFD_HAAR_PHASE_TIMING
        if (lane == 0) atomicAdd(&g_haar_warp[warp], static_cast<unsigned long long>(clock64() - it0_));
#endif
        __syncthreads();
        AFD_HAAR_MARK(3);
    }
    if (since_flush) flush();
}

static bool make_stream_schedule(int n10, int workers, int max_groups, signed char (*blocks)[kMaxBlocksPerWarp], signed char (*groups)[kMaxGroupsPerWarp],
                                 double block_cost, bool early_pairs = false) {
    if (n10 > workers * kMaxBlocksPerWarp || 32 > workers * max_groups) return false;
    double load[kLinMaxWarps] = {0};
    int nb[kLinMaxWarps] = {0}, ng[kLinMaxWarps] = {0};
    for (int w = 0; w < kLinMaxWarps; ++w) {
        for (int t = 0; t < kMaxBlocksPerWarp; ++t) blocks[w][t] = -1;
        for (int t = 0; t < kMaxGroupsPerWarp; ++t) groups[w][t] = -1;
        if (w >= workers) { nb[w] = kMaxBlocksPerWarp; ng[w] = max_groups; }
    }
    auto least = [&](const int* cnt, int cap) {
        int best = -1;
        for (int w = 0; w < kLinMaxWarps; ++w)
            if (cnt[w] < cap && (best < 0 || load[w] < load[best])) best = w;
        return best;
    };
    const int pairs = early_pairs && n10 > workers && n10 <= 2 * workers && kMaxBlocksPerWarp >= 2 ? n10 - workers : 0;
    if (pairs > 0) {
        // Linear kernel: the warps that own two blocks are the critical path of a clip (pass C, then both blocks, each as soon as
        // its chunk has landed), so they take the blocks that land FIRST (w and pairs + w: chunks 0 / 1); the one-block warps run two
        // node groups of pass C before they need their block and take the later ones, the last (appended rows) among them.
        for (int w = 0; w < workers; ++w) {
            if (w < pairs) {
                blocks[w][nb[w]++] = static_cast<signed char>(w);
                blocks[w][nb[w]++] = static_cast<signed char>(pairs + w);
                load[w] += 2 * block_cost;
            } else {
                blocks[w][nb[w]++] = static_cast<signed char>(pairs + w);
                load[w] += block_cost;
            }
        }
    } else {
        for (int b = 0; b < n10; ++b) {
            const int w = least(nb, kMaxBlocksPerWarp);
            if (w < 0) return false;
            blocks[w][nb[w]++] = static_cast<signed char>(b);
            load[w] += block_cost;
        }
    }
    for (int g = 0; g < 32; ++g) {
        const int w = least(ng, max_groups);
        if (w < 0) return false;
        groups[w][ng[w]++] = static_cast<signed char>(g);
        load[w] += 1.0;
    }
    return true;
}

static bool make_linear_plan(const HaarFastPlan& fp, HaarLinearPlan* lp) {
    lp->fast = fp;
    lp->ext_block = (fp.n5 - 1) >> 5;
    if (fp.extA > 32 || ((fp.N + fp.extA - 1) >> 10) != lp->ext_block || ((fp.N - 1) >> 10) != lp->ext_block) return false;
    for (int e = 0; e < fp.extB; ++e)
        if ((fp.tabB[e] >> 5) != lp->ext_block || ((fp.n5 + e) >> 5) != lp->ext_block) return false;
    for (int e = 0; e < fp.extA; ++e)
        if ((fp.tabA[e] >> 10) != lp->ext_block) return false;
    if (lp->ext_block != fp.n10 - 1) return false;                        // the tail lives in the last block (and last chunk)
    lp->buf_floats = (1024 * fp.n10 + 2 + 1023) / 1024 * 1024;            // linear clip incl. appended rows, + shift; 4096-byte multiple
    for (int c = 0; c <= kLinChunks; ++c) lp->chunk_first[c] = static_cast<int>(static_cast<long long>(fp.n10) * c / kLinChunks);
    // measured (phase table): a block's passes A + B take ~1.75 k cycles, a node group's pass C ~1.3 k
    return make_stream_schedule(fp.n10, kLinWarps, kLinGroups < kMaxGroupsPerWarp ? kLinGroups : kMaxGroupsPerWarp, lp->blocks, lp->groups,
                                AFD_HAAR_BLOCK_COST, AFD_HAAR_EARLY_PAIRS != 0);
}

// Greedy schedule: blocks (cost 2.4) then node groups (cost 1) to the least loaded warp.
static bool make_stream_plan(const HaarFastPlan& fp, HaarStreamPlan* sp) {
    sp->fast = fp;
    if (fp.n10 > kWorkerWarps * kMaxBlocksPerWarp) return false;
    sp->ext_block = (fp.n5 - 1) >> 5;
    if (fp.extA > 32 || ((fp.N + fp.extA - 1) >> 10) != sp->ext_block || (fp.N >> 10) != sp->ext_block) return false;
    for (int e = 0; e < fp.extB; ++e)
        if ((fp.tabB[e] >> 5) != sp->ext_block || ((fp.n5 + e) >> 5) != sp->ext_block) return false;
    for (int e = 0; e < fp.extA; ++e)
        if ((fp.tabA[e] >> 10) != sp->ext_block) return false;     // appended samples are copied inside their block
    double load[kStreamWarps] = {0};
    int nb[kStreamWarps] = {0}, ng[kStreamWarps] = {0};
    for (int w = kWorkerWarps; w < kStreamWarps; ++w) { nb[w] = kMaxBlocksPerWarp; ng[w] = kMaxGroupsPerWarp; }   // loader warps take no work
    for (int w = 0; w < kStreamWarps; ++w) {
        for (int t = 0; t < kMaxBlocksPerWarp; ++t) sp->blocks[w][t] = -1;
        for (int t = 0; t < kMaxGroupsPerWarp; ++t) sp->groups[w][t] = -1;
    }
    auto least = [&](const int* cnt, int cap) {
        int best = -1;
        for (int w = 0; w < kStreamWarps; ++w)
            if (cnt[w] < cap && (best < 0 || load[w] < load[best])) best = w;
        return best;
    };
    for (int b = 0; b < fp.n10; ++b) {
        const int w = least(nb, kMaxBlocksPerWarp);
        if (w < 0) return false;
        sp->blocks[w][nb[w]++] = static_cast<signed char>(b);
        load[w] += 2.4;
    }
    for (int g = 0; g < 32; ++g) {
        const int w = least(ng, kMaxGroupsPerWarp);
        if (w < 0) return false;
        sp->groups[w][ng[w]++] = static_cast<signed char>(g);
        load[w] += 1.0;
    }
    return true;
}

// Appended elements that make `levels` halvings of a node of `n` elements block-regular: whenever the current
// level holds an odd number m of elements (each covering `w` base elements) the reflect rule pairs the last one
// with element m-2, i.e. the base range of element m-2 is appended.  tab[e] = source of appended base element n+e.
static int build_extension(int n, int levels, int* tab) {
    int ext = 0, m = n, w = 1;
    for (int j = 0; j < levels; ++j) {
        if (m & 1) {
            for (int t = 0; t < w; ++t) tab[ext + t] = w * (m - 2) + t;
            ext += w;
            ++m;
        }
        m >>= 1;
        w <<= 1;
    }
    return ext;
}

static bool make_fast_plan(int64_t N, int level, HaarFastPlan* p) {
    if (level < 11 || level > 14 || N < 2048) return false;
    p->N = static_cast<int>(N);
    p->L = level;
    int n[kHaarMaxLevel + 1];
    n[0] = p->N;
    for (int l = 1; l <= level; ++l) n[l] = (n[l - 1] + 1) / 2;
    for (int l = 0; l < level; ++l) if (n[l] < 2) return false;
    p->n5 = n[5]; p->n10 = n[10]; p->nL = n[level];
    p->extA = build_extension(n[0], 5, p->tabA);
    p->extB = build_extension(n[5], 5, p->tabB);
    p->extC = build_extension(n[10], level - 10, p->tabC);
    if (n[0] + p->extA != 32 * n[5] || n[5] + p->extB != 32 * n[10] || n[10] + p->extC != (n[level] << (level - 10))) return false;
    // sources must be original elements (never appended ones) and, for pass B, among the last 32 elements
    for (int e = 0; e < p->extA; ++e) if (p->tabA[e] < 0 || p->tabA[e] >= n[0]) return false;
    for (int e = 0; e < p->extB; ++e) if (p->tabB[e] < n[5] - 32 || p->tabB[e] < 0 || p->tabB[e] >= n[5]) return false;
    for (int e = 0; e < p->extC; ++e) if (p->tabC[e] < 0 || p->tabC[e] >= n[10]) return false;
    const long long logical_max = 1024LL * n[10];     // = 32 * (n5 + extB) >= n0 + extA; pass C redirects instead of storing
    const long long physical = logical_max + ((logical_max + 31) / 32) * 4 + 64;
    if (physical * 4 > 228 * 1024 / 2 - 1024) return false;       // two CTAs per SM or the generic kernel
    p->buf_floats = static_cast<int>(physical);
    p->final_scale = static_cast<float>(pow(0.70710678118654752440, level));
    return true;
}

__global__ void add_count_kernel(long long* count, long long v) { *count += v; }

}  // namespace afd

using namespace afd;

extern "C" int afd_haar_fingerprint_accum(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int level,
                                          double* sums, int64_t* count, void* stream) {
    if (!sums || (!x && B != 0)) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_accum: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_accum: bad B/N/stride");
    if (level < 1 || level > kHaarMaxLevel)
        return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_accum: level %d not in 1..%d", level, kHaarMaxLevel);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HaarFastPlan fp;
    if (N <= (1 << 24) && make_fast_plan(N, level, &fp)) {
        if (B == 0) return AFD_OK;
        int dev = 0, sms = kNumSmsFallback;
        AFD_CUDA_TRY(cudaGetDevice(&dev));
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t smem = 4ull * fp.buf_floats;
        const char* impl = getenv("AFD_HAAR_IMPL");       // "stream" / "fast": the earlier kernels (A/B, cross-checks)
        HaarLinearPlan lp;
        if (!impl && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_row_stride & 1) == 0 && make_linear_plan(fp, &lp) &&
            2 * 4ull * lp.buf_floats + 8 * 2 * kLinChunks + 4096 <= static_cast<size_t>(kMaxSmemPerCta)) {
            long long grid = sms;
            if (grid > B) grid = B;
            const size_t lsmem = 2 * 4ull * lp.buf_floats + 8 * 2 * kLinChunks + 4096;      // + alignment slack
#define AFD_HAAR_LINEAR(KK, HH)                                                                                      \
            {                                                                                                        \
                static thread_local bool configured[16] = {false};                                                   \
                if (dev >= 16 || !configured[dev]) {                                                                 \
                    AFD_CUDA_TRY(cudaFuncSetAttribute(haar_linear_kernel<KK, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta)); \
                    if (dev < 16) configured[dev] = true;                                                            \
                }                                                                                                    \
                haar_linear_kernel<KK, HH><<<static_cast<unsigned>(grid), kLinThreads, lsmem, s>>>(               \
                    x, static_cast<long long>(x_row_stride), static_cast<long long>(B), sums, lp);                    \
            }
            bool head = level == 14 && fp.n10 == 22 && fp.nL == 2 && fp.extC == 10;
            for (int e = 0; head && e < 10; ++e) head = fp.tabC[e] == head_elem(22 + e);
            switch (level - 10) {
                case 1: AFD_HAAR_LINEAR(1, false) break;
                case 2: AFD_HAAR_LINEAR(2, false) break;
                case 3: AFD_HAAR_LINEAR(3, false) break;
                case 4: if (head) AFD_HAAR_LINEAR(4, true) else AFD_HAAR_LINEAR(4, false) break;
            }
#undef AFD_HAAR_LINEAR
            AFD_CUDA_TRY(cudaGetLastError());
#if AFD_HAAR_PHASE_TIMING
            {
                unsigned long long h[8];
                cudaDeviceSynchronize();
                cudaMemcpyFromSymbol(h, g_haar_phase, sizeof(h));
                const double clips = static_cast<double>(B);
                unsigned long long hw[32];
                cudaMemcpyFromSymbol(hw, g_haar_warp, sizeof(hw));
                fprintf(stderr, "haar linear: cycles from iteration start to barrier arrival per warp:");
                for (int w = 0; w < 16; ++w) fprintf(stderr, " %.0f", hw[w] / static_cast<double>(B));
                fprintf(stderr, "\n");
                memset(hw, 0, sizeof(hw));
                cudaMemcpyToSymbol(g_haar_warp, hw, sizeof(hw));
                {
                    static unsigned long long wp[32][4];
                    cudaMemcpyFromSymbol(wp, g_haar_wphase, sizeof(wp));
                    for (int w = 0; w < 20; ++w)
                        fprintf(stderr, "  warp %2d: C %5.0f  wait %5.0f  AB %5.0f  barrier %5.0f\n", w, wp[w][0] / clips, wp[w][1] / clips,
                                wp[w][2] / clips, wp[w][3] / clips);
                    memset(wp, 0, sizeof(wp));
                    cudaMemcpyToSymbol(g_haar_wphase, wp, sizeof(wp));
                }
                fprintf(stderr, "haar linear phases (cycles per clip; C, wait, AB, barrier) warp0: %.0f %.0f %.0f %.0f | warp15: %.0f %.0f %.0f %.0f\n",
                        h[0] / clips, h[1] / clips, h[2] / clips, h[3] / clips, h[4] / clips, h[5] / clips, h[6] / clips, h[7] / clips);
                memset(h, 0, sizeof(h));
                cudaMemcpyToSymbol(g_haar_phase, h, sizeof(h));
            }
#endif
            if (count) {
                add_count_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(count), static_cast<long long>(B) * fp.nL);
                AFD_CUDA_TRY(cudaGetLastError());
            }
            return AFD_OK;
        }
        HaarStreamPlan sp;
        if (2 * smem + 2 * 8 * kStreamWarps * kMaxBlocksPerWarp <= static_cast<size_t>(kMaxSmemPerCta) && !(impl && strcmp(impl, "fast") == 0) &&
            (reinterpret_cast<uintptr_t>(x) & 7) == 0 && (x_row_stride & 1) == 0 && make_stream_plan(fp, &sp)) {
            long long grid = sms;
            if (grid > B) grid = B;
#define AFD_HAAR_STREAM(KK)                                                                                          \
            case KK: {                                                                                               \
                static thread_local bool configured[16] = {false};                                                   \
                if (dev >= 16 || !configured[dev]) {                                                                 \
                    AFD_CUDA_TRY(cudaFuncSetAttribute(haar_stream_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta)); \
                    if (dev < 16) configured[dev] = true;                                                            \
                }                                                                                                    \
                haar_stream_kernel<KK><<<static_cast<unsigned>(grid), kStreamThreads, 2 * smem + 2 * 8 * kStreamWarps * kMaxBlocksPerWarp, s>>>(                \
                    x, static_cast<long long>(x_row_stride), static_cast<long long>(B), sums, sp);                    \
                break;                                                                                               \
            }
            switch (level - 10) { AFD_HAAR_STREAM(1) AFD_HAAR_STREAM(2) AFD_HAAR_STREAM(3) AFD_HAAR_STREAM(4) }
#undef AFD_HAAR_STREAM
            AFD_CUDA_TRY(cudaGetLastError());
#if AFD_HAAR_PHASE_TIMING
            {
                unsigned long long h[8];
                cudaDeviceSynchronize();
                cudaMemcpyFromSymbol(h, g_haar_phase, sizeof(h));
                const double clips = static_cast<double>(B);
                fprintf(stderr, "haar stream phases (cycles per clip; issue+C, wait, AB, barrier) warp0: %.0f %.0f %.0f %.0f | warp15: %.0f %.0f %.0f %.0f\n",
                        h[0] / clips, h[1] / clips, h[2] / clips, h[3] / clips, h[4] / clips, h[5] / clips, h[6] / clips, h[7] / clips);
                memset(h, 0, sizeof(h));
                cudaMemcpyToSymbol(g_haar_phase, h, sizeof(h));
            }
#endif
            if (count) {
                add_count_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(count), static_cast<long long>(B) * fp.nL);
                AFD_CUDA_TRY(cudaGetLastError());
            }
            return AFD_OK;
        }
        long long grid = 2LL * sms;
        if (grid > B) grid = B;
#define AFD_HAAR_FAST(KK)                                                                                            \
        case KK: {                                                                                                   \
            static thread_local bool configured[16] = {false};                                                       \
            if (dev >= 16 || !configured[dev]) {                                                                     \
                AFD_CUDA_TRY(cudaFuncSetAttribute(haar_fast_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta)); \
                AFD_CUDA_TRY(cudaFuncSetAttribute(haar_fast_kernel<KK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
                if (dev < 16) configured[dev] = true;                                                                \
            }                                                                                                        \
            haar_fast_kernel<KK><<<static_cast<unsigned>(grid), kFastThreads, smem, s>>>(                            \
                x, static_cast<long long>(x_row_stride), static_cast<long long>(B), sums, fp);                        \
            break;                                                                                                   \
        }
        switch (level - 10) { AFD_HAAR_FAST(1) AFD_HAAR_FAST(2) AFD_HAAR_FAST(3) AFD_HAAR_FAST(4) }
#undef AFD_HAAR_FAST
        AFD_CUDA_TRY(cudaGetLastError());
        if (count) {
            add_count_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(count), static_cast<long long>(B) * fp.nL);
            AFD_CUDA_TRY(cudaGetLastError());
        }
        return AFD_OK;
    }
    HaarPlan plan;
    plan.N = static_cast<int>(N);
    plan.L = level;
    plan.n[0] = plan.N;
    for (int l = 1; l <= level; ++l) {
        plan.n[l] = (plan.n[l - 1] + 1) / 2;
        if (l < level && plan.n[l] < 2)
            return fail(AFD_ERR_REFLECT_PAD, "afd_haar_fingerprint_accum: node length %d at level %d is too short for reflect padding", plan.n[l], l);
    }
    // in-place pass l -> l+1 (l = 0 .. L-2) touches positions below 2^l * (n[l] rounded up to even)
    long long need = N;
    for (int l = 0; l + 1 < level; ++l) {
        const long long fl = (static_cast<long long>(plan.n[l]) + (plan.n[l] & 1)) << l;
        need = need > fl ? need : fl;
    }
    if (N > (1 << 24)) return fail(AFD_ERR_UNSUPPORTED, "afd_haar_fingerprint_accum: clip too long");
    plan.buf_floats = static_cast<int>((need + 3) / 4 * 4);
    const long long smem = 4LL * plan.buf_floats;
    if (smem > kMaxSmemPerCta)
        return fail(AFD_ERR_UNSUPPORTED, "afd_haar_fingerprint_accum: tree needs %lld bytes of shared memory, limit %d", smem, kMaxSmemPerCta);
    if (B == 0) return AFD_OK;
    int dev = 0, sms = kNumSmsFallback;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static thread_local bool configured[16] = {false};
    if (dev >= 16 || !configured[dev]) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(haar_fingerprint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));
        if (dev < 16) configured[dev] = true;
    }
    const int per_sm = (smem + 1024) * 2 <= 228 * 1024 ? 2 : 1;
    long long grid = static_cast<long long>(sms) * per_sm;
    if (grid > B) grid = B;
    haar_fingerprint_kernel<<<static_cast<unsigned>(grid), kHaarThreads, static_cast<size_t>(smem), s>>>(
        x, static_cast<long long>(x_row_stride), static_cast<long long>(B), sums, plan);
    AFD_CUDA_TRY(cudaGetLastError());
    if (count) {
        add_count_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(count), static_cast<long long>(B) * plan.n[level]);
        AFD_CUDA_TRY(cudaGetLastError());
    }
    return AFD_OK;
}
