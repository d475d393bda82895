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
constexpr int kFlushEvery = 32;     // clips between flushes of the fp32 register accumulators into the fp64 sums

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
