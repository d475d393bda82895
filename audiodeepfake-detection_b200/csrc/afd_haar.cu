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

__global__ void add_count_kernel(long long* count, long long v) { *count += v; }

}  // namespace afd

using namespace afd;

extern "C" int afd_haar_fingerprint_accum(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int level,
                                          double* sums, int64_t* count, void* stream) {
    if (!sums || (!x && B != 0)) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_accum: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_accum: bad B/N/stride");
    if (level < 1 || level > kHaarMaxLevel)
        return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_accum: level %d not in 1..%d", level, kHaarMaxLevel);
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
    cudaStream_t s = static_cast<cudaStream_t>(stream);
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
