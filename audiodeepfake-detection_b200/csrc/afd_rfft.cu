// Mean-spectrum ("rFFT") fingerprint for sm_100a.
//
// Replaces scripts/freq_visual/fingerprints.py:51-62 of the reference:
//     freq_clips = np.fft.rfft(clip_array, axis=-1); (mask with use = all bins: identity)
//     masked_time_mean = np.mean(np.fft.irfft(masked_freq), 0)[0]
//     mean_abs_fft = np.abs(np.fft.rfft(masked_time_mean))
// rfft -> irfft over all bins is the identity on an even-length clip and the mean is linear, so the 2 * n_clips
// transforms of the reference collapse to ONE transform of the mean clip.  What is left per clip is a column sum --
// a single coalesced, HBM-bound pass over the clips (88,200 bytes per clip read, nothing written) -- followed by one
// real DFT of the N-sample mean, evaluated directly in double precision (N = 22050 = 2 * 3^2 * 5^2 * 7^2 has no
// power-of-two path and the one-off N^2 / 2 multiply-adds take a few milliseconds).
#include <math.h>

#include "afd_common.cuh"

namespace afd {

constexpr int kSumThreads = 256;
constexpr int kSumCols = 4;                 // columns per thread, kSumThreads apart (coalesced 128-byte warp loads)
constexpr int kSumTile = kSumThreads * kSumCols;

// sums[n] += sum over the slab's clips of x[b][n].  grid = (column tiles, slabs); clips b = slab, slab + S, ...
__global__ void __launch_bounds__(kSumThreads)
clip_sum_kernel(const float* __restrict__ x, long long x_row_stride, long long B, int N, double* __restrict__ sums) {
    const int c0 = blockIdx.x * kSumTile + threadIdx.x;
    float acc[kSumCols];
#pragma unroll
    for (int j = 0; j < kSumCols; ++j) acc[j] = 0.f;
    const long long S = gridDim.y;
    long long b = blockIdx.y;
    // four clips in flight per thread: 16 independent 4-byte loads
    for (; b + 3 * S < B; b += 4 * S) {
        float v[4][kSumCols];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* row = x + (b + u * S) * x_row_stride;
#pragma unroll
            for (int j = 0; j < kSumCols; ++j) {
                const int c = c0 + j * kSumThreads;
                v[u][j] = c < N ? __ldcs(row + c) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < kSumCols; ++j) acc[j] += (v[0][j] + v[1][j]) + (v[2][j] + v[3][j]);
    }
    for (; b < B; b += S) {
        const float* row = x + b * x_row_stride;
#pragma unroll
        for (int j = 0; j < kSumCols; ++j) {
            const int c = c0 + j * kSumThreads;
            if (c < N) acc[j] += __ldcs(row + c);
        }
    }
#pragma unroll
    for (int j = 0; j < kSumCols; ++j) {
        const int c = c0 + j * kSumThreads;
        if (c < N) atomicAdd(sums + c, static_cast<double>(acc[j]));
    }
}

__global__ void count_add_kernel(long long* count, long long add) { *count += add; }

// tw[j] = exp(-2 pi i j / N)
__global__ void twiddle_kernel(double2* __restrict__ tw, int N) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double s, c;
    sincospi(2.0 * static_cast<double>(j) / static_cast<double>(N), &s, &c);
    tw[j] = make_double2(c, -s);
}

constexpr int kDftThreads = 128;
constexpr int kDftTile = 1024;

// mag[k] = | sum_n scale * x[n] * exp(-2 pi i k n / N) |,  k = 0 .. N/2; the phase index (k n) mod N is exact.
__global__ void __launch_bounds__(kDftThreads)
rdft_magnitude_kernel(const double* __restrict__ x, int N, double scale, const double2* __restrict__ tw,
                      double* __restrict__ mag) {
    __shared__ double xs[kDftTile];
    const int k = blockIdx.x * kDftThreads + threadIdx.x;
    const int bins = N / 2 + 1;
    const int kk = k < bins ? k : 0;
    double re = 0.0, im = 0.0;
    long long idx = 0;                                   // (kk * n) mod N for the current n
    for (int n0 = 0; n0 < N; n0 += kDftTile) {
        const int len = min(kDftTile, N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += kDftThreads) xs[i] = x[n0 + i];
        __syncthreads();
        for (int i = 0; i < len; ++i) {
            const double2 w = tw[idx];
            re = fma(xs[i], w.x, re);
            im = fma(xs[i], w.y, im);
            idx += kk;
            if (idx >= N) idx -= N;
        }
    }
    if (k < bins) mag[k] = scale * sqrt(re * re + im * im);
}

}  // namespace afd

using namespace afd;

extern "C" int afd_clip_sum_accum(const float* x, int64_t B, int64_t N, int64_t x_row_stride, double* sums,
                                  int64_t* count, void* stream) {
    if ((!x || !sums) && B != 0) return fail(AFD_ERR_INVALID_ARG, "afd_clip_sum_accum: null pointer");
    if (B < 0 || N < 1 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_clip_sum_accum: bad B/N/stride");
    if (N > (1LL << 30)) return fail(AFD_ERR_UNSUPPORTED, "afd_clip_sum_accum: clip too long");
    if (B == 0) return AFD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = kNumSmsFallback;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = static_cast<int>((N + kSumTile - 1) / kSumTile);
    // enough slabs for ~8 resident CTAs per SM, never more than one slab per clip, at most 65535 (gridDim.y)
    long long slabs = (8LL * sms + tiles - 1) / tiles;
    if (slabs > B) slabs = B;
    if (slabs > 65535) slabs = 65535;
    if (slabs < 1) slabs = 1;
    clip_sum_kernel<<<dim3(tiles, static_cast<unsigned>(slabs)), kSumThreads, 0, s>>>(
        x, static_cast<long long>(x_row_stride), static_cast<long long>(B), static_cast<int>(N), sums);
    AFD_CUDA_TRY(cudaGetLastError());
    if (count) {
        count_add_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(count), static_cast<long long>(B));
        AFD_CUDA_TRY(cudaGetLastError());
    }
    return AFD_OK;
}

extern "C" int afd_rdft_magnitude(const double* x, int64_t N, double scale, double* mag, void* stream) {
    if (!x || !mag) return fail(AFD_ERR_INVALID_ARG, "afd_rdft_magnitude: null pointer");
    if (N < 1 || N > (1 << 24)) return fail(AFD_ERR_INVALID_ARG, "afd_rdft_magnitude: N out of range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double2* tw = nullptr;
    AFD_CUDA_TRY(cudaMallocAsync(&tw, sizeof(double2) * N, s));
    const int n = static_cast<int>(N);
    twiddle_kernel<<<(n + 255) / 256, 256, 0, s>>>(tw, n);
    const int bins = n / 2 + 1;
    rdft_magnitude_kernel<<<(bins + kDftThreads - 1) / kDftThreads, kDftThreads, 0, s>>>(x, n, scale, tw, mag);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tw, s);
    if (e != cudaSuccess) return cuda_fail(e, "rdft_magnitude_kernel launch");
    return AFD_OK;
}
