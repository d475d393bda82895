// Host-buffer entry points: stream frames through the device in chunks, overlapping the H2D copy of chunk
// i+1, the kernel of chunk i and the D2H copy of chunk i-1 on separate streams.  This is what a CPU-side
// caller of the reference (numpy in / numpy out, e.g. scripts/freq_visual/fingerprints.py) would use.
#include <algorithm>
#include <functional>

#include "afd_common.cuh"

namespace afd {

constexpr int kSlots = 3;

struct Slot {
    cudaStream_t stream = nullptr;
    float* d_in = nullptr;
    float* d_out = nullptr;
};

// launch(d_in, d_out, nb, stream) must enqueue the device transform of nb frames.
static int run_pipelined(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride, float* out_host,
                         int64_t out_row_floats, int device, int64_t chunk,
                         const std::function<int(const float*, float*, int64_t, cudaStream_t)>& launch) {
    int prev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&prev));
    AFD_CUDA_TRY(cudaSetDevice(device));
    if (chunk <= 0) chunk = 512;
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(B, 1));
    Slot slots[kSlots];
    int rc = AFD_OK;
    auto cleanup = [&]() {
        for (auto& s : slots) {
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        cudaSetDevice(prev);
    };
    const int nslots = static_cast<int>(std::min<int64_t>(kSlots, (B + chunk - 1) / chunk));
    for (int i = 0; i < nslots && rc == AFD_OK; ++i) {
        cudaError_t e = cudaStreamCreateWithFlags(&slots[i].stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&slots[i].d_in, sizeof(float) * chunk * N);
        if (e == cudaSuccess && out_row_floats > 0) e = cudaMalloc(&slots[i].d_out, sizeof(float) * chunk * out_row_floats);
        if (e != cudaSuccess) rc = cuda_fail(e, "pipeline slot allocation");
    }
    int64_t done = 0;
    for (int64_t c = 0; rc == AFD_OK && done < B; ++c) {
        Slot& s = slots[c % nslots];
        const int64_t nb = std::min(chunk, B - done);
        cudaError_t e = cudaMemcpy2DAsync(s.d_in, sizeof(float) * N, x_host + done * x_row_stride,
                                          sizeof(float) * x_row_stride, sizeof(float) * N, nb,
                                          cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
        rc = launch(s.d_in, s.d_out, nb, s.stream);
        if (rc != AFD_OK) break;
        if (out_row_floats > 0) {
            e = cudaMemcpyAsync(out_host + done * out_row_floats, s.d_out, sizeof(float) * nb * out_row_floats,
                                cudaMemcpyDeviceToHost, s.stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy"); break; }
        }
        done += nb;
    }
    for (int i = 0; i < nslots; ++i) {
        if (!slots[i].stream) continue;
        cudaError_t e = cudaStreamSynchronize(slots[i].stream);
        if (e != cudaSuccess && rc == AFD_OK) rc = cuda_fail(e, "pipeline synchronize");
    }
    cleanup();
    return rc;
}

}  // namespace afd

using namespace afd;

extern "C" int afd_wpt_forward_host(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride,
                                    const double* dec_lo_host, int F, int level, int order, float power,
                                    int log_scale, float log_offset, int sign_channel, float* out_host,
                                    int64_t* T_out, int device, int64_t chunk_frames) {
    if (!x_host || !out_host || !dec_lo_host) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward_host: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward_host: bad B/N/stride");
    int64_t T = 0;
    int rc = afd_wpt_out_len(N, F, level, &T);
    if (rc != AFD_OK) return rc;
    if (T_out) *T_out = T;
    if (level < 1 || level > 30) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward_host: bad level");
    const int64_t C = (log_scale && sign_channel) ? 2 : 1;
    const int64_t row = C * T * (int64_t(1) << level);
    if (B == 0) return AFD_OK;
    return run_pipelined(x_host, B, N, x_row_stride, out_host, row, device, chunk_frames,
                         [&](const float* d_in, float* d_out, int64_t nb, cudaStream_t s) {
                             return afd_wpt_forward(d_in, nb, N, N, dec_lo_host, F, level, order, power, log_scale,
                                                    log_offset, sign_channel, d_out, nullptr, s);
                         });
}

extern "C" int afd_stft_power_host(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride, int n_fft,
                                   int hop, float power, int log_scale, float log_offset, float* out_host,
                                   int device, int64_t chunk_frames) {
    if (!x_host || !out_host) return fail(AFD_ERR_INVALID_ARG, "afd_stft_power_host: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_stft_power_host: bad B/N/stride");
    int64_t frames = 0, bins = 0;
    int rc = afd_stft_out_shape(N, n_fft, hop, &frames, &bins);
    if (rc != AFD_OK) return rc;
    if (B == 0) return AFD_OK;
    return run_pipelined(x_host, B, N, x_row_stride, out_host, frames * bins, device, chunk_frames,
                         [&](const float* d_in, float* d_out, int64_t nb, cudaStream_t s) {
                             return afd_stft_power(d_in, nb, N, N, n_fft, hop, power, log_scale, log_offset, d_out, s);
                         });
}

extern "C" int afd_haar_fingerprint_host(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride, int level,
                                         double* sums_host, int64_t* count_host, int device, int64_t chunk_frames) {
    if (!x_host || !sums_host) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_host: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_host: bad B/N/stride");
    if (level < 1 || level > 14) return fail(AFD_ERR_INVALID_ARG, "afd_haar_fingerprint_host: bad level");
    int prev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&prev));
    AFD_CUDA_TRY(cudaSetDevice(device));
    const size_t P = size_t(1) << level;
    double* d_sums = nullptr;
    long long* d_count = nullptr;
    cudaError_t e = cudaMalloc(&d_sums, P * sizeof(double) + sizeof(long long));
    if (e != cudaSuccess) { cudaSetDevice(prev); return cuda_fail(e, "cudaMalloc(sums)"); }
    d_count = reinterpret_cast<long long*>(d_sums + P);
    cudaMemset(d_sums, 0, P * sizeof(double) + sizeof(long long));
    // every chunk accumulates into the same device sums: atomics make the slots' kernels commute
    int rc = B == 0 ? AFD_OK
                    : run_pipelined(x_host, B, N, x_row_stride, nullptr, 0, device, chunk_frames,
                                    [&](const float* d_in, float*, int64_t nb, cudaStream_t s) {
                                        return afd_haar_fingerprint_accum(d_in, nb, N, N, level, d_sums,
                                                                          reinterpret_cast<int64_t*>(d_count), s);
                                    });
    if (rc == AFD_OK) {
        AFD_CUDA_TRY(cudaSetDevice(device));
        e = cudaMemcpy(sums_host, d_sums, P * sizeof(double), cudaMemcpyDeviceToHost);
        long long cnt = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&cnt, d_count, sizeof(long long), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "D2H copy of sums");
        if (count_host) *count_host = cnt;
    }
    cudaFree(d_sums);
    cudaSetDevice(prev);
    return rc;
}
