// Host-buffer entry points: stream frames through the device in chunks, overlapping the H2D copy of chunk
// i+1, the kernel of chunk i and the D2H copy of chunk i-1 on separate streams.  This is what a CPU-side
// caller of the reference (numpy in / numpy out, e.g. scripts/freq_visual/fingerprints.py) would use.
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>

#include "afd_common.cuh"

namespace afd {

#ifndef AFD_HOST_SLOTS
#define AFD_HOST_SLOTS 4
#endif
constexpr int kSlots = AFD_HOST_SLOTS;

// Per-device pipeline state, created on first use and kept for the life of the process: streams and device
// staging buffers are reused by every host-buffer call (cudaMalloc / cudaFree per call would serialise the
// device and cost more than the copies of a small batch).  One host call at a time per device.
struct Pipeline {
    std::mutex mu;
    cudaStream_t stream[kSlots] = {};
    float* d_in[kSlots] = {};
    float* d_out[kSlots] = {};
    size_t in_cap[kSlots] = {};
    size_t out_cap[kSlots] = {};
};

static std::mutex g_pipe_mutex;
static std::map<int, Pipeline*> g_pipes;

static Pipeline* pipeline_for(int device) {
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    auto it = g_pipes.find(device);
    if (it != g_pipes.end()) return it->second;
    Pipeline* p = new Pipeline();
    g_pipes[device] = p;
    return p;
}

static cudaError_t ensure(float** buf, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return cudaSuccess;
    if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc(buf, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

// launch(d_in, d_out, nb, stream) must enqueue the device transform of nb frames.
static int run_pipelined(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride, float* out_host,
                         int64_t out_row_floats, int device, int64_t chunk,
                         const std::function<int(const float*, float*, int64_t, cudaStream_t)>& launch) {
    int prev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&prev));
    AFD_CUDA_TRY(cudaSetDevice(device));
    if (chunk <= 0) chunk = 512;
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(B, 1));
    Pipeline* pl = pipeline_for(device);
    std::lock_guard<std::mutex> lock(pl->mu);
    int rc = AFD_OK;
    const int nslots = static_cast<int>(std::min<int64_t>(kSlots, (B + chunk - 1) / chunk));
    for (int i = 0; i < nslots && rc == AFD_OK; ++i) {
        cudaError_t e = cudaSuccess;
        if (!pl->stream[i]) e = cudaStreamCreateWithFlags(&pl->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = ensure(&pl->d_in[i], &pl->in_cap[i], sizeof(float) * chunk * N);
        if (e == cudaSuccess && out_row_floats > 0)
            e = ensure(&pl->d_out[i], &pl->out_cap[i], sizeof(float) * chunk * out_row_floats);
        if (e != cudaSuccess) rc = cuda_fail(e, "pipeline slot allocation");
    }
    // Ramp-up: the device-to-host copies (the longer direction for the feature tensors) cannot start before the first
    // chunk's host-to-device copy and kernel are through, so the first chunks are small (chunk / 8, / 4, / 2) and the pipeline
    // fills in ~1/8 of a chunk's copy time.
    int64_t done = 0;
    for (int64_t c = 0; rc == AFD_OK && done < B; ++c) {
        const int i = static_cast<int>(c % nslots);
        cudaStream_t s = pl->stream[i];
        const int64_t ramp = c < 3 ? std::max<int64_t>(chunk >> (3 - c), 1) : chunk;
        const int64_t nb = std::min(ramp, B - done);
        cudaError_t e;
        if (x_row_stride == N)
            e = cudaMemcpyAsync(pl->d_in[i], x_host + done * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, s);
        else
            e = cudaMemcpy2DAsync(pl->d_in[i], sizeof(float) * N, x_host + done * x_row_stride,
                                  sizeof(float) * x_row_stride, sizeof(float) * N, nb, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
        rc = launch(pl->d_in[i], pl->d_out[i], nb, s);
        if (rc != AFD_OK) break;
        if (out_row_floats > 0) {
            e = cudaMemcpyAsync(out_host + done * out_row_floats, pl->d_out[i], sizeof(float) * nb * out_row_floats,
                                cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy"); break; }
        }
        done += nb;
    }
    for (int i = 0; i < nslots; ++i) {
        if (!pl->stream[i]) continue;
        cudaError_t e = cudaStreamSynchronize(pl->stream[i]);
        if (e != cudaSuccess && rc == AFD_OK) rc = cuda_fail(e, "pipeline synchronize");
    }
    cudaSetDevice(prev);
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
    e = cudaMemset(d_sums, 0, P * sizeof(double) + sizeof(long long));
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);   // the pipeline streams are non-blocking: they would not wait for it
    if (e != cudaSuccess) { cudaFree(d_sums); cudaSetDevice(prev); return cuda_fail(e, "cudaMemset(sums)"); }
    // every chunk accumulates into the same device sums: atomics make the slots' kernels commute
    int rc = B == 0 ? AFD_OK
                    : run_pipelined(x_host, B, N, x_row_stride, nullptr, 0, device, chunk_frames,
                                    [&](const float* d_in, float*, int64_t nb, cudaStream_t s) {
                                        return afd_haar_fingerprint_accum(d_in, nb, N, N, level, d_sums,
                                                                          reinterpret_cast<int64_t*>(d_count), s);
                                    });
    if (rc == AFD_OK) {
        e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaMemcpy(sums_host, d_sums, P * sizeof(double), cudaMemcpyDeviceToHost);
        long long cnt = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&cnt, d_count, sizeof(long long), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "D2H copy of sums");
        if (count_host) *count_host = cnt;
    }
    cudaFree(d_sums);
    cudaSetDevice(prev);
    return rc;
}
