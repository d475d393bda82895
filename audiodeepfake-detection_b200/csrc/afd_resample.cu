// Device-side sample-rate conversion for the frame cutter: the reference resamples every loaded window with
// torchaudio.functional.resample(audio, sample_rate, resample_rate) on the host (src/audiofakedetect/data_loader.py:341-344)
// before the transform sees it.  torchaudio's algorithm (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) is a
// polyphase FIR: with o = orig / gcd, m = new / gcd, base = min(o, m) * rolloff, width = ceil(6 * o / base),
//     y[f * m + p] = sum_{j < K} k[p][j] * xz[f * o + j - width],   K = 2 * width + o,   xz = x zero-extended,
//     k[p][j] = sinc(pi t) * cos^2(pi t / 12) * base / o,   t = clamp((-p / m + (j - width) / o) * base, -6, 6),
// for f * m + p < ceil(m * n / o).  The tap table is built once per (orig, new) on the host (float32, torchaudio's order) and kept on
// the device (transposed, [K][m], so that consecutive outputs read consecutive taps); the kernel stages the input span of a
// tile of outputs in shared memory and every thread accumulates its outputs from it.
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "afd_common.cuh"

namespace afd {

namespace {

struct ResampleTable {
    float* taps = nullptr;   // device [K][m]
    int o = 0, m = 0, width = 0, K = 0;
};

std::mutex g_rs_mutex;
std::map<std::tuple<int, int, int>, ResampleTable> g_rs_tables;   // (device, orig, new)

long long gcd_ll(long long a, long long b) { return b == 0 ? a : gcd_ll(b, a % b); }

int get_table(int dev, int orig, int fresh, ResampleTable* out) {
    std::lock_guard<std::mutex> lock(g_rs_mutex);
    auto key = std::make_tuple(dev, orig, fresh);
    auto it = g_rs_tables.find(key);
    if (it != g_rs_tables.end()) { *out = it->second; return AFD_OK; }
    const int g = static_cast<int>(gcd_ll(orig, fresh));
    ResampleTable t;
    t.o = orig / g;
    t.m = fresh / g;
    const double lpw = 6.0, rolloff = 0.99;
    const double base = (t.o < t.m ? t.o : t.m) * rolloff;
    t.width = static_cast<int>(ceil(lpw * t.o / base));
    t.K = 2 * t.width + t.o;
    if (static_cast<long long>(t.K) * t.m > (1LL << 26)) return fail(AFD_ERR_UNSUPPORTED, "afd_resample: tap table too large (%d x %d)", t.K, t.m);
    std::vector<float> host(static_cast<size_t>(t.K) * t.m);
    // torchaudio evaluates the taps in the waveform's dtype (float32), operation by operation; the same order here keeps
    // the table within an ulp of its (a float64 evaluation differs from it by up to 2e-5 of the signal for 640:441)
    const float pi = static_cast<float>(3.14159265358979323846);
    const float basef = static_cast<float>(base), scalef = static_cast<float>(base / t.o), lpwf = static_cast<float>(lpw);
    for (int p = 0; p < t.m; ++p)
        for (int j = 0; j < t.K; ++j) {
            const float idx = static_cast<float>(j - t.width) / static_cast<float>(t.o);
            float tt = static_cast<float>(-p) / static_cast<float>(t.m) + idx;
            tt = tt * basef;
            tt = tt < -lpwf ? -lpwf : (tt > lpwf ? lpwf : tt);
            const float c = cosf(tt * pi / lpwf / 2.0f);
            const float window = c * c;
            const float a = tt * pi;
            const float sinc = a == 0.0f ? 1.0f : sinf(a) / a;
            host[static_cast<size_t>(j) * t.m + p] = sinc * (window * scalef);
        }
    cudaError_t e = cudaMalloc(&t.taps, host.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(t.taps, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { if (t.taps) cudaFree(t.taps); return cuda_fail(e, "resample tap table"); }
    g_rs_tables[key] = t;
    *out = t;
    return AFD_OK;
}

constexpr int kRsThreads = 256;
constexpr int kRsPerThread = 4;
constexpr int kRsTile = kRsThreads * kRsPerThread;

__global__ void __launch_bounds__(kRsThreads)
resample_kernel(const float* __restrict__ x, long long x_row_stride, long long n_in, float* __restrict__ out,
                long long out_row_stride, long long n_out, const float* __restrict__ taps, int o, int m, int width, int K,
                int span_max) {
    extern __shared__ __align__(16) float xs[];
    const long long o0 = static_cast<long long>(blockIdx.x) * kRsTile;
    const long long o1 = min(o0 + kRsTile, n_out);                 // outputs [o0, o1)
    const float* xr = x + blockIdx.y * x_row_stride;
    float* yr = out + blockIdx.y * out_row_stride;
    const long long f_first = o0 / m, f_last = (o1 - 1) / m;
    const long long s0 = f_first * o - width;                      // input index staged at xs[0]
    const int span = static_cast<int>((f_last - f_first) * o + K);
    for (int i = threadIdx.x; i < span && i < span_max; i += kRsThreads) {
        const long long s = s0 + i;
        xs[i] = (s >= 0 && s < n_in) ? __ldg(xr + s) : 0.f;
    }
    __syncthreads();
    float acc[kRsPerThread];
    int base[kRsPerThread], phase[kRsPerThread];
#pragma unroll
    for (int u = 0; u < kRsPerThread; ++u) {
        const long long oi = o0 + threadIdx.x + u * kRsThreads;
        const long long f = oi / m;
        phase[u] = static_cast<int>(oi - f * m);
        base[u] = oi < o1 ? static_cast<int>((f - f_first) * o) : 0;
        acc[u] = 0.f;
    }
    for (int j = 0; j < K; ++j) {
        const float* tj = taps + static_cast<long long>(j) * m;
#pragma unroll
        for (int u = 0; u < kRsPerThread; ++u) acc[u] = fmaf(__ldg(tj + phase[u]), xs[base[u] + j], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < kRsPerThread; ++u) {
        const long long oi = o0 + threadIdx.x + u * kRsThreads;
        if (oi < o1) yr[oi] = acc[u];
    }
}

}  // namespace

}  // namespace afd

using namespace afd;

extern "C" int afd_resample_out_len(int64_t n_in, int orig_freq, int new_freq, int64_t* n_out) {
    if (n_in < 0 || orig_freq < 1 || new_freq < 1 || !n_out) return fail(AFD_ERR_INVALID_ARG, "afd_resample_out_len: bad argument");
    const long long g = gcd_ll(orig_freq, new_freq);
    const long long o = orig_freq / g, m = new_freq / g;
    *n_out = (m * n_in + o - 1) / o;                               // ceil(new * length / orig)
    return AFD_OK;
}

extern "C" int afd_resample(const float* x, int64_t B, int64_t n_in, int64_t x_row_stride, int orig_freq, int new_freq,
                            float* out, int64_t out_row_stride, void* stream) {
    if ((!x || !out) && B != 0) return fail(AFD_ERR_INVALID_ARG, "afd_resample: null pointer");
    if (B < 0 || n_in < 1 || x_row_stride < n_in || orig_freq < 1 || new_freq < 1)
        return fail(AFD_ERR_INVALID_ARG, "afd_resample: bad B/n/stride/rates");
    if (B > 65535) return fail(AFD_ERR_UNSUPPORTED, "afd_resample: more than 65535 rows per call");
    int64_t n_out = 0;
    afd_resample_out_len(n_in, orig_freq, new_freq, &n_out);
    if (out_row_stride < n_out) return fail(AFD_ERR_INVALID_ARG, "afd_resample: out_row_stride %lld < %lld output samples", (long long)out_row_stride, (long long)n_out);
    if (B == 0 || n_out == 0) return AFD_OK;
    int dev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    ResampleTable t;
    int rc = get_table(dev, orig_freq, new_freq, &t);
    if (rc != AFD_OK) return rc;
    // input span of one tile of outputs: (frames touched - 1) * o + K
    const long long frames = (kRsTile + t.m - 1) / t.m + 1;
    const long long span = frames * t.o + t.K;
    if (span * 4 > 200 * 1024) return fail(AFD_ERR_UNSUPPORTED, "afd_resample: ratio %d:%d needs %lld bytes of shared memory per tile", t.o, t.m, span * 4);
    const size_t smem = static_cast<size_t>(span) * 4;
    static thread_local bool configured[16] = {false};
    if (smem > 48 * 1024 && (dev >= 16 || !configured[dev])) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev < 16) configured[dev] = true;
    }
    dim3 grid(static_cast<unsigned>((n_out + kRsTile - 1) / kRsTile), static_cast<unsigned>(B));
    resample_kernel<<<grid, kRsThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<long long>(x_row_stride), static_cast<long long>(n_in), out, static_cast<long long>(out_row_stride),
        static_cast<long long>(n_out), t.taps, t.o, t.m, t.width, t.K, static_cast<int>(span));
    AFD_CUDA_TRY(cudaGetLastError());
    return AFD_OK;
}
