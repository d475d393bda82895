// Fused STFT power spectrogram for sm_100a: framing + reflect padding + Hann window + real DFT of arbitrary
// length + |X|^power (+ log) in one kernel.
//
// Replaces torchaudio.transforms.Spectrogram(n_fft, hop_length, power) -> torch.stft(center=True,
// pad_mode="reflect", window=hann_window(n_fft) [periodic], onesided=True, normalized=False) ->
// abs().pow(power) and the optional log(spec + 1e-12)   (reference wavelet_math.py:47,63-66).
//
// The reference's n_fft is 2*num_of_scales-1 = 511 = 7*73, so no radix-2 path exists.  The DFT is evaluated
// with Bluestein's chirp-z identity  jk = (j^2 + k^2 - (k-j)^2)/2 :
//     X[k] = conj(c[k]) * sum_j (x_w[j] conj(c[j])) * c[k-j],     c[m] = exp(i*pi*m^2/n)
// i.e. one circular convolution of length M = 1024 >= 2n-1, done with two 1024-point complex FFTs.
// Two real STFT frames are packed into one complex transform (z = x1 + i*x2) and separated afterwards by
// conjugate symmetry, so one frame costs one 1024-point FFT.
//
// Mapping: ONE WARP owns one frame pair from load to store; there is no block-level barrier in the main
// loop.  A 1024-point FFT is a 32x32 decomposition: every lane runs a 32-point radix-2 DIF FFT entirely in
// registers (compile-time twiddles), the warp transposes through its private 8.25 KB padded shared-memory
// tile, and the lanes run the second 32-point FFT.  The spectrum of the chirp (with the 1/M of the inverse
// transform folded in) is computed on the host in double precision once per (n_fft) and cached per device.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "afd_common.cuh"

namespace afd {

constexpr int kFftM = 1024;
constexpr int kStftWarps = 8;
constexpr int kTileStride = 33;                        // float2 row stride of the 32x32 transpose tile
constexpr int kTileFloat2 = 32 * kTileStride;          // per warp

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// compile-time twiddle W32^k = exp(-2*pi*i*k/32) for the in-register FFT
template <int K>
__device__ __forceinline__ float2 mul_w32(float2 v) {
    if (K == 0) return v;
    if (K == 8) return make_float2(v.y, -v.x);                 // * (-i)
    constexpr float kC[16] = {
        1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
        0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
        -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
        -0.92387953251128674f, -0.98078528040323043f};
    constexpr float kS[16] = {
        0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f,
        0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f,
        0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
        0.38268343236508977f, 0.19509032201612825f};
    // (x + iy)(c - is) = (xc + ys) + i(yc - xs)
    return make_float2(fmaf(v.x, kC[K], v.y * kS[K]), fmaf(v.y, kC[K], -v.x * kS[K]));
}

// One radix-2 DIF stage over v[0..31]: butterflies of span HALF; twiddle step STEP = 16/HALF.
template <int HALF, int J>
struct Dif32Stage {
    __device__ static __forceinline__ void run(float2 (&v)[32]) {
        if constexpr (J < 16) {
            constexpr int grp = J / HALF;
            constexpr int pos = J % HALF;
            constexpr int i0 = grp * 2 * HALF + pos;
            constexpr int i1 = i0 + HALF;
            const float2 a = v[i0], b = v[i1];
            v[i0] = make_float2(a.x + b.x, a.y + b.y);
            v[i1] = mul_w32<pos * (16 / HALF)>(make_float2(a.x - b.x, a.y - b.y));
            Dif32Stage<HALF, J + 1>::run(v);
        }
    }
};

// Forward 32-point FFT in registers.  Output is in bit-reversed order: v[brev5(k)] = X[k].
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) {
    Dif32Stage<16, 0>::run(v);
    Dif32Stage<8, 0>::run(v);
    Dif32Stage<4, 0>::run(v);
    Dif32Stage<2, 0>::run(v);
    Dif32Stage<1, 0>::run(v);
}

__host__ __device__ constexpr int brev5(int k) {
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// Forward 1024-point FFT of the warp's data.
//   input : v[r] = x[lane + 32 r]                         (register resident)
//   output: v[q] = X[lane + 32 q]                         (register resident, natural order)
// tw2d[kr*32 + l] = exp(-2*pi*i*l*kr/1024) (shared memory, conflict-free by lane).  tile: warp-private 32x33 float2.
__device__ __forceinline__ void fft1024_warp(float2 (&v)[32], float2* __restrict__ tile,
                                             const float2* __restrict__ tw2d, int lane) {
    // pass A: FFT over r for column `lane`  ->  Y[lane][kr], then twiddle W1024^(lane*kr), store tile[lane][kr]
    fft32_dif(v);
#pragma unroll
    for (int kr = 0; kr < 32; ++kr) {
        float2 y = v[brev5(kr)];
        if (kr != 0) y = cmul(y, tw2d[kr * 32 + lane]);
        tile[lane * kTileStride + kr] = y;
    }
    __syncwarp();
    // pass B: lane = kr; FFT over l -> X[kr + 32 kl]
#pragma unroll
    for (int l = 0; l < 32; ++l) v[l] = tile[l * kTileStride + lane];
    __syncwarp();
    fft32_dif(v);
    // v[brev5(kl)] = X[lane + 32 kl]  -> put into natural register order through compile-time renaming
    float2 t[32];
#pragma unroll
    for (int kl = 0; kl < 32; ++kl) t[kl] = v[brev5(kl)];
#pragma unroll
    for (int kl = 0; kl < 32; ++kl) v[kl] = t[kl];
}

struct StftParams {
    int n_fft, hop, frames, bins, N, pad;
    int pairs_per_row;      // ceil(frames / 2)
    long long total_pairs;  // B * pairs_per_row
    float power, log_offset;
    int log_scale, square;
    int normalize, store;   // extended epilogue (afd_stft_power_ex)
    float nmean, nrstd;
    double* moments;
};

// Tables (device, per n_fft): [0,n)      a_tab[j] = hann[j] * conj(c[j])        (float2)
//                             [n,2n)     post[k]  = conj(c[k])                  (float2)
//                             [2n,2n+M)  bspec[q] = FFT_M(chirp)[q] / M         (float2)
//                             [2n+M, +M) tw2d[kr*32+l] = exp(-2 pi i l kr / M)  (float2)
template <bool EXT>
__global__ void __launch_bounds__(kStftWarps * 32, 2)
stft_bluestein_kernel(const float* __restrict__ x, long long x_row_stride, float* __restrict__ out,
                      const float2* __restrict__ tables, const __grid_constant__ StftParams p) {
    extern __shared__ __align__(16) float2 smem2[];
    const int n = p.n_fft;
    float2* s_tw = smem2;                         // [M]
    float2* s_bspec = s_tw + kFftM;               // [M]
    float2* s_tiles = s_bspec + kFftM;            // [warps][32*33]
    const float2* g_atab = tables;
    const float2* g_post = tables + n;
    for (int i = threadIdx.x; i < kFftM; i += blockDim.x) {
        s_bspec[i] = tables[2 * n + i];
        s_tw[i] = tables[2 * n + kFftM + i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float2* tile = s_tiles + warp * kTileFloat2;
    const long long warps_total = static_cast<long long>(gridDim.x) * kStftWarps;
    const bool square = p.square != 0;
    float mom_s = 0.f, mom_q = 0.f;

    for (long long pr = static_cast<long long>(blockIdx.x) * kStftWarps + warp; pr < p.total_pairs; pr += warps_total) {
        const long long b = pr / p.pairs_per_row;
        const int f0 = static_cast<int>(pr - b * p.pairs_per_row) * 2;     // first frame of the pair
        const bool has2 = (f0 + 1) < p.frames;
        const float* xr = x + b * x_row_stride;
        const int s0 = f0 * p.hop - p.pad;                                 // sample index of tap 0, frame f0
        // ---- load, window, pre-chirp:  v[r] = (x1 + i x2)[j] * hann[j] * conj(c[j]),  j = lane + 32 r
        float2 v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int j = lane + 32 * r;
            float2 val = make_float2(0.f, 0.f);
            if (j < n) {
                int i1 = s0 + j;
                i1 = i1 < 0 ? -i1 : i1;
                i1 = i1 >= p.N ? 2 * (p.N - 1) - i1 : i1;
                int i2 = s0 + p.hop + j;
                i2 = i2 < 0 ? -i2 : i2;
                i2 = i2 >= p.N ? 2 * (p.N - 1) - i2 : i2;
                const float x1 = __ldg(xr + i1);
                const float x2 = has2 ? __ldg(xr + i2) : 0.f;
                val = cmul(make_float2(x1, x2), __ldg(g_atab + j));
            }
            v[r] = val;
        }
        // ---- A = FFT(a);  A *= Bspec;  y = IFFT(A) via conj(FFT(conj(.)))  (1/M folded into Bspec)
        fft1024_warp(v, tile, s_tw, lane);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const float2 t = cmul(v[q], s_bspec[lane + 32 * q]);
            v[q] = make_float2(t.x, -t.y);
        }
        fft1024_warp(v, tile, s_tw, lane);
        // v[q] = conj(y[lane + 32 q]);  Z[k] = conj(c[k]) * y[k], k < n
        // ---- stash Z in the tile (natural order, k = lane + 32 q -> tile[k + k/32]) to pair k with n-k
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int k = lane + 32 * q;
            if (k < n) {
                const float2 y = make_float2(v[q].x, -v[q].y);
                tile[k + (k >> 5)] = cmul(y, __ldg(g_post + k));
            }
        }
        __syncwarp();
        // ---- separate the two real frames, power, log, store (bins are contiguous in memory)
        float* o1 = out + (b * p.frames + f0) * static_cast<long long>(p.bins);
        for (int k = lane; k < p.bins; k += 32) {
            const float2 zk = tile[k + (k >> 5)];
            const int km = (k == 0) ? 0 : n - k;
            const float2 zm = tile[km + (km >> 5)];
            // X1 = (Z[k] + conj(Z[n-k]))/2 ; X2 = (Z[k] - conj(Z[n-k]))/(2i)
            const float x1r = 0.5f * (zk.x + zm.x), x1i = 0.5f * (zk.y - zm.y);
            const float x2r = 0.5f * (zk.y + zm.y), x2i = -0.5f * (zk.x - zm.x);
            float p1 = fmaf(x1r, x1r, x1i * x1i);
            float p2 = fmaf(x2r, x2r, x2i * x2i);
            if (!square) {
                p1 = powf(sqrtf(p1), p.power);
                p2 = powf(sqrtf(p2), p.power);
            }
            if (p.log_scale) {
                p1 = __logf(p1 + p.log_offset);
                p2 = __logf(p2 + p.log_offset);
            }
            if (EXT && p.moments) {
                mom_s += p1; mom_q = fmaf(p1, p1, mom_q);
                if (has2) { mom_s += p2; mom_q = fmaf(p2, p2, mom_q); }
            }
            if (EXT && p.normalize) {
                p1 = (p1 - p.nmean) * p.nrstd;
                p2 = (p2 - p.nmean) * p.nrstd;
            }
            if (!EXT || p.store) {
                st_cs(o1 + k, p1);
                if (has2) st_cs(o1 + p.bins + k, p2);
            }
        }
        __syncwarp();
    }
    if (EXT && p.moments) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mom_s += __shfl_xor_sync(0xffffffffu, mom_s, o);
            mom_q += __shfl_xor_sync(0xffffffffu, mom_q, o);
        }
        if (lane == 0) {
            atomicAdd(p.moments, static_cast<double>(mom_s));
            atomicAdd(p.moments + 1, static_cast<double>(mom_q));
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
struct TableKey {
    int dev, n_fft;
    bool operator<(const TableKey& o) const { return dev != o.dev ? dev < o.dev : n_fft < o.n_fft; }
};
static std::mutex g_tab_mutex;
static std::map<TableKey, float2*> g_tables;

static void fft_host(std::vector<double>& re, std::vector<double>& im) {  // in-place radix-2, size power of 2
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * M_PI * double(k) / double(len);
                const double wr = cos(ang), wi = sin(ang);
                const double ur = re[i + k], ui = im[i + k];
                const double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                const double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
            }
        }
    }
}

static int get_tables(int dev, int n, float2** out) {
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    auto it = g_tables.find({dev, n});
    if (it != g_tables.end()) { *out = it->second; return AFD_OK; }
    const int M = kFftM;
    std::vector<float2> h(2 * n + 2 * M);
    std::vector<double> cr(n), ci(n);
    for (int m = 0; m < n; ++m) {
        const long long q = (static_cast<long long>(m) * m) % (2LL * n);   // exact phase reduction
        const double ang = M_PI * double(q) / double(n);
        cr[m] = cos(ang); ci[m] = sin(ang);
        const double hann = 0.5 - 0.5 * cos(2.0 * M_PI * double(m) / double(n));  // periodic Hann
        h[m] = make_float2(float(hann * cr[m]), float(-hann * ci[m]));
        h[n + m] = make_float2(float(cr[m]), float(-ci[m]));
    }
    std::vector<double> br(M, 0.0), bi(M, 0.0);
    for (int m = 0; m < n; ++m) {
        br[m] = cr[m]; bi[m] = ci[m];
        if (m) { br[M - m] = cr[m]; bi[M - m] = ci[m]; }
    }
    fft_host(br, bi);
    for (int q = 0; q < M; ++q) h[2 * n + q] = make_float2(float(br[q] / M), float(bi[q] / M));
    for (int kr = 0; kr < 32; ++kr)
        for (int l = 0; l < 32; ++l) {
            const double ang = -2.0 * M_PI * double(l * kr) / double(M);
            h[2 * n + M + kr * 32 + l] = make_float2(float(cos(ang)), float(sin(ang)));
        }
    float2* d = nullptr;
    AFD_CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(float2)));
    AFD_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    g_tables[{dev, n}] = d;
    *out = d;
    return AFD_OK;
}

// afd_stft_pfa.cu: prime-factor / tensor-core path for n_fft = 511
bool stft_pfa511_supported(const float* x, int64_t N, int n_fft, int hop, const float* out);
int stft_pfa511_launch(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int hop, float power, int log_scale,
                       float log_offset, const StftExtras& ex, float* out, cudaStream_t stream);
bool stft_tc511_supported(const float* x, int64_t B, int64_t N, int n_fft, int hop, const float* out);
int stft_tc511_launch(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int hop, float power, int log_scale,
                      float log_offset, const StftExtras& ex, float* out, cudaStream_t stream);

}  // namespace afd

using namespace afd;

extern "C" int afd_stft_out_shape(int64_t N, int n_fft, int hop, int64_t* frames, int64_t* bins) {
    if (N < 1 || n_fft < 2 || hop < 1 || !frames || !bins) return fail(AFD_ERR_INVALID_ARG, "afd_stft_out_shape: bad argument");
    *frames = 1 + (N + 2 * (n_fft / 2) - n_fft) / hop;   // torch.stft(center=True): 1 + (N + 2*pad - n_fft) // hop
    *bins = n_fft / 2 + 1;
    return AFD_OK;
}

static int stft_power_impl(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int n_fft, int hop,
                           float power, int log_scale, float log_offset, const StftExtras& ex, float* out, void* stream) {
    if ((!x || (!out && !ex.moments)) && B != 0) return fail(AFD_ERR_INVALID_ARG, "afd_stft_power: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N || hop < 1) return fail(AFD_ERR_INVALID_ARG, "afd_stft_power: bad B/N/stride/hop");
    if (n_fft < 2) return fail(AFD_ERR_INVALID_ARG, "afd_stft_power: n_fft must be >= 2");
    if (2 * n_fft - 1 > kFftM)
        return fail(AFD_ERR_UNSUPPORTED, "afd_stft_power: n_fft %d needs a chirp-z length above %d (n_fft <= 512 supported)", n_fft, kFftM);
    if (n_fft / 2 >= N)
        return fail(AFD_ERR_REFLECT_PAD, "afd_stft_power: reflect padding %d must be smaller than the signal length %lld", n_fft / 2, (long long)N);
    if (N > (1LL << 30)) return fail(AFD_ERR_UNSUPPORTED, "afd_stft_power: signal too long");
    if (B == 0) return AFD_OK;
    {
        // n_fft = 511 (the reference's 2*num_of_scales-1) takes the prime-factor tensor-core kernels: the tcgen05 / TMEM
        // kernel (afd_stft_tc.cu); the mma.sync kernel (afd_stft_pfa.cu) is kept as the cross-check.  AFD_STFT_IMPL=pfa forces the mma.sync kernel, AFD_STFT_IMPL=bluestein the generic chirp-z
        // kernel (the parity tests cover all three).
        const char* impl = getenv("AFD_STFT_IMPL");
        const bool force_generic = impl && strcmp(impl, "bluestein") == 0;
        const bool force_pfa = impl && strcmp(impl, "pfa") == 0;
        if (!force_generic && !force_pfa && stft_tc511_supported(x, B, N, n_fft, hop, out))
            return stft_tc511_launch(x, B, N, x_row_stride, hop, power, log_scale, log_offset, ex, out,
                                     static_cast<cudaStream_t>(stream));
        if (!force_generic && stft_pfa511_supported(x, N, n_fft, hop, out))
            return stft_pfa511_launch(x, B, N, x_row_stride, hop, power, log_scale, log_offset, ex, out,
                                      static_cast<cudaStream_t>(stream));
    }
    int dev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    float2* tables = nullptr;
    int rc = get_tables(dev, n_fft, &tables);
    if (rc != AFD_OK) return rc;
    StftParams p;
    p.n_fft = n_fft; p.hop = hop; p.N = static_cast<int>(N); p.pad = n_fft / 2;
    p.frames = static_cast<int>(1 + (N + 2 * (n_fft / 2) - n_fft) / hop);
    p.bins = n_fft / 2 + 1;
    p.pairs_per_row = (p.frames + 1) / 2;
    p.total_pairs = B * static_cast<long long>(p.pairs_per_row);
    p.power = power; p.log_offset = log_offset; p.log_scale = log_scale ? 1 : 0; p.square = (power == 2.0f);
    p.normalize = ex.normalize; p.nmean = ex.nmean; p.nrstd = ex.nrstd; p.moments = ex.moments; p.store = out != nullptr;
    const int smem = static_cast<int>(sizeof(float2)) * (2 * kFftM + kStftWarps * kTileFloat2);
    const bool ext = ex.normalize || ex.moments || !out;
    auto kern = ext ? stft_bluestein_kernel<true> : stft_bluestein_kernel<false>;
    static thread_local bool configured[2][16] = {{false}, {false}};
    if (dev >= 16 || !configured[ext][dev]) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev < 16) configured[ext][dev] = true;
    }
    int sms = kNumSmsFallback;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (p.total_pairs + kStftWarps - 1) / kStftWarps;
    const long long max_blocks = 2LL * sms;   // persistent: 2 CTAs (16 warps) per SM, register-bound
    if (blocks > max_blocks) blocks = max_blocks;
    kern<<<static_cast<unsigned>(blocks), kStftWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<long long>(x_row_stride), out, tables, p);
    AFD_CUDA_TRY(cudaGetLastError());
    return AFD_OK;
}

extern "C" int afd_stft_power(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int n_fft, int hop,
                              float power, int log_scale, float log_offset, float* out, void* stream) {
    if (!out && B != 0) return fail(AFD_ERR_INVALID_ARG, "afd_stft_power: null pointer");
    return stft_power_impl(x, B, N, x_row_stride, n_fft, hop, power, log_scale, log_offset, StftExtras{}, out, stream);
}

extern "C" int afd_stft_power_ex(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int n_fft, int hop,
                                 float power, int log_scale, float log_offset, const float* norm_mean_std_host,
                                 double* feat_moments, float* out, void* stream) {
    StftExtras ex{};
    ex.moments = feat_moments;
    if (norm_mean_std_host) {
        const float mean = norm_mean_std_host[0], std = norm_mean_std_host[1];
        if (!(std > 0.f) || !isfinite(mean) || !isfinite(std))
            return fail(AFD_ERR_INVALID_ARG, "afd_stft_power_ex: needs a finite mean and a positive std");
        ex.normalize = 1; ex.nmean = mean; ex.nrstd = 1.0f / std;
    }
    return stft_power_impl(x, B, N, x_row_stride, n_fft, hop, power, log_scale, log_offset, ex, out, stream);
}
