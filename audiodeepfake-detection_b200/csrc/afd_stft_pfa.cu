// STFT power spectrogram for n_fft = 511 = 7 * 73 on sm_100a: prime-factor (Good-Thomas) real DFT whose
// 73-point stage runs on the tensor cores.
//
// Replaces torchaudio.transforms.Spectrogram(511, hop, power) -> torch.stft(center=True, pad_mode="reflect",
// window=hann_window(511) [periodic], onesided=True) -> abs().pow(power) and the optional log(spec + 1e-12)
// of the reference (wavelet_math.py:47,63-66) for the reference's only STFT size (n_fft = 2*256-1).
//
// Algorithm.  With n = (73 n1 + 7 n2) mod 511 and k = (k mod 7, k mod 73) = (j, k2) the 511-point DFT of the
// windowed frame xw factors WITHOUT twiddles into 7-point DFTs over n1 and 73-point DFTs over n2:
//     X[k] = sum_n2 W73^(n2 k2) * ( sum_n1 W7^(n1 j) xw[n1, n2] ).
//   (1) fold:    p[n1, m] = xw[n1, m] + xw[n1, 73-m],  q[n1, m] = xw[n1, m] - xw[n1, 73-m]   (m = 1..36; p[., 0] = xw[., 0])
//   (2) DFT-7:   over n1, on the real rows p[., m] and q[., m]  ->  seven real sequences per frame
//                (u0, u1, v1, u2, v2, u3, v3): real / imaginary parts of the j = 0..3 outputs (j = 4..6 are conjugates)
//   (3) DFT-73:  of every sequence as two small real GEMMs against cos / -sin matrices (K = 37 / 36 padded to 40,
//                N = 37 padded to 40):  Re = P * C,  Im = Q * S.   These run as mma.sync m16n8k8 TF32 tensor-core
//                tiles with the 3xTF32 error-compensated split (hi*hi + hi*lo + lo*hi), i.e. fp32-level accuracy.
//   (4) combine: Z_j[k2] = U + iV, Z_j[73-k2] = conj(U) + i conj(V);  bin k = CRT(j, k2') (or 511 - k),
//                |Z|^2 (-> power, log) -> out[b][0][t][bin].
// Per STFT frame this is ~1.0 k tensor-core MACs-rows plus ~6 k scalar operations instead of the ~67 k scalar
// operations of the Bluestein kernel (afd_stft.cu), which stays as the path for every other n_fft.
//
// Mapping.  A work unit = 16 consecutive STFT frames of one signal = the 16 rows of an MMA tile; the 7 sequences are
// 7 row tiles, so step (4) pairs accumulators that sit in the SAME thread (same fragment slot of the u_j and v_j
// tiles).  One unit is processed by a GROUP of 4 warps (128 threads, named barriers); 3 groups per CTA, one
// persistent CTA per SM:
//     stage   : the unit's 15*hop + 511 samples (reflect padding materialised) -> shared memory, 16-byte cp.async;
//               the NEXT unit's samples are prefetched while the current unit is in the tensor-core phase
//     pre     : thread = (m, row subset): window * samples, fold, two 7-point real DFTs -> A tile in shared memory
//     gemm    : warp j owns the sequences (u_j, v_j): ldmatrix A, split hi/lo in registers, B fragments (pre-split,
//               fragment-ordered) from shared memory, 3 MMAs per tile product
//     epilogue: combine, |.|^2, log -> 16 x 256 tile in shared memory (aliases the A tile) -> 128-bit coalesced stores.
#include <math.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "afd_common.cuh"

namespace afd {

constexpr int kPfaN = 511;
constexpr int kPfaRows = 16;               // STFT frames per unit
constexpr int kPfaGroups = 3;              // unit groups per CTA
constexpr int kPfaGroupThreads = 128;
constexpr int kPfaThreads = kPfaGroups * kPfaGroupThreads;
constexpr int kPfaAStride = 84;            // floats per A-tile row (80 used): 8 consecutive rows hit 32 distinct banks
constexpr int kPfaAFloats = 7 * kPfaRows * kPfaAStride;     // 9408
constexpr int kPfaRawFloats = 4160;        // >= 15*hop + 511 + 6  (hop <= 242)
constexpr int kPfaOutStride = 260;         // floats per row of the output tile (16-byte aligned, bank-skewed)
constexpr int kPfaBFloats = 2 * 5 * 5 * 32 * 4;             // fragment-ordered cos / -sin tables, hi+lo
constexpr int kPfaWFloats = 2 * 37 * 8;    // window for the (n1, m) and (n1, 73-m) samples, 8-float rows
constexpr int kPfaMaxHop = (kPfaRawFloats - kPfaN - 6) / (kPfaRows - 1);

struct PfaParams {
    int hop, frames, N, pad, units_per_row, vec_ok;
    long long total_units;
    float power, log_offset;
    int log_scale, square;
    // extended epilogue (afd_stft_power_ex)
    int normalize, store;
    float nmean, nrstd;
    double* moments;           // device [2]: sum / sum of squares of the features before normalisation
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kPfaGroupThreads) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 7-point DFT of a real sequence p[0..6]: r[0] = X0, (r[2j-1], r[2j]) = (Re Xj, Im Xj), j = 1..3.
__device__ __forceinline__ void dft7_real(const float (&p)[7], float (&r)[7]) {
    constexpr float c1 = 0.62348980185873353f, c2 = -0.22252093395631440f, c3 = -0.90096886790241913f;
    constexpr float s1 = 0.78183148246802981f, s2 = 0.97492791218182361f, s3 = 0.43388373911755812f;
    const float a1 = p[1] + p[6], a2 = p[2] + p[5], a3 = p[3] + p[4];
    const float b1 = p[1] - p[6], b2 = p[2] - p[5], b3 = p[3] - p[4];
    r[0] = (p[0] + a1) + (a2 + a3);
    r[1] = fmaf(c3, a3, fmaf(c2, a2, fmaf(c1, a1, p[0])));
    r[3] = fmaf(c1, a3, fmaf(c3, a2, fmaf(c2, a1, p[0])));
    r[5] = fmaf(c2, a3, fmaf(c1, a2, fmaf(c3, a1, p[0])));
    r[2] = -fmaf(s3, b3, fmaf(s2, b2, s1 * b1));
    r[4] = -fmaf(-s1, b3, fmaf(-s3, b2, s2 * b1));
    r[6] = -fmaf(s2, b3, fmaf(-s1, b2, s3 * b1));
}

// Stage the samples of unit (b, t0) into `raw`: raw[aoff + i] = x~[t0*hop - pad + i], x~ = reflect extension.
// Returns nothing; the caller waits with cp.async.wait_group.
__device__ __forceinline__ void pfa_stage(const float* __restrict__ xrow, long long g0, float* __restrict__ raw, int t0,
                                          int valid, const PfaParams& p, int gt) {
    const int S0 = t0 * p.hop - p.pad;                       // first sample of the unit (may be negative)
    const int len = (valid - 1) * p.hop + kPfaN;
    const int aoff = static_cast<int>((g0 + S0) & 3);        // element misalignment of x~[S0] (g0 = b*stride, >= 0 after +S0? see below)
    const int nq = (aoff + len + 3) >> 2;
    const int N = p.N;
    for (int q = gt; q < nq; q += kPfaGroupThreads) {
        const int s0 = S0 - aoff + 4 * q;
        float* dst = raw + 4 * q;
        if (p.vec_ok && s0 >= 0 && s0 + 3 < N) {
            cp_async_16(dst, xrow + s0);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int s = s0 + e;
                s = s < 0 ? -s : s;
                s = s >= N ? 2 * (N - 1) - s : s;
                s = max(0, min(s, N - 1));
                cp_async_4(dst + e, xrow + s);
            }
        }
    }
    cp_async_commit();
}

template <bool PAIR>
__device__ __forceinline__ void pfa_gemm(uint32_t a_addr, const float4* __restrict__ btab, int lane,
                                         float (&acc)[2][2][5][4]) {
#pragma unroll
    for (int gm = 0; gm < 2; ++gm) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
            uint32_t au[4], av[4], aul[4], avl[4];
            const uint32_t addr = a_addr + (gm * 40 + ks * 8) * 4;
            ldmatrix_x4(addr, au);
            if (PAIR) ldmatrix_x4(addr + kPfaRows * kPfaAStride * 4, av);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t hi = au[i] & 0xffffe000u;
                aul[i] = __float_as_uint(__uint_as_float(au[i]) - __uint_as_float(hi));
                au[i] = hi;
                if (PAIR) {
                    const uint32_t hv = av[i] & 0xffffe000u;
                    avl[i] = __float_as_uint(__uint_as_float(av[i]) - __uint_as_float(hv));
                    av[i] = hv;
                }
            }
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                const float4 b = btab[((gm * 5 + ks) * 5 + nt) * 32 + lane];      // {b0 hi, b1 hi, b0 lo, b1 lo}
                const uint32_t b0h = __float_as_uint(b.x), b1h = __float_as_uint(b.y);
                const uint32_t b0l = __float_as_uint(b.z), b1l = __float_as_uint(b.w);
                mma_tf32(acc[0][gm][nt], aul, b0h, b1h);
                mma_tf32(acc[0][gm][nt], au, b0l, b1l);
                mma_tf32(acc[0][gm][nt], au, b0h, b1h);
                if (PAIR) {
                    mma_tf32(acc[1][gm][nt], avl, b0h, b1h);
                    mma_tf32(acc[1][gm][nt], av, b0l, b1l);
                    mma_tf32(acc[1][gm][nt], av, b0h, b1h);
                }
            }
        }
    }
}

// MODE 0: power 2 + log (the reference's configuration), 1: power 2, linear, 2: any power / log flag (runtime)
template <int MODE>
__device__ __forceinline__ float pfa_finish(float re, float im, const PfaParams& p) {
    float v = fmaf(re, re, im * im);
    if (MODE == 0) return ln_approx(v + p.log_offset);
    if (MODE == 1) return v;
    if (!p.square) v = powf(sqrtf(v), p.power);
    if (p.log_scale) v = __logf(v + p.log_offset);
    return v;
}

// tables (device, float): [0, kPfaBFloats) B fragments; then kPfaWFloats window values.
template <bool EXT, int MODE>
__global__ void __launch_bounds__(kPfaThreads, 1)
stft_pfa511_kernel(const float* __restrict__ x, long long x_row_stride, float* __restrict__ out,
                   const float* __restrict__ tables, const __grid_constant__ PfaParams p) {
    extern __shared__ __align__(16) float smem[];
    float* const s_b = smem;                                   // kPfaBFloats
    float* const s_w = s_b + kPfaBFloats;                      // kPfaWFloats
    const int tid = threadIdx.x;
    const int grp = tid / kPfaGroupThreads;
    const int gt = tid - grp * kPfaGroupThreads;
    const int wj = gt >> 5;                                    // warp in group = j
    const int lane = tid & 31;
    float* const s_a = s_w + kPfaWFloats + grp * (kPfaAFloats + kPfaRawFloats);
    float* const s_raw = s_a + kPfaAFloats;
    float* const s_out = s_a;                                  // aliases the A tile after the tensor-core phase

    for (int i = tid; i < kPfaBFloats + kPfaWFloats; i += kPfaThreads) smem[i] = __ldg(tables + i);
    for (int i = gt; i < kPfaAFloats + kPfaRawFloats; i += kPfaGroupThreads) s_a[i] = 0.f;
    __syncthreads();

    // ---- per-thread constants
    // pre-stage role: thread (m, sub) handles rows sub, sub+3, ...; threads 111..127 zero the padding columns
    const bool pre_active = gt < 111;
    const int pm = gt % 37;
    const int psub = gt / 37;
    int offA[7], offB[7];
#pragma unroll
    for (int n1 = 0; n1 < 7; ++n1) {
        int a = 73 * n1 + 7 * pm;
        a = a >= kPfaN ? a - kPfaN : a;
        int b = 73 * n1 - 7 * pm;
        b = b < 0 ? b + kPfaN : b;
        offA[n1] = a;
        offB[n1] = b;
    }
    // epilogue role: warp j, fragment slot (g, tig); output bins of (j, k2) and (j, 73-k2), k2 = 8 nt + 2 tig + c
    const int fg = lane >> 2, ftig = lane & 3;
    uint32_t bins[5];
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int k2 = 8 * nt + 2 * ftig + c;
            int ka = (365 * wj + 147 * k2) % kPfaN;
            ka = ka > 255 ? kPfaN - ka : ka;
            int kb = (365 * wj + 147 * (73 - k2)) % kPfaN;
            kb = kb > 255 ? kPfaN - kb : kb;
            packed |= static_cast<uint32_t>(ka & 255) << (8 * c);
            packed |= static_cast<uint32_t>(kb & 255) << (16 + 8 * c);
        }
        bins[nt] = packed;
    }
    const int s_first = wj == 0 ? 0 : 2 * wj - 1;              // sequence index of u_j (v_j follows)
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lcol = (lane >> 4) * 4;
    const uint32_t a_addr = smem_u32(s_a) + ((s_first * kPfaRows + lrow) * kPfaAStride + lcol) * 4;
    const float4* const btab = reinterpret_cast<const float4*>(s_b);

    float mom_s = 0.f, mom_q = 0.f;
    const float n_rs = (EXT && p.normalize) ? p.nrstd : 1.f, n_dm = (EXT && p.normalize) ? -p.nmean * p.nrstd : 0.f;
    const long long gstride = static_cast<long long>(gridDim.x) * kPfaGroups;
    long long unit = static_cast<long long>(blockIdx.x) * kPfaGroups + grp;
    if (unit < p.total_units) {
        const long long b = unit / p.units_per_row;
        const int t0 = static_cast<int>(unit - b * p.units_per_row) * kPfaRows;
        pfa_stage(x + b * x_row_stride, b * x_row_stride, s_raw, t0, min(kPfaRows, p.frames - t0), p, gt);
    }
    for (; unit < p.total_units; unit += gstride) {
        const long long b = unit / p.units_per_row;
        const int t0 = static_cast<int>(unit - b * p.units_per_row) * kPfaRows;
        const int valid = min(kPfaRows, p.frames - t0);
        const int aoff = static_cast<int>((b * x_row_stride + (t0 * p.hop - p.pad)) & 3);
        cp_async_wait<0>();
        group_bar(grp);                                        // samples landed; A / out tile free again
        // ---------------------------------------------------------------- pre-stage: window, fold, DFT-7
        if (pre_active) {
            float wA[7], wB[7];
#pragma unroll
            for (int n1 = 0; n1 < 7; ++n1) {
                wA[n1] = s_w[pm * 8 + n1];
                wB[n1] = s_w[(37 + pm) * 8 + n1];
            }
            for (int r = psub; r < valid; r += 3) {
                const float* fr = s_raw + aoff + r * p.hop;
                float pp[7], qq[7], P[7], Q[7];
#pragma unroll
                for (int n1 = 0; n1 < 7; ++n1) {
                    const float xa = fr[offA[n1]] * wA[n1];
                    const float xb = fr[offB[n1]] * wB[n1];
                    pp[n1] = xa + xb;
                    qq[n1] = xa - xb;
                }
                dft7_real(pp, P);
                dft7_real(qq, Q);
                float* arow = s_a + r * kPfaAStride + pm;
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    arow[s * kPfaRows * kPfaAStride] = P[s];
                    if (pm != 0) arow[s * kPfaRows * kPfaAStride + 40] = Q[s];
                }
            }
        } else {
            // padding columns 37..40 and 77..79 of all 112 rows (the output tile of the previous unit aliased them)
            for (int i = gt - 111; i < 7 * kPfaRows * 7; i += kPfaGroupThreads - 111) {
                const int row = i / 7, c = i - row * 7;
                s_a[row * kPfaAStride + (c < 4 ? 37 + c : 73 + c)] = 0.f;
            }
        }
        group_bar(grp);                                        // A tile complete; raw buffer free
        {
            const long long nu = unit + gstride;
            if (nu < p.total_units) {
                const long long nb = nu / p.units_per_row;
                const int nt0 = static_cast<int>(nu - nb * p.units_per_row) * kPfaRows;
                pfa_stage(x + nb * x_row_stride, nb * x_row_stride, s_raw, nt0, min(kPfaRows, p.frames - nt0), p, gt);
            }
        }
        // ---------------------------------------------------------------- tensor-core phase: DFT-73 of (u_j, v_j)
        float acc[2][2][5][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2)
#pragma unroll
                for (int nt = 0; nt < 5; ++nt)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][g2][nt][c] = 0.f;
        if (wj == 0) pfa_gemm<false>(a_addr, btab, lane, acc);
        else pfa_gemm<true>(a_addr, btab, lane, acc);
        group_bar(grp);                                        // every warp is done reading the A tile
        // ---------------------------------------------------------------- combine, power, log -> output tile
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k2 = 8 * nt + 2 * ftig + (c & 1);
                const int row = fg + 8 * (c >> 1);
                if (k2 <= 36) {
                    const float ure = acc[0][0][nt][c], uim = acc[0][1][nt][c];
                    const float vre = acc[1][0][nt][c], vim = acc[1][1][nt][c];      // zero for j = 0
                    const uint32_t bp = bins[nt] >> (8 * (c & 1));
                    float* orow = s_out + row * kPfaOutStride;
                    orow[bp & 255u] = pfa_finish<MODE>(ure - vim, uim + vre, p);
                    if (wj != 0 && k2 != 0) orow[(bp >> 16) & 255u] = pfa_finish<MODE>(ure + vim, vre - uim, p);
                }
            }
        }
        group_bar(grp);
        // ---------------------------------------------------------------- coalesced 128-bit stores of the valid rows
        {
            float* og = out + (b * p.frames + t0) * 256LL;
            for (int i = gt; i < valid * 64; i += kPfaGroupThreads) {
                const int row = i >> 6, c4 = (i & 63) * 4;
                float4 v = *reinterpret_cast<const float4*>(s_out + row * kPfaOutStride + c4);
                if (EXT && p.moments) {
                    mom_s += (v.x + v.y) + (v.z + v.w);
                    mom_q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, mom_q))));
                }
                if (EXT && p.normalize) {
                    v.x = fmaf(v.x, n_rs, n_dm); v.y = fmaf(v.y, n_rs, n_dm);
                    v.z = fmaf(v.z, n_rs, n_dm); v.w = fmaf(v.w, n_rs, n_dm);
                }
                if (!EXT || p.store) st_cs4(reinterpret_cast<float4*>(og + row * 256 + c4), v);
            }
        }
    }
    cp_async_wait<0>();
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
static std::mutex g_pfa_mutex;
static std::map<int, float*> g_pfa_tables;

static float tf32_trunc(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u &= 0xffffe000u;
    memcpy(&f, &u, 4);
    return f;
}

static int pfa_tables(int dev, float** out) {
    std::lock_guard<std::mutex> lock(g_pfa_mutex);
    auto it = g_pfa_tables.find(dev);
    if (it != g_pfa_tables.end()) { *out = it->second; return AFD_OK; }
    std::vector<float> h(kPfaBFloats + kPfaWFloats, 0.f);
    auto bval = [](int gm, int k, int n) -> double {
        if (k > 36 || n > 36) return 0.0;
        const int ph = (k * n) % 73;                                     // exact phase reduction
        const double ang = 2.0 * M_PI * double(ph) / 73.0;
        if (gm == 0) return cos(ang);
        return k == 0 ? 0.0 : -sin(ang);
    };
    for (int gm = 0; gm < 2; ++gm)
        for (int ks = 0; ks < 5; ++ks)
            for (int nt = 0; nt < 5; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tig = lane & 3;
                    const int n = 8 * nt + g;
                    float* e = &h[((((gm * 5 + ks) * 5 + nt) * 32) + lane) * 4];
                    for (int i = 0; i < 2; ++i) {
                        const double v = bval(gm, 8 * ks + tig + 4 * i, n);
                        const float hi = tf32_trunc(static_cast<float>(v));
                        e[i] = hi;
                        e[2 + i] = static_cast<float>(v - static_cast<double>(hi));
                    }
                }
    auto hann = [](int n) { return 0.5 - 0.5 * cos(2.0 * M_PI * double(n) / double(kPfaN)); };   // periodic Hann
    for (int m = 0; m < 37; ++m)
        for (int n1 = 0; n1 < 7; ++n1) {
            const int a = (73 * n1 + 7 * m) % kPfaN;
            const int b = ((73 * n1 - 7 * m) % kPfaN + kPfaN) % kPfaN;
            h[kPfaBFloats + m * 8 + n1] = static_cast<float>(hann(a));
            h[kPfaBFloats + (37 + m) * 8 + n1] = m == 0 ? 0.f : static_cast<float>(hann(b));
        }
    float* d = nullptr;
    AFD_CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(float)));
    AFD_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    g_pfa_tables[dev] = d;
    *out = d;
    return AFD_OK;
}

bool stft_pfa511_supported(const float* x, int64_t N, int n_fft, int hop, const float* out) {
    return n_fft == kPfaN && hop >= 1 && hop <= kPfaMaxHop && N > kPfaN / 2 &&
           (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0;
}

int stft_pfa511_launch(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int hop, float power, int log_scale,
                       float log_offset, const StftExtras& ex, float* out, cudaStream_t stream) {
    int dev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    float* tables = nullptr;
    int rc = pfa_tables(dev, &tables);
    if (rc != AFD_OK) return rc;
    PfaParams p;
    p.hop = hop; p.N = static_cast<int>(N); p.pad = kPfaN / 2;
    p.frames = static_cast<int>(1 + (N + 2 * (kPfaN / 2) - kPfaN) / hop);
    p.units_per_row = (p.frames + kPfaRows - 1) / kPfaRows;
    p.total_units = B * static_cast<long long>(p.units_per_row);
    p.vec_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0 ? 1 : 0;
    p.power = power; p.log_offset = log_offset; p.log_scale = log_scale ? 1 : 0; p.square = (power == 2.0f);
    p.normalize = ex.normalize; p.nmean = ex.nmean; p.nrstd = ex.nrstd; p.moments = ex.moments; p.store = out != nullptr;
    const int smem = static_cast<int>(sizeof(float)) *
                     (kPfaBFloats + kPfaWFloats + kPfaGroups * (kPfaAFloats + kPfaRawFloats));
    const bool ext = ex.normalize || ex.moments || !out;
    const int mode = p.square ? (p.log_scale ? 0 : 1) : 2;
    using Kern = void (*)(const float*, long long, float*, const float*, const PfaParams);
    static const Kern kerns[2][3] = {
        {stft_pfa511_kernel<false, 0>, stft_pfa511_kernel<false, 1>, stft_pfa511_kernel<false, 2>},
        {stft_pfa511_kernel<true, 0>, stft_pfa511_kernel<true, 1>, stft_pfa511_kernel<true, 2>}};
    Kern kern = kerns[ext][mode];
    static thread_local bool configured[2][3][16] = {};
    if (dev >= 16 || !configured[ext][mode][dev]) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev < 16) configured[ext][mode][dev] = true;
    }
    int sms = kNumSmsFallback;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (p.total_units + kPfaGroups - 1) / kPfaGroups;
    if (blocks > sms) blocks = sms;
    kern<<<static_cast<unsigned>(blocks), kPfaThreads, smem, stream>>>(x, static_cast<long long>(x_row_stride), out, tables, p);
    AFD_CUDA_TRY(cudaGetLastError());
    return AFD_OK;
}

}  // namespace afd
