// Fused wavelet-packet analysis tree + feature epilogue for sm_100a.
//
// Replaces the ptwt.WaveletPacket / per-node loop / stack / log epilogue of the reference
// (src/audiofakedetect/wavelet_math.py:182-218).  Per node the reference does
//     x~ = reflect_pad(x, F-2 left, F-2 (+1 if len odd) right);  y[k] = sum_m h[m] * x~[2k+1-m]
// for the low-pass h = dec_lo and the high-pass g = dec_hi, recursively to `level`, then orders the
// leaves by Gray code, stacks them P-innermost and applies log(|c|^power + 1e-12).
//
// Design (see DESIGN.md "wpt_tree_kernel"):
//   * One frame is handled by TWO CTAs: CTA h (0/1) owns the sub-tree under the level-1 node 'a'/'d'.
//     Each CTA streams the frame through shared memory once, applying only its own level-1 filter, so no
//     FMA is duplicated; the second read of the frame is an L2 hit.  A half-tree needs ~106 KB of shared
//     memory for the headline configs (level 8, N = 22050), so two CTAs are resident per SM and one CTA's
//     load / store phases overlap the other's FMA phases.
//   * All intermediate levels live in two ping-pong shared-memory regions; only the input frame and the
//     final features touch HBM.
//   * Work item = R consecutive output pairs of one node: the thread loads the 2R+F-2 input window with
//     128-bit LDS (conflict-free for R/2 odd), runs 2*R*F FFMAs whose tap operands come straight from the
//     kernel-parameter constant bank, and writes the R low / R high outputs with 64-bit STS.
//   * Items whose window crosses a node end take a reflect-indexed scalar path; interior and edge items are
//     enumerated separately so warps do not diverge between the two paths.
//   * Epilogue: leaves are read node-major from shared memory, permuted to frequency order
//     (natural index = p ^ (p >> 1)), transformed by log(|c|^power + offset) and stored as full 128-byte
//     rows of out[b][c][t][p].
#include "afd_common.cuh"

namespace afd {

constexpr int kMaxLevel = 12;
constexpr int kThreads = 256;

template <int F>
struct Taps {
    float lo[F];
    float hi[F];
};

struct WptPlan {
    int N;                      // samples per frame
    int L;                      // tree depth
    int n[kMaxLevel + 1];       // n[l] = node length at level l (n[0] = N)
    int stride[kMaxLevel + 1];  // node stride (floats) of level l inside its region
    int base[kMaxLevel + 1];    // offset of node 0 of level l inside its region
    int region_floats[2];       // region 0 holds odd levels, region 1 even levels + level-1 staging
    int stage_chunks;           // number of level-1 staging chunks
    int stage_kc;               // level-1 outputs per chunk (multiple of R)
};

struct Epilogue {
    float power;
    float log_offset;
    int log_scale;
    int sign_channel;
    int order;
    int square;  // power == 2
};

// ------------------------------------------------------------------------------------------------
// FIR core: R output pairs from a register window.  w[j] = x~[2*k0 + 2 - F + j].
// output r uses x~[2(k0+r)+1-m] = w[2r + F-1-m].
// ------------------------------------------------------------------------------------------------
template <int F, int R, bool LO, bool HI, int WLEN>
__device__ __forceinline__ void fir(const float (&w)[WLEN], const Taps<F>& taps, float (&lo)[R], float (&hi)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int m = 0; m < F; ++m) {
            const float xv = w[2 * r + F - 1 - m];
            if (LO) a = fmaf(taps.lo[m], xv, a);
            if (HI) d = fmaf(taps.hi[m], xv, d);
        }
        lo[r] = a;
        hi[r] = d;
    }
}

template <int F, int R>
struct Win {
    static constexpr int W = 2 * R + F - 2;   // window length
    static constexpr int NV = (W + 3) / 4;    // float4 loads
    static constexpr int WLEN = 4 * NV;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * (n - 1) - i : i;
    return min(max(i, 0), n - 1);  // only reached by windows of outputs that are never stored
}

// One tree level: `parents` nodes of length n_in in `in` -> 2*parents nodes of length n_out in `out`.
template <int F, int R>
__device__ __forceinline__ void tree_level(const float* __restrict__ in, float* __restrict__ out, int parents,
                                           int n_in, int n_out, int in_stride, int out_stride,
                                           const Taps<F>& taps) {
    using WN = Win<F, R>;
    const int tid = threadIdx.x;
    const int C = (n_out + R - 1) / R;                 // chunks per node
    const int CL = ((F - 2) / 2 + R - 1) / R;          // chunks whose window starts left of the node
    int CIe = n_in / (2 * R);                          // first chunk whose window ends right of the node
    CIe = min(max(CIe, CL), C);
    const int NI = CIe - CL;                           // interior chunks per node
    const int NE = C - NI;                             // edge chunks per node

    // ---- interior items: aligned 128-bit window loads, all R outputs valid
    for (int it = tid; it < parents * NI; it += kThreads) {
        const int node = it / NI;
        const int k0 = (CL + it - node * NI) * R;
        const float4* src = reinterpret_cast<const float4*>(in + node * in_stride + (2 * k0 + 2 - F));
        float w[WN::WLEN];
#pragma unroll
        for (int v = 0; v < WN::NV; ++v) {
            const float4 q = src[v];
            w[4 * v + 0] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
        }
        float lo[R], hi[R];
        fir<F, R, true, true>(w, taps, lo, hi);
        float2* dlo = reinterpret_cast<float2*>(out + (2 * node) * out_stride + k0);
        float2* dhi = reinterpret_cast<float2*>(out + (2 * node + 1) * out_stride + k0);
#pragma unroll
        for (int r = 0; r < R / 2; ++r) {
            dlo[r] = make_float2(lo[2 * r], lo[2 * r + 1]);
            dhi[r] = make_float2(hi[2 * r], hi[2 * r + 1]);
        }
    }
    // ---- edge items: reflect-indexed scalar loads, guarded stores.  Threads are visited in reverse so the
    // warps left idle by the last interior pass pick these up first.
    for (int it = kThreads - 1 - tid; it < parents * NE; it += kThreads) {
        const int node = it / NE;
        const int e = it - node * NE;
        const int k0 = (e < CL ? e : CIe + (e - CL)) * R;
        const float* nb = in + node * in_stride;
        float w[WN::WLEN];
#pragma unroll
        for (int j = 0; j < WN::W; ++j) w[j] = nb[reflect_idx(2 * k0 + 2 - F + j, n_in)];
        float lo[R], hi[R];
        fir<F, R, true, true>(w, taps, lo, hi);
        float* dlo = out + (2 * node) * out_stride + k0;
        float* dhi = dlo + out_stride;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (k0 + r < n_out) { dlo[r] = lo[r]; dhi[r] = hi[r]; }
        }
    }
}

// Level 1 for one staged chunk: outputs [kb, ke) of the CTA's own filter.
// Sample s of the frame sits at stage[s - s0]  (s0 already folds the alignment offset).
template <int F, int R, bool HIGH>
__device__ __forceinline__ void level1_chunk(const float* __restrict__ stage, int s0, int s_end, int N,
                                             float* __restrict__ out, int kb, int ke, int n_out,
                                             const Taps<F>& taps) {
    using WN = Win<F, R>;
    const int tid = threadIdx.x;
    const int C = (ke - kb + R - 1) / R;
    int CL = 0;
    if (2 * kb + 2 - F < 0) CL = ((F - 2) / 2 - kb + R - 1) / R;  // windows reaching left of sample 0
    int CIe = (s_end - 2 * kb) / (2 * R);                           // window end must stay inside [.., s_end)
    CL = min(CL, C);
    CIe = min(max(CIe, CL), C);
    const int NI = CIe - CL;
    const int NE = C - NI;

    for (int it = tid; it < NI; it += kThreads) {
        const int k0 = kb + (CL + it) * R;
        const float4* src = reinterpret_cast<const float4*>(stage + (2 * k0 + 2 - F - s0));
        float w[WN::WLEN];
#pragma unroll
        for (int v = 0; v < WN::NV; ++v) {
            const float4 q = src[v];
            w[4 * v + 0] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
        }
        float lo[R], hi[R];
        fir<F, R, !HIGH, HIGH>(w, taps, lo, hi);
        float2* dst = reinterpret_cast<float2*>(out + k0);
#pragma unroll
        for (int r = 0; r < R / 2; ++r)
            dst[r] = HIGH ? make_float2(hi[2 * r], hi[2 * r + 1]) : make_float2(lo[2 * r], lo[2 * r + 1]);
    }
    for (int it = kThreads - 1 - tid; it < NE; it += kThreads) {
        const int k0 = kb + (it < CL ? it : CIe + (it - CL)) * R;
        float w[WN::WLEN];
#pragma unroll
        for (int j = 0; j < WN::W; ++j) w[j] = stage[reflect_idx(2 * k0 + 2 - F + j, N) - s0];
        float lo[R], hi[R];
        fir<F, R, !HIGH, HIGH>(w, taps, lo, hi);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (k0 + r < ke && k0 + r < n_out) out[k0 + r] = HIGH ? hi[r] : lo[r];
    }
}

template <int F, int R>
__global__ void __launch_bounds__(kThreads, 2)
wpt_tree_kernel(const float* __restrict__ x, long long x_row_stride, float* __restrict__ out,
                const __grid_constant__ WptPlan plan, const __grid_constant__ Taps<F> taps,
                const __grid_constant__ Epilogue ep) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int b = blockIdx.x >> 1;
    const int half = blockIdx.x & 1;
    const int L = plan.L;
    const int N = plan.N;
    // region 0 (odd levels) starts at smem, region 1 (even levels / level-1 staging) right behind it
    const int r1 = plan.region_floats[0];

    // ------------------------------------------------------------------ level 1 (frame -> own node)
    {
        const float* xg = x + static_cast<long long>(b) * x_row_stride;
        float* stage = smem + r1;
        float* l1 = smem + plan.base[1];
        const int n1 = plan.n[1];
        const bool vec2 = ((reinterpret_cast<uintptr_t>(xg) & 7) == 0);
        for (int c = 0; c < plan.stage_chunks; ++c) {
            const int kb = c * plan.stage_kc;
            const int ke = min(n1, kb + plan.stage_kc);
            const int s_lo = max(0, 2 * kb + 2 - F) & ~1;     // first staged sample (even)
            const int s_end = min(N, 2 * ke);                  // one past the last staged sample
            const int c_off = (F - 2 + s_lo) & 3;              // makes interior windows 16-byte aligned
            const int count = s_end - s_lo;
            if (vec2) {
                const int pairs = count >> 1;
                for (int i = tid; i < pairs; i += kThreads)
                    cp_async_8(stage + c_off + 2 * i, xg + s_lo + 2 * i);
                if ((count & 1) && tid == 0) cp_async_4(stage + c_off + count - 1, xg + s_lo + count - 1);
            } else {
                for (int i = tid; i < count; i += kThreads) cp_async_4(stage + c_off + i, xg + s_lo + i);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            if (half)
                level1_chunk<F, R, true>(stage, s_lo - c_off, s_end, N, l1, kb, ke, n1, taps);
            else
                level1_chunk<F, R, false>(stage, s_lo - c_off, s_end, N, l1, kb, ke, n1, taps);
            __syncthreads();
        }
    }
    // ------------------------------------------------------------------ levels 2..L inside shared memory
    for (int l = 2; l <= L; ++l) {
        const float* in = smem + ((l - 1) & 1 ? 0 : r1) + plan.base[l - 1];
        float* o = smem + (l & 1 ? 0 : r1) + plan.base[l];
        tree_level<F, R>(in, o, 1 << (l - 2), plan.n[l - 1], plan.n[l], plan.stride[l - 1], plan.stride[l], taps);
        __syncthreads();
    }
    // ------------------------------------------------------------------ epilogue: leaves -> out[b][c][t][p]
    {
        const float* leaves = smem + (L & 1 ? 0 : r1) + plan.base[L];
        const int Ts = plan.stride[L];
        const int T = plan.n[L];
        const int Ph = 1 << (L - 1);        // packets owned by this CTA
        const int P = Ph << 1;
        const int C = (ep.log_scale && ep.sign_channel) ? 2 : 1;
        float* ob = out + static_cast<long long>(b) * C * T * P + half * Ph;
        const bool square = ep.square != 0;
        const int total = T * Ph;
        for (int e = tid; e < total; e += kThreads) {
            const int t = e >> (L - 1);
            const int pl = e & (Ph - 1);
            const int p = half * Ph + pl;
            const int nat = (ep.order == AFD_ORDER_FREQ ? (p ^ (p >> 1)) : p) & (Ph - 1);
            const float c = leaves[nat * Ts + t];
            float* dst = ob + static_cast<long long>(t) * P + pl;
            if (ep.log_scale) {
                st_cs(dst, log_power(c, ep.power, ep.log_offset, square));
                if (C == 2) st_cs(dst + static_cast<long long>(T) * P, c < 0.f ? -1.f : 1.f);
            } else {
                st_cs(dst, c);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Builds the shared-memory plan.  Returns 0 or a negative AFD_ERR code.
static int make_plan(int64_t N, int F, int L, int R, WptPlan* p) {
    p->N = static_cast<int>(N);
    p->L = L;
    p->n[0] = static_cast<int>(N);
    for (int l = 1; l <= L; ++l) p->n[l] = (p->n[l - 1] + F - 1) / 2;
    for (int l = 0; l < L; ++l)
        if (p->n[l] < F - 1 + (p->n[l] & 1) || p->n[l] < 2)
            return fail(AFD_ERR_REFLECT_PAD,
                        "node length %d at level %d is not longer than the reflect padding of a %d-tap filter",
                        p->n[l], l, F);
    int need[2] = {0, 0};
    for (int l = 1; l <= L; ++l) {
        const int nodes = 1 << (l - 1);
        if (l == L) {
            // leaves: never read by vector loads; stride == 2 (mod 4) keeps 64-bit stores aligned and the
            // transposing epilogue reads at most 2-way bank conflicted
            int s = round_up(p->n[l], 2);
            if ((s & 3) == 0) s += 2;
            p->stride[l] = s;
            p->base[l] = 0;
        } else {
            p->stride[l] = round_up(p->n[l] + 4, 4);   // +4: slack for the 128-bit window over-read
            p->base[l] = (F - 2) & 3;                  // puts x~[-(F-2)] of every node on a 16-byte boundary
        }
        const int fl = p->base[l] + nodes * p->stride[l] + 4;
        need[l & 1 ? 0 : 1] = need[l & 1 ? 0 : 1] > fl ? need[l & 1 ? 0 : 1] : fl;
    }
    // level-1 staging lives in region 1.  Give it at least what the even levels need anyway; grow it (fewer
    // chunks) only while the whole CTA stays within half an SM.
    const int half_sm_floats = (kMaxSmemPerCta / 2 - 1024) / 4;
    int stage_cap = need[1];
    const int whole = round_up(static_cast<int>(N) + 16, 4);
    if (need[0] + whole <= half_sm_floats) stage_cap = stage_cap > whole ? stage_cap : whole;
    const int min_cap = 4 * R + 2 * F + 16;
    if (stage_cap < min_cap) stage_cap = min_cap;
    // chunk: kc outputs need 2*kc + F - 2 samples (+8 alignment / over-read slack)
    int kc_max = (stage_cap - F - 8) / 2;
    kc_max = kc_max / R * R;
    if (kc_max < R) return fail(AFD_ERR_UNSUPPORTED, "staging buffer too small");
    int chunks = (p->n[1] + kc_max - 1) / kc_max;
    int kc = round_up((p->n[1] + chunks - 1) / chunks, R);
    if (kc > kc_max) kc = kc_max;
    chunks = (p->n[1] + kc - 1) / kc;
    // the last chunk must contain the samples its reflected right edge reads: s_lo(last) <= N - F - 2R - 2
    while (chunks > 1 && 2 * ((chunks - 1) * kc) + 2 - F > p->n[0] - F - 2 * R - 2) {
        // rebalance: shrink kc so the last chunk is not tiny
        kc -= R;
        if (kc < R) return fail(AFD_ERR_UNSUPPORTED, "cannot chunk level-1 staging");
        chunks = (p->n[1] + kc - 1) / kc;
    }
    p->stage_chunks = chunks;
    p->stage_kc = kc;
    const int stage_need = 2 * kc + F + 8;
    need[1] = need[1] > stage_need ? need[1] : stage_need;
    p->region_floats[0] = round_up(need[0] + 4, 4);
    p->region_floats[1] = round_up(need[1] + 4, 4);
    const long long bytes = 4LL * (p->region_floats[0] + p->region_floats[1]);
    if (bytes > kMaxSmemPerCta)
        return fail(AFD_ERR_UNSUPPORTED,
                    "wavelet-packet tree (N=%lld, F=%d, level=%d) needs %lld bytes of shared memory per CTA, limit %d",
                    static_cast<long long>(N), F, L, bytes, kMaxSmemPerCta);
    return AFD_OK;
}

template <int F, int R>
static int launch(const float* x, int64_t B, int64_t x_row_stride, float* out, const WptPlan& plan,
                  const float* dec_lo, const Epilogue& ep, cudaStream_t stream) {
    Taps<F> taps;
    for (int k = 0; k < F; ++k) {
        taps.lo[k] = dec_lo[k];
        taps.hi[k] = ((k & 1) ? 1.f : -1.f) * dec_lo[F - 1 - k];   // dec_hi[k] = (-1)^(k+1) dec_lo[F-1-k]
    }
    const int smem = 4 * (plan.region_floats[0] + plan.region_floats[1]);
    auto kern = wpt_tree_kernel<F, R>;
    static thread_local int configured_smem[16] = {0};  // per device
    int dev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 16 ? configured_smem[dev] < smem : true) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        if (dev < 16) configured_smem[dev] = kMaxSmemPerCta;
    }
    const long long grid = 2 * B;
    if (grid > 2147483647LL) return fail(AFD_ERR_UNSUPPORTED, "batch too large for one launch");
    kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(x, static_cast<long long>(x_row_stride), out,
                                                                   plan, taps, ep);
    AFD_CUDA_TRY(cudaGetLastError());
    return AFD_OK;
}

// chunk width per filter length: R/2 odd keeps the 128-bit window loads and 64-bit stores conflict-free;
// larger R amortises the F-2 halo, smaller R bounds the register window (2R+F-2 floats).
constexpr int pick_r(int F) { return F <= 12 ? 14 : (F <= 32 ? 10 : 6); }

template <int F>
static int dispatch_one(const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                        const float* dec_lo, const Epilogue& ep, cudaStream_t stream) {
    constexpr int R = pick_r(F);
    WptPlan plan;
    int rc = make_plan(N, F, L, R, &plan);
    if (rc != AFD_OK) return rc;
    return launch<F, R>(x, B, x_row_stride, out, plan, dec_lo, ep, stream);
}

}  // namespace afd

using namespace afd;

extern "C" int afd_wpt_out_len(int64_t N, int F, int level, int64_t* T_out) {
    if (N < 1 || F < 2 || (F & 1) || level < 0 || !T_out) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_out_len: bad argument");
    int64_t n = N;
    for (int l = 0; l < level; ++l) n = (n + F - 1) / 2;
    *T_out = n;
    return AFD_OK;
}

extern "C" int afd_wpt_forward(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                               const float* dec_lo_host, int F, int level, int order, float power,
                               int log_scale, float log_offset, int sign_channel, float* out,
                               int64_t* T_out, void* stream) {
    if (!x || !out || !dec_lo_host) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: bad B/N/stride");
    if (F < 2 || F > 64 || (F & 1)) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: filter length %d not in {2,4,..,64}", F);
    if (level < 1 || level > kMaxLevel) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: level %d not in 1..%d", level, kMaxLevel);
    if (order != AFD_ORDER_FREQ && order != AFD_ORDER_NATURAL) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: bad order");
    if (N > (1 << 28)) return fail(AFD_ERR_UNSUPPORTED, "afd_wpt_forward: frame too long");
    int64_t T = 0;
    afd_wpt_out_len(N, F, level, &T);
    if (T_out) *T_out = T;
    if (B == 0) return AFD_OK;
    Epilogue ep;
    ep.power = power;
    ep.log_offset = log_offset;
    ep.log_scale = log_scale ? 1 : 0;
    ep.sign_channel = sign_channel ? 1 : 0;
    ep.order = order;
    ep.square = (power == 2.0f);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define AFD_CASE(FF) case FF: return dispatch_one<FF>(x, B, N, x_row_stride, out, level, dec_lo_host, ep, s);
    switch (F) {
        AFD_CASE(2) AFD_CASE(4) AFD_CASE(6) AFD_CASE(8) AFD_CASE(10) AFD_CASE(12) AFD_CASE(14) AFD_CASE(16)
        AFD_CASE(18) AFD_CASE(20) AFD_CASE(22) AFD_CASE(24) AFD_CASE(26) AFD_CASE(28) AFD_CASE(30) AFD_CASE(32)
        AFD_CASE(34) AFD_CASE(36) AFD_CASE(38) AFD_CASE(40) AFD_CASE(42) AFD_CASE(44) AFD_CASE(46) AFD_CASE(48)
        AFD_CASE(50) AFD_CASE(52) AFD_CASE(54) AFD_CASE(56) AFD_CASE(58) AFD_CASE(60) AFD_CASE(62) AFD_CASE(64)
    }
#undef AFD_CASE
    return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: unsupported filter length %d", F);
}
