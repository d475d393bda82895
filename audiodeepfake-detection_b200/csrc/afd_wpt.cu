// Fused wavelet-packet analysis tree + feature epilogue for sm_100a.
//
// Replaces the ptwt.WaveletPacket / per-node loop / stack / log epilogue of the reference
// (src/audiofakedetect/wavelet_math.py:182-218).  Per node the reference does
//     x~ = reflect_pad(x, F-2 left, F-2 (+1 if len odd) right);  y[k] = sum_m h[m] * x~[2k+1-m]
// for the low-pass h = dec_lo and the high-pass g = dec_hi, recursively to `level`, then orders the
// leaves by Gray code, stacks them P-innermost and applies log(|c|^power + 1e-12).
//
// Design (DESIGN.md "wpt_tree_kernel"):
//   * One frame is handled by TWO persistent CTAs: CTA h (0/1) owns the sub-tree under the level-1 node
//     'a'/'d' and applies only its own level-1 filter, so no FMA is duplicated; the second read of the frame
//     is an L2 hit.  A half-tree needs ~100 KB of shared memory for the headline configs (level 8, N = 22050),
//     so two CTAs are resident per SM and one CTA's load / store phases overlap the other's FMA phases.
//   * Intermediate levels live in two ping-pong shared-memory regions (odd levels in A, even levels and the
//     frame staging buffers in B).  Every node is stored WITH its reflect padding materialised (F-2 mirrored
//     samples left, F-2 (+1) right), written by the threads that produce the mirrored coefficients.  Every
//     work item of the next level is therefore a plain aligned window: no index reflection on the read side.
//   * Work item = R consecutive output pairs of one node: 128-bit LDS of the 2R+F-2 window (conflict-free for
//     R/2 odd), 2*R*F FFMAs whose tap operands come from the constant bank / uniform registers, 64-bit STS.
//   * The LAST level is never stored: lanes map to consecutive parent nodes, each thread keeps its 2*RL leaf
//     coefficients in registers, applies log(|c|^power + offset), and writes out[b][c][t][2q..2q+1] so that a
//     warp covers 256 contiguous bytes of a feature row per store (frequency order: the children of natural
//     node m sit at positions 2*igray(m) + {parity(m), 1-parity(m)}).
//   * The frame is staged through two cp.async buffers (chunk j+1 in flight while chunk j is filtered) and the
//     first chunk of the CTA's NEXT frame is prefetched while the last level runs.
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
    int stride[kMaxLevel + 1];  // padded node stride (floats) of stored levels 1..L-1 (and level 1 when L == 1)
    int region_b;               // float offset of region B (region A starts at 0)
    int buf_floats;             // floats per staging buffer (two of them at the start of region B)
    int kc;                     // level-1 outputs per staging chunk (multiple of R)
    int nch;                    // staging chunks per frame
    int groups;                 // the last stored level is produced / consumed in `groups` slices (1, 2 or 4)
    int smem_floats;            // total dynamic shared memory in floats
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
// FIR core: R outputs (of one or both filters) from a register window.  w[j] = x~[2*k0 + 2 - F + j];
// output r uses x~[2(k0+r)+1-m] = w[2r + F-1-m].
// ------------------------------------------------------------------------------------------------
template <int F, int R, int WLEN>
__device__ __forceinline__ void fir2(const float (&w)[WLEN], const Taps<F>& taps, float (&lo)[R], float (&hi)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int m = 0; m < F; ++m) {
            const float xv = w[2 * r + F - 1 - m];
            a = fmaf(taps.lo[m], xv, a);
            d = fmaf(taps.hi[m], xv, d);
        }
        lo[r] = a;
        hi[r] = d;
    }
}

template <int F, int R, int WLEN>
__device__ __forceinline__ void fir1(const float (&w)[WLEN], const float (&t)[F], float (&y)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float a = 0.f;
#pragma unroll
        for (int m = 0; m < F; ++m) a = fmaf(t[m], w[2 * r + F - 1 - m], a);
        y[r] = a;
    }
}

template <int F, int R>
struct Win {
    static constexpr int W = 2 * R + F - 2;   // window length
    static constexpr int NV = (W + 3) / 4;    // float4 loads
    static constexpr int WLEN = 4 * NV;
};

template <int NV>
__device__ __forceinline__ void load_window(const float* __restrict__ p, float (&w)[4 * NV]) {
    const float4* src = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float4 q = src[v];
        w[4 * v + 0] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
    }
}

// Chunk classification of a node with n_out coefficients, stored with padl / padr mirrored samples:
// chunk c (outputs cR .. cR+R-1) is "interior" when it holds no mirrored coefficient and is fully valid.
struct Split {
    int C, CL, CIe, NI, NE;
    unsigned magicNI, magicNE;     // floor(i / d) = umulhi(i, magic) for i, d < 2^16
};
__host__ __device__ inline unsigned magic_of(int d) { return d > 0 ? static_cast<unsigned>(0xFFFFFFFFu / static_cast<unsigned>(d)) + 1u : 0u; }

__device__ __forceinline__ Split make_split(int n_out, int R, int padl) {
    Split s;
    s.C = (n_out + R - 1) / R;
    const int padr = padl + (n_out & 1);
    int cl = padl == 0 ? 0 : padl / R + 1;
    int cie = (n_out - 1 - padr) / R;            // chunks c < cie end before the first right-mirrored coefficient
    if (n_out - 1 - padr < 0) cie = 0;
    cl = min(cl, s.C);
    cie = min(max(cie, cl), s.C);
    s.CL = cl; s.CIe = cie; s.NI = cie - cl; s.NE = s.C - s.NI;
    s.magicNI = magic_of(s.NI);
    s.magicNE = magic_of(s.NE);
    return s;
}
__device__ __forceinline__ int fast_div(int i, int d, unsigned magic) { return d == 1 ? i : static_cast<int>(__umulhi(static_cast<unsigned>(i), magic)); }

// Guarded store of R coefficients of a child node plus their mirror images into the node's padding.
// `node` points at the first padding sample; coefficient k lives at node[padl + k].
template <int R>
__device__ __forceinline__ void edge_store(float* __restrict__ node, const float (&v)[R], int k0, int n_out, int padl) {
    const int padr = padl + (n_out & 1);
    float* pos = node + padl;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int k = k0 + r;
        if (k < n_out) {
            pos[k] = v[r];
            if (k >= 1 && k <= padl) pos[-k] = v[r];
            const int mr = n_out - 1 - k;
            if (mr >= 1 && mr <= padr) pos[n_out - 1 + mr] = v[r];
        }
    }
}

template <int R>
__device__ __forceinline__ void vec_store(float* __restrict__ dst, const float (&v)[R]) {
    float2* d = reinterpret_cast<float2*>(dst);
#pragma unroll
    for (int r = 0; r < R / 2; ++r) d[r] = make_float2(v[2 * r], v[2 * r + 1]);
}

// One stored tree level: `parents` padded nodes in `in` -> 2*parents padded nodes in `out`.
template <int F, int R>
__device__ __forceinline__ void mid_level(const float* __restrict__ in, float* __restrict__ out, int parents,
                                          int n_out, int in_stride, int out_stride, const Taps<F>& taps) {
    using WN = Win<F, R>;
    constexpr int padl = F - 2;
    const Split sp = make_split(n_out, R, padl);
    const int n_int = parents * sp.NI;
    const int total = parents * sp.C;
    for (int it = threadIdx.x; it < total; it += kThreads) {
        const bool edge = it >= n_int;
        int node, c;
        if (!edge) {
            node = fast_div(it, sp.NI, sp.magicNI);
            c = sp.CL + it - node * sp.NI;
        } else {
            const int e = it - n_int;
            node = fast_div(e, sp.NE, sp.magicNE);
            const int ee = e - node * sp.NE;
            c = ee < sp.CL ? ee : sp.CIe + (ee - sp.CL);
        }
        const int k0 = c * R;
        float w[WN::WLEN];
        load_window<WN::NV>(in + node * in_stride + 2 * k0, w);
        float lo[R], hi[R];
        fir2<F, R>(w, taps, lo, hi);
        float* d0 = out + (2 * node) * out_stride;
        if (!edge) {
            vec_store<R>(d0 + padl + k0, lo);
            vec_store<R>(d0 + out_stride + padl + k0, hi);
        } else {
            edge_store<R>(d0, lo, k0, n_out, padl);
            edge_store<R>(d0 + out_stride, hi, k0, n_out, padl);
        }
    }
}

// Level 1 for one staged chunk: outputs [kb, ke) of the CTA's own filter -> padded level-1 node.
// Sample 2*kb + 2 - F of the (reflect-extended) frame sits at buf[0].
template <int F, int R>
__device__ __forceinline__ void level1_chunk(const float* __restrict__ buf, float* __restrict__ node, int kb, int ke,
                                             int n_out, const Split& sp, const float (&t)[F]) {
    using WN = Win<F, R>;
    constexpr int padl = F - 2;
    const int items = (ke - kb + R - 1) / R;
    const int c0 = kb / R;
    for (int i = threadIdx.x; i < items; i += kThreads) {
        const int c = c0 + i;
        const int k0 = c * R;
        float w[WN::WLEN];
        load_window<WN::NV>(buf + 2 * i * R, w);
        float y[R];
        fir1<F, R>(w, t, y);
        if (c >= sp.CL && c < sp.CIe) vec_store<R>(node + padl + k0, y);
        else edge_store<R>(node, y, k0, n_out, padl);
    }
}
// Chunks are cut at multiples of R, so only the last chunk (ke == n_out) holds a partial item.

__device__ __forceinline__ unsigned igray(unsigned x) {
    x ^= x >> 1; x ^= x >> 2; x ^= x >> 4; x ^= x >> 8;
    return x;
}

// Last level: `parents` padded nodes (level L-1) -> features in global memory.  Lanes map to parents.
template <int F, int RL>
__device__ __forceinline__ void last_level(const float* __restrict__ in, int parents, int lg_parents, int in_stride,
                                           int T, int parent_base, float* __restrict__ out_b, int P,
                                           const Taps<F>& taps, const Epilogue& ep) {
    using WN = Win<F, RL>;
    const int chunks = (T + RL - 1) / RL;
    const int total = parents * chunks;
    const bool square = ep.square != 0;
    const bool two = ep.log_scale && ep.sign_channel;
    const long long ch1 = static_cast<long long>(T) * P;
    for (int it = threadIdx.x; it < total; it += kThreads) {
        const int m = it & (parents - 1);
        const int c = it >> lg_parents;
        const int k0 = c * RL;
        float w[WN::WLEN];
        load_window<WN::NV>(in + m * in_stride + 2 * k0, w);
        float lo[RL], hi[RL];
        fir2<F, RL>(w, taps, lo, hi);
        const unsigned pf = static_cast<unsigned>(parent_base + m);      // natural index at level L-1
        unsigned q = pf;
        bool swap = false;
        if (ep.order == AFD_ORDER_FREQ) {
            q = igray(pf);
            swap = (q & 1u) != 0;                                      // parity(pf) = lsb of igray(pf)
        }
        float* o = out_b + static_cast<long long>(k0) * P + 2 * q;
#pragma unroll
        for (int r = 0; r < RL; ++r) {
            if (k0 + r < T) {
                const float c0 = swap ? hi[r] : lo[r];
                const float c1 = swap ? lo[r] : hi[r];
                float2 v = make_float2(c0, c1);
                if (ep.log_scale)
                    v = make_float2(log_power(c0, ep.power, ep.log_offset, square),
                                    log_power(c1, ep.power, ep.log_offset, square));
                __stcs(reinterpret_cast<float2*>(o + static_cast<long long>(r) * P), v);
                if (two)
                    __stcs(reinterpret_cast<float2*>(o + ch1 + static_cast<long long>(r) * P),
                           make_float2(c0 < 0.f ? -1.f : 1.f, c1 < 0.f ? -1.f : 1.f));
            }
        }
    }
}

// Stage chunk j of the frame (with the frame's own reflect padding) into `buf` with cp.async.
template <int F>
__device__ __forceinline__ void issue_chunk(const float* __restrict__ xg, float* __restrict__ buf, int j,
                                            const WptPlan& plan) {
    const int N = plan.N;
    const int kb = j * plan.kc;
    const int ke = min(plan.n[1], kb + plan.kc);
    const int s_start = 2 * kb + 2 - F;           // sample index stored at buf[0] (even)
    const int s_end = 2 * ke;                     // one past the last sample any stored output needs
    const int r_lo = max(s_start, 0);
    const int r_hi = min(s_end, N);               // real samples [r_lo, r_hi)
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(xg) & 7) == 0) {     // r_lo is even: 8-byte copies
        const int pairs = (r_hi - r_lo) >> 1;
        for (int i = tid; i < pairs; i += kThreads)
            cp_async_8(buf + (r_lo - s_start) + 2 * i, xg + r_lo + 2 * i);
        if (((r_hi - r_lo) & 1) && tid == 0) cp_async_4(buf + (r_hi - 1 - s_start), xg + r_hi - 1);
    } else {
        for (int i = r_lo + tid; i < r_hi; i += kThreads) cp_async_4(buf + (i - s_start), xg + i);
    }
    // reflect padding of the frame itself: x~[-i] = x[i], x~[N-1+i] = x[N-1-i]
    for (int s = s_start + tid; s < 0; s += kThreads) cp_async_4(buf + (s - s_start), xg - s);
    for (int s = max(N, s_start) + tid; s < s_end; s += kThreads) cp_async_4(buf + (s - s_start), xg + (2 * (N - 1) - s));
    cp_async_commit();
}

template <int F, int R, int RL>
__global__ void __launch_bounds__(kThreads, 2)
wpt_tree_kernel(const float* __restrict__ x, long long x_row_stride, long long B, float* __restrict__ out,
                const __grid_constant__ WptPlan plan, const __grid_constant__ Taps<F> taps,
                const __grid_constant__ Epilogue ep) {
    extern __shared__ __align__(16) float smem[];
    constexpr int padl = F - 2;
    const int L = plan.L;
    const int half = blockIdx.x & 1;                      // gridDim.x is even: constant per CTA
    float* const regA = smem;                              // odd levels
    float* const regB = smem + plan.region_b;              // even levels; staging buffers at its start
    float* const buf0 = regB;
    float* const buf1 = regB + plan.buf_floats;
    const int P = 1 << L;
    const int C = (ep.log_scale && ep.sign_channel) ? 2 : 1;
    const int T = plan.n[L];
    const int n1 = plan.n[1];
    const Split sp1 = make_split(n1, R, padl);
    // the CTA's level-1 filter
    float t1[F];
#pragma unroll
    for (int m = 0; m < F; ++m) t1[m] = half ? taps.hi[m] : taps.lo[m];
    // region B is idle while the last level runs iff the last level reads region A
    const bool can_prefetch = (L == 1) || (((L - 1) & 1) == 1);
    bool prefetched = false;
    bool first = true;

    for (long long wk = blockIdx.x; wk < 2 * B; wk += gridDim.x) {
        const long long b = wk >> 1;
        const float* xg = x + b * x_row_stride;
        if (!prefetched) {
            if (!first) __syncthreads();                   // region B may still be read by the previous last level
            issue_chunk<F>(xg, buf0, 0, plan);
        }
        first = false;
        prefetched = false;
        // ---------------------------------------------------------------- level 1 (frame -> own padded node in A)
        for (int j = 0; j < plan.nch; ++j) {
            cp_async_wait<0>();
            __syncthreads();
            if (j + 1 < plan.nch) issue_chunk<F>(xg, ((j + 1) & 1) ? buf1 : buf0, j + 1, plan);
            const int kb = j * plan.kc;
            const int ke = min(n1, kb + plan.kc);
            level1_chunk<F, R>((j & 1) ? buf1 : buf0, regA, kb, ke, n1, sp1, t1);
        }
        __syncthreads();
        float* out_b = out + b * C * static_cast<long long>(T) * P;
        if (L == 1) {
            // the level-1 node is the output: epilogue straight from shared memory (rare configuration)
            const long long nb = wk + gridDim.x;
            if (nb < 2 * B) { issue_chunk<F>(x + (nb >> 1) * x_row_stride, buf0, 0, plan); prefetched = true; }
            const bool square = ep.square != 0;
            for (int e = threadIdx.x; e < T; e += kThreads) {
                const float c = regA[padl + e];
                float* dst = out_b + static_cast<long long>(e) * P + half;
                if (ep.log_scale) {
                    st_cs(dst, log_power(c, ep.power, ep.log_offset, square));
                    if (C == 2) st_cs(dst + static_cast<long long>(T) * P, c < 0.f ? -1.f : 1.f);
                } else {
                    st_cs(dst, c);
                }
            }
            continue;
        }
        // ---------------------------------------------------------------- stored levels 2 .. L-1
        const int G = plan.groups;
        const int last_full = G > 1 ? L - 2 : L - 1;
        for (int l = 2; l <= last_full; ++l) {
            const float* in = ((l - 1) & 1) ? regA : regB;
            float* o = (l & 1) ? regA : regB;
            mid_level<F, R>(in, o, 1 << (l - 2), plan.n[l], plan.stride[l - 1], plan.stride[l], taps);
            __syncthreads();
        }
        // ---------------------------------------------------------------- last level (+ grouped level L-1)
        const int nodes_lm1 = 1 << (L - 2);                 // level L-1 nodes of this half tree
        const int per_group = nodes_lm1 / G;
        int lg = 0;
        while ((1 << lg) < per_group) ++lg;
        const float* lm1 = ((L - 1) & 1) ? regA : regB;
        for (int g = 0; g < G; ++g) {
            if (G > 1) {
                if (g > 0) __syncthreads();                 // level L-1 slice is being re-used
                const float* in = ((L - 2) & 1) ? regA : regB;
                float* o = ((L - 1) & 1) ? regA : regB;
                const int par = per_group / 2;              // level L-2 parents of this slice
                mid_level<F, R>(in + g * par * plan.stride[L - 2], o, par, plan.n[L - 1], plan.stride[L - 2],
                                plan.stride[L - 1], taps);
                __syncthreads();
            }
            if (g == G - 1 && can_prefetch) {
                const long long nb = wk + gridDim.x;
                if (nb < 2 * B) { issue_chunk<F>(x + (nb >> 1) * x_row_stride, buf0, 0, plan); prefetched = true; }
            }
            last_level<F, RL>(lm1, per_group, lg, plan.stride[L - 1], T, half * nodes_lm1 + g * per_group, out_b, P,
                              taps, ep);
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Builds the shared-memory plan for `ctas_per_sm` resident CTAs.  Returns 0 or a negative AFD_ERR code.
static int make_plan(int64_t N, int F, int L, int R, int RL, int ctas_per_sm, WptPlan* p) {
    p->N = static_cast<int>(N);
    p->L = L;
    p->n[0] = static_cast<int>(N);
    for (int l = 1; l <= L; ++l) p->n[l] = (p->n[l - 1] + F - 1) / 2;
    for (int l = 0; l < L; ++l)
        if (p->n[l] < F - 1 + (p->n[l] & 1) || p->n[l] < 2)
            return fail(AFD_ERR_REFLECT_PAD,
                        "node length %d at level %d is not longer than the reflect padding of a %d-tap filter",
                        p->n[l], l, F);
    const int padl = F - 2;
    const int limit_floats = (ctas_per_sm == 2 ? (228 * 1024 / 2 - 1024) : kMaxSmemPerCta) / 4;
    const int tail = 2 * (R > RL ? R : RL) + F + 16;          // over-read slack behind the last node of a region
    const int stored = L == 1 ? 1 : L - 1;                    // levels kept in shared memory
    for (int l = 1; l <= stored; ++l) {
        int s = round_up(p->n[l] + 2 * padl + (p->n[l] & 1), 4);
        if (l == L - 1 && ((s >> 2) & 1) == 0) s += 4;        // lanes map to nodes in the last level: odd 16-byte stride
        p->stride[l] = s;
    }
    // level-1 staging: balanced chunks of at most kThreads items
    const int n1 = p->n[1];
    p->nch = (n1 + kThreads * R - 1) / (kThreads * R);
    p->kc = round_up((n1 + p->nch - 1) / p->nch, R);
    p->nch = (n1 + p->kc - 1) / p->kc;
    p->buf_floats = round_up(2 * p->kc + F + 8, 4);
    for (int G = 1; G <= 4; G *= 2) {
        if (G > 1 && (L < 4 || (1 << (L - 2)) / G < 2)) break;
        int need[2] = {0, 2 * p->buf_floats};                 // [0] = region A (odd levels), [1] = region B
        for (int l = 1; l <= stored; ++l) {
            int nodes = 1 << (l - 1);
            if (l == L - 1 && G > 1) nodes /= G;
            const int fl = nodes * p->stride[l] + tail;
            int& r = need[(l & 1) ? 0 : 1];
            r = r > fl ? r : fl;
        }
        p->groups = G;
        p->region_b = round_up(need[0], 4);
        p->smem_floats = p->region_b + round_up(need[1], 4);
        if (p->smem_floats <= limit_floats) return AFD_OK;
    }
    return AFD_ERR_UNSUPPORTED;   // caller retries with one CTA per SM or reports
}

template <int F, int R, int RL>
static int launch(const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                  const float* dec_lo, const Epilogue& ep, cudaStream_t stream) {
    WptPlan plan;
    int ctas = 2;
    int rc = make_plan(N, F, L, R, RL, 2, &plan);
    if (rc == AFD_ERR_UNSUPPORTED) {
        ctas = 1;
        rc = make_plan(N, F, L, R, RL, 1, &plan);
        if (rc == AFD_ERR_UNSUPPORTED)
            return fail(AFD_ERR_UNSUPPORTED,
                        "wavelet-packet tree (N=%lld, F=%d, level=%d) needs %lld bytes of shared memory per CTA, limit %d",
                        static_cast<long long>(N), F, L, 4LL * plan.smem_floats, kMaxSmemPerCta);
    }
    if (rc != AFD_OK) return rc;
    Taps<F> taps;
    for (int k = 0; k < F; ++k) {
        taps.lo[k] = dec_lo[k];
        taps.hi[k] = ((k & 1) ? 1.f : -1.f) * dec_lo[F - 1 - k];   // dec_hi[k] = (-1)^(k+1) dec_lo[F-1-k]
    }
    const int smem = 4 * plan.smem_floats;
    auto kern = wpt_tree_kernel<F, R, RL>;
    static thread_local bool configured[16] = {false};  // per device
    int dev = 0, sms = kNumSmsFallback;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16 || !configured[dev]) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        if (dev < 16) configured[dev] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = 2LL * sms * ctas / 2 * 2;             // persistent: every resident slot, even count
    if (ctas == 1) grid = sms / 2 * 2;
    if (grid > 2 * B) grid = 2 * B;
    kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(x, static_cast<long long>(x_row_stride),
                                                                   static_cast<long long>(B), out, plan, taps, ep);
    AFD_CUDA_TRY(cudaGetLastError());
    return AFD_OK;
}

// R: outputs per work item in the stored levels (R/2 odd keeps the 128-bit window loads and 64-bit stores
// conflict-free; larger R amortises the F-2 halo, smaller R bounds the register window of 2R+F-2 floats).
// RL: outputs per work item in the last level (even).
constexpr int pick_r(int F) { return F <= 24 ? 14 : (F <= 40 ? 10 : 6); }
constexpr int pick_rl(int F) { return F <= 12 ? 12 : (F <= 24 ? 14 : (F <= 40 ? 10 : 6)); }

template <int F>
static int dispatch_one(const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                        const float* dec_lo, const Epilogue& ep, cudaStream_t stream) {
    return launch<F, pick_r(F), pick_rl(F)>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream);
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
    if (!dec_lo_host || ((!x || !out) && B != 0)) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: null pointer");
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: bad B/N/stride");
    if (F < 2 || F > 64 || (F & 1)) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: filter length %d not in {2,4,..,64}", F);
    if (level < 1 || level > kMaxLevel) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: level %d not in 1..%d", level, kMaxLevel);
    if (order != AFD_ORDER_FREQ && order != AFD_ORDER_NATURAL) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward: bad order");
    if (N > (1 << 20)) return fail(AFD_ERR_UNSUPPORTED, "afd_wpt_forward: frame too long");
    if (B > (1LL << 30)) return fail(AFD_ERR_UNSUPPORTED, "afd_wpt_forward: batch too large");
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
