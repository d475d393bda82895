// Fused wavelet-packet analysis tree + feature epilogue for sm_100a.
//
// Replaces the ptwt.WaveletPacket / per-node loop / stack / log epilogue of the reference
// (src/audiofakedetect/wavelet_math.py:182-218).  Per node the reference does
//     x~ = reflect_pad(x, F-2 left, F-2 (+1 if len odd) right);  y[k] = sum_m h[m] * x~[2k+1-m]
// for the low-pass h = dec_lo and the high-pass g = dec_hi, recursively to `level`, then orders the
// leaves by Gray code, stacks them P-innermost and applies log(|c|^power + 1e-12).
//
// Two kernels share the work-item code below.  wpt_frame_kernel (r2, further down) serves every filter that has a usable
// lattice when the whole tree fits one CTA: one 512-thread CTA per frame, level 1 as a lattice pass over the staged frame,
// two 256-thread groups for the half trees.  wpt_tree_kernel (r1: two CTAs per frame, described next) serves the rest:
// direct-form filters (F > 32, coif5), long frames, very deep trees.
//
// Design of wpt_tree_kernel:
//   * One frame is handled by TWO persistent CTAs: CTA h (0/1) owns the sub-tree under the level-1 node
//     'a'/'d' and applies only its own level-1 filter (direct form), so no FMA is duplicated; the second read of
//     the frame is an L2 hit.  A half-tree needs ~100 KB of shared memory for the headline configs (level 8,
//     N = 22050), so two CTAs are resident per SM and one CTA's load / store phases overlap the other's FMAs.
//   * Levels >= 2 evaluate the filter PAIR as a paraunitary lattice (afd_lattice.cu): J = F/2 plane rotations
//     separated by unit delays, run in place on the register window, F FFMAs per (lo, hi) pair instead of 2F.
//     The rotations are unscaled (u += t v', v = v' - t u); the product of the stage cosines is folded into the
//     stores of the levels where the pending factor leaves [1e-9, 1e9] and into the epilogue of the last level.
//     Filters whose lattice is not trustworthy (residual, reflection, F > 32) use the direct form.
//   * Intermediate levels live in two ping-pong shared-memory regions (odd levels in A, even levels and the
//     frame staging buffers in B).  Every node is stored with room for its reflect padding (F-2 samples left,
//     F-2 (+1) right), so every work item of the next level is a plain aligned window.  The right padding is
//     materialised (by the threads that produce the mirrored coefficients, or by a copy pass for F >= 16); the
//     left padding is not written: its only reader, chunk 0 of the next level, mirrors its own register window
//     (see ReflOk below; nodes with an item size R < F/2 keep the left mirror stores).
//   * Work item = R consecutive output pairs of one node: 128-bit LDS of the 2R+F-2 window (conflict-free for
//     R/2 odd), the FIR / lattice on registers with tap operands from the constant bank, 64-bit STS.  R is picked
//     per level (two instantiations) so that the item count fills whole rounds of the 256 threads.
//   * The LAST level is never stored: lanes map to consecutive parent nodes, each thread keeps its leaf
//     coefficients in registers, applies log(c^2 + offset) and writes out[b][c][t][2q..2q+1] so that a
//     warp covers 256 contiguous bytes of a feature row per store (frequency order: the children of natural
//     node m sit at positions 2*igray(m) + {parity(m), 1-parity(m)}).
//   * The frame is staged through two cp.async buffers (chunk j+1 in flight while chunk j is filtered) and the
//     first chunk of the CTA's NEXT frame is prefetched while the last level runs.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "afd_common.cuh"

namespace afd {

constexpr int kMaxLevel = 12;
#ifndef AFD_WPT_THREADS
#define AFD_WPT_THREADS 256
#endif
constexpr int kThreads = AFD_WPT_THREADS;
constexpr int kMaxPasses = 24;

template <int F>
struct Coefs {
    float lo[F];
    float hi[F];
    float t[F / 2];     // lattice stage tangents (stage 0 first)
};

// One barrier-delimited step of the half tree after level 1.
struct Pass {
    int kind;          // 0: stored level (shared -> shared), 1: last level (shared -> features)
    int in_off;        // float offsets into dynamic shared memory
    int out_off;
    int parents;       // parent nodes handled (power of two)
    int lg_parents;
    int n_out;         // child node length
    int in_stride;
    int out_stride;
    int rsel;          // which of the two item sizes
    float mul;         // factor folded into the stores (1: none)
    int parent_base;   // natural index (within the half tree) of the first parent, last level only
    int n_in;          // parent node length
    int prefetch;      // the next frame's first chunk may be staged while this pass runs
    int sync_before;   // barrier needed before this pass even if the previous pass ended with one
    // chunk classification of the child nodes for the chosen item size (make_split, evaluated on the host)
    int C, CL, CIe, CR, NI, NE;
    unsigned magicNI;
    unsigned magicC;   // floor(i / C) for the uniform item mapping
    int in_group_off;  // frame kernel: float offset of group 1's input / output nodes relative to group 0's (each group keeps
    int out_group_off; // its half tree in its own half of a region: the groups work on different levels at the same time)
};

struct WptPlan {
    int N;                      // samples per frame
    int L;                      // tree depth
    int n1;                     // level-1 node length
    int T;                      // leaf length
    int stride1;                // padded stride of the level-1 node
    int region_b;               // float offset of region B (region A starts at 0)
    int buf_floats;             // floats per staging buffer (two of them at the start of region B)
    int kc;                     // level-1 outputs per staging chunk (multiple of R1)
    int nch;                    // staging chunks per frame
    int npass;
    int smem_floats;            // total dynamic shared memory in floats
    int root_split;             // frame kernel: root float index where group 1's half of region B starts (multiple of 4)
    int stagger;                // frame kernel: cycles group 1 idles after level 1 so that the groups run out of phase
    Pass pass[kMaxPasses];
};

struct Epilogue {
    float power;
    float log_offset;
    int log_scale;
    int sign_channel;
    int order;
    int square;  // power == 2
    // ---- extended epilogue (afd_wpt_forward_ex); all off / null for afd_wpt_forward
    const float* node_scale;   // device [P]: coefficient of output column p is multiplied by node_scale[p] (block norm)
    double* node_stats;        // device [3][P]: sum c, sum c^2, max |c| of the raw coefficients per output column
    double* feat_moments;      // device [C][2]: sum / sum of squares of the features written to each channel
    int normalize;             // features leave as (v - nmean[c]) * nrstd[c]
    float nmean[2];
    float nrstd[2];
    int stats_simple;          // every thread keeps one column pair for the whole launch: statistics stay in registers
    int store;                 // 0: statistics only, nothing is written to `out`
};

// Per-thread running statistics (registers for the whole persistent loop when Epilogue::stats_simple).
struct ThreadStats {
    float s[2], q[2], m[2];    // sum, sum of squares, max |c| of the thread's (lo, hi) columns
    int col[2];                // output columns they belong to (-1: nothing accumulated)
    float fs[2], fq[2];        // feature sum / sum of squares per channel
};

__device__ __forceinline__ void atomic_max_nonneg(double* p, double v) {
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__double_as_longlong(v)));
}

__device__ __forceinline__ void flush_node_stats(const Epilogue& ep, int P, ThreadStats& ts) {
    if (ts.col[0] < 0) return;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        atomicAdd(ep.node_stats + ts.col[i], static_cast<double>(ts.s[i]));
        atomicAdd(ep.node_stats + P + ts.col[i], static_cast<double>(ts.q[i]));
        atomic_max_nonneg(ep.node_stats + 2 * P + ts.col[i], static_cast<double>(ts.m[i]));
        ts.s[i] = 0.f; ts.q[i] = 0.f; ts.m[i] = 0.f;
    }
    ts.col[0] = -1;
}

__device__ __forceinline__ void flush_feat_moments(const Epilogue& ep, int C, ThreadStats& ts) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c >= C) break;
        float a = ts.fs[c], b = ts.fq[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(ep.feat_moments + 2 * c, static_cast<double>(a));
            atomicAdd(ep.feat_moments + 2 * c + 1, static_cast<double>(b));
        }
        ts.fs[c] = 0.f; ts.fq[c] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// Direct form: R outputs (of one or both filters) from a register window.  w[j] = x~[2*k0 + 2 - F + j];
// output r uses x~[2(k0+r)+1-m] = w[2r + F-1-m].
// ------------------------------------------------------------------------------------------------
template <int F, int R, int WLEN>
__device__ __forceinline__ void fir2(const float (&w)[WLEN], const Coefs<F>& cf, float (&lo)[R], float (&hi)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int m = 0; m < F; ++m) {
            const float xv = w[2 * r + F - 1 - m];
            a = fmaf(cf.lo[m], xv, a);
            d = fmaf(cf.hi[m], xv, d);
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

// ------------------------------------------------------------------------------------------------
// Lattice form, in place on the same window.  Pair i (i = 0 .. R+J-2) is p[k0-(J-1)+i] = (x~[2k+1], x~[2k])
// = (w[2i+1], w[2i]); channel u lives in the odd slots, channel v in the even slots.
//   stage 0:          u_i = a + t0 b,            v_i = b - t0 a
//   stage m (1..J-1): u_i = u_i + tm v_{i-1},    v_i = v_{i-1} - tm u_i(old)       (descending i: in place)
// After stage J-1 pairs J-1 .. R+J-2 hold (lo, hi)[k0 .. k0+R-1] / prod(cos theta_m).
// ------------------------------------------------------------------------------------------------
template <int F, int R, int WLEN>
__device__ __forceinline__ void lattice2(float (&w)[WLEN], const Coefs<F>& cf, float (&lo)[R], float (&hi)[R]) {
    constexpr int J = F / 2;
    constexpr int NP = R + J - 1;
    static_assert(2 * NP <= WLEN, "window too short");
    {
        const float t0 = cf.t[0];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const float a = w[2 * i + 1], b = w[2 * i];
            w[2 * i + 1] = fmaf(t0, b, a);
            w[2 * i] = fmaf(-t0, a, b);
        }
    }
#pragma unroll
    for (int m = 1; m < J; ++m) {
        const float tm = cf.t[m];
#pragma unroll
        for (int i = NP - 1; i >= m; --i) {
            const float vd = w[2 * (i - 1)];
            const float uo = w[2 * i + 1];
            w[2 * i + 1] = fmaf(tm, vd, uo);
            w[2 * i] = fmaf(-tm, uo, vd);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        lo[r] = w[2 * (J - 1 + r) + 1];
        hi[r] = w[2 * (J - 1 + r)];
    }
}

// ------------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100: fma.rn.f32x2 -> FFMA2).  One FFMA2 does two FMAs in ONE issue slot at the same
// pipe FLOP rate as two FFMAs (tools/probes/ffma2.cu: 73.6 vs 72.6 TFLOP/s), and these kernels are issue-bound, so
// every FFMA pair folded into an FFMA2 is a freed slot.  64-bit operands are even-aligned register pairs: the
// mov.b64 packs / unpacks below cost nothing when the register allocator can place the halves adjacently.
// AFD_WPT_FFMA2=0 builds the scalar forms (A/B measurements).  Measured on B200 (gpurun_out/ab_ffma2.log): sym5 (F = 10)
// 2 % faster, coif4 (F = 24) 1.5 % slower (the kernel is latency / barrier bound, not issue bound, and the packed
// long-filter forms hold more live registers), so the packed forms are used for F <= kPackedMaxF only.
// ------------------------------------------------------------------------------------------------
#ifndef AFD_WPT_FFMA2
#define AFD_WPT_FFMA2 1
#endif
#ifndef AFD_WPT_PACKED_MAXF
#define AFD_WPT_PACKED_MAXF 16
#endif
constexpr int kPackedMaxF = AFD_WPT_PACKED_MAXF;
// Left reflect padding without stores (r1d): the only reader of a node's left padding is the consumer chunk 0 of the
// next level (when every item size R >= F/2), and for that chunk the reflection x~[-i] = x[i] is a COMPILE-TIME
// permutation of its own register window (w[j] = w[2F-4-j], j < F-2).  So the producers skip the left mirror stores
// (lanes across nodes: 8-way bank conflicts, F-2 scalar stores per channel) and chunk 0 fixes its window up in registers.
// The right padding is still materialised: its reflection point depends on the node length (run time).
#ifndef AFD_WPT_REFLECT_REGS
#define AFD_WPT_REFLECT_REGS 1
#endif
template <int F, int... Rs>
struct ReflOk {
    static constexpr bool value = AFD_WPT_REFLECT_REGS && ((Rs >= F / 2) && ...);
};
// Right padding (AFD_WPT_REFLECT_RIGHT, off): its reflection point is the node length, known only at run time, so the
// chunks whose window crosses the node's end would re-read the mirrored samples with predicated scalar loads
// (w[j] = src[2 jn - j], jn = window index of the node's last sample) instead of the producers storing them.  Measured
// (same-box A/B, bit-identical): coif4 +20 % time, sym5 +1.5 % -- ~70 predicated loads per edge item cost more than the
// F-1 conflicted stores they replace -- so the right padding stays materialised.
#ifndef AFD_WPT_REFLECT_RIGHT
#define AFD_WPT_REFLECT_RIGHT 0
#endif
// Right padding by a copy pass (AFD_WPT_MIRROR_COPY): after the barrier that ends a stored level all threads copy
// pos[n-1+i] = pos[n-1-i] (lanes walk along the padding of a node: conflict free) and a second barrier follows; the
// right-edge items then carry no mirror stores at all.
// Measured (same-box A/B, bit-identical): coif4 (F = 24) -2.3 % time, sym5 (F = 10) +0.3 %: the two extra barriers per level
// pay only when the padding is long, so the copy pass is used for F >= kMirrorCopyMinF.
#ifndef AFD_WPT_MIRROR_COPY
#define AFD_WPT_MIRROR_COPY 1
#endif
#ifndef AFD_WPT_MIRROR_COPY_MINF
#define AFD_WPT_MIRROR_COPY_MINF 16
#endif
constexpr int kMirrorCopyMinF = AFD_WPT_MIRROR_COPY_MINF;
template <int F>
struct MirrorCopy {
    static constexpr bool value = AFD_WPT_MIRROR_COPY != 0 && F >= kMirrorCopyMinF;
};
// Uniform items (r2): with the left padding gone (ReflOk) the only items that differ from the plain interior item are the
// ones holding a node's last coefficients: partial chunks (guarded scalar stores) and the right-mirror stores.  A warp that
// holds one such item runs both store paths back to back and the whole CTA waits for it at the level's barrier (probe
// tools/probes/phase_overlap.cu: a level of pure items takes 1945 cycles per CTA, the real level 2 with ONE edge item 2733,
// level 7 with 64 of them 4240).  So every item stores all of its R coefficients with vector stores -- the ones past the
// node's end are garbage computed from garbage and land in the node's tail space, which the plan sizes for it -- and the
// right padding is always written by the copy pass that follows the level's barrier.
#ifndef AFD_WPT_UNIFORM
#define AFD_WPT_UNIFORM 1
#endif
template <int F, bool UNI>
struct CopyPass {
    static constexpr bool value = UNI || MirrorCopy<F>::value;
};
// With the left padding gone the first chunks of a node are ordinary interior items (lanes along the node, vector stores).
#ifndef AFD_WPT_REFL_INTERIOR
#define AFD_WPT_REFL_INTERIOR 1
#endif
#ifndef AFD_WPT_STRIDE_CONGRUENCE
#define AFD_WPT_STRIDE_CONGRUENCE 1
#endif
#ifndef AFD_WPT_PAIRED_STORE
#define AFD_WPT_PAIRED_STORE 1      // r2: sym5 -3.6 % time, coif4 -1.6 % (same-box A/B, bit-identical)
#endif
#ifndef AFD_WPT_KO_MIRRORS
#define AFD_WPT_KO_MIRRORS 0      // knock-out timing: 1 skips the mirror (padding) stores of the edge items -- wrong results
#endif
#ifndef AFD_WPT_PACKED_FIR_MAXF
#define AFD_WPT_PACKED_FIR_MAXF 32   // level-1 direct form: packed for F <= this (coif4: 1 % faster, same-box A/B); lattice: kPackedMaxF
#endif
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Direct form, one filter, packed over tap pairs: y[r] = sum_q (t[2q] w[e+1] + t[2q+1] w[e]), e = 2r + F-2 - 2q (even),
// so (w[e], w[e+1]) is an aligned pair of the 128-bit window loads.  t2[q] = (t[2q+1], t[2q]).  F/2 FFMA2 + 1 FADD per output.
template <int F, int R, int WLEN>
__device__ __forceinline__ void fir1_packed(const float (&w)[WLEN], const u64 (&t2)[F / 2], float (&y)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        u64 acc = fma2(pk2(w[2 * r + F - 2], w[2 * r + F - 1]), t2[0], pk2(0.f, 0.f));
#pragma unroll
        for (int q = 1; q < F / 2; ++q) {
            const int e = 2 * r + F - 2 - 2 * q;
            acc = fma2(pk2(w[e], w[e + 1]), t2[q], acc);
        }
        float a, b;
        upk2(acc, a, b);
        y[r] = a + b;
    }
}

// Lattice with the middle stages packed: pair indices i and i + H (H = ceil(NP / 2)) share one 64-bit operand, so the
// unit delay (index shift by one) applies to whole operands.  Stage 0 (reads the window registers as loaded) and the
// last stage (its outputs go to paired 64-bit shared-memory stores of adjacent coefficients) stay scalar, which lets
// the register allocator form the pairs without moves.  The first half computes indices below the stage number that
// the scalar form skips: (J-2) * H packed rotations instead of sum (NP - m) scalar ones.
template <int F, int R, int WLEN>
__device__ __forceinline__ void lattice2_packed(const float (&w)[WLEN], const Coefs<F>& cf, float (&lo)[R], float (&hi)[R]) {
    constexpr int J = F / 2;
    constexpr int NP = R + J - 1;
    constexpr int H = (NP + 1) / 2;
    static_assert(J >= 3, "packed lattice needs at least one middle stage");
    static_assert(2 * NP <= WLEN, "window too short");
    u64 U[H], V[H];
    {
        const float t0 = cf.t[0];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float a0 = w[2 * i + 1], b0 = w[2 * i];
            float u1 = 0.f, v1 = 0.f;
            if (i + H < NP) {
                const float a1 = w[2 * (i + H) + 1], b1 = w[2 * (i + H)];
                u1 = fmaf(t0, b1, a1);
                v1 = fmaf(-t0, a1, b1);
            }
            U[i] = pk2(fmaf(t0, b0, a0), u1);
            V[i] = pk2(fmaf(-t0, a0, b0), v1);
        }
    }
#pragma unroll
    for (int m = 1; m < J - 1; ++m) {
        const float tm = cf.t[m];
        const u64 tp = pk2(tm, tm), tn = pk2(-tm, -tm);
        float vl, vh;
        upk2(V[H - 1], vl, vh);
        const u64 vm1 = pk2(0.f, vl);                  // (index -1: never used, index H-1)
#pragma unroll
        for (int i = H - 1; i >= 0; --i) {
            const u64 vp = i > 0 ? V[i - 1] : vm1;
            const u64 un = fma2(tp, vp, U[i]);
            V[i] = fma2(tn, U[i], vp);
            U[i] = un;
        }
    }
    {
        const float tl = cf.t[J - 1];
        float u[2 * H], v[2 * H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            upk2(U[i], u[i], u[i + H]);
            upk2(V[i], v[i], v[i + H]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = J - 1 + r;
            lo[r] = fmaf(tl, v[i - 1], u[i]);
            hi[r] = fmaf(-tl, u[i], v[i - 1]);
        }
    }
}

#ifndef AFD_WPT_SH2_ALIGNED
#define AFD_WPT_SH2_ALIGNED 1
#endif
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

// The same window from an address two floats behind a 16-byte boundary (frame kernel: a root staged with a 2-float shift).
// 64-bit loads of lanes 2R floats apart conflict two ways (R even), so the window is loaded as NV + 1 aligned 128-bit vectors
// from p - 2 (the two floats in front of the window and up to two behind it belong to the staged root and its tail).
template <int NV>
__device__ __forceinline__ void load_window_f2(const float* __restrict__ p, float (&w)[4 * NV]) {
#if AFD_WPT_SH2_ALIGNED
    const float4* src = reinterpret_cast<const float4*>(p - 2);
    float4 q = src[0];
    w[0] = q.z; w[1] = q.w;
#pragma unroll
    for (int v = 1; v < NV; ++v) {
        q = src[v];
        w[4 * v - 2] = q.x; w[4 * v - 1] = q.y; w[4 * v] = q.z; w[4 * v + 1] = q.w;
    }
    q = src[NV];
    w[4 * NV - 2] = q.x; w[4 * NV - 1] = q.y;
#else
    const float2* src = reinterpret_cast<const float2*>(p);
#pragma unroll
    for (int v = 0; v < 2 * NV; ++v) {
        const float2 q = src[v];
        w[2 * v] = q.x; w[2 * v + 1] = q.y;
    }
#endif
}

template <int F, int R, bool LAT, bool REFL, bool SH2 = false>
__device__ __forceinline__ void filter_pair(const float* __restrict__ src, const Coefs<F>& cf, float (&lo)[R],
                                            float (&hi)[R], bool first, int jn) {
    using WN = Win<F, R>;
    float w[WN::WLEN];
    if constexpr (SH2) load_window_f2<WN::NV>(src, w);
    else load_window<WN::NV>(src, w);
    if constexpr (REFL && AFD_WPT_REFLECT_RIGHT) {
        if (jn < WN::W - 1) {              // the window crosses the node's end: x~[n-1+i] = x[n-1-i], i = 1 .. F-2 (+1)
            const unsigned padr = static_cast<unsigned>(F - 2 + ((jn + 1) & 1));     // jn = n + F - 3 - 2 k0: n odd <=> jn even
#pragma unroll
            for (int j = 1; j < WN::W; ++j)
                if (static_cast<unsigned>(j - jn - 1) < padr) w[j] = src[2 * jn - j];
        }
    }
    if constexpr (REFL) {
        if (first) {                       // chunk 0: the left padding is the mirror image of the window's own samples
#pragma unroll
            for (int j = 0; j < F - 2; ++j) w[j] = w[2 * F - 4 - j];
        }
    }
    if constexpr (LAT && AFD_WPT_FFMA2 && F >= 6 && F <= kPackedMaxF) lattice2_packed<F, R>(w, cf, lo, hi);
    else if constexpr (LAT) lattice2<F, R>(w, cf, lo, hi);
    else fir2<F, R>(w, cf, lo, hi);
}

// Chunk classification of a node with n_out coefficients, stored with padl / padr mirrored samples:
// chunk c (outputs cR .. cR+R-1) is "interior" when it holds no mirrored coefficient and is fully valid.
struct Split {
    int C, CL, CIe, NI, NE;
    int CR;               // first chunk holding a right-mirrored or invalid coefficient (unclamped CIe)
    unsigned magicNI;     // floor(i / NI) = umulhi(i, magic) for i, NI < 2^16
};
__host__ __device__ inline unsigned magic_of(int d) { return d > 0 ? static_cast<unsigned>(0xFFFFFFFFu / static_cast<unsigned>(d)) + 1u : 0u; }

__host__ __device__ inline Split make_split(int n_out, int R, int padl, bool no_left = false) {
    Split s;
    s.C = (n_out + R - 1) / R;
    const int padr = padl + (n_out & 1);
    int cl = (padl == 0 || no_left) ? 0 : padl / R + 1;    // no_left: the left padding is not stored (ReflOk): no left-edge chunks
    int cie = (n_out - 1 - padr) / R;            // chunks c < cie end before the first right-mirrored coefficient
    if (n_out - 1 - padr < 0) cie = 0;
    s.CR = cie;
    cl = cl < s.C ? cl : s.C;
    cie = cie > cl ? cie : cl;
    cie = cie < s.C ? cie : s.C;
    s.CL = cl; s.CIe = cie; s.NI = cie - cl; s.NE = s.C - s.NI;
    s.magicNI = magic_of(s.NI);
    return s;
}
__device__ __forceinline__ int fast_div(int i, int d, unsigned magic) { return d == 1 ? i : static_cast<int>(__umulhi(static_cast<unsigned>(i), magic)); }

template <int R>
__device__ __forceinline__ void vec_store(float* __restrict__ dst, const float (&v)[R]) {
    float2* d = reinterpret_cast<float2*>(dst);
#pragma unroll
    for (int r = 0; r < R / 2; ++r) d[r] = make_float2(v[2 * r], v[2 * r + 1]);
}

// Generic guarded store of R coefficients of a child node plus their mirror images into the node's padding.
// `node` points at the first padding sample; coefficient k lives at node[padl + k].
template <int R, bool REFL, bool MC>
__device__ __forceinline__ void edge_store(float* __restrict__ node, const float (&v)[R], int k0, int n_out, int padl) {
    const int padr = padl + (n_out & 1);
    float* pos = node + padl;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int k = k0 + r;
        if (k < n_out) {
            pos[k] = v[r];
            if (!REFL && !AFD_WPT_KO_MIRRORS && k >= 1 && k <= padl) pos[-k] = v[r];
            const int mr = n_out - 1 - k;
            if (!(REFL && AFD_WPT_REFLECT_RIGHT) && !MC && !AFD_WPT_KO_MIRRORS && mr >= 1 && mr <= padr) pos[n_out - 1 + mr] = v[r];
        }
    }
}

// Left-edge chunk C (compile time) of a pair of sibling nodes: fully valid, mirrors k = 1 .. padl to -k.
template <int R, int PADL, int C, bool REFL>
__device__ __forceinline__ void left_store(float* __restrict__ plo, float* __restrict__ phi, const float (&lo)[R],
                                           const float (&hi)[R]) {
    vec_store<R>(plo + C * R, lo);
    vec_store<R>(phi + C * R, hi);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int k = C * R + r;
        if (!REFL && !AFD_WPT_KO_MIRRORS && k >= 1 && k <= PADL) {
            plo[-k] = lo[r];
            phi[-k] = hi[r];
        }
    }
}

// Right-edge chunk of a pair of sibling nodes (no left-mirrored coefficient inside): outputs k0+r are valid for
// r <= q = n_out-1-k0 and mirrored to n_out-1+mr (mr = q - r) for 1 <= mr <= padr; one predicate pair serves both
// channels.  plo / phi point at coefficient 0.
template <int R, bool REFL, bool MC>
__device__ __forceinline__ void right_store(float* __restrict__ plo, float* __restrict__ phi, const float (&lo)[R],
                                            const float (&hi)[R], int k0, int n_out, int padl) {
    const int padr = padl + (n_out & 1);
    const int q = n_out - 1 - k0;
    float* dlo = plo + k0;
    float* dhi = phi + k0;
    float* mlo = plo + (n_out - 1 + q);
    float* mhi = phi + (n_out - 1 + q);
    if (q >= R - 1) {                      // every coefficient of the chunk exists
        vec_store<R>(dlo, lo);
        vec_store<R>(dhi, hi);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r <= q) { dlo[r] = lo[r]; dhi[r] = hi[r]; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (!(REFL && AFD_WPT_REFLECT_RIGHT) && !MC && !AFD_WPT_KO_MIRRORS && static_cast<unsigned>(q - r - 1) < static_cast<unsigned>(padr)) { mlo[-r] = lo[r]; mhi[-r] = hi[r]; }
}

template <int R>
__device__ __forceinline__ void scale_all(float (&lo)[R], float (&hi)[R], float mul) {
#pragma unroll
    for (int r = 0; r < R; ++r) { lo[r] *= mul; hi[r] *= mul; }
}

// Right reflect padding of `nodes` stored nodes of length n: pos[n-1+i] = pos[n-1-i], i = 1 .. F-2 (+1 if n is odd).
// E = next power of two >= F-1 consecutive threads serve one node (shifts only; lanes walk along the padding).
__host__ __device__ constexpr int pow2_ge(int v) { int e = 1; while (e < v) e <<= 1; return e; }
template <int F>
__device__ __forceinline__ void mirror_copy(float* __restrict__ base, int nodes, int n, int stride,
                                            int tid = threadIdx.x, int nthr = kThreads) {
    constexpr int padl = F - 2;
    constexpr int E = pow2_ge(F - 1) < 128 ? pow2_ge(F - 1) : 128;
    const int padr = padl + (n & 1);
    if (padr == 0) return;
    const int e = tid & (E - 1);
    for (int mr = e + 1; mr <= padr; mr += E)                       // one trip (F - 1 <= 128)
        for (int node = tid / E; node < nodes; node += nthr / E) {
            float* pos = base + node * stride + padl + (n - 1);
            pos[mr] = pos[-mr];
        }
}

// One stored tree level with uniform items (see AFD_WPT_UNIFORM): item = (node, chunk), lanes walk along a node.
template <int F, int R, bool LAT, bool SH2 = false>
__device__ __forceinline__ void mid_level_uniform(const float* __restrict__ in, float* __restrict__ out, const Pass& ps,
                                                  const Coefs<F>& cf, int tid = threadIdx.x, int nthr = kThreads) {
    constexpr int padl = F - 2;
    const int total = ps.parents * ps.C;
    const bool do_mul = ps.mul != 1.0f;
    for (int it = tid; it < total; it += nthr) {
        const int node = fast_div(it, ps.C, ps.magicC);
        const int c = it - node * ps.C;
        const int k0 = c * R;
        float lo[R], hi[R];
        filter_pair<F, R, LAT, true, SH2>(in + node * ps.in_stride + 2 * k0, cf, lo, hi, c == 0, 0);
        if (do_mul) scale_all<R>(lo, hi, ps.mul);
        float* d0 = out + (2 * node) * ps.out_stride + padl + k0;
        vec_store<R>(d0, lo);
        vec_store<R>(d0 + ps.out_stride, hi);
    }
}

// One stored tree level: `parents` padded nodes in `in` -> 2*parents padded nodes in `out`.
// Interior chunks first (lanes walk along a node), then the edge chunks type-major (lanes walk across nodes).
template <int F, int R, bool LAT, bool REFL>
__device__ __forceinline__ void mid_level(const float* __restrict__ in, float* __restrict__ out, const Pass& ps,
                                          const Coefs<F>& cf) {
    constexpr int padl = F - 2;
    const int n_out = ps.n_out;
    Split sp;
    sp.C = ps.C; sp.CL = ps.CL; sp.CIe = ps.CIe; sp.CR = ps.CR; sp.NI = ps.NI; sp.NE = ps.NE; sp.magicNI = ps.magicNI;
    const int parents = ps.parents;
    const int n_int = parents * sp.NI;
    const int total = parents * sp.C;
    const bool do_mul = ps.mul != 1.0f;
    for (int it = threadIdx.x; it < total; it += kThreads) {
        const bool edge = it >= n_int;
        int node, c;
        if (!edge) {
            node = fast_div(it, sp.NI, sp.magicNI);
            c = sp.CL + it - node * sp.NI;
        } else {
            const int e = it - n_int;
            node = e & (parents - 1);
            const int ee = e >> ps.lg_parents;
            c = ee < sp.CL ? ee : sp.CIe + (ee - sp.CL);
        }
        const int k0 = c * R;
        float lo[R], hi[R];
        filter_pair<F, R, LAT, REFL>(in + node * ps.in_stride + 2 * k0, cf, lo, hi, c == 0, ps.n_in + F - 3 - 2 * k0);
        if (do_mul) scale_all<R>(lo, hi, ps.mul);
        float* d0 = out + (2 * node) * ps.out_stride;
        float* d1 = d0 + ps.out_stride;
        if (!edge) {
            vec_store<R>(d0 + padl + k0, lo);
            vec_store<R>(d1 + padl + k0, hi);
        } else {
            const bool left = c < sp.CL, right = c >= sp.CR;
            if (left && !right && c == 0) left_store<R, padl, 0, REFL>(d0 + padl, d1 + padl, lo, hi);
            else if (left && !right && c == 1) left_store<R, padl, 1, REFL>(d0 + padl, d1 + padl, lo, hi);
            else if (right && !left) right_store<R, REFL, MirrorCopy<F>::value>(d0 + padl, d1 + padl, lo, hi, k0, n_out, padl);
            else {
                edge_store<R, REFL, MirrorCopy<F>::value>(d0, lo, k0, n_out, padl);
                edge_store<R, REFL, MirrorCopy<F>::value>(d1, hi, k0, n_out, padl);
            }
        }
    }
}

// Level 1 for one staged chunk: outputs [kb, ke) of the CTA's own filter -> padded level-1 node.
// Sample 2*kb + 2 - F of the (reflect-extended) frame sits at buf[0].
template <int F, int R, bool REFL, bool UNI>
__device__ __forceinline__ void level1_chunk(const float* __restrict__ buf, float* __restrict__ node, int kb, int ke,
                                             int n_out, const Split& sp, const float (&t)[F], const u64 (&t2)[F / 2]) {
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
        if constexpr (AFD_WPT_FFMA2 != 0 && F <= AFD_WPT_PACKED_FIR_MAXF) fir1_packed<F, R>(w, t2, y);
        else fir1<F, R>(w, t, y);
        if (UNI || (c >= sp.CL && c < sp.CIe)) vec_store<R>(node + padl + k0, y);
        else edge_store<R, REFL, MirrorCopy<F>::value>(node, y, k0, n_out, padl);
    }
}
// Chunks are cut at multiples of R, so only the last chunk (ke == n_out) holds a partial item.

__device__ __forceinline__ unsigned igray(unsigned x) {
    x ^= x >> 1; x ^= x >> 2; x ^= x >> 4; x ^= x >> 8;
    return x;
}

// Last level: `parents` padded nodes (level L-1) -> features in global memory.  Lanes map to parents.
template <int F, int RL, bool LAT, bool EXT, bool REFL>
__device__ __forceinline__ void last_level(const float* __restrict__ in, const Pass& ps, int T, int half_base,
                                           float* __restrict__ out_b, int P, const Coefs<F>& cf, const Epilogue& ep,
                                           ThreadStats& ts, int tid = threadIdx.x, int nthr = kThreads) {
    const int parents = ps.parents;
    const int chunks = (T + RL - 1) / RL;
    const int total = parents * chunks;
    const bool two = ep.log_scale && ep.sign_channel;
    const bool do_mul = ps.mul != 1.0f;
    const int mode = !ep.log_scale ? 0 : (ep.square ? 1 : 2);    // 0 raw, 1 log(c^2 + off), 2 log(|c|^p + off)
    const float off = ep.log_offset;
    const long long ch1 = static_cast<long long>(T) * P;
    constexpr bool extended = EXT;                                   // afd_wpt_forward_ex instantiation
    for (int it = tid; it < total; it += nthr) {
        const int m = it & (parents - 1);
        const int c = it >> ps.lg_parents;
        const int k0 = c * RL;
        float lo[RL], hi[RL];
        filter_pair<F, RL, LAT, REFL>(in + m * ps.in_stride + 2 * k0, cf, lo, hi, c == 0, ps.n_in + F - 3 - 2 * k0);
        if (do_mul) scale_all<RL>(lo, hi, ps.mul);
        const unsigned pf = static_cast<unsigned>(half_base + ps.parent_base + m);      // natural index at level L-1
        unsigned q = pf;
        bool swap = false;
        if (ep.order == AFD_ORDER_FREQ) {
            q = igray(pf);
            swap = (q & 1u) != 0;                                      // parity(pf) = lsb of igray(pf)
        }
        // children of natural node m sit at columns 2q + {0, 1}; `swap` exchanges them: resolved by addressing
        const int col_lo = static_cast<int>(2 * q) + (swap ? 1 : 0);
        const int col_hi = static_cast<int>(2 * q) + (swap ? 0 : 1);
        float* o_lo = out_b + static_cast<long long>(k0) * P + col_lo;
        float* o_hi = out_b + static_cast<long long>(k0) * P + col_hi;
        const int nv = T - k0;                                         // rows r < nv exist
        if constexpr (extended) {
            if (ep.node_stats) {                                       // statistics of the raw coefficients
                if (ts.col[0] != col_lo) {                             // taken once per launch when stats_simple
                    flush_node_stats(ep, P, ts);
                    ts.col[0] = col_lo; ts.col[1] = col_hi;
                }
#pragma unroll
                for (int r = 0; r < RL; ++r)
                    if (r < nv) {
                        ts.s[0] += lo[r]; ts.q[0] = fmaf(lo[r], lo[r], ts.q[0]); ts.m[0] = fmaxf(ts.m[0], fabsf(lo[r]));
                        ts.s[1] += hi[r]; ts.q[1] = fmaf(hi[r], hi[r], ts.q[1]); ts.m[1] = fmaxf(ts.m[1], fabsf(hi[r]));
                    }
            }
            if (ep.node_scale) {
                const float s_lo = __ldg(ep.node_scale + col_lo), s_hi = __ldg(ep.node_scale + col_hi);
#pragma unroll
                for (int r = 0; r < RL; ++r) { lo[r] *= s_lo; hi[r] *= s_hi; }
            }
        }
        if (two) {
            float sl = 1.f, sh = 1.f, dl = 0.f;                        // sign channel: (s - mean1) * rstd1
            if (extended && ep.normalize) { sl = ep.nrstd[1]; sh = ep.nrstd[1]; dl = -ep.nmean[1] * ep.nrstd[1]; }
#pragma unroll
            for (int r = 0; r < RL; ++r)
                if (r < nv) {
                    const float a = lo[r] < 0.f ? -1.f : 1.f, d = hi[r] < 0.f ? -1.f : 1.f;
                    if (extended && ep.feat_moments) { ts.fs[1] += a + d; ts.fq[1] += 2.f; }
                    if (!extended || ep.store) {
                        __stcs(o_lo + ch1 + r * P, extended ? fmaf(a, sl, dl) : a);   // r * P < T * P < 2^31
                        __stcs(o_hi + ch1 + r * P, extended ? fmaf(d, sh, dl) : d);
                    }
                }
        }
        if (mode == 1) {
#pragma unroll
            for (int r = 0; r < RL; ++r) { lo[r] = ln_approx(fmaf(lo[r], lo[r], off)); hi[r] = ln_approx(fmaf(hi[r], hi[r], off)); }
        } else if (mode == 2) {
#pragma unroll
            for (int r = 0; r < RL; ++r) { lo[r] = log_power(lo[r], ep.power, off, false); hi[r] = log_power(hi[r], ep.power, off, false); }
        }
        if constexpr (extended) {
            if (ep.feat_moments) {
#pragma unroll
                for (int r = 0; r < RL; ++r)
                    if (r < nv) {
                        ts.fs[0] += lo[r] + hi[r];
                        ts.fq[0] = fmaf(lo[r], lo[r], fmaf(hi[r], hi[r], ts.fq[0]));
                    }
            }
            if (ep.normalize) {
                const float rs = ep.nrstd[0], dm = -ep.nmean[0] * ep.nrstd[0];
#pragma unroll
                for (int r = 0; r < RL; ++r) { lo[r] = fmaf(lo[r], rs, dm); hi[r] = fmaf(hi[r], rs, dm); }
            }
        }
        if (!extended || ep.store) {
#if AFD_WPT_PAIRED_STORE
            // the two children of a parent sit in adjacent columns 2q, 2q + 1 (in either order): one 64-bit streaming store per row
            // writes whole sectors and halves the store instructions (48 -> 24 per item); the order is resolved by two selects
            float2* o2 = reinterpret_cast<float2*>(out_b + static_cast<long long>(k0) * P + 2 * static_cast<int>(q));
#pragma unroll
            for (int r = 0; r < RL; ++r)
                if (r < nv) __stcs(o2 + r * (P / 2), make_float2(swap ? hi[r] : lo[r], swap ? lo[r] : hi[r]));
#else
#pragma unroll
            for (int r = 0; r < RL; ++r)
                if (r < nv) {
                    __stcs(o_lo + r * P, lo[r]);
                    __stcs(o_hi + r * P, hi[r]);
                }
#endif
        }
    }
    if (extended && ep.node_stats && !ep.stats_simple) flush_node_stats(ep, P, ts);
}

// Stage chunk j of the frame (with the frame's own reflect padding) into `buf` with cp.async.
template <int F>
__device__ __forceinline__ void issue_chunk(const float* __restrict__ xg, float* __restrict__ buf, int j,
                                            const WptPlan& plan, const int nthr = kThreads, const int tid_in = -1) {
    const int N = plan.N;
    const int kb = j * plan.kc;
    const int ke = min(plan.n1, kb + plan.kc);
    const int s_start = 2 * kb + 2 - F;           // sample index stored at buf[0] (even)
    const int s_end = 2 * ke;                     // one past the last sample any stored output needs
    const int r_lo = max(s_start, 0);
    const int r_hi = min(s_end, N);               // real samples [r_lo, r_hi)
    const int tid = tid_in >= 0 ? tid_in : static_cast<int>(threadIdx.x);
    {
        // widest copy both addresses allow (a frame is 88,200 B, so odd frames are only 8-byte aligned)
        const char* src = reinterpret_cast<const char*>(xg + r_lo);
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(buf + (r_lo - s_start)));
        const int n = r_hi - r_lo;
        const unsigned mis = static_cast<unsigned>(reinterpret_cast<uintptr_t>(src)) | dst;
        if (((static_cast<unsigned>(reinterpret_cast<uintptr_t>(src)) ^ dst) & 15) == 0) {
            // congruent modulo 16 bytes: single floats up to the first 16-byte boundary, 16-byte copies, single floats
            const int head = min(n, static_cast<int>(((16u - (dst & 15u)) & 15u) >> 2));
            const int units = (n - head) >> 2;
            const char* s = src + 4 * head + 16 * tid;
            uint32_t d = dst + 4 * head + 16 * tid;
            for (int i = tid; i < units; i += nthr, s += 16 * nthr, d += 16 * nthr)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(s));
            if (tid < head) cp_async_4(buf + (r_lo - s_start) + tid, xg + r_lo + tid);
            if (tid < n - head - 4 * units) cp_async_4(buf + (r_lo - s_start) + head + 4 * units + tid, xg + r_lo + head + 4 * units + tid);
        } else if ((mis & 7) == 0) {
            const int units = n >> 1;
            const char* s = src + 8 * tid;
            uint32_t d = dst + 8 * tid;
            for (int i = tid; i < units; i += nthr, s += 8 * nthr, d += 8 * nthr)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(s));
            if ((n & 1) && tid == 0) cp_async_4(buf + (r_hi - 1 - s_start), xg + r_hi - 1);
        } else {
            for (int i = r_lo + tid; i < r_hi; i += nthr) cp_async_4(buf + (i - s_start), xg + i);
        }
    }
    // reflect padding of the frame itself: x~[-i] = x[i], x~[N-1+i] = x[N-1-i]
    for (int s = s_start + tid; s < 0; s += nthr) cp_async_4(buf + (s - s_start), xg - s);
    for (int s = max(N, s_start) + tid; s < s_end; s += nthr) cp_async_4(buf + (s - s_start), xg + (2 * (N - 1) - s));
    cp_async_commit();
}

// Phase timing (debug builds, -DAFD_WPT_PHASE_TIMING=1): thread 0 of every CTA adds the clock64() span of each
// barrier-delimited phase to g_wpt_phase[]; launch() prints the table after every launch.  Slots: 0 wait for the staged
// chunk (cp.async + barrier), 1 level-1 filtering, 2 barrier + copy pass after level 1, 3+pi pass pi (incl. its closing
// barrier for stored levels; the last level is closed by the next frame's first barrier, i.e. slot 0), 30 frames, 31 total.
#ifndef AFD_WPT_PHASE_TIMING
#define AFD_WPT_PHASE_TIMING 0
#endif
#if AFD_WPT_PHASE_TIMING
static __device__ unsigned long long g_wpt_phase[32];
#define AFD_PHASE_MARK(slot)                                                                       \
    do {                                                                                           \
        if (threadIdx.x == 0) {                                                                    \
            const long long now_ = clock64();                                                      \
            atomicAdd(&g_wpt_phase[slot], static_cast<unsigned long long>(now_ - phase_t0_));      \
            phase_t0_ = now_;                                                                      \
        }                                                                                          \
    } while (0)
#else
#define AFD_PHASE_MARK(slot) do { } while (0)
#endif

template <int F, int R1, int RA, int RB, int RLA, int RLB, bool LAT, bool EXT>
__global__ void __launch_bounds__(kThreads, 2)
wpt_tree_kernel(const float* __restrict__ x, long long x_row_stride, long long B, float* __restrict__ out,
                const __grid_constant__ WptPlan plan, const __grid_constant__ Coefs<F> cf,
                const __grid_constant__ Epilogue ep) {
    extern __shared__ __align__(16) float smem[];
    constexpr int padl = F - 2;
    constexpr bool REFL = ReflOk<F, RA, RB, RLA, RLB>::value;      // left padding by register reflection (see ReflOk)
    constexpr bool UNI = REFL && AFD_WPT_UNIFORM != 0;               // uniform items + copy pass (see AFD_WPT_UNIFORM)
    constexpr bool COPY = CopyPass<F, UNI>::value;
    const int L = plan.L;
    const int half = blockIdx.x & 1;                      // gridDim.x is even: constant per CTA
    float* const regA = smem;                              // level-1 node at its start
    float* const regB = smem + plan.region_b;              // staging buffers at its start
    float* const buf0 = regB;
    float* const buf1 = regB + plan.buf_floats;
    const int P = 1 << L;
    const int C = (ep.log_scale && ep.sign_channel) ? 2 : 1;
    const int T = plan.T;
    const int n1 = plan.n1;
    const Split sp1 = make_split(n1, R1, padl, REFL && AFD_WPT_REFL_INTERIOR);
    const int half_base = L >= 2 ? half << (L - 2) : 0;   // natural index of the half tree's first level-(L-1) node
    // the CTA's level-1 filter
    float t1[F];
#pragma unroll
    for (int m = 0; m < F; ++m) t1[m] = half ? cf.hi[m] : cf.lo[m];
    u64 t1p[F / 2];                                        // (t[2q+1], t[2q]) for the packed direct form
#pragma unroll
    for (int q = 0; q < F / 2; ++q) t1p[q] = pk2(t1[2 * q + 1], t1[2 * q]);
    bool prefetched = false;
    bool first = true;
    ThreadStats ts;
    ts.s[0] = ts.s[1] = ts.q[0] = ts.q[1] = ts.m[0] = ts.m[1] = 0.f;
    ts.fs[0] = ts.fs[1] = ts.fq[0] = ts.fq[1] = 0.f;
    ts.col[0] = ts.col[1] = -1;
#if AFD_WPT_PHASE_TIMING
    long long phase_t0_ = clock64();
    const long long phase_start_ = phase_t0_;
#endif

    for (long long wk = blockIdx.x; wk < 2 * B; wk += gridDim.x) {
        const long long b = wk >> 1;
        const float* xg = x + b * x_row_stride;
        const long long nb = wk + gridDim.x;
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
            AFD_PHASE_MARK(0);
            if (j + 1 < plan.nch) issue_chunk<F>(xg, ((j + 1) & 1) ? buf1 : buf0, j + 1, plan);
            const int kb = j * plan.kc;
            const int ke = min(n1, kb + plan.kc);
            level1_chunk<F, R1, REFL, UNI>((j & 1) ? buf1 : buf0, regA, kb, ke, n1, sp1, t1, t1p);
            AFD_PHASE_MARK(1);
        }
        __syncthreads();
        if constexpr (COPY) {
            if (L > 1) {
                mirror_copy<F>(regA, 1, n1, plan.stride1);
                __syncthreads();
            }
        }
        AFD_PHASE_MARK(2);
        float* out_b = out + b * C * static_cast<long long>(T) * P;
        if (L == 1) {
            // the level-1 node is the output: epilogue straight from shared memory (rare configuration)
            if (nb < 2 * B) { issue_chunk<F>(x + (nb >> 1) * x_row_stride, buf0, 0, plan); prefetched = true; }
            const bool square = ep.square != 0;
            if (EXT && ep.node_stats) { ts.col[0] = half; ts.col[1] = half; }      // slot 1 stays empty (adds zeros)
            for (int e = threadIdx.x; e < T; e += kThreads) {
                float c = regA[padl + e];
                if (EXT && ep.node_stats) { ts.s[0] += c; ts.q[0] = fmaf(c, c, ts.q[0]); ts.m[0] = fmaxf(ts.m[0], fabsf(c)); }
                if (EXT && ep.node_scale) c *= __ldg(ep.node_scale + half);
                float* dst = out_b + static_cast<long long>(e) * P + half;
                float v = ep.log_scale ? log_power(c, ep.power, ep.log_offset, square) : c;
                if (EXT && ep.feat_moments) { ts.fs[0] += v; ts.fq[0] = fmaf(v, v, ts.fq[0]); }
                if (EXT && ep.normalize) v = (v - ep.nmean[0]) * ep.nrstd[0];
                if (!EXT || ep.store) st_cs(dst, v);
                if (C == 2) {
                    float sg = c < 0.f ? -1.f : 1.f;
                    if (EXT && ep.feat_moments) { ts.fs[1] += sg; ts.fq[1] += 1.f; }
                    if (EXT && ep.normalize) sg = (sg - ep.nmean[1]) * ep.nrstd[1];
                    if (!EXT || ep.store) st_cs(dst + static_cast<long long>(T) * P, sg);
                }
            }
            if (EXT && ep.node_stats) flush_node_stats(ep, P, ts);
            continue;
        }
        // ---------------------------------------------------------------- levels 2 .. L as planned passes
        for (int pi = 0; pi < plan.npass; ++pi) {
            const Pass& ps = plan.pass[pi];
            if (ps.sync_before) __syncthreads();
            if (ps.kind == 0) {
                if constexpr (UNI) {
                    if (ps.rsel == 0) mid_level_uniform<F, RA, LAT>(smem + ps.in_off, smem + ps.out_off, ps, cf);
                    else mid_level_uniform<F, RB, LAT>(smem + ps.in_off, smem + ps.out_off, ps, cf);
                } else {
                    if (ps.rsel == 0) mid_level<F, RA, LAT, REFL>(smem + ps.in_off, smem + ps.out_off, ps, cf);
                    else mid_level<F, RB, LAT, REFL>(smem + ps.in_off, smem + ps.out_off, ps, cf);
                }
                __syncthreads();
                if constexpr (COPY) {
                    mirror_copy<F>(smem + ps.out_off, 2 * ps.parents, ps.n_out, ps.out_stride);
                    __syncthreads();
                }
                AFD_PHASE_MARK(3 + pi);
            } else {
                if (ps.prefetch && nb < 2 * B) {
                    issue_chunk<F>(x + (nb >> 1) * x_row_stride, buf0, 0, plan);
                    prefetched = true;
                }
                if (ps.rsel == 0) last_level<F, RLA, LAT, EXT, REFL>(smem + ps.in_off, ps, T, half_base, out_b, P, cf, ep, ts);
                else last_level<F, RLB, LAT, EXT, REFL>(smem + ps.in_off, ps, T, half_base, out_b, P, cf, ep, ts);
                AFD_PHASE_MARK(3 + pi);
            }
        }
#if AFD_WPT_PHASE_TIMING
        if (threadIdx.x == 0) atomicAdd(&g_wpt_phase[30], 1ull);
#endif
    }
#if AFD_WPT_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_wpt_phase[31], static_cast<unsigned long long>(clock64() - phase_start_));
#endif
    cp_async_wait<0>();
    if constexpr (EXT) {
        if (ep.node_stats) flush_node_stats(ep, P, ts);   // stats_simple: the one flush of the launch
        if (ep.feat_moments) flush_feat_moments(ep, C, ts);
    }
}


// ------------------------------------------------------------------------------------------------
// Frame kernel (r2): ONE persistent CTA of 512 threads per frame, for every filter the lattice serves.
//
// The two-CTA kernel above splits a frame by level-1 FILTER, so level 1 runs in the direct form (F FMAs per output, the
// other CTA computes the other channel from its own copy of the frame) and is staged in four barrier-delimited chunks:
// 21 % (sym5) / 22 % (coif4) of the step.  Here the whole frame is staged ONCE as a padded root node (reflect padding
// included) and level 1 is an ordinary lattice pass over it -- F FMAs per output PAIR, one round of the 512 threads,
// half the staging traffic.  From level 2 on the two half trees are independent: thread group g (256 threads) owns the
// sub-tree under the level-1 node g and synchronises on its own named barrier, so the groups drift apart exactly like the
// two CTAs did (one group's shared-memory phases overlap the other's FMAs).  The CTA-wide barriers left are the ones
// around level 1 and the one that frees region B for the prefetch of the next frame while the last level runs.
// Shared memory: region A = odd levels of the FULL tree, region B = root + even levels (~185 KB sym5, ~215 KB coif4).
// ------------------------------------------------------------------------------------------------
#ifndef AFD_WPT_FRAME_KERNEL
#define AFD_WPT_FRAME_KERNEL 1
#endif
#ifndef AFD_WPT_STAGGER_DEFAULT
#define AFD_WPT_STAGGER_DEFAULT 1200
#endif
constexpr int kFrameThreads = 512;
constexpr int kGroupThreads = 256;

__device__ __forceinline__ void group_barrier(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
}

// Staging of a frame as the padded root node by ONE warp: a single bulk copy (TMA engine; no thread waits for it and nothing
// enters the LSU queue -- the 5512 cp.async.16 of the first version kept the issuing group busy for 4.7 k cycles per frame,
// 16 % of the step, in front of its last level) plus a few scalar copies.  Sample i of the frame sits at root[F - 2 + d + i];
// d (0 or 2 floats) makes source and destination congruent mod 16 bytes (a frame is 88,200 B: odd frames are 8-byte aligned).
// By threads: the samples in front of the first / behind the last whole 16-byte unit and the right reflect padding
// x~[N-1+i] = x[N-1-i]; the left padding is never read (register reflection of the first item).
template <int F>
__device__ __forceinline__ int root_shift(const float* xg) {
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(xg) & 15u) >> 2);       // bulk staging: 0 or 2 (host-checked)
    const int d = (a - (F - 2)) & 3;
    return (d & 1) ? 0 : d;                                // 4-byte aligned rows (cp.async staging only): no shift, narrow copies
}
// The root is staged in two parts, one per half-tree group: part g is what lies in group g's half of region B (root index
// < / >= split, a multiple of four floats, so both parts start on 16-byte boundaries in global and shared memory), and group g
// may stage it as soon as IT has finished with that half.  Each part is one warp's work: `AFD_WPT_ROOT_COPIES` bulk copies on
// the shared mbarrier (initialised to two arrivals, one per part) plus the scalar copies at the part's outer end.
#ifndef AFD_WPT_ROOT_COPIES
#define AFD_WPT_ROOT_COPIES 1
#endif
template <int F>
__device__ __forceinline__ void stage_root_bulk(const float* __restrict__ xg, float* __restrict__ root, int N, uint32_t bar, int lane,
                                                int part, int split) {
    constexpr int padl = F - 2;
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(xg) & 15u) >> 2);
    const int d = (a - padl) & 3;
    const int s0 = (4 - a) & 3;                             // first sample at a 16-byte boundary
    const int units = (N - s0) >> 2;
    float* const dst = root + padl + d;                     // position of sample 0
    int sm = split - padl - d;                              // first sample of part 1: (sm - s0) is a multiple of 4
    sm = sm < s0 ? s0 : (sm > s0 + 4 * units ? s0 + 4 * units : sm);
    const int b0 = part == 0 ? s0 : sm, b1 = part == 0 ? sm : s0 + 4 * units;        // this part's bulk range, in samples
    if (lane == 0) {
        // the bulk copies first: the scalar copies below wait for their loads (a DRAM round trip) before they can store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = 4u * static_cast<uint32_t>(b1 - b0);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const uint32_t piece = ((bytes / AFD_WPT_ROOT_COPIES) + 15u) & ~15u;
        uint32_t off = 0;
#pragma unroll
        for (int c = 0; c < AFD_WPT_ROOT_COPIES; ++c) {
            const uint32_t nb = (c == AFD_WPT_ROOT_COPIES - 1 || off + piece > bytes) ? bytes - off : piece;
            if (nb)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst + b0)) + off),
                               "l"(reinterpret_cast<const char*>(xg + b0) + off), "r"(nb), "r"(bar) : "memory");
            off += nb;
        }
    }
    // by threads (visible to the CTA through the frame's first barrier, not through the mbarrier)
    if (part == 0) {
        if (lane < s0) dst[lane] = __ldg(xg + lane);
    } else {
        for (int s = s0 + 4 * units + lane; s < N; s += 32) dst[s] = __ldg(xg + s);
        const int padr = padl + 1 + 3;                      // right padding incl. the odd-length sample and the item over-read's first floats
        for (int j = lane; j < padr && j < N - 1; j += 32) dst[N + j] = __ldg(xg + (N - 2 - j));
    }
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// Bulk staging pays for the short filters (sym5: -3 % time); for coif4 it is 1.9 % SLOWER than the cp.async staging (the
// issue time it removes was hidden behind the other group's work, and half of the frames pay 8-byte window loads in level 1).
#ifndef AFD_WPT_BULK_ROOT_MAXF
#define AFD_WPT_BULK_ROOT_MAXF 16
#endif
template <int F>
struct BulkRoot {
    static constexpr bool value = F <= AFD_WPT_BULK_ROOT_MAXF;
};
// cp.async staging (longer filters): place the frame so that global and shared addresses agree modulo 16 bytes (16-byte copies
// for every 8-byte aligned frame; without it coif4's root, 22 floats behind a 16-byte boundary, is staged in 8-byte pieces).
#ifndef AFD_WPT_CPASYNC_ALIGN
#define AFD_WPT_CPASYNC_ALIGN 1
#endif

template <int F, int R0, int RA, int RB, int RLA, int RLB, bool EXT>
__global__ void __launch_bounds__(kFrameThreads, 1)
wpt_frame_kernel(const float* __restrict__ x, long long x_row_stride, long long B, float* __restrict__ out,
                 const __grid_constant__ WptPlan plan, const __grid_constant__ Coefs<F> cf,
                 const __grid_constant__ Epilogue ep) {
    extern __shared__ __align__(16) float smem[];
    static_assert(ReflOk<F, R0, RA, RB, RLA, RLB>::value, "the frame kernel needs register reflection (item sizes >= F/2)");
    const int L = plan.L;
    const int tid = threadIdx.x;
    const int group = tid >> 8;                            // half tree under level-1 node 'a' (0) / 'd' (1)
    const int gt = tid & (kGroupThreads - 1);
    float* const root = smem + plan.region_b;              // padded frame: sample i at root[F - 2 + i]
    const int P = 1 << L;
    const int C = (ep.log_scale && ep.sign_channel) ? 2 : 1;
    const int T = plan.T;
    const int half_base = L >= 2 ? group << (L - 2) : 0;   // natural index of the half tree's first level-(L-1) node
    bool prefetched = false;
    bool first = true;
    // three words behind the planned regions: groups that have finished with region B (monotonic count), and per group
    // whether it was the second one to arrive
    const uint32_t root_bar = static_cast<uint32_t>(__cvta_generic_to_shared(smem + plan.smem_floats));    // mbarrier: the root's bulk copy
    unsigned int& s_arrivals = *reinterpret_cast<unsigned int*>(smem + plan.smem_floats + 2);
    int* const s_second = reinterpret_cast<int*>(smem + plan.smem_floats + 3);
    if (tid == 0) {
        s_arrivals = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(root_bar), "r"(2));       // one arrival per part of the root
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t root_parity = 0;
    ThreadStats ts;
    ts.s[0] = ts.s[1] = ts.q[0] = ts.q[1] = ts.m[0] = ts.m[1] = 0.f;
    ts.fs[0] = ts.fs[1] = ts.fq[0] = ts.fq[1] = 0.f;
    ts.col[0] = ts.col[1] = -1;
#if AFD_WPT_PHASE_TIMING
    long long phase_t0_ = clock64();
    const long long phase_start_ = phase_t0_;
#endif

    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const float* xg = x + b * x_row_stride;
        const long long nb = b + gridDim.x;
        constexpr bool kBulk = BulkRoot<F>::value;
        constexpr bool kShift = kBulk || AFD_WPT_CPASYNC_ALIGN;     // sample 0 at root[F - 2 + root_shift(frame)]
        if (!prefetched) {
            if (!first) __syncthreads();                   // region B may still be read by a group's last passes
            if constexpr (kBulk) { if (gt < 32) stage_root_bulk<F>(xg, root, plan.N, root_bar, gt, group, plan.root_split); }
            else issue_chunk<F>(xg, root + (kShift ? root_shift<F>(xg) : 0), 0, plan, kFrameThreads);
        }
        first = false;
        prefetched = false;
#if AFD_WPT_PHASE_TIMING
        const long long w0_ = clock64();
#endif
        if constexpr (kBulk) {
            mbar_wait_parity(root_bar, root_parity);       // the bulk copy has landed ...
            root_parity ^= 1u;
        } else {
            cp_async_wait<0>();
        }
#if AFD_WPT_PHASE_TIMING
        const long long w1_ = clock64();
#endif
        __syncthreads();                                   // ... and so have the staging warp's scalar copies
#if AFD_WPT_PHASE_TIMING
        if (gt == 0) {          // per group: wait for the copy, wait for the other group
            atomicAdd(&g_wpt_phase[16 + 4 * group], static_cast<unsigned long long>(w1_ - w0_));
            atomicAdd(&g_wpt_phase[17 + 4 * group], static_cast<unsigned long long>(clock64() - w1_));
        }
#endif
        const bool root_sh2 = kShift && root_shift<F>(xg) != 0;
        AFD_PHASE_MARK(0);
        float* out_b = out + b * C * static_cast<long long>(T) * P;
        // ---------------------------------------------------------------- passes
        for (int pi = 0; pi < plan.npass; ++pi) {
            const Pass& ps = plan.pass[pi];
            if (pi == 0) {
                // level 1: the root's two children, all 512 threads
                if (ps.kind == 0) {
                    if (kShift && root_sh2) mid_level_uniform<F, R0, true, kShift>(smem + ps.in_off + 2, smem + ps.out_off, ps, cf, tid, kFrameThreads);
                    else mid_level_uniform<F, R0, true, false>(smem + ps.in_off, smem + ps.out_off, ps, cf, tid, kFrameThreads);
                    __syncthreads();
                    mirror_copy<F>(smem + ps.out_off, 2, ps.n_out, ps.out_stride, tid, kFrameThreads);
                    __syncthreads();
                    // The barrier leaves both groups in phase: all 16 warps would load, then filter, then store together and
                    // the shared-memory pipe and the FMA pipe would take turns.  Group 1 idles for a fraction of a pass, so
                    // that from here on one group's loads / stores overlap the other's FMAs (like two independent CTAs).
                    if (group == 1 && plan.stagger > 0) {
                        const long long t0 = clock64();
                        while (clock64() - t0 < plan.stagger) { }
                    }
                } else {
                    // L == 1 is served by the two-CTA kernel (make_frame_plan refuses it)
                    last_level<F, RLA, true, EXT, true>(smem + ps.in_off, ps, T, 0, out_b, P, cf, ep, ts, tid, kFrameThreads);
                }
                AFD_PHASE_MARK(1);
                continue;
            }
            // levels >= 2: group g works on the nodes under level-1 node g
            const float* in = smem + ps.in_off + group * ps.in_group_off;
            if (ps.kind == 0) {
                float* o = smem + ps.out_off + group * ps.out_group_off;
                if (ps.rsel == 0) mid_level_uniform<F, RA, true>(in, o, ps, cf, gt, kGroupThreads);
                else mid_level_uniform<F, RB, true>(in, o, ps, cf, gt, kGroupThreads);
                group_barrier(group);
                mirror_copy<F>(o, 2 * ps.parents, ps.n_out, ps.out_stride, gt, kGroupThreads);
                group_barrier(group);
            } else {
                if (ps.prefetch) {
                    if constexpr (kBulk) {
                        // This group's half of region B (root + even levels) is free: it has produced level L-1 (closed by a
                        // group barrier).  It stages its part of the next frame without waiting for the other group.
                        if (nb < B) {
                            if (gt < 32) stage_root_bulk<F>(x + nb * x_row_stride, root, plan.N, root_bar, gt, group, plan.root_split);
                            prefetched = true;
                        }
                    } else {
                        // Region B is free once BOTH groups have produced level L-1.  No CTA-wide barrier (it would put the
                        // groups back in phase): the group that gets here second stages the next frame.
                        if (gt == 0) s_second[group] = static_cast<int>(atomicAdd(&s_arrivals, 1u) & 1u);
                        group_barrier(group);
                        if (nb < B) {
                            const float* xn = x + nb * x_row_stride;
                            if (s_second[group]) issue_chunk<F>(xn, root + (kShift ? root_shift<F>(xn) : 0), 0, plan, kGroupThreads, gt);
                            prefetched = true;
                        }
                    }
                }
                if (ps.rsel == 0) last_level<F, RLA, true, EXT, true>(in, ps, T, half_base, out_b, P, cf, ep, ts, gt, kGroupThreads);
                else last_level<F, RLB, true, EXT, true>(in, ps, T, half_base, out_b, P, cf, ep, ts, gt, kGroupThreads);
            }
            AFD_PHASE_MARK(1 + pi);
        }
#if AFD_WPT_PHASE_TIMING
        if (threadIdx.x == 0) atomicAdd(&g_wpt_phase[30], 1ull);
#endif
    }
#if AFD_WPT_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_wpt_phase[31], static_cast<unsigned long long>(clock64() - phase_start_));
#endif
    cp_async_wait<0>();
    if constexpr (EXT) {
        if (ep.node_stats) flush_node_stats(ep, P, ts);
        if (ep.feat_moments) flush_feat_moments(ep, C, ts);
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

struct Tuning {
    int R1, RA, RB, RLA, RLB;
    bool lat;
    double halo;      // extra pairs per item a lattice item pays: (J-1)/2; 0 for the direct form
};

// items * per-item cost in units of one output pair (lattice: R + (J-1)/2 rotations-columns; direct: R)
static double level_cost(int nodes, int n_out, int R, double halo) {
    const long long items = static_cast<long long>(nodes) * ((n_out + R - 1) / R);
    const long long rounds = (items + kThreads - 1) / kThreads;
    return static_cast<double>(rounds) * (R + halo);
}

// Builds the shared-memory plan and the pass list for `ctas_per_sm` resident CTAs.
static int make_plan(int64_t N, int F, int L, const Tuning& tu, int ctas_per_sm, double level_scale, WptPlan* p,
                     int shave_bytes = 0) {
    int n[kMaxLevel + 1], stride[kMaxLevel + 1];
    p->N = static_cast<int>(N);
    p->L = L;
    n[0] = static_cast<int>(N);
    for (int l = 1; l <= L; ++l) n[l] = (n[l - 1] + F - 1) / 2;
    for (int l = 0; l < L; ++l)
        if (n[l] < F - 1 + (n[l] & 1) || n[l] < 2)
            return fail(AFD_ERR_REFLECT_PAD,
                        "node length %d at level %d is not longer than the reflect padding of a %d-tap filter",
                        n[l], l, F);
    p->n1 = n[1];
    p->T = n[L];
    const int padl = F - 2;
    int minR = tu.RA;
    for (int r : {tu.RB, tu.RLA, tu.RLB}) minR = r < minR ? r : minR;
    const bool refl_ok = AFD_WPT_REFLECT_REGS != 0 && minR >= F / 2;      // == ReflOk<F, RA, RB, RLA, RLB> of the launched kernel
    const int limit_floats = ((ctas_per_sm == 2 ? (228 * 1024 / 2 - 1024) : kMaxSmemPerCta) - shave_bytes) / 4;
    int maxR = tu.R1;
    for (int r : {tu.RA, tu.RB, tu.RLA, tu.RLB}) maxR = r > maxR ? r : maxR;
    const int tail = round_up(2 * maxR + 1, 4);               // over-read slack behind the last node of a region:
                                                              // the last item's window ends <= 2R+1 floats past its node
    const int stored = L == 1 ? 1 : L - 1;                    // levels kept in shared memory
    const bool uniform = refl_ok && AFD_WPT_UNIFORM != 0;     // == UNI of the launched kernel
    for (int l = 1; l <= stored; ++l) {
        int s = round_up(n[l] + 2 * padl + (n[l] & 1), 4);
        if (uniform) {
            // The producer's last chunk stores all of its R coefficients: the node needs R-1 floats of tail space (>= the
            // right padding).  The left padding is never written and its only reader replaces what it read (ReflOk), so
            // it aliases the previous node's tail.
            const int r_prod = l == 1 ? tu.R1 : (tu.RA > tu.RB ? tu.RA : tu.RB);
            const int padr = padl + (n[l] & 1);
            const int tail_l = padr > r_prod - 1 ? padr : r_prod - 1;
            s = round_up(n[l] + (padl > tail_l ? padl : tail_l), 4);
        }
        if (l >= 2 && ((s >> 2) & 1) == 0) s += 4;            // lanes walk across nodes in edge / last-level items: odd 16-byte stride
        stride[l] = s;
    }
    p->stride1 = stride[1];
    // level-1 staging: balanced chunks of at most kThreads items
    const int n1 = n[1];
    p->nch = (n1 + kThreads * tu.R1 - 1) / (kThreads * tu.R1);
    p->kc = round_up((n1 + p->nch - 1) / p->nch, tu.R1);
    p->nch = (n1 + p->kc - 1) / p->kc;
    p->buf_floats = round_up(2 * p->kc + F + 8, 4);
    int G = 1;
    int need[2];
    for (;; G *= 2) {
        if (G > 4 || (G > 1 && (L < 4 || (1 << (L - 2)) / G < 2))) return AFD_ERR_UNSUPPORTED;   // caller retries / reports
        need[0] = 0; need[1] = 2 * p->buf_floats;             // [0] = region A (odd levels), [1] = region B
        for (int l = 1; l <= stored; ++l) {
            int nodes = 1 << (l - 1);
            if (l == L - 1 && G > 1) nodes /= G;
            const int fl = nodes * stride[l] + tail + (uniform ? padl : 0);   // aliased left padding: the first node's is extra
            int& r = need[(l & 1) ? 0 : 1];
            r = r > fl ? r : fl;
        }
        p->region_b = round_up(need[0], 4);
        p->smem_floats = p->region_b + round_up(need[1], 4);
        if (p->smem_floats <= limit_floats) break;
    }
    if (L == 1) { p->npass = 0; return AFD_OK; }
    // ---- pass list.  `pending`: true coefficient = pending * stored value (the unscaled lattice defers its cosines)
    auto region = [&](int l) { return (l & 1) ? 0 : p->region_b; };
    double pending = 1.0;
    int np = 0;
    auto pick = [&](int nodes, int n_out, int ra, int rb) {
        return level_cost(nodes, n_out, rb, tu.halo) < level_cost(nodes, n_out, ra, tu.halo) ? 1 : 0;
    };
    auto fill_split = [&](Pass& ps, int R) {
        const Split sp = make_split(ps.n_out, R, padl, refl_ok && AFD_WPT_REFL_INTERIOR);
        ps.C = sp.C; ps.CL = sp.CL; ps.CIe = sp.CIe; ps.CR = sp.CR; ps.NI = sp.NI; ps.NE = sp.NE; ps.magicNI = sp.magicNI;
        ps.magicC = magic_of(sp.C);
    };
    auto stored_mul = [&]() {                                 // factor for a stored level
        if (!tu.lat) return 1.0f;
        pending *= level_scale;
        if (fabs(pending) < 1e-9 || fabs(pending) > 1e9) { const float m = static_cast<float>(pending); pending = 1.0; return m; }
        return 1.0f;
    };
    const int last_full = G > 1 ? L - 2 : L - 1;
    for (int l = 2; l <= last_full; ++l) {
        Pass& ps = p->pass[np++];
        ps = Pass{};
        ps.kind = 0; ps.in_off = region(l - 1); ps.out_off = region(l);
        ps.parents = 1 << (l - 2); ps.lg_parents = l - 2; ps.n_out = n[l]; ps.n_in = n[l - 1];
        ps.in_stride = stride[l - 1]; ps.out_stride = stride[l];
        ps.rsel = pick(ps.parents, ps.n_out, tu.RA, tu.RB);
        ps.mul = stored_mul();
        fill_split(ps, ps.rsel ? tu.RB : tu.RA);
    }
    const int nodes_lm1 = 1 << (L - 2);                       // level L-1 nodes of this half tree
    const int per_group = nodes_lm1 / G;
    float grouped_mul = 1.0f;
    if (G > 1) grouped_mul = stored_mul();                    // same factor for every slice of level L-1
    const float last_mul = tu.lat ? static_cast<float>(pending * level_scale) : 1.0f;
    const bool can_prefetch = ((L - 1) & 1) == 1;             // region B is idle while the last level reads region A
    for (int g = 0; g < G; ++g) {
        if (G > 1) {
            Pass& ps = p->pass[np++];
            ps = Pass{};
            const int par = per_group / 2;                    // level L-2 parents of this slice
            ps.kind = 0; ps.in_off = region(L - 2) + g * par * stride[L - 2]; ps.out_off = region(L - 1);
            ps.parents = par; ps.lg_parents = ilog2(par); ps.n_out = n[L - 1]; ps.n_in = n[L - 2];
            ps.in_stride = stride[L - 2]; ps.out_stride = stride[L - 1];
            ps.rsel = pick(par, ps.n_out, tu.RA, tu.RB);
            ps.mul = grouped_mul;
            fill_split(ps, ps.rsel ? tu.RB : tu.RA);
            ps.sync_before = g > 0;                           // the level L-1 slice is being re-used
        }
        Pass& ps = p->pass[np++];
        ps = Pass{};
        ps.kind = 1; ps.in_off = region(L - 1);
        ps.parents = per_group; ps.lg_parents = ilog2(per_group); ps.n_out = n[L]; ps.n_in = n[L - 1];
        ps.in_stride = stride[L - 1];
        ps.rsel = pick(per_group, n[L], tu.RLA, tu.RLB);
        ps.mul = last_mul;
        ps.parent_base = g * per_group;
        ps.prefetch = (g == G - 1 && can_prefetch) ? 1 : 0;
    }
    p->npass = np;
    return AFD_OK;
}

// What afd_wpt_plan_info reports (host-side introspection of the launch configuration).
struct PlanReport {
    int smem_bytes, ctas_per_sm, lattice, passes, nch, kc;
    int pass_items[kMaxPasses];      // work items of each pass
    int pass_r[kMaxPasses];          // item size chosen for each pass
};

// Everything about a launch that depends only on (device, N, level, taps): planned once and reused.  A training / eval
// loop calls the transform with the same shape every step, and at small batches (BASELINE configs[0]: 128 frames, a 23 us
// kernel) the planning, the occupancy query and the attribute calls would otherwise cost several times the kernel.
template <int F>
struct LaunchCache {
    bool valid = false;
    int dev = -1, L = 0, ctas = 0, sms = 0, stats_simple = 0;
    int64_t N = 0;
    double taps[F];
    WptPlan plan;
    Coefs<F> cf;
};

template <int F, int R1, int RA, int RB, int RLA, int RLB, bool LAT, bool EXT>
static int launch(const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                  const double* dec_lo, const LatticeInfo& lat, const Epilogue& ep, cudaStream_t stream,
                  PlanReport* report) {
    auto kern = wpt_tree_kernel<F, R1, RA, RB, RLA, RLB, LAT, EXT>;
    static thread_local LaunchCache<F> cache;              // one per instantiation and host thread
    int dev = 0;
    if (!report) AFD_CUDA_TRY(cudaGetDevice(&dev));
    if (report || !cache.valid || cache.dev != dev || cache.N != N || cache.L != L ||
        memcmp(cache.taps, dec_lo, sizeof(double) * F) != 0) {
        cache.valid = false;
        WptPlan& plan = cache.plan;
        Tuning tu{R1, RA, RB, RLA, RLB, LAT, LAT ? (F / 2 - 1) * 0.5 : 0.0};
        int ctas = 2;
        int rc = make_plan(N, F, L, tu, 2, lat.scale, &plan);
        if (rc == AFD_ERR_UNSUPPORTED) {
            ctas = 1;
            rc = make_plan(N, F, L, tu, 1, lat.scale, &plan);
            if (rc == AFD_ERR_UNSUPPORTED)
                return fail(AFD_ERR_UNSUPPORTED,
                            "wavelet-packet tree (N=%lld, F=%d, level=%d) needs %lld bytes of shared memory per CTA, limit %d",
                            static_cast<long long>(N), F, L, 4LL * plan.smem_floats, kMaxSmemPerCta);
        }
        if (rc != AFD_OK) return rc;
        if (report) {
            report->smem_bytes = 4 * plan.smem_floats; report->ctas_per_sm = ctas; report->lattice = LAT ? 1 : 0;
            report->passes = plan.npass; report->nch = plan.nch; report->kc = plan.kc;
            for (int i = 0; i < plan.npass; ++i) {
                const Pass& ps = plan.pass[i];
                const int r = ps.kind == 0 ? (ps.rsel ? RB : RA) : (ps.rsel ? RLB : RLA);
                report->pass_r[i] = r;
                report->pass_items[i] = ps.parents * ((ps.n_out + r - 1) / r);
            }
            return AFD_OK;
        }
        Coefs<F>& cf = cache.cf;
        for (int k = 0; k < F; ++k) {
            cf.lo[k] = static_cast<float>(dec_lo[k]);
            cf.hi[k] = static_cast<float>(((k & 1) ? 1.0 : -1.0) * dec_lo[F - 1 - k]);   // dec_hi[k] = (-1)^(k+1) dec_lo[F-1-k]
        }
        for (int m = 0; m < F / 2; ++m) cf.t[m] = LAT ? static_cast<float>(lat.tan_theta[m]) : 0.f;
        static thread_local bool configured[16] = {false};  // per device
        int sms = kNumSmsFallback;
        if (dev >= 16 || !configured[dev]) {
            AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));
            AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
            if (dev < 16) configured[dev] = true;
        }
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (ctas == 2) {
            // the plan may sit exactly on the two-CTAs-per-SM boundary: if the driver disagrees, re-plan with head-room
            int resident = 0;
            AFD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kThreads, 4 * plan.smem_floats));
            if (resident < 2 && make_plan(N, F, L, tu, 2, lat.scale, &plan, 4096) != AFD_OK) {
                ctas = 1;
                rc = make_plan(N, F, L, tu, 1, lat.scale, &plan);
                if (rc != AFD_OK) return rc;
            }
        }
        {   // derived from the plan that is actually launched (the occupancy fallback above may have re-planned)
            int last_passes = 0, parents = 0;
            for (int i = 0; i < plan.npass; ++i)
                if (plan.pass[i].kind == 1) { ++last_passes; parents = plan.pass[i].parents; }
            cache.stats_simple = (last_passes == 1 && parents <= kThreads && kThreads % parents == 0) ? 1 : 0;
        }
        cache.dev = dev; cache.N = N; cache.L = L; cache.ctas = ctas; cache.sms = sms;
        memcpy(cache.taps, dec_lo, sizeof(double) * F);
        cache.valid = true;
    }
    Epilogue epk = ep;
    epk.stats_simple = cache.stats_simple;
    const int smem = 4 * cache.plan.smem_floats;
    long long grid = 2LL * cache.sms * cache.ctas / 2 * 2;             // persistent: every resident slot, even count
    if (cache.ctas == 1) grid = cache.sms / 2 * 2;
    if (grid > 2 * B) grid = 2 * B;
    kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(x, static_cast<long long>(x_row_stride),
                                                                   static_cast<long long>(B), out, cache.plan, cache.cf, epk);
    AFD_CUDA_TRY(cudaGetLastError());
#if AFD_WPT_PHASE_TIMING
    {
        unsigned long long h[32];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_wpt_phase, sizeof(h));
        const double half_frames = static_cast<double>(h[30] ? h[30] : 1);
        fprintf(stderr, "wpt phases F=%d B=%lld (thread-0 cycles per half frame): ", F, static_cast<long long>(B));
        for (int i = 0; i < 3 + cache.plan.npass; ++i) fprintf(stderr, "%s%.0f", i ? " " : "", h[i] / half_frames);
        fprintf(stderr, " | total %.0f\n", h[31] / half_frames);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_wpt_phase, h, sizeof(h));
    }
#endif
    return AFD_OK;
}


// Plan of the frame kernel: full tree, level l = 2^l nodes (the nodes of half tree g are the contiguous half g), root =
// the padded frame at the start of region B.  pass[0] is level 1 (all threads); pass[l-1], l >= 2, carries the PER-GROUP
// parent count and group-0 offsets.  Returns AFD_ERR_UNSUPPORTED when the tree does not fit one CTA.
static int make_frame_plan(int64_t N, int F, int L, int R0, const Tuning& tu, double level_scale, WptPlan* p,
                           bool tune_last_stride = true) {
    int n[kMaxLevel + 1], stride[kMaxLevel + 1];
    p->N = static_cast<int>(N);
    p->L = L;
    n[0] = static_cast<int>(N);
    for (int l = 1; l <= L; ++l) n[l] = (n[l - 1] + F - 1) / 2;
    for (int l = 0; l < L; ++l)
        if (n[l] < F - 1 + (n[l] & 1) || n[l] < 2)
            return fail(AFD_ERR_REFLECT_PAD,
                        "node length %d at level %d is not longer than the reflect padding of a %d-tap filter",
                        n[l], l, F);
    p->n1 = n[1];
    p->T = n[L];
    const int padl = F - 2;
    int maxR = R0;
    for (int r : {tu.RA, tu.RB, tu.RLA, tu.RLB}) maxR = r > maxR ? r : maxR;
    const int tail = round_up(2 * maxR + 1, 4);
    // root: staged by issue_chunk(j = 0) with kc = n1: samples 2 - F .. 2 n1 - 1, over-read by the last item's window
    p->kc = n[1];
    p->nch = 1;
    const int c1 = (n[1] + R0 - 1) / R0;
    if (c1 > kFrameThreads) return AFD_ERR_UNSUPPORTED;
    if (L < 2) return AFD_ERR_UNSUPPORTED;                   // a single level: the two-CTA kernel's special case
    p->buf_floats = round_up(2 * c1 * R0 + F + 8 + 4, 4);      // + the root's alignment shift
    stride[0] = p->buf_floats;
    for (int l = 1; l < L; ++l) {
        const int r_prod = l == 1 ? R0 : (tu.RA > tu.RB ? tu.RA : tu.RB);
        const int padr = padl + (n[l] & 1);
        const int tail_l = padr > r_prod - 1 ? padr : r_prod - 1;
        int s = round_up(n[l] + (padl > tail_l ? padl : tail_l), 4);
        if (l == L - 1) {
            if (l >= 2 && ((s >> 2) & 1) == 0) s += 4;        // lanes walk across nodes in the last level: odd 16-byte stride
            if (l >= 2 && tune_last_stride) {
                // ... and of the four residues that leaves modulo 32 floats, the one whose 64-bit stores (the producing pass:
                // lanes R floats apart along a node, C items per node) need the fewest shared-memory wavefronts
                const int parents_p = 1 << (l - 2);
                const int rp = level_cost(parents_p, n[l], tu.RB, tu.halo) < level_cost(parents_p, n[l], tu.RA, tu.halo) ? tu.RB : tu.RA;
                const int cp = (n[l] + rp - 1) / rp;
                int best = s, best_w = 1 << 30;
                for (int cand = s; cand < s + 32; cand += 8) {
                    int waves = 0;
                    for (int h0 = 0; h0 < parents_p * cp; h0 += 16) {         // one wavefront serves 16 lanes x 8 bytes
                        int cnt[16] = {0}, worst = 0;
                        for (int it = h0; it < h0 + 16 && it < parents_p * cp; ++it) {
                            const int b = ((it / cp) * (cand / 2) + (it % cp) * (rp / 2)) & 15;
                            worst = ++cnt[b] > worst ? cnt[b] : worst;
                        }
                        waves += worst;
                    }
                    if (waves < best_w) { best_w = waves; best = cand; }
                }
                s = best;
            }
        } else if (AFD_WPT_STRIDE_CONGRUENCE && l >= 2) {
            // The consumer (level l + 1) walks its lanes along a node, 2R floats apart (conflict-free for R/2 odd), and jumps
            // to the next node after C items: with stride == C * 2R (mod 32 floats) the jump continues the same progression.
            const int parents_c = 1 << (l - 1);                                          // per group, as consumer parents
            const int rc = level_cost(parents_c, n[l + 1], tu.RB, tu.halo) < level_cost(parents_c, n[l + 1], tu.RA, tu.halo) ? tu.RB : tu.RA;
            const int cc = (n[l + 1] + rc - 1) / rc;
            const int want = (cc * 2 * rc) & 31;
            while ((s & 31) != want) s += 4;
        }
        stride[l] = s;
    }
    // Each group keeps its half tree in its own half of a region: the groups run out of phase, so one group may write
    // level l + 2 while the other still reads level l of the same region, and the per-level layouts do not nest.
    int half[2] = {0, 0};                                     // [0] = region A (odd levels), [1] = region B (even levels)
    for (int l = 1; l < L; ++l) {
        const int fl = (1 << (l - 1)) * stride[l] + tail + padl;
        int& r = half[(l & 1) ? 0 : 1];
        r = r > fl ? r : fl;
    }
    half[0] = round_up(half[0], 4);
    half[1] = round_up(half[1], 4);
    if (L > 1) stride[1] = half[0];                           // the two level-1 nodes are one group half apart
    p->stride1 = L > 1 ? stride[1] : 0;
    const int root_floats = p->buf_floats + tail;
    p->region_b = 2 * half[0];
    p->root_split = half[1];
    p->smem_floats = p->region_b + (2 * half[1] > root_floats ? 2 * half[1] : round_up(root_floats, 4));
    if (4LL * (p->smem_floats + 8) > kMaxSmemPerCta) return AFD_ERR_UNSUPPORTED;      // + the kernel's mbarrier and sync words
    auto region = [&](int l) { return (l & 1) ? 0 : p->region_b; };
    double pending = 1.0;
    auto stored_mul = [&]() {
        pending *= level_scale;
        if (fabs(pending) < 1e-9 || fabs(pending) > 1e9) { const float m = static_cast<float>(pending); pending = 1.0; return m; }
        return 1.0f;
    };
    auto pick = [&](int nodes, int n_out, int ra, int rb) {
        return level_cost(nodes, n_out, rb, tu.halo) < level_cost(nodes, n_out, ra, tu.halo) ? 1 : 0;
    };
    int np = 0;
    for (int l = 1; l < L; ++l) {
        Pass& ps = p->pass[np++];
        ps = Pass{};
        ps.kind = 0; ps.in_off = region(l - 1); ps.out_off = region(l);
        ps.parents = l == 1 ? 1 : 1 << (l - 2);               // per group from level 2 on
        ps.lg_parents = l == 1 ? 0 : l - 2;
        ps.n_out = n[l]; ps.n_in = n[l - 1];
        ps.in_stride = stride[l - 1]; ps.out_stride = stride[l];
        ps.in_group_off = l == 1 ? 0 : half[((l - 1) & 1) ? 0 : 1];
        ps.out_group_off = half[(l & 1) ? 0 : 1];
        ps.rsel = l == 1 ? 0 : pick(ps.parents, ps.n_out, tu.RA, tu.RB);
        const int R = l == 1 ? R0 : (ps.rsel ? tu.RB : tu.RA);
        ps.mul = stored_mul();
        ps.C = (ps.n_out + R - 1) / R;
        ps.magicC = magic_of(ps.C);
        if (l >= 2 && ps.parents * ps.C > 4 * kGroupThreads) return AFD_ERR_UNSUPPORTED;   // degenerate shapes: old kernel
    }
    Pass& ps = p->pass[np++];
    ps = Pass{};
    ps.kind = 1; ps.in_off = region(L - 1);
    ps.parents = L == 1 ? 1 : 1 << (L - 2); ps.lg_parents = L == 1 ? 0 : L - 2;
    ps.n_out = n[L]; ps.n_in = n[L - 1];
    ps.in_stride = stride[L - 1];
    ps.in_group_off = L == 1 ? 0 : half[((L - 1) & 1) ? 0 : 1];
    ps.rsel = L == 1 ? 0 : pick(ps.parents, n[L], tu.RLA, tu.RLB);
    ps.mul = static_cast<float>(pending * level_scale);
    ps.parent_base = 0;
    ps.prefetch = ((L - 1) & 1) == 1 ? 1 : 0;                 // region B (root) is idle while the last level reads region A
    p->npass = np;
    return AFD_OK;
}

template <int F, int R0, int RA, int RB, int RLA, int RLB, bool EXT>
static int launch_frame(const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                        const double* dec_lo, const LatticeInfo& lat, const Epilogue& ep, cudaStream_t stream,
                        PlanReport* report, bool* taken) {
    auto kern = wpt_frame_kernel<F, R0, RA, RB, RLA, RLB, EXT>;
    static thread_local LaunchCache<F> cache;
    static thread_local int cache_ok = 0;                  // 1: plan fits, -1: shape not served by this kernel
    int dev = 0;
    *taken = false;
    if (!report) AFD_CUDA_TRY(cudaGetDevice(&dev));
    if (report || !cache.valid || cache.dev != dev || cache.N != N || cache.L != L ||
        memcmp(cache.taps, dec_lo, sizeof(double) * F) != 0) {
        cache.valid = false;
        Tuning tu{R0, RA, RB, RLA, RLB, true, (F / 2 - 1) * 0.5};
        int rc = make_frame_plan(N, F, L, R0, tu, lat.scale, &cache.plan, true);
        if (rc == AFD_ERR_UNSUPPORTED) rc = make_frame_plan(N, F, L, R0, tu, lat.scale, &cache.plan, false);   // the tuned stride may not fit
        if (rc == AFD_ERR_UNSUPPORTED) cache_ok = -1;
        else if (rc != AFD_OK) return rc;
        else cache_ok = 1;
        if (report) {
            if (cache_ok < 0) return AFD_OK;
            const WptPlan& plan = cache.plan;
            report->smem_bytes = 4 * plan.smem_floats; report->ctas_per_sm = 1; report->lattice = 1;
            report->passes = plan.npass; report->nch = 1; report->kc = plan.kc;
            for (int i = 0; i < plan.npass; ++i) {
                const Pass& ps = plan.pass[i];
                const int r = i == 0 && ps.kind == 0 ? R0 : (ps.kind == 0 ? (ps.rsel ? RB : RA) : (ps.rsel ? RLB : RLA));
                report->pass_r[i] = r;
                report->pass_items[i] = ps.parents * ((ps.n_out + r - 1) / r);
            }
            *taken = true;
            return AFD_OK;
        }
        if (cache_ok > 0) {
            const char* stg = getenv("AFD_WPT_STAGGER");  // tuning knob (cycles); default measured on B200
            // sweeps (tools/gpu_stagger_sweep.sh, B = 4096, level 8): db2 / db3 best at 1000, sym4 / sym5 / coif2 at 1200, db7 / sym8 at
            // 1600 cycles; the filters staged with cp.async (F > 16) at 0 - 400
            cache.plan.stagger = stg ? atoi(stg)
                                     : (F <= 6 ? AFD_WPT_STAGGER_DEFAULT * 5 / 6
                                               : F <= 12 ? AFD_WPT_STAGGER_DEFAULT : F <= 16 ? AFD_WPT_STAGGER_DEFAULT * 4 / 3 : AFD_WPT_STAGGER_DEFAULT / 3);
            Coefs<F>& cf = cache.cf;
            for (int k = 0; k < F; ++k) {
                cf.lo[k] = static_cast<float>(dec_lo[k]);
                cf.hi[k] = static_cast<float>(((k & 1) ? 1.0 : -1.0) * dec_lo[F - 1 - k]);
            }
            for (int m = 0; m < F / 2; ++m) cf.t[m] = static_cast<float>(lat.tan_theta[m]);
            static thread_local bool configured[16] = {false};
            if (dev >= 16 || !configured[dev]) {
                AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));
                AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                  cudaSharedmemCarveoutMaxShared));
                if (dev < 16) configured[dev] = true;
            }
            int sms = kNumSmsFallback;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int parents = cache.plan.pass[cache.plan.npass - 1].parents;
            cache.stats_simple = (L >= 2 && parents <= kGroupThreads && kGroupThreads % parents == 0) ? 1 : 0;
            cache.sms = sms;
        }
        cache.dev = dev; cache.N = N; cache.L = L;
        memcpy(cache.taps, dec_lo, sizeof(double) * F);
        cache.valid = true;
    }
    if (cache_ok < 0) return AFD_OK;                       // not taken: the caller falls back to the two-CTA kernel
    // a root staged by a bulk copy: every frame must start on an 8-byte boundary (16-byte units, 0 or 2 floats of shift)
    if (BulkRoot<F>::value && ((reinterpret_cast<uintptr_t>(x) & 7) != 0 || (x_row_stride & 1) != 0)) return AFD_OK;
    *taken = true;
    Epilogue epk = ep;
    epk.stats_simple = cache.stats_simple;
    long long grid = cache.sms;
    if (grid > B) grid = B;
    kern<<<static_cast<unsigned>(grid), kFrameThreads, 4 * (cache.plan.smem_floats + 8), stream>>>(
        x, static_cast<long long>(x_row_stride), static_cast<long long>(B), out, cache.plan, cache.cf, epk);
    AFD_CUDA_TRY(cudaGetLastError());
#if AFD_WPT_PHASE_TIMING
    {
        unsigned long long h[32];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_wpt_phase, sizeof(h));
        const double frames = static_cast<double>(h[30] ? h[30] : 1);
        fprintf(stderr, "wpt frame-kernel phases F=%d B=%lld (thread-0 cycles per frame: wait, level 1, level 2, ...): ", F,
                static_cast<long long>(B));
        for (int i = 0; i < 1 + cache.plan.npass; ++i) fprintf(stderr, "%s%.0f", i ? " " : "", h[i] / frames);
        fprintf(stderr, " | total %.0f | group 0: copy wait %.0f, barrier %.0f; group 1: copy wait %.0f, barrier %.0f\n", h[31] / frames,
                h[16] / frames, h[17] / frames, h[20] / frames, h[21] / frames);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_wpt_phase, h, sizeof(h));
    }
#endif
    return AFD_OK;
}

// Item sizes.  R/2 odd keeps the 128-bit window loads and 64-bit stores of consecutive lanes conflict-free; larger
// R amortises the F-2 halo (and, for the lattice, its (J-1)/2 extra rotation columns), smaller R bounds the
// register window of 2R+F-2 floats.  Two sizes per kernel let the plan fill whole rounds of 256 threads.
constexpr int pick_r1(int F) { return F <= 24 ? 14 : (F <= 40 ? 10 : 6); }
constexpr int pick_rd(int F) { return F <= 40 ? 10 : 6; }           // direct form, F > 32
// second last-level item size: a quarter of the leaf length of the headline shape (N = 22050, level 8: T ~ 85 + F)
#ifndef AFD_WPT_RA
#define AFD_WPT_RA 22
#define AFD_WPT_RB 26
#endif
#ifndef AFD_WPT_RLA
#define AFD_WPT_RLA 14
#endif
#ifndef AFD_WPT_R0
#define AFD_WPT_R0 22             // level-1 item size of the frame kernel: 11029 / 22 = 502 items = one round of 512 threads
#endif
#ifndef AFD_WPT_RLB_CHUNKS
#define AFD_WPT_RLB_CHUNKS 4      // chunks per leaf of the headline shape: 64 parents x 4 chunks = 256 items
#endif
constexpr int pick_rlb(int F) { return 2 * ((85 + F + 2 * AFD_WPT_RLB_CHUNKS - 1) / (2 * AFD_WPT_RLB_CHUNKS)); }

template <int F, bool EXT>
static int dispatch_one(const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                        const double* dec_lo, const Epilogue& ep, cudaStream_t stream, PlanReport* report) {
    LatticeInfo lat{};
    lat.scale = 1.0;
    if constexpr (F <= 32) {
        static thread_local double seen[F];                    // the factorisation of the last taps seen by this thread
        static thread_local LatticeInfo seen_lat{};
        static thread_local int seen_rc = -1;
        if (seen_rc < 0 || memcmp(seen, dec_lo, sizeof(double) * F) != 0) {
            seen_lat = LatticeInfo{};
            seen_lat.scale = 1.0;
            seen_rc = lattice_factor(dec_lo, F, &seen_lat) == AFD_OK ? 1 : 0;
            memcpy(seen, dec_lo, sizeof(double) * F);
        }
        lat = seen_lat;
        if (seen_rc == 1 && lat.usable) {
            if constexpr (AFD_WPT_FRAME_KERNEL != 0 && AFD_WPT_UNIFORM != 0 && AFD_WPT_THREADS == 256 &&
                          ReflOk<F, AFD_WPT_R0, AFD_WPT_RA, AFD_WPT_RB, AFD_WPT_RLA, pick_rlb(F)>::value) {
                bool taken = false;
                const int rc = launch_frame<F, AFD_WPT_R0, AFD_WPT_RA, AFD_WPT_RB, AFD_WPT_RLA, pick_rlb(F), EXT>(
                    x, B, N, x_row_stride, out, L, dec_lo, lat, ep, stream, report, &taken);
                if (rc != AFD_OK || taken) return rc;
            }
            return launch<F, pick_r1(F), AFD_WPT_RA, AFD_WPT_RB, AFD_WPT_RLA, pick_rlb(F), true, EXT>(x, B, N, x_row_stride, out, L, dec_lo, lat, ep, stream, report);
        }
        return launch<F, pick_r1(F), 14, 10, 14, 10, false, EXT>(x, B, N, x_row_stride, out, L, dec_lo, lat, ep, stream, report);
    } else {
        return launch<F, pick_r1(F), pick_rd(F), pick_rd(F), pick_rd(F), pick_rd(F), false, EXT>(
            x, B, N, x_row_stride, out, L, dec_lo, lat, ep, stream, report);
    }
}


// Filter lengths are compiled in four groups, each for the plain (afd_wpt_forward) and the extended
// (afd_wpt_forward_ex) epilogue: eight translation units, built in parallel.
using WptGroupFn = int (*)(int F, const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                           const double* dec_lo, const Epilogue& ep, cudaStream_t stream, PlanReport* report);
#define AFD_WPT_GROUP_DECL(NAME)                                                                                     \
    int NAME(int F, const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,                   \
             const double* dec_lo, const Epilogue& ep, cudaStream_t stream, PlanReport* report);
AFD_WPT_GROUP_DECL(wpt_group0)     // F = 2 .. 16
AFD_WPT_GROUP_DECL(wpt_group1)     // F = 18 .. 32
AFD_WPT_GROUP_DECL(wpt_group2)     // F = 34 .. 48
AFD_WPT_GROUP_DECL(wpt_group3)     // F = 50 .. 64
AFD_WPT_GROUP_DECL(wpt_xgroup0)    // the same with the extended epilogue
AFD_WPT_GROUP_DECL(wpt_xgroup1)
AFD_WPT_GROUP_DECL(wpt_xgroup2)
AFD_WPT_GROUP_DECL(wpt_xgroup3)

#define AFD_WPT_GROUP(NAME, F0, EXT)                                                                                 \
    int NAME(int F, const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,                   \
             const double* dec_lo, const Epilogue& ep, cudaStream_t stream, PlanReport* report) {                    \
        switch (F) {                                                                                                 \
            case F0: return dispatch_one<F0, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report);        \
            case F0 + 2: return dispatch_one<F0 + 2, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
            case F0 + 4: return dispatch_one<F0 + 4, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
            case F0 + 6: return dispatch_one<F0 + 6, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
            case F0 + 8: return dispatch_one<F0 + 8, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
            case F0 + 10: return dispatch_one<F0 + 10, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
            case F0 + 12: return dispatch_one<F0 + 12, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
            case F0 + 14: return dispatch_one<F0 + 14, EXT>(x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report); \
        }                                                                                                            \
        return fail(AFD_ERR_INVALID_ARG, "unsupported filter length %d", F);                                         \
    }

}  // namespace afd
