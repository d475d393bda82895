// STFT power spectrogram for n_fft = 511 = 7 * 73 on sm_100a, tcgen05 version: the prime-factor real DFT of
// afd_stft_pfa.cu with the 73-point stage on the 5th-generation tensor cores (tcgen05.mma kind::tf32, operands in
// shared memory, accumulators in tensor memory) and a warp-specialised persistent CTA around it.
//
// Replaces torchaudio.transforms.Spectrogram(511, hop, power) -> torch.stft(center=True, pad_mode="reflect",
// window=hann_window(511) [periodic], onesided=True) -> abs().pow(power) and the optional log(spec + 1e-12)
// of the reference (wavelet_math.py:47,63-66).  Algorithm (fold, DFT-7, DFT-73 as two real GEMMs against cos / -sin,
// combine): see the header of afd_stft_pfa.cu; the arithmetic outside the GEMM is the same.
//
// Mapping.  A unit = 16 consecutive rows of the flattened (signal, frame) index, i.e. 16 consecutive output rows; it may
// straddle two signals (two staged sample segments), so only the last unit of a launch is partial.  Its seven real sequences (u0, u1, v1, u2, v2, u3, v3)
// are the rows of ONE 128-row MMA tile: row = 32 j + 16 h + f  (j = 0..3, h = 0: u_j / 1: v_j, f = frame; rows
// 16..31 stay zero), columns = [P (m = 0..36, padded to 40) | Q (m = 1..36 at 41..76, padded to 80)].
//     D[:, 0:48]  = A[:, 0:40]  x C      (Re part, C[m][k2] =  cos(2 pi m k2 / 73), N padded 37 -> 48)
//     D[:, 48:96] = A[:, 40:80] x S      (Im part, S[m][k2] = -sin(2 pi m k2 / 73))
// Each product is 3xTF32 error compensated: A and the tables are split into a TF32-exact high part and an fp32 remainder
// (two shared-memory copies each) and D accumulates lo*hi + hi*lo + hi*hi in fp32 -> fp32-level accuracy.  30
// tcgen05.mma (M = 128, N = 48, K = 8) per unit, issued by one thread.
// Shared-memory operand layout: the canonical K-major no-swizzle layout (8-row x 16-byte core matrices); consecutive
// 16-byte K chunks of the A tile are 144 bytes apart so that the 37 lanes that write one row segment hit distinct banks.
//
// Roles of the 16 warps of the one CTA per SM (each role walks the same unit sequence):
//     warps 8..14  producers : stage the unit's samples (cp.async, next unit prefetched), window, fold, two 7-point real
//                              DFTs per (frame, m), hi / lo split, store into the A tiles, arrive on `a_full`
//     warp 18      MMA issuer: waits `a_full` (+ `d_empty` of the accumulator buffer), issues the 30 MMAs, commits to
//                              `a_free` (producers may overwrite A) and `d_full` (accumulators ready)
//     warps 0..7   epilogue  : warp w reads TMEM lanes 32 (w % 4) .. +31 (= j), half w / 4 of the k2 range;
//                              tcgen05.ld, partner exchange (u <-> v sit 16 lanes apart), |Z|^2, log -> 16 x 256 tile in
//                              shared memory (double buffered) -> 128-bit coalesced streaming stores.
// Accumulators are double buffered in TMEM (2 x 96 of 256 allocated columns), so unit n's epilogue overlaps unit
// n+1's producer phase; the tensor phase itself is ~0.5 us per unit.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "afd_common.cuh"

namespace afd {

namespace tc {

constexpr int kN = 511;
constexpr int kRows = 16;                     // STFT frames per unit
constexpr int kEpiWarps = 8, kPreWarps = 10;
constexpr int kEpiThreads = kEpiWarps * 32;   // 256
constexpr int kPreThreads = kPreWarps * 32;   // 320: two (frame, m) items per thread (frames r, r + 8), 296 used
constexpr int kThreads = kEpiThreads + kPreThreads + 32;   // + MMA warp = 608

// A tile (bytes): 128 rows x 80 columns of fp32, K-major core matrices
constexpr int kALbo = 144;                    // distance of consecutive 16-byte K chunks
constexpr int kASbo = 20 * kALbo;             // distance of consecutive 8-row groups (2880)
constexpr int kATile = 16 * kASbo;            // 46080
// B tiles (bytes): 48 rows (k2) x 40 columns (m)
constexpr int kBLbo = 128;
constexpr int kBSbo = 10 * kBLbo;             // 1280
constexpr int kBTile = 6 * kBSbo;             // 7680
constexpr int kRawFloats = 4432;              // two segments: >= 14 * hop + 2 * (511 + 6)  (hop <= 242); one: 15 * hop + 517
constexpr int kMaxHop = (kRawFloats - 2 * (kN + 6)) / (kRows - 2);
constexpr int kRawBufs = 4;                   // staging depth: the bulk loads run kRawBufs units ahead of the producers
constexpr int kOutStride = 260;               // floats per row of the output tile
constexpr int kOutFloats = kRows * kOutStride;
constexpr int kWFloats = 7 * 74;               // window, [n1][m] for the (n1, m) samples then [n1][37 + m] for (n1, 73 - m)

// shared-memory map (bytes)
constexpr int kOffAHi = 0;
constexpr int kOffALo = kOffAHi + kATile;
constexpr int kOffB = kOffALo + kATile;                      // C hi, C lo, S hi, S lo
constexpr int kOffRaw = kOffB + 4 * kBTile;                  // two staging buffers
constexpr int kOffOut = kOffRaw + kRawBufs * kRawFloats * 4;  // two output tiles
constexpr int kOffW = kOffOut + 2 * kOutFloats * 4;
constexpr int kOffBar = kOffW + kWFloats * 4;                // mbarriers + TMEM base address
constexpr int kSmemBytes = kOffBar + 128 + 8 * kRawBufs;
constexpr int kTableFloats = 4 * kBTile / 4 + kWFloats;      // device table: B tiles (byte-exact smem image) + window

constexpr int kTmemCols = 512;
constexpr int kDStride = 256;                 // TMEM columns between the two accumulator buffers (2 parts x 96 used)
constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);   // f32 accumulate, tf32 x tf32, K-major, M = 128
constexpr uint32_t kIdesc48 = kIdescBase | ((48u >> 3) << 17), kIdesc96 = kIdescBase | ((96u >> 3) << 17);

struct Params {
    int hop, frames, N, pad, B;
    long long total_units, total_rows;
    float power, log_offset;
    int log_scale, square;
    int normalize, store;
    float nmean, nrstd;
    double* moments;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");   // suspend-time hint (ns): sleep, do not spin
        if (spin > (1u << 20)) __trap();       // a lost arrival must not hang the GPU
    }
}
// Knock-out timing builds (tools/stft_tc_knockout.py; results are wrong by construction, only the times matter):
// 1 two of the 20 MMAs, 2 no epilogue math / tile stores, 3 no producer loads / DFT-7, 4 no A-tile stores, 5 no proxy fences,
// 7 all of 1-4 (pipeline skeleton), 8 = 7 without the output row stores.
#ifndef AFD_TC_KO
#define AFD_TC_KO 0
#endif
#define AFD_KO(n) (AFD_TC_KO == (n) || ((n) <= 4 && AFD_TC_KO >= 7))
#ifdef AFD_TC_PROF
#define PROF_DECL() long long _pacc[6] = {0, 0, 0, 0, 0, 0}; long long _pt0 = 0
#define PROF_START() _pt0 = clock64()
#define PROF_LAP(i) do { const long long _n = clock64(); _pacc[i] += _n - _pt0; _pt0 = _n; } while (0)
#define PROF_PRINT(cond, what)                                                                                      \
    if (blockIdx.x == 0 && (cond) && it > 0)                                                                        \
        printf("tid %d units %d %s: %lld %lld %lld %lld %lld %lld\n", tid, it, what, _pacc[0] / it, _pacc[1] / it,   \
               _pacc[2] / it, _pacc[3] / it, _pacc[4] / it, _pacc[5] / it)
#else
#define PROF_DECL()
#define PROF_START()
#define PROF_LAP(i)
#define PROF_PRINT(cond, what)
#endif
// Whole-warp wait: lane 0 polls, the other lanes park at the warp barrier (no issue slots burnt by 31 spinning lanes).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle, version 1 (sm_100) shared-memory matrix descriptor
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// Walks the CTA's unit sequence (unit = first, first + stride, ...): (signal b, frame t) of the unit's first row, without
// divisions in the loop.  A unit holds n0 = min(16, frames - t) rows of signal b and, when b + 1 exists, 16 - n0 rows of b + 1.
struct UnitWalk {
    int b, t;
    int db, dt, frames;
    __device__ __forceinline__ UnitWalk(int first, int stride, int frames_) {
        frames = frames_;
        const long long G = static_cast<long long>(first) * kRows, D = static_cast<long long>(stride) * kRows;
        b = static_cast<int>(G / frames); t = static_cast<int>(G - static_cast<long long>(b) * frames);
        db = static_cast<int>(D / frames); dt = static_cast<int>(D - static_cast<long long>(db) * frames);
    }
    __device__ __forceinline__ void next() {
        b += db; t += dt;
        if (t >= frames) { t -= frames; ++b; }
    }
    __device__ __forceinline__ UnitWalk peek() const { UnitWalk n = *this; n.next(); return n; }
    __device__ __forceinline__ int n0() const { return min(kRows, frames - t); }
    __device__ __forceinline__ int n1(int B) const { return b + 1 < B ? kRows - n0() : 0; }
};

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

// Staging of a unit's samples: raw[i] = x~[s_begin + i], i in [0, 4 nq), x~ = reflect extension, s_begin = S0 - aoff
// (S0 = t0 * hop - pad, aoff = misalignment of x~[S0] against 16 bytes; g0 = absolute address of the row in elements).  The 16-byte aligned in-range middle part
// [lo4, hi4) arrives as ONE bulk copy (TMA engine, full-line shared-memory writes, issued by the MMA thread); the few
// remaining samples (alignment slack, reflect padding of the first / last unit of a signal) are written by the producers.
struct StageGeom {
    int s_begin, lo4, hi4, end;        // sample indices: buffer start, bulk range, buffer end (exclusive)
};
__device__ __forceinline__ StageGeom stage_geom(long long g0, int t0, int valid, const Params& p) {
    StageGeom g;
    const int S0 = t0 * p.hop - p.pad;
    const int len = (valid - 1) * p.hop + kN;
    const int aoff = static_cast<int>((g0 + S0) & 3);
    g.s_begin = S0 - aoff;
    g.end = g.s_begin + ((aoff + len + 3) & ~3);
    g.lo4 = g.s_begin < 0 ? g.s_begin + ((-g.s_begin + 3) & ~3) : g.s_begin;
    const int top = min(g.end, p.N);
    g.hi4 = g.s_begin + ((top - g.s_begin) & ~3);
    if (g.hi4 < g.lo4) g.hi4 = g.lo4;
    return g;
}
__device__ __forceinline__ uint32_t bulk_bytes(const StageGeom& g) { return static_cast<uint32_t>(g.hi4 - g.lo4) * 4u; }
__device__ __forceinline__ void bulk_issue(const float* __restrict__ xrow, const StageGeom& g, uint32_t raw_addr, uint32_t bar) {
    if (bulk_bytes(g))
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(raw_addr + static_cast<uint32_t>(g.lo4 - g.s_begin) * 4u), "l"(xrow + g.lo4), "r"(bulk_bytes(g)), "r"(bar) : "memory");
}
// Both segments of a unit: one arrival carrying the byte count, then the (up to) two bulk copies.  One thread.
__device__ __forceinline__ void stage_bulk(const float* __restrict__ x, long long xbase4, long long stride, const UnitWalk& w,
                                           const Params& p, uint32_t raw_addr, uint32_t bar) {
    const long long g0 = w.b * stride;
    const StageGeom a = stage_geom(xbase4 + g0, w.t, w.n0(), p);
    const int n1 = w.n1(p.B);
    StageGeom c = a;
    if (n1 > 0) c = stage_geom(xbase4 + g0 + stride, 0, n1, p);
    const uint32_t bytes = bulk_bytes(a) + (n1 > 0 ? bulk_bytes(c) : 0u);
    if (bytes) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    else mbar_arrive(bar);
    bulk_issue(x + g0, a, raw_addr, bar);
    if (n1 > 0) bulk_issue(x + g0 + stride, c, raw_addr + static_cast<uint32_t>(a.end - a.s_begin) * 4u, bar);
}
__device__ __forceinline__ void rest_one(const float* __restrict__ xrow, const StageGeom& g, float* __restrict__ raw, int N, int lane) {
    const int nlo = g.lo4 - g.s_begin, nhi = g.end - g.hi4;
    for (int e = lane; e < nlo + nhi; e += 32) {
        const int i = e < nlo ? e : (g.hi4 - g.s_begin) + (e - nlo);
        int sidx = g.s_begin + i;
        sidx = sidx < 0 ? -sidx : sidx;
        sidx = sidx >= N ? 2 * (N - 1) - sidx : sidx;
        sidx = max(0, min(sidx, N - 1));
        raw[i] = __ldg(xrow + sidx);
    }
}
// The samples the bulk copies leave out (alignment slack, reflect padding of a signal's first / last frames).  One warp.
__device__ __forceinline__ void stage_rest(const float* __restrict__ x, long long xbase4, long long stride, const UnitWalk& w,
                                           const Params& p, float* __restrict__ raw, int lane) {
    const long long g0 = w.b * stride;
    const StageGeom a = stage_geom(xbase4 + g0, w.t, w.n0(), p);
    rest_one(x + g0, a, raw, p.N, lane);
    const int n1 = w.n1(p.B);
    if (n1 > 0) rest_one(x + g0 + stride, stage_geom(xbase4 + g0 + stride, 0, n1, p), raw + (a.end - a.s_begin), p.N, lane);
}

// MODE 0: power 2 + log (the reference's configuration), 1: power 2, linear, 2: any power / log flag (runtime)
template <int MODE>
__device__ __forceinline__ float finish(float re, float im, const Params& p) {
    float v = fmaf(re, re, im * im);
    if (MODE == 0) return ln_approx(v + p.log_offset);
    if (MODE == 1) return v;
    if (!p.square) v = powf(sqrtf(v), p.power);
    if (p.log_scale) v = __logf(v + p.log_offset);
    return v;
}

// Epilogue of one unit for the warps of k2 half KH: accumulator columns -> output tile.
// Accumulator block of a part (Re at +0, Im at +96): columns [0, 48) = A_hi B_hi + A_lo B_hi, [48, 96) = A_hi B_lo.
#ifndef AFD_TC_EPI_PIPELINED
#define AFD_TC_EPI_PIPELINED 0      // r2: measured, no effect (347.4 vs 346.9 us, bit-identical): the TMEM load latency is not what the epilogue waits for
#endif
template <int MODE, int KH>
__device__ __forceinline__ void combine(uint32_t taddr, uint32_t bar_d_empty, float* __restrict__ orow, const uint32_t (&binpk)[5],
                                        bool lane_live, int h, const Params& p) {
    constexpr int kLo = KH ? 19 : 0, kHi = KH ? 36 : 18;
#if AFD_TC_EPI_PIPELINED
    // Six rounds of four accumulator columns, the tensor-memory loads of round r + 1 in flight while round r is finished
    // (three rounds of eight columns exposed the full tcgen05.ld latency three times per unit and warp).
    float re[2][4], im[2][4], re2[2][4], im2[2][4];
    constexpr int kBase = KH ? 16 : 0;
    tc_ld4(taddr + kBase, re[0]);
    tc_ld4(taddr + 48 + kBase, re2[0]);
    tc_ld4(taddr + 96 + kBase, im[0]);
    tc_ld4(taddr + 144 + kBase, im2[0]);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const int col0 = kBase + 4 * r;
        tc_ld_wait();
        if (r < 5) {
            tc_ld4(taddr + col0 + 4, re[(r + 1) & 1]);
            tc_ld4(taddr + 48 + col0 + 4, re2[(r + 1) & 1]);
            tc_ld4(taddr + 96 + col0 + 4, im[(r + 1) & 1]);
            tc_ld4(taddr + 144 + col0 + 4, im2[(r + 1) & 1]);
        } else {                               // every accumulator of the unit is in registers: release the buffer
            tc_fence_before();
            mbar_arrive(bar_d_empty);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k2 = col0 + i;
            if (k2 < kLo || k2 > kHi) continue;
            if (AFD_KO(2) && k2 != kLo) continue;
            const float rr = re[r & 1][i] + re2[r & 1][i], q = im[r & 1][i] + im2[r & 1][i];
            const float pre = __shfl_xor_sync(0xffffffffu, rr, 16);
            const float pim = __shfl_xor_sync(0xffffffffu, q, 16);
            const float val = finish<MODE>(rr - pim, q + pre, p);
            const int idx = k2 - kLo;
            const uint32_t bin = (binpk[idx >> 2] >> (8 * (idx & 3))) & 255u;
            const bool live = lane_live && !(h == 1 && k2 == 0);
            if (live) orow[bin] = val;
        }
    }
#else
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int col0 = (KH ? 16 : 0) + 8 * c;
        float re[8], im[8], re2[8], im2[8];
        tc_ld8(taddr + col0, re);
        tc_ld8(taddr + 48 + col0, re2);
        tc_ld8(taddr + 96 + col0, im);
        tc_ld8(taddr + 144 + col0, im2);
        tc_ld_wait();
        if (c == 2) {                          // every accumulator of the unit is in registers: release the buffer
            tc_fence_before();
            mbar_arrive(bar_d_empty);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k2 = col0 + i;
            if (k2 < kLo || k2 > kHi) continue;
            if (AFD_KO(2) && k2 != kLo) continue;
            const float r = re[i] + re2[i], q = im[i] + im2[i];
            const float pre = __shfl_xor_sync(0xffffffffu, r, 16);
            const float pim = __shfl_xor_sync(0xffffffffu, q, 16);
            const float val = finish<MODE>(r - pim, q + pre, p);
            const int idx = k2 - kLo;
            const uint32_t bin = (binpk[idx >> 2] >> (8 * (idx & 3))) & 255u;
            const bool live = lane_live && !(h == 1 && k2 == 0);
            if (live) orow[bin] = val;
        }
    }
#endif
}

template <bool EXT, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
stft_tc511_kernel(const float* __restrict__ x, long long x_row_stride, float* __restrict__ out,
                  const float* __restrict__ tables, const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    float* const s_w = reinterpret_cast<float*>(smem + kOffW);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_p_full = s_base + kOffBar, bar_q_full = bar_p_full + 8;
    const uint32_t bar_p_free = bar_p_full + 16, bar_q_free = bar_p_full + 24;
    const uint32_t bar_d_full = bar_p_full + 32, bar_d_empty = bar_p_full + 48;      // two each
    uint32_t* const s_tmem = reinterpret_cast<uint32_t*>(smem + kOffBar + 64);
    const uint32_t bar_raw_full = bar_p_full + 128;                                    // kRawBufs: bulk copies of the staging buffers

    // ---- one-time setup: tables -> shared memory, A tiles zeroed, barriers, tensor memory
    {
        const float4* src = reinterpret_cast<const float4*>(tables);
        float4* dstB = reinterpret_cast<float4*>(smem + kOffB);
        for (int i = tid; i < 4 * kBTile / 16; i += kThreads) dstB[i] = __ldg(src + i);
        for (int i = tid; i < kWFloats; i += kThreads) s_w[i] = __ldg(tables + kBTile + i);  // window follows the B tiles
        float4* a = reinterpret_cast<float4*>(smem + kOffAHi);
        for (int i = tid; i < 2 * kATile / 16; i += kThreads) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* o = reinterpret_cast<float4*>(smem + kOffRaw);
        for (int i = tid; i < (kRawBufs * kRawFloats + 2 * kOutFloats) / 4; i += kThreads) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        mbar_init(bar_p_full, kPreThreads);
        mbar_init(bar_q_full, kPreThreads);
        mbar_init(bar_p_free, 1);
        mbar_init(bar_q_free, 1);
        mbar_init(bar_d_full, 1);
        mbar_init(bar_d_full + 8, 1);
        mbar_init(bar_d_empty, kEpiThreads);
        mbar_init(bar_d_empty + 8, kEpiThreads);
        for (int i = 0; i < kRawBufs; ++i) mbar_init(bar_raw_full + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kThreads / 32 - 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes of the tables / zeros -> async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int first_unit = blockIdx.x;                    // total_units < 2^31 (checked on the host)
    const int ustride = gridDim.x;
    const int total_units = static_cast<int>(p.total_units);
    const long long xbase4 = static_cast<long long>(reinterpret_cast<uintptr_t>(x) >> 2);   // rows are aligned by ABSOLUTE address
    PROF_DECL();

    if (warp < kEpiWarps) {
        // ============================================================ epilogue warps
        const int j = warp & 3, kh = warp >> 2;
        const int h = lane >> 4, f = lane & 15;
        const bool lane_live = !(h == 1 && j == 0);
        uint32_t binpk[5] = {0u, 0u, 0u, 0u, 0u};
        {
            const int lo = kh ? 19 : 0, cnt = kh ? 18 : 19;
#pragma unroll
            for (int idx = 0; idx < 19; ++idx) {               // compile-time indices keep binpk in registers
                const int k2 = lo + idx;
                int kk = (365 * j + 147 * (h ? 73 - k2 : k2)) % kN;
                kk = kk > 255 ? kN - kk : kk;
                if (idx < cnt) binpk[idx >> 2] |= static_cast<uint32_t>(kk & 255) << (8 * (idx & 3));
            }
        }
        float mom_s = 0.f, mom_q = 0.f;
        const float n_rs = (EXT && p.normalize) ? p.nrstd : 1.f, n_dm = (EXT && p.normalize) ? -p.nmean * p.nrstd : 0.f;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(32 * j) << 16);
        int it = 0;
        for (int unit = first_unit; unit < total_units; unit += ustride, ++it) {
            const long long G0 = static_cast<long long>(unit) * kRows;      // first output row of the unit
            const int valid = static_cast<int>(min(static_cast<long long>(kRows), p.total_rows - G0));
            const int buf = it & 1;
            float* const s_out = reinterpret_cast<float*>(smem + kOffOut) + buf * kOutFloats;
            PROF_START();
            if (warp == 0) mbar_wait_warp(bar_d_full + 8 * buf, (it >> 1) & 1);   // one warp polls ...
            named_bar(3, kEpiThreads);                           // ... the others park at a hardware barrier (no issue slots)
            PROF_LAP(0);
            tc_fence_after();
            const uint32_t taddr = lane_addr + buf * kDStride;
            if (kh == 0) combine<MODE, 0>(taddr, bar_d_empty + 8 * buf, s_out + f * kOutStride, binpk, lane_live, h, p);
            else combine<MODE, 1>(taddr, bar_d_empty + 8 * buf, s_out + f * kOutStride, binpk, lane_live, h, p);
            PROF_LAP(1);
            if constexpr (!EXT) {
                // plain features: the valid rows leave as bulk copies (one 1 KB row per lane of warp 0), asynchronously
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // tile writes -> async proxy
                if (warp == 0 && lane < kRows) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // unit it-1's rows are out of the other tile
                named_bar(1, kEpiThreads);                       // the unit's tile is complete; the other tile is free
                PROF_LAP(2);
                if (warp == 0 && lane < kRows) {
                    if (lane < valid && AFD_TC_KO != 8) {
                        float* og = out + (G0 + lane) * 256LL;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1024;"
                                     ::"l"(og), "r"(smem_u32(s_out + lane * kOutStride)) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
            named_bar(1, kEpiThreads);                           // the unit's tile is complete
            PROF_LAP(2);
            float* og = out + G0 * 256LL;
            for (int i = tid; i < valid * 64; i += kEpiThreads) {
                const int row = i >> 6, c4 = (i & 63) * 4;
                float4 v = *reinterpret_cast<const float4*>(s_out + row * kOutStride + c4);
                if (p.moments) {
                    mom_s += (v.x + v.y) + (v.z + v.w);
                    mom_q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, mom_q))));
                }
                if (p.normalize) {
                    v.x = fmaf(v.x, n_rs, n_dm); v.y = fmaf(v.y, n_rs, n_dm);
                    v.z = fmaf(v.z, n_rs, n_dm); v.w = fmaf(v.w, n_rs, n_dm);
                }
                if (p.store) st_cs4(reinterpret_cast<float4*>(og + row * 256 + c4), v);
            }
            }
            PROF_LAP(3);
            // the other tile is rewritten only after every thread passed the next unit's barrier
        }
        if (!EXT && warp == 0 && lane < kRows) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        PROF_PRINT((tid & 31) == 5, "epi: wait d_full / combine / bar / store");   // one thread of every epilogue warp
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
    } else if (warp < kEpiWarps + kPreWarps) {
        // ============================================================ producer warps: thread = (frames r0 and r0 + 8, column m)
        // warps 0..7: r0 = warp, m = lane (0..31): data and A-tile accesses of a warp walk along m and are bank-conflict
        // free; warps 8..9: the remaining m = 32..36 (40 threads).  Two independent frames per thread give the scheduler
        // twice the instruction-level parallelism per warp; the window values of column m stay in registers.
        const int pt = tid - kEpiThreads;
        const int pw = pt >> 5;
        const int pe = (pw - 8) * 32 + lane;
        const bool active = pw < 8 || pe < 40;
        const int r0 = pw < 8 ? pw : (pe < 40 ? pe & 7 : 0);            // frames fastest: banks 4 r0 + (m & 3) are distinct
        const int pm = pw < 8 ? lane : (pe < 40 ? 32 + (pe >> 3) : 0);
        // Sample (n1, m) of a frame sits at offA[n1] = (73 n1 + 7 m) mod 511, its fold partner (n1, 73 - m) at
        // (73 n1 - 7 m) mod 511 = 511 - offA[(7 - n1) mod 7], and the periodic Hann window is symmetric, so one offset and
        // one window value per n1 serve both.  (m = 0: partner == sample, pp = 2 xa exactly; the table row m = 0 is 0.5.)
        int offA[7];
        float wA[7];
#pragma unroll
        for (int n1 = 0; n1 < 7; ++n1) {
            int a = 73 * n1 + 7 * pm;
            a = a >= kN ? a - kN : a;
            offA[n1] = a;
            wA[n1] = s_w[74 * n1 + pm];
        }
        const int wrap0 = pm == 0 ? 0 : kN;                    // (n1, m) = (0, 0) is its own partner at offset 0
        unsigned char* const arow = smem + kOffAHi + (pm >> 2) * kALbo + (pm & 3) * 4 + r0 * 16;     // frame r0 + 8: + kASbo
        constexpr int kQ = 10 * kALbo;                                                      // column 40 + m
        float* const raw0 = reinterpret_cast<float*>(smem + kOffRaw);
        UnitWalk uw(first_unit, ustride, p.frames);
        if (first_unit < total_units && pw == kPreWarps - 1) stage_rest(x, xbase4, x_row_stride, uw, p, raw0, lane);
        int it = 0;
        for (int unit = first_unit; unit < total_units; unit += ustride, ++it, uw.next()) {
            const int n0 = uw.n0();
            const int valid = n0 + uw.n1(p.B);
            // where the thread's frames start in the staging buffer: segment 0 (signal b) or segment 1 (signal b + 1)
            const int abs0 = static_cast<int>(xbase4 + uw.b * x_row_stride);
            const int aoff0 = (abs0 + (uw.t * p.hop - p.pad)) & 3;
            const int off1 = (aoff0 + (n0 - 1) * p.hop + kN + 3) & ~3;
            const int aoff1 = (abs0 + static_cast<int>(x_row_stride) - p.pad) & 3;
            const float* raw = raw0 + (it % kRawBufs) * kRawFloats;
            PROF_START();
            mbar_wait_warp(bar_raw_full + 8 * (it % kRawBufs), (it / kRawBufs) & 1);     // the unit's bulk copy has landed
            named_bar(2, kPreThreads);                           // ... and its edge samples; the other buffer is no longer read
            PROF_LAP(0);
            if (pw == kPreWarps - 1 && unit + ustride < total_units)     // the warp with the fewest items writes the edge samples
                stage_rest(x, xbase4, x_row_stride, uw.peek(), p, raw0 + ((it + 1) % kRawBufs) * kRawFloats, lane);
            // the frames are computed BEFORE the A tiles are claimed: this overlaps the previous unit's MMAs
            float P[2][7], Q[2][7];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = r0 + 8 * rr;
                // rows past the end of the batch (last unit only) read frame 0 of the buffer: finite values, never stored
                const int fo = r < n0 ? aoff0 + r * p.hop : (r < valid ? off1 + aoff1 + (r - n0) * p.hop : 0);
                const float* fr = raw + fo;
                float pp[7], qq[7];
#pragma unroll
                for (int n1 = 0; n1 < 7; ++n1) {
                    if (AFD_KO(3)) { pp[n1] = fr[0]; qq[n1] = fr[1]; continue; }
                    const int nb = (7 - n1) % 7;
                    const float xa = fr[offA[n1]] * wA[n1];
                    const float xb = fr[(n1 == 0 ? wrap0 : kN) - offA[nb]] * wA[nb];
                    pp[n1] = xa + xb;
                    qq[n1] = xa - xb;
                }
                dft7_real(pp, P[rr]);
                dft7_real(qq, Q[rr]);
            }
            PROF_LAP(1);
            if (it > 0) mbar_wait_warp(bar_p_free, (it - 1) & 1);     // the previous unit's Re MMAs have read the P block
            PROF_LAP(2);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
                if (active) {
#pragma unroll
                    for (int s = 0; s < (AFD_KO(4) ? 1 : 7); ++s) {
                        const int row8 = (s == 0 ? 0 : 4 * ((s + 1) >> 1) + 2 * ((s + 1) & 1)) + rr;   // (32 j + 16 h) / 8 + frame / 8
                        const float hi = __uint_as_float(__float_as_uint(P[rr][s]) & 0xffffe000u);
                        *reinterpret_cast<float*>(arow + row8 * kASbo) = hi;
                        *reinterpret_cast<float*>(arow + row8 * kASbo + kATile) = P[rr][s] - hi;
                    }
                }
            if (!AFD_KO(5)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_p_full);
            PROF_LAP(3);
            if (it > 0) mbar_wait_warp(bar_q_free, (it - 1) & 1);     // ... the Im MMAs the Q block
            PROF_LAP(4);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
                if (active) {                                    // column 40 (m = 0) meets an all-zero table row: any finite value
#pragma unroll
                    for (int s = 0; s < (AFD_KO(4) ? 1 : 7); ++s) {
                        const int row8 = (s == 0 ? 0 : 4 * ((s + 1) >> 1) + 2 * ((s + 1) & 1)) + rr;
                        const float hi = __uint_as_float(__float_as_uint(Q[rr][s]) & 0xffffe000u);
                        *reinterpret_cast<float*>(arow + row8 * kASbo + kQ) = hi;
                        *reinterpret_cast<float*>(arow + row8 * kASbo + kQ + kATile) = Q[rr][s] - hi;
                    }
                }
            if (!AFD_KO(5)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_q_full);
            PROF_LAP(5);
        }
        PROF_PRINT(pt == 0 || pt == 300, "pre: raw wait+bar / stage+compute / wait p_free / sts P / wait q_free / sts Q");
    } else if (lane == 0) {
        // ============================================================ MMA issuer (one thread)
        const uint32_t a_hi = s_base + kOffAHi, a_lo = s_base + kOffALo;
        const uint32_t b0 = s_base + kOffB;
        const uint32_t raw_addr = s_base + kOffRaw;
        UnitWalk lw(first_unit, ustride, p.frames);              // walks two units ahead: the staging loads
        auto load_unit = [&](int n) {                            // bulk copies of the CTA's n-th unit into buffer n & 1
            stage_bulk(x, xbase4, x_row_stride, lw, p, raw_addr + (n % kRawBufs) * kRawFloats * 4, bar_raw_full + 8 * (n % kRawBufs));
            lw.next();
        };
        for (int n = 0; n < kRawBufs; ++n)
            if (static_cast<long long>(first_unit) + static_cast<long long>(n) * ustride < total_units) load_unit(n);
        int it = 0;
        for (int unit = first_unit; unit < total_units; unit += ustride, ++it) {
            const int buf = it & 1;
#pragma unroll
            for (int part = 0; part < 2; ++part) {               // 0: Re = P x C, 1: Im = Q x S
                PROF_START();
                mbar_wait(part ? bar_q_full : bar_p_full, it & 1);
                if (part == 0 && it >= 2) mbar_wait(bar_d_empty + 8 * buf, ((it >> 1) - 1) & 1);
                PROF_LAP(2 * part);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * kDStride + part * 96;
                const uint32_t bh = b0 + (2 * part) * kBTile;     // [hi tile | lo tile] = 96 rows of one N = 96 operand
                const uint32_t ak = part * 10 * kALbo;
#pragma unroll
                for (int ks = 0; ks < (AFD_KO(1) ? 1 : 5); ++ks)     // A_hi x [B_hi | B_lo] -> columns [0, 96)
                    tc_mma_tf32(d, make_desc(a_hi + ak + 2 * ks * kALbo, kALbo, kASbo), make_desc(bh + 2 * ks * kBLbo, kBLbo, kBSbo),
                                kIdesc96, ks != 0);
#pragma unroll
                for (int ks = 0; ks < (AFD_KO(1) ? 0 : 5); ++ks)     // A_lo x B_hi -> columns [0, 48)
                    tc_mma_tf32(d, make_desc(a_lo + ak + 2 * ks * kALbo, kALbo, kASbo), make_desc(bh + 2 * ks * kBLbo, kBLbo, kBSbo),
                                kIdesc48, 1);
                tc_commit(part ? bar_q_free : bar_p_free);
                // p_full(it) has completed: every producer has finished reading staging buffer it % kRawBufs -> refill it
                // (after the MMAs are issued: the address arithmetic of the bulk copies stays off the A-tile round trip)
                if (part == 0 && static_cast<long long>(unit) + static_cast<long long>(kRawBufs) * ustride < total_units)
                    load_unit(it + kRawBufs);
                if (part == 1) tc_commit(bar_d_full + 8 * buf);
                PROF_LAP(2 * part + 1);
            }
        }
        PROF_PRINT(true, "mma: wait p_full(+d_empty) / issue Re / wait q_full / issue Im");
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kThreads / 32 - 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
static std::mutex g_mutex;
static std::map<int, float*> g_tables;

static float tf32_trunc(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u &= 0xffffe000u;
    memcpy(&f, &u, 4);
    return f;
}

static int get_tables(int dev, float** out) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_tables.find(dev);
    if (it != g_tables.end()) { *out = it->second; return AFD_OK; }
    std::vector<float> h(kTableFloats, 0.f);
    auto bval = [](int part, int m, int k2) -> double {
        if (m > 36 || k2 > 36) return 0.0;
        const int ph = (m * k2) % 73;                                    // exact phase reduction
        const double ang = 2.0 * M_PI * double(ph) / 73.0;
        if (part == 0) return m == 0 ? 0.5 : cos(ang);                   // the producers deliver 2 * p[., 0]
        return m == 0 ? 0.0 : -sin(ang);
    };
    for (int part = 0; part < 2; ++part)
        for (int n = 0; n < 48; ++n)
            for (int k = 0; k < 40; ++k) {
                const double v = bval(part, k, n);
                const float hi = tf32_trunc(static_cast<float>(v));
                const float lo = static_cast<float>(v - static_cast<double>(hi));
                const int byte = (n & 7) * 16 + (n >> 3) * kBSbo + (k >> 2) * kBLbo + (k & 3) * 4;
                h[((2 * part) * kBTile + byte) / 4] = hi;
                h[((2 * part + 1) * kBTile + byte) / 4] = lo;
            }
    auto hann = [](int n) { return 0.5 - 0.5 * cos(2.0 * M_PI * double(n) / double(kN)); };   // periodic Hann
    float* w = h.data() + kBTile;                                         // 4 * kBTile bytes = kBTile floats
    for (int m = 0; m < 37; ++m)
        for (int n1 = 0; n1 < 7; ++n1) {
            const int a = (73 * n1 + 7 * m) % kN;
            const int b = ((73 * n1 - 7 * m) % kN + kN) % kN;
            w[n1 * 74 + m] = static_cast<float>(hann(a));
            w[n1 * 74 + 37 + m] = m == 0 ? 0.f : static_cast<float>(hann(b));
        }
    float* d = nullptr;
    AFD_CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(float)));
    AFD_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    g_tables[dev] = d;
    *out = d;
    return AFD_OK;
}

}  // namespace tc

bool stft_tc511_supported(const float* x, int64_t B, int64_t N, int n_fft, int hop, const float* out) {
    const int64_t frames = 1 + (N + 2 * (tc::kN / 2) - tc::kN) / (hop > 0 ? hop : 1);
    if (B >= (1LL << 31) - 1 || B * frames >= (1LL << 31) - (1 << 20)) return false;      // 32-bit row / unit indices
    return n_fft == tc::kN && hop >= 1 && hop <= tc::kMaxHop && N > tc::kN / 2 && frames >= tc::kRows &&   // a unit spans <= 2 signals
           (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0;
}

int stft_tc511_launch(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int hop, float power, int log_scale,
                      float log_offset, const StftExtras& ex, float* out, cudaStream_t stream) {
    using namespace tc;
    int dev = 0;
    AFD_CUDA_TRY(cudaGetDevice(&dev));
    float* tables = nullptr;
    int rc = get_tables(dev, &tables);
    if (rc != AFD_OK) return rc;
    Params p;
    p.hop = hop; p.N = static_cast<int>(N); p.pad = kN / 2;
    p.frames = static_cast<int>(1 + (N + 2 * (kN / 2) - kN) / hop);
    p.B = static_cast<int>(B);
    p.total_rows = B * static_cast<long long>(p.frames);
    p.total_units = (p.total_rows + kRows - 1) / kRows;
    if (p.total_rows >= (1LL << 31) - (1 << 20) || B >= (1LL << 31) - 1) return fail(AFD_ERR_UNSUPPORTED, "afd_stft_power: batch too large for one launch");
    p.power = power; p.log_offset = log_offset; p.log_scale = log_scale ? 1 : 0; p.square = (power == 2.0f);
    p.normalize = ex.normalize; p.nmean = ex.nmean; p.nrstd = ex.nrstd; p.moments = ex.moments; p.store = out != nullptr;
    const bool ext = ex.normalize || ex.moments || !out;
    const int mode = p.square ? (p.log_scale ? 0 : 1) : 2;
    using Kern = void (*)(const float*, long long, float*, const float*, const Params);
    static const Kern kerns[2][3] = {
        {stft_tc511_kernel<false, 0>, stft_tc511_kernel<false, 1>, stft_tc511_kernel<false, 2>},
        {stft_tc511_kernel<true, 0>, stft_tc511_kernel<true, 1>, stft_tc511_kernel<true, 2>}};
    Kern kern = kerns[ext][mode];
    static thread_local bool configured[2][3][16] = {};
    if (dev >= 16 || !configured[ext][mode][dev]) {
        AFD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        if (dev < 16) configured[ext][mode][dev] = true;
    }
    int sms = kNumSmsFallback;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = p.total_units < sms ? p.total_units : sms;
    kern<<<static_cast<unsigned>(blocks), kThreads, kSmemBytes, stream>>>(x, static_cast<long long>(x_row_stride), out, tables, p);
    AFD_CUDA_TRY(cudaGetLastError());
    return AFD_OK;
}

}  // namespace afd
