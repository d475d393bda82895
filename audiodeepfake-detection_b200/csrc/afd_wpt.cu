// C ABI of the wavelet-packet transform (kernel and planning: afd_wpt_kernel.cuh; instantiations: afd_wpt_g*.cu).
#include "afd_wpt_kernel.cuh"

using namespace afd;

static int wpt_dispatch(int F, const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                        const double* dec_lo, const Epilogue& ep, cudaStream_t stream, PlanReport* report) {
    const WptGroupFn groups[4] = {wpt_group0, wpt_group1, wpt_group2, wpt_group3};
    return groups[(F - 2) / 16](F, x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report);
}

extern "C" int afd_wpt_out_len(int64_t N, int F, int level, int64_t* T_out) {
    if (N < 1 || F < 2 || (F & 1) || level < 0 || !T_out) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_out_len: bad argument");
    int64_t n = N;
    for (int l = 0; l < level; ++l) n = (n + F - 1) / 2;
    *T_out = n;
    return AFD_OK;
}

extern "C" int afd_wpt_forward(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                               const double* dec_lo_host, int F, int level, int order, float power,
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
    return wpt_dispatch(F, x, B, N, x_row_stride, out, level, dec_lo_host, ep, s, nullptr);
}

extern "C" int afd_wpt_plan_info(int64_t N, const double* dec_lo_host, int F, int level, int* smem_bytes,
                                 int* ctas_per_sm, int* lattice, int* passes, int* pass_items, int* pass_r) {
    if (!dec_lo_host || N < 2 || F < 2 || F > 64 || (F & 1) || level < 1 || level > kMaxLevel)
        return fail(AFD_ERR_INVALID_ARG, "afd_wpt_plan_info: bad argument");
    PlanReport rep{};
    Epilogue ep{};
    const int rc = wpt_dispatch(F, nullptr, 1, N, N, nullptr, level, dec_lo_host, ep, nullptr, &rep);
    if (rc != AFD_OK) return rc;
    if (smem_bytes) *smem_bytes = rep.smem_bytes;
    if (ctas_per_sm) *ctas_per_sm = rep.ctas_per_sm;
    if (lattice) *lattice = rep.lattice;
    if (passes) *passes = rep.passes;
    for (int i = 0; i < rep.passes; ++i) {
        if (pass_items) pass_items[i] = rep.pass_items[i];
        if (pass_r) pass_r[i] = rep.pass_r[i];
    }
    return AFD_OK;
}
