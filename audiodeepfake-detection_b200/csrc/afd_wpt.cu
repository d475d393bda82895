// C ABI of the wavelet-packet transform (kernel and planning: afd_wpt_kernel.cuh; instantiations: afd_wpt_g*.cu).
#include "afd_wpt_kernel.cuh"

using namespace afd;

static int wpt_dispatch(int F, const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* out, int L,
                        const double* dec_lo, const Epilogue& ep, cudaStream_t stream, PlanReport* report) {
    const WptGroupFn groups[2][4] = {{wpt_group0, wpt_group1, wpt_group2, wpt_group3},
                                     {wpt_xgroup0, wpt_xgroup1, wpt_xgroup2, wpt_xgroup3}};
    const bool extended = ep.node_stats || ep.node_scale || ep.feat_moments || ep.normalize || !ep.store;
    return groups[extended && !report ? 1 : 0][(F - 2) / 16](F, x, B, N, x_row_stride, out, L, dec_lo, ep, stream, report);
}

extern "C" int afd_wpt_out_len(int64_t N, int F, int level, int64_t* T_out) {
    if (N < 1 || F < 2 || (F & 1) || level < 0 || !T_out) return fail(AFD_ERR_INVALID_ARG, "afd_wpt_out_len: bad argument");
    int64_t n = N;
    for (int l = 0; l < level; ++l) n = (n + F - 1) / 2;
    *T_out = n;
    return AFD_OK;
}

static int wpt_forward_impl(const char* who, const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                            const double* dec_lo_host, int F, int level, int order, float power, int log_scale,
                            float log_offset, int sign_channel, Epilogue ep, float* out, int64_t* T_out, void* stream) {
    const bool stats_only = out == nullptr && (ep.node_stats != nullptr || ep.feat_moments != nullptr);
    if (!dec_lo_host || ((!x || (!out && !stats_only)) && B != 0)) return fail(AFD_ERR_INVALID_ARG, "%s: null pointer", who);
    if (B < 0 || N < 2 || x_row_stride < N) return fail(AFD_ERR_INVALID_ARG, "%s: bad B/N/stride", who);
    if (F < 2 || F > 64 || (F & 1)) return fail(AFD_ERR_INVALID_ARG, "%s: filter length %d not in {2,4,..,64}", who, F);
    if (level < 1 || level > kMaxLevel) return fail(AFD_ERR_INVALID_ARG, "%s: level %d not in 1..%d", who, level, kMaxLevel);
    if (order != AFD_ORDER_FREQ && order != AFD_ORDER_NATURAL) return fail(AFD_ERR_INVALID_ARG, "%s: bad order", who);
    if (N > (1 << 20)) return fail(AFD_ERR_UNSUPPORTED, "%s: frame too long", who);
    if (B > (1LL << 30)) return fail(AFD_ERR_UNSUPPORTED, "%s: batch too large", who);
    int64_t T = 0;
    afd_wpt_out_len(N, F, level, &T);
    if (T_out) *T_out = T;
    if (B == 0) return AFD_OK;
    ep.power = power;
    ep.log_offset = log_offset;
    ep.log_scale = log_scale ? 1 : 0;
    ep.sign_channel = sign_channel ? 1 : 0;
    ep.order = order;
    ep.square = (power == 2.0f);
    ep.store = out != nullptr ? 1 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return wpt_dispatch(F, x, B, N, x_row_stride, out, level, dec_lo_host, ep, s, nullptr);
}

extern "C" int afd_wpt_forward(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                               const double* dec_lo_host, int F, int level, int order, float power,
                               int log_scale, float log_offset, int sign_channel, float* out,
                               int64_t* T_out, void* stream) {
    return wpt_forward_impl("afd_wpt_forward", x, B, N, x_row_stride, dec_lo_host, F, level, order, power, log_scale,
                            log_offset, sign_channel, Epilogue{}, out, T_out, stream);
}

extern "C" int afd_wpt_forward_ex(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                                  const double* dec_lo_host, int F, int level, int order, float power,
                                  int log_scale, float log_offset, int sign_channel, const float* node_scale,
                                  const float* norm_mean_std_host, double* node_stats, double* feat_moments,
                                  float* out, int64_t* T_out, void* stream) {
    Epilogue ep{};
    ep.node_scale = node_scale;
    ep.node_stats = node_stats;
    ep.feat_moments = feat_moments;
    if (norm_mean_std_host) {
        const int C = (log_scale && sign_channel) ? 2 : 1;
        ep.normalize = 1;
        for (int c = 0; c < C; ++c) {
            const float mean = norm_mean_std_host[2 * c], std = norm_mean_std_host[2 * c + 1];
            if (!(std > 0.f) || !isfinite(mean) || !isfinite(std))
                return fail(AFD_ERR_INVALID_ARG, "afd_wpt_forward_ex: channel %d needs a finite mean and a positive std", c);
            ep.nmean[c] = mean;
            ep.nrstd[c] = 1.0f / std;
        }
    }
    return wpt_forward_impl("afd_wpt_forward_ex", x, B, N, x_row_stride, dec_lo_host, F, level, order, power,
                            log_scale, log_offset, sign_channel, ep, out, T_out, stream);
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
