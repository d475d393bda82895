// Paraunitary lattice factorisation of an orthogonal two-channel analysis bank (host side, double precision).
//
// The analysis step the reference executes through ptwt (wavelet_math.py:182) is, per node,
//     lo[k] = sum_m h[m] x~[2k+1-m],   hi[k] = sum_m g[m] x~[2k+1-m],   g[m] = (-1)^(m+1) h[F-1-m].
// With the polyphase pairs p[k] = (x~[2k+1], x~[2k]) and the 2x2 tap blocks A_j = [[h[2j], h[2j+1]], [g[2j], g[2j+1]]]
// this is  y[k] = sum_j A_j p[k-j],  i.e. the polyphase matrix E(z) = sum_j A_j z^-j.  For an orthonormal wavelet
// E(z) is paraunitary and factors into J = F/2 plane rotations separated by one-sample delays of the second
// channel (Vaidyanathan):
//     E(z) = R_{J-1} L(z) R_{J-2} L(z) ... L(z) R_0,      L(z) = diag(1, z^-1),   R_m = c_m [[1, t_m], [-t_m, 1]].
// Evaluated that way one output PAIR costs J rotations = F FFMAs instead of the 2F of the direct form.  The kernels
// run the rotations unscaled (only t_m = tan(theta_m)); the product of the c_m is applied once, later.
//
// Steps: (1) peel the stages off E(z) (order reduction from whichever end block is better conditioned),
// (2) polish the angles with Levenberg-damped Gauss-Newton on the tap residual -- tabulated filters are paraunitary
// only to ~1e-11, which the peeling amplifies by 1/|end tap| (5e-7 for coif4 before, 5e-10 after polishing),
// (3) report residual, growth and sign so that the caller can fall back to the direct form when the lattice is not
// trustworthy for a given filter.
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "afd_common.cuh"

namespace afd {

namespace {

struct M2 {
    double a, b, c, d;   // [[a, b], [c, d]]
};
inline M2 mul(const M2& x, const M2& y) {
    return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}

std::vector<M2> tap_blocks(const double* h, int F) {
    const int J = F / 2;
    std::vector<M2> A(J);
    auto g = [&](int k) { return ((k & 1) ? 1.0 : -1.0) * h[F - 1 - k]; };
    for (int j = 0; j < J; ++j) A[j] = {h[2 * j], h[2 * j + 1], g(2 * j), g(2 * j + 1)};
    return A;
}

M2 stage_matrix(double theta, bool reflect) {
    const double c = cos(theta), s = sin(theta);
    return reflect ? M2{c, s, s, -c} : M2{c, s, -s, c};
}

// E(z) blocks of R_{J-1} L ... L R_0
std::vector<M2> synth(const std::vector<double>& theta, bool reflect0) {
    const int J = static_cast<int>(theta.size());
    std::vector<M2> E{stage_matrix(theta[0], reflect0)};
    for (int m = 1; m < J; ++m) {
        std::vector<M2> D(E.size() + 1, M2{0, 0, 0, 0});
        for (size_t j = 0; j < E.size(); ++j) {
            D[j].a += E[j].a; D[j].b += E[j].b;              // first channel: no delay
            D[j + 1].c += E[j].c; D[j + 1].d += E[j].d;      // second channel: one sample later
        }
        const M2 R = stage_matrix(theta[m], false);
        for (auto& blk : D) blk = mul(R, blk);
        E.swap(D);
    }
    return E;
}

double residual(const std::vector<double>& theta, bool reflect0, const std::vector<M2>& A, std::vector<double>* r) {
    const std::vector<M2> E = synth(theta, reflect0);
    double worst = 0;
    if (r) r->resize(4 * A.size());
    for (size_t j = 0; j < A.size(); ++j) {
        const double d[4] = {E[j].a - A[j].a, E[j].b - A[j].b, E[j].c - A[j].c, E[j].d - A[j].d};
        for (int q = 0; q < 4; ++q) {
            if (r) (*r)[4 * j + q] = d[q];
            worst = fmax(worst, fabs(d[q]));
        }
    }
    return worst;
}

// Solve (H + lam * mean(diag H) * I) x = -g by Gaussian elimination with partial pivoting; false if singular.
bool solve_damped(std::vector<double> H, std::vector<double> g, int n, double lam, std::vector<double>* x) {
    double tr = 0;
    for (int i = 0; i < n; ++i) tr += H[i * n + i];
    for (int i = 0; i < n; ++i) { H[i * n + i] += lam * tr / n; g[i] = -g[i]; }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i) if (fabs(H[i * n + k]) > fabs(H[piv * n + k])) piv = i;
        if (fabs(H[piv * n + k]) < 1e-300) return false;
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(H[k * n + j], H[piv * n + j]);
            std::swap(g[k], g[piv]);
        }
        for (int i = k + 1; i < n; ++i) {
            const double f = H[i * n + k] / H[k * n + k];
            for (int j = k; j < n; ++j) H[i * n + j] -= f * H[k * n + j];
            g[i] -= f * g[k];
        }
    }
    x->assign(n, 0.0);
    for (int i = n - 1; i >= 0; --i) {
        double s = g[i];
        for (int j = i + 1; j < n; ++j) s -= H[i * n + j] * (*x)[j];
        (*x)[i] = s / H[i * n + i];
    }
    return true;
}

}  // namespace

static int lattice_factor_uncached(const double* h, int F, LatticeInfo* info) {
    const int J = F / 2;
    if (F < 2 || (F & 1) || J > kMaxLatticeStages) return AFD_ERR_INVALID_ARG;
    std::vector<M2> A = tap_blocks(h, F);
    const std::vector<M2> A_ref = A;
    std::vector<double> theta(J, 0.0);
    // ---- (1) order reduction: E_m(z) = R_m L(z) E_{m-1}(z)
    for (int m = J - 1; m >= 1; --m) {
        const M2& A0 = A[0];
        const M2& Am = A[m];
        const double n0a = hypot(A0.a, A0.c), n0b = hypot(A0.b, A0.d);
        const double nma = hypot(Am.a, Am.c), nmb = hypot(Am.b, Am.d);
        double ax, ay;   // unit column direction of A_0; A_m's columns are orthogonal to it
        if (fmax(n0a, n0b) >= fmax(nma, nmb)) {
            if (n0a >= n0b) { ax = A0.a / n0a; ay = A0.c / n0a; } else { ax = A0.b / n0b; ay = A0.d / n0b; }
        } else {
            double cx, cy;
            if (nma >= nmb) { cx = Am.a / nma; cy = Am.c / nma; } else { cx = Am.b / nmb; cy = Am.d / nmb; }
            ax = cy; ay = -cx;
        }
        if (!(isfinite(ax) && isfinite(ay))) return AFD_ERR_UNSUPPORTED;
        // R_m^T = [[ax, ay], [-ay, ax]]  =>  R_m = [[ax, -ay], [ay, ax]] = [[c, s], [-s, c]] with c = ax, s = -ay
        theta[m] = atan2(-ay, ax);
        const M2 RT{ax, ay, -ay, ax};
        std::vector<M2> B(m + 1);
        for (int j = 0; j <= m; ++j) B[j] = mul(RT, A[j]);
        std::vector<M2> C(m);
        for (int j = 0; j < m; ++j) C[j] = {B[j].a, B[j].b, B[j + 1].c, B[j + 1].d};
        A.swap(C);
    }
    const double det0 = A[0].a * A[0].d - A[0].b * A[0].c;
    const bool reflect0 = det0 < 0;
    theta[0] = atan2(A[0].b, A[0].a);
    // ---- (2) Gauss-Newton polish on the tap residual
    std::vector<double> r;
    double err = residual(theta, reflect0, A_ref, &r);
    for (int it = 0; it < 40 && err > 1e-15; ++it) {
        const int n = J, mrows = 4 * J;
        std::vector<double> Jm(static_cast<size_t>(mrows) * n), rp, rm;
        for (int i = 0; i < n; ++i) {
            std::vector<double> tp = theta, tm = theta;
            tp[i] += 1e-7; tm[i] -= 1e-7;
            residual(tp, reflect0, A_ref, &rp);
            residual(tm, reflect0, A_ref, &rm);
            for (int q = 0; q < mrows; ++q) Jm[static_cast<size_t>(q) * n + i] = (rp[q] - rm[q]) / 2e-7;
        }
        std::vector<double> H(static_cast<size_t>(n) * n, 0.0), g(n, 0.0);
        for (int q = 0; q < mrows; ++q)
            for (int i = 0; i < n; ++i) {
                g[i] += Jm[static_cast<size_t>(q) * n + i] * r[q];
                for (int j = 0; j < n; ++j) H[static_cast<size_t>(i) * n + j] += Jm[static_cast<size_t>(q) * n + i] * Jm[static_cast<size_t>(q) * n + j];
            }
        bool improved = false;
        const double lams[] = {0.0, 1e-12, 1e-9, 1e-6, 1e-3};
        for (double lam : lams) {
            std::vector<double> step;
            if (!solve_damped(H, g, n, lam, &step)) continue;
            std::vector<double> cand = theta, rc;
            for (int i = 0; i < n; ++i) cand[i] += step[i];
            const double e2 = residual(cand, reflect0, A_ref, &rc);
            if (e2 < err) { theta.swap(cand); r.swap(rc); err = e2; improved = true; break; }
        }
        if (!improved) break;
    }
    // ---- (3) report
    info->stages = J;
    info->reflect0 = reflect0 ? 1 : 0;
    info->residual = err;
    info->scale = 1.0;
    info->max_abs_tan = 0.0;
    for (int m = 0; m < J; ++m) {
        const double c = cos(theta[m]);
        info->tan_theta[m] = tan(theta[m]);
        info->scale *= c;
        info->max_abs_tan = fmax(info->max_abs_tan, fabs(info->tan_theta[m]));
    }
    info->usable = (!reflect0 && err <= 2e-9 && isfinite(info->max_abs_tan) && info->max_abs_tan < 1e7 &&
                    fabs(info->scale) > 1e-12) ? 1 : 0;
    return AFD_OK;
}

// The factorisation costs ~1 ms for long filters: cache it per tap vector (the transform modules call with the
// same wavelet every batch).
int lattice_factor(const double* h, int F, LatticeInfo* info) {
    static std::mutex mu;
    static std::map<std::vector<double>, LatticeInfo> cache;
    if (F < 2 || (F & 1)) return AFD_ERR_INVALID_ARG;
    std::vector<double> key(h, h + F);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *info = it->second; return AFD_OK; }
    }
    const int rc = lattice_factor_uncached(h, F, info);
    if (rc != AFD_OK) return rc;
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() < 256) cache[key] = *info;
    return AFD_OK;
}

}  // namespace afd

using namespace afd;

extern "C" int afd_wpt_lattice_info(const double* dec_lo_host, int F, double* tan_theta, double* scale,
                                    double* residual_out, int* usable) {
    if (!dec_lo_host || F < 2 || (F & 1) || F > 2 * kMaxLatticeStages)
        return fail(AFD_ERR_INVALID_ARG, "afd_wpt_lattice_info: bad argument");
    std::vector<double> h(F);
    for (int k = 0; k < F; ++k) h[k] = dec_lo_host[k];
    LatticeInfo info;
    const int rc = lattice_factor(h.data(), F, &info);
    if (rc != AFD_OK) return fail(rc, "afd_wpt_lattice_info: factorisation failed");
    if (tan_theta) for (int m = 0; m < info.stages; ++m) tan_theta[m] = info.tan_theta[m];
    if (scale) *scale = info.scale;
    if (residual_out) *residual_out = info.residual;
    if (usable) *usable = info.usable;
    return AFD_OK;
}
