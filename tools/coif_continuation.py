"""Continue the coiflet family to orders PyWavelets tabulates but nobody here can recall digit by digit (coif6 .. coif10,
swept by the reference's scripts/start_exps.sh:26-31).

The coifN design equations -- sum h = sqrt 2, orthonormality of the even shifts, 2N vanishing moments of psi, 2N - 1 of phi
about k = 2N (rec_lo indexing) -- have several real solutions for every order.  The 4N linear equations are removed by a
null-space parametrisation (80-digit QR), the 3N quadratic ones are solved by damped Gauss-Newton from many starts around
the zero-padded table of order N - 1 (float64 search, 80-digit polish), and the solution NEAREST to that padded table is
kept.  The rule is validated on the orders whose PyWavelets tables are known: from coif2 it returns coif3, from coif3
coif4 (both to table precision) and from coif4 the exact coiflet 1.1e-5 away from PyWavelets' (low precision) coif5, each
time with a 6x .. 9x margin to the second-nearest solution.

Usage:  python tools/coif_continuation.py [first_order=6] [last_order=10]   (minutes per order for N >= 7)
Prints CONTINUED_COIF for tools/gen_wavelets.py.
"""
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import gen_wavelets as G  # noqa: E402

mp.mp.dps = 80


def linear_system(N):
    F = 6 * N
    rows, rhs = [[mp.mpf(1)] * F], [mp.sqrt(2)]
    for p in range(2 * N):          # psi moments (shifted and scaled monomials span the same space)
        rows.append([(-1) ** k * (((mp.mpf(k) - 2 * N) / (3 * N)) ** p if p else 1) for k in range(F)])
        rhs.append(0)
    for p in range(1, 2 * N):       # phi moments about k = 2N
        rows.append([((mp.mpf(k) - 2 * N) / (3 * N)) ** p for k in range(F)])
        rhs.append(0)
    return mp.matrix(rows), mp.matrix(rhs)


def nullspace(N):
    """x_p + Z c spans every filter satisfying the linear equations."""
    A, b = linear_system(N)
    Q, R = mp.qr(A.T)
    m = A.rows
    y = mp.lu_solve(R[:m, :m].T, b)
    return Q[:, :m] * y, Q[:, m:]


def ortho_resid(r):
    F = len(r)
    return np.array([np.dot(r[:F - 2 * m], r[2 * m:]) - (1.0 if m == 0 else 0.0) for m in range(F // 2)])


def ortho_jac(r):
    F = len(r)
    J = np.zeros((F // 2, F))
    for m in range(F // 2):
        J[m, :F - 2 * m] += r[2 * m:]
        J[m, 2 * m:] += r[:F - 2 * m]
    return J


def candidates(N, guess, xp, Z, trials=150, sigma=0.04, seed=0):
    xpf = np.array([float(v) for v in xp])
    Zf = np.array([[float(Z[i, j]) for j in range(Z.cols)] for i in range(Z.rows)])
    c0 = Zf.T @ (guess - xpf)
    rng = np.random.default_rng(seed)
    out = []
    for t in range(trials):
        c = c0 + (rng.standard_normal(c0.shape) * sigma if t else 0)
        for _ in range(200):
            r = xpf + Zf @ c
            dc = np.linalg.lstsq(ortho_jac(r) @ Zf, -ortho_resid(r), rcond=1e-12)[0]
            n = np.max(np.abs(dc))
            c = c + (min(1.0, 0.1 / n) if n > 0 else 1.0) * dc
            if n < 1e-7:
                break
        r = xpf + Zf @ c
        if np.max(np.abs(ortho_resid(r))) < 1e-7 and not any(np.max(np.abs(r - s)) < 1e-3 for s in out):
            out.append(r)
    return out


def polish(N, r0, xp, Z):
    F = 6 * N
    c = Z.T * (mp.matrix([mp.mpf(float(v)) for v in r0]) - xp)

    def resid(r):
        return mp.matrix([sum(r[k] * r[k + 2 * m] for k in range(F - 2 * m)) - (1 if m == 0 else 0) for m in range(F // 2)])

    for _ in range(60):
        r = xp + Z * c
        J = mp.matrix(F // 2, F)
        for m in range(F // 2):
            for k in range(F - 2 * m):
                J[m, k] += r[k + 2 * m]
                J[m, k + 2 * m] += r[k]
        Jr = J * Z
        dc = mp.lu_solve(Jr.T * Jr, -(Jr.T * resid(r)))
        c = c + dc
        if mp.norm(dc) < mp.mpf(10) ** -60:
            break
    r = xp + Z * c
    return [r[i] for i in range(F)], mp.norm(resid(r))


def continue_family(first=6, last=10, start=None):
    prev = np.array((start or G.RECALLED_COIF[first - 1])[::-1], dtype=float)     # rec_lo orientation
    out = {}
    for N in range(first, last + 1):
        guess = np.concatenate([[0, 0], prev, [0, 0, 0, 0]])                        # peak moves from 2(N-1) to 2N
        xp, Z = nullspace(N)
        sols = []
        for s in candidates(N, guess, xp, Z):
            r, res = polish(N, s, xp, Z)
            if res < mp.mpf(10) ** -50:
                sols.append(r)
        dist = lambda r: float(mp.sqrt(sum((a - mp.mpf(float(b))) ** 2 for a, b in zip(r, guess))))   # noqa: E731
        sols.sort(key=dist)
        uniq = [sols[0]]
        for r in sols[1:]:
            if all(max(abs(a - b) for a, b in zip(r, u)) > 1e-6 for u in uniq):
                uniq.append(r)
        print("coif%d: %d distinct solutions, distances to the padded coif%d: %s" %
              (N, len(uniq), N - 1, [round(dist(r), 4) for r in uniq[:4]]), file=sys.stderr, flush=True)
        prev = np.array([float(v) for v in uniq[0]])
        out[N] = [mp.nstr(v, 25) for v in uniq[0][::-1]]                            # dec_lo orientation
    return out


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    start = None
    if first - 1 not in G.RECALLED_COIF:
        start = [float(v) for v in G.CONTINUED_COIF[first - 1]]
    tables = continue_family(first, last, start)
    print("CONTINUED_COIF = {")
    for N, v in tables.items():
        print("    %d: [" % N)
        for i in range(0, len(v), 3):
            print("        " + ", ".join('"%s"' % x for x in v[i:i + 3]) + ",")
        print("    ],")
    print("}")
