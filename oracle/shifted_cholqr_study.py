"""NumPy study behind cuda_compute._shifted_cholesky_r (test infrastructure, not product code).

Iterated shifted CholeskyQR (shifted CholeskyQR3 of Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa, SIAM J. Sci.
Comput. 42 (2020), repeated while the condition bound stays large) with the decision rule of the device path: the
same bound sqrt(|L|_1 |L|_inf |L^-1|_1 |L^-1|_inf), explicit triangular inverses, accept below 30, plain pass below
1e7, at most three shifted passes, shift 11 n (n + 1) u trace(G) escalated by factors of 100 up to the theoretical
11 (m n + n (n + 1)) u trace(G) when Cholesky breaks down.  Prints, for graded spectra up to cond 1e16, rank-deficient and zero blocks, the
backward error |R^T R - A^T A| / |A^T A|, the distance to LAPACK's Householder R (rows sign-canonical) and the pass
sequence.   python -m oracle.shifted_cholqr_study
"""
import numpy as np


def graded(m, n, kappa, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return (U * np.logspace(0, -np.log10(kappa), n)) @ V.T


def factor(G):
    try:
        L = np.linalg.cholesky(G)
    except np.linalg.LinAlgError:
        return None, None, np.inf
    Li = np.linalg.inv(L)
    k = np.sqrt(np.linalg.norm(L, 1) * np.linalg.norm(L, np.inf) * np.linalg.norm(Li, 1) * np.linalg.norm(Li, np.inf))
    return L.T, Li.T, k


def shifted_cholesky_r(A, accept=30.0, plain=1e7, max_shift=3):
    m, n = A.shape
    u = 2.0 ** -53
    cur, racc, shifts, log = A, np.eye(n), 0, []
    for _ in range(max_shift + 3):
        G = cur.T @ cur
        R, Ri, k = factor(G)
        if k <= accept:
            log.append("done")
            return R @ racc, log
        if k > plain:
            if shifts >= max_shift:
                log.append("bail")
                return None, log
            shifts += 1
            trace = np.trace(G)
            if not np.isfinite(trace) or trace <= 0:
                log.append("bail")
                return None, log
            full = 11.0 * (m * n + n * (n + 1)) * u          # the theoretical shift; start smaller, escalate on breakdown
            c = min(11.0 * n * (n + 1) * u, full)
            while True:
                R, Ri, ks = factor(G + c * trace * np.eye(n))
                if np.isfinite(ks) or c >= full:
                    break
                c = min(c * 100.0, full)
                log.append("retry")
            log.append("shift")
            if not np.isfinite(ks):
                return None, log
        else:
            log.append("plain")
        racc = R @ racc
        cur = cur @ Ri
    log.append("bail")
    return None, log


def main():
    for (m, n, kappa) in [(40000, 64, 1e7), (40000, 64, 1e9), (40000, 64, 1e13), (200000, 128, 1e9), (200000, 128, 1e12),
                          (40000, 64, 1e15), (40000, 64, 1e16)]:
        X = graded(m, n, kappa, 3)
        R, log = shifted_cholesky_r(X)
        G = X.T @ X
        if R is None:
            print(m, n, "%.0e" % kappa, "no convergence", log)
            continue
        Rw = np.linalg.qr(X, mode="r")
        sg = np.sign(np.diag(R)) * np.sign(np.diag(Rw))
        print(m, n, "%.0e" % kappa, "gram err %.2e   R vs LAPACK %.2e" % (
            np.linalg.norm(R.T @ R - G) / np.linalg.norm(G), np.linalg.norm(R * sg[:, None] - Rw) / np.linalg.norm(Rw)), log)
    rng = np.random.default_rng(77)
    X = rng.standard_normal((30000, 32))
    X[:, 7] = X[:, 3]
    X[:, 20] = 0.5 * X[:, 1] - 2.0 * X[:, 2]
    R, log = shifted_cholesky_r(X)
    G = X.T @ X
    print("rank deficient:", log, "gram err %.2e" % (np.linalg.norm(R.T @ R - G) / np.linalg.norm(G)))
    print("zeros:", shifted_cholesky_r(np.zeros((4096, 16)))[1])


if __name__ == "__main__":
    main()
