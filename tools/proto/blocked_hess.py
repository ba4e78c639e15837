"""numpy prototype of the blocked Hessenberg reduction + blocked Q exactly as pb_bhess.cuh schedules it
(dlahr2 panels, Ytop GEMM, unified column tiles with R/L updates, backward WY accumulation of Q).
Development aid: validates index conventions against scipy before the CUDA transcription."""
import numpy as np
import scipy.linalg as sl

NB = 32

def vrule(A, k, r, p):
    """V(r, p) of the panel starting at column k: unit lower trapezoidal, stored below the subdiagonal of A."""
    c = k + p
    if r == c + 1:
        return 1.0
    if r > c + 1:
        return A[r, c]
    return 0.0

def vmat(A, k, ib, rows):
    return np.array([[vrule(A, k, r, p) for p in range(ib)] for r in rows])

def larfg(alpha, x):
    xn = np.linalg.norm(x)
    if xn == 0.0:
        return alpha, x * 0, 0.0
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    tau = (beta - alpha) / beta
    return beta, x / (alpha - beta), tau

def blocked_hess(A, ilo, ihi):
    n = A.shape[0]
    A = A.copy()
    tau = np.zeros(n)
    Ts = {}
    k = ilo
    last = ihi - 2          # last reflector column
    while k <= last:
        ib = min(NB, last - k + 1)
        Y = np.zeros((n, ib))
        T = np.zeros((ib, ib))
        for j in range(ib):
            c = k + j
            rows = np.arange(k + 1, ihi + 1)
            b = A[rows, c].copy()
            if j > 0:
                vrow = np.array([vrule(A, k, c, p) for p in range(j)])
                b -= Y[rows, :j] @ vrow
                V = vmat(A, k, j, rows)
                w = V.T @ b
                w = T[:j, :j].T @ w
                b -= V @ w
            # reflector from b: alpha at row c+1
            ia = c + 1 - (k + 1)
            beta, v, t = larfg(b[ia], b[ia + 1:])
            tau[c] = t
            A[rows[:ia], c] = b[:ia]
            A[c + 1, c] = beta
            A[c + 2:ihi + 1, c] = v
            vs = np.zeros(n)
            vs[c + 1] = 1.0
            vs[c + 2:ihi + 1] = v
            y = A[k + 1:ihi + 1, c + 1:ihi + 1] @ vs[c + 1:ihi + 1]
            if j > 0:
                Vc = vmat(A, k, j, np.arange(c + 1, ihi + 1))
                w2 = Vc.T @ vs[c + 1:ihi + 1]
                y -= Y[k + 1:ihi + 1, :j] @ w2
                T[:j, j] = -t * (T[:j, :j] @ w2)
            T[j, j] = t
            Y[k + 1:ihi + 1, j] = t * y
        Ts[k] = (ib, T.copy())
        # Ytop
        Vall = vmat(A, k, ib, np.arange(k + 1, ihi + 1))
        VT = Vall @ T
        Y[:k + 1, :] = A[:k + 1, k + 1:ihi + 1] @ VT
        # column tiles: panel columns (top rows, R only), trailing columns (R all rows + L)
        for cc in range(k + 1, n):
            vrow = np.array([vrule(A, k, cc, p) for p in range(ib)])
            if cc <= k + ib - 1:
                A[:k + 1, cc] -= Y[:k + 1, :] @ vrow
            else:
                if cc <= ihi:
                    A[:ihi + 1, cc] -= Y[:ihi + 1, :] @ vrow
                w = Vall.T @ A[k + 1:ihi + 1, cc]
                w = T.T @ w
                A[k + 1:ihi + 1, cc] -= Vall @ w
        k += ib
    return A, tau, Ts

def blocked_q(A, ilo, ihi, Ts):
    n = A.shape[0]
    Q = np.eye(n)
    for k in sorted(Ts.keys(), reverse=True):
        ib, T = Ts[k]
        Vall = vmat(A, k, ib, np.arange(k + 1, ihi + 1))
        for cc in range(k + 1, ihi + 1):
            w = Vall.T @ Q[k + 1:ihi + 1, cc]
            w = T @ w
            Q[k + 1:ihi + 1, cc] -= Vall @ w
    return Q

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, ilo, ihi in [(70, 0, 69), (131, 3, 120), (40, 0, 39), (33, 0, 32), (5, 0, 4), (3, 0, 2), (64, 10, 12)]:
        A0 = rng.standard_normal((n, n))
        # emulate a balanced matrix: zero below the diagonal outside ilo..ihi
        for c in range(n):
            for r in range(c + 1, n):
                if c < ilo or r > ihi:
                    A0[r, c] = 0.0
        A, tau, Ts = blocked_hess(A0, ilo, ihi)
        H = np.triu(A, -1)
        Q = blocked_q(A, ilo, ihi, Ts)
        e1 = np.linalg.norm(Q.T @ A0 @ Q - H) / np.linalg.norm(A0)
        e2 = np.linalg.norm(Q.T @ Q - np.eye(n))
        # LAPACK's own reflectors for comparison (same algorithm => same H up to rounding)
        H2, Q2 = sl.hessenberg(A0, calc_q=True)
        print(n, ilo, ihi, "resid %.2e orth %.2e" % (e1, e2), "| vs scipy H diff %.2e" % np.linalg.norm(np.abs(H) - np.abs(H2)) if ilo == 0 and ihi == n - 1 else "")
