"""
lapack_stages.py -- TEST INFRASTRUCTURE ONLY. Not part of the product.

Stage-level oracles: ctypes bindings to the individual LAPACK 3.12.0 routines that
`dgeev_` (the reference's only numerical call, /root/reference/src/paladin.hpp:248-254)
is built from, taken from the SAME library the unmodified reference is linked against
in oracle/_ref (scipy's bundled OpenBLAS 0.3.31.dev, LP64, symbols prefixed `scipy_`).
Used by tests/ to check each device stage (balance, Hessenberg, orghr, QR iteration,
eigenvector solve) separately; never imported from paladin_b200/.
"""
import ctypes
import glob
import os

import numpy as np

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        import scipy
        d = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
        _LIB = ctypes.CDLL(sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so")))[0])
    return _LIB


def _i(x):
    return ctypes.byref(ctypes.c_int(x))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def set_threads(n):
    lib().scipy_openblas_set_num_threads(int(n))


def dgebal(A, job=b"B"):
    """-> (A_balanced, ilo, ihi (1-based), scale)"""
    A = np.array(A, dtype=np.float64, order="F")
    n = A.shape[0]
    ilo, ihi, info = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    scale = np.zeros(max(n, 1))
    lib().scipy_dgebal_(job, _i(n), _dp(A), _i(max(n, 1)), ctypes.byref(ilo), ctypes.byref(ihi), _dp(scale),
                        ctypes.byref(info), 1)
    assert info.value == 0
    return A, ilo.value, ihi.value, scale[:n]


def dgehrd(A, ilo, ihi):
    """ilo/ihi 1-based. -> (A_out with reflectors, tau[n-1])"""
    A = np.array(A, dtype=np.float64, order="F")
    n = A.shape[0]
    tau = np.zeros(max(n - 1, 1))
    lwork = max(1, 64 * n)
    work = np.zeros(lwork)
    info = ctypes.c_int()
    lib().scipy_dgehrd_(_i(n), _i(ilo), _i(ihi), _dp(A), _i(max(n, 1)), _dp(tau), _dp(work), _i(lwork),
                        ctypes.byref(info))
    assert info.value == 0
    return A, tau[:max(n - 1, 0)]


def dorghr(A, tau, ilo, ihi):
    Q = np.array(A, dtype=np.float64, order="F")
    n = Q.shape[0]
    lwork = max(1, 64 * n)
    work = np.zeros(lwork)
    info = ctypes.c_int()
    t = np.zeros(max(n - 1, 1))
    t[:len(tau)] = tau
    lib().scipy_dorghr_(_i(n), _i(ilo), _i(ihi), _dp(Q), _i(max(n, 1)), _dp(t), _dp(work), _i(lwork),
                        ctypes.byref(info))
    assert info.value == 0
    return Q


def dhseqr(H, ilo, ihi, job=b"S", compz=b"V", Z=None):
    """-> (T, wr, wi, Z, info)"""
    H = np.array(H, dtype=np.float64, order="F")
    n = H.shape[0]
    wr, wi = np.zeros(max(n, 1)), np.zeros(max(n, 1))
    if Z is None:
        Z = np.eye(n)
    Z = np.array(Z, dtype=np.float64, order="F")
    lwork = max(1, 64 * n)
    work = np.zeros(lwork)
    info = ctypes.c_int()
    lib().scipy_dhseqr_(job, compz, _i(n), _i(ilo), _i(ihi), _dp(H), _i(max(n, 1)), _dp(wr), _dp(wi), _dp(Z),
                        _i(max(n, 1)), _dp(work), _i(lwork), ctypes.byref(info), 1, 1)
    return H, wr[:n], wi[:n], Z, info.value


def dlahqr(H, ilo, ihi, wantt=True, wantz=True, Z=None):
    H = np.array(H, dtype=np.float64, order="F")
    n = H.shape[0]
    wr, wi = np.zeros(max(n, 1)), np.zeros(max(n, 1))
    if Z is None:
        Z = np.eye(n)
    Z = np.array(Z, dtype=np.float64, order="F")
    info = ctypes.c_int()
    lib().scipy_dlahqr_(_i(1 if wantt else 0), _i(1 if wantz else 0), _i(n), _i(ilo), _i(ihi), _dp(H), _i(max(n, 1)),
                        _dp(wr), _dp(wi), _i(1), _i(n), _dp(Z), _i(max(n, 1)), ctypes.byref(info))
    return H, wr[:n], wi[:n], Z, info.value


def dtrevc3(T, VR=None, side=b"R", howmny=b"B"):
    """Right (or left) eigenvectors of quasi-triangular T, back-transformed by VR when howmny='B'."""
    T = np.array(T, dtype=np.float64, order="F")
    n = T.shape[0]
    V = np.array(np.eye(n) if VR is None else VR, dtype=np.float64, order="F")
    dummy = np.zeros((1, 1), order="F")
    sel = np.zeros(max(n, 1), dtype=np.int32)
    lwork = max(1, n + 2 * n * 64)
    work = np.zeros(lwork)
    m, info = ctypes.c_int(), ctypes.c_int()
    if side == b"R":
        lib().scipy_dtrevc3_(side, howmny, _ip(sel), _i(n), _dp(T), _i(max(n, 1)), _dp(dummy), _i(1), _dp(V),
                             _i(max(n, 1)), _i(n), ctypes.byref(m), _dp(work), _i(lwork), ctypes.byref(info), 1, 1)
    else:
        lib().scipy_dtrevc3_(side, howmny, _ip(sel), _i(n), _dp(T), _i(max(n, 1)), _dp(V), _i(max(n, 1)), _dp(dummy),
                             _i(1), _i(n), ctypes.byref(m), _dp(work), _i(lwork), ctypes.byref(info), 1, 1)
    assert info.value == 0
    return V


def dgebak(V, ilo, ihi, scale, job=b"B", side=b"R"):
    V = np.array(V, dtype=np.float64, order="F")
    n = V.shape[0]
    s = np.array(scale, dtype=np.float64)
    info = ctypes.c_int()
    lib().scipy_dgebak_(job, side, _i(n), _i(ilo), _i(ihi), _dp(s), _i(n), _dp(V), _i(max(n, 1)), ctypes.byref(info),
                        1, 1)
    assert info.value == 0
    return V


def dgeev(A, jobvl=b"N", jobvr=b"V"):
    """Raw dgeev_ exactly as the reference calls it (workspace query, then solve)."""
    A = np.array(A, dtype=np.float64, order="F")
    n = A.shape[0]
    ld = max(n, 1)
    wr, wi = np.zeros(ld), np.zeros(ld)
    VL = np.zeros((ld, ld), order="F")
    VR = np.zeros((ld, ld), order="F")
    wk = np.zeros(1)
    info = ctypes.c_int()
    f = lib().scipy_dgeev_
    f(jobvl, jobvr, _i(n), _dp(A), _i(ld), _dp(wr), _dp(wi), _dp(VL), _i(ld), _dp(VR), _i(ld), _dp(wk), _i(-1),
      ctypes.byref(info), 1, 1)
    lwork = max(1, int(wk[0]))
    work = np.zeros(lwork)
    f(jobvl, jobvr, _i(n), _dp(A), _i(ld), _dp(wr), _dp(wi), _dp(VL), _i(ld), _dp(VR), _i(ld), _dp(work), _i(lwork),
      ctypes.byref(info), 1, 1)
    return wr[:n], wi[:n], VL[:n, :n], VR[:n, :n], info.value
