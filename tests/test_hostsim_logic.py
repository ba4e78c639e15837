"""CPU-only: the group-templated device code (paladin_b200/csrc/*.cuh) instantiated on the host
(tests/hostsim) against LAPACK 3.12 (oracle/lapack_stages.py): same eigenvalue order as dlahqr for
n <= 75, tolerance parity beyond, and -- with real threads under ThreadSanitizer -- no missing barrier."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import lapack_stages as ls
from oracle import paladin_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "hostsim"))
import build as hostsim_build  # noqa: E402

dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.fixture(scope="module")
def hs():
    return ctypes.CDLL(hostsim_build.build())


def _geev(hs, A, wantvr=1):
    n = A.shape[0]
    A2 = np.array(A, order="F")
    wr, wi, V = np.zeros(n), np.zeros(n), np.zeros((n, n), order="F")
    info = hs.hs_geev(n, dp(A2), dp(wr), dp(wi), wantvr, dp(V))
    return wr, wi, V, info


def _cases(rng, n):
    A = rng.standard_normal((n, n))
    yield A, 1.0
    if n >= 5:
        B = A.copy()
        B[rng.random((n, n)) < 0.6] = 0
        B[:, 2] *= 1e5
        B[1, :] *= 1e-4
        yield B, 1.0
    yield A * 1e-200, 1e-200   # dgeev scales such input up / down before balancing
    yield A * 1e200, 1e200


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 10, 17, 33, 40, 64, 75, 100, 150])
def test_geev_pipeline_matches_lapack(hs, n):
    rng = np.random.default_rng(n)
    for q, (A, s) in enumerate(_cases(rng, n)):
        wr, wi, V, info = _geev(hs, A)
        wr0, wi0, _, VR0, info0 = ls.dgeev(A)
        assert info == 0 and info0 == 0
        A, wr, wi, wr0, wi0 = A / s, wr / s, wi / s, wr0 / s, wi0 / s  # compare in the unscaled frame
        nrm = max(1.0, np.linalg.norm(A, 2))
        w, w0 = wr + 1j * wi, wr0 + 1j * wi0
        if n <= 75 and q == 0:  # LAPACK runs dlahqr here too: identical order, pairs +imag first
            assert np.abs(w - w0).max() <= 1e-11 * nrm
        assert po.match_eigenvalues(w, w0) <= 1e-9 * nrm
        assert po.backward_error(A, wr, wi, V) <= 100
        wre, wie, _, infoe = _geev(hs, A * s, 0)
        assert infoe == 0 and po.match_eigenvalues((wre + 1j * wie) / s, w0) <= 1e-9 * nrm


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 17, 31, 32, 33, 40, 48, 63, 64])
def test_warp_register_sweep_matches_lapack(hs, n):
    """The lane-explicit Francis sweep of a single warp (pb_small.cuh::francis_sweep_lanes), 32 lanes walked by
    one host thread: whole pipeline against dgeev (same eigenvalue order as dlahqr), and lahqr alone on
    sub-blocks ilo..ihi with and without the Schur form: Z' H Z = T, T quasi-triangular, Z orthogonal."""
    rng = np.random.default_rng([11, n])
    for q, (A, s) in enumerate(_cases(rng, n)):
        A2 = np.array(A, order="F")
        wr, wi, V = np.zeros(n), np.zeros(n), np.zeros((n, n), order="F")
        assert hs.hs_geev_warp(n, dp(A2), dp(wr), dp(wi), 1, dp(V)) == 0
        wr0, wi0, _, VR0, info0 = ls.dgeev(A)
        As, w, w0 = A / s, (wr + 1j * wi) / s, (wr0 + 1j * wi0) / s
        nrm = max(1.0, np.linalg.norm(As, 2))
        if q == 0:
            assert np.abs(w - w0).max() <= 1e-11 * nrm
        assert po.match_eigenvalues(w, w0) <= 1e-9 * nrm
        assert po.backward_error(As, wr / s, wi / s, V) <= 100
        A2 = np.array(A, order="F")
        wre, wie = np.zeros(n), np.zeros(n)
        assert hs.hs_geev_warp(n, dp(A2), dp(wre), dp(wie), 0, dp(V)) == 0
        assert po.match_eigenvalues((wre + 1j * wie) / s, w0) <= 1e-9 * nrm
    for trial in range(6):
        H0 = np.triu(rng.standard_normal((n, n)), -1)
        ilo = int(rng.integers(0, max(1, n // 3)))
        ihi = int(rng.integers(max(ilo, n - 1 - n // 3), n))
        if trial == 0:
            ilo, ihi = 0, n - 1
        # isolate the active block the way dgebal / an earlier deflation would
        H0[ilo:, :ilo] = 0
        H0[ihi + 1:, :ihi + 1] = 0
        H0[np.tril_indices(n, -1)] *= (np.abs(np.subtract.outer(np.arange(n), np.arange(n))[np.tril_indices(n, -1)]) == 1)
        for wantt in (0, 1):
            H = np.array(H0, order="F")
            Z = np.eye(n, order="F")
            wr, wi = np.zeros(n), np.zeros(n)
            assert hs.hs_lahqr_warp(wantt, wantt, n, ilo, ihi, dp(H), dp(wr), dp(wi), dp(Z)) == 0
            w = (wr + 1j * wi)[ilo:ihi + 1]
            w0 = np.linalg.eigvals(H0[ilo:ihi + 1, ilo:ihi + 1])
            nrm = max(1.0, np.linalg.norm(H0, 2))
            assert po.match_eigenvalues(w, w0) <= 1e-9 * nrm
            if wantt:
                assert np.abs(Z.T @ Z - np.eye(n)).max() <= 1e-13 * n
                assert np.abs(Z.T @ H0 @ Z - H).max() <= 1e-13 * n * nrm
                assert np.abs(np.tril(H, -2)).max() == 0.0
                sub = np.abs(np.diag(H, -1))[ilo:ihi]  # subdiagonal of the active block
                assert not np.any((sub[:-1] != 0) & (sub[1:] != 0))  # 2 x 2 blocks do not touch


@pytest.mark.parametrize("n", [2, 3, 5, 16, 33, 64])
def test_two_hessenberg_matrices_share_one_array(hs, n):
    """The values-only n <= 64 kernel iterates on two Hessenberg matrices stored in one array (the second transposed,
    pb::lahqr hrs / band_only): neither run may read or write a cell of the other (poisoned cells outside both)."""
    rng = np.random.default_rng([21, n])
    H0 = np.asfortranarray(np.triu(rng.standard_normal((n, n)), -1))
    H1 = np.asfortranarray(np.triu(rng.standard_normal((n, n)), -1))
    w = [np.zeros(n) for _ in range(4)]
    assert hs.hs_lahqr_pair(n, dp(H0), dp(H1), dp(w[0]), dp(w[1]), dp(w[2]), dp(w[3])) == 0
    for H, wr, wi in ((H0, w[0], w[1]), (H1, w[2], w[3])):
        assert po.match_eigenvalues(wr + 1j * wi, np.linalg.eigvals(H)) <= 1e-9 * max(1.0, np.linalg.norm(H, 2))


def _hessenberg_case(n, seed, sparse=False):
    rng = np.random.default_rng([seed, n])
    A = rng.standard_normal((n, n))
    if sparse:
        A[rng.random((n, n)) < 0.7] = 0
    B, ilo, ihi, sc = ls.dgebal(np.asfortranarray(A))
    H, tau = ls.dgehrd(B, ilo, ihi)
    Q = ls.dorghr(H, tau, ilo, ihi)
    # what dgehrd leaves below the subdiagonal (reflector storage) must be ignored by the QR stage
    Hj = np.asfortranarray(np.triu(H, -1) + np.tril(rng.standard_normal((n, n)), -2))
    return B, ilo, ihi, np.asfortranarray(np.triu(H, -1)), Hj, Q


# last entry: chains + 100 * shifts per sweep + 10000 * AED window order (0 = no AED)
MS_CONFIGS = [(8, 49, 48, 128, 1), (8, 49, 48, 256, 2), (4, 17, 16, 7, 1), (2, 10, 6, 5, 3), (16, 97, 64, 64, 1),
              (8, 49, 48, 128, 2 + 100 * 32 + 10000 * 48), (4, 17, 16, 7, 1 + 10000 * 16), (2, 10, 6, 5, 3 + 10000 * 9),
              (8, 49, 20, 64, 2 + 10000 * 30)]


@pytest.mark.parametrize("n", [10, 49, 50, 80, 130, 200])
@pytest.mark.parametrize("cfg", MS_CONFIGS)
def test_multishift_qr_matches_lapack_dhseqr(hs, n, cfg):
    nb, W, wss, tl, nch = cfg
    for sparse in (False, True):
        B, ilo, ihi, H, Hj, Q = _hessenberg_case(n, 5, sparse)
        T0, wr0, wi0, Z0, info0 = ls.dhseqr(H, ilo, ihi, Z=Q)
        sl = slice(ilo - 1, ihi)
        for wantt in (1, 0):
            H2, Z2 = Hj.copy(order="F"), Q.copy(order="F")
            wr, wi = np.zeros(n), np.zeros(n)
            info = hs.hs_msqr(wantt, wantt, n, ilo - 1, ihi - 1, dp(H2), dp(wr), dp(wi), dp(Z2), nb, W, wss, tl, nch)
            assert info == 0 and info0 == 0
            nrm = max(1.0, np.linalg.norm(B, 2))
            assert po.match_eigenvalues(wr[sl] + 1j * wi[sl], wr0[sl] + 1j * wi0[sl]) <= 1e-9 * nrm
            if wantt:
                T = np.triu(H2, -1)
                assert np.abs(Z2 @ T @ Z2.T - B).max() <= 1e-12 * nrm * n
                assert np.abs(Z2.T @ Z2 - np.eye(n)).max() <= 1e-12 * n
                sub = np.diag(T, -1)
                for k in np.nonzero(sub)[0]:  # standardised 2x2 blocks (dlanv2), never two in a row
                    assert T[k, k] == T[k + 1, k + 1] and T[k, k + 1] * T[k + 1, k] < 0
                    assert k + 1 >= n - 1 or sub[k + 1] == 0


def test_block_swaps_match_lapack_dtrexc(hs):
    """pb::aed_trexc_up (dtrexc/dlaexc restated, pb_aed.cuh) against LAPACK: same eigenvalue order, orthogonal
    accumulated factor, similarity preserved."""
    from scipy.linalg import schur
    rng = np.random.default_rng(123)
    moved = 0
    for n in (2, 3, 5, 8, 13, 24, 48):
        for trial in range(6):
            A = rng.standard_normal((n, n))
            if trial >= 4:
                A = np.triu(A)  # real spectrum: 1x1 blocks only
            T, Q = schur(A, output="real")
            T = np.asfortranarray(T); Q = np.asfortranarray(Q)
            starts = [k for k in range(n) if k == 0 or T[k, k - 1] == 0]
            for ifst in starts[1:]:
                for ilst in starts:
                    if ilst >= ifst:
                        continue
                    T2, Q2 = T.copy(order="F"), Q.copy(order="F")
                    at = hs.hs_trexc_up(n, dp(T2), dp(Q2), ifst, ilst)
                    assert at == ilst
                    assert np.abs(np.tril(T2, -2)).max() == 0 if n > 2 else True
                    sub = np.diag(T2, -1)
                    for k in np.nonzero(sub)[0]:
                        assert T2[k, k] == T2[k + 1, k + 1] and T2[k, k + 1] * T2[k + 1, k] < 0
                        assert k + 1 >= n - 1 or sub[k + 1] == 0
                    nrm = max(1.0, np.abs(A).max())
                    assert np.abs(Q2 @ T2 @ Q2.T - A).max() <= 1e-13 * nrm * n
                    assert np.abs(Q2.T @ Q2 - np.eye(n)).max() <= 1e-13 * n
                    # the moved block's eigenvalue(s) now sit at ilst; the blocks it passed keep their relative order
                    def blocks(M):
                        out, k = [], 0
                        while k < n:
                            if k + 1 < n and M[k + 1, k] != 0:
                                out.append(complex(M[k, k], np.sqrt(abs(M[k, k + 1] * M[k + 1, k])))); k += 2
                            else:
                                out.append(complex(M[k, k], 0.0)); k += 1
                        return out
                    b0, b2 = blocks(T), blocks(T2)
                    i0, i1 = starts.index(ifst), starts.index(ilst)
                    want = b0[:i1] + [b0[i0]] + b0[i1:i0] + b0[i0 + 1:]
                    assert len(want) == len(b2)
                    assert np.abs(np.array(want) - np.array(b2)).max() <= 1e-10 * nrm * n
                    moved += 1
    assert moved > 200


def test_aed_reduces_sweeps(hs):
    """AED (pb_aed.cuh) must deflate most eigenvalues and at least halve the number of sweeps at n = 256."""
    n = 256
    B, ilo, ihi, H, Hj, Q = _hessenberg_case(n, 9)
    hs.hs_ms_counters.restype = ctypes.POINTER(ctypes.c_int)
    T0, wr0, wi0, Z0, info0 = ls.dhseqr(H, ilo, ihi, Z=Q)
    sweeps = {}
    for aed in (0, 48):
        H2, Z2 = Hj.copy(order="F"), Q.copy(order="F")
        wr, wi = np.zeros(n), np.zeros(n)
        assert hs.hs_msqr(1, 1, n, ilo - 1, ihi - 1, dp(H2), dp(wr), dp(wi), dp(Z2), 8, 49, 48, 128, 2 + 100 * 32 + 10000 * aed) == 0
        c = hs.hs_ms_counters()
        sweeps[aed] = c[1]
        if aed:
            assert c[12] >= n // 2
        nrm = max(1.0, np.linalg.norm(B, 2))
        assert po.match_eigenvalues(wr + 1j * wi, wr0 + 1j * wi0) <= 1e-9 * nrm
        T = np.triu(H2, -1)
        assert np.abs(Z2 @ T @ Z2.T - B).max() <= 1e-12 * nrm * n
    assert 2 * sweeps[48] <= sweeps[0]


def test_multishift_geev_pipeline(hs):
    rng = np.random.default_rng(77)
    for n in (65, 100, 257):
        A = rng.standard_normal((n, n))
        for wantvr in (1, 0):
            A2 = np.array(A, order="F")
            wr, wi, V = np.zeros(n), np.zeros(n), np.zeros((n, n), order="F")
            info = hs.hs_geev_ms(n, dp(A2), dp(wr), dp(wi), wantvr, dp(V), 8, 49, 48, 256, 2)
            wr0, wi0, _, VR0, info0 = ls.dgeev(A)
            assert info == 0
            assert po.match_eigenvalues(wr + 1j * wi, wr0 + 1j * wi0) <= 1e-9 * max(1.0, np.linalg.norm(A, 2))
            if wantvr:
                assert po.backward_error(A, wr, wi, V) <= 100


def test_left_eigenvectors_match_lapack(hs):
    rng = np.random.default_rng(31)
    for n in (1, 2, 3, 6, 20, 50, 90):
        A = rng.standard_normal((n, n))
        if n >= 6:
            A[rng.random((n, n)) < 0.4] = 0
            A[:, 2] *= 1e6
        A = np.asfortranarray(A)
        B, ilo, ihi, sc = ls.dgebal(A)
        H, tau = ls.dgehrd(B, ilo, ihi)
        Q = ls.dorghr(H, tau, ilo, ihi)
        T, wr, wi, Z, info = ls.dhseqr(np.triu(H, -1), ilo, ihi, Z=Q)
        V0 = ls.dtrevc3(T, Z, side=b"L")
        V = Z.copy(order="F")
        hs.hs_trevc_left(n, dp(np.asfortranarray(T)), dp(V))
        assert np.abs(V - V0).max() <= 1e-10 * max(1.0, np.abs(V0).max())
        Vb0 = ls.dgebak(V0, ilo, ihi, sc, side=b"L")
        sc2 = sc.copy()
        for k in list(range(0, ilo - 1)) + list(range(ihi, n)):
            sc2[k] -= 1
        hs.hs_gebak_left(n, ilo - 1, ihi - 1, dp(sc2), dp(V))
        assert np.abs(V - Vb0).max() <= 1e-10 * max(1.0, np.abs(Vb0).max())
        # full pipeline, both sides, against dgeev
        A2 = np.array(A, order="F")
        w1, w2 = np.zeros(n), np.zeros(n)
        VL, VR = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
        assert hs.hs_geev_lr(n, dp(A2), dp(w1), dp(w2), 1, dp(VL), 1, dp(VR)) == 0
        wr0, wi0, VL0, VR0, info0 = ls.dgeev(A, jobvl=b"V", jobvr=b"V")
        assert po.match_eigenvalues(w1 + 1j * w2, wr0 + 1j * wi0) <= 1e-9 * max(1.0, np.linalg.norm(A, 2))
        assert po.backward_error(A, w1, w2, VR) <= 100 and po.left_backward_error(A, w1, w2, VL) <= 100


def test_balancing_with_precomputed_sums_matches_lapack(hs):
    """pb::gebal with its work array (row / column sums of a sweep taken in two coalesced passes until the first rescaling,
    used by the prep kernels from 128 active rows on) against LAPACK dgebal: same ilo/ihi, same power-of-two scale
    factors, same balanced matrix -- for inputs that need no scaling, one rescaling, and many."""
    rng = np.random.default_rng(12)
    for n in (130, 200, 300):
        for kind in range(4):
            A = rng.standard_normal((n, n))
            if kind == 1:
                A[:, 7] *= 1e6
            elif kind == 2:
                A *= np.outer(2.0 ** rng.integers(-20, 20, n), 2.0 ** rng.integers(-20, 20, n))
            elif kind == 3:
                A[rng.random((n, n)) < 0.97] = 0      # isolated rows / columns: ilo > 1, ihi < n
                A[:, 3] *= 1e5
            A = np.asfortranarray(A)
            B, ilo, ihi, sc = ls.dgebal(A)
            for threaded in (False, True):
                A2 = A.copy(order="F")
                ilo2, ihi2, sc2 = ctypes.c_int(), ctypes.c_int(), np.zeros(n)
                if threaded:
                    hs.ht_gebal_work(4, n, dp(A2), ctypes.byref(ilo2), ctypes.byref(ihi2), dp(sc2))
                else:
                    hs.hs_gebal_work(n, dp(A2), ctypes.byref(ilo2), ctypes.byref(ihi2), dp(sc2))
                assert (ilo2.value, ihi2.value) == (ilo - 1, ihi - 1), (n, kind)
                scl = sc.copy()
                for k in list(range(0, ilo - 1)) + list(range(ihi, n)):
                    scl[k] -= 1
                assert np.array_equal(scl, sc2), (n, kind, threaded)
                assert np.array_equal(B, A2), (n, kind, threaded)


def test_stage_functions_match_lapack(hs):
    rng = np.random.default_rng(3)
    for n in (6, 20, 50):
        A = rng.standard_normal((n, n))
        A[rng.random((n, n)) < 0.5] = 0
        A[:, 2] *= 1e6
        A = np.asfortranarray(A)
        B, ilo, ihi, sc = ls.dgebal(A)
        A2 = A.copy(order="F")
        ilo2, ihi2, sc2 = ctypes.c_int(), ctypes.c_int(), np.zeros(n)
        hs.hs_gebal(n, dp(A2), ctypes.byref(ilo2), ctypes.byref(ihi2), dp(sc2))
        assert (ilo2.value, ihi2.value) == (ilo - 1, ihi - 1)
        assert np.array_equal(B, A2)
        scl = sc.copy()
        for k in list(range(0, ilo - 1)) + list(range(ihi, n)):
            scl[k] -= 1
        assert np.array_equal(scl, sc2)
        # trevc on LAPACK's own Schur form
        H, tau = ls.dgehrd(B, ilo, ihi)
        Q = ls.dorghr(H, tau, ilo, ihi)
        T, wr, wi, Z, info = ls.dhseqr(np.triu(H, -1), ilo, ihi, Z=Q)
        V0 = ls.dtrevc3(T, Z)
        V = Z.copy(order="F")
        hs.hs_trevc_right(n, dp(np.asfortranarray(T)), dp(V))
        assert np.abs(V - V0).max() <= 1e-10 * max(1.0, np.abs(V0).max())
        Vb0 = ls.dgebak(V0, ilo, ihi, sc)
        hs.hs_gebak_right(n, ilo - 1, ihi - 1, dp(sc2), dp(V))
        assert np.abs(V - Vb0).max() <= 1e-10 * max(1.0, np.abs(Vb0).max())


def test_dlaln2_near_singular_and_scaling(hs):
    # tiny pivots are perturbed to smin, huge right-hand sides are scaled: x solves (A - wI) x = scale*b
    a = np.array([1.0, 0.5, -0.25, 1.0])
    for na, nw, wr, wi in [(1, 1, 1.0, 0.0), (1, 2, 1.0, 1e-300), (2, 1, 0.3, 0.0), (2, 2, 0.3, 0.7), (2, 2, 1.0, 0.0)]:
        b = np.array([1e300, -2e300, 3e300, 1e299])
        x = np.zeros(4)
        scale, xnorm = ctypes.c_double(), ctypes.c_double()
        hs.hs_dlaln2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p]
        hs.hs_dlaln2(0, na, nw, 1e-8, a.ctypes.data, b.ctypes.data, wr, wi, x.ctypes.data, ctypes.byref(scale),
                     ctypes.byref(xnorm))
        assert np.all(np.isfinite(x)) and 0 < scale.value <= 1.0


@pytest.mark.timeout(900)
def test_threaded_groups_are_race_free_under_tsan():
    """4-5 real threads per group, pthread barrier behind sync(): ThreadSanitizer reports any access pair
    that is not ordered by a barrier, i.e. a missing __syncthreads()/__syncwarp() in the device code."""
    so = hostsim_build.build(tsan=True)
    tsan_rt = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan_rt):
        pytest.skip("libtsan not available")
    code = r"""
import ctypes, numpy as np, sys
sys.path.insert(0, %r)
from oracle import lapack_stages as ls, paladin_oracle as po
hs = ctypes.CDLL(%r)
dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
rng = np.random.default_rng(1)
for n, nt, (nb, W, wss, tl, nch) in [(30, 4, (2, 10, 6, 5, 2)), (40, 3, (3, 14, 8, 6, 1)), (26, 5, (2, 11, 7, 4, 3)), (44, 4, (2, 13, 8, 6, 2 + 10000 * 12)), (36, 5, (3, 14, 6, 5, 1 + 10000 * 9))]:
    A = rng.standard_normal((n, n))
    for wantvr in (0, 1):
        A2 = np.array(A, order='F'); wr = np.zeros(n); wi = np.zeros(n); V = np.zeros((n, n), order='F')
        info = hs.ht_geev_ms(nt, n, dp(A2), dp(wr), dp(wi), wantvr, dp(V), nb, W, wss, tl, nch)
        wr0, wi0, _, VR0, info0 = ls.dgeev(A)
        assert info == 0
        assert po.match_eigenvalues(wr + 1j * wi, wr0 + 1j * wi0) <= 1e-9 * max(1, np.linalg.norm(A, 2))
        if wantvr: assert po.backward_error(A, wr, wi, V) <= 100
for n, nt in [(5, 3), (14, 4)]:
    A = rng.standard_normal((n, n))
    A2 = np.array(A, order='F'); wr = np.zeros(n); wi = np.zeros(n)
    VL = np.zeros((n, n), order='F'); VR = np.zeros((n, n), order='F')
    assert hs.ht_geev_lr(nt, n, dp(A2), dp(wr), dp(wi), 1, dp(VL), 1, dp(VR)) == 0
    assert po.backward_error(A, wr, wi, VR) <= 100 and po.left_backward_error(A, wr, wi, VL) <= 100
for n, nt in [(1, 3), (2, 3), (3, 4), (7, 4), (12, 3), (20, 5)]:
    for trial in range(2):
        A = rng.standard_normal((n, n))
        if trial == 1 and n >= 5:
            A[rng.random((n, n)) < 0.6] = 0; A[:, 2] *= 1e5; A[1, :] *= 1e-4
        for wantvr in (0, 1):
            A2 = np.array(A, order='F'); wr = np.zeros(n); wi = np.zeros(n); V = np.zeros((n, n), order='F')
            info = hs.ht_geev(nt, n, dp(A2), dp(wr), dp(wi), wantvr, dp(V))
            wr0, wi0, _, VR0, info0 = ls.dgeev(A)
            assert info == 0
            assert po.match_eigenvalues(wr + 1j * wi, wr0 + 1j * wi0) <= 1e-9 * max(1, np.linalg.norm(A, 2))
            if wantvr: assert po.backward_error(A, wr, wi, V) <= 100
import ctypes as _ct
for n, nt in [(140, 4)]:
    A = np.asfortranarray(rng.standard_normal((n, n))); A[:, 5] *= 1e4
    lo, hi, sc = _ct.c_int(), _ct.c_int(), np.zeros(n)
    hs.ht_gebal_work(nt, n, dp(A), _ct.byref(lo), _ct.byref(hi), dp(sc))
print('DONE')
""" % (os.path.dirname(HERE), so)
    env = dict(os.environ, LD_PRELOAD=tsan_rt, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=850)
    assert "DONE" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ThreadSanitizer: data race" not in r.stderr, r.stderr[:6000]
