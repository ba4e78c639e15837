"""
paladin_oracle.py -- TEST INFRASTRUCTURE ONLY. Not part of the product.

CPU oracle for the paladin hot path (MatrixMarket COO -> dense -> dgeev -> unpack
-> spectrum / eigenvector text). Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it, and only as the
checker or the timed CPU arm -- never from paladin_b200/.

What is restated here and what it follows (citations into /root/reference):

* parse / scatter / measures / partition: thin numpy wrappers over the C
  restatement in oracle/paladin_oracle.cpp (src/square-matrix.hpp:101-163,
  :236-268; src/measure-util.hpp:104-111; src/paladin.hpp:141-155).
* dgeev: the arithmetic of the path lives in a third-party library that is NOT
  in the reference tree ("system LAPACK", CMakeLists.txt:98-100, no version
  pinned). The oracle instance is the LAPACK 3.12.0 bundled in scipy's OpenBLAS
  0.3.31.dev (scipy 1.18.x wheel, symbol scipy_dgeev_), i.e. the very library
  oracle/_ref links the unmodified reference against. It is reached through
  scipy.linalg.lapack.dgeev with the reference's call convention
  (src/paladin.hpp:247-254: jobvl/jobvr in {'N','V'}, lda=ldvl=ldvr=N).
* unpack: real-packed LAPACK output -> complex (src/paladin.hpp:256-333).
* writers: src/spectrum-util.hpp:119-124 (spectrum) and :134-196 (vectors),
  default ostream precision 6 == printf("%g").

Pinned by tests/test_oracle_golden.py against every *.exact file of the
reference's ctest suite (copied as fixtures into tests/golden/ref_ctest) and
against vectors produced by the unmodified reference binary built in
oracle/_ref (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MEASURE_CODES = {"dim": 0, "nnz": 1, "dcb": 2, "zds": 3, "sps": 4, "spc": 5}


def _lib():
    """Load (building on first use) the C restatement."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "_ref", "libpaladin_oracle.so")
    src = os.path.join(_HERE, "paladin_oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    c_int_p = ctypes.POINTER(ctypes.c_int)
    c_dbl_p = ctypes.POINTER(ctypes.c_double)
    lib.oracle_mm_header.restype = ctypes.c_long
    lib.oracle_mm_header.argtypes = [ctypes.c_char_p, ctypes.c_long, c_int_p, c_int_p]
    lib.oracle_mm_entries.restype = ctypes.c_int
    lib.oracle_mm_entries.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                      c_int_p, c_int_p, c_dbl_p]
    lib.oracle_scatter_dense.restype = None
    lib.oracle_scatter_dense.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p]
    lib.oracle_measure.restype = ctypes.c_double
    lib.oracle_measure.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.oracle_partition.restype = None
    lib.oracle_partition.argtypes = [ctypes.c_int, c_dbl_p, ctypes.c_int, c_int_p, c_int_p]
    _LIB = lib
    return lib


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------------------------
# MatrixMarket parse + dense scatter   (reference src/square-matrix.hpp:101-180)
# ----------------------------------------------------------------------------------------------
def parse_mm(text):
    """bytes -> (n, nnz, i[int32], j[int32], v[float64]) in file order, 1-based indices."""
    if isinstance(text, str):
        text = text.encode()
    n = ctypes.c_int()
    nnz = ctypes.c_int()
    off = _lib().oracle_mm_header(text, len(text), ctypes.byref(n), ctypes.byref(nnz))
    if off < 0:
        raise ValueError("malformed MatrixMarket header (code %d)" % off)
    ii = np.zeros(nnz.value, dtype=np.int32)
    jj = np.zeros(nnz.value, dtype=np.int32)
    vv = np.zeros(nnz.value, dtype=np.float64)
    got = _lib().oracle_mm_entries(text, len(text), off, nnz.value, _ip(ii), _ip(jj), _dp(vv))
    if got != nnz.value:
        raise ValueError("file holds %d entries, size line promised %d" % (got, nnz.value))
    return n.value, nnz.value, ii, jj, vv


def read_mm(path):
    with open(path, "rb") as f:
        return parse_mm(f.read())


def scatter_dense(n, ii, jj, vv):
    """Zero-filled column-major n*n array (1-D, Fortran order), last writer wins."""
    ii = np.ascontiguousarray(ii, dtype=np.int32)
    jj = np.ascontiguousarray(jj, dtype=np.int32)
    vv = np.ascontiguousarray(vv, dtype=np.float64)
    dense = np.empty(n * n, dtype=np.float64)
    _lib().oracle_scatter_dense(n, len(vv), _ip(ii), _ip(jj), _dp(vv), _dp(dense))
    return dense


def dense_matrix(n, ii, jj, vv):
    """Same as scatter_dense but shaped (n, n), Fortran-ordered."""
    return scatter_dense(n, ii, jj, vv).reshape((n, n), order="F")


# ----------------------------------------------------------------------------------------------
# load measures + static balancer   (src/square-matrix.hpp:236-268, src/paladin.hpp:141-155)
# ----------------------------------------------------------------------------------------------
def measure(n, nnz, kind="nnz"):
    return _lib().oracle_measure(int(n), int(nnz), MEASURE_CODES[kind])


def partition(measures, P):
    """-> (order, color): listing index and pod of each SORTED position."""
    m = np.ascontiguousarray(measures, dtype=np.float64)
    order = np.zeros(len(m), dtype=np.int32)
    color = np.zeros(len(m), dtype=np.int32)
    _lib().oracle_partition(len(m), _dp(m), int(P), _ip(order), _ip(color))
    return order, color


# ----------------------------------------------------------------------------------------------
# dgeev (third-party LAPACK, see module docstring) with the reference's call convention
# ----------------------------------------------------------------------------------------------
def dgeev(A, left=False, right=False):
    """A: (n,n) array. Returns wr, wi, VL, VR (n,n Fortran; zeros when not requested), info."""
    from scipy.linalg import lapack
    A = np.array(A, dtype=np.float64, order="F")
    n = A.shape[0]
    wr, wi, vl, vr, info = lapack.dgeev(A, compute_vl=1 if left else 0, compute_vr=1 if right else 0,
                                        overwrite_a=1)
    VL = np.asfortranarray(vl) if left else np.zeros((n, n), order="F")
    VR = np.asfortranarray(vr) if right else np.zeros((n, n), order="F")
    return wr, wi, VL, VR, info


# ----------------------------------------------------------------------------------------------
# eigenpair unpack   (src/paladin.hpp:256-333)
# ----------------------------------------------------------------------------------------------
def unpack(wr, wi, V):
    """Real-packed LAPACK vectors -> list of n complex vectors, following the reference loop
    exactly (pair detected iff wi[k] == -wi[k+1] and wi[k] != 0; last column handled apart)."""
    n = len(wr)
    vecs = []
    skip = False
    for c in range(n - 1):
        if not skip:
            if wi[c] == -wi[c + 1] and wi[c] != 0:
                vecs.append(V[:, c] + 1j * V[:, c + 1])
                # the reference builds the conjugate as (re, -im): keeps signed zeros
                vecs.append(V[:, c] + 1j * (-V[:, c + 1]))
                skip = True
            else:
                vecs.append(V[:, c] + 1j * np.zeros(n))
                skip = False
        else:
            skip = False
    if not skip:
        vecs.append(V[:, n - 1] + 1j * np.zeros(n))
    return vecs


def eigenvalues(wr, wi):
    return np.asarray(wr) + 1j * np.asarray(wi)


# ----------------------------------------------------------------------------------------------
# writers   (src/spectrum-util.hpp:119-124, :134-196)
# ----------------------------------------------------------------------------------------------
def _g(x):
    return "%g" % x


def format_spectrum(wr, wi):
    return "".join("%s %s\n" % (_g(a), _g(b)) for a, b in zip(wr, wi)).encode()


def format_vectors(parts, threshold=1.0e-3):
    """parts: list of (re, im) float arrays per eigenvector (kept separate so that signed
    zeros survive). Entry kept iff |x|^2 > threshold * max|x|^2 (squared magnitudes against the
    un-squared threshold, as the reference does)."""
    n = len(parts)
    lines = []
    count = 0
    for c, (re, im) in enumerate(parts):
        mag = re * re + im * im
        big = mag.max() if len(mag) else 0.0
        for r in range(len(re)):
            if mag[r] > threshold * big:
                count += 1
                lines.append("%d %d %s %s\n" % (r + 1, c + 1, _g(re[r]), _g(im[r])))
    head = "%%MatrixMarket matrix coordinate complex general\n" + "%d %d %d\n" % (n, n, count)
    return (head + "".join(lines)).encode()


def unpack_parts(wr, wi, V):
    """Like unpack() but returns (re, im) float arrays so -0.0 imaginary parts print as '-0'."""
    n = len(wr)
    parts = []
    skip = False
    zero = np.zeros(n)
    for c in range(n - 1):
        if not skip:
            if wi[c] == -wi[c + 1] and wi[c] != 0:
                parts.append((V[:, c].copy(), V[:, c + 1].copy()))
                parts.append((V[:, c].copy(), -V[:, c + 1]))
                skip = True
            else:
                parts.append((V[:, c].copy(), zero.copy()))
                skip = False
        else:
            skip = False
    if not skip:
        parts.append((V[:, n - 1].copy(), zero.copy()))
    return parts


# ----------------------------------------------------------------------------------------------
# comparison tools used by the parity tests
# ----------------------------------------------------------------------------------------------
def match_eigenvalues(w_test, w_ref):
    """Conjugate-aware pairing: both spectra are sorted by (real, |imag|, sign of imag) after
    which a greedy nearest-neighbour pass repairs order flips among near-equal real parts.
    Returns max_k |w_test[p(k)] - w_ref[k]| for the best pairing found."""
    a = np.array(w_test, dtype=np.complex128)
    b = np.array(w_ref, dtype=np.complex128)
    if len(a) != len(b):
        return np.inf
    if len(a) == 0:
        return 0.0
    if not (np.all(np.isfinite(a)) and np.all(np.isfinite(b))):
        return np.inf
    used = np.zeros(len(a), dtype=bool)
    worst = 0.0
    # exact assignment would be Hungarian; nearest-unused is enough at the 1e-9 level because
    # eigenvalue gaps of the test matrices are many orders larger than the tolerance.
    order = np.argsort(-np.abs(b.imag) - np.abs(b.real))
    for k in order:
        d = np.abs(a - b[k])
        d[used] = np.inf
        p = int(np.argmin(d))
        used[p] = True
        worst = max(worst, float(d[p]))
    return worst


def backward_error(A, wr, wi, VR):
    """max over eigenpairs of ||A v - lambda v||_2 / (n ||A||_1 eps ||v||_2), v from real-packed VR."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    eps = np.finfo(np.float64).eps
    vecs = unpack(wr, wi, VR)
    lam = eigenvalues(wr, wi)
    V = np.array(vecs).T
    R = A @ V - V * lam[None, :]
    nrmA = max(np.linalg.norm(A, 1), np.finfo(np.float64).tiny)
    vn = np.linalg.norm(V, axis=0)
    if np.any(vn == 0) or not np.all(np.isfinite(V)):
        return np.inf
    return float(np.max(np.linalg.norm(R, axis=0) / vn) / (n * nrmA * eps))


def left_backward_error(A, wr, wi, VL):
    """Same for left vectors: ||u^H A - lambda u^H|| / (n ||A|| eps ||u||)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    eps = np.finfo(np.float64).eps
    vecs = unpack(wr, wi, VL)
    lam = eigenvalues(wr, wi)
    U = np.array(vecs).T
    R = U.conj().T @ A - lam[:, None] * U.conj().T
    nrmA = max(np.linalg.norm(A, 1), np.finfo(np.float64).tiny)
    un = np.linalg.norm(U, axis=0)
    if np.any(un == 0) or not np.all(np.isfinite(U)):
        return np.inf
    return float(np.max(np.linalg.norm(R, axis=1) / un) / (n * nrmA * eps))


# ----------------------------------------------------------------------------------------------
# synthetic workloads of SURVEY.md section 8(d): text format "%d %d %.17g\n"
# ----------------------------------------------------------------------------------------------
def random_dense_coo(n, seed):
    """Dense iid N(0,1) matrix as COO triplets in column-major file order."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    jj, ii = np.meshgrid(np.arange(1, n + 1, dtype=np.int32), np.arange(1, n + 1, dtype=np.int32))
    # column-major traversal: i fastest
    ii = np.ascontiguousarray(ii.T.reshape(-1))
    jj = np.ascontiguousarray(jj.T.reshape(-1))
    vv = np.ascontiguousarray(A.T.reshape(-1))
    return ii, jj, vv


def random_sparse_coo(n, density, seed):
    """nnz = round(density*n*n) distinct positions, values N(0,1), shuffled order."""
    rng = np.random.default_rng(seed)
    nnz = max(1, int(round(density * n * n)))
    pos = rng.choice(n * n, size=nnz, replace=False)
    ii = (pos % n + 1).astype(np.int32)
    jj = (pos // n + 1).astype(np.int32)
    vv = rng.standard_normal(nnz)
    return ii, jj, vv


def mm_text(n, ii, jj, vv, comments=()):
    head = "%%MatrixMarket matrix coordinate real general\n"
    head += "".join("%" + c + "\n" for c in comments)
    head += "%d %d %d\n" % (n, n, len(vv))
    body = "".join("%d %d %.17g\n" % (i, j, v) for i, j, v in zip(ii, jj, vv))
    return (head + body).encode()
