"""
ctypes bindings of include/paladin_gpu.h. Nothing here computes: every numerical entry forwards to
libpaladin_gpu.so and raises PaladinGpuError when the library reports a failure (no CPU fallback).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NUM_STAGES = 8
STAGE_NAMES = ("scatter", "balance", "hessenberg", "qr", "vectors", "small", "generic", "mid")
PHASE_NAMES = ("balance", "hessenberg", "form_q", "qr_total", "trevc", "bak_normalise", "vec_gemm", "hess_update", "ms_scan", "ms_shifts",
               "ms_chase", "ms_strips", "ms_small", "ms_small_u", "ms_aed_window", "ms_aed_update")

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)
_c_flt_p = ctypes.POINTER(ctypes.c_float)


class PaladinGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("paladin_gpu error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    # PALADIN_GPU_LIB: another build of the SAME library (tuning variants compiled with different -D switches, see
    # tools/build_variants.sh); never a different implementation
    return os.environ.get("PALADIN_GPU_LIB") or os.path.join(_HERE, "lib", "libpaladin_gpu.so")


def lib():
    """Loads libpaladin_gpu.so (built by __graft_entry__.build() / paladin_b200/csrc/Makefile)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise PaladinGpuError(-1, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback)" % p)
    L = ctypes.CDLL(p)
    L.paladin_gpu_device_count.restype = ctypes.c_int
    L.paladin_gpu_init.argtypes = [ctypes.c_int]
    L.paladin_gpu_shutdown.argtypes = [ctypes.c_int]
    L.paladin_gpu_last_error.restype = ctypes.c_char_p
    L.paladin_gpu_version.restype = ctypes.c_char_p
    L.paladin_gpu_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    L.paladin_gpu_host_free.argtypes = [ctypes.c_void_p]
    L.paladin_gpu_geev_batch.argtypes = [ctypes.c_int, ctypes.c_int, _c_int_p, _c_i64_p, _c_int_p, _c_int_p, _c_dbl_p,
                                         ctypes.c_char, ctypes.c_char, ctypes.POINTER(_c_dbl_p),
                                         ctypes.POINTER(_c_dbl_p), ctypes.POINTER(_c_dbl_p), ctypes.POINTER(_c_dbl_p),
                                         _c_int_p]
    L.paladin_gpu_batch_create.argtypes = [ctypes.c_int, ctypes.c_int, _c_int_p, _c_i64_p, ctypes.c_char,
                                           ctypes.c_char, ctypes.POINTER(ctypes.c_void_p)]
    L.paladin_gpu_batch_upload.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_scatter.argtypes = [ctypes.c_void_p]
    L.paladin_gpu_batch_upload_dense.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_get_dense.argtypes = [ctypes.c_void_p, ctypes.c_int, _c_dbl_p]
    L.paladin_gpu_batch_solve.argtypes = [ctypes.c_void_p]
    L.paladin_gpu_batch_download.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_stage_ms.argtypes = [ctypes.c_void_p, _c_flt_p, _c_int_p]
    L.paladin_gpu_batch_destroy.argtypes = [ctypes.c_void_p]
    L.paladin_gpu_batch_vector_filter.argtypes = [ctypes.c_void_p, ctypes.c_char, ctypes.c_double, ctypes.c_void_p,
                                                  ctypes.c_void_p]
    L.paladin_gpu_batch_upload_text.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_upload_coo_one.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_get_coo.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_format_vectors.argtypes = [ctypes.c_void_p, ctypes.c_char, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.paladin_gpu_batch_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    L.paladin_gpu_batch_timer_start.argtypes = [ctypes.c_void_p]
    L.paladin_gpu_batch_timer_stop.argtypes = [ctypes.c_void_p, _c_flt_p]
    L.paladin_gpu_timer_start.argtypes = [ctypes.c_int]
    L.paladin_gpu_timer_stop.argtypes = [ctypes.c_int, _c_flt_p]
    L.paladin_gpu_dgeev_.restype = None
    L.paladin_gpu_dgeev_.argtypes = [ctypes.c_char_p, ctypes.c_char_p, _c_int_p, _c_dbl_p, _c_int_p, _c_dbl_p, _c_dbl_p,
                                     _c_dbl_p, _c_int_p, _c_dbl_p, _c_int_p, _c_dbl_p, _c_int_p, _c_int_p]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise PaladinGpuError(rc, lib().paladin_gpu_last_error().decode())


def version():
    return lib().paladin_gpu_version().decode()


def device_count():
    return lib().paladin_gpu_device_count()


def timer_start(device=0):
    _check(lib().paladin_gpu_timer_start(device))


def timer_stop(device=0):
    ms = ctypes.c_float()
    _check(lib().paladin_gpu_timer_stop(device, ctypes.byref(ms)))
    return float(ms.value)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class PinnedArray:
    """numpy view over page-locked host memory from paladin_gpu_host_alloc."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        _check(lib().paladin_gpu_host_alloc(ctypes.byref(p), self.nbytes))
        self._p = p
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().paladin_gpu_host_free(self._p)
            self._p = None


class Batch:
    """Staged API: create -> upload -> scatter -> solve -> download (include/paladin_gpu.h)."""

    def __init__(self, n, nnz_offsets, jobvl="N", jobvr="N", device=0):
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.nnz_offsets = np.ascontiguousarray(nnz_offsets, dtype=np.int64)
        assert len(self.nnz_offsets) == len(self.n) + 1
        self.count = len(self.n)
        self.jobvr = jobvr
        self.jobvl = jobvl
        self.total_n = int(self.n.astype(np.int64).sum())
        self.total_nn = int((self.n.astype(np.int64) ** 2).sum())
        h = ctypes.c_void_p()
        _check(lib().paladin_gpu_batch_create(device, self.count, self.n.ctypes.data_as(_c_int_p),
                                              self.nnz_offsets.ctypes.data_as(_c_i64_p), jobvl.encode(), jobvr.encode(),
                                              ctypes.byref(h)))
        self._h = h

    def upload(self, coo_i, coo_j, coo_v):
        assert coo_i.dtype == np.int32 and coo_j.dtype == np.int32 and coo_v.dtype == np.float64
        _check(lib().paladin_gpu_batch_upload(self._h, _ptr(coo_i), _ptr(coo_j), _ptr(coo_v)))

    def upload_text(self, text, text_off):
        """Raw MatrixMarket entry lines of the batch (bytes / uint8 array) -> COO on the device. Returns needs_host[count]."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, dtype=np.uint8)
        off = np.ascontiguousarray(text_off, dtype=np.int64)
        assert len(off) == self.count + 1
        need = np.zeros(max(self.count, 1), dtype=np.int32)
        _check(lib().paladin_gpu_batch_upload_text(self._h, ctypes.c_void_p(buf.ctypes.data) if len(buf) else None, _ptr(off), _ptr(need)))
        return need[:self.count]

    def upload_coo_one(self, k, coo_i, coo_j, coo_v):
        coo_i = np.ascontiguousarray(coo_i, dtype=np.int32)
        coo_j = np.ascontiguousarray(coo_j, dtype=np.int32)
        coo_v = np.ascontiguousarray(coo_v, dtype=np.float64)
        _check(lib().paladin_gpu_batch_upload_coo_one(self._h, int(k), _ptr(coo_i), _ptr(coo_j), _ptr(coo_v)))

    def get_coo(self):
        nnz = int(self.nnz_offsets[-1] - self.nnz_offsets[0])
        i = np.zeros(nnz, dtype=np.int32)
        j = np.zeros(nnz, dtype=np.int32)
        v = np.zeros(nnz, dtype=np.float64)
        _check(lib().paladin_gpu_batch_get_coo(self._h, _ptr(i), _ptr(j), _ptr(v)))
        return i, j, v

    def format_vectors(self, side="R", threshold=1e-3):
        """-> (list of body bytes per matrix, kept[count], needs_host[count]) of the eigenvector files, formatted on the device."""
        cap = 48 * max(self.total_nn, 1) + 64
        text = np.zeros(cap, dtype=np.uint8)
        off = np.zeros(self.count + 1, dtype=np.int64)
        kept = np.zeros(max(self.count, 1), dtype=np.int64)
        need = np.zeros(max(self.count, 1), dtype=np.int32)
        _check(lib().paladin_gpu_batch_format_vectors(self._h, side.encode(), float(threshold), 0, self.count, _ptr(text), cap, _ptr(off), _ptr(kept),
                                                      _ptr(need)))
        bodies = [text[off[k]:off[k + 1]].tobytes() for k in range(self.count)]
        return bodies, kept[:self.count], need[:self.count]

    def scatter(self):
        _check(lib().paladin_gpu_batch_scatter(self._h))

    def upload_dense(self, dense):
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        _check(lib().paladin_gpu_batch_upload_dense(self._h, _ptr(dense)))

    def get_dense(self, k):
        out = np.empty(int(self.n[k]) ** 2, dtype=np.float64)
        _check(lib().paladin_gpu_batch_get_dense(self._h, k, out.ctypes.data_as(_c_dbl_p)))
        return out

    def solve(self):
        _check(lib().paladin_gpu_batch_solve(self._h))

    def download(self, wr=None, wi=None, vr=None, info=None, vl=None):
        wr = np.empty(self.total_n) if wr is None else wr
        wi = np.empty(self.total_n) if wi is None else wi
        info = np.empty(self.count, dtype=np.int32) if info is None else info
        if self.jobvr == "V" and vr is None:
            vr = np.empty(self.total_nn)
        if self.jobvl == "V" and vl is None:
            vl = np.empty(self.total_nn)
        _check(lib().paladin_gpu_batch_download(self._h, _ptr(wr), _ptr(wi), _ptr(vl) if self.jobvl == "V" else None,
                                                _ptr(vr) if self.jobvr == "V" else None, _ptr(info)))
        self.vl = vl  # left vectors of the last download (None unless jobvl == 'V')
        return wr, wi, vr, info

    def vector_filter(self, side="R", threshold=1e-3):
        """-> (max |x|^2 per eigenvalue index, entries kept per matrix) of the reference's eigenvector writer, on the device."""
        max2 = np.zeros(self.total_n)
        kept = np.zeros(self.count, dtype=np.int64)
        _check(lib().paladin_gpu_batch_vector_filter(self._h, side.encode(), float(threshold), _ptr(max2), _ptr(kept)))
        return max2, kept

    def stage_ms(self):
        ms = (ctypes.c_float * NUM_STAGES)()
        ln = (ctypes.c_int * NUM_STAGES)()
        _check(lib().paladin_gpu_batch_stage_ms(self._h, ms, ln))
        return {STAGE_NAMES[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_STAGES)}

    def timer_start(self):
        _check(lib().paladin_gpu_batch_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float()
        _check(lib().paladin_gpu_batch_timer_stop(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def phase_cycles(self):
        cyc = (ctypes.c_longlong * 16)()
        _check(lib().paladin_gpu_batch_phase_cycles(self._h, cyc))
        return {PHASE_NAMES[i]: int(cyc[i]) for i in range(16) if PHASE_NAMES[i]}

    def destroy(self):
        if self._h is not None:
            lib().paladin_gpu_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def split_results(n, wr, wi, vr):
    """Concatenated outputs -> per-matrix (wr, wi, VR[n,n] Fortran-ordered or None)."""
    out = []
    o1 = 0
    o2 = 0
    for nk in n:
        nk = int(nk)
        V = None
        if vr is not None:
            V = vr[o2:o2 + nk * nk].reshape((nk, nk), order="F")
        out.append((wr[o1:o1 + nk], wi[o1:o1 + nk], V))
        o1 += nk
        o2 += nk * nk
    return out


def geev_batch(n, nnz_offsets, coo_i, coo_j, coo_v, jobvr="N", device=0, jobvl="N"):
    """One-shot entry paladin_gpu_geev_batch with host buffers. Returns (list of (wr, wi, VR), info); with
    jobvl='V' the list holds (wr, wi, VR, VL)."""
    n = np.ascontiguousarray(n, dtype=np.int32)
    off = np.ascontiguousarray(nnz_offsets, dtype=np.int64)
    coo_i = np.ascontiguousarray(coo_i, dtype=np.int32)
    coo_j = np.ascontiguousarray(coo_j, dtype=np.int32)
    coo_v = np.ascontiguousarray(coo_v, dtype=np.float64)
    count = len(n)
    wr = [np.zeros(int(k)) for k in n]
    wi = [np.zeros(int(k)) for k in n]
    vr = [np.zeros((int(k), int(k)), order="F") if jobvr == "V" else None for k in n]
    vl = [np.zeros((int(k), int(k)), order="F") if jobvl == "V" else None for k in n]
    info = np.zeros(max(count, 1), dtype=np.int32)
    PP = _c_dbl_p * max(count, 1)
    pwr = PP(*[a.ctypes.data_as(_c_dbl_p) for a in wr])
    pwi = PP(*[a.ctypes.data_as(_c_dbl_p) for a in wi])
    pvr = PP(*[(a.ctypes.data_as(_c_dbl_p) if a is not None else None) for a in vr])
    pvl = PP(*[(a.ctypes.data_as(_c_dbl_p) if a is not None else None) for a in vl])
    _check(lib().paladin_gpu_geev_batch(device, count, n.ctypes.data_as(_c_int_p), off.ctypes.data_as(_c_i64_p),
                                        coo_i.ctypes.data_as(_c_int_p), coo_j.ctypes.data_as(_c_int_p),
                                        coo_v.ctypes.data_as(_c_dbl_p), jobvl.encode(), jobvr.encode(), pwr, pwi,
                                        pvl if jobvl == "V" else None, pvr if jobvr == "V" else None,
                                        info.ctypes.data_as(_c_int_p)))
    if jobvl == "V":
        return list(zip(wr, wi, vr, vl)), info[:count]
    return list(zip(wr, wi, vr)), info[:count]


def dgeev(A, jobvr="V", jobvl="N"):
    """Single dense matrix through the Fortran-signature shim paladin_gpu_dgeev_ (query + solve, exactly the
    two calls of reference src/paladin.hpp:247-254). Returns wr, wi, VR, info (and VL as a fifth value when
    jobvl='V')."""
    A = np.array(A, dtype=np.float64, order="F")
    n = A.shape[0]
    ld = max(n, 1)
    wr, wi = np.zeros(ld), np.zeros(ld)
    VL = np.zeros((ld, ld), order="F")
    VR = np.zeros((ld, ld), order="F")
    wk = np.zeros(1)
    info = ctypes.c_int(0)
    ci = lambda v: ctypes.byref(ctypes.c_int(v))
    dp = lambda a: a.ctypes.data_as(_c_dbl_p)
    f = lib().paladin_gpu_dgeev_
    f(jobvl.encode(), jobvr.encode(), ci(n), dp(A), ci(ld), dp(wr), dp(wi), dp(VL), ci(ld), dp(VR), ci(ld), dp(wk), ci(-1),
      ctypes.byref(info))
    lwork = max(1, int(wk[0]))
    work = np.zeros(lwork)
    f(jobvl.encode(), jobvr.encode(), ci(n), dp(A), ci(ld), dp(wr), dp(wi), dp(VL), ci(ld), dp(VR), ci(ld), dp(work), ci(lwork),
      ctypes.byref(info))
    if jobvl == "V":
        return wr[:n], wi[:n], VR[:n, :n], info.value, VL[:n, :n]
    return wr[:n], wi[:n], VR[:n, :n], info.value
