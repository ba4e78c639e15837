"""
ctypes access to the CPU-side host library (paladin_b200/lib/libpaladin_host.so, built from
paladin_b200/host/paladin_host.cpp): the reference's reader, load measures, balancer and writers as the GPU path
uses them. Used by tests/ and by bench.py to shard the listing across ranks exactly as the paladin-gpu
executable shards it across GPUs.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def exe_path():
    return os.path.join(_HERE, "lib", "paladin-gpu")


def lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(_HERE, "lib", "libpaladin_host.so")
        if not os.path.exists(p):
            raise RuntimeError("%s is missing: run __graft_entry__.build()" % p)
        L = ctypes.CDLL(p)
        L.ph_parse_mm.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p]
        L.ph_measure.restype = ctypes.c_double
        L.ph_measure.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        L.ph_partition.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.ph_format.restype = ctypes.c_long
        L.ph_format.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int,
                                ctypes.c_char_p, ctypes.c_long]
        _LIB = L
    return _LIB


def parse_mm(path):
    n = ctypes.c_int()
    nnz = lib().ph_parse_mm(path.encode(), ctypes.byref(n), 0, None, None, None)
    if nnz < 0:
        raise ValueError("cannot parse %s" % path)
    ii = np.zeros(nnz, dtype=np.int32)
    jj = np.zeros(nnz, dtype=np.int32)
    vv = np.zeros(nnz, dtype=np.float64)
    lib().ph_parse_mm(path.encode(), ctypes.byref(n), nnz, ii.ctypes.data, jj.ctypes.data, vv.ctypes.data)
    return n.value, nnz, ii, jj, vv


def parse_batch(paths, threads=4):
    """The executable's batch parser (paladin_host.cpp::parse_batch): files -> concatenated COO triplets."""
    L = lib()
    L.ph_parse_batch.restype = ctypes.c_longlong
    L.ph_parse_batch.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    blob = "\n".join(paths).encode()
    total = L.ph_parse_batch(blob, len(paths), threads, None, None, 0, None, None, None)
    if total < 0:
        raise ValueError("cannot parse the batch")
    n = np.zeros(len(paths), dtype=np.int32)
    off = np.zeros(len(paths) + 1, dtype=np.int64)
    ii = np.zeros(total, dtype=np.int32)
    jj = np.zeros(total, dtype=np.int32)
    vv = np.zeros(total, dtype=np.float64)
    L.ph_parse_batch(blob, len(paths), threads, n.ctypes.data, off.ctypes.data, total, ii.ctypes.data, jj.ctypes.data, vv.ctypes.data)
    return n, off, ii, jj, vv


def measure(n, nnz, kind="nnz"):
    return lib().ph_measure(int(n), int(nnz), kind.encode())


def partition(measures, pods):
    """-> (order, colour) of the SORTED positions, exactly as the paladin-gpu executable assigns pods."""
    m = np.ascontiguousarray(measures, dtype=np.float64)
    order = np.zeros(len(m), dtype=np.int32)
    colour = np.zeros(len(m), dtype=np.int32)
    lib().ph_partition(len(m), m.ctypes.data, int(pods), order.ctypes.data, colour.ctypes.data)
    return order, colour


def pod_of_rank(measures, world, rank):
    """Listing indices owned by `rank` (in sorted order), the N>1 sharding rule of bench.py."""
    order, colour = partition(measures, world)
    return order[colour == rank]


def format_output(wr, wi, vr, which, threshold=1e-3):
    """which: 0 spectrum file, 1 right-eigenvector file -> bytes"""
    n = len(wr)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    wi = np.ascontiguousarray(wi, dtype=np.float64)
    v = None if vr is None else np.asfortranarray(vr, dtype=np.float64)
    vp = None if v is None else v.ctypes.data
    need = lib().ph_format(n, wr.ctypes.data, wi.ctypes.data, vp, threshold, which, None, 0)
    buf = ctypes.create_string_buffer(max(need, 1))
    lib().ph_format(n, wr.ctypes.data, wi.ctypes.data, vp, threshold, which, buf, need)
    return buf.raw[:need]
