#!/usr/bin/env python
"""Development aid: stage 1 alone (PALADIN_GPU_DEBUG_STOP_AFTER_PREP) on a few orders; prints ||Q^T Q - I|| and
||Q^T A Q - H|| / ||A|| per matrix. usage: check_hess.py n1 n2 ..."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PALADIN_GPU_DEBUG_STOP_AFTER_PREP"] = "1"
import paladin_b200 as pg
from oracle import paladin_oracle as po

sizes = [int(x) for x in sys.argv[1:]] or [1026, 1300, 1537, 2050]
mats = [(n,) + po.random_dense_coo(n, [93, q]) for q, n in enumerate(sizes)]
n = np.array([m[0] for m in mats], dtype=np.int32)
off = np.concatenate([[0], np.cumsum([len(m[3]) for m in mats])]).astype(np.int64)
b = pg.Batch(n, off, jobvr="V")
b.upload(np.concatenate([m[1] for m in mats]), np.concatenate([m[2] for m in mats]), np.concatenate([m[3] for m in mats]))
b.scatter()
b.solve()
Hs = [b.get_dense(k) for k in range(len(mats))]
wr, wi, vr, info = b.download()
o = 0
for m, Hraw in zip(mats, Hs):
    nk = m[0]
    A = po.dense_matrix(*m)
    Hfull = Hraw.reshape((nk, nk), order="F")
    H = np.triu(Hfull, -1)
    Q = vr[o:o + nk * nk].reshape((nk, nk), order="F")
    o += nk * nk
    nrm = np.linalg.norm(A, 2)
    # where does H deviate from scipy's? (first bad column)
    import scipy.linalg as sl
    H2 = sl.hessenberg(A)
    dcol = np.abs(np.abs(H) - np.abs(H2)).max(axis=0)
    badc = np.nonzero(dcol > 1e-8 * nrm)[0]
    print("n=%d orth %.2e resid %.2e first bad column %s (of %d bad) nan=%d" % (
        nk, np.linalg.norm(Q.T @ Q - np.eye(nk)), np.linalg.norm(Q.T @ A @ Q - H) / nrm, badc[:1], len(badc), int(np.isnan(Hfull).sum())))
