#!/usr/bin/env python
"""Development aid: a few large matrices (BASELINE.json configs[4]: 3072x3072 full decomposition, streamed subset)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paladin_b200 as pg

def run(n, count, job="V"):
    nn = n * n
    ii = np.tile(np.arange(1, n + 1, dtype=np.int32), n)
    jj = np.repeat(np.arange(1, n + 1, dtype=np.int32), n)
    vals = [np.random.default_rng([5, k]).standard_normal(nn) for k in range(count)]
    nvec = np.full(count, n, dtype=np.int32)
    off = np.arange(count + 1, dtype=np.int64) * nn
    b = pg.Batch(nvec, off, jobvr=job)
    b.upload(np.tile(ii, count), np.tile(jj, count), np.concatenate(vals))
    b.timer_start()
    b.scatter(); b.solve()
    ms = b.timer_stop()
    wr, wi, vr, info = b.download()
    st = b.stage_ms()
    cyc = b.phase_cycles()
    b.destroy()
    A = vals[0].reshape((n, n), order="F")
    V = vr[:nn].reshape((n, n), order="F")
    lam = wr[:n] + 1j * wi[:n]
    # residual of a few pairs
    res = []
    k = 0
    while k < n and len(res) < 6:
        if wi[k] == 0:
            v = V[:, k]; res.append(np.linalg.norm(A @ v - wr[k] * v)); k += 1
        else:
            v = V[:, k] + 1j * V[:, k + 1]; res.append(np.linalg.norm(A @ v - lam[k] * v)); k += 2
    tr = abs(lam.sum() - np.trace(A)) / max(1.0, abs(np.trace(A)))
    print("n=%d count=%d: %.1f ms -> %.2f decomp/s, info %s, max residual/(n eps ||A||) %.2f, trace err %.1e, stages %s" % (
        n, count, ms, count / ms * 1e3, info.tolist(), max(res) / (n * 2.2e-16 * np.linalg.norm(A, 2)), tr,
        {k_: round(v_[0], 1) for k_, v_ in st.items() if v_[0] > 0}), flush=True)
    print("   phase Mcycles (summed over recording CTAs):", {k_: round(v_ / 1e6, 1) for k_, v_ in cyc.items() if v_}, flush=True)

if __name__ == "__main__":
    cases = [(1536, 8), (3072, 4)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    for n, count in cases:
        run(n, count)
