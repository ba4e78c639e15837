#!/usr/bin/env python
"""Development aid: throughput of the parity configurations that are not the headline bench line
(BASELINE.json configs[1]: 12,800 dense 64x64, eigenvalues only; plus a mixed-size batch)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paladin_b200 as pg

def run(n, count, job, distinct=64):
    nn = n * n
    ii = np.tile(np.arange(1, n + 1, dtype=np.int32), n)
    jj = np.repeat(np.arange(1, n + 1, dtype=np.int32), n)
    vals = [np.random.default_rng([2, k]).standard_normal(nn) for k in range(distinct)]
    ci = np.tile(ii, count); cj = np.tile(jj, count)
    cv = np.concatenate([vals[k % distinct] for k in range(count)])
    nvec = np.full(count, n, dtype=np.int32)
    off = np.arange(count + 1, dtype=np.int64) * nn
    b = pg.Batch(nvec, off, jobvr=job)
    b.upload(ci, cj, cv)
    for _ in range(2):
        b.scatter(); b.solve()
    b.timer_start()
    reps = 3
    for _ in range(reps):
        b.scatter(); b.solve()
    ms = b.timer_stop() / reps
    wr, wi, vr, info = b.download()
    st = b.stage_ms()
    b.destroy()
    print("n=%d count=%d job=%s: %.2f ms/step -> %.0f decomp/s, info!=0: %d, stages %s" % (
        n, count, job, ms, count / ms * 1e3, int((info != 0).sum()), {k: round(v[0], 2) for k, v in st.items() if v[0] > 0}), flush=True)

if __name__ == "__main__":
    run(64, 12800, "N")
    run(64, 12800, "V")
    run(32, 12800, "N")
    run(128, 2048, "V")
    run(256, 592, "V")
    run(1024, 37, "V")
