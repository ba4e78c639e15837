#!/usr/bin/env python
"""Wall-clock of the drop-in executable `paladin-gpu` on a listing of dense n x n matrices, for a few values of
--inflight (batches of one GPU whose library calls / writers overlap, paladin_host.cpp::run_pod).

  python tools/bench_exe.py [--n 512] [--count 592] [--distinct 4] [--batch 74] [--inflight 1,2,3] [--right]

Prints one JSON line per setting: matrices/s over the executable's own "total" timer (max over pods) and over the
wall clock of the process (which includes CUDA context creation and the listing pre-pass).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paladin_b200 import host as ph  # noqa: E402


def write_dense(path, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n))
    jj, ii = np.divmod(np.arange(n * n), n)  # column-major order, like the reference's fixtures
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, n * n))
        f.write("".join("%d %d %.17g\n" % (i + 1, j + 1, a[i, j]) for i, j in zip(ii, jj)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--count", type=int, default=592)
    ap.add_argument("--distinct", type=int, default=4)
    ap.add_argument("--batch", type=int, default=74)
    ap.add_argument("--inflight", default="1,2,3")
    ap.add_argument("--right", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        names = []
        for q in range(a.distinct):
            write_dense(os.path.join(d, "m%d.matrix" % q), a.n, 100 + q)
            names.append("m%d.matrix" % q)
        with open(os.path.join(d, "listing"), "w") as f:
            f.write("\n".join(names[k % a.distinct] for k in range(a.count)))
        for k in [int(x) for x in a.inflight.split(",")]:
            cmd = [ph.exe_path(), "--listing=" + os.path.join(d, "listing"), "--rootdir=" + d, "--gpus=%d" % a.gpus,
                   "--batch=%d" % a.batch, "--inflight=%d" % k, "--measure=dim"] + (["--right"] if a.right else [])
            best = None
            for _ in range(2):  # first run warms the page cache / driver
                t0 = time.time()
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
                wall = time.time() - t0
                if r.returncode != 0:
                    print(r.stdout[-1500:], r.stderr[-1500:])
                    sys.exit(1)
                m = re.search(r"^max\t(\S+)\t(\S+)\t(\S+)", r.stdout, re.M)
                rd, eig, tot = (float(x) for x in m.groups())
                if best is None or tot < best[2]:
                    best = (rd, eig, tot, wall)
            print(json.dumps({"n": a.n, "count": a.count, "batch": a.batch, "inflight": k, "right": a.right, "gpus": a.gpus,
                              "read_s": best[0], "eig_s": best[1], "total_s": best[2], "wall_s": round(best[3], 3),
                              "matrices_per_s": round(a.count / best[2], 1)}), flush=True)


if __name__ == "__main__":
    main()
