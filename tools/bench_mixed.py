#!/usr/bin/env python
"""Mixed-size heterogeneous batch (BASELINE.json configs[3], SURVEY.md 8d "Config 4"): M matrices, n log-uniform in
[32, 3072], COO density uniform in [0.01, 1], values N(0,1), matrix k from default_rng([4, k]); sharded over the
GPUs by the reference's own static balancer (paladin_b200.host.partition = measures -> std::sort -> greedy colouring)
with the chosen load measure, full decomposition (eigenvalues + right eigenvectors) through the staged C-ABI.

  python tools/bench_mixed.py [--per-gpu 256] [--measure dcb] [--nmax 3072]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_mixed.py ...

Prints one JSON line (rank 0): decompositions/s over all GPUs (max-over-ranks device time), FP64 work rate by the
26.33 n^3 convention, per-rank busy time and the reference's load-imbalance statistic (max / mean - 1).
A development / parity-configuration tool, not the headline bench (bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paladin_b200 as pg  # noqa: E402
from paladin_b200 import host as ph  # noqa: E402


def matrix_shape(k, nmin, nmax):
    rng = np.random.default_rng([4, k, 0])
    n = int(round(np.exp(rng.uniform(np.log(nmin), np.log(nmax)))))
    dens = float(rng.uniform(0.01, 1.0))
    return n, max(1, int(round(dens * n * n)))


def matrix_coo(k, n, nnz, shuffle):
    rng = np.random.default_rng([4, k, 1])
    nn = n * n
    if nnz >= nn:
        pos = np.arange(nn, dtype=np.int64)
    else:
        pos = np.sort(rng.choice(nn, size=nnz, replace=False)) if nn <= 4_000_000 else np.flatnonzero(rng.random(nn) < nnz / nn)
    if shuffle:
        rng.shuffle(pos)
    vals = rng.standard_normal(len(pos))
    return (pos % n + 1).astype(np.int32), (pos // n + 1).astype(np.int32), vals


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=256)
    ap.add_argument("--measure", default="dcb", choices=["nnz", "dim", "dcb", "zds", "sps", "spc"])
    ap.add_argument("--nmin", type=int, default=32)
    ap.add_argument("--nmax", type=int, default=3072)
    ap.add_argument("--shuffle", action="store_true", help="entries in shuffled order (order-exact scatter fallback)")
    ap.add_argument("--job", default="V", choices=["N", "V"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M = args.per_gpu * world
    shapes = [matrix_shape(k, args.nmin, args.nmax) for k in range(M)]
    measures = [ph.measure(n, nnz, args.measure) for n, nnz in shapes]
    mine = [int(k) for k in ph.pod_of_rank(measures, world, rank)]
    t0 = time.time()
    coo = [matrix_coo(k, shapes[k][0], shapes[k][1], args.shuffle) for k in mine]
    nvec = np.array([shapes[k][0] for k in mine], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum([len(c[2]) for c in coo])]).astype(np.int64)
    ii = np.concatenate([c[0] for c in coo]); jj = np.concatenate([c[1] for c in coo]); vv = np.concatenate([c[2] for c in coo])
    gen_s = time.time() - t0
    b = pg.Batch(nvec, off, jobvr=args.job, device=local)
    b.upload(ii, jj, vv)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    b.timer_start()
    b.scatter()
    b.solve()
    ms = b.timer_stop()
    out = b.download()
    info = out[-1]
    st = b.stage_ms()
    b.destroy()
    fl = float(sum((26.33 if args.job == "V" else 10.0) * float(shapes[k][0]) ** 3 for k in mine))
    t = torch.tensor([ms, fl, float(len(mine)), float((info != 0).sum())], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    if dist is not None:
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        times = [float(x[0]) for x in allt]
        tmax = max(times)
        line = {"workload": "mixed batch: %d matrices, n log-uniform [%d, %d], density U[0.01, 1], job %s, measure %s%s" % (
                    M, args.nmin, args.nmax, args.job, args.measure, ", shuffled entries" if args.shuffle else ""),
                "n_gpus": world, "value": M / (tmax * 1e-3), "unit": "decompositions/s", "ms": tmax,
                "tflops_equiv": sum(float(x[1]) for x in allt) / (tmax * 1e-3) / 1e12,
                "per_rank_ms": [round(x, 1) for x in times], "per_rank_matrices": [int(x[2]) for x in allt],
                "load_imbalance": tmax / (sum(times) / len(times)) - 1.0,
                "failed": int(sum(float(x[3]) for x in allt)),
                "rank0_stage_ms": {k: round(v[0], 1) for k, v in st.items() if v[0] > 0}, "rank0_generate_s": round(gen_s, 1),
                "largest_n": int(max(s[0] for s in shapes))}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
