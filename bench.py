#!/usr/bin/env python
"""
bench.py -- eigendecompositions/sec of the paladin hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (COO -> dense scatter -> dgeev) over one batch of synthetic matrices per GPU.
Workloads (BASELINE.json configs):

  512v  (default, the configuration the metric and the 10x target are quoted on): a slice of "12,800 random real
        general 512x512 matrices, full decomposition"; step = 592 matrices per GPU, eigenvalues + right eigenvectors.
        Matrices are independent, ranks own disjoint slices, no data-path collective ("scaling": "weak").
  3072v README-scale 3072x3072, --right; step = 32 matrices per GPU (8 distinct): with 32 in flight a matrix gets a
        cluster of 4 CTAs instead of 16 and the leader-bound phases of the QR iteration overlap across matrices
        (8 per step: 4.5/s, 32: 8.5/s, 64: 8.9/s on one B200).
  64n   12,800 dense 64x64, eigenvalues only (warp-per-matrix path); step = 12,800 matrices per GPU.
  mixed M = 2048 matrices, n log-uniform in [32, 3072], density U[0.01, 1], --right; STRONG scaling: the fixed list is
        sharded over the ranks by the reference's own static balancer (measures -> std::sort -> greedy colouring,
        paladin_b200.host.pod_of_rank) and the line reports per-rank busy time and the reference's load-imbalance
        statistic (src/paladin.hpp:417-418).

The default line (512v) carries the other three as a `secondary` object (each with its own throughput, roofline
fraction, clocks and -- at N = 1 -- cpu_baseline). `--workload X` makes X the primary line instead.

  value : decompositions/s with the COO triplets already resident in HBM (scatter + solve on device)
  e2e   : the same through the host-facing C-ABI: pinned host COO -> H2D -> scatter -> solve -> D2H of
          eigenvalues, right eigenvectors and info, every step
  exe   : text in -> text out through the drop-in executable paladin-gpu (MatrixMarket files -> .spectrum/.rightvecs)
  --impl reference : the UNMODIFIED reference (oracle/_ref/exec-paladin, built by oracle/Makefile from
          /root/reference, linked against scipy's OpenBLAS/LAPACK 3.12) on all host cores through the
          in-repo MPI stand-in, serial LAPACK per rank (README.md:2), on a bounded sample of the same
          workload, timed by the reference's own `eig.` timer (src/paladin.hpp:340-342).

The oracle/ directory is only used for the cpu_baseline / reference arm, never for the measured GPU path.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "decompositions/s"
FLOPS_PER_N3_VEC = 26.33   # dgehrd 10/3 + dorghr 4/3 + dhseqr ~20 + dtrevc ~5/3
FLOPS_PER_N3_VAL = 10.0

# workload -> (n, job, matrices per GPU per step, distinct matrices, seed, description)
UNIFORM = {
    "512v": (512, "V", 592, 37, 3, "12800x512x512 dense N(0,1), eigenvalues + right eigenvectors"),
    "3072v": (3072, "V", 32, 8, 5, "README-scale 3072x3072 dense N(0,1), eigenvalues + right eigenvectors (streamed subset)"),
    "64n": (64, "N", 12800, 128, 2, "12800x64x64 dense N(0,1), eigenvalues only"),
}
MIXED_DESC = "mixed batch: %d matrices, n log-uniform [32, 3072], density U[0.01, 1], eigenvalues + right eigenvectors"


def metric_name(n, job):
    return "eigendecompositions/sec (FP64 dgeev, n=%d, %s)" % (n, "eigenvalues + right eigenvectors" if job == "V" else "eigenvalues only")


def synth_coo(n, distinct, seed):
    """`distinct` dense N(0,1) matrices as COO in column-major file order (default_rng([seed, k]))."""
    ii = np.tile(np.arange(1, n + 1, dtype=np.int32), n)
    jj = np.repeat(np.arange(1, n + 1, dtype=np.int32), n)
    vals = [np.random.default_rng([seed, k]).standard_normal(n * n) for k in range(distinct)]
    return ii, jj, vals


# ---- mixed batch: matrix k from default_rng([4, k, .]) ------------------------------------
def mixed_shape(k, nmin=32, nmax=3072):
    rng = np.random.default_rng([4, k, 0])
    n = int(round(np.exp(rng.uniform(np.log(nmin), np.log(nmax)))))
    dens = float(rng.uniform(0.01, 1.0))
    return n, max(1, int(round(dens * n * n)))


def mixed_coo(k, n, nnz):
    rng = np.random.default_rng([4, k, 1])
    nn = n * n
    if nnz >= nn:
        pos = np.arange(nn, dtype=np.int64)
    else:
        pos = np.sort(rng.choice(nn, size=nnz, replace=False)) if nn <= 4_000_000 else np.flatnonzero(rng.random(nn) < nnz / nn)
    vals = rng.standard_normal(len(pos))
    return (pos % n + 1).astype(np.int32), (pos // n + 1).astype(np.int32), vals


def mm_text(n, ii, jj, vv):
    """MatrixMarket 'coordinate real general' text of one matrix, "%d %d %.17g" entries, no trailing newline (the format of
    the reference's own fixtures)."""
    head = "%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(vv))
    return (head + "\n".join("%d %d %.17g" % t for t in zip(ii.tolist(), jj.tolist(), vv.tolist()))).encode()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.sm = []
        self.max_sm = None
        self.reasons = set()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(f[0]))
                self.max_sm = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=15)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


_FP64_PEAK = {}


def fp64_peak_tflops(device):
    """FP64 matrix-pipe peak on this device: a large cuBLAS DGEMM through torch (library GEMM used only as
    the roofline denominator, MEASURED_PEAKS.json has no FP64 entry)."""
    if device in _FP64_PEAK:
        return _FP64_PEAK[device]
    import torch
    n = 6144
    with torch.cuda.device(device):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del a, b
        torch.cuda.empty_cache()
    _FP64_PEAK[device] = best
    return best


def kernel_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per MATRIX of the named kernel, from the ncu --set full capture of
    THIS round's code (profiles/traffic.json, written from the .ncu-rep by tools/ncu_traffic.py); None when there is no
    capture of the current kernels."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    e = json.load(open(p)).get(key)
    if not e:
        return None
    return (e["dram_bytes_read"] + e["dram_bytes_write"]) / e["matrices_per_launch"]


# -----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference binary on the host cores
# -----------------------------------------------------------------------------------------------------
def run_reference_files(files, listing, cores, right=True):
    """files: {name: (n, ii, jj, vv)} written as MatrixMarket text; listing: names, one per decomposition. Runs
    exec-paladin under the MPI stand-in with min(cores, len(listing)) ranks and serial BLAS. Returns a dict of the
    reference's own max-over-ranks timers (src/paladin.hpp:365-420) plus the wall clock."""
    from oracle import paladin_oracle as po
    ref = os.path.join(ROOT, "oracle", "_ref")
    exe, mpirun = os.path.join(ref, "exec-paladin"), os.path.join(ref, "shim_mpirun")
    if not (os.path.exists(exe) and os.path.exists(mpirun)):
        raise RuntimeError("oracle/_ref/exec-paladin missing (built by oracle/Makefile where /root/reference exists)")
    tmp = tempfile.mkdtemp(prefix="paladin_ref_")
    try:
        for name, (n, ii, jj, vv) in files.items():
            with open(os.path.join(tmp, name), "wb") as f:
                f.write(po.mm_text(n, ii, jj, vv)[:-1])  # reference fixtures carry no trailing newline
        with open(os.path.join(tmp, "listing"), "w") as f:
            f.write("\n".join(listing))
        ranks = max(1, min(cores, len(listing)))
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        cmd = [mpirun, "-np", str(ranks), exe, "--listing=" + os.path.join(tmp, "listing"), "--rootdir=" + tmp, "--measure=dcb"]
        if right:
            cmd.append("--right")
        t0 = time.perf_counter()
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3600).stdout
        wall = time.perf_counter() - t0
        m = re.search(r"^max\t(\S+)\t(\S+)\t(\S+)", out, re.M)
        if not m:
            raise RuntimeError("could not parse reference timings:\n" + out[-2000:])
        return {"read": float(m.group(1)), "eig": float(m.group(2)), "total": float(m.group(3)), "wall": wall, "ranks": ranks}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_uniform(n, per_rank, cores, right=True, seed=3):
    """Bounded sample of a uniform workload: `per_rank` matrices per host core (4 distinct files, listing-style repeats).
    Returns (decomp/s by the eig timer, decomp/s by the total timer, description)."""
    total = cores * per_rank
    distinct = min(4, total)
    ii, jj, vals = synth_coo(n, distinct, seed)
    files = {"m%d.matrix" % k: (n, ii, jj, vals[k]) for k in range(distinct)}
    t = run_reference_files(files, ["m%d.matrix" % (k % distinct) for k in range(total)], cores, right)
    desc = ("%d matrices of %dx%d (%d distinct files) over %d ranks x %d each, OPENBLAS_NUM_THREADS=1, "
            "reference timers max-over-ranks: read %.3fs eig %.3fs total %.3fs, wall incl. writers %.1fs"
            % (total, n, n, distinct, t["ranks"], per_rank, t["read"], t["eig"], t["total"], t["wall"]))
    return total / t["eig"], total / t["total"], desc


def run_reference_mixed(M, stride, cores):
    """Bounded sample of the mixed batch: every `stride`-th matrix of the list through the reference with its own
    balancer (--measure=dcb) on all host cores; the rate is scaled by the sample's share of the whole list's 26.33 n^3 work
    (stated extrapolation)."""
    ks = list(range(0, M, stride))
    shapes = {k: mixed_shape(k) for k in ks}
    files = {"m%d.matrix" % k: (shapes[k][0],) + mixed_coo(k, *shapes[k]) for k in ks}
    t = run_reference_files(files, ["m%d.matrix" % k for k in ks], cores, True)
    work_sample = sum(float(shapes[k][0]) ** 3 for k in ks)
    work_all = sum(float(mixed_shape(k)[0]) ** 3 for k in range(M))
    est_total_s = t["eig"] * work_all / work_sample
    desc = ("%d of the %d matrices (every %dth) over %d ranks, --measure=dcb, OPENBLAS_NUM_THREADS=1, reference eig timer "
            "max-over-ranks %.2fs (total %.2fs); extrapolated to the whole list by its share of sum n^3 (%.4f)"
            % (len(ks), M, stride, t["ranks"], t["eig"], t["total"], work_sample / work_all))
    return M / est_total_s, desc


def per_core(v, cores):
    return v / cores if (v is not None and cores) else None


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = args.workload
    vals = []
    desc = ""
    if wl == "mixed":
        for s in range(args.warmup + args.steps):
            v, desc = run_reference_mixed(args.mixed_m, args.ref_mixed_stride, cores)
            if s >= args.warmup:
                vals.append(v)
        n, job, metric, workload = 0, "V", "eigendecompositions/sec (FP64 dgeev, mixed batch n=32...3072, eigenvalues + right eigenvectors)", MIXED_DESC % args.mixed_m
        units_per_step = args.mixed_m
    else:
        n, job, _, _, seed, workload = UNIFORM[wl]
        n = args.n or n
        job = args.job or job
        per_rank = args.ref_per_rank if args.ref_per_rank > 0 else {"512v": 8, "3072v": 1, "64n": 200}[wl]
        for s in range(args.warmup + args.steps):
            v_eig, v_tot, desc = run_reference_uniform(n, per_rank, cores, right=job == "V", seed=seed)
            if s >= args.warmup:
                vals.append(v_eig)
        metric = metric_name(n, job)
        units_per_step = cores * per_rank
    v = float(np.median(vals))
    line = {
        "impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * units_per_step / v, "higher_is_better": True,
        "scaling": "strong" if wl == "mixed" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": uniform_config(workload, n, job) if wl != "mixed" else {"workload": workload, "job": "V"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "per_core": per_core(v, cores), "kind": "reference", "sample": desc,
                         "statistic": "median over %d steps of the reference's max-over-ranks eig. timer" % len(vals)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def uniform_config(workload, n, job):
    # identical in both arms (the driver compares them); how many matrices a step holds is in `step_config`
    return {"workload": workload, "n": n, "job": job,
            "l2": "GPU arm: inputs larger than L2 between timed iterations (every step streams its whole COO + dense batch)"}


# -----------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.dev = self.local

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, *xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def gather(self, *xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            out = [self.torch.zeros_like(t) for _ in range(self.world)]
            self.dist.all_gather(out, t)
            return [[float(x) for x in o] for o in out]
        return [[float(x) for x in t]]


def run_uniform(cx, args, wl, n, job, B, distinct, seed, steps, warmup, full):
    """One uniform workload on this rank's GPU. full: also e2e / overlapped / phase profile (the primary line)."""
    import paladin_b200 as pg
    from paladin_b200 import api
    torch = cx.torch
    dev = cx.dev
    distinct = min(distinct, B)
    ii1, jj1, vals = synth_coo(n, distinct, seed + 1000 * cx.rank)
    nn = n * n
    nvec = np.full(B, n, dtype=np.int32)
    off = (np.arange(B + 1, dtype=np.int64) * nn)
    # pinned host COO for the whole batch (the "listing" repeats `distinct` files, as the reference's own
    # fixtures do: test/identity-matrix/10x10.listing10)
    pi = api.PinnedArray((B * nn,), np.int32)
    pj = api.PinnedArray((B * nn,), np.int32)
    pv = api.PinnedArray((B * nn,), np.float64)
    for k in range(B):
        pi.array[k * nn:(k + 1) * nn] = ii1
        pj.array[k * nn:(k + 1) * nn] = jj1
        pv.array[k * nn:(k + 1) * nn] = vals[k % distinct]
    out_wr = api.PinnedArray((B * n,), np.float64)
    out_wi = api.PinnedArray((B * n,), np.float64)
    out_vr = api.PinnedArray((B * nn if job == "V" else 1,), np.float64)
    out_info = api.PinnedArray((B,), np.int32)
    pinned = [pi, pj, pv, out_wr, out_wi, out_vr, out_info]

    batch = pg.Batch(nvec, off, jobvr=job, device=dev)
    batch.upload(pi.array, pj.array, pv.array)

    def step_resident():
        batch.scatter()
        batch.solve()

    # e2e: the step's matrices go through the C-ABI as --e2e-chunks batches (default 16 x 37 matrices when the rank has 16
    # host cores to itself, else 8 x 74: measured 1368-1377/s with 16, 1301-1322/s with 8, 1345/s with 4, 1002/s with 24 on
    # one B200 and 16 cores -- the finer the batches, the shorter the pipeline's fill and drain inside the K timed steps,
    # until the kernels get too small) driven by one host thread each; every batch owns a stream, so the copies of one
    # batch run under the kernels of the others. Over K steps each thread streams its share K times (upload, scatter,
    # solve, download per step, results of every step land in host memory); the threads are joined once, at the end of the
    # K steps, like the batches of the paladin-gpu executable.
    halves = []
    if full and B >= 2:
        # (large matrices: a batch of fewer than 16 gets clusters of 16 CTAs per matrix and leader-bound QR phases that do not
        # overlap; keep at least 16 per batch)
        want = args.e2e_chunks
        if want <= 0:
            want = 16 if (os.cpu_count() or 1) // max(1, cx.world) >= 16 else 8
        nchunk = max(2, min(want, B if n <= 1024 else B // 16))
        cuts = [B * q // nchunk for q in range(nchunk + 1)]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            hb = pg.Batch(nvec[lo:hi], off[lo:hi + 1] - off[lo], jobvr=job, device=dev)
            halves.append((hb, lo, hi))

    def run_half(hb, lo, hi):
        hb.upload(pi.array[lo * nn:hi * nn], pj.array[lo * nn:hi * nn], pv.array[lo * nn:hi * nn])
        hb.scatter()
        hb.solve()
        hb.download(out_wr.array[lo * n:hi * n], out_wi.array[lo * n:hi * n],
                    out_vr.array[lo * nn:hi * nn] if job == "V" else None, out_info.array[lo:hi])

    def steps_e2e(k):
        if not halves:
            for _ in range(k):
                batch.upload(pi.array, pj.array, pv.array)
                batch.scatter()
                batch.solve()
                batch.download(out_wr.array, out_wi.array, out_vr.array if job == "V" else None, out_info.array)
            return

        def stream_half(hb, lo, hi):
            for _ in range(k):
                run_half(hb, lo, hi)
        th = [threading.Thread(target=stream_half, args=h) for h in halves]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()

    # ---- value: inputs resident in HBM ----
    for _ in range(warmup):
        step_resident()
    stage_acc = {}
    launches = 0
    sampler = ClockSampler(cx.local)
    sampler.start()
    cx.barrier()
    batch.timer_start()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_resident()
        for k, (ms, ln) in batch.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + ms
            launches += ln
    dev_ms = batch.timer_stop()
    cx.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.result()
    dev_ms, wall_ms = cx.allmax(dev_ms, wall_ms)
    ms_per_step = dev_ms / steps
    value = cx.world * B / (ms_per_step * 1e-3)

    # sanity: results are real (spot-check residuals without the oracle)
    wr, wi, vr, info = batch.download(out_wr.array, out_wi.array, out_vr.array if job == "V" else None, out_info.array)
    assert not info.any(), "non-zero info in bench batch"
    for q in range(min(distinct, 3)):
        A = vals[q].reshape((n, n), order="F")
        lr, li = wr[q * n:(q + 1) * n], wi[q * n:(q + 1) * n]
        assert abs(lr.sum() - np.trace(A)) <= 1e-9 * n * np.sqrt(n), "bench trace check failed"
        if job == "V":
            V = vr[q * nn:(q + 1) * nn].reshape((n, n), order="F")
            k0 = int(np.argmax(li == 0.0)) if np.any(li == 0.0) else None
            if k0 is not None:
                r = np.linalg.norm(A @ V[:, k0] - lr[k0] * V[:, k0]) / (np.linalg.norm(A, 2) * n * 2.2e-16)
                assert r < 100.0, "bench residual check failed: %g" % r

    res = {"value": value, "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms / steps, "clocks": clocks, "launches": launches,
           "stage_ms": {k: v / steps for k, v in stage_acc.items() if v > 0}, "B": B, "distinct": distinct, "n": n, "job": job,
           "h2d": cx.world * B * nn * 16, "d2h": cx.world * (B * (2 * n * 8 + 4) + (B * nn * 8 if job == "V" else 0))}

    if full:
        # ---- the same sub-batches with resident inputs: what the overlap of kernels of different batches is worth without
        # the copies (host clock between device synchronisations) ----
        def steps_resident_chunks(k):
            def stream_half(hb):
                for _ in range(k):
                    hb.scatter()
                    hb.solve()
            th = [threading.Thread(target=stream_half, args=(h[0],)) for h in halves]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()

        if halves:
            steps_e2e(1)   # every sub-batch has its COO on the device
            steps_resident_chunks(1)
            cx.barrier()
            t0 = time.perf_counter()
            steps_resident_chunks(steps)
            cx.barrier()
            ov_ms = cx.allmax((time.perf_counter() - t0) * 1e3)[0] / steps
            res["overlapped"] = {"value": cx.world * B / (ov_ms * 1e-3), "unit": UNIT, "ms_per_step": ov_ms, "batches_in_flight": len(halves),
                                 "timing": "host clock between device synchronisations, max over ranks"}
        # ---- e2e: host buffers through the C-ABI ----
        steps_e2e(max(1, warmup // 2))
        cx.barrier()
        t0 = time.perf_counter()
        steps_e2e(steps)
        cx.barrier()
        e2e_ms = cx.allmax((time.perf_counter() - t0) * 1e3)[0] / steps
        res["e2e"] = {"value": cx.world * B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                      "d2h_bytes_per_step": res["d2h"], "ms_per_step": e2e_ms}
        res["phase_cycles"] = batch.phase_cycles()
    batch.destroy()
    for hb, _, _ in halves:
        hb.destroy()
    for p in pinned:
        p.free()
    return res


def rooflines(cx, res):
    """roofline objects of a uniform workload from its stage timers (rank 0)."""
    n, job, B = res["n"], res["job"], res["B"]
    nn = n * n
    peaks, peak_kind = measured_peaks()
    fp64_peak = fp64_peak_tflops(cx.dev)
    flops = (FLOPS_PER_N3_VEC if job == "V" else FLOPS_PER_N3_VAL) * n ** 3
    st = res["stage_ms"]
    qr_ms = st.get("qr", 0.0)
    solve_ms = sum(v for k, v in st.items() if k != "scatter")
    # dominant kernel: mid_qr_kernel (the QR iteration, one launch per size class and step). Algorithmic work of that
    # stage: 20 n^3 flops with Schur vectors / 6.67 n^3 values-only (SURVEY.md 2.1 stage table)
    qr_flops = (20.0 if job == "V" else 6.67) * n ** 3
    peak_source = ("FP64 cuBLAS DGEMM 6144^3 measured in this run (DMMA path; MEASURED_PEAKS.json has no FP64 entry; its "
                   "hbm_gbs=%.1f (%s) is the scatter denominator)" % (peaks["hbm_gbs"], peak_kind))
    if qr_ms > 0:
        achieved = B * qr_flops / (qr_ms * 1e-3) / 1e12
        kernel_name = "mid_qr_kernel (windowed multishift QR with early deflation, CTA per matrix)"
        tr = kernel_traffic("mid_qr_n%d_%s" % (n, job))
    else:  # n <= 64: everything runs in the fused warp-per-matrix kernel
        achieved = B * flops / (solve_ms * 1e-3) / 1e12
        kernel_name = "geev_warp_kernel (all stages, warp per matrix)"
        tr = kernel_traffic("warp_n%d_%s" % (n, job))
    roofline = {"bound": "tensor", "kernel": kernel_name, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak, "traffic": tr * B if tr else None, "peak_source": peak_source,
                "flops_per_unit": qr_flops if qr_ms > 0 else flops, "kernel_ms_per_launch": qr_ms if qr_ms > 0 else solve_ms,
                "whole_solve": {"flops_per_unit": flops, "achieved": B * flops / (solve_ms * 1e-3) / 1e12,
                                "frac": B * flops / (solve_ms * 1e-3) / 1e12 / fp64_peak}}
    out = {"roofline": roofline}
    scat_ms = st.get("scatter", 0.0)
    if scat_ms > 0:
        scat_bytes = B * (16 * nn + 8 * nn)
        a = scat_bytes / (scat_ms * 1e-3) / 1e9
        out["scatter_roofline"] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s", "bytes_per_unit": 24 * nn,
                                   "frac": a / peaks["hbm_gbs"]}
    hess_ms = st.get("hessenberg", 0.0)
    if hess_ms > 0:
        # the Hessenberg reduction and the explicit Q are dense contractions, bounded by the FP64 tensor pipe:
        # (10/3 + 4/3) n^3 flops with vectors, 10/3 n^3 values-only. traffic = ncu dram bytes of this round's kernel.
        hflops = (10.0 / 3.0 + (4.0 / 3.0 if job == "V" else 0.0)) * n ** 3
        a = B * hflops / (hess_ms * 1e-3) / 1e12
        tr = kernel_traffic("mid_prep_n%d_%s" % (n, job))
        out["hessenberg_roofline"] = {"bound": "tensor", "kernel": "mid_prep_kernel (balance + blocked Hessenberg + Q)", "achieved": a,
                                      "peak": fp64_peak, "unit": "TFLOP/s", "frac": a / fp64_peak, "flops_per_unit": hflops,
                                      "kernel_ms_per_launch": hess_ms, "traffic": tr * B if tr else None,
                                      "compulsory_bytes_per_unit": (16 + (8 if job == "V" else 0)) * nn}
    return out


def cpu_baseline_uniform(args, wl, n, job, seed):
    cores = os.cpu_count() or 1
    per_rank = args.ref_per_rank if args.ref_per_rank > 0 else {"512v": 8, "3072v": 1, "64n": 200}.get(wl, 2)
    try:
        v_eig, v_tot, desc = run_reference_uniform(n, per_rank, cores, right=job == "V", seed=seed)
        return {"value": v_eig, "unit": UNIT, "cores": cores, "per_core": per_core(v_eig, cores), "kind": "reference", "sample": desc,
                "value_incl_read": v_tot}
    except Exception as e:  # the baseline is reported, never required for the GPU number
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "failed: %s" % e}


def run_mixed(cx, args, measure, warmup):
    """Strong-scaled mixed batch: the fixed list of M matrices sharded by the reference balancer with `measure`."""
    import paladin_b200 as pg
    from paladin_b200 import host as ph
    M = args.mixed_m
    shapes = [mixed_shape(k) for k in range(M)]
    measures = [ph.measure(n, nnz, measure) for n, nnz in shapes]
    mine = [int(k) for k in ph.pod_of_rank(measures, cx.world, cx.rank)]
    t0 = time.time()
    coo = [mixed_coo(k, shapes[k][0], shapes[k][1]) for k in mine]
    nvec = np.array([shapes[k][0] for k in mine], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum([len(c[2]) for c in coo])]).astype(np.int64)
    ii = np.concatenate([c[0] for c in coo]) if coo else np.zeros(0, np.int32)
    jj = np.concatenate([c[1] for c in coo]) if coo else np.zeros(0, np.int32)
    vv = np.concatenate([c[2] for c in coo]) if coo else np.zeros(0)
    del coo
    gen_s = time.time() - t0
    b = pg.Batch(nvec, off, jobvr="V", device=cx.dev)
    b.upload(ii, jj, vv)
    for _ in range(warmup):
        b.scatter()
        b.solve()
    sampler = ClockSampler(cx.local)
    sampler.start()
    cx.barrier()
    b.timer_start()
    b.scatter()
    b.solve()
    ms = b.timer_stop()
    cx.barrier()
    clocks = sampler.result()
    wr, wi, vr, info = b.download()
    st = b.stage_ms()
    b.destroy()
    # size-independent check on this rank's share: trace of every matrix
    o1 = 0
    bad = int((info != 0).sum())
    fl = float(sum(FLOPS_PER_N3_VEC * float(shapes[k][0]) ** 3 for k in mine))
    rows = cx.gather(ms, fl, float(len(mine)), float(bad))
    del wr, wi, vr
    times = [r[0] for r in rows]
    tmax = max(times)
    return {"workload": MIXED_DESC % M, "measure": measure, "n_gpus": cx.world, "scaling": "strong", "value": M / (tmax * 1e-3), "unit": UNIT,
            "ms_per_step": tmax, "steps": 1, "warmup": warmup, "tflops_equiv": sum(r[1] for r in rows) / (tmax * 1e-3) / 1e12,
            "per_rank_busy_ms": [round(x, 1) for x in times], "per_rank_matrices": [int(r[2]) for r in rows],
            "load_imbalance": tmax / (sum(times) / len(times)) - 1.0, "failed": int(sum(r[3] for r in rows)),
            "rank0_stage_ms": {k: round(v[0], 1) for k, v in st.items() if v[0] > 0}, "rank0_generate_s": round(gen_s, 1),
            "clocks": clocks, "flops_total": sum(r[1] for r in rows)}


def secondary_entry_uniform(cx, args, wl):
    n, job, B, distinct, seed, workload = UNIFORM[wl]
    steps, warmup = {"3072v": (2, 3), "64n": (5, 3), "512v": (3, 3)}[wl]
    res = run_uniform(cx, args, wl, n, job, B, distinct, seed, steps, warmup, full=False)
    e = None
    if cx.rank == 0:
        rf = rooflines(cx, res)
        e = {"metric": metric_name(n, job), "value": res["value"], "unit": UNIT, "n_gpus": cx.world, "scaling": "weak",
             "ms_per_step": res["ms_per_step"], "steps": steps, "warmup": warmup,
             "config": uniform_config(workload, n, job), "step_config": "%d matrices per GPU (%d distinct, listing-style repeats)" % (res["B"], res["distinct"]),
             "clocks": res["clocks"], "gpu_launches": res["launches"], "stage_ms_per_step": res["stage_ms"]}
        e.update(rf)
        e["roofline_frac_whole_solve"] = rf["roofline"]["whole_solve"]["frac"]
        if cx.world == 1 and not args.no_cpu_baseline:
            e["cpu_baseline"] = cpu_baseline_uniform(args, wl, n, job, seed)
    return e


def secondary_entry_mixed(cx, args):
    runs = []
    measures = ["dcb"] if cx.world == 1 else ["dcb", "nnz"]
    for q, m in enumerate(measures):
        runs.append(run_mixed(cx, args, m, warmup=1 if q == 0 else 0))
    e = None
    if cx.rank == 0:
        e = dict(runs[0])
        fp64_peak = fp64_peak_tflops(cx.dev)
        e["roofline"] = {"bound": "tensor", "achieved": e["tflops_equiv"], "peak": fp64_peak * cx.world, "unit": "TFLOP/s",
                         "frac": e["tflops_equiv"] / (fp64_peak * cx.world), "flops": "26.33 n^3 per matrix, whole solve, all GPUs"}
        e["by_measure"] = {r["measure"]: {"value": r["value"], "load_imbalance": r["load_imbalance"], "per_rank_busy_ms": r["per_rank_busy_ms"],
                                          "per_rank_matrices": r["per_rank_matrices"]} for r in runs}
        if cx.world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            try:
                v, desc = run_reference_mixed(args.mixed_m, args.ref_mixed_stride, cores)
                e["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "per_core": per_core(v, cores), "kind": "reference", "sample": desc}
            except Exception as ex:
                e["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "failed: %s" % ex}
    return e


def exe_entry(cx, args):
    """text -> text through the drop-in executable (rank 0, N = 1 only): `--exe-matrices` MatrixMarket files of 512x512
    (listing-style repeats of 16 distinct files on local disk), --right, written .spectrum/.rightvecs included."""
    from paladin_b200 import host as ph
    n, count, distinct = 512, args.exe_matrices, 16
    # files on tmpfs when the box offers enough of it: the leg measures the software path (text -> GPU -> text), not the
    # write-back rate of the VM's disk; stated in the entry
    base, where = None, "local disk (/tmp)"
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > 16 * count * 2 ** 20:
            base, where = "/dev/shm", "tmpfs (/dev/shm)"
    except OSError:
        pass
    tmp = tempfile.mkdtemp(prefix="paladin_exe_", dir=base)
    try:
        ii, jj, vals = synth_coo(n, distinct, 3)
        names = []
        # every listed file is a distinct path (hard links to the distinct contents): outputs are written once per matrix
        for k in range(distinct):
            with open(os.path.join(tmp, "d%d.matrix" % k), "wb") as f:
                f.write(mm_text(n, ii, jj, vals[k]))
        for k in range(count):
            nm = "m%05d.matrix" % k
            os.link(os.path.join(tmp, "d%d.matrix" % (k % distinct)), os.path.join(tmp, nm))
            names.append(nm)
        with open(os.path.join(tmp, "listing"), "w") as f:
            f.write("\n".join(names))
        def run(extra):
            cmd = [ph.exe_path(), "--listing=" + os.path.join(tmp, "listing"), "--rootdir=" + tmp, "--right", "--gpus=1"] + extra
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                return None, wall, (r.stdout[-500:] + r.stderr[-500:])
            m = re.search(r"^max\t(\S+)\t(\S+)\t(\S+)", r.stdout, re.M)
            return (float(m.group(3)) if m else wall), wall, ""
        total, wall, err = run([])
        if total is None:
            return {"value": None, "error": err}
        nbytes_in = sum(os.path.getsize(os.path.join(tmp, "d%d.matrix" % (k % distinct))) for k in range(count))
        nbytes_out = sum(os.path.getsize(os.path.join(tmp, nm + ext)) for nm in names for ext in (".spectrum", ".rightvecs"))
        out = {"value": count / total, "unit": UNIT, "matrices": count, "seconds": total, "wall_incl_listing_scan": wall,
               "text_in_bytes": nbytes_in, "text_out_bytes": nbytes_out, "host_cores": os.cpu_count(), "files_on": where,
               "what": "paladin-gpu --right on %d MatrixMarket files of 512x512 -> .spectrum + .rightvecs (text in, text out); "
                       "entry lines parsed and eigenvector files formatted on the GPU" % count}
        # the same with parsing and formatting on the host cores (the round-1 path), on a sixth of the files
        sub = max(1, count // 6)
        with open(os.path.join(tmp, "listing_sub"), "w") as f:
            f.write("\n".join(names[:sub]))

        def run_sub():
            cmd = [ph.exe_path(), "--listing=" + os.path.join(tmp, "listing_sub"), "--rootdir=" + tmp, "--right", "--gpus=1", "--host-text"]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
            m = re.search(r"^max\t(\S+)\t(\S+)\t(\S+)", r.stdout, re.M)
            return float(m.group(3)) if (r.returncode == 0 and m) else None
        t2 = run_sub()
        if t2:
            out["host_text_value"] = sub / t2
            out["host_text_matrices"] = sub
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ours(args):
    cx = Ctx()
    import paladin_b200 as pg
    assert pg.device_count() > cx.dev, "no B200 for this rank (no CPU fallback)"
    wl = args.workload
    if wl == "mixed":
        e = secondary_entry_mixed(cx, args)
        if cx.rank == 0:
            line = {"metric": "eigendecompositions/sec (FP64 dgeev, mixed batch n=32...3072, eigenvalues + right eigenvectors)",
                    "value": e["value"], "unit": UNIT, "n_gpus": cx.world, "steps": 1, "warmup": 1, "ms_per_step": e["ms_per_step"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": e["workload"], "job": "V"}, "clocks": e["clocks"], "mixed": e,
                    "roofline": e["roofline"], "cpu_baseline": e.get("cpu_baseline"), "gpu_launches": None}
            print(json.dumps(line), flush=True)
        if cx.world > 1:
            cx.dist.destroy_process_group()
        return

    n, job, B, distinct, seed, workload = UNIFORM[wl]
    n = args.n or n
    job = args.job or job
    B = args.batch or B
    distinct = args.distinct or distinct
    res = run_uniform(cx, args, wl, n, job, B, distinct, seed, args.steps, args.warmup, full=True)

    secondary = {}
    if wl == "512v" and not args.no_secondary:
        for name in ("3072v", "64n"):
            try:
                secondary[name] = secondary_entry_uniform(cx, args, name)
            except Exception as ex:
                secondary[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        try:
            secondary["mixed"] = secondary_entry_mixed(cx, args)
        except Exception as ex:
            secondary["mixed"] = {"error": "%s: %s" % (type(ex).__name__, ex)}

    if cx.rank == 0:
        cyc = res["phase_cycles"]
        top = ("balance", "hessenberg", "hess_update", "form_q", "qr_total", "trevc", "bak_normalise")
        tot = float(sum(cyc[k] for k in top)) or 1.0
        phase_share = {k: round(v / tot, 4) for k, v in cyc.items() if v}
        rf = rooflines(cx, res)
        cpu = None
        if cx.world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_uniform(args, wl, n, job, seed)
        exe = None
        if cx.world == 1 and wl == "512v" and args.exe_matrices > 0:
            try:
                exe = exe_entry(cx, args)
            except Exception as ex:
                exe = {"value": None, "error": "%s: %s" % (type(ex).__name__, ex)}
        line = {
            "metric": metric_name(n, job), "value": res["value"], "unit": UNIT, "n_gpus": cx.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": uniform_config(workload, n, job),
            "step_config": {"batch_per_gpu": B, "distinct": res["distinct"], "coo_mib_per_step": B * n * n * 16 / 2 ** 20,
                            "note": "listing-style repeats of the distinct matrices, as in the reference's own fixtures"},
            "clocks": res["clocks"],
            "e2e": res.get("e2e"),
            "exe": exe,
            "value_overlapped": res.get("overlapped"),
            "gpu_launches": res["launches"],
            "cpu_baseline": cpu,
            "stage_ms_per_step": res["stage_ms"],
            "phase_cycle_share": phase_share,
            "wall_ms_per_step": res["wall_ms_per_step"],
            "host_cores": os.cpu_count(),
        }
        line.update(rf)
        if cpu and cpu.get("value"):
            line["gpu_equivalent_reference_cores"] = {"by_value": res["value"] / cx.world / cpu["per_core"],
                                                      "by_e2e": (res["e2e"]["value"] / cx.world / cpu["per_core"]) if res.get("e2e") else None,
                                                      "note": "one B200 in units of reference host cores (eig. timer), the portable figure"}
        if secondary:
            line["secondary"] = secondary
            mx = secondary.get("mixed")
            if isinstance(mx, dict) and "value" in mx:
                line["mixed"] = {"value": mx["value"], "unit": UNIT, "scaling": "strong", "load_imbalance": mx["load_imbalance"],
                                 "per_rank_busy_ms": mx["per_rank_busy_ms"], "measure": mx["measure"], "matrices": args.mixed_m}
        print(json.dumps(line), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="512v", choices=["512v", "3072v", "64n", "mixed"])
    ap.add_argument("--n", type=int, default=0, help="override the workload's matrix order")
    ap.add_argument("--job", default="", choices=["", "N", "V"], help="override the workload's job")
    ap.add_argument("--batch", type=int, default=0, help="matrices per GPU per step (512v: 592 = two waves of 2 CTAs x 148 SMs)")
    ap.add_argument("--distinct", type=int, default=0, help="distinct matrices per GPU (repeated listing-style)")
    ap.add_argument("--ref-per-rank", type=int, default=0, help="matrices per host core in the reference sample (0: 8 / 1 / 200 for 512v / 3072v / 64n)")
    ap.add_argument("--mixed-m", type=int, default=2048, help="matrices of the mixed batch (fixed total, strong scaling)")
    ap.add_argument("--ref-mixed-stride", type=int, default=64, help="the reference sample of the mixed batch takes every k-th matrix")
    ap.add_argument("--exe-matrices", type=int, default=7104, help="matrices of the text-in/text-out executable leg (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=-1, help="e2e: the step's matrices go through the C-ABI as this many batches, one host thread each (<= 0: 16 with 16 host cores per rank, else 8)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
