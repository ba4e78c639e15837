#!/usr/bin/env python
"""
bench.py -- eigendecompositions/sec of the paladin hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (COO -> dense scatter -> dgeev with right eigenvectors) over one
batch of B synthetic 512x512 matrices per GPU: a slice of the headline workload "12,800 random real
general 512x512 matrices, full decomposition" (BASELINE.json configs[2]). Matrices are independent,
so ranks own disjoint slices and there is no data-path collective ("scaling": "weak").

  value : decompositions/s with the COO triplets already resident in HBM (scatter + solve on device)
  e2e   : the same through the host-facing C-ABI: pinned host COO -> H2D -> scatter -> solve -> D2H of
          eigenvalues, right eigenvectors and info, every step
  --impl reference : the UNMODIFIED reference (oracle/_ref/exec-paladin, built by oracle/Makefile from
          /root/reference, linked against scipy's OpenBLAS/LAPACK 3.12) on all host cores through the
          in-repo MPI stand-in, serial LAPACK per rank (README.md:2), on a bounded sample of the same
          workload, timed by the reference's own `eig.` timer (src/paladin.hpp:340-342).

The oracle/ directory is only used for the cpu_baseline / reference arm, never for the measured GPU path.
"""
import argparse
import ctypes
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

METRIC = "eigendecompositions/sec (FP64 dgeev, n=512, eigenvalues + right eigenvectors)"
UNIT = "decompositions/s"
FLOPS_PER_N3_VEC = 26.33   # dgehrd 10/3 + dorghr 4/3 + dhseqr ~20 + dtrevc ~5/3
FLOPS_PER_N3_VAL = 10.0


def synth_coo(n, distinct, seed):
    """`distinct` dense N(0,1) matrices as COO in column-major file order."""
    ii = np.tile(np.arange(1, n + 1, dtype=np.int32), n)
    jj = np.repeat(np.arange(1, n + 1, dtype=np.int32), n)
    vals = [np.random.default_rng([seed, k]).standard_normal(n * n) for k in range(distinct)]
    return ii, jj, vals


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


def fp64_peak_tflops(device):
    """FP64 matrix-pipe peak on this device: a large cuBLAS DGEMM through torch (library GEMM used only as
    the roofline denominator, MEASURED_PEAKS.json has no FP64 entry)."""
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
    return best


# -----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference binary on the host cores
# -----------------------------------------------------------------------------------------------------
def run_reference_sample(n, per_rank, cores, right=True, seed=3):
    """Writes `distinct` MatrixMarket files, lists each rank's share, runs exec-paladin under the MPI
    stand-in with one rank per core and serial BLAS. Returns (decomp/s by the eig timer, decomp/s by total
    timer, description)."""
    from oracle import paladin_oracle as po
    ref = os.path.join(ROOT, "oracle", "_ref")
    exe, mpirun = os.path.join(ref, "exec-paladin"), os.path.join(ref, "shim_mpirun")
    if not (os.path.exists(exe) and os.path.exists(mpirun)):
        raise RuntimeError("oracle/_ref/exec-paladin missing (built by oracle/Makefile where /root/reference exists)")
    tmp = tempfile.mkdtemp(prefix="paladin_ref_")
    try:
        distinct = min(4, cores * per_rank)
        ii, jj, vals = synth_coo(n, distinct, seed)
        for k in range(distinct):
            with open(os.path.join(tmp, "m%d.matrix" % k), "wb") as f:
                f.write(po.mm_text(n, ii, jj, vals[k])[:-1])  # reference fixtures carry no trailing newline
        total = cores * per_rank
        with open(os.path.join(tmp, "listing"), "w") as f:
            f.write("\n".join("m%d.matrix" % (k % distinct) for k in range(total)))
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        cmd = [mpirun, "-np", str(cores), exe, "--listing=" + os.path.join(tmp, "listing"), "--rootdir=" + tmp]
        if right:
            cmd.append("--right")
        t0 = time.perf_counter()
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3600).stdout
        wall = time.perf_counter() - t0
        m = re.search(r"^max\t(\S+)\t(\S+)\t(\S+)", out, re.M)
        if not m:
            raise RuntimeError("could not parse reference timings:\n" + out[-2000:])
        t_read, t_eig, t_total = float(m.group(1)), float(m.group(2)), float(m.group(3))
        desc = ("%d matrices of %dx%d (%d distinct files) over %d ranks x %d each, OPENBLAS_NUM_THREADS=1, "
                "reference timers max-over-ranks: read %.3fs eig %.3fs total %.3fs, wall incl. writers %.1fs"
                % (total, n, n, distinct, cores, per_rank, t_read, t_eig, t_total, wall))
        return total / t_eig, total / t_total, desc
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    desc = ""
    for s in range(args.warmup + args.steps):
        v_eig, v_tot, desc = run_reference_sample(args.n, args.ref_per_rank, cores, right=args.job == "V")
        if s >= args.warmup:
            vals.append(v_eig)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * cores * args.ref_per_rank / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "12800x512x512 dense N(0,1), eigenvalues + right eigenvectors (bounded sample)",
                   "n": args.n, "job": args.job},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# -----------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    import paladin_b200 as pg
    from paladin_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = local
    assert pg.device_count() > dev, "no B200 for this rank (no CPU fallback)"

    n, B, job = args.n, args.batch, args.job
    distinct = min(args.distinct, B)
    ii1, jj1, vals = synth_coo(n, distinct, 3 + 1000 * rank)
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

    batch = pg.Batch(nvec, off, jobvr=job, device=dev)
    batch.upload(pi.array, pj.array, pv.array)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        batch.scatter()
        batch.solve()

    # e2e: the step's matrices go through the C-ABI as --e2e-chunks batches (default 8 x 74 matrices: four of them fill
    # the 296 CTA slots of the QR kernel) driven by one host thread each; every batch owns a stream, so the copies of one
    # batch run under the kernels of the others -- and the HBM-bound Hessenberg kernels of one batch under the
    # latency-bound QR kernels of another (measured: 2 batches 1283/s, 4: 1318, 8: 1388, 16: 1236). Over K steps each thread streams its
    # half K times (upload, scatter, solve, download per step, results of every step land in host memory); the threads
    # are joined once, at the end of the K steps, like the batches of the paladin-gpu executable.
    halves = []
    if B >= 2:
        nchunk = max(2, min(args.e2e_chunks, B))
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
    for _ in range(args.warmup):
        step_resident()
    stage_acc = {}
    launches = 0
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    batch.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        for k, (ms, ln) in batch.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + ms
            launches += ln
    dev_ms = batch.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.result()
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # sanity: results are real (spot-check a residual without the oracle)
    wr, wi, vr, info = batch.download(out_wr.array, out_wi.array, out_vr.array if job == "V" else None, out_info.array)
    assert not info.any(), "non-zero info in bench batch"
    if job == "V":
        A = vals[0].reshape((n, n), order="F")
        V = vr[:nn].reshape((n, n), order="F")
        k0 = int(np.argmax(wi[:n] == 0.0)) if np.any(wi[:n] == 0.0) else None
        if k0 is not None:
            r = np.linalg.norm(A @ V[:, k0] - wr[k0] * V[:, k0]) / (np.linalg.norm(A, 2) * n * 2.2e-16)
            assert r < 100.0, "bench residual check failed: %g" % r

    # ---- the same sub-batches with resident inputs (their COO stays in HBM from the uploads above): what the overlap of
    # kernels of different batches is worth without the copies (host clock between device synchronisations) ----
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

    overlapped = None
    if halves:
        steps_e2e(1)   # every sub-batch has its COO on the device
        steps_resident_chunks(1)
        barrier()
        t0 = time.perf_counter()
        steps_resident_chunks(args.steps)
        barrier()
        ov_ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ov_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ov_ms = float(t[0]) / args.steps
        overlapped = {"value": world * B / (ov_ms * 1e-3), "unit": UNIT, "ms_per_step": ov_ms, "batches_in_flight": len(halves),
                      "timing": "host clock between device synchronisations, max over ranks"}

    # ---- e2e: host buffers through the C-ABI ----
    steps_e2e(max(1, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    steps_e2e(args.steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0]) / args.steps
    e2e_value = world * B / (e2e_ms * 1e-3)
    h2d = world * B * nn * 16                                                   # whole job, like `value`
    d2h = world * (B * (2 * n * 8 + 4) + (B * nn * 8 if job == "V" else 0))

    if rank == 0:
        cyc = batch.phase_cycles()
        top = ("balance", "hessenberg", "form_q", "qr_total", "trevc", "bak_normalise")
        tot = float(sum(cyc[k] for k in top)) or 1.0
        phase_share = {k: round(v / tot, 4) for k, v in cyc.items() if v}
        peaks, peak_kind = measured_peaks()
        flops = (FLOPS_PER_N3_VEC if job == "V" else FLOPS_PER_N3_VAL) * n ** 3
        # dominant kernel: mid_qr_kernel (the QR iteration, one launch per step). Algorithmic work of that stage:
        # 20 n^3 flops with Schur vectors / 6.67 n^3 values-only (SURVEY.md 2.1 stage table)
        qr_ms = stage_acc.get("qr", 0.0) / args.steps
        solve_ms = sum(v for k, v in stage_acc.items() if k != "scatter") / args.steps
        qr_flops = (20.0 if job == "V" else 6.67) * n ** 3
        fp64_peak = fp64_peak_tflops(dev)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "qr_traffic.json")
        if os.path.exists(tpath) and n == 512 and job == "V":
            tj = json.load(open(tpath))  # ncu --set full capture of the same kernel, scaled to this launch's matrix count
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["matrices_per_launch"] * B
        if qr_ms > 0:
            achieved = B * qr_flops / (qr_ms * 1e-3) / 1e12
            kernel_name = "mid_qr_kernel (windowed multishift QR with early deflation, CTA per matrix)"
        else:  # n <= 64: everything runs in the fused warp-per-matrix kernel
            achieved = B * flops / (solve_ms * 1e-3) / 1e12
            kernel_name = "geev_warp_kernel (all stages)"
        roofline = {"bound": "tensor", "kernel": kernel_name,
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                    "traffic": traffic,
                    "peak_source": "FP64 cuBLAS DGEMM 6144^3 measured in this run (DMMA path; MEASURED_PEAKS.json has no FP64 "
                                   "entry; its hbm_gbs=%.1f (%s) is the scatter denominator)" % (peaks["hbm_gbs"], peak_kind),
                    "flops_per_unit": qr_flops if qr_ms > 0 else flops,
                    "kernel_ms_per_launch": qr_ms if qr_ms > 0 else solve_ms,
                    "whole_solve": {"flops_per_unit": flops, "achieved": B * flops / (solve_ms * 1e-3) / 1e12,
                                    "frac": B * flops / (solve_ms * 1e-3) / 1e12 / fp64_peak}}
        scat_ms = stage_acc.get("scatter", 0.0) / args.steps
        scat_bytes = B * (16 * nn + 8 * nn)
        scatter = {"bound": "hbm", "achieved": scat_bytes / (scat_ms * 1e-3) / 1e9 if scat_ms > 0 else None,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "bytes_per_unit": 24 * nn}
        if scatter["achieved"]:
            scatter["frac"] = scatter["achieved"] / scatter["peak"]
        # Hessenberg + explicit Q stage: one read + one write of the trailing matrix per column (8 n^3 bytes) plus the
        # backward accumulation of Q (1.07 + 0.36 GB per 512^2 matrix, confirmed by ncu: 422 GB per 296)
        hess_ms = stage_acc.get("hessenberg", 0.0) / args.steps
        hess_bytes = (8.0 + (8.0 / 3.0 if job == "V" else 0.0)) * n ** 3
        hessenberg = {"bound": "hbm", "achieved": B * hess_bytes / (hess_ms * 1e-3) / 1e9 if hess_ms > 0 else None,
                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "bytes_per_unit": hess_bytes}
        if hessenberg["achieved"]:
            hessenberg["frac"] = hessenberg["achieved"] / hessenberg["peak"]
        hessenberg["traffic"] = None
        if os.path.exists(tpath) and n == 512 and job == "V":
            pj_ = json.load(open(tpath)).get("prep_r01z")
            if pj_:  # ncu dram__bytes of mid_prep_kernel<16>, scaled to this launch's matrix count (L2 absorbs the rest)
                hessenberg["traffic"] = (pj_["dram_bytes_read"] + pj_["dram_bytes_write"]) / pj_["matrices_per_launch"] * B
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                v_eig, v_tot, desc = run_reference_sample(n, args.ref_per_rank, cores, right=job == "V")
                cpu = {"value": v_eig, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc,
                       "value_incl_read": v_tot}
            except Exception as e:  # the baseline is reported, never required for the GPU number
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "12800x512x512 dense N(0,1), eigenvalues + right eigenvectors; step = %d matrices per GPU "
                                   "(%d distinct, listing-style repeats)" % (B, distinct),
                       "n": n, "job": job, "batch_per_gpu": B, "l2": "inputs larger than L2 (%.0f MiB COO per step)" % (B * nn * 16 / 2 ** 20)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "value_overlapped": overlapped,
            "gpu_launches": launches,
            "roofline": roofline, "scatter_roofline": scatter, "hessenberg_roofline": hessenberg, "cpu_baseline": cpu,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_acc.items() if v > 0},
            "phase_cycle_share": phase_share,
            "wall_ms_per_step": wall_ms / args.steps,
        }
        print(json.dumps(line), flush=True)
    batch.destroy()
    for hb, _, _ in halves:
        hb.destroy()
    for p in (pi, pj, pv, out_wr, out_wi, out_vr, out_info):
        p.free()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--job", default="V", choices=["N", "V"])
    ap.add_argument("--batch", type=int, default=592, help="matrices per GPU per step (two waves of 2 CTAs x 148 SMs)")
    ap.add_argument("--distinct", type=int, default=37, help="distinct matrices per GPU (repeated listing-style)")
    ap.add_argument("--ref-per-rank", type=int, default=2, help="matrices per host core in the reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="e2e: the step's matrices go through the C-ABI as this many batches, one host thread each")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
