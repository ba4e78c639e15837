"""CPU-only: the host side of the GPU path (paladin_b200/host, C++) against the oracle and the reference's golden
files: MatrixMarket parse bit-exact, load measures and pod colouring bit-exact (same libstdc++ std::sort), output
formats byte-exact, N>1 sharding over world_size-2 gloo."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import paladin_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ph():
    from paladin_b200 import host
    if not os.path.exists(host.exe_path()):
        import __graft_entry__
        __graft_entry__.build()
    host.lib()
    return host


ALL_MATS = sorted(glob.glob(os.path.join(GOLD, "ref_*", "**", "*.matrix"), recursive=True))


@pytest.mark.parametrize("path", ALL_MATS, ids=lambda p: os.path.relpath(p, GOLD))
def test_parser_bit_exact(ph, path):
    n, nnz, ii, jj, vv = ph.parse_mm(path)
    n0, nnz0, i0, j0, v0 = po.read_mm(path)
    assert (n, nnz) == (n0, nnz0)
    assert ii.tobytes() == i0.tobytes() and jj.tobytes() == j0.tobytes() and vv.tobytes() == v0.tobytes()


def test_parser_rejects_what_the_reference_mangles(ph, tmp_path):
    bad = tmp_path / "long-index.matrix"
    bad.write_text("%%MatrixMarket matrix coordinate real general\n3 3 1\n1234567 1 2.0")
    with pytest.raises(ValueError):
        ph.parse_mm(str(bad))
    notsq = tmp_path / "notsquare.matrix"
    notsq.write_text("%%MatrixMarket matrix coordinate real general\n3 4 1\n1 1 2.0")
    with pytest.raises(ValueError):
        ph.parse_mm(str(notsq))


def test_measures_and_partition_match_reference(ph):
    ref = json.load(open(os.path.join(GOLD, "ref_partition.json")))
    n, nnz = ref["n"], ref["nnz"]
    for key, want in ref["cases"].items():
        kind, P = key.split("/")
        m = [ph.measure(a, b, kind) for a, b in zip(n, nnz)]
        assert m == [po.measure(a, b, kind) for a, b in zip(n, nnz)]
        order, colour = ph.partition(m, int(P))
        assert order.tolist() == want["order"], key
        assert colour.tolist() == want["color"], key


def test_partition_random_ties_match_oracle(ph):
    rng = np.random.default_rng(2)
    for M, P in [(1, 1), (7, 3), (100, 8), (1000, 4), (513, 2)]:
        meas = rng.integers(1, 6, size=M).astype(np.float64) * 4096  # many ties: the unstable-sort order matters
        o1, c1 = ph.partition(meas, P)
        o0, c0 = po.partition(meas, P)
        assert o1.tolist() == o0.tolist() and c1.tolist() == c0.tolist()


CTEST = sorted(glob.glob(os.path.join(GOLD, "ref_ctest", "*", "*.matrix")))
RUNS = sorted(glob.glob(os.path.join(GOLD, "ref_run", "*.matrix")))


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_batch_parser_concatenates_in_listing_order(ph, threads):
    # the executable's threaded batch parser against the single-file reader, file by file and bit for bit
    paths = (CTEST + RUNS) * 2
    n, off, ii, jj, vv = ph.parse_batch(paths, threads)
    assert off[0] == 0 and off[-1] == len(vv)
    for k, p in enumerate(paths):
        n1, nnz1, i1, j1, v1 = ph.parse_mm(p)
        assert n[k] == n1 and off[k + 1] - off[k] == nnz1
        sl = slice(off[k], off[k + 1])
        assert np.array_equal(ii[sl], i1) and np.array_equal(jj[sl], j1)
        assert v1.tobytes() == vv[sl].tobytes()


@pytest.mark.parametrize("path", CTEST + RUNS, ids=lambda p: os.path.relpath(p, GOLD))
def test_writers_byte_exact(ph, path):
    # eigen data from the oracle's dgeev, formatted by the product's writers, against the reference's own files
    n, nnz, ii, jj, vv = po.read_mm(path)
    A = po.dense_matrix(n, ii, jj, vv)
    wr, wi, VL, VR, info = po.dgeev(A, left=True, right=True)
    assert ph.format_output(wr, wi, None, 0) == open(path + ".spectrum.exact", "rb").read()
    assert ph.format_output(wr, wi, VR, 1) == open(path + ".rightvecs.exact", "rb").read()
    assert ph.format_output(wr, wi, VL, 2) == open(path + ".leftvecs.exact", "rb").read()


def test_executable_partition_dump(ph, tmp_path):
    d = os.path.join(GOLD, "ref_ctest", "multiple-matrices")
    out = tmp_path / "partition.txt"
    r = subprocess.run([ph.exe_path(), "--listing=" + os.path.join(d, "listing"), "--rootdir=" + d, "--gpus=4",
                        "--dump-partition=" + str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [ln.split(" ", 2) for ln in out.read_text().splitlines()]
    names = open(os.path.join(d, "listing")).read().split("\n")
    meas = []
    for nm in names:
        n, nnz, *_ = po.read_mm(os.path.join(d, nm))
        meas.append(po.measure(n, nnz, "nnz"))
    order, colour = po.partition(meas, 4)
    assert [int(r_[1]) for r_ in rows] == colour.tolist()
    assert [r_[2] for r_ in rows] == [d + "/" + names[k] for k in order]
    # unknown keys warn, never abort; --left is refused loudly
    r = subprocess.run([ph.exe_path(), "--listing=" + os.path.join(d, "listing"), "--rootdir=" + d, "--gpus=2", "--load-measure=dim",
                        "--dump-partition=" + str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "not a recognized option" in r.stdout
    # without a GPU the executable refuses to compute (no CPU fallback)
    import paladin_b200 as pg
    if pg.device_count() == 0:
        r = subprocess.run([ph.exe_path(), "--listing=" + os.path.join(d, "listing"), "--rootdir=" + d, "--left"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np
import torch, torch.distributed as dist
from paladin_b200 import host
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
n = rng.integers(32, 3073, size=301)
nnz = (rng.uniform(0.01, 1.0, size=301) * n * n).astype(np.int64)
meas = [host.measure(a, b, "nnz") for a, b in zip(n, nnz)]
mine = host.pod_of_rank(meas, world, rank)
got = [None] * world
dist.all_gather_object(got, mine.tolist())
if rank == 0:
    flat = sorted(x for g in got for x in g)
    assert flat == list(range(301)), "pods are not a partition of the listing"
    loads = [sum(meas[i] for i in g) for g in got]
    assert max(loads) / (sum(loads) / world) - 1.0 < 0.05, loads
    print("SHARDING-OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_sharding_over_gloo(ph, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr=127.0.0.1",
                        "--master-port=29517", str(script)], capture_output=True, text=True, timeout=300)
    assert "SHARDING-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_fast_number_parse_is_bit_identical_to_atof(ph, tmp_path):
    # every spelling the C library accepts must convert to the same bits as the reference's atof/atoi calls
    rng = np.random.default_rng(9)
    vals = []
    for k in range(3000):
        x = float(rng.standard_normal()) * 10.0 ** int(rng.integers(-320, 309))
        fmt = ["%.17g", "%.16e", "%.8g", "%g", "%.3f", "%.25e", "%+.17g"][k % 7]
        s = fmt % x
        if len(s) <= 31:
            vals.append(s)
    vals += ["0", "-0", "-0.0", "1e400", "-1e400", "4.9e-324", "2.4e-324", "1e-400", "nan", "-nan", "inf", "-inf", "Infinity", "NaN",
             "0x1p3", "0x1.8p1", "1.5D0", "1.5d3", ".5", "5.", "-.5e1", "+3", "+-1", "1e", "1e+", "abc", "", "  7.25", "\t2.5", "1,5",
             "9007199254740993", "0.1e-1x", "123456789012345678901234567890"]
    n = 80
    assert len(vals) <= n * n
    lines = ["%d %d %s" % (k % n + 1, k // n + 1, v) for k, v in enumerate(vals)]
    path = tmp_path / "numbers.matrix"
    path.write_text("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(vals)) + "\n".join(lines))
    n1, nnz1, i1, j1, v1 = ph.parse_mm(str(path))
    n0, nnz0, i0, j0, v0 = po.read_mm(str(path))
    assert (n1, nnz1) == (n0, nnz0) and i1.tobytes() == i0.tobytes() and j1.tobytes() == j0.tobytes()
    bad = [vals[k] for k in range(nnz0) if v1[k:k + 1].tobytes() != v0[k:k + 1].tobytes()]
    assert not bad, bad[:10]
