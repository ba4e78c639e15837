#!/usr/bin/env python
"""
Regenerates tests/golden/ from the reference. Run in the build container only
(needs /root/reference and oracle/_ref built by `make -C oracle`); the GPU box
never runs this -- it only reads the committed fixtures.

  ref_ctest/            verbatim copies of the reference's ctest DATA fixtures
                        (reference test/{canonical-2x2,identity-matrix,multiple-matrices}:
                        *.matrix, listings, *.exact) -- data, not source.
  ref_run/              outputs of the UNMODIFIED reference binary (oracle/_ref/test-paladin,
                        --left --right) on seeded random MatrixMarket files, plus the dense
                        matrix the reference's reader produced for each (ref_probe dense).
  ref_parser/           reader edge cases (comments, duplicates, shuffled order, exponent
                        spellings) with the reference reader's dense result.
  ref_partition.json    pod colouring computed by the reference constructor
                        (ref_probe partition, PALADIN_SHIM_FAKE_SIZE=P) for a synthetic listing,
                        every --measure, P in {1,2,3,4,8}.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import paladin_oracle as po  # noqa: E402

REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw).stdout.decode()


def copy_ctest():
    dst = os.path.join(HERE, "ref_ctest")
    shutil.rmtree(dst, ignore_errors=True)
    for d in ("canonical-2x2", "identity-matrix", "multiple-matrices"):
        shutil.copytree(os.path.join(REF, "test", d), os.path.join(dst, d))


def dense_dump(path):
    with tempfile.NamedTemporaryFile(suffix=".bin") as tmp:
        run([os.path.join(BIN, "ref_probe"), "dense", path, tmp.name])
        raw = open(tmp.name, "rb").read()
    n, nnz = np.frombuffer(raw[:8], dtype=np.int32)
    return int(n), int(nnz), np.frombuffer(raw[8:], dtype=np.float64).copy()


def ref_runs():
    dst = os.path.join(HERE, "ref_run")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    names = []
    cases = [(3, 1.0), (7, 1.0), (16, 1.0), (33, 1.0), (48, 1.0), (24, 0.3), (40, 0.1)]
    for k, (n, dens) in enumerate(cases):
        if dens >= 1.0:
            ii, jj, vv = po.random_dense_coo(n, [11, k])
        else:
            ii, jj, vv = po.random_sparse_coo(n, dens, [11, k])
        name = "rand-n%d-d%d.matrix" % (n, int(dens * 100))
        with open(os.path.join(dst, name), "wb") as f:
            f.write(po.mm_text(n, ii, jj, vv, comments=(" seeded fixture %d" % k,)))
        names.append(name)
    with open(os.path.join(dst, "listing"), "w") as f:
        f.write("\n".join(names))
    run([os.path.join(BIN, "test-paladin"), "--listing=" + os.path.join(dst, "listing"), "--rootdir=" + dst,
         "--left", "--right"])
    for name in names:
        for ext in ("spectrum", "leftvecs", "rightvecs"):
            os.rename(os.path.join(dst, name + "." + ext), os.path.join(dst, name + "." + ext + ".exact"))
        n, nnz, dense = dense_dump(os.path.join(dst, name))
        np.save(os.path.join(dst, name + ".dense.npy"), dense)


def parser_cases():
    dst = os.path.join(HERE, "ref_parser")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    cases = {
        "duplicates.matrix": "%%MatrixMarket matrix coordinate real general\n% dup (2,1) twice: last wins\n"
                             "3 3 5\n1 1 1.5\n2 1 -2\n3 3 4e0\n2 1 7.25\n1 3 -0\n",
        "shuffled.matrix": "%MatrixMarket matrix coordinate real general\n%\n%\n4 4 6\n4 4 1\n1 4 2.5e-3\n"
                           "3 2 -1E+2\n2 2 0.1\n1 1 3.0000000000000001\n4 1 1.7976931348623157e308",
        "spaces-in-sizeline.matrix": "%%MatrixMarket matrix coordinate real general\n2  2   3\n1 1 1\n"
                                     "2 2 2\n1 2 -3.5\n",
        "empty.matrix": "%%MatrixMarket matrix coordinate real general\n5 5 0\n",
        "fortran-exponent.matrix": "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 2 1.5D0\n"
                                   "2 1 +2.5e+01\n",
    }
    for name, text in cases.items():
        path = os.path.join(dst, name)
        with open(path, "w") as f:
            f.write(text)
        n, nnz, dense = dense_dump(path)
        np.save(path + ".dense.npy", dense)


def partitions():
    rng = np.random.default_rng(2024)
    M = 57
    n = np.concatenate([rng.integers(2, 3000, size=M - 12), np.full(12, 64)]).astype(int)
    dens = rng.uniform(0.01, 1.0, size=M)
    nnz = np.maximum(1, np.round(dens * n * n)).astype(int)
    nnz[-12:] = 4096  # ties, the common case
    out = {"n": n.tolist(), "nnz": nnz.tolist(), "cases": {}}
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        for k in range(M):
            name = "m%03d" % k
            with open(os.path.join(tmp, name), "w") as f:
                f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n[k], n[k], nnz[k]))
            names.append(name)
        with open(os.path.join(tmp, "listing"), "w") as f:
            f.write("\n".join(names))
        for kind in ("dim", "nnz", "dcb", "zds", "sps", "spc"):
            for P in (1, 2, 3, 4, 8):
                env = dict(os.environ, PALADIN_SHIM_FAKE_SIZE=str(P))
                txt = run([os.path.join(BIN, "ref_probe"), "partition", "--listing=" + os.path.join(tmp, "listing"),
                           "--rootdir=" + tmp, "--measure=" + kind], env=env)
                lines = txt.split("@@PARTITION")[1].strip().split("\n")[1:]
                order = [int(l.split()[1].rsplit("/m", 1)[1]) for l in lines]
                color = [int(l.split()[0]) for l in lines]
                out["cases"]["%s/%d" % (kind, P)] = {"order": order, "color": color}
    with open(os.path.join(HERE, "ref_partition.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    copy_ctest()
    ref_runs()
    parser_cases()
    partitions()
    print("golden fixtures regenerated under", HERE)
