"""
Pins the CPU oracle (oracle/paladin_oracle.{py,cpp}) against
  * every *.exact file of the reference's own ctest suite (tests/golden/ref_ctest), byte for byte,
  * outputs of the unmodified reference binary on seeded matrices (tests/golden/ref_run),
  * the reference reader's dense result on parser edge cases (tests/golden/ref_parser),
  * the reference constructor's pod colouring (tests/golden/ref_partition.json).
CPU only.
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import paladin_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_files(matrix_path):
    n, nnz, ii, jj, vv = po.read_mm(matrix_path)
    A = po.dense_matrix(n, ii, jj, vv)
    wr, wi, VL, VR, info = po.dgeev(A, left=True, right=True)
    assert info == 0
    return (po.format_spectrum(wr, wi),
            po.format_vectors(po.unpack_parts(wr, wi, VL)),
            po.format_vectors(po.unpack_parts(wr, wi, VR)))


CTEST = sorted(glob.glob(os.path.join(GOLD, "ref_ctest", "*", "*.matrix")))
RUNS = sorted(glob.glob(os.path.join(GOLD, "ref_run", "*.matrix")))


def test_fixture_inventory():
    exact = glob.glob(os.path.join(GOLD, "ref_ctest", "*", "*.exact"))
    assert len(CTEST) == 10 and len(exact) == 30  # 20 distinct matrices x 3 outputs = 60 listed compares
    assert len(RUNS) == 7


@pytest.mark.parametrize("path", CTEST + RUNS, ids=lambda p: os.path.relpath(p, GOLD))
def test_oracle_reproduces_reference_files(path):
    spec, left, right = _oracle_files(path)
    assert spec == open(path + ".spectrum.exact", "rb").read()
    assert left == open(path + ".leftvecs.exact", "rb").read()
    assert right == open(path + ".rightvecs.exact", "rb").read()


@pytest.mark.parametrize("path", RUNS + sorted(glob.glob(os.path.join(GOLD, "ref_parser", "*.matrix"))),
                         ids=lambda p: os.path.relpath(p, GOLD))
def test_oracle_parse_and_scatter_bit_exact(path):
    n, nnz, ii, jj, vv = po.read_mm(path)
    dense = po.scatter_dense(n, ii, jj, vv)
    want = np.load(path + ".dense.npy")
    assert dense.shape == want.shape
    assert dense.tobytes() == want.tobytes()


def test_oracle_partition_matches_reference():
    ref = json.load(open(os.path.join(GOLD, "ref_partition.json")))
    n, nnz = ref["n"], ref["nnz"]
    for key, want in ref["cases"].items():
        kind, P = key.split("/")
        m = [po.measure(a, b, kind) for a, b in zip(n, nnz)]
        order, color = po.partition(m, int(P))
        assert order.tolist() == want["order"], key
        assert color.tolist() == want["color"], key


def test_listing_partition_of_ctest_suite():
    # the reference's own heterogeneous listing (multiple-matrices) through sort + LPT at 4 ranks:
    # 9 10x10 (nnz=10) first, then the 16 2x2 (nnz=4); greedy first-minimum colouring.
    d = os.path.join(GOLD, "ref_ctest", "multiple-matrices")
    names = open(os.path.join(d, "listing")).read().split("\n")
    meas = []
    for nm in names:
        n, nnz, *_ = po.read_mm(os.path.join(d, nm))
        meas.append(po.measure(n, nnz, "nnz"))
    order, color = po.partition(meas, 4)
    assert sorted(order.tolist()) == list(range(25))
    assert color.tolist()[:9] == [0, 1, 2, 3, 0, 1, 2, 3, 0]
    loads = np.zeros(4)
    for s, c in zip(order, color):
        loads[c] += meas[s]
    assert loads.max() - loads.min() <= 10


def test_match_and_backward_error_tools():
    rng = np.random.default_rng(5)
    A = rng.standard_normal((30, 30))
    wr, wi, VL, VR, info = po.dgeev(A, left=True, right=True)
    w = po.eigenvalues(wr, wi)
    assert po.match_eigenvalues(w[::-1], w) == 0.0
    assert po.backward_error(A, wr, wi, VR) < 10
    assert po.left_backward_error(A, wr, wi, VL) < 10
    assert po.match_eigenvalues(w + 1e-3, w) > 9e-4
