"""CPU tests: the oracle is pinned against (i) the reference's own GPU outputs recorded in
tests/golden/ (written by make_golden.py through oracle/_ref on a B200) and (ii) scipy."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import mh_spgemm_b200  # noqa: F401
from mh_spgemm_b200 import generators as G
from mh_spgemm_b200.csr import CSR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP64_RTOL = 1e-12  # BASELINE.json north_star tolerance for fp64 values


def sha(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dt).tobytes()).hexdigest()


@pytest.mark.parametrize("name", list(cases.SMALL))
def test_oracle_matches_reference_full(orc, name):
    """Full CSR recorded from the reference kernels: row_ptr / col_idx bit-exact, values 1e-12."""
    A, B = cases.SMALL[name]()
    B = A if B is None else B
    ref = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert np.array_equal(Cp, ref["ptr"].astype(np.int64))
    assert np.array_equal(Cc, ref["col"])
    np.testing.assert_allclose(Cv, ref["val"], rtol=FP64_RTOL, atol=0)
    # family 1: the reference's B mask matrix (tiles canonicalised to ascending order)
    tp, tc, tm = orc.mask_matrix(B)
    assert np.array_equal(tp, ref["tileptr"])
    assert np.array_equal(tc, ref["tilecol"])
    assert np.array_equal(tm, ref["tilemask"])
    # mask-formulated symbolic == column-formulated symbolic
    Cp2, _ = orc.symbolic_mask(A, B, (tp, tc, tm))
    assert np.array_equal(Cp, Cp2)


@pytest.mark.parametrize("name", ["dense_rows", "rmat_s14", "fem_small"])
def test_oracle_matches_reference_checksums(orc, name):
    A, B = cases.LARGE[name]()
    B = A if B is None else B
    meta = json.load(open(os.path.join(GOLD, f"ref_{name}.json")))
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert int(Cp[-1]) == meta["nnz"]
    assert sha(Cp, np.int32) == meta["sha_ptr"]
    assert sha(Cc, np.int32) == meta["sha_col"]
    assert Cv.sum() == pytest.approx(meta["sum_val"], rel=1e-10)
    w = (np.arange(Cv.size, dtype=np.int64) % 97 + 1).astype(np.float64)
    assert (Cv * w).sum() == pytest.approx(meta["sum_weighted"], rel=1e-10)
    tp, tc, tm = orc.mask_matrix(B)
    assert sha(tp, np.int32) == meta["sha_tileptr"]
    assert sha(tc, np.int32) == meta["sha_tilecol"]
    assert sha(tm, np.uint32) == meta["sha_tilemask"]


def test_reference_fault_cases_are_recorded():
    """The reference faults on all-small-row inputs (Poisson); the fixture records that."""
    s = json.load(open(os.path.join(GOLD, "ref_summary.json")))
    assert s["poisson_32"].startswith("reference faulted")
    assert all(v == "ok" for k, v in s.items() if k != "poisson_32")


@pytest.mark.parametrize("make", [
    lambda: G.poisson2d(24), lambda: G.fem3d(3, 4, 5, 2, seed=3), lambda: G.rmat(11, 2000, 9000, seed=4),
    lambda: G.uniform_random(500, 400, 3000, seed=5), lambda: G.triangular_grid(20), lambda: G.road_grid(30),
])
def test_oracle_matches_scipy(orc, make):
    A = make()
    B = A if A.M == A.N else A.transpose()
    assert A.is_canonical()
    Cp, Cc, Cv = orc.spgemm(A, B)
    S = A.to_scipy() @ B.to_scipy()
    S.sort_indices()
    # scipy's product is structural too (no pruning of cancellations without eliminate_zeros)
    assert np.array_equal(S.indptr.astype(np.int64), Cp)
    assert np.array_equal(S.indices, Cc)
    np.testing.assert_allclose(Cv, S.data, rtol=FP64_RTOL, atol=0)
    assert orc.intprod(A, B) == int(np.diff(B.ptr)[A.col].sum())
    assert np.array_equal(orc.row_intprod(A, B).sum(), orc.intprod(A, B))


def test_structural_zeros_are_kept(orc):
    """Cancellation must not drop an entry (inc/numeric.cuh:237-241 never tests for zero)."""
    A = CSR(2, 2, [0, 2, 2], [0, 1], [1.0, -1.0])
    B = CSR(2, 2, [0, 1, 2], [0, 0], [1.0, 1.0])
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert Cp.tolist() == [0, 1, 1] and Cc.tolist() == [0] and Cv.tolist() == [0.0]


def test_poisson_sizes_match_survey(orc):
    """65,536 rows, 326,656 nnz, 1,629,192 products, 846,852 nnz(C)."""
    A = G.poisson2d(256)
    assert (A.M, A.nnz) == (65536, 326656)
    assert orc.intprod(A, A) == 1629192
    assert int(orc.symbolic(A, A)[-1]) == 846852


def test_empty_and_ragged(orc):
    E = CSR(4, 4, np.zeros(5, np.int32), [], [])
    Cp, Cc, Cv = orc.spgemm(E, E)
    assert Cp.tolist() == [0] * 5 and Cc.size == 0
    A = CSR(3, 3, [0, 0, 3, 3], [0, 1, 2], [1.0, 2.0, 3.0])  # empty first and last rows
    Cp, Cc, Cv = orc.spgemm(A, A)
    assert Cp.tolist() == [0, 0, 3, 3] and Cc.tolist() == [0, 1, 2] and Cv.tolist() == [2.0, 4.0, 6.0]


@pytest.mark.parametrize("name", list(cases.TRANSPOSE))
def test_oracle_transpose_matches_reference(orc, name):
    """The AAT mode's B operand: the oracle's transpose against vectors recorded from the
    reference's own matrix_transposition (src/utils.cpp:20-46; make_golden_transpose.py)."""
    A = cases.TRANSPOSE[name]()
    ref = np.load(os.path.join(GOLD, f"ref_transpose_{name}.npz"))
    T = orc.transpose(A)
    assert T.M == A.N and T.N == A.M
    assert np.array_equal(T.ptr, ref["ptr"]) and np.array_equal(T.col, ref["col"])
    assert np.array_equal(T.val, ref["val"])  # values are moved, not computed: bit-for-bit
    assert T.is_canonical()
    N = A.transpose()  # the numpy transpose of the host container agrees too
    assert np.array_equal(N.ptr, T.ptr) and np.array_equal(N.col, T.col) and np.array_equal(N.val, T.val)
    T32 = orc.transpose(A.astype(np.float32))
    assert T32.val.dtype == np.float32 and np.array_equal(T32.col, T.col)
