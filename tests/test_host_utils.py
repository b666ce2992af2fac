"""CPU tests of the host-side helpers that feed the path: the CSR container (inc/CSR.h:4-44
contract: 0-based, rows sorted ascending, duplicate-free) and the synthetic generators whose
shapes SURVEY.md 8(d) fixes."""
import numpy as np
import pytest

import mh_spgemm_b200  # noqa: F401
from mh_spgemm_b200 import generators as G
from mh_spgemm_b200.csr import CSR


def test_poisson_config1_shape():
    A = G.poisson2d(256)
    assert (A.M, A.N, A.nnz) == (65536, 65536, 326656) and A.is_canonical()
    assert np.all(A.val[A.col == np.repeat(np.arange(A.M), np.diff(A.ptr))] == 4.0)
    assert set(np.unique(A.val)) == {-1.0, 4.0}


def test_cant_like_config2_shape():
    A = G.fem3d(8, 8, 325, 3, seed=1)
    assert (A.M, A.nnz) == (62400, 4238388) and A.is_canonical()
    assert int(np.diff(A.ptr).max()) == 81
    ip = int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
    assert ip == 302542020  # intermediate products of C = A*A
    assert A.val.min() >= 0.5 and A.val.max() < 1.5  # all positive: no cancellation in the products
    # the three dof rows of a node share one column list (what the twin kernels exploit)
    r = 3 * 1234
    c0 = A.col[A.ptr[r]:A.ptr[r + 1]]
    assert np.array_equal(c0, A.col[A.ptr[r + 1]:A.ptr[r + 2]]) and np.array_equal(c0, A.col[A.ptr[r + 2]:A.ptr[r + 3]])


def test_rmat_is_seeded_and_canonical():
    A = G.rmat(14, 12000, 60000, seed=3)
    B = G.rmat(14, 12000, 60000, seed=3)
    assert A.is_canonical() and A.M == A.N == 12000
    assert np.array_equal(A.ptr, B.ptr) and np.array_equal(A.col, B.col) and np.array_equal(A.val, B.val)
    D = G.rmat(14, 12000, 60000, seed=4)
    assert D.nnz != A.nnz or not np.array_equal(D.col, A.col)  # another seed, another matrix


@pytest.mark.parametrize("make", [lambda: G.banded_random(3000, 7, 200), lambda: G.uniform_random(700, 900, 5000),
                                  lambda: G.triangular_grid(25), lambda: G.road_grid(40),
                                  lambda: G.with_dense_rows(G.banded_random(2000, 5, 50), 3, 1500)])
def test_generators_give_canonical_csr(make):
    A = make()
    assert A.is_canonical()
    assert A.ptr[0] == 0 and A.ptr[-1] == A.nnz and A.col.min() >= 0 and A.col.max() < A.N
    assert A.ptr.dtype == np.int32 and A.col.dtype == np.int32


def test_suite_names_cover_the_16_shapes():
    names = set(G.SUITE)
    for want in ("pdb1HYS", "pwtk", "webbase-1M", "cage12", "cant", "hood", "rma10", "scircuit", "shipsec1", "cop20k_A",
                 "mac_econ_fwd500", "offshore", "wb-edu", "cage15", "GAP-road", "delaunay_n24"):
        assert want in names
    A = G.suite("mac_econ_fwd500")
    assert A.M == 206500 and A.is_canonical()


def test_csr_rows_transpose_from_coo():
    rng = np.random.default_rng(0)
    A = CSR.from_coo(50, 70, rng.integers(0, 50, 400), rng.integers(0, 70, 400), rng=rng)
    assert A.is_canonical()  # duplicates merged, rows sorted
    S = A.to_scipy()
    T = A.transpose()
    assert T.is_canonical() and (T.M, T.N) == (70, 50)
    assert abs(T.to_scipy() - S.T).max() == 0
    blk = A.rows(10, 25)
    assert blk.M == 15 and blk.N == 70 and blk.ptr[0] == 0
    assert abs(blk.to_scipy() - S[10:25]).max() == 0
    assert A.astype(np.float32).val.dtype == np.float32
    # a row with unsorted or repeated columns is not canonical
    bad = CSR(1, 5, np.array([0, 2], np.int32), np.array([3, 1], np.int32), np.ones(2))
    dup = CSR(1, 5, np.array([0, 2], np.int32), np.array([2, 2], np.int32), np.ones(2))
    assert not bad.is_canonical() and not dup.is_canonical()


def test_perturbed_fem_breaks_the_twins():
    """fem3d_perturbed keeps the cant-like shape (diagonal intact, ~10 % of the rest dropped) but no
    two consecutive rows share a column list any more."""
    A, P = G.fem3d(4, 4, 12, 3, seed=3), G.fem3d_perturbed(4, 4, 12, 3, drop=0.10, seed=3)
    assert P.is_canonical() and (P.M, P.N) == (A.M, A.N)
    assert 0.86 * A.nnz < P.nnz < 0.94 * A.nnz
    rows = np.repeat(np.arange(P.M), np.diff(P.ptr))
    assert int((rows == P.col).sum()) == P.M  # the diagonal survives
    twins = lambda X: sum(np.array_equal(X.col[X.ptr[r - 1]:X.ptr[r]], X.col[X.ptr[r]:X.ptr[r + 1]])  # noqa: E731
                          for r in range(1, X.M))
    assert twins(A) == A.M // 3 * 2 and twins(P) < A.M // 50
    Q = G.fem3d_perturbed(4, 4, 12, 3, drop=0.10, seed=3)
    assert np.array_equal(P.col, Q.col) and np.array_equal(P.val, Q.val)  # seeded


def test_rmat_device_generator_is_seeded_and_canonical():
    """The torch-side R-MAT generator of the full-size configs[4] run (here on the CPU device): canonical
    CSR, same matrix for the same seed (every rank of a job draws its own copy), R-MAT skew present."""
    A = G.rmat_device(12, 1 << 12, 16 << 12, 0.45, 0.15, 0.15, seed=5, device="cpu")
    B = G.rmat_device(12, 1 << 12, 16 << 12, 0.45, 0.15, 0.15, seed=5, device="cpu")
    assert A.is_canonical() and (A.M, A.N) == (4096, 4096) and 0.8 * (16 << 12) < A.nnz <= (16 << 12)
    assert np.array_equal(A.ptr, B.ptr) and np.array_equal(A.col, B.col) and np.array_equal(A.val, B.val)
    assert A.val.min() >= 0.5 and A.val.max() < 1.5
    deg = np.diff(A.ptr)
    assert deg[:64].sum() > 8 * deg[-64:].sum()  # a = .45: the head rows are the heavy ones
    C = G.rmat_device(12, 1 << 12, 16 << 12, 0.45, 0.15, 0.15, seed=6, device="cpu")
    assert C.nnz != A.nnz or not np.array_equal(C.col, A.col)
