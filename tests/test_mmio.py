"""CPU tests of the Matrix Market reader against the semantics of the reference's
readMtxFile (inc/mmio_read.h:34-158)."""
import numpy as np
import pytest

import mh_spgemm_b200  # noqa: F401
from mh_spgemm_b200 import generators as G
from mh_spgemm_b200.mmio import read_mtx, write_mtx


def test_general_real_roundtrip(tmp_path):
    A = G.uniform_random(40, 30, 200, seed=1)
    p = tmp_path / "a.mtx"
    write_mtx(str(p), A)
    B, sym = read_mtx(str(p))
    assert not sym and (B.M, B.N) == (40, 30)
    assert np.array_equal(B.ptr, A.ptr) and np.array_equal(B.col, A.col) and np.array_equal(B.val, A.val)
    assert B.is_canonical()


def test_symmetric_expansion_and_one_based(tmp_path):
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% comment\n3 3 4\n1 1 2.0\n2 1 -1.0\n3 2 5.0\n3 3 1.5\n")
    A, sym = read_mtx(str(p))
    assert sym
    assert A.to_scipy().toarray().tolist() == [[2.0, -1.0, 0.0], [-1.0, 0.0, 5.0], [0.0, 5.0, 1.5]]
    assert A.ptr.tolist() == [0, 2, 4, 6] and A.col.tolist() == [0, 1, 0, 2, 1, 2]


def test_pattern_integer_complex(tmp_path):
    p = tmp_path / "p.mtx"
    p.write_text("%%MatrixMarket matrix coordinate pattern general\n2 2 2\n2 2\n1 2\n")
    A, _ = read_mtx(str(p))
    assert A.col.tolist() == [1, 1] and A.val.tolist() == [1.0, 1.0]
    p.write_text("%%MatrixMarket matrix coordinate integer general\n2 2 2\n1 1 7\n2 1 -3\n")
    A, _ = read_mtx(str(p))
    assert A.val.tolist() == [7.0, -3.0]
    p.write_text("%%MatrixMarket matrix coordinate complex hermitian\n2 2 2\n1 1 1.5 0.0\n2 1 2.5 -4.0\n")
    A, sym = read_mtx(str(p))
    assert not sym  # isSymmetric is mm_is_symmetric only; the expansion still happens
    assert A.to_scipy().toarray().tolist() == [[1.5, 2.5], [2.5, 0.0]]  # real part only, mirrored unchanged


def test_bad_banner(tmp_path):
    p = tmp_path / "b.mtx"
    p.write_text("hello\n1 1 1\n1 1 1\n")
    with pytest.raises(ValueError):
        read_mtx(str(p))
