"""Host-side CSR container mirroring the reference's ``class CSR`` (inc/CSR.h:4-44).

Field names follow the reference (``M, N, nnz, ptr, col, val``): 0-based int32 indices,
rows sorted ascending and duplicate-free (assumed by the reference at
inc/Form_mask_matrix_B.cuh:442-447 and enforced by its reader, inc/mmio_read.h:150).
Device twins live in the handle of :mod:`mh_spgemm_b200.api`, not here.
"""
from __future__ import annotations

import numpy as np


class CSR:
    __slots__ = ("M", "N", "ptr", "col", "val")

    def __init__(self, M, N, ptr, col, val):
        self.M = int(M)
        self.N = int(N)
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val)
        if self.val.dtype not in (np.float64, np.float32):
            self.val = self.val.astype(np.float64)
        assert self.ptr.shape == (self.M + 1,)
        assert self.col.shape == self.val.shape == (int(self.ptr[-1]),)

    @property
    def nnz(self) -> int:
        return int(self.ptr[-1])

    def astype(self, dt) -> "CSR":
        return CSR(self.M, self.N, self.ptr, self.col, self.val.astype(dt))

    def rows(self, r0: int, r1: int) -> "CSR":
        """Row block [r0, r1) as its own CSR (local row_ptr starting at 0)."""
        lo, hi = int(self.ptr[r0]), int(self.ptr[r1])
        return CSR(r1 - r0, self.N, self.ptr[r0:r1 + 1] - lo, self.col[lo:hi], self.val[lo:hi])

    def transpose(self) -> "CSR":
        """CSR transpose (the reference's matrix_transposition, src/utils.cpp:20-46)."""
        order = np.argsort(self.col, kind="stable")
        rows = np.repeat(np.arange(self.M, dtype=np.int32), np.diff(self.ptr))
        ptr = np.zeros(self.N + 1, np.int64)
        np.cumsum(np.bincount(self.col, minlength=self.N), out=ptr[1:])
        return CSR(self.N, self.M, ptr, rows[order], self.val[order])

    def is_canonical(self) -> bool:
        """Sorted, duplicate-free rows with in-range columns."""
        if self.nnz == 0:
            return True
        if self.col.min() < 0 or self.col.max() >= self.N:
            return False
        d = np.diff(self.col.astype(np.int64))
        starts = self.ptr[1:-1]
        starts = starts[(starts > 0) & (starts < self.nnz)]
        ok = d > 0
        ok[starts - 1] = True
        return bool(ok.all())

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val, self.col, self.ptr), shape=(self.M, self.N))

    @staticmethod
    def from_coo(M, N, rows, cols, vals=None, rng=None, dtype=np.float64) -> "CSR":
        """Canonical CSR from COO: duplicates merged (first kept), rows sorted."""
        key = np.asarray(rows, np.int64) * np.int64(N) + np.asarray(cols, np.int64)
        if vals is None:
            key = np.unique(key)
            v = (rng.random(key.size) + 0.5).astype(dtype) if rng is not None else np.ones(key.size, dtype)
        else:
            key, idx = np.unique(key, return_index=True)
            v = np.asarray(vals, dtype)[idx]
        r = key // N
        c = (key - r * N).astype(np.int32)
        ptr = np.zeros(M + 1, np.int64)
        np.cumsum(np.bincount(r, minlength=M), out=ptr[1:])
        return CSR(M, N, ptr, c, v)
