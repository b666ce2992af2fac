"""oracle -- TEST INFRASTRUCTURE ONLY (ctypes front-end of oracle/gustavson.c and oracle/_ref).

May be imported only by tests/, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs.  The product package never imports it.

* :class:`Oracle`    -- host Gustavson restatement (gcc + OpenMP), see gustavson.c for the
  reference file:line each function follows.
* :class:`CuSparse`  -- cusparseSpGEMM with a selectable algorithm (DEFAULT / ALG1 / ALG2 / ALG3;
  oracle/cusparse_check.cu, own code); needs a GPU at call time.
* :class:`Reference` -- the UNMODIFIED reference kernels rebuilt for sm_100
  (oracle/_ref/libmhref.so, built by ``make -C oracle ref`` where /root/reference exists);
  needs a GPU at call time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libmhref.so")
CSK_SO = os.path.join(_HERE, "_build", "libcusparse_check.so")
REF_TREE = "/root/reference"

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the C restatement, and the reference where its tree is present."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all", "cusparse"])
    if ref and os.path.isdir(REF_TREE):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "-j8"])


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """Host Gustavson SpGEMM and the per-stage quantities of the reference pipeline."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_max_threads.restype = C.c_int
        L.orc_intprod.restype = C.c_int64
        L.orc_intprod.argtypes = [C.c_int, _i32p, _i32p, _i32p]
        L.orc_row_intprod.argtypes = [C.c_int, _i32p, _i32p, _i32p, _i64p]
        L.orc_mask_count.restype = C.c_int64
        L.orc_mask_count.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p]
        L.orc_mask_fill.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _u32p]
        L.orc_row_tileflop.argtypes = [C.c_int, _i32p, _i32p, _i32p, _i64p]
        L.orc_symbolic.restype = C.c_int
        L.orc_symbolic.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i64p]
        L.orc_symbolic_mask.restype = C.c_int
        L.orc_symbolic_mask.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _u32p, _i64p, _i64p]
        for name, fp in (("orc_numeric_f64", _f64p), ("orc_numeric_f32", _f32p)):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, fp, _i32p, _i32p, fp, _i64p, _i32p, fp]
        for name, fp in (("orc_compare_f64", _f64p), ("orc_compare_f32", _f32p)):
            f = getattr(L, name)
            f.restype = C.c_int64
            f.argtypes = [C.c_int, _i64p, _i32p, fp, _i64p, _i32p, fp, C.c_double, C.POINTER(C.c_int64)]
        L.orc_transpose.restype = None
        L.orc_transpose.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_void_p, C.c_int, _i32p, _i32p, C.c_void_p]
        self.L = L

    def transpose(self, A):
        """T = A^T as the reference's host transpose builds it (src/utils.cpp:20-46)."""
        from mh_spgemm_b200.csr import CSR
        Tp = np.zeros(A.N + 1, np.int32)
        Tc = np.zeros(max(A.nnz, 1), np.int32)
        Tv = np.zeros(max(A.nnz, 1), A.val.dtype)
        self.L.orc_transpose(A.M, A.N, A.ptr, A.col, A.val.ctypes.data, A.val.dtype.itemsize, Tp, Tc,
                             Tv.ctypes.data)
        return CSR(A.N, A.M, Tp, Tc[:A.nnz], Tv[:A.nnz])

    @property
    def threads(self) -> int:
        return int(self.L.orc_max_threads())

    def intprod(self, A, B) -> int:
        return int(self.L.orc_intprod(A.M, A.ptr, A.col, B.ptr))

    def row_intprod(self, A, B):
        out = np.zeros(A.M, np.int64)
        self.L.orc_row_intprod(A.M, A.ptr, A.col, B.ptr, out)
        return out

    def mask_matrix(self, B):
        tileptr = np.zeros(B.M + 1, np.int32)
        nt = int(self.L.orc_mask_count(B.M, B.N, B.ptr, B.col, tileptr))
        tilecol = np.zeros(max(nt, 1), np.int32)
        tilemask = np.zeros(max(nt, 1), np.uint32)
        self.L.orc_mask_fill(B.M, B.N, B.ptr, B.col, tileptr, tilecol, tilemask)
        return tileptr, tilecol[:nt], tilemask[:nt]

    def row_tileflop(self, A, tileptr):
        out = np.zeros(A.M, np.int64)
        self.L.orc_row_tileflop(A.M, A.ptr, A.col, _c(tileptr, np.int32), out)
        return out

    def symbolic(self, A, B):
        Cp = np.zeros(A.M + 1, np.int64)
        rc = self.L.orc_symbolic(A.M, A.N, B.N, A.ptr, A.col, B.ptr, B.col, Cp)
        assert rc == 0
        return Cp

    def symbolic_mask(self, A, B, mask=None):
        tileptr, tilecol, tilemask = mask if mask is not None else self.mask_matrix(B)
        Cp = np.zeros(A.M + 1, np.int64)
        ctiles = np.zeros(A.M, np.int64)
        tc = _c(tilecol, np.int32) if len(tilecol) else np.zeros(1, np.int32)
        tm = _c(tilemask, np.uint32) if len(tilemask) else np.zeros(1, np.uint32)
        rc = self.L.orc_symbolic_mask(A.M, B.N, A.ptr, A.col, _c(tileptr, np.int32), tc, tm, Cp, ctiles)
        assert rc == 0
        return Cp, ctiles

    def numeric(self, A, B, Cp):
        dt = A.val.dtype
        nnz = int(Cp[-1])
        Cc = np.zeros(max(nnz, 1), np.int32)
        Cv = np.zeros(max(nnz, 1), dt)
        f = self.L.orc_numeric_f64 if dt == np.float64 else self.L.orc_numeric_f32
        rc = f(A.M, A.N, B.N, A.ptr, A.col, A.val, B.ptr, B.col, _c(B.val, dt), _c(Cp, np.int64), Cc, Cv)
        assert rc == 0, "oracle numeric: row_ptr inconsistent with the structural product"
        return Cc[:nnz], Cv[:nnz]

    def spgemm(self, A, B):
        """Host Gustavson C = A*B -> (Cp int64[M+1], Cc int32[nnz] ascending per row, Cv)."""
        Cp = self.symbolic(A, B)
        Cc, Cv = self.numeric(A, B, Cp)
        return Cp, Cc, Cv

    def compare(self, M, c1, c2, rtol):
        """Number of mismatching entries between two (ptr, col, val) triples (0 == equal)."""
        (p1, k1, v1), (p2, k2, v2) = c1, c2
        dt = np.asarray(v1).dtype
        f = self.L.orc_compare_f64 if dt == np.float64 else self.L.orc_compare_f32
        first = C.c_int64(-1)

        def pad(a, d):
            a = _c(a, d)
            return a if a.size else np.zeros(1, d)

        bad = f(M, _c(p1, np.int64), pad(k1, np.int32), pad(v1, dt), _c(p2, np.int64), pad(k2, np.int32),
                pad(v2, dt), float(rtol), C.byref(first))
        return int(bad), int(first.value)


class CuSparse:
    """cusparseSpGEMM (generic API) as an independent oracle; ALG2 / ALG3 bound the work buffers
    (SURVEY.md 8f row 3).  Columns inside a row are sorted here before they are returned."""

    ALGS = {"default": 0, "alg1": 1, "alg2": 2, "alg3": 3}

    def __init__(self):
        if not os.path.exists(CSK_SO):
            subprocess.check_call(["make", "-s", "-C", _HERE, "cusparse"])
        L = C.CDLL(CSK_SO)
        L.csk_free.argtypes = [C.c_void_p]
        L.csk_spgemm.restype = C.c_int
        L.csk_spgemm.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, C.c_int,
                                 C.c_float, _i32p, C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_double)),
                                 C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
        self.L = L

    def spgemm(self, A, B, alg="default", chunk_fraction=0.2):
        """-> dict(ptr, col, val, nnz, ms) or raises MemoryError / RuntimeError."""
        Cp = np.zeros(A.M + 1, np.int32)
        cc = C.POINTER(C.c_int)()
        cv = C.POINTER(C.c_double)()
        nnz = C.c_longlong(0)
        ms = C.c_double(0)
        rc = self.L.csk_spgemm(A.M, A.N, B.N, A.ptr, A.col, _c(A.val, np.float64), B.ptr, B.col,
                               _c(B.val, np.float64), self.ALGS[alg], float(chunk_fraction), Cp, C.byref(cc),
                               C.byref(cv), C.byref(nnz), C.byref(ms))
        if rc == -2:
            raise MemoryError(f"cusparseSpGEMM {alg}: insufficient resources")
        if rc != 0:
            raise RuntimeError(f"cusparseSpGEMM {alg} failed")
        n = int(nnz.value)
        col = np.ctypeslib.as_array(cc, shape=(max(n, 1),))[:n].astype(np.int32, copy=True)
        val = np.ctypeslib.as_array(cv, shape=(max(n, 1),))[:n].astype(np.float64, copy=True)
        self.L.csk_free(C.cast(cc, C.c_void_p))
        self.L.csk_free(C.cast(cv, C.c_void_p))
        rows = np.repeat(np.arange(A.M, dtype=np.int64), np.diff(Cp))
        order = np.lexsort((col, rows))
        return dict(ptr=Cp, col=col[order], val=val[order], nnz=n, ms=float(ms.value))


class Reference:
    """The reference's own GPU kernels (oracle/_ref/libmhref.so). Needs a CUDA device."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                f"{REF_SO} missing: run `make -C oracle ref` where {REF_TREE} exists")
        L = C.CDLL(REF_SO)
        ip = C.POINTER(C.c_int)
        dp = C.POINTER(C.c_double)
        up = C.POINTER(C.c_uint)
        L.mhref_free.argtypes = [C.c_void_p]
        L.mhref_spgemm.restype = C.c_int
        L.mhref_spgemm.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p,
                                   C.c_int, C.c_int, C.c_int, _i32p, C.POINTER(ip), C.POINTER(dp),
                                   C.POINTER(C.c_int), dp, dp, _f64p,
                                   C.c_void_p, C.POINTER(ip), C.POINTER(up), dp]
        L.mhref_cusparse.restype = C.c_int
        L.mhref_cusparse.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p,
                                     C.c_int, C.c_int, _i32p, C.POINTER(ip), C.POINTER(dp),
                                     C.POINTER(C.c_int), dp, dp]
        L.mhref_transpose.restype = C.c_int
        L.mhref_transpose.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        self.L = L

    def transpose(self, A):
        """The reference's own host transpose (src/utils.cpp:20-46); needs no GPU."""
        from mh_spgemm_b200.csr import CSR
        Tp = np.zeros(A.N + 1, np.int32)
        Tc = np.zeros(max(A.nnz, 1), np.int32)
        Tv = np.zeros(max(A.nnz, 1), np.float64)
        rc = self.L.mhref_transpose(A.M, A.N, A.ptr, A.col, _c(A.val, np.float64), Tp, Tc, Tv)
        assert rc == 0
        return CSR(A.N, A.M, Tp, Tc[:A.nnz], Tv[:A.nnz])

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def _take(self, p, n, dt):
        if n == 0:
            out = np.zeros(0, dt)
        else:
            out = np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)
        self.L.mhref_free(C.cast(p, C.c_void_p))
        return out

    def spgemm(self, A, B, reps=1, warmup=0, e2e_reps=0, want_mask=False):
        """Run MH_spgemm. Returns dict(ptr, col, val, nnz, ms_device, ms_e2e, stage_ms[, mask])."""
        assert A.val.dtype == np.float64, "the reference is compiled for VALUE_TYPE double"
        Cp = np.zeros(A.M + 1, np.int32)
        cc = C.POINTER(C.c_int)()
        cv = C.POINTER(C.c_double)()
        nnz = C.c_int(0)
        msd, mse, msmin = C.c_double(0), C.c_double(0), C.c_double(0)
        stage = np.zeros(7, np.float64)
        tileptr = np.zeros(B.M + 1, np.int32)
        tc = C.POINTER(C.c_int)()
        tm = C.POINTER(C.c_uint)()
        rc = self.L.mhref_spgemm(
            A.M, A.N, B.N, A.ptr, A.col, A.val, B.ptr, B.col, B.val, reps, warmup, e2e_reps, Cp,
            C.byref(cc), C.byref(cv), C.byref(nnz), C.byref(msd), C.byref(mse), stage,
            tileptr.ctypes.data if want_mask else None,
            C.byref(tc) if want_mask else None, C.byref(tm) if want_mask else None, C.byref(msmin))
        if rc != 0:
            raise RuntimeError("reference MH_spgemm failed")
        n = int(nnz.value)
        out = dict(ptr=Cp, col=self._take(cc, n, np.int32), val=self._take(cv, n, np.float64), nnz=n,
                   ms_device=float(msd.value), ms_e2e=float(mse.value), stage_ms=stage,
                   ms_device_min=float(msmin.value))
        if want_mask:
            nt = int(tileptr[-1])
            out["mask"] = (tileptr, self._take(tc, nt, np.int32), self._take(tm, nt, np.uint32))
        return out

    def cusparse(self, A, B, reps=1, warmup=0):
        Cp = np.zeros(A.M + 1, np.int32)
        cc = C.POINTER(C.c_int)()
        cv = C.POINTER(C.c_double)()
        nnz = C.c_int(0)
        msd, msmin = C.c_double(0), C.c_double(0)
        rc = self.L.mhref_cusparse(A.M, A.N, B.N, A.ptr, A.col, A.val, B.ptr, B.col, B.val, reps, warmup,
                                   Cp, C.byref(cc), C.byref(cv), C.byref(nnz), C.byref(msd), C.byref(msmin))
        if rc != 0:
            raise RuntimeError("cuSPARSE SpGEMM failed")
        n = int(nnz.value)
        return dict(ptr=Cp, col=self._take(cc, n, np.int32), val=self._take(cv, n, np.float64), nnz=n,
                    ms_device=float(msd.value), ms_device_min=float(msmin.value))
