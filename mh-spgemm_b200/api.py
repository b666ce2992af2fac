"""Host-side mirror of the reference's operator interface over the C ABI (ctypes).

The reference's public surface for the hot path is ``CSR`` (inc/CSR.h) plus
``MH_spgemm(A, B, C, Timing, Tool)`` (src/main.cu:12).  The same names are kept here:

* :class:`Tool`       -- the workspace handle (Tool::allocate / release, src/Tool.cu).
* :func:`MH_spgemm`   -- C = A*B with host CSR in, host CSR out, through
  ``mhb_spgemm_host_*`` (CSR::H2D + MH_spgemm + CSR::D2H).
* ``Tool.symbolic`` / ``Tool.numeric`` -- the two phases on device-resident arrays
  (torch tensors are used only as device-memory owners).

There is no CPU fallback: if the CUDA library is missing or no device is present,
construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .csr import CSR

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmhb_spgemm.so")

SYM_BINS = ["EMPTY", "BM_G8", "BM_WARP", "BM_BLOCK", "H_G8", "H_WARP", "H_BLOCK_S", "H_BLOCK_L", "H_GLOBAL", "TINY", "TINY_S", "TINY_M", "H_G16"]
NUM_BINS = ["EMPTY", "WIN_G8", "WIN_WARP", "WIN_BLOCK_S", "WIN_BLOCK_L", "H_G8", "H_WARP_S", "H_WARP_L",
            "H_BLOCK_S", "H_BLOCK_L", "H_GLOBAL", "TINY", "TINY_S", "TINY_M", "H_WARP_XS", "H_WARP_M", "WIN_COMPACT",
            "H_BLOCK_M", "H_BLOCK_XS"]

# every symbol include/mhb_spgemm.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "mhb_version", "mhb_create", "mhb_destroy", "mhb_last_error", "mhb_set_stream", "mhb_set_option",
    "mhb_symbolic", "mhb_numeric_f64", "mhb_numeric_f32", "mhb_spgemm_f64", "mhb_spgemm_f32",
    "mhb_spgemm_into_f64", "mhb_spgemm_into_f32", "mhb_shard_spgemm_into_f64", "mhb_shard_spgemm_into_f32",
    "mhb_spgemm_into_begin_f64", "mhb_spgemm_into_begin_f32", "mhb_spgemm_into_end",
    "mhb_shard_spgemm_into_begin_f64", "mhb_shard_spgemm_into_begin_f32", "mhb_shard_spgemm_into_end",
    "mhb_shard_repost_size", "mhb_get_device_scalars",
    "mhb_device_free", "mhb_device_alloc", "mhb_memcpy_h2d", "mhb_memcpy_d2h", "mhb_spgemm_host_f64", "mhb_spgemm_host_f32", "mhb_host_alloc", "mhb_host_free",
    "mhb_form_mask_matrix_B", "mhb_get_row_info", "mhb_get_bins", "mhb_get_timing", "mhb_get_stats",
    "mhb_transpose_f64", "mhb_transpose_f32", "mhb_get_stream",
    "mhb_shard_create", "mhb_shard_destroy", "mhb_shard_last_error", "mhb_shard_set_A", "mhb_shard_export",
    "mhb_shard_import", "mhb_shard_own_B", "mhb_shard_image", "mhb_shard_exchange", "mhb_shard_publish",
    "mhb_shard_pull", "mhb_shard_barrier",
    "mhb_shard_symbolic", "mhb_shard_numeric_f64", "mhb_shard_numeric_f32", "mhb_shard_post_size",
    "mhb_shard_offsets", "mhb_nccl_unique_id", "mhb_shard_init_nccl", "mhb_shard_broadcast",
]
SHARD_BLOB_BYTES = 128
ERR_CUDA, ERR_ARG, ERR_OVERFLOW, ERR_NOMEM, ERR_CAPACITY = 1, 2, 3, 4, 5


class Timing(C.Structure):
    """Per-stage ms of the last call; field names follow the reference's Timing (inc/Timing.h:6-12)."""
    _fields_ = [(n, C.c_double) for n in ("mem_alloc", "Form_mask_matrix_B", "symbolic_binning", "Calculate_C_nnz",
                                          "Malloc_C_col_val", "numeric_binning", "Numeric", "total")]

    def getTotal(self) -> float:
        """The reference's convention (src/Timing.cpp:39-42): everything except the mask build."""
        return (self.Calculate_C_nnz + self.Malloc_C_col_val + self.Numeric + self.symbolic_binning
                + self.numeric_binning + self.mem_alloc)

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [("intprod", C.c_longlong), ("tileflop", C.c_longlong), ("ntiles_B", C.c_longlong),
                ("nnzC", C.c_longlong), ("sym_bin_size", C.c_int * 24), ("num_bin_size", C.c_int * 24),
                ("gpu_launches", C.c_int), ("hash_probes", C.c_longlong), ("sym_hash_probes", C.c_longlong),
                ("speculative_launches", C.c_int), ("speculative_misses", C.c_int), ("fused_calls", C.c_int)]

    def as_dict(self):
        return dict(intprod=self.intprod, tileflop=self.tileflop, ntiles_B=self.ntiles_B, nnzC=self.nnzC,
                    sym_bins=dict(zip(SYM_BINS, list(self.sym_bin_size))),
                    num_bins=dict(zip(NUM_BINS, list(self.num_bin_size))), gpu_launches=self.gpu_launches,
                    hash_probes=self.hash_probes, sym_hash_probes=self.sym_hash_probes,
                    speculative_launches=self.speculative_launches, speculative_misses=self.speculative_misses,
                    fused_calls=self.fused_calls)


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        try:  # a fresh checkout: compile in-tree with nvcc (seconds); never a CPU substitute
            from .build import build
            build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"{LIB_PATH} not built and nvcc build failed ({e}): run "
                               "`python mh-spgemm_b200/build.py` (there is no CPU fallback)") from e
    L = C.CDLL(LIB_PATH)
    vp, ip, ll = C.c_void_p, C.c_int, C.c_longlong
    L.mhb_version.restype = C.c_char_p
    L.mhb_last_error.restype = C.c_char_p
    L.mhb_last_error.argtypes = [vp]
    L.mhb_create.argtypes = [C.POINTER(vp), ip]
    L.mhb_destroy.argtypes = [vp]
    L.mhb_set_stream.argtypes = [vp, vp]
    L.mhb_set_option.argtypes = [vp, C.c_char_p, ll]
    L.mhb_symbolic.argtypes = [vp, ip, ip, ip, ip, vp, vp, ip, vp, vp, vp, C.POINTER(ll)]
    for n in ("mhb_numeric_f64", "mhb_numeric_f32"):
        getattr(L, n).argtypes = [vp, vp, vp, vp, vp]
    for n in ("mhb_spgemm_f64", "mhb_spgemm_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, ip, vp, vp, vp,
                                  C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(ll)]
    for n in ("mhb_spgemm_into_f64", "mhb_spgemm_into_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, ip, vp, vp, vp, vp, vp, vp, ll, C.POINTER(ll)]
    for n in ("mhb_shard_spgemm_into_f64", "mhb_shard_spgemm_into_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, vp, vp, vp, vp, ll, C.POINTER(ll)]
    for n in ("mhb_spgemm_into_begin_f64", "mhb_spgemm_into_begin_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, ip, vp, vp, vp, vp, vp, vp, ll]
    for n in ("mhb_shard_spgemm_into_begin_f64", "mhb_shard_spgemm_into_begin_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, vp, vp, vp, vp, ll]
    L.mhb_spgemm_into_end.argtypes = [vp, C.POINTER(ll)]
    L.mhb_shard_spgemm_into_end.argtypes = [vp, C.POINTER(ll)]
    L.mhb_device_free.argtypes = [vp]
    L.mhb_device_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.mhb_memcpy_h2d.argtypes = [vp, vp, C.c_size_t]
    L.mhb_memcpy_d2h.argtypes = [vp, vp, C.c_size_t]
    for n in ("mhb_spgemm_host_f64", "mhb_spgemm_host_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, ip, vp, vp, vp, vp, vp, vp,
                                  C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(ll)]
    for n in ("mhb_transpose_f64", "mhb_transpose_f32"):
        getattr(L, n).argtypes = [vp, ip, ip, ip, vp, vp, vp, vp, vp, vp]
    L.mhb_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.mhb_shard_create.argtypes = [C.POINTER(vp), vp, ip, ip, ip, ip, ip, C.POINTER(ll)]
    L.mhb_shard_destroy.argtypes = [vp]
    L.mhb_shard_last_error.argtypes = [vp]
    L.mhb_shard_last_error.restype = C.c_char_p
    L.mhb_shard_set_A.argtypes = [vp, ip, ip, vp, vp, vp]
    L.mhb_shard_export.argtypes = [vp, ip, vp]
    L.mhb_shard_import.argtypes = [vp, ip, vp]
    L.mhb_shard_own_B.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(ll)]
    L.mhb_shard_image.argtypes = [vp, C.POINTER(ip), C.POINTER(ip), C.POINTER(ll), C.POINTER(ll)]
    L.mhb_shard_exchange.argtypes = [vp]
    L.mhb_shard_publish.argtypes = [vp]
    L.mhb_shard_pull.argtypes = [vp]
    L.mhb_shard_barrier.argtypes = [vp]
    L.mhb_shard_symbolic.argtypes = [vp, ip, ip, vp, C.POINTER(ll)]
    L.mhb_shard_numeric_f64.argtypes = [vp, vp, vp, vp]
    L.mhb_shard_numeric_f32.argtypes = [vp, vp, vp, vp]
    L.mhb_shard_post_size.argtypes = [vp, ll]
    L.mhb_shard_repost_size.argtypes = [vp, ll]
    L.mhb_get_device_scalars.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mhb_shard_offsets.argtypes = [vp, C.POINTER(ll), C.POINTER(ll), C.POINTER(ll)]
    L.mhb_nccl_unique_id.argtypes = [vp]
    L.mhb_shard_init_nccl.argtypes = [vp, vp]
    L.mhb_shard_broadcast.argtypes = [vp, vp, C.c_size_t, ip]
    L.mhb_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.mhb_host_free.argtypes = [vp]
    L.mhb_form_mask_matrix_B.argtypes = [vp, ip, ip, ip, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                         C.POINTER(ll)]
    L.mhb_get_row_info.argtypes = [vp, C.POINTER(vp)]
    L.mhb_get_bins.argtypes = [vp, ip, C.POINTER(ip), C.POINTER(vp), C.POINTER(ip)]
    L.mhb_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.mhb_get_stats.argtypes = [vp, C.POINTER(Stats)]
    for n in ABI_SYMBOLS:
        if n not in ("mhb_version", "mhb_last_error", "mhb_shard_last_error"):
            getattr(L, n).restype = ip
    return L


class MhbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mhb status {code}: {msg}")
        self.code = code


def _hp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class PinnedCSR:
    """A CSR whose three arrays live in pinned host memory (what CSR::H2D should copy from)."""

    def __init__(self, lib, A: CSR):
        self._lib = lib
        self.M, self.N, self.nnz, self.dtype = A.M, A.N, A.nnz, A.val.dtype
        self._ptrs = []
        self.ptr = self._pin(A.ptr)
        self.col = self._pin(A.col)
        self.val = self._pin(A.val)

    def _pin(self, a: np.ndarray) -> np.ndarray:
        p = C.c_void_p()
        rc = self._lib.mhb_host_alloc(C.byref(p), max(a.nbytes, 1))
        if rc:
            raise MhbError(rc, "pinned allocation failed")
        self._ptrs.append(p)
        buf = (C.c_char * max(a.nbytes, 1)).from_address(p.value)
        out = np.frombuffer(buf, dtype=a.dtype, count=a.size)
        out[:] = a
        return out

    def close(self):
        for p in self._ptrs:
            self._lib.mhb_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tool:
    """Workspace + stream handle (the reference's ``Tool``, inc/Tool.h); one per device/thread."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.mhb_create(C.byref(self.h), device)
        if rc:
            raise MhbError(rc, "mhb_create failed: no usable CUDA device (there is no CPU fallback)")
        self.device = device
        self._keep = None

    # -- lifetime ---------------------------------------------------------------------
    def release(self):
        if self.h:
            self.L.mhb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise MhbError(rc, self.L.mhb_last_error(self.h).decode())

    def set_option(self, key: str, value: int):
        self._chk(self.L.mhb_set_option(self.h, key.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr: int | None):
        """None -> the handle's own non-blocking stream; an integer -> that cudaStream_t.  0 is the
        legacy default stream (what ``torch.cuda.current_stream().cuda_stream`` returns when no other
        stream was made current): it is passed on as cudaStreamLegacy, because a NULL pointer means
        "own stream" in the C ABI -- before this distinction, binding to torch's default stream
        silently left the handle on its own stream, unordered against the caller's work."""
        if cuda_stream_ptr is None:
            p = 0
        elif int(cuda_stream_ptr) == 0:
            p = 1  # cudaStreamLegacy
        else:
            p = int(cuda_stream_ptr)
        self._chk(self.L.mhb_set_stream(self.h, C.c_void_p(p)))

    def pin(self, A: CSR) -> PinnedCSR:
        return PinnedCSR(self.L, A)

    @property
    def timing(self) -> Timing:
        t = Timing()
        self._chk(self.L.mhb_get_timing(self.h, C.byref(t)))
        return t

    @property
    def stats(self) -> dict:
        s = Stats()
        self._chk(self.L.mhb_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    # -- host buffers in, host buffers out (CSR::H2D + MH_spgemm + CSR::D2H) --------------
    def spgemm_host(self, A, B=None, copy: bool = True) -> CSR:
        """C = A*B.  A, B: CSR or PinnedCSR (B=None or B is A -> C = A*A, uploaded once)."""
        B = A if B is None else B
        dt = np.dtype(A.val.dtype)
        if A.N != B.M:
            raise ValueError(f"inner dimensions differ: A is {A.M}x{A.N}, B is {B.M}x{B.N}")
        if np.dtype(B.val.dtype) != dt or dt not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError(f"A and B must share one value type (float64 or float32), got {dt} and {B.val.dtype}")
        f = self.L.mhb_spgemm_host_f64 if dt == np.float64 else self.L.mhb_spgemm_host_f32
        cp, cc, cv, nnz = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        self._chk(f(self.h, A.M, A.N, B.N, _hp(A.ptr), _hp(A.col), _hp(A.val), _hp(B.ptr), _hp(B.col),
                    _hp(B.val), C.byref(cp), C.byref(cc), C.byref(cv), C.byref(nnz)))
        n = int(nnz.value)

        def view(p, count, dtype):
            if count == 0:
                return np.zeros(0, dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p.value)
            a = np.frombuffer(buf, dtype=dtype, count=count)
            return a.copy() if copy else a

        return CSR(A.M, B.N, view(cp, A.M + 1, np.int32), view(cc, n, np.int32), view(cv, n, dt))

    # -- device-resident phases (torch tensors own the memory) ----------------------------
    def spgemm_sliced(self, A: CSR, B: CSR | None = None, cap: int = 2**31 - 1):
        """C = A*B for products whose nnz(C) may exceed the int32 contract: A's rows are cut
        into slices with at most `cap` intermediate products each (nnz(C slice) <= cap), every
        slice is one SpGEMM with its own local int32 row_ptr, and the caller gets the slices
        plus their int64 offsets.  Returns [(r0, r1, CSR slice)], offsets (int64)."""
        from .distributed import row_work, slice_rows_fast
        B = A if B is None else B
        work = row_work(A, B)
        out, offs, run = [], [], 0
        for r0, r1 in slice_rows_fast(work, 0, A.M, cap):
            C = self.spgemm_host(A.rows(r0, r1), B)
            out.append((r0, r1, C))
            offs.append(run)
            run += C.nnz
        return out, np.array(offs + [run], np.int64)

    def symbolic(self, M, K, N, dA_ptr, dA_col, dB_ptr, dB_col):
        """-> (dC_ptr int32[M+1], nnzC).  Inputs: int32 device arrays (DeviceArray or CUDA
        torch tensors -- anything with data_ptr()/numel())."""
        dC_ptr = _alloc_like(dA_ptr, M + 1, np.int32)
        nnz = C.c_longlong()
        self._keep = (dA_ptr, dA_col, dB_ptr, dB_col, dC_ptr)
        self._chk(self.L.mhb_symbolic(self.h, M, K, N, dA_col.numel(), dA_ptr.data_ptr(), dA_col.data_ptr(),
                                      dB_col.numel(), dB_ptr.data_ptr(), dB_col.data_ptr(), dC_ptr.data_ptr(),
                                      C.byref(nnz)))
        return dC_ptr, int(nnz.value)

    def numeric(self, dA_val, dB_val, nnzC):
        """-> (dC_col int32, dC_val) device arrays (first nnzC entries valid), using the last
        symbolic pattern."""
        dC_col = _alloc_like(dA_val, max(nnzC, 1), np.int32)
        dC_val = _alloc_like(dA_val, max(nnzC, 1), None)
        self.numeric_into(dA_val, dB_val, dC_col, dC_val)
        return dC_col, dC_val

    def numeric_into(self, dA_val, dB_val, dC_col, dC_val):
        f = self.L.mhb_numeric_f64 if _itemsize(dA_val) == 8 else self.L.mhb_numeric_f32
        self._chk(f(self.h, dA_val.data_ptr(), dB_val.data_ptr(), dC_col.data_ptr(), dC_val.data_ptr()))

    def spgemm_into(self, M, K, N, dA_ptr, dA_col, dA_val, dB_ptr, dB_col, dB_val, dC_ptr, dC_col, dC_val) -> int:
        """MH_spgemm into caller-owned C arrays (mhb_spgemm_into_*): one host synchronisation per
        call in steady state.  Returns nnz(C); raises MhbError with code ERR_CAPACITY (row_ptr valid,
        .nnzC on the exception) when dC_col / dC_val are too small."""
        f = self.L.mhb_spgemm_into_f64 if _itemsize(dA_val) == 8 else self.L.mhb_spgemm_into_f32
        nnz = C.c_longlong()
        self._keep = (dA_ptr, dA_col, dB_ptr, dB_col, dC_ptr)
        cap = min(dC_col.numel(), dC_val.numel()) if dC_col is not None else 0
        rc = f(self.h, M, K, N, dA_col.numel(), dA_ptr.data_ptr(), dA_col.data_ptr(), dA_val.data_ptr(),
               dB_col.numel(), dB_ptr.data_ptr(), dB_col.data_ptr(), dB_val.data_ptr(), dC_ptr.data_ptr(),
               dC_col.data_ptr() if cap else None, dC_val.data_ptr() if cap else None, cap, C.byref(nnz))
        if rc == ERR_CAPACITY:
            e = MhbError(rc, self.L.mhb_last_error(self.h).decode())
            e.nnzC = int(nnz.value)
            raise e
        self._chk(rc)
        return int(nnz.value)

    def spgemm_into_begin(self, M, K, N, dA_ptr, dA_col, dA_val, dB_ptr, dB_col, dB_val, dC_ptr, dC_col, dC_val):
        """First half of spgemm_into: queues the SpGEMM on the handle's stream and returns without
        waiting (steady state).  Follow with spgemm_into_end()."""
        f = self.L.mhb_spgemm_into_begin_f64 if _itemsize(dA_val) == 8 else self.L.mhb_spgemm_into_begin_f32
        self._keep = (dA_ptr, dA_col, dA_val, dB_ptr, dB_col, dB_val, dC_ptr, dC_col, dC_val)
        cap = min(dC_col.numel(), dC_val.numel()) if dC_col is not None else 0
        self._chk(f(self.h, M, K, N, dA_col.numel(), dA_ptr.data_ptr(), dA_col.data_ptr(), dA_val.data_ptr(),
                    dB_col.numel(), dB_ptr.data_ptr(), dB_col.data_ptr(), dB_val.data_ptr(), dC_ptr.data_ptr(),
                    dC_col.data_ptr() if cap else None, dC_val.data_ptr() if cap else None, cap))

    def spgemm_into_end(self) -> int:
        """Second half: synchronises, verifies; returns nnz(C) or raises (ERR_CAPACITY carries .nnzC)."""
        nnz = C.c_longlong()
        rc = self.L.mhb_spgemm_into_end(self.h, C.byref(nnz))
        if rc == ERR_CAPACITY:
            e = MhbError(rc, self.L.mhb_last_error(self.h).decode())
            e.nnzC = int(nnz.value)
            raise e
        self._chk(rc)
        return int(nnz.value)

    def transpose_device(self, M, N, dA_ptr, dA_col, dA_val):
        """T = A^T on the device (mhb_transpose_*): -> (dT_ptr[N+1], dT_col[nnz], dT_val[nnz])."""
        nnz = dA_col.numel()
        dT_ptr = _alloc_like(dA_ptr, N + 1, np.int32)
        dT_col = _alloc_like(dA_ptr, max(nnz, 1), np.int32)
        dT_val = _alloc_like(dA_val, max(nnz, 1), None)
        f = self.L.mhb_transpose_f64 if _itemsize(dA_val) == 8 else self.L.mhb_transpose_f32
        self._chk(f(self.h, M, N, nnz, dA_ptr.data_ptr(), dA_col.data_ptr(), dA_val.data_ptr(),
                    dT_ptr.data_ptr(), dT_col.data_ptr(), dT_val.data_ptr()))
        return dT_ptr, dT_col, dT_val

    def transpose(self, A: CSR) -> CSR:
        """Host CSR in, host CSR out through the device transpose (matrix_transposition,
        src/utils.cpp:20-46, as a library call)."""
        dp, dc, dv = DeviceArray(A.ptr), DeviceArray(A.col), DeviceArray(A.val)
        tp, tc, tv = self.transpose_device(A.M, A.N, dp, dc, dv)
        T = CSR(A.N, A.M, tp.numpy(), tc.numpy()[:A.nnz], tv.numpy()[:A.nnz])
        for d in (dp, dc, dv, tp, tc, tv):
            d.free()
        return T

    def mask_matrix_B(self, K, N, dB_ptr, dB_col):
        """Family 1 alone -> (tileptr[K+1], tilecol[nt], tilemask[nt]) as numpy arrays."""
        tp, tc, tm, nt = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        self._chk(self.L.mhb_form_mask_matrix_B(self.h, K, N, dB_col.numel(), dB_ptr.data_ptr(), dB_col.data_ptr(),
                                                C.byref(tp), C.byref(tc), C.byref(tm), C.byref(nt)))
        n = int(nt.value)
        return (_d2h(tp.value, K + 1, np.int32), _d2h(tc.value, n, np.int32), _d2h(tm.value, n, np.uint32))

    def row_info(self, M):
        p = C.c_void_p()
        self._chk(self.L.mhb_get_row_info(self.h, C.byref(p)))
        return _d2h(p.value, 4 * M, np.int32).reshape(M, 4)

    def bins(self, which: int, M: int):
        nb, p = C.c_int(), C.c_void_p()
        off = (C.c_int * 25)()
        self._chk(self.L.mhb_get_bins(self.h, which, C.byref(nb), C.byref(p), off))
        return int(nb.value), _d2h(p.value, M, np.int32), np.array(list(off), np.int32)


_lib_cache = None


def _lib():
    global _lib_cache
    if _lib_cache is None:
        _lib_cache = load_library()
    return _lib_cache


def _d2h(dev_ptr: int, count: int, dtype) -> np.ndarray:
    """Copy `count` items from a raw device pointer to a new numpy array (CSR::D2H)."""
    out = np.zeros(count, dtype)
    if count:
        rc = _lib().mhb_memcpy_d2h(_hp(out), C.c_void_p(dev_ptr), out.nbytes)
        if rc:
            raise MhbError(rc, "mhb_memcpy_d2h failed")
    return out


class DeviceView:
    """A non-owning view of `count` items at a raw device pointer (e.g. a rank's shard of B
    inside its peer window); same duck type as DeviceArray.  upload() is CSR::H2D's copy."""

    def __init__(self, ptr: int, count: int, dtype):
        self.ptr, self.count, self.dtype = int(ptr), int(count), np.dtype(dtype)
        self.nbytes = self.count * self.dtype.itemsize

    def data_ptr(self) -> int:
        return self.ptr

    def numel(self) -> int:
        return self.count

    def numpy(self) -> np.ndarray:
        return _d2h(self.ptr, self.count, self.dtype)

    def upload(self, host: np.ndarray):
        host = np.ascontiguousarray(host, self.dtype)
        assert host.size == self.count
        if self.nbytes:
            rc = _lib().mhb_memcpy_h2d(C.c_void_p(self.ptr), _hp(host), self.nbytes)
            if rc:
                raise MhbError(rc, "mhb_memcpy_h2d failed")


class DeviceArray:
    """Minimal owner of a device buffer for callers without torch (CSR::H2D, src/CSR.cu:97-105)."""

    def __init__(self, host: np.ndarray | None = None, count: int = 0, dtype=np.int32):
        if host is not None:
            host = np.ascontiguousarray(host)
            count, dtype = host.size, host.dtype
        self.count, self.dtype = int(count), np.dtype(dtype)
        self.nbytes = self.count * self.dtype.itemsize
        p = C.c_void_p()
        rc = _lib().mhb_device_alloc(C.byref(p), max(self.nbytes, 1))
        if rc:
            raise MhbError(rc, "mhb_device_alloc failed")
        self.ptr = p.value
        if host is not None and self.nbytes:
            rc = _lib().mhb_memcpy_h2d(C.c_void_p(self.ptr), _hp(host), self.nbytes)
            if rc:
                raise MhbError(rc, "mhb_memcpy_h2d failed")

    def data_ptr(self) -> int:
        return self.ptr

    def numel(self) -> int:
        return self.count

    def numpy(self) -> np.ndarray:
        return _d2h(self.ptr, self.count, self.dtype)

    def free(self):
        if self.ptr:
            _lib().mhb_device_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _itemsize(x) -> int:
    if isinstance(x, DeviceArray):
        return x.dtype.itemsize
    return x.element_size()


def _alloc_like(x, count, dtype):
    """Device buffer of `count` items in the same framework as x (dtype None -> x's dtype)."""
    if isinstance(x, DeviceArray):
        return DeviceArray(count=count, dtype=x.dtype if dtype is None else dtype)
    import torch
    tdt = x.dtype if dtype is None else {np.int32: torch.int32}[dtype]
    return torch.empty(count, dtype=tdt, device=x.device)


_default_tool: Tool | None = None


def MH_spgemm(A: CSR, B: CSR | None = None, tools: Tool | None = None) -> CSR:
    """Drop-in for the reference's ``MH_spgemm(A, B, C, Timing, Tool)`` on host CSR objects:
    returns C (ptr exclusive-scanned, col ascending per row); ``tools.timing`` holds the
    per-stage times that the reference returns through its ``Timing`` argument."""
    global _default_tool
    if tools is None:
        if _default_tool is None:
            _default_tool = Tool(0)
        tools = _default_tool
    return tools.spgemm_host(A, B)
