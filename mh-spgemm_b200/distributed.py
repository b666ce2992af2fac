"""Row-sharded SpGEMM across the GPUs of one box: one process per GPU, torch.distributed
(NCCL over NVLink / NVSwitch; gloo in the CPU tests) for the plumbing.

The reference is single-GPU (no NCCL/MPI/peer copies).  Gustavson
rows are independent, so C(i,:) needs A(i,:) and all of B: A is partitioned into contiguous
row blocks balanced by intermediate-product count (the k_calculate_flop_tmp quantity,
inc/Form_mask_matrix_B.cuh:56-95), B is broadcast from its owner, every rank runs the
single-GPU pipeline on its block and keeps its own CSR slice of C; the global row_ptr is
the concatenation shifted by the all-gathered slice sizes (int64).  There is no collective
in the symbolic or numeric phase.
"""
from __future__ import annotations

import numpy as np

from .csr import CSR


def row_work(A: CSR, B: CSR) -> np.ndarray:
    """Intermediate products per row of A (int64) -- the balancing weight."""
    blen = np.diff(B.ptr).astype(np.int64)
    w = blen[A.col]
    out = np.zeros(A.M, np.int64)
    nz = np.diff(A.ptr) > 0
    if w.size:
        out[nz] = np.add.reduceat(w, A.ptr[:-1][nz])
    return out


def row_cost(work: np.ndarray) -> np.ndarray:
    """Balancing weight per row of A: its intermediate products, inflated for long rows.
    Measured on 8 B200 (profiles/r2_scaling.md): long rows go through larger hash tables, more probe
    rounds and longer sorts, and the rank that owns the hub rows set the step time.  R-MAT scale 24
    split by the first model (p * (1 + min(p, 6000) / 1500), fitted at scale 22): the ranks spent
    49 / 62 / 71 / 84 / 136 ps per product at 341 / 600 / 1 010 / 1 362 / 2 647 products per row on
    average -- linear, 36 + 0.038 p -- and rank 0 (the hub rows) took 128 ms against 91 ms for the rest.
    cost = p * (1 + min(p, 12000) / 950) is that line."""
    w = work.astype(np.float64)
    return w * (1.0 + np.minimum(w, 12000.0) / 950.0)


def snap_to_pattern_change(A: CSR, bounds: np.ndarray) -> np.ndarray:
    """Move every interior boundary forward to the next row whose column list differs from the
    row before it, so that a run of twin rows (the dof rows of one FEM node) is never cut
    between two ranks: the numeric kernel folds up to three twin rows into one pass, and a
    block that starts in the middle of a run loses that (8 B200, r2s: 0.41 ms instead of
    0.29 ms for the numeric kernel on the ranks whose first row was not a node boundary)."""
    b = np.array(bounds, np.int64)
    for g in range(1, len(b) - 1):
        r = int(b[g])
        while 0 < r < A.M:
            a0, a1, a2 = int(A.ptr[r - 1]), int(A.ptr[r]), int(A.ptr[r + 1])
            if a2 - a1 != a1 - a0 or a2 == a1 or not np.array_equal(A.col[a0:a1], A.col[a1:a2]):
                break
            r += 1
        b[g] = r
    return np.maximum.accumulate(b)


def partition_rows(work: np.ndarray, nparts: int, nnz_cap: int | None = None) -> np.ndarray:
    """Boundaries b[0..nparts] of contiguous row blocks whose work sums are as equal as the
    g/G quantiles of the prefix sum allow.  Rows without work are free."""
    M = work.size
    pre = np.concatenate([[0], np.cumsum(work, dtype=np.float64 if work.dtype.kind == "f" else np.int64)])
    total = pre[-1]
    b = np.zeros(nparts + 1, np.int64)
    b[-1] = M
    for g in range(1, nparts):
        target = total * g / nparts if work.dtype.kind == "f" else int(total) * g // nparts
        b[g] = int(np.searchsorted(pre, target, side="left"))
    b = np.maximum.accumulate(np.minimum(b, M))
    return b


def pack_b(B: CSR):
    """One contiguous byte image of B (val | ptr | col, 8-byte aligned first) so that the
    broadcast is a single collective."""
    import torch
    nv = B.val.nbytes
    npb = B.ptr.nbytes
    pad = (-(nv + npb)) % 8
    buf = np.empty(nv + npb + pad + B.col.nbytes, np.uint8)
    buf[:nv] = B.val.view(np.uint8)
    buf[nv:nv + npb] = B.ptr.view(np.uint8)
    buf[nv + npb + pad:] = B.col.view(np.uint8)
    return torch.from_numpy(buf), (nv, npb, pad)


def b_views(buf, K: int, nnzB: int, val_dtype):
    """(ptr, col, val) tensor views into a packed B image on the device."""
    import torch
    isz = torch.tensor([], dtype=val_dtype).element_size()
    nv = nnzB * isz
    npb = (K + 1) * 4
    pad = (-(nv + npb)) % 8
    val = buf[:nv].view(val_dtype)
    ptr = buf[nv:nv + npb].view(torch.int32)
    col = buf[nv + npb + pad:nv + npb + pad + nnzB * 4].view(torch.int32)
    return ptr, col, val


def exchange_B(Bbuf, world: int, src: int = 0):
    """The one exchange step of the path: replicate B's packed image from its owner
    (ncclBroadcast over NVLink / NVSwitch; gloo in the CPU tests)."""
    import torch.distributed as dist
    if world > 1:
        dist.broadcast(Bbuf, src=src)
    return Bbuf


def slice_offsets(nnz_local: int, rank: int, world: int, device, on_host: bool = True):
    """All-gather the per-rank nnz(C slice) (int64) -> (offset of this rank's slice in the
    global col/val arrays, total nnz(C)).  on_host=False returns 0-dim device tensors and
    does not synchronise (the step then ends without a host round trip)."""
    if world == 1:
        return 0, int(nnz_local)
    import torch
    import torch.distributed as dist
    mine = torch.tensor([nnz_local], dtype=torch.int64, device=device)
    if world > 1:
        allnnz = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allnnz, mine)
    else:
        allnnz = mine
    if not on_host:
        return allnnz[:rank].sum(), allnnz.sum()
    sizes = allnnz.cpu().numpy()
    return int(sizes[:rank].sum()), int(sizes.sum())


class SliceSizes:
    """Per-step all-gather of the ranks' nnz(C slice) with preallocated buffers: one small
    non-blocking H2D from pinned memory and one NCCL all-gather; the exclusive prefix (the
    slice offsets of the global CSR) is taken lazily by whoever assembles C."""

    def __init__(self, rank: int, world: int, device):
        import torch
        self.rank, self.world = rank, world
        if world > 1:
            self.host = torch.zeros(1, dtype=torch.int64).pin_memory() if device.type == "cuda" else torch.zeros(1, dtype=torch.int64)
            self.mine = torch.zeros(1, dtype=torch.int64, device=device)
            self.all = torch.zeros(world, dtype=torch.int64, device=device)

    def gather(self, nnz_local: int):
        import torch.distributed as dist
        self.nnz = int(nnz_local)
        if self.world > 1:
            self.host[0] = self.nnz
            self.mine.copy_(self.host, non_blocking=True)
            dist.all_gather_into_tensor(self.all, self.mine)
        return self

    def offsets(self):
        """(offset of this rank's slice, total nnz(C)) on the host (synchronises)."""
        if self.world == 1:
            return 0, self.nnz
        sizes = self.all.cpu().numpy()
        return int(sizes[:self.rank].sum()), int(sizes.sum())


class RangeExchange:
    """B exchange when B is row-sharded like A (the natural layout for C = A*A): rank r owns
    B rows [bounds[r], bounds[r+1]) and needs the rows [k0, k1) that the columns of its A
    block reference.  Every step it receives the missing pieces of col/val from their
    owners (grouped NCCL send/recv over NVLink) into one contiguous image of rows [k0, k1).
    Banded / FEM inputs exchange only a halo; for graphs whose blocks reference all of B
    this is an all-gather.  The row offsets of the gathered range depend only on the
    pattern and are rebuilt once per plan."""

    def __init__(self, rank, world, bounds, kranges, B_ptr_host, val_dtype, device):
        import torch
        self.rank, self.world, self.device = rank, world, device
        bp = np.asarray(B_ptr_host, np.int64)
        self.k0, self.k1 = int(kranges[rank][0]), int(kranges[rank][1])
        self.K_local = self.k1 - self.k0
        self.nnz_local = int(bp[self.k1] - bp[self.k0])
        self.ptr = torch.from_numpy((bp[self.k0:self.k1 + 1] - bp[self.k0]).astype(np.int32)).to(device)
        self.col = torch.empty(max(self.nnz_local, 1), dtype=torch.int32, device=device)
        self.val = torch.empty(max(self.nnz_local, 1), dtype=val_dtype, device=device)
        own0 = int(bp[bounds[rank]])
        self.recv, self.send, self.local = [], [], None
        for o in range(world):  # pieces I need, by owner
            ra, rb = max(self.k0, int(bounds[o])), min(self.k1, int(bounds[o + 1]))
            if ra >= rb:
                continue
            dst = (int(bp[ra] - bp[self.k0]), int(bp[rb] - bp[self.k0]))
            if o == rank:
                self.local = (int(bp[ra]) - own0, int(bp[rb]) - own0, dst[0], dst[1])
            elif dst[1] > dst[0]:
                self.recv.append((o, dst[0], dst[1]))
        for d in range(world):  # pieces of mine that others need
            if d == rank:
                continue
            ra, rb = max(int(kranges[d][0]), int(bounds[rank])), min(int(kranges[d][1]), int(bounds[rank + 1]))
            if ra < rb and bp[rb] > bp[ra]:
                self.send.append((d, int(bp[ra]) - own0, int(bp[rb]) - own0))
        self.bytes_received = sum(b - a for _, a, b in self.recv) * (4 + torch.tensor([], dtype=val_dtype).element_size())

    def own_views(self):
        """Views of this rank's own shard INSIDE the gathered image.  A caller that keeps its
        shard of B there (instead of in separate arrays) saves the local copy of run()."""
        if self.local is None:
            return self.col[:0], self.val[:0]
        sa, sb, da, db = self.local
        return self.col[da:db], self.val[da:db]

    def run(self, own_col, own_val):
        """own_col / own_val: this rank's shard of B (device).  Returns (ptr, col, val) of B
        rows [k0, k1); A's column indices must be shifted by -k0 (done once by the caller)."""
        import torch.distributed as dist
        ops = []
        for d, a, b in self.send:
            ops.append(dist.P2POp(dist.isend, own_col[a:b], d))
            ops.append(dist.P2POp(dist.isend, own_val[a:b], d))
        for o, a, b in self.recv:
            ops.append(dist.P2POp(dist.irecv, self.col[a:b], o))
            ops.append(dist.P2POp(dist.irecv, self.val[a:b], o))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if self.local is not None:
            sa, sb, da, db = self.local
            if own_col.data_ptr() != self.col[da:db].data_ptr():  # shard not already in place
                self.col[da:db].copy_(own_col[sa:sb])
                self.val[da:db].copy_(own_val[sa:sb])
        for r in reqs:
            r.wait()
        return self.ptr, self.col, self.val


def column_range(A: CSR) -> tuple[int, int]:
    """[k0, k1): rows of B referenced by the columns of A (the halo-aware exchange range)."""
    if A.nnz == 0:
        return 0, 0
    return int(A.col.min()), int(A.col.max()) + 1


class ShardedSpGEMM:
    """C = A*B with A row-sharded over the ranks.  Each rank owns a Tool (one handle per
    device) and the device arrays of its row block."""

    def __init__(self, tool, rank: int, world: int, device=None):
        import torch
        self.tool, self.rank, self.world = tool, rank, world
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.sizes = SliceSizes(rank, world, self.device)
        self._bind_stream()

    def _bind_stream(self):
        """The exchange (torch.distributed) runs on torch's current stream; the Tool must launch
        on the same stream, or symbolic / numeric could read B before the exchange has landed
        and torch's caching allocator would not know about the writes into ccol / cval."""
        if self.device.type == "cuda":
            import torch
            self.tool.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, A_blk, Bbuf, K: int, N: int, nnzB: int, val_dtype, src: int = 0):
        """One sharded SpGEMM.  A_blk = (M_local, ptr, col, val) device tensors of this
        rank's rows; Bbuf = packed B image (valid on `src`, receive buffer elsewhere).
        Returns (C_ptr, C_col, C_val, slice_offset, total_nnz)."""
        import torch
        self._bind_stream()
        exchange_B(Bbuf, self.world, src)
        bp, bc, bv = b_views(Bbuf, K, nnzB, val_dtype)
        Ml, ap, ac, av = A_blk
        cp, nnz = self.tool.symbolic(Ml, K, N, ap, ac, bp, bc)
        ccol = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)
        cval = torch.empty(max(nnz, 1), dtype=val_dtype, device=self.device)
        self.tool.numeric_into(av, bv, ccol, cval)
        off, total = slice_offsets(nnz, self.rank, self.world, self.device, on_host=True)
        return cp, ccol[:nnz], cval[:nnz], off, total

    def step_range(self, A_blk_shifted, plan: RangeExchange, own_col, own_val, N: int, val_dtype,
                   offsets_on_host: bool = True):
        """Same as step() with B row-sharded like A: gather the referenced row range of B
        (RangeExchange), multiply the local block (whose columns were shifted by -k0)."""
        import torch
        self._bind_stream()
        bp, bc, bv = plan.run(own_col, own_val)
        Ml, ap, ac, av = A_blk_shifted
        cp, nnz = self.tool.symbolic(Ml, plan.K_local, N, ap, ac, bp, bc[:plan.nnz_local])
        ccol = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)
        cval = torch.empty(max(nnz, 1), dtype=val_dtype, device=self.device)
        self.tool.numeric_into(av, bv, ccol, cval)
        if offsets_on_host:
            off, total = slice_offsets(nnz, self.rank, self.world, self.device, on_host=True)
        else:  # sizes stay on the device; SliceSizes.offsets() turns them into offsets on demand
            off, total = None, self.sizes.gather(nnz)
        return cp, ccol[:nnz], cval[:nnz], off, total


class Shard:
    """One rank of the row-sharded SpGEMM behind the C ABI (``mhb_shard_*``, include/mhb_spgemm.h):
    B is row-sharded like A, every rank keeps a contiguous image of the B rows its block references
    in a CUDA-IPC window and pulls the pieces it does not own out of the owners' windows over
    NVLink -- one-sided, no send/recv pairing, no host-side collective inside a step.

    `allgather(bytes) -> list[bytes]` is the caller's transport for the two fixed-size set-up
    blobs (default: torch.distributed.all_gather_object on the default group)."""

    def __init__(self, tool, rank: int, world: int, K: int, N: int, val_dtype, bounds, allgather=None):
        import ctypes as C
        from .api import MhbError
        self.C, self.MhbError = C, MhbError
        self.tool, self.L, self.rank, self.world = tool, tool.L, rank, world
        self.val_dtype = np.dtype(val_dtype)
        self.s = C.c_void_p()
        b = (C.c_longlong * (world + 1))(*[int(x) for x in bounds])
        rc = self.L.mhb_shard_create(C.byref(self.s), tool.h, rank, world, K, N, self.val_dtype.itemsize, b)
        if rc:
            raise MhbError(rc, "mhb_shard_create failed (bad bounds / world > 8?)")
        self._allgather = allgather or _allgather_object
        self._keep = None

    def _chk(self, rc):
        if rc:
            raise self.MhbError(rc, self.L.mhb_shard_last_error(self.s).decode())

    def close(self):
        if self.s:
            self.L.mhb_shard_destroy(self.s)
            self.s = self.C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up (collective) ---------------------------------------------------------------
    def build(self, M_local: int, dA_ptr, dA_col, dBown_ptr):
        """Plan of this rank: A's block (device int32 arrays, GLOBAL column ids; kept by
        reference) and the row_ptr of its shard of B (rebased to 0).  Collective: two
        all-gathers of 128-byte blobs through the caller's transport."""
        from .api import SHARD_BLOB_BYTES
        C = self.C
        self._keep = (dA_ptr, dA_col, dBown_ptr)
        self.M_local = M_local
        self._chk(self.L.mhb_shard_set_A(self.s, M_local, dA_col.numel(), dA_ptr.data_ptr(), dA_col.data_ptr(),
                                         dBown_ptr.data_ptr()))
        for phase in (1, 2):
            blob = (C.c_char * SHARD_BLOB_BYTES)()
            self._chk(self.L.mhb_shard_export(self.s, phase, blob))
            parts = self._allgather(bytes(blob))
            assert len(parts) == self.world and all(len(x) == SHARD_BLOB_BYTES for x in parts)
            self._chk(self.L.mhb_shard_import(self.s, phase, b"".join(parts)))
        return self

    def own_B(self):
        """(col view, val view) of this rank's shard of B inside its window: the caller writes
        GLOBAL column ids and values there (DeviceView.upload from the host, or a device copy)."""
        from .api import DeviceView
        C = self.C
        pc, pv, n = C.c_void_p(), C.c_void_p(), C.c_longlong()
        self._chk(self.L.mhb_shard_own_B(self.s, C.byref(pc), C.byref(pv), C.byref(n)))
        return DeviceView(pc.value or 0, n.value, np.int32), DeviceView(pv.value or 0, n.value, self.val_dtype)

    def image(self):
        """(k0, k1, nnz_image, halo_bytes_per_step): the B rows this rank's SpGEMM reads."""
        C = self.C
        k0, k1, n, hb = C.c_int(), C.c_int(), C.c_longlong(), C.c_longlong()
        self._chk(self.L.mhb_shard_image(self.s, C.byref(k0), C.byref(k1), C.byref(n), C.byref(hb)))
        return k0.value, k1.value, n.value, hb.value

    # ---- per step (stream-ordered on the Tool's stream) ---------------------------------------
    def exchange(self):
        self._chk(self.L.mhb_shard_exchange(self.s))

    def publish(self):
        """First half of exchange(): flag stores only, never waits."""
        self._chk(self.L.mhb_shard_publish(self.s))

    def pull(self):
        """Second half of exchange(): spins until the owners have published -- ranks that share
        one GPU must put a host barrier between publish() and pull() (include/mhb_spgemm.h)."""
        self._chk(self.L.mhb_shard_pull(self.s))

    def barrier(self):
        self._chk(self.L.mhb_shard_barrier(self.s))

    def symbolic(self, r_lo: int, r_hi: int, dC_ptr):
        nnz = self.C.c_longlong()
        self._chk(self.L.mhb_shard_symbolic(self.s, r_lo, r_hi, dC_ptr.data_ptr(), self.C.byref(nnz)))
        return int(nnz.value)

    def numeric_into(self, dA_val, dC_col, dC_val):
        f = self.L.mhb_shard_numeric_f64 if self.val_dtype.itemsize == 8 else self.L.mhb_shard_numeric_f32
        self._chk(f(self.s, dA_val.data_ptr(), dC_col.data_ptr(), dC_val.data_ptr()))

    def spgemm_into(self, r_lo: int, r_hi: int, dA_val, dC_ptr, dC_col, dC_val) -> int:
        """Fused symbolic + numeric of rows [r_lo, r_hi) into caller-owned C arrays
        (mhb_shard_spgemm_into_*): one host synchronisation per step in steady state.  Raises
        MhbError(code ERR_CAPACITY, .nnzC) when dC_col / dC_val are too small."""
        from .api import ERR_CAPACITY
        f = self.L.mhb_shard_spgemm_into_f64 if self.val_dtype.itemsize == 8 else self.L.mhb_shard_spgemm_into_f32
        nnz = self.C.c_longlong()
        cap = min(dC_col.numel(), dC_val.numel()) if dC_col is not None else 0
        rc = f(self.s, r_lo, r_hi, dA_val.data_ptr(), dC_ptr.data_ptr(), dC_col.data_ptr() if cap else None,
               dC_val.data_ptr() if cap else None, cap, self.C.byref(nnz))
        if rc == ERR_CAPACITY:
            e = self.MhbError(rc, self.L.mhb_shard_last_error(self.s).decode())
            e.nnzC = int(nnz.value)
            raise e
        self._chk(rc)
        return int(nnz.value)

    def spgemm_into_begin(self, r_lo: int, r_hi: int, dA_val, dC_ptr, dC_col, dC_val):
        f = (self.L.mhb_shard_spgemm_into_begin_f64 if self.val_dtype.itemsize == 8
             else self.L.mhb_shard_spgemm_into_begin_f32)
        cap = min(dC_col.numel(), dC_val.numel()) if dC_col is not None else 0
        self._chk(f(self.s, r_lo, r_hi, dA_val.data_ptr(), dC_ptr.data_ptr(), dC_col.data_ptr() if cap else None,
                    dC_val.data_ptr() if cap else None, cap))

    def spgemm_into_end(self) -> int:
        from .api import ERR_CAPACITY
        nnz = self.C.c_longlong()
        rc = self.L.mhb_shard_spgemm_into_end(self.s, self.C.byref(nnz))
        if rc == ERR_CAPACITY:
            e = self.MhbError(rc, self.L.mhb_shard_last_error(self.s).decode())
            e.nnzC = int(nnz.value)
            raise e
        self._chk(rc)
        return int(nnz.value)

    def post_size(self, nnz_local: int):
        """nnz_local = -1: post the device-side nnz(C) of the SpGEMM just queued by spgemm_into_begin."""
        self._chk(self.L.mhb_shard_post_size(self.s, int(nnz_local)))

    def repost_size(self, nnz_local: int):
        """Confirm (or, after a redo, supply) the size of the step already posted from the device."""
        self._chk(self.L.mhb_shard_repost_size(self.s, int(nnz_local)))

    def offsets(self):
        """(offset of this rank's slice in the global col / val arrays, total nnz(C), all sizes)."""
        C = self.C
        off, tot = C.c_longlong(), C.c_longlong()
        sizes = (C.c_longlong * self.world)()
        self._chk(self.L.mhb_shard_offsets(self.s, C.byref(off), C.byref(tot), sizes))
        return int(off.value), int(tot.value), [int(x) for x in sizes]

    # ---- NCCL from C++ (the broadcast layout) ---------------------------------------------------
    def init_nccl(self):
        """ncclCommInitRank inside the library; the 128-byte id travels through `allgather`."""
        C = self.C
        ident = (C.c_char * 128)()
        if self.rank == 0:
            rc = self.L.mhb_nccl_unique_id(ident)
            if rc:
                raise self.MhbError(rc, "mhb_nccl_unique_id failed (libnccl.so.2 not loadable?)")
        ident = self._allgather(bytes(ident))[0]
        self._chk(self.L.mhb_shard_init_nccl(self.s, ident))

    def broadcast(self, dbuf, nbytes: int, root: int = 0):
        self._chk(self.L.mhb_shard_broadcast(self.s, dbuf.data_ptr(), nbytes, root))


def _allgather_object(blob: bytes):
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [blob]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, blob)
    return out


def slice_rows_by_products(work: np.ndarray, r0: int, r1: int, cap: int = 2**31 - 1) -> list[tuple[int, int]]:
    """Cut rows [r0, r1) into consecutive slices whose intermediate-product sums stay <= cap.
    nnz(C slice) <= products(slice), so every slice honours the int32 CSR contract
    (SURVEY.md 'int32 limits': C of the 16 M-row R-MAT exceeds 2^31 entries)."""
    out, start, acc = [], r0, 0
    for r in range(r0, r1):
        w = int(work[r])
        if acc + w > cap and r > start:
            out.append((start, r))
            start, acc = r, 0
        acc += w
    out.append((start, r1))
    return out


def slice_rows_fast(work: np.ndarray, r0: int, r1: int, cap: int = 2**31 - 1) -> list[tuple[int, int]]:
    """Vectorised version of slice_rows_by_products (greedy cuts via searchsorted)."""
    pre = np.concatenate([[0], np.cumsum(work[r0:r1], dtype=np.int64)])
    out, start = [], 0
    n = r1 - r0
    while start < n:
        end = int(np.searchsorted(pre, pre[start] + cap, side="right")) - 1
        end = min(max(end, start + 1), n)
        out.append((r0 + start, r0 + end))
        start = end
    return out or [(r0, r1)]


def concat_slices(slices):
    """Host-side concatenation of per-rank CSR slices [(ptr, col, val), ...] by row_ptr
    offset (int64 global row_ptr) -- used by the parity tests."""
    ptrs, cols, vals = [], [], []
    off = 0
    for p, c, v in slices:
        p = np.asarray(p, np.int64)
        ptrs.append(p[:-1] + off)
        off += int(p[-1])
        cols.append(np.asarray(c))
        vals.append(np.asarray(v))
    ptrs.append(np.array([off], np.int64))
    return np.concatenate(ptrs), np.concatenate(cols), np.concatenate(vals)
