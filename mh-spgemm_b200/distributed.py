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


def partition_rows(work: np.ndarray, nparts: int, nnz_cap: int | None = None) -> np.ndarray:
    """Boundaries b[0..nparts] of contiguous row blocks whose work sums are as equal as the
    g/G quantiles of the prefix sum allow.  Rows without work are free."""
    M = work.size
    pre = np.concatenate([[0], np.cumsum(work, dtype=np.int64)])
    total = int(pre[-1])
    b = np.zeros(nparts + 1, np.int64)
    b[-1] = M
    for g in range(1, nparts):
        target = total * g // nparts
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


def slice_offsets(nnz_local: int, rank: int, world: int, device):
    """All-gather the per-rank nnz(C slice) (int64) -> (offset of this rank's slice in the
    global col/val arrays, total nnz(C))."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor([nnz_local], dtype=torch.int64, device=device)
    if world > 1:
        allnnz = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allnnz, mine)
    else:
        allnnz = mine
    sizes = allnnz.cpu().numpy()
    return int(sizes[:rank].sum()), int(sizes.sum())


class ShardedSpGEMM:
    """C = A*B with A row-sharded over the ranks.  Each rank owns a Tool (one handle per
    device) and the device arrays of its row block."""

    def __init__(self, tool, rank: int, world: int, device=None):
        import torch
        self.tool, self.rank, self.world = tool, rank, world
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def step(self, A_blk, Bbuf, K: int, N: int, nnzB: int, val_dtype, src: int = 0):
        """One sharded SpGEMM.  A_blk = (M_local, ptr, col, val) device tensors of this
        rank's rows; Bbuf = packed B image (valid on `src`, receive buffer elsewhere).
        Returns (C_ptr, C_col, C_val, slice_offset, total_nnz)."""
        import torch
        exchange_B(Bbuf, self.world, src)
        bp, bc, bv = b_views(Bbuf, K, nnzB, val_dtype)
        Ml, ap, ac, av = A_blk
        cp, nnz = self.tool.symbolic(Ml, K, N, ap, ac, bp, bc)
        ccol = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)
        cval = torch.empty(max(nnz, 1), dtype=val_dtype, device=self.device)
        self.tool.numeric_into(av, bv, ccol, cval)
        off, total = slice_offsets(nnz, self.rank, self.world, self.device)
        return cp, ccol[:nnz], cval[:nnz], off, total


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
