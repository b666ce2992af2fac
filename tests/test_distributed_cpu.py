"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row partition balanced by
intermediate products, the packed-B broadcast, slice offsets by all-gather, and the
concatenation by row_ptr offset.  The per-rank SpGEMM is the oracle here (the CUDA path
cannot run without a GPU); the -m gpu suite runs the same flow on devices."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mh_spgemm_b200  # noqa: F401
from mh_spgemm_b200 import distributed as D, generators as G
from mh_spgemm_b200.csr import CSR


def test_partition_balances_products():
    A = G.rmat(13, 8000, 40000, seed=3)
    w = D.row_work(A, A)
    assert w.sum() == int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
    for parts in (1, 2, 4, 8):
        b = D.partition_rows(w, parts)
        assert b[0] == 0 and b[-1] == A.M and np.all(np.diff(b) >= 0)
        sums = np.add.reduceat(np.append(w, 0), np.minimum(b[:-1], A.M))[:parts]
        sums = np.array([w[b[g]:b[g + 1]].sum() for g in range(parts)])
        assert sums.sum() == w.sum()
        # no block exceeds the ideal share by more than the heaviest single row
        assert sums.max() <= w.sum() / parts + w.max()


def test_partition_degenerate():
    assert D.partition_rows(np.zeros(10, np.int64), 4).tolist()[-1] == 10
    assert D.partition_rows(np.array([5], np.int64), 3).tolist() == [0, 0, 0, 1] or \
        D.partition_rows(np.array([5], np.int64), 3)[-1] == 1


def test_pack_roundtrip():
    B = G.uniform_random(50, 70, 300, seed=4)
    buf, _ = D.pack_b(B)
    p, c, v = D.b_views(buf, B.M, B.nnz, torch.float64)
    assert np.array_equal(p.numpy(), B.ptr) and np.array_equal(c.numpy(), B.col)
    assert np.array_equal(v.numpy(), B.val)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    orc = Oracle()
    A = G.rmat(12, 4000, 20000, seed=5)  # same seeded input on every rank
    B = A
    bounds = D.partition_rows(D.row_work(A, B), world)
    blk = A.rows(int(bounds[rank]), int(bounds[rank + 1]))
    packed, _ = D.pack_b(B)
    buf = packed if rank == 0 else torch.zeros_like(packed)
    D.exchange_B(buf, world, src=0)
    bp, bc, bv = D.b_views(buf, B.M, B.nnz, torch.float64)
    Bl = CSR(B.M, B.N, bp.numpy(), bc.numpy(), bv.numpy())
    Cp, Cc, Cv = orc.spgemm(blk, Bl)
    off, total = D.slice_offsets(int(Cp[-1]), rank, world, torch.device("cpu"))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((Cp, Cc, Cv, off, total), gathered, dst=0)
    if rank == 0:
        fp, fc, fv = orc.spgemm(A, B)
        gp, gc, gv = D.concat_slices([(g[0], g[1], g[2]) for g in gathered])
        ok = (np.array_equal(gp, fp) and np.array_equal(gc, fc) and np.array_equal(gv, fv)
              and all(g[4] == int(fp[-1]) for g in gathered)
              and [g[3] for g in gathered] == [int(fp[bounds[r]]) for r in range(world)])
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_sharded_flow():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
