"""Seeded synthetic CSR generators shaped like the reference's benchmark inputs.

The reference ships no matrices; its suite is a list of SuiteSparse
names (16matrix.txt:1-16).  These generators produce canonical CSR (sorted,
duplicate-free rows -- what inc/mmio_read.h:113-150 guarantees for the reference) with
rows / nnz / row-length distribution matched to those shapes.
All use ``numpy.random.default_rng(seed)`` (PCG64) so inputs are identical on every box.
"""
from __future__ import annotations

import numpy as np

from .csr import CSR


def poisson2d(n: int = 256, dtype=np.float64) -> CSR:
    """2-D 5-point Poisson on an n x n grid: diag 4, off-diag -1 (BASELINE configs[0]).
    A*A is exact in floating point (small integers)."""
    M = n * n
    idx = np.arange(M, dtype=np.int64)
    x, y = idx % n, idx // n
    rows = [idx]
    cols = [idx]
    vals = [np.full(M, 4.0)]
    for ok, off in ((x > 0, -1), (x < n - 1, 1), (y > 0, -n), (y < n - 1, n)):
        rows.append(idx[ok])
        cols.append(idx[ok] + off)
        vals.append(np.full(int(ok.sum()), -1.0))
    return CSR.from_coo(M, M, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), dtype=dtype)


def fem3d(nx: int = 8, ny: int = 8, nz: int = 325, dof: int = 3, seed: int = 1, dtype=np.float64) -> CSR:
    """27-point stencil on an nx*ny*nz node grid with dense dof x dof blocks ('cant'-like,
    BASELINE configs[1]): default 62,400 rows, 4,238,388 nnz, values U[0.5,1.5)."""
    nn = nx * ny * nz
    node = np.arange(nn, dtype=np.int64)
    x, y, z = node % nx, (node // nx) % ny, node // (nx * ny)
    nr, nc = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny)
                      & (z + dz >= 0) & (z + dz < nz))
                nr.append(node[ok])
                nc.append(node[ok] + dx + dy * nx + dz * nx * ny)
    nr, nc = np.concatenate(nr), np.concatenate(nc)
    a, b = np.meshgrid(np.arange(dof), np.arange(dof), indexing="ij")
    rows = (nr[:, None] * dof + a.ravel()[None, :]).ravel()
    cols = (nc[:, None] * dof + b.ravel()[None, :]).ravel()
    return CSR.from_coo(nn * dof, nn * dof, rows, cols, rng=np.random.default_rng(seed), dtype=dtype)


def fem3d_perturbed(nx: int = 8, ny: int = 8, nz: int = 325, dof: int = 3, drop: float = 0.10, seed: int = 1,
                    dtype=np.float64) -> CSR:
    """fem3d with `drop` of its off-diagonal entries removed at random (diagonal kept): the dof
    rows of a node are no longer exact twins and the runs of a row are ragged -- the cant-like
    shape without the regularity the twin-row kernels key on ("best-case-only")."""
    A = fem3d(nx, ny, nz, dof, seed, dtype)
    rng = np.random.default_rng(seed + 1000)
    rows = np.repeat(np.arange(A.M, dtype=np.int64), np.diff(A.ptr))
    keep = (rng.random(A.nnz) >= drop) | (rows == A.col)
    ptr = np.zeros(A.M + 1, np.int64)
    np.cumsum(np.bincount(rows[keep], minlength=A.M), out=ptr[1:])
    return CSR(A.M, A.N, ptr, A.col[keep], A.val[keep])


def rmat(scale: int = 20, n: int = 1_000_005, draws: int = 3_300_000, a: float = 0.48, b: float = 0.17,
         c: float = 0.17, seed: int = 2, dtype=np.float64) -> CSR:
    """R-MAT edges on a 2^scale grid, cut to n rows/cols, duplicates merged
    ('webbase-1M'-like, BASELINE configs[2]); values U[0.5,1.5)."""
    rng = np.random.default_rng(seed)
    r = np.zeros(draws, np.int64)
    cidx = np.zeros(draws, np.int64)
    for _ in range(scale):
        u = rng.random(draws)
        rbit = u >= a + b                      # quadrants c,d -> lower half
        cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        r = (r << 1) | rbit
        cidx = (cidx << 1) | cbit
    keep = (r < n) & (cidx < n)
    return CSR.from_coo(n, n, r[keep], cidx[keep], rng=rng, dtype=dtype)


def rmat_device(scale: int, n: int, draws: int, a: float, b: float, c: float, seed: int, device, dtype=np.float64) -> CSR:
    """The same R-MAT recipe drawn and canonicalised with torch on `device` (Philox stream of a seeded
    torch.Generator, so every rank of a job gets the same matrix without shipping it): what makes the
    full-size BASELINE configs[4] (scale 24: 268 M draws) practical -- numpy needs minutes per process for
    it while an 8-GPU box waits.  A DIFFERENT random stream than rmat(): the two give different matrices of
    the same distribution.  Returns host arrays (the bench partitions on the host)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    r = torch.zeros(draws, dtype=torch.int32, device=dev)
    ci = torch.zeros(draws, dtype=torch.int32, device=dev)
    for _ in range(scale):
        u = torch.rand(draws, generator=g, device=dev, dtype=torch.float32)
        rbit = u >= a + b
        cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        r = (r << 1) | rbit.to(torch.int32)
        ci = (ci << 1) | cbit.to(torch.int32)
        del u, rbit, cbit
    keep = (r < n) & (ci < n)
    key = r[keep].to(torch.int64) * n + ci[keep].to(torch.int64)
    del r, ci, keep
    key = torch.unique(key)  # sorted, duplicates merged
    rows = key // n
    cols = (key - rows * n).to(torch.int32)
    del key
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    del rows
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    vals = torch.rand(cols.numel(), generator=g, device=dev, dtype=tdt) + 0.5
    return CSR(n, n, ptr.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy().astype(dtype, copy=False))


def banded_random(M: int, per_row: int, halfband: int, seed: int = 3, dtype=np.float64) -> CSR:
    """Near-regular random rows with locality (cage/mac_econ-like): per_row draws per row
    within +-halfband of the diagonal, plus the diagonal."""
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(M, dtype=np.int64), per_row)
    cols = rows + rng.integers(-halfband, halfband + 1, rows.size)
    cols = np.clip(cols, 0, M - 1)
    d = np.arange(M, dtype=np.int64)
    return CSR.from_coo(M, M, np.concatenate([rows, d]), np.concatenate([cols, d]), rng=rng, dtype=dtype)


def uniform_random(M: int, N: int, nnz: int, seed: int = 4, dtype=np.float64) -> CSR:
    """Uniformly random (Erdos-Renyi) M x N pattern -- rectangular inputs for A*B tests."""
    rng = np.random.default_rng(seed)
    return CSR.from_coo(M, N, rng.integers(0, M, nnz), rng.integers(0, N, nnz), rng=rng, dtype=dtype)


def triangular_grid(n: int, seed: int = 5, dtype=np.float64) -> CSR:
    """6-neighbour triangulated n x n grid (delaunay_n24-like at n=4096)."""
    M = n * n
    idx = np.arange(M, dtype=np.int64)
    x, y = idx % n, idx // n
    rows, cols = [idx], [idx]
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1)):
        ok = (x + dx >= 0) & (x + dx < n) & (y + dy >= 0) & (y + dy < n)
        rows.append(idx[ok])
        cols.append(idx[ok] + dx + dy * n)
    return CSR.from_coo(M, M, np.concatenate(rows), np.concatenate(cols),
                        rng=np.random.default_rng(seed), dtype=dtype)


def road_grid(n: int, keep: float = 0.6, seed: int = 6, dtype=np.float64) -> CSR:
    """Thinned symmetric 2-D grid at degree ~2.4 (GAP-road-like); no diagonal."""
    rng = np.random.default_rng(seed)
    M = n * n
    idx = np.arange(M, dtype=np.int64)
    x, y = idx % n, idx // n
    rows, cols = [], []
    for ok, off in ((x < n - 1, 1), (y < n - 1, n)):
        src = idx[ok]
        sel = rng.random(src.size) < keep
        rows += [src[sel], src[sel] + off]
        cols += [src[sel] + off, src[sel]]
    return CSR.from_coo(M, M, np.concatenate(rows), np.concatenate(cols), rng=rng, dtype=dtype)


def with_dense_rows(A: CSR, nrows: int, length: int, seed: int = 7) -> CSR:
    """Overwrite `nrows` rows with `length` random columns each (stress for the large-row
    and global-memory paths)."""
    rng = np.random.default_rng(seed)
    pick = rng.choice(A.M, nrows, replace=False)
    rows = np.repeat(np.arange(A.M, dtype=np.int64), np.diff(A.ptr))
    keep = ~np.isin(rows, pick)
    er = np.repeat(pick.astype(np.int64), length)
    ec = rng.integers(0, A.N, er.size)
    return CSR.from_coo(A.M, A.N, np.concatenate([rows[keep], er]),
                        np.concatenate([A.col[keep].astype(np.int64), ec]), rng=rng, dtype=A.val.dtype)


# name -> (constructor, kwargs): synthetic analogs of the 16matrix.txt suite shapes
# (SURVEY.md section 8d targets in the comment: rows / nnz).
SUITE = {
    "pdb1HYS":          (fem3d, dict(nx=6, ny=6, nz=202, dof=5, seed=11)),      # 36,417 / 4.34M
    "pwtk":             (fem3d, dict(nx=4, ny=4, nz=4540, dof=3, seed=12)),     # 217,918 / 11.6M
    "webbase-1M":       (rmat, dict(scale=20, n=1_000_005, draws=3_300_000, seed=2)),
    "cage12":           (banded_random, dict(M=130_228, per_row=15, halfband=3000, seed=13)),
    "cant":             (fem3d, dict(nx=8, ny=8, nz=325, dof=3, seed=1)),       # 62,451 / 4.0M
    "hood":             (fem3d, dict(nx=3, ny=4, nz=6126, dof=3, seed=14)),     # 220,542 / 10.8M
    "rma10":            (fem3d, dict(nx=4, ny=4, nz=976, dof=3, seed=15)),      # 46,835 / 2.37M
    "scircuit":         (rmat, dict(scale=18, n=170_998, draws=1_700_000, a=0.40, b=0.20, c=0.20, seed=16)),
    "shipsec1":         (fem3d, dict(nx=4, ny=4, nz=2935, dof=3, seed=17)),     # 140,874 / 7.8M
    "cop20k_A":         (banded_random, dict(M=121_192, per_row=21, halfband=20000, seed=18)),
    "mac_econ_fwd500":  (banded_random, dict(M=206_500, per_row=5, halfband=500, seed=19)),
    "offshore":         (banded_random, dict(M=259_789, per_row=15, halfband=5000, seed=20)),
    "wb-edu":           (rmat, dict(scale=24, n=9_845_725, draws=60_000_000, a=0.46, b=0.17, c=0.17, seed=21)),
    "cage15":           (banded_random, dict(M=5_154_859, per_row=18, halfband=50000, seed=22)),
    "GAP-road":         (road_grid, dict(n=4894, keep=0.6, seed=23)),           # 23.9M rows
    "delaunay_n24":     (triangular_grid, dict(n=4096, seed=24)),               # 16.8M rows
}


def suite(name: str, **over) -> CSR:
    fn, kw = SUITE[name]
    kw = dict(kw, **over)
    return fn(**kw)
