"""Matrix Market coordinate reader with the semantics of the reference's readMtxFile
(inc/mmio_read.h:34-158): 1-based -> 0-based indices, `symmetric` / `hermitian` files
expanded by mirroring every off-diagonal entry with the same value, `pattern` -> 1.0,
`complex` -> real part, `integer` -> converted, rows sorted by (column, value); entries
are NOT merged (the format forbids duplicates).  `skew-symmetric` is read as stored, like
the reference (mm_is_symmetric is false for it)."""
from __future__ import annotations

import numpy as np

from .csr import CSR


def read_mtx(path: str, dtype=np.float64) -> tuple[CSR, bool]:
    """-> (A, is_symmetric).  Raises ValueError on a malformed banner (the reference prints
    'Could not process Matrix Market banner.' and returns -1)."""
    with open(path, "rb") as f:
        banner = f.readline().decode().strip().split()
        if len(banner) < 5 or banner[0] != "%%MatrixMarket" or banner[1].lower() != "matrix":
            raise ValueError("Could not process Matrix Market banner.")
        fmt, field, sym = banner[2].lower(), banner[3].lower(), banner[4].lower()
        if fmt != "coordinate":
            raise ValueError("only coordinate (sparse) Matrix Market files are supported")
        line = f.readline()
        while line.startswith(b"%") or not line.strip():
            line = f.readline()
        m, n, nz = (int(x) for x in line.split()[:3])
        data = np.loadtxt(f, ndmin=2, dtype=np.float64, max_rows=nz) if nz else np.zeros((0, 3))
    if data.shape[0] != nz:
        raise ValueError(f"expected {nz} entries, found {data.shape[0]}")
    r = data[:, 0].astype(np.int64) - 1
    c = data[:, 1].astype(np.int64) - 1
    if field == "pattern":
        v = np.ones(nz)
    elif field in ("real", "integer", "complex"):
        v = data[:, 2].copy()  # complex: only the real part is stored (mmio_read.h:96-99)
    else:
        raise ValueError(f"unsupported field {field}")
    symmetric = sym in ("symmetric", "hermitian")
    if symmetric:
        off = r != c
        r, c, v = np.concatenate([r, c[off]]), np.concatenate([c, r[off]]), np.concatenate([v, v[off]])
    order = np.lexsort((v, c, r))  # per row: sort by (column, value) like std::sort on pairs
    r, c, v = r[order], c[order], v[order]
    ptr = np.zeros(m + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=m), out=ptr[1:])
    return CSR(m, n, ptr, c, v.astype(dtype)), sym == "symmetric"


def write_mtx(path: str, A: CSR, symmetric_lower: bool = False, field: str = "real") -> None:
    """Write A as 'matrix coordinate <field> general' (or the lower triangle as 'symmetric')."""
    rows = np.repeat(np.arange(A.M), np.diff(A.ptr))
    cols, vals = A.col.astype(np.int64), A.val
    if symmetric_lower:
        keep = rows >= cols
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {'symmetric' if symmetric_lower else 'general'}\n")
        f.write(f"{A.M} {A.N} {rows.size}\n")
        for i, j, x in zip(rows, cols, vals):
            if field == "pattern":
                f.write(f"{i + 1} {j + 1}\n")
            elif field == "integer":
                f.write(f"{i + 1} {j + 1} {int(x)}\n")
            else:
                f.write(f"{i + 1} {j + 1} {float(x)!r}\n")
