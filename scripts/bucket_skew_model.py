"""Host-side MODEL (not a measurement) of the rank-inside-bucket loop of k_num_hash_list on the webbase-like
R-MAT: for every row of the 128 ... 1 024-slot bins, the keys of the C row are dealt to the kernel's linear
buckets ((col - first col) >> sh, S/4 buckets for the row's table S) and the loop's trip count -- the sum over
buckets of size^2 -- is compared with what finer buckets and what equal-depth (sampled-splitter) buckets would
give.  python scripts/bucket_skew_model.py  -> table for profiles/r2_bucket_skew_model.md"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mh_spgemm_b200  # noqa: E402,F401
from mh_spgemm_b200 import generators as G  # noqa: E402


def ceil_log2(v):
    return 0 if v <= 1 else int(v - 1).bit_length()


def main():
    A = G.rmat()
    S = sp.csr_matrix((np.ones(A.nnz), A.col, A.ptr), shape=(A.M, A.N))
    C = (S @ S).tocsr()
    C.sort_indices()
    n_all = np.diff(C.indptr)
    rng = np.random.default_rng(1)
    print("| bin (nnz of the C row) | rows | mean nnz | loop trips per key: S/4 linear (ships) | S/2 linear | S linear | S/4 equal-depth | keys in the largest bucket (S/4 linear), mean / max |")
    print("|---|---|---|---|---|---|---|---|")
    for lo, hi in ((25, 80), (81, 160), (161, 320), (321, 640), (641, 1280)):
        rows = np.flatnonzero((n_all >= lo) & (n_all <= hi))
        pick = rng.choice(rows, min(rows.size, 4000), replace=False)
        acc = {k: 0.0 for k in ("q4", "q2", "q1", "eq")}
        keys = 0
        big = []
        for r in pick:
            c = C.indices[C.indptr[r]:C.indptr[r + 1]].astype(np.int64)
            n = c.size
            lS = max(5, ceil_log2((n * 8 + 4) // 5))
            W = int(c[-1] - c[0] + 1)
            for name, bits in (("q4", lS - 2), ("q2", lS - 1), ("q1", lS)):
                sh = max(0, ceil_log2(W) - bits)
                cnt = np.bincount((c - c[0]) >> sh)
                acc[name] += float((cnt.astype(np.float64) ** 2).sum())
                if name == "q4":
                    big.append(int(cnt.max()))
            nb = 1 << (lS - 2)
            per = -(-n // nb)  # equal-depth buckets: ceil(n / NB) keys each
            full, rest = divmod(n, per)
            acc["eq"] += full * per * per + rest * rest
            keys += n
        print(f"| {lo}-{hi} | {rows.size} | {n_all[rows].mean():.0f} | {acc['q4'] / keys:.2f} | {acc['q2'] / keys:.2f} | "
              f"{acc['q1'] / keys:.2f} | {acc['eq'] / keys:.2f} | {np.mean(big):.1f} / {np.max(big)} |")


if __name__ == "__main__":
    main()
