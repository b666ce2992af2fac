"""Seeded inputs of the golden fixtures (shared by make_golden.py and the tests)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mh_spgemm_b200  # noqa: E402,F401
from mh_spgemm_b200 import generators as G  # noqa: E402

# name -> () -> (A, B or None).  "full": the whole reference output is stored (small
# cases); "sum": only nnz + SHA-256 of row_ptr / col_idx + value checksums are stored.
SMALL = {
    "fem_3x3x6x2": lambda: (G.fem3d(3, 3, 6, 2, seed=31), None),
    "rmat_s10": lambda: (G.rmat(10, 1000, 4000, seed=32), None),
    "uniform_300": lambda: (G.uniform_random(300, 300, 2400, seed=33), None),
    "rect_200x300x500": lambda: (G.uniform_random(200, 300, 1500, seed=34), G.uniform_random(300, 500, 2500, seed=35)),
    "banded_400": lambda: (G.banded_random(400, 6, 12, seed=36), None),
}
LARGE = {
    "dense_rows": lambda: (G.with_dense_rows(G.uniform_random(3000, 3000, 30000, seed=8), 6, 1500, seed=9), None),
    "rmat_s14": lambda: (G.rmat(14, 16000, 60000, seed=6), None),
    "fem_small": lambda: (G.fem3d(4, 4, 10, 3, seed=5), None),
    "F_cant_like": lambda: (G.fem3d(), None),
    "R_webbase_like": lambda: (G.rmat(), None),
}
# The reference itself faults (illegal memory access in its numeric stage) when no C row
# has more than 22 nnz: its k_init_group_size<2> launch gets an empty grid
# (src/main.cu:49-50) and the pending launch error makes the following CUB scan
# (src/main.cu:55) return early, so C.nnz is garbage.  Poisson inputs are therefore
# pinned by the host oracle + scipy only.
REFERENCE_FAULTS = {
    "poisson_32": lambda: (G.poisson2d(32), None),
}

# Inputs of the transpose fixtures (the reference's host matrix_transposition, src/utils.cpp:20-46,
# run in the build container by make_golden_transpose.py -- it needs no GPU).
TRANSPOSE = {
    "rect_300x200": lambda: G.uniform_random(300, 200, 2500, seed=3),
    "rmat_s10": lambda: G.rmat(10, 1000, 4000, seed=32),
    "wide_40x70000": lambda: G.uniform_random(40, 70000, 3000, seed=41),   # 3 radix passes, many empty columns
    "tall_5000x9": lambda: G.uniform_random(5000, 9, 20000, seed=42),      # 1 radix pass, long T rows
    "fem_3x3x6x2": lambda: G.fem3d(3, 3, 6, 2, seed=31),
}

# BASELINE configs[3]: the twelve suite analogs that fit a test run (the four large shapes --
# wb-edu, cage15, GAP-road, delaunay_n24 -- are in SUITE_LARGE and run only with MHB_SLOW=1).
SUITE12 = ["pdb1HYS", "pwtk", "webbase-1M", "cage12", "cant", "hood", "rma10", "scircuit", "shipsec1", "cop20k_A",
           "mac_econ_fwd500", "offshore"]
SUITE_LARGE = ["wb-edu", "cage15", "GAP-road", "delaunay_n24"]
SUITE = {name: (lambda name=name: (G.suite(name), None)) for name in SUITE12 + SUITE_LARGE}
