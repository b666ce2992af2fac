"""Command-line driver equivalent to the reference's `./spgemm <file.mtx>`
(src/main.cu:74-217): read a Matrix Market file, B = A (or A^T with --aat, the reference's
compile-time AAT switch, inc/common.h:37 + src/utils.cpp:20-46), run MH_spgemm once after a
warm-up, print the reference's report lines.

    python -m mh_spgemm_b200.cli matrix.mtx [--aat] [--iters N] [--out C.mtx]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import api
from .mmio import read_mtx, write_mtx


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="spgemm")
    ap.add_argument("file")
    ap.add_argument("--aat", action="store_true", help="C = A * A^T instead of A * A")
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--out", default=None, help="write C as a Matrix Market file")
    args = ap.parse_args(argv)
    try:
        A, is_sym = read_mtx(args.file)
    except (OSError, ValueError) as e:
        print(str(e))
        return -1
    if not args.aat and A.M != A.N:
        print("C=AA must have rowA = colA. Exit.")  # src/main.cu:92-96
        return 0
    print("--------------------------SpGEMM Start!!!--------------------------")
    B = A.transpose() if (args.aat and not is_sym) else A
    int_result = int(np.diff(B.ptr).astype(np.int64)[A.col].sum())  # src/main.cu:102-107
    name = os.path.splitext(os.path.basename(args.file))[0]
    print(f"Matrix {name} ({A.M} , {B.N}) nnz:{A.nnz}")
    print(f"SpGEMM intermediate result = {int_result}")
    try:
        tools = api.Tool(0)
        tools.spgemm_host(A, B)  # warm-up (the reference warms the GPU with a dummy kernel)
        tot = {}
        for _ in range(max(args.iters, 1)):
            C = tools.spgemm_host(A, B)
            for k, v in tools.timing.as_dict().items():
                tot[k] = tot.get(k, 0.0) + v / max(args.iters, 1)
        print(f"C.nnz = {C.nnz}")
        print("  -------------time-------------")
        for label, key in (("mem_alloc: \t\t", "mem_alloc"), ("form_mask_matrix_B: ", "Form_mask_matrix_B"),
                           ("symbolic_binning: \t", "symbolic_binning"), ("calculate_C_nnz: \t", "Calculate_C_nnz"),
                           ("malloc_C_col_val: \t", "Malloc_C_col_val"), ("numeric_binning: \t", "numeric_binning"),
                           ("numeric: \t\t", "Numeric")):
            print(f"    {label}{tot[key]:.3f}ms")
        print("  ------------------------------")
        total = tot["total"]  # mask build included (the reference's getTotal() leaves it out)
        print(f"MH-SpGEMM runtime is {total:.3f}ms, Gflops is {2.0 * int_result / (total * 1e6):.2f}")
        if args.out:
            write_mtx(args.out, C)
    except api.MhbError as e:
        print(f"MH-SpGEMM failed!!! {e}")
        return 1
    print("--------------------------SpGEMM   End!!!--------------------------")
    return 0


if __name__ == "__main__":
    sys.exit(main())
