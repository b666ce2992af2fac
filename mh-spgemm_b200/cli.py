"""Command-line driver equivalent to the reference's `./spgemm <file.mtx>`
(src/main.cu:74-217): read a Matrix Market file, B = A (or A^T with --aat, the reference's
compile-time AAT switch, inc/common.h:37 + src/utils.cpp:20-46), run MH_spgemm once after a
warm-up, print the reference's report lines.

    python -m mh_spgemm_b200.cli matrix.mtx [--aat] [--iters N] [--out C.mtx] [--write DIR]
    python -m mh_spgemm_b200.cli --list 16matrix.txt --root ../matrix [--write DIR]   # process.sh

--write DIR appends the Gflops figure (two decimals, one per line) to DIR/Gflops_MH-SpGEMM.csv,
the reference's WRITE switch (inc/common.h:80, src/main.cu:201-213).  --list runs every matrix
named in a list file from ROOT/<name>/<name>.mtx, as process.sh:21-37 does (missing files are
warned about and skipped; the first failing matrix stops the run).
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
    ap.add_argument("file", nargs="?")
    ap.add_argument("--list", default=None, help="file with one matrix name per line (process.sh)")
    ap.add_argument("--root", default="../matrix", help="directory holding <name>/<name>.mtx (with --list)")
    ap.add_argument("--write", default=None, help="append Gflops to DIR/Gflops_MH-SpGEMM.csv (WRITE)")
    ap.add_argument("--aat", action="store_true", help="C = A * A^T instead of A * A")
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--out", default=None, help="write C as a Matrix Market file")
    args = ap.parse_args(argv)
    if args.list:
        return run_list(args)
    if not args.file:
        print("Invalid Arguments.")  # src/main.cu:82-86
        print("Usage:\t ./spgemm <Input File>")
        return -1
    try:
        A, is_sym = read_mtx(args.file)
    except (OSError, ValueError) as e:
        print(str(e))
        return -1
    if not args.aat and A.M != A.N:
        print("C=AA must have rowA = colA. Exit.")  # src/main.cu:92-96
        return 0
    print("--------------------------SpGEMM Start!!!--------------------------")
    try:
        tools = api.Tool(0)
        # AAT: B = A^T by the device-side transpose (mhb_transpose_*; the reference transposes on
        # the host, src/utils.cpp:20-46); a symmetric file needs none (src/main.cu:98-101)
        B = tools.transpose(A) if (args.aat and not is_sym) else A
    except api.MhbError as e:
        print(f"MH-SpGEMM failed!!! {e}")
        return 1
    int_result = int(np.diff(B.ptr).astype(np.int64)[A.col].sum())  # src/main.cu:102-107
    name = os.path.splitext(os.path.basename(args.file))[0]
    print(f"Matrix {name} ({A.M} , {B.N}) nnz:{A.nnz}")
    print(f"SpGEMM intermediate result = {int_result}")
    gflops = 0.0
    try:
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
        gflops = 2.0 * int_result / (total * 1e6)
        print(f"MH-SpGEMM runtime is {total:.3f}ms, Gflops is {gflops:.2f}")
        if args.out:
            write_mtx(args.out, C)
    except api.MhbError as e:
        print(f"MH-SpGEMM failed!!! {e}")  # the reference prints and records Gflops = 0 (src/main.cu:141-145)
        if not args.write:
            return 1
    if args.write:
        try:
            os.makedirs(args.write, exist_ok=True)
            with open(os.path.join(args.write, "Gflops_MH-SpGEMM.csv"), "a") as f:
                f.write(f"{gflops:.2f}\n")
        except OSError:
            print("Unable to open Gflops_MH-SpGEMM.csv")
            return 1
    print("--------------------------SpGEMM   End!!!--------------------------")
    return 0


def run_list(args) -> int:
    """process.sh:1-40 -- every matrix of a list file, in order."""
    try:
        names = [ln.strip() for ln in open(args.list) if ln.strip()]
    except OSError:
        print(f"Error: Matrix list file {args.list} not found.")
        return 1
    print(f"Total matrices to process: {len(names)}")
    for count, name in enumerate(names, 1):
        path = os.path.join(args.root, name, name + ".mtx")
        if not os.path.isfile(path):
            print(f"Warning: File not found: {path}")
            continue
        print(f"[{count}/{len(names)}] Processing: {path}")
        argv = [path] + (["--aat"] if args.aat else []) + ["--iters", str(args.iters)]
        if args.write:
            argv += ["--write", args.write]
        if main(argv) != 0:
            print(f"Error: Failed to process {path}")
            return 1
    print("All listed matrices processed successfully.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
