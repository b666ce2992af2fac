"""Generate the golden fixtures by running the UNMODIFIED reference kernels
(oracle/_ref/libmhref.so, rebuilt for sm_100) on a B200:

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden [--suite]
    cp gpurun_out/golden/* tests/golden/

--suite records the twelve suite analogs of BASELINE configs[3] (ref_suite_<name>.json).

Each case runs in its own process so a fault inside the reference cannot poison the rest.
Small cases store the full CSR of C; large cases store nnz, SHA-256 of row_ptr and
col_idx and value checksums.  The reference's values are nondeterministic in the last
bits (atomic accumulation order); its structure is deterministic.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402


def checksums(ptr, col, val):
    w = (np.arange(val.size, dtype=np.int64) % 97 + 1).astype(np.float64)
    return dict(nnz=int(ptr[-1]), sha_ptr=hashlib.sha256(np.ascontiguousarray(ptr, np.int32).tobytes()).hexdigest(),
                sha_col=hashlib.sha256(np.ascontiguousarray(col, np.int32).tobytes()).hexdigest(),
                sum_val=float(val.sum()), sum_abs=float(np.abs(val).sum()), sum_weighted=float((val * w).sum()))


def run_case(name, outdir):
    from oracle import Reference
    table = {**cases.SMALL, **cases.LARGE, **cases.REFERENCE_FAULTS}
    suite = name.startswith("suite_")
    A, B = cases.SUITE[name[6:]]() if suite else table[name]()
    B = A if B is None else B
    R = Reference().spgemm(A, B, reps=1, warmup=0, e2e_reps=0, want_mask=True)
    tp, tc, tm = R["mask"]
    # the reference emits a row's tiles in hash-slot order: canonicalise per row
    order = np.lexsort((tc, np.repeat(np.arange(B.M), np.diff(tp))))
    meta = checksums(R["ptr"], R["col"], R["val"])
    meta.update(M=A.M, K=A.N, N=B.N, nnzA=A.nnz, nnzB=B.nnz, ntiles_B=int(tp[-1]),
                sha_tileptr=hashlib.sha256(tp.astype(np.int32).tobytes()).hexdigest(),
                sha_tilecol=hashlib.sha256(tc[order].astype(np.int32).tobytes()).hexdigest(),
                sha_tilemask=hashlib.sha256(tm[order].astype(np.uint32).tobytes()).hexdigest())
    if name in cases.SMALL:
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), ptr=R["ptr"], col=R["col"], val=R["val"],
                            tileptr=tp, tilecol=tc[order], tilemask=tm[order])
    with open(os.path.join(outdir, f"ref_{name}.json"), "w") as f:
        json.dump(meta, f, indent=1)


def main():
    outdir = sys.argv[1]
    os.makedirs(outdir, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[2] != "--suite":
        return run_case(sys.argv[2], outdir)
    summary = {}
    names = [*cases.SMALL, *cases.LARGE, *cases.REFERENCE_FAULTS]
    if len(sys.argv) > 2:  # --suite: only the twelve suite analogs of BASELINE configs[3]
        names = ["suite_" + n for n in cases.SUITE12]
    for name in names:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), outdir, name], capture_output=True, text=True,
                           timeout=900)
        ok = p.returncode == 0 and os.path.exists(os.path.join(outdir, f"ref_{name}.json"))
        summary[name] = "ok" if ok else ("reference faulted: " + (p.stdout + p.stderr).strip().splitlines()[0][:200]
                                         if (p.stdout + p.stderr).strip() else "reference faulted")
        print(name, summary[name], flush=True)
    with open(os.path.join(outdir, "ref_suite_summary.json" if len(sys.argv) > 2 else "ref_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
