"""BASELINE configs[3]: C = A*A (fp64) on a synthetic analog of each 16matrix.txt shape --
ours vs the reference kernels (oracle/_ref) vs cuSPARSE SpGEMM on the same B200, structure
checked against the reference's output.  Writes JSON lines and a markdown table.

    python scripts/suite.py gpurun_out/suite [name ...]
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mh_spgemm_b200  # noqa: E402
from mh_spgemm_b200 import generators as G  # noqa: E402
from mh_spgemm_b200.csr import CSR  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def child(kind, path):
    from oracle import Reference
    z = np.load(path)
    A = CSR(int(z["M"]), int(z["N"]), z["ptr"], z["col"], z["val"])
    R = Reference()
    if kind == "ref":
        r = R.spgemm(A, A, reps=9, warmup=2, e2e_reps=0)
    else:
        r = R.cusparse(A, A, reps=5, warmup=1)
    print("CHILD", json.dumps(dict(nnz=r["nnz"], ms=r["ms_device_min"], ms_median=r["ms_device"], sha_ptr=sha(r["ptr"]), sha_col=sha(r["col"]),
                                   sum=float(r["val"].sum()))))


def run_child(kind, path, timeout=900):
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", kind, path], capture_output=True,
                           text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return dict(error="timeout")
    for ln in p.stdout.splitlines():
        if ln.startswith("CHILD"):
            return json.loads(ln[6:])
    err = [x for x in (p.stdout + p.stderr).splitlines() if x.strip() and not x.startswith("C.nnz")]
    return dict(error=(err[-1][:160] if err else f"exit {p.returncode}"))


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3])
    outdir = sys.argv[1]
    os.makedirs(outdir, exist_ok=True)
    names = sys.argv[2:] or list(G.SUITE)
    from mh_spgemm_b200 import api
    tool = api.Tool(0)
    rows = []
    for name in names:
        t0 = time.time()
        A = G.suite(name)
        tgen = time.time() - t0
        ip = int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
        path = f"/dev/shm/suite_{name}.npz"
        np.savez(path, M=A.M, N=A.N, ptr=A.ptr, col=A.col, val=A.val)
        rec = dict(name=name, rows=A.M, nnz=A.nnz, intprod=ip, gen_s=round(tgen, 1))
        try:
            dAp, dAc, dAv = api.DeviceArray(A.ptr), api.DeviceArray(A.col), api.DeviceArray(A.val)
            ts = []
            for _ in range(5):
                dCp, nnzC = tool.symbolic(A.M, A.N, A.N, dAp, dAc, dAp, dAc)
                dCc, dCv = tool.numeric(dAv, dAv, nnzC)
                ts.append(tool.timing.total)
                if _ < 4:
                    dCc.free(), dCv.free(), dCp.free()
            ms = float(np.median(ts[1:]))
            cp, cc, cv = dCp.numpy(), dCc.numpy()[:nnzC], dCv.numpy()[:nnzC]
            rec.update(nnzC=nnzC, ours_ms=round(ms, 3), ours_gflops=round(2 * ip / ms / 1e6, 1),
                       stage=tool.timing.as_dict(), sha_ptr=sha(cp), sha_col=sha(cc), sum=float(cv.sum()))
            for d in (dCp, dCc, dCv, dAp, dAc, dAv):
                d.free()
        except Exception as e:  # noqa: BLE001
            rec["ours_error"] = str(e)[:200]
        for kind in (() if os.environ.get("SUITE_OURS_ONLY") else ("ref", "cusparse")):
            r = run_child(kind, path)
            if "error" in r:
                rec[kind + "_error"] = r["error"]
            else:
                rec[kind + "_ms"] = round(r["ms"], 3)  # best repetition (their host paths are noisy)
                rec[kind + "_ms_median"] = round(r["ms_median"], 3)
                rec[kind + "_gflops"] = round(2 * ip / r["ms"] / 1e6, 1)
                same = r["sha_ptr"] == rec.get("sha_ptr") and r["sha_col"] == rec.get("sha_col")
                rec[kind + "_structure_equal"] = bool(same)
                rec[kind + "_sum_rel"] = abs(r["sum"] - rec.get("sum", 0)) / max(abs(r["sum"]), 1e-300)
        os.remove(path)
        rows.append(rec)
        print(json.dumps({k: v for k, v in rec.items() if k != "stage"}), {k: round(v, 3) for k, v in rec.get("stage", {}).items()}, flush=True)
        with open(os.path.join(outdir, "suite.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    with open(os.path.join(outdir, "suite.md"), "w") as f:
        f.write("| analog of | rows | nnz | products | nnz(C) | ours ms | ours GFLOPS | reference ms | reference GFLOPS | "
                "cuSPARSE ms | cuSPARSE GFLOPS | speed-up vs ref | structure == ref |\n|" + "---|" * 13 + "\n")
        for r in rows:
            ref = r.get("ref_ms")
            f.write(f"| {r['name']} | {r['rows']} | {r['nnz']} | {r['intprod']} | {r.get('nnzC', '-')} | "
                    f"{r.get('ours_ms', r.get('ours_error', '-'))} | {r.get('ours_gflops', '-')} | "
                    f"{ref if ref else r.get('ref_error', '-')} | {r.get('ref_gflops', '-')} | "
                    f"{r.get('cusparse_ms', r.get('cusparse_error', '-'))} | {r.get('cusparse_gflops', '-')} | "
                    f"{round(ref / r['ours_ms'], 2) if ref and 'ours_ms' in r else '-'} | "
                    f"{r.get('ref_structure_equal', '-')} |\n")


if __name__ == "__main__":
    main()
