#!/usr/bin/env python
"""bench.py -- SpGEMM GFLOPS (2 x intermediate products / s) and HBM-roofline fraction for
C = A*A, fp64, on N B200 GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload F|P|R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one complete SpGEMM (B mask build, binning, symbolic, nnz hand-off + allocation
of C, numeric) with A and B already resident in HBM.  N=1 runs BASELINE.json configs[1]
(the cant-like FEM matrix, 62,400 rows / 4.24 M nnz / 302.5 M products); N>1 runs the same
per-GPU work on a matrix N times longer (weak scaling): A is row-sharded by
intermediate-product count; inside every step B is exchanged with NCCL -- by default B is
row-sharded like A and each rank gathers the row range of B its block references
(--exchange range: a halo for FEM inputs, an all-gather for graphs), or B lives on rank 0
and is broadcast (--exchange broadcast).

`--impl reference` times the UNMODIFIED reference kernels rebuilt for sm_100
(oracle/_ref, through MH_spgemm) on the same input; if that library is absent it times
the host Gustavson oracle instead.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import mh_spgemm_b200  # noqa: E402,F401
from mh_spgemm_b200 import generators as G  # noqa: E402
from mh_spgemm_b200.csr import CSR  # noqa: E402

METRIC = "SpGEMM GFLOPS (2*intprod/s), C=A*A fp64"
UNIT = "GFLOPS (2*intprod/s)"


SPIN_CYCLES = 400_000  # ~0.2 ms at 1.965 GHz


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name: str, scale: int = 1) -> tuple[CSR, dict]:
    if name == "F":
        A = G.fem3d(8, 8, 325 * scale, 3, seed=1)
        desc = f"configs[1] cant-like FEM 27-pt 8x8x{325 * scale} x3dof, C=A*A"
    elif name == "FP":
        A = G.fem3d_perturbed(8, 8, 325 * scale, 3, drop=0.10, seed=1)
        desc = f"cant-like FEM 27-pt 8x8x{325 * scale} x3dof with 10 % of the off-diagonal entries dropped at random, C=A*A"
    elif name == "P":
        n = int(os.environ.get("MHB_POISSON_N", "256"))
        A = G.poisson2d(n)
        desc = f"configs[0] Poisson {n}x{n} 5-pt, C=A*A"
    elif name == "R":
        A = G.rmat()
        desc = "configs[2] webbase-like R-MAT scale 20, C=A*A"
    elif name == "G":
        # BASELINE configs[4] at a reduced scale: R-MAT, 2^S rows, 16 * 2^S draws, a=.45 b=c=.15
        # (S=24 is the full config; the default S=22 keeps host-side generation in seconds)
        S = int(os.environ.get("MHB_RMAT_SCALE", "22"))
        A = G.rmat(scale=S, n=1 << S, draws=16 << S, a=0.45, b=0.15, c=0.15, seed=5)
        desc = f"configs[4] R-MAT scale {S} ({1 << S} rows, {16 << S} draws, a=.45 b=c=.15), C=A*A, row-sharded"
    else:
        raise SystemExit(f"unknown workload {name}")
    return A, {"workload": desc, "rows": A.M, "nnzA": A.nnz}


def bytes_alg(A: CSR, B: CSR, nnzC: int, w: int = 8) -> int:
    """Compulsory CSR->CSR traffic: A, B read once, C written once."""
    return ((4 * (A.M + 1) + A.nnz * (4 + w)) + (4 * (B.M + 1) + B.nnz * (4 + w))
            + (4 * (A.M + 1) + nnzC * (4 + w)))


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~5 ms while the timed region
    runs (the nvidia-smi line of B200_PROFILING.md, in-process so that a region of a few
    milliseconds still gets samples)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.stop_flag, self.t = index, [], set(), False, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def mark(self):
        """Forget what was sampled so far (warm-up); keep sampling."""
        self.samples, self.reasons = [], set()

    def stop(self) -> dict:
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


def cpu_model() -> str:
    """CPU model of the box the baseline ran on (SURVEY.md 8d asks for it beside the thread count)."""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(A: CSR, B: CSR, intprod: int, budget_s: float = 12.0) -> dict:
    """Host Gustavson (the oracle port) on the box's cores: a reported baseline, not the target."""
    from oracle import Oracle
    o = Oracle()
    o.spgemm(A, B)  # warm (page faults of the per-thread accumulators)
    ts = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end and len(ts) < 20:
        t0 = time.perf_counter()
        o.spgemm(A, B)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return {"value": round(2.0 * intprod / t / 1e9, 3), "unit": UNIT, "cores": o.threads, "kind": "port",
            "sample": f"whole workload x{len(ts)} (median), host Gustavson symbolic+numeric, OpenMP",
            "ms_per_step": round(t * 1e3, 3), "cpu": cpu_model()}


# ---------------------------------------------------------------------------------------
L2_NOTE = ("flushed between timed steps (256 MiB written, then read back); a 0.2 ms spin kernel between the flush and "
           "the start event lets the host queue the step ahead of the device")

# numeric bin -> the kernel that serves it (csrc/mhb_capi.cu launch_numeric_bins)
NUM_KERNEL = {"WIN_COMPACT": "k_num_compact_rowtwins", "WIN_WARP": "k_num_win_group<32>", "WIN_G8": "k_num_win_group<8>",
              "WIN_BLOCK_S": "k_num_win_block", "WIN_BLOCK_L": "k_num_win_block", "H_G8": "k_num_hash_group<8>",
              "H_WARP_XS": "k_num_hash_list<wrows>", "H_WARP_S": "k_num_hash_list<wrows>", "H_WARP_M": "k_num_hash_list",
              "H_WARP_L": "k_num_hash_list", "H_BLOCK_S": "k_num_hash_list", "H_BLOCK_M": "k_num_hash_list", "H_BLOCK_XS": "k_num_hash_list",
              "H_BLOCK_L": "k_num_hash_block",
              "H_GLOBAL": "k_num_hash_block(pool)", "TINY": "k_num_tiny", "TINY_S": "k_num_tiny", "TINY_M": "k_num_tiny"}


def captured_traffic(workload: str, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from a
    committed `ncu --set full` capture (profiles/r2_traffic.json names the capture); never a
    literal in this file, and null for anything that has no capture."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(workload)
    except (OSError, ValueError):
        rec = None
    if not rec or world != 1:
        return None, None
    return int(rec["traffic"]), {k: rec[k] for k in ("kernel", "capture") if k in rec}


def run_reference(args, rank, out_stream=sys.stdout):
    """--impl reference: the reference's own kernels (oracle/_ref) on the SAME matrix our arm
    multiplies at this N (the reference is single-GPU: rank 0 runs the whole N x matrix)."""
    if rank != 0:
        return
    A, cfg = make_workload(args.workload, scale=max(1, args.gpus))
    intprod = int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
    base = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if args.workload == "G" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "parallelism": "single GPU (the reference has no multi-GPU path)"}
    from oracle import Reference
    out = None
    if Reference.available():
        # separate process: a fault inside the reference must not take the bench down
        code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import mh_spgemm_b200\n"
                "import bench; from oracle import Reference\n"
                "A, _ = bench.make_workload(%r, scale=%d)\n"
                "R = Reference().spgemm(A, A, reps=%d, warmup=%d, e2e_reps=%d)\n"
                "print('REFJSON', json.dumps(dict(nnz=R['nnz'], ms_device=R['ms_device'], ms_e2e=R['ms_e2e'], ms_min=R['ms_device_min'])))\n"
                % (ROOT, args.workload, max(1, args.gpus), args.steps, args.warmup, max(2, min(args.steps, 5))))
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500)
        for ln in p.stdout.splitlines():
            if ln.startswith("REFJSON"):
                out = json.loads(ln[8:])
        if out is None:
            base["reference_error"] = (p.stdout + p.stderr)[-400:]
    if out is not None:
        nnzC = out["nnz"]
        h2d = 2 * (4 * (A.M + 1) + 12 * A.nnz)  # the reference uploads A and its deep copy B
        d2h = 4 * (A.M + 1) + 12 * nnzC
        v = 2.0 * intprod / out["ms_device"] / 1e6
        e = 2.0 * intprod / out["ms_e2e"] / 1e6
        base.update(value=round(v, 3), ms_per_step=round(out["ms_device"], 4),
                    ms_per_step_best=round(out.get("ms_min", 0.0), 4),
                    config=dict(cfg, intprod=intprod, nnzC=nnzC),
                    l2="not flushed by the harness: every MH_spgemm call allocates its workspace and C afresh "
                       "(9 cudaMalloc, 12 stream creations), which is the reference's own stock path",
                    e2e={"value": round(e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": round(out["ms_e2e"], 3)},
                    cpu_baseline={"value": round(v, 3), "unit": UNIT, "cores": 1, "kind": "reference",
                                  "sample": "whole workload; reference CUDA kernels rebuilt for sm_100 "
                                            "(oracle/_ref), MH_spgemm end to end incl. its mask build and "
                                            "per-call allocations, std::chrono, median"},
                    gpu_launches=0)
    else:
        from oracle import Oracle
        nnzC = int(Oracle().symbolic(A, A)[-1])
        cb = cpu_baseline(A, A, intprod)
        base.update(value=cb["value"], ms_per_step=cb["ms_per_step"], cpu_baseline=cb,
                    config=dict(cfg, intprod=intprod, nnzC=nnzC),
                    e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0)
    print(json.dumps(base), file=out_stream, flush=True)


# ---------------------------------------------------------------------------------------
def check_slice_against_oracle(Ablk, B, cp, cc, cv, rtol=1e-12) -> dict:
    """This rank's slice of C against the host oracle on the same rows (outside the timed
    region): row_ptr and col_idx bit-exact, values within rtol."""
    from oracle import Oracle
    orc = Oracle()
    Cp, Cc, Cv = orc.spgemm(Ablk, B)
    ok_struct = bool(np.array_equal(cp.astype(np.int64), Cp) and np.array_equal(cc, Cc))
    bad = -1
    if ok_struct:
        bad, _ = orc.compare(Ablk.M, (cp, cc, cv), (Cp, Cc, Cv), rtol)
    return {"structure": ok_struct, "values_bad": int(bad), "nnz": int(Cp[-1])}


def check_slice_invariants(torch, Aslice, B, cp, cc, cv) -> dict:
    """Size-independent checks for slices too large for the host oracle (the large R-MAT):
    column ids ascending inside every row and in range; row sums equal A (B 1) (linearity)."""
    nnz = int(cc.numel())
    ok_sorted = True
    if nnz > 1:
        d = cc[1:] - cc[:-1]
        starts = cp[1:-1].long()
        starts = starts[(starts > 0) & (starts < nnz)]
        d[starts - 1] = 1
        ok_sorted = bool((d > 0).all().item()) and int(cc.min().item()) >= 0 and int(cc.max().item()) < B.N
        del d
    blen = np.diff(B.ptr)
    y = np.add.reduceat(np.append(B.val, 0.0), np.minimum(B.ptr[:-1], B.nnz))
    y[blen == 0] = 0.0
    t = Aslice.val * y[Aslice.col]
    want = np.add.reduceat(np.append(t, 0.0), np.minimum(Aslice.ptr[:-1], Aslice.nnz))
    want[np.diff(Aslice.ptr) == 0] = 0.0
    lens = (cp[1:] - cp[:-1]).long()
    got = torch.segment_reduce(cv[:nnz], "sum", lengths=lens, unsafe=True).cpu().numpy()
    got[np.diff(cp.cpu().numpy()) == 0] = 0.0
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))) if want.size else 0.0
    return {"sorted_in_range": ok_sorted, "rowsum_max_rel_err": err}


def perturbed_fem(tool, torch, dev, steps=10) -> dict:
    """The headline shape without its regularity, next to the headline number: the cant-like
    matrix with 10 % of its off-diagonal entries dropped at random (the dof rows of a node stop
    being exact twins).  Same call, same flush and event timing as the main loop, fewer steps;
    structure and values checked against the host oracle."""
    from oracle import Oracle
    A, cfg = make_workload("FP")
    ip = int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
    a_ptr, a_col, a_val = (torch.from_numpy(x).to(dev) for x in (A.ptr, A.col, A.val))
    cp = torch.empty(A.M + 1, dtype=torch.int32, device=dev)
    orc = Oracle()
    Cp, Cc, Cv = orc.spgemm(A, A)
    cc = torch.empty(int(Cp[-1]), dtype=torch.int32, device=dev)
    cv = torch.empty(int(Cp[-1]), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    ms = []
    for k in range(3 + steps):
        flush.fill_(k & 0xFF)
        torch.sum(flush.view(torch.int64), dim=0, keepdim=True, out=sink)
        torch.cuda._sleep(SPIN_CYCLES)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tool.spgemm_into_begin(A.M, A.N, A.N, a_ptr, a_col, a_val, a_ptr, a_col, a_val, cp, cc, cv)
        e1.record(stream)
        nnz = tool.spgemm_into_end()
        torch.cuda.synchronize()
        if k >= 3:
            ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms))
    st = tool.stats
    ok = (nnz == int(Cp[-1]) and np.array_equal(cp.cpu().numpy().astype(np.int64), Cp)
          and np.array_equal(cc.cpu().numpy(), Cc))
    bad, _ = orc.compare(A.M, (Cp, Cc, cv.cpu().numpy()), (Cp, Cc, Cv), 1e-12) if ok else (-1, None)
    return {"workload": cfg["workload"], "rows": A.M, "nnzA": A.nnz, "intprod": ip, "nnzC": int(nnz),
            "ms_per_step": round(t, 4), "gflops": round(2.0 * ip / t / 1e6, 1), "numeric_ms": round(tool.timing.Numeric, 4),
            "parity": "bit-exact structure, values 1e-12" if ok and bad == 0 else "FAILED",
            "num_bins": {k: v for k, v in st["num_bins"].items() if v}}


def suite_breadth(tool, budget_s=150.0) -> dict:
    """BASELINE configs[3] in one number: C = A*A on the twelve small suite analogs, ours
    (device time, mask build included, median of 3) vs the reference's kernels on the same box
    (oracle/_ref in a subprocess, best of 5), structure compared by SHA-256.  Outside the timed
    region; stops when the time budget is used up."""
    import hashlib
    from mh_spgemm_b200 import api
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    from oracle import Reference
    if not Reference.available():
        return {"unavailable": "oracle/_ref not built"}
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, np.int32).tobytes()).hexdigest()  # noqa: E731
    t_end = time.time() + budget_s
    rows, ratios = {}, []
    for name in cases.SUITE12:
        if time.time() > t_end:
            break
        A = G.suite(name)
        ip = int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
        dAp, dAc, dAv = api.DeviceArray(A.ptr), api.DeviceArray(A.col), api.DeviceArray(A.val)
        ts = []
        for it in range(4):
            dCp, nnzC = tool.symbolic(A.M, A.N, A.N, dAp, dAc, dAp, dAc)
            dCc, dCv = tool.numeric(dAv, dAv, nnzC)
            ts.append(tool.timing.total)
            if it < 3:
                dCc.free(), dCv.free(), dCp.free()
        ours_ms = float(np.median(ts[1:]))
        mine = (sha(dCp.numpy()), sha(dCc.numpy()[:nnzC]))
        for d in (dCp, dCc, dCv, dAp, dAc, dAv):
            d.free()
        path = f"/dev/shm/mhb_suite_{os.getpid()}.npz"
        np.savez(path, M=A.M, N=A.N, ptr=A.ptr, col=A.col, val=A.val)
        code = ("import sys, json, hashlib, numpy as np; sys.path.insert(0, %r); import mh_spgemm_b200\n"
                "from mh_spgemm_b200.csr import CSR; from oracle import Reference\n"
                "z = np.load(%r); A = CSR(int(z['M']), int(z['N']), z['ptr'], z['col'], z['val'])\n"
                "R = Reference().spgemm(A, A, reps=5, warmup=1, e2e_reps=0)\n"
                "h = lambda a: hashlib.sha256(np.ascontiguousarray(a, np.int32).tobytes()).hexdigest()\n"
                "print('REFJSON', json.dumps(dict(ms=R['ms_device_min'], sp=h(R['ptr']), sc=h(R['col']))))\n" % (ROOT, path))
        rec = {"ours_ms": round(ours_ms, 3), "ours_gflops": round(2 * ip / ours_ms / 1e6, 1)}
        try:
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
            r = [json.loads(ln[8:]) for ln in p.stdout.splitlines() if ln.startswith("REFJSON")]
            if r:
                rec.update(ref_ms=round(r[0]["ms"], 3), speedup=round(r[0]["ms"] / ours_ms, 2),
                           structure_equal=bool((r[0]["sp"], r[0]["sc"]) == mine))
                ratios.append(r[0]["ms"] / ours_ms)
            else:
                rec["ref_error"] = (p.stdout + p.stderr)[-160:]
        except subprocess.TimeoutExpired:
            rec["ref_error"] = "timeout"
        finally:
            if os.path.exists(path):
                os.remove(path)
        rows[name] = rec
    geo = float(np.exp(np.mean(np.log(ratios)))) if ratios else None
    return {"suite_geomean_vs_reference": None if geo is None else round(geo, 3), "matrices": len(ratios),
            "all_structure_equal": all(r.get("structure_equal", False) for r in rows.values() if "ref_ms" in r),
            "per_matrix": rows}


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="F", choices=["F", "FP", "P", "R", "G"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-suite", action="store_true", help="skip the 12-matrix breadth check (N=1, workload F only)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="mhb_set_option on the handle before the run (A/B measurements), e.g. --opt compact_rows=0")
    ap.add_argument("--no-perturbed", action="store_true", help="skip the perturbed-FEM side measurement (N=1, workload F)")
    ap.add_argument("--contract", default="fused", choices=["fused", "two-phase"],
                    help="'fused' = mhb_spgemm_into_* (C arrays kept by the caller, one host synchronisation per "
                         "SpGEMM); 'two-phase' = mhb_symbolic, then mhb_numeric_* (the reference's hand-off as two calls)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "broadcast", "sendrecv"],
                    help="N>1: 'peer' = B row-sharded like A, every rank pulls the B rows its block references out of "
                         "the owners' CUDA-IPC windows (mhb_shard_*, one-sided over NVLink); 'broadcast' = B on "
                         "rank 0, ncclBroadcast from C++ every step (the north-star wording); 'sendrecv' = the "
                         "round-1 path, grouped NCCL send/recv driven by torch.distributed")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # ONE JSON line on stdout: native libraries (NCCL prints its version banner on fd 1 when
    # NCCL_DEBUG is set in the environment) are pointed at stderr for the rest of the run
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, real_stdout)

    import torch
    import torch.distributed as dist
    from mh_spgemm_b200 import api
    from mh_spgemm_b200.distributed import (RangeExchange, Shard, ShardedSpGEMM, SliceSizes, b_views, column_range,
                                            pack_b, partition_rows, row_cost, row_work, slice_rows_fast,
                                            snap_to_pattern_change)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = load_peaks()

    # ---- workload: identical seeded matrix on every rank, rows sharded by product count ----
    strong = args.workload == "G"  # fixed matrix split over the ranks (strong scaling)
    rmat_scale = int(os.environ.get("MHB_RMAT_SCALE", "22"))
    if strong and os.environ.get("MHB_RMAT_GEN", "device" if rmat_scale >= 23 else "host") == "device":
        # full-size configs[4]: every rank draws the SAME matrix on its own GPU (seeded Philox stream) --
        # the host generator needs minutes per process at scale 24 while the whole box waits
        S = rmat_scale
        A = G.rmat_device(S, 1 << S, 16 << S, 0.45, 0.15, 0.15, seed=5, device=dev)
        torch.cuda.empty_cache()
        cfg = {"workload": f"configs[4] R-MAT scale {S} ({1 << S} rows, {16 << S} draws, a=.45 b=c=.15; drawn on the device, "
                           f"torch Philox seed 5), C=A*A, row-sharded", "rows": A.M, "nnzA": A.nnz}
        if world > 1:  # the ranks must hold the same matrix
            chk = torch.tensor([float(A.nnz), float(A.col[::4097].astype(np.int64).sum()), float(A.val[::4099].sum())],
                               dtype=torch.float64, device=dev)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                raise SystemExit("device-side R-MAT differs between the ranks")
    elif strong:
        # the large R-MAT takes the host a minute to generate: the first process to need it writes
        # it to shared memory, every other rank (and later runs on the same box) reads it back --
        # the same seeded input either way
        path = f"/dev/shm/mhb_bench_G_s{os.environ.get('MHB_RMAT_SCALE', '22')}.npz"
        if rank == 0 and not os.path.exists(path):
            A, cfg = make_workload(args.workload, scale=1)
            np.savez(path + ".tmp.npz", M=A.M, N=A.N, ptr=A.ptr, col=A.col, val=A.val, desc=cfg["workload"])
            os.replace(path + ".tmp.npz", path)
        if world > 1:
            dist.barrier()
        z = np.load(path)
        A = CSR(int(z["M"]), int(z["N"]), z["ptr"], z["col"], z["val"])
        cfg = {"workload": str(z["desc"]), "rows": A.M, "nnzA": A.nnz}
    else:
        A, cfg = make_workload(args.workload, scale=world)
    B = A
    work = row_work(A, B)
    intprod = int(work.sum())
    # row blocks balanced by a per-row cost (products, inflated for long rows), never cutting a
    # run of twin rows
    bounds = snap_to_pattern_change(A, partition_rows(row_cost(work), world))
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    Ablk = A.rows(r0, r1)
    tool = api.Tool(local)
    # one explicit stream for everything: the flush, the events and the library's kernels.  (Until r2m
    # the handle was "bound" to torch's default stream, whose handle 0 the C ABI reads as "own
    # stream": the first kernels of a step then overlapped the tail of the L2 flush.)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    tool.set_stream(stream.cuda_stream)
    for kv in args.opt:
        key, _, val = kv.partition("=")
        tool.set_option(key, int(val))
    dt = torch.float64
    a_ptr, a_col, a_val = (torch.from_numpy(Ablk.ptr).to(dev), torch.from_numpy(Ablk.col).to(dev),
                           torch.from_numpy(Ablk.val).to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    flush_sink = torch.zeros(1, dtype=torch.int64, device=dev)
    # rows of this rank cut into slices of < 2^31 products (local int32 row_ptr per slice, int64 offsets)
    slices = [(a - r0, b - r0) for a, b in slice_rows_fast(work, r0, r1, cap=(1 << 31) - 1)] if r1 > r0 else [(0, 0)]
    mode = "single" if world == 1 else args.exchange
    sh = sizes = None
    exch_bytes = 0
    k0, k1 = 0, B.M

    c_buf = {}  # the caller's C arrays, kept across steps and regrown only when a slice needs more room

    def grow(buf, nnz):
        buf[1] = buf[2] = None
        buf[1] = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        buf[2] = torch.empty(max(nnz, 1), dtype=dt, device=dev)

    def multiply(into, begin=None, end=None):
        """One SpGEMM per slice into the caller's C arrays.  Queues the work and returns fin(), which
        waits for it and gives (last slice, the rank's nnz).
        C.ptr / C.col / C.val are caller-owned (the hand-off of src/main.cu:55-60): a caller that
        multiplies repeatedly keeps its buffers, so they are allocated on the first step (and
        whenever nnz outgrows them: the call reports MHB_ERR_CAPACITY with the size it needs) and
        reused afterwards; with one slice per rank nothing is shared between slices.
        With begin / end (the two halves of the fused call) the last slice is only QUEUED here, so
        that the caller's stop event lands right behind the last kernel instead of behind the
        host's wake-up from the call's final synchronisation."""
        total, last = 0, None
        for i, (s0, s1) in enumerate(slices):
            key = i if len(slices) == 1 else 0  # many slices (C beyond int32): one set of buffers, reused in turn
            if key not in c_buf or c_buf[key][0].numel() < s1 - s0 + 1:
                c_buf[key] = [torch.empty(s1 - s0 + 1, dtype=torch.int32, device=dev), None, None]
            buf = c_buf[key]
            cp = buf[0][:s1 - s0 + 1]
            if begin is not None and i == len(slices) - 1 and buf[1] is not None:
                begin(s0, s1, cp, buf[1], buf[2])

                def fin(total=total, s0=s0, s1=s1, cp=cp, buf=buf):
                    try:
                        nnz = end()
                    except api.MhbError as e:
                        if e.code != api.ERR_CAPACITY:
                            raise
                        grow(buf, e.nnzC)
                        nnz = into(s0, s1, cp, buf[1], buf[2])
                    return (s0, s1, cp, buf[1][:nnz], buf[2][:nnz]), total + nnz
                return fin
            try:
                nnz = into(s0, s1, cp, buf[1], buf[2])
            except api.MhbError as e:
                if e.code != api.ERR_CAPACITY:
                    raise
                grow(buf, e.nnzC)
                nnz = into(s0, s1, cp, buf[1], buf[2])
            total += nnz
            last = (s0, s1, cp, buf[1][:nnz], buf[2][:nnz])
        return lambda: (last, total)

    def two_phase(sym, num):
        """The reference's contract as two calls (symbolic -> caller sizes C -> numeric), behind the
        same signature as the fused call: --contract two-phase."""
        def into(s0, s1, cp, cc, cv):
            nnz = sym(s0, s1, cp)
            if cc is None or cc.numel() < nnz:
                e = api.MhbError(api.ERR_CAPACITY, "grow C")
                e.nnzC = nnz
                raise e
            num(cc, cv)
            return nnz
        return into

    fused = args.contract == "fused"
    # step_queue() puts one step on the stream and returns fin(), which waits for it and gives
    # (last slice, the rank's nnz).  With the fused contract the stop event of a timed step is
    # recorded between the two, i.e. right behind the last kernel (device_bracket = True).
    device_bracket = fused and mode in ("single", "peer")
    if mode == "single":
        if fused:
            into = lambda s0, s1, cp, cc, cv: tool.spgemm_into(s1 - s0, B.M, B.N, a_ptr[s0:], a_col, a_val, a_ptr, a_col,
                                                               a_val, cp, cc, cv)
            begin = lambda s0, s1, cp, cc, cv: tool.spgemm_into_begin(s1 - s0, B.M, B.N, a_ptr[s0:], a_col, a_val, a_ptr,
                                                                      a_col, a_val, cp, cc, cv)
            end = tool.spgemm_into_end
        else:
            into = two_phase(lambda s0, s1, cp: tool_symbolic_into(tool, s1 - s0, B.M, B.N, a_ptr[s0:], a_col, a_ptr,
                                                                   a_col, cp),
                             lambda cc, cv: tool.numeric_into(a_val, a_val, cc, cv))
            begin = end = None

        def step_queue():  # B = A: one copy on the device, aliasing visible to the library
            return multiply(into, begin, end)
    elif mode == "peer":
        Bown = B.rows(r0, r1)  # B = A is sharded like A
        dBp = torch.from_numpy(Bown.ptr).to(dev)
        sh = Shard(tool, rank, world, B.M, B.N, np.float64, bounds).build(Ablk.M, a_ptr, a_col, dBp)
        col_own, val_own = sh.own_B()
        col_own.upload(Bown.col)
        val_own.upload(Bown.val)
        k0, k1, _, exch_bytes = sh.image()

        trace = []  # MHB_BENCH_TRACE=1: device time of exchange / multiply / size post per step
        if fused:
            into = lambda s0, s1, cp, cc, cv: sh.spgemm_into(s0, s1, a_val, cp, cc, cv)
            begin = lambda s0, s1, cp, cc, cv: sh.spgemm_into_begin(s0, s1, a_val, cp, cc, cv)
            end = sh.spgemm_into_end
        else:
            into = two_phase(lambda s0, s1, cp: sh.symbolic(s0, s1, cp), lambda cc, cv: sh.numeric_into(a_val, cc, cv))
            begin = end = None
        # one slice per rank: the slice size is posted to the peers from the device, behind the
        # SpGEMM's kernels and inside the timed region, and confirmed from the host afterwards
        device_post = fused and len(slices) == 1

        def step_queue():
            tr = os.environ.get("MHB_BENCH_TRACE") == "1"
            if tr:
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                e[0].record(stream)
            sh.exchange()
            if tr:
                e[1].record(stream)
            fin = multiply(into, begin, end)
            if tr:
                e[2].record(stream)
            queued = device_post and c_buf.get(0, [None, None])[1] is not None
            if queued:
                sh.post_size(-1)
            if tr:
                e[3].record(stream)
                trace.append(e)

            def finish():
                last, total = fin()
                if queued:
                    sh.repost_size(total)
                else:
                    sh.post_size(total)
                return last, total
            return finish
    elif mode == "broadcast":
        sh = Shard(tool, rank, world, B.M, B.N, np.float64, bounds)
        sh.init_nccl()
        sizes = SliceSizes(rank, world, dev)
        packed, _ = pack_b(B)
        Bbuf = packed.to(dev) if rank == 0 else torch.empty_like(packed, device=dev)
        exch_bytes = 0 if rank == 0 else packed.numel()
        bp, bc, bv = b_views(Bbuf, B.M, B.nnz, dt)
        if fused:
            into = lambda s0, s1, cp, cc, cv: tool.spgemm_into(s1 - s0, B.M, B.N, a_ptr[s0:], a_col, a_val, bp, bc, bv,
                                                               cp, cc, cv)
        else:
            into = two_phase(lambda s0, s1, cp: tool_symbolic_into(tool, s1 - s0, B.M, B.N, a_ptr[s0:], a_col, bp, bc, cp),
                             lambda cc, cv: tool.numeric_into(a_val, bv, cc, cv))

        def step_queue():
            sh.broadcast(Bbuf, Bbuf.numel(), 0)  # ncclBroadcast issued from C++ on the Tool's stream
            last, total = multiply(into)()
            sizes.gather(total)
            return lambda: (last, total)
    else:  # sendrecv: the round-1 exchange (torch.distributed grouped send/recv), kept as the fallback
        kr = [column_range(A.rows(int(bounds[r]), int(bounds[r + 1]))) for r in range(world)]
        plan = RangeExchange(rank, world, bounds, kr, B.ptr, dt, dev)
        a_shift = a_col - plan.k0
        k0, k1 = plan.k0, plan.k1
        exch_bytes = plan.bytes_received
        own_col, own_val = plan.own_views()
        if own_col.numel() == a_col.numel():
            own_col.copy_(a_col)
            own_val.copy_(a_val)
        else:
            own_col, own_val = a_col, a_val
        sizes = SliceSizes(rank, world, dev)

        def step_queue():
            bp, bc, bv = plan.run(own_col, own_val)
            last, total = multiply(two_phase(
                lambda s0, s1, cp: tool_symbolic_into(tool, s1 - s0, plan.K_local, B.N, a_ptr[s0:], a_shift, bp,
                                                      bc[:plan.nnz_local], cp),
                lambda cc, cv: tool.numeric_into(a_val, bv, cc, cv)))()
            sizes.gather(total)
            return lambda: (last, total)

    def one_step():
        return step_queue()()

    def totals():
        if mode == "single":
            return 0, nnz_local, [nnz_local]
        if mode == "peer":
            return sh.offsets()
        off, tot = sizes.offsets()
        return off, tot, None

    # NVML is initialised (and its first, slow queries made) BEFORE anything is timed: r2e showed
    # the first timed step of the OTHER rank taking 37 ms while rank 0 sat in nvmlInit
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # the very first call is timed too: cold workspace (allocations), cold caches, module load
    cold0, cold1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cold0.record(stream)
    last, nnz_local = one_step()
    cold1.record(stream)
    torch.cuda.synchronize()
    cold_ms = cold0.elapsed_time(cold1)
    for _ in range(max(args.warmup - 1, 0)):
        last, nnz_local = one_step()
    slice_off, nnzC_total, all_sizes = totals()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        last, nnz_local = one_step()  # one more untimed step AFTER the host barrier: the first step behind
        torch.cuda.synchronize()      # an NCCL barrier ran 20-50 % long on one rank or the other (r2i)
    if rank == 0:
        sampler.mark()  # samples from here on belong to the timed region
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    num_ms, launches = [], 0
    torch.cuda.synchronize()
    import gc
    gc.collect()
    gc.disable()  # no collector pause inside a 0.5 ms step
    for k in range(args.steps):
        # evict L2 between timed iterations (outside the event pair): write 256 MiB, then read it
        # back so that L2 is left full of CLEAN lines -- after a write-only flush the step also
        # pays for draining ~126 MB of dirty lines (r2e: 0.05 ms in front of the first kernel)
        flush.fill_(k & 0xFF)
        # the sum lands in flush_sink directly: `flush_sink.copy_(x.sum())` put a copy-engine
        # memcpy between the flush and the start event, and the first kernels behind it ran
        # 0.03-0.06 ms long (r2k: mem_alloc 0.034 / mask 0.098 ms at N=1 against 0.006 / 0.062 ms
        # behind the barrier kernel of the N>1 runs)
        torch.sum(flush.view(torch.int64), dim=0, keepdim=True, out=flush_sink)
        # ... and a ~0.2 ms spin kernel in front of the start event, so that the host has the step's
        # launches queued before the device reaches them (nvbench's blocking kernel, by time because
        # the call under test ends with its own synchronisation).  Without it the first kernels of a
        # step wait for the HOST, which comes out of the previous step's synchronisation and the
        # Python around it just in time (r2k/r2l: mem_alloc 0.035 ms for a 3 us kernel, mask build
        # 0.10 ms against 0.062 ms behind the peer barrier of the N>1 runs)
        torch.cuda._sleep(SPIN_CYCLES)
        if mode == "peer":
            sh.barrier()       # ranks leave the flush together: no start-time skew inside the event pair
        ev[k][0].record(stream)
        fin = step_queue()
        if device_bracket:
            ev[k][1].record(stream)  # behind the last kernel of the step; fin() waits for it
        last, nnz_local = fin()
        if not device_bracket:
            ev[k][1].record(stream)
        num_ms.append(tool.timing.Numeric)
        launches += tool.stats["gpu_launches"] * len(slices) + (3 if mode == "peer" else 0)
    torch.cuda.synchronize()
    gc.enable()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    own_ms = float(step_ms.mean().item())
    per_rank = None
    if world > 1:
        allsteps = [torch.zeros_like(step_ms) for _ in range(world)]
        dist.all_gather(allsteps, step_ms)
        tmg = tool.timing
        info = torch.tensor([float(Ablk.M), float(Ablk.nnz), float(work[r0:r1].sum()), float(nnz_local),
                             tmg.total, tmg.Numeric, tmg.Calculate_C_nnz, tmg.Form_mask_matrix_B],
                            dtype=torch.float64, device=dev)
        allinfo = [torch.zeros_like(info) for _ in range(world)]
        dist.all_gather(allinfo, info)
        per_rank = [{"mean_ms": round(float(t.mean().item()), 4), "median_ms": round(float(t.median().item()), 4),
                     "max_ms": round(float(t.max().item()), 4), "rows": int(i[0].item()), "nnzA": int(i[1].item()),
                     "products": int(i[2].item()), "nnzC": int(i[3].item()),
                     "stage_total_ms": round(float(i[4].item()), 4), "numeric_ms": round(float(i[5].item()), 4),
                     "symbolic_ms": round(float(i[6].item()), 4), "mask_ms": round(float(i[7].item()), 4),
                     "steps_ms": [round(float(x), 3) for x in t.tolist()]} for t, i in zip(allsteps, allinfo)]
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)  # max over ranks, per step
    step_ms = step_ms.cpu().numpy()
    ms = float(step_ms.mean())
    timing = tool.timing.as_dict()
    stats = tool.stats

    # ---- parity, outside the timed region, on every rank and at every N ----
    parity = None
    if not args.no_parity:
        s0, s1, cp, cc, cv = last
        if strong:
            inv = check_slice_invariants(torch, Ablk.rows(s0, s1), B, cp, cc, cv)
            okv = inv["sorted_in_range"] and inv["rowsum_max_rel_err"] < 1e-9
            mine = torch.tensor([1.0 if okv else 0.0, inv["rowsum_max_rel_err"]], dtype=torch.float64, device=dev)
        else:
            chk = check_slice_against_oracle(Ablk, B, cp.cpu().numpy(), cc.cpu().numpy(), cv.cpu().numpy())
            okv = chk["structure"] and chk["values_bad"] == 0 and chk["nnz"] == nnz_local
            mine = torch.tensor([1.0 if okv else 0.0, float(max(chk["values_bad"], 0))], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, mine)
        else:
            allv = [mine]
        oks = [bool(v[0].item() > 0.5) for v in allv]
        sizes_ok = True
        if not strong:  # slice sizes / offsets against the oracle's global row_ptr
            from oracle import Oracle
            gp = Oracle().symbolic(A, B) if rank == 0 else None
            if rank == 0:
                want = [int(gp[int(bounds[r + 1])] - gp[int(bounds[r])]) for r in range(world)]
                sizes_ok = (nnzC_total == int(gp[-1])) and (all_sizes is None or list(all_sizes) == want)
        if strong:
            parity = {"checked": "every rank's last slice: columns ascending and in range, row sums == A(B 1)",
                      "structure": "sorted + in range" if all(oks) else "FAILED",
                      "rowsum_max_rel_err": max(float(v[1].item()) for v in allv), "ranks_ok": oks,
                      "nnzC_total": nnzC_total}
        else:
            parity = {"checked": "every rank's slice vs the host oracle on the same rows; slice sizes vs the oracle's global row_ptr",
                      "structure": "bit-exact" if all(oks) and sizes_ok else "FAILED", "values_rtol": 1e-12,
                      "values_out_of_tol": int(sum(float(v[1].item()) for v in allv)), "ranks_ok": oks,
                      "slice_sizes_ok": bool(sizes_ok)}
        if not (all(oks) and sizes_ok):
            if rank == 0:
                print(json.dumps({"error": "parity check failed", "parity": parity}), file=real_stdout, flush=True)
            raise SystemExit(3)

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H inside) ----
    if strong:
        # the per-rank product can exceed the int32 contract of one host call; the device path
        # above is the measurement for this workload
        e2e_mean, h2d, d2h = None, 0, 0
    else:
        if world == 1:
            PA = PB = tool.pin(Ablk)
            h2d = 4 * (Ablk.M + 1) + 12 * Ablk.nnz
        else:  # a rank uploads its block of A and only the rows [k0, k1) of B that the block references
            if mode in ("single", "broadcast"):
                k0, k1 = column_range(Ablk)
            from mh_spgemm_b200.csr import CSR as _CSR
            PA = tool.pin(_CSR(Ablk.M, k1 - k0, Ablk.ptr, Ablk.col - k0, Ablk.val))
            Bimg = B.rows(k0, k1)
            PB = tool.pin(Bimg)
            h2d = 4 * (Ablk.M + 1) + 12 * Ablk.nnz + 4 * (Bimg.M + 1) + 12 * Bimg.nnz
        tool.set_stream(None)
        e2e_ms = []
        for k in range(args.warmup + args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            Ch = tool.spgemm_host(PA, PB, copy=False)
            t1 = time.perf_counter()
            if k >= args.warmup:
                e2e_ms.append((t1 - t0) * 1e3)
        e2e_t = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_mean = float(e2e_t.item())
        d2h = 4 * (Ablk.M + 1) + 12 * Ch.nnz
        tool.set_stream(stream.cuda_stream)

    if rank == 0:
        ba = bytes_alg(A, B, nnzC_total)
        kern_ms = float(np.mean(num_ms))
        # per-rank share of the compulsory traffic for the kernel's roofline (rank 0's last slice)
        # (a rank that cuts its block into slices: the A rows of the LAST slice, whose numeric time kern_ms is)
        A_last = Ablk if len(slices) == 1 else Ablk.rows(last[0], last[1])
        ba_rank = bytes_alg(A_last, B if mode in ("single", "broadcast") else B.rows(k0, k1), int(last[3].numel()))
        achieved = ba_rank / (kern_ms * 1e-3) / 1e9
        nb = {k: v for k, v in stats["num_bins"].items() if v and k != "EMPTY"}
        kernels = sorted({NUM_KERNEL.get(k, k) for k in nb})
        traffic, capture = captured_traffic(args.workload, world)
        par = {"single": "single GPU",
               "peer": f"A row-sharded x{world} by per-row cost (products, long rows weighted up); B row-sharded like A, every rank pulls the B rows its "
                       "block references out of the owners' CUDA-IPC windows (one-sided, NVLink), C ABI mhb_shard_*",
               "broadcast": f"A row-sharded x{world} by product count; B ncclBroadcast from rank 0 each step (C++), "
                            "slice sizes by NCCL all-gather",
               "sendrecv": f"A row-sharded x{world}; B row range by grouped NCCL send/recv (torch.distributed)"}[mode]
        line = {
            "metric": METRIC, "value": round(2.0 * intprod / ms / 1e6, 3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(cfg, intprod=intprod, nnzC=nnzC_total),
            "l2": L2_NOTE, "parallelism": par,
            "call": ("mhb_spgemm_into_f64: caller-owned C arrays kept across steps, one host synchronisation per SpGEMM"
                     if fused else "mhb_symbolic + mhb_numeric_f64: two calls, host reads nnz(C) in between"),
            "options": args.opt or None,
            "fused_calls": stats.get("fused_calls"), "speculative_misses": stats.get("speculative_misses"),
            "timed_region": ("start event | exchange, every kernel of the SpGEMM, slice-size post | stop event, all queued "
                             "before the host waits (mhb_spgemm_into_begin / _end)" if device_bracket else
                             "start event | the whole step including its host synchronisations | stop event"), "exchange_bytes_received_rank0": exch_bytes, "slices_rank0": len(slices),
            "roofline": {"bound": "hbm",
                         "kernel": "numeric: " + " + ".join(kernels) + ("" if len(kernels) == 1 else " (bins run concurrently)"),
                         "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "traffic_from_capture": capture,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ba_rank,
                         "kernel_ms": round(kern_ms, 4),
                         "pipeline_frac": round(ba / (ms * 1e-3) / 1e9 / peak / world, 4)},
            "e2e": (None if e2e_mean is None else
                    {"value": round(2.0 * intprod / e2e_mean / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_mean, 3)}),
            "parity": parity,
            "gpu_launches": launches, "clocks": clocks,
            "stage_ms": {k: round(v, 4) for k, v in timing.items()},
            "rank0_ms_per_step": round(own_ms, 4), "per_rank": per_rank,
            "rank0_phase_ms": (None if not (mode == "peer" and trace) else
                               {n: round(float(np.median([t[i].elapsed_time(t[i + 1]) for t in trace[-args.steps:]])), 4)
                                for i, n in enumerate(("exchange", "symbolic+alloc+numeric", "post_size"))}),
            # the reference's own total leaves the mask build out (src/Timing.cpp:39-42)
            "ms_per_step_reference_convention": round(ms - timing.get("Form_mask_matrix_B", 0.0), 4),
            "cold_first_call_ms": round(cold_ms, 3),
            "bins": {"sym": {k: v for k, v in stats["sym_bins"].items() if v},
                     "num": {k: v for k, v in stats["num_bins"].items() if v}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(A, B, intprod)
        if world == 1 and args.workload == "F" and not args.no_perturbed:
            line["perturbed_fem"] = perturbed_fem(tool, torch, dev)
        if world == 1 and args.workload == "F" and not args.no_suite:
            line["suite"] = suite_breadth(tool)
        print(json.dumps(line), file=real_stdout, flush=True)
    if sh is not None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sh.close()
    if world > 1:
        dist.destroy_process_group()


def tool_symbolic_into(tool, M, K, N, a_ptr, a_col, b_ptr, b_col, cp):
    """mhb_symbolic into a caller-owned row_ptr (the Tool wrapper allocates its own)."""
    import ctypes as C
    nnz = C.c_longlong()
    tool._keep = (a_ptr, a_col, b_ptr, b_col, cp)
    tool._chk(tool.L.mhb_symbolic(tool.h, M, K, N, a_col.numel(), a_ptr.data_ptr(), a_col.data_ptr(), b_col.numel(),
                                  b_ptr.data_ptr(), b_col.data_ptr(), cp.data_ptr(), C.byref(nnz)))
    return int(nnz.value)


if __name__ == "__main__":
    main()
