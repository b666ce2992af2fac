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
    """SM clock and throttle reasons sampled through NVML every ~2 ms while the timed region
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
            time.sleep(0.002)

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
def run_reference(args, rank):
    """--impl reference: the reference's own kernels (oracle/_ref) on the N=1 workload."""
    if rank != 0:
        return
    A, cfg = make_workload(args.workload)
    intprod = int(np.diff(A.ptr).astype(np.int64)[A.col].sum())
    base = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(cfg, parallelism="single GPU (the reference has no multi-GPU path)")}
    from oracle import Reference
    out = None
    if Reference.available():
        # separate process: a fault inside the reference must not take the bench down
        code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import mh_spgemm_b200\n"
                "import bench; from oracle import Reference\n"
                "A, _ = bench.make_workload(%r)\n"
                "R = Reference().spgemm(A, A, reps=%d, warmup=%d, e2e_reps=%d)\n"
                "print('REFJSON', json.dumps(dict(nnz=R['nnz'], ms_device=R['ms_device'], ms_e2e=R['ms_e2e'], ms_min=R['ms_device_min'])))\n"
                % (ROOT, args.workload, args.steps, args.warmup, max(2, min(args.steps, 5))))
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
                    e2e={"value": round(e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": round(out["ms_e2e"], 3)},
                    cpu_baseline={"value": round(v, 3), "unit": UNIT, "cores": 1, "kind": "reference",
                                  "sample": "whole workload; reference CUDA kernels rebuilt for sm_100 "
                                            "(oracle/_ref), MH_spgemm end to end incl. its mask build and "
                                            "per-call allocations, std::chrono, median"},
                    gpu_launches=0)
    else:
        cb = cpu_baseline(A, A, intprod)
        base.update(value=cb["value"], ms_per_step=cb["ms_per_step"], cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0)
    print(json.dumps(base), flush=True)


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="F", choices=["F", "P", "R", "G"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="range", choices=["range", "broadcast"],
                    help="N>1: 'range' = B row-sharded like A, each rank gathers the B rows its block "
                         "references (halo for FEM, all-gather for graphs); 'broadcast' = B on rank 0, "
                         "ncclBroadcast every step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from mh_spgemm_b200 import api
    from mh_spgemm_b200.distributed import (RangeExchange, ShardedSpGEMM, column_range, pack_b, partition_rows,
                                            row_work)
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
    A, cfg = make_workload(args.workload, scale=world)
    B = A
    work = row_work(A, B)
    intprod = int(work.sum())
    bounds = partition_rows(work, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    Ablk = A.rows(r0, r1)
    tool = api.Tool(local)
    stream = torch.cuda.current_stream()
    tool.set_stream(stream.cuda_stream)
    dt = torch.float64
    a_dev = (Ablk.M, torch.from_numpy(Ablk.ptr).to(dev), torch.from_numpy(Ablk.col).to(dev),
             torch.from_numpy(Ablk.val).to(dev))
    sh = ShardedSpGEMM(tool, rank, world, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    use_range = world > 1 and args.exchange == "range"
    if use_range:
        kr = [column_range(A.rows(int(bounds[r]), int(bounds[r + 1]))) for r in range(world)]
        plan = RangeExchange(rank, world, bounds, kr, B.ptr, dt, dev)
        a_shift = (a_dev[0], a_dev[1], a_dev[2] - plan.k0, a_dev[3])
        exch_bytes = plan.bytes_received
        # B = A is sharded like A: this rank's shard of B is its own block of A.  It is kept
        # inside the gathered image (own_views), so a step only receives the halo pieces.
        own_col, own_val = plan.own_views()
        if own_col.numel() == a_dev[2].numel():
            own_col.copy_(a_dev[2])
            own_val.copy_(a_dev[3])
        else:  # the block does not reference all of its own rows of B: keep the shard separate
            own_col, own_val = a_dev[2], a_dev[3]
        nnz_box = {}

        def one_step():
            out_ = sh.step_range(a_shift, plan, own_col, own_val, B.N, dt, offsets_on_host=False)
            nnz_box["total"] = out_[4]
            return out_
    else:
        packed, _ = pack_b(B)
        Bbuf = packed.to(dev) if rank == 0 else torch.empty_like(packed, device=dev)
        exch_bytes = 0 if world == 1 else packed.numel()

        def one_step():
            return sh.step(a_dev, Bbuf, B.M, B.N, B.nnz, dt, src=0)

    if strong:
        # nnz(C) of a rank can exceed int32: cut the local rows into slices of < 2^31 products,
        # one SpGEMM per slice (local int32 row_ptr, int64 offsets), C slices are not retained
        from mh_spgemm_b200.distributed import b_views, exchange_B, slice_offsets, slice_rows_fast
        slices = slice_rows_fast(work, r0, r1, cap=(1 << 31) - 1)
        slices = [(a - r0, b - r0) for a, b in slices]

        def one_step():  # noqa: F811
            if use_range:
                bp, bc, bv = plan.run(own_col, own_val)
                bc = bc[:plan.nnz_local]
                ac, K_loc = a_shift[2], plan.K_local
            else:
                exchange_B(Bbuf, world, 0)
                bp, bc, bv = b_views(Bbuf, B.M, B.nnz, dt)
                ac, K_loc = a_dev[2], B.M
            total_local = 0
            for s0, s1 in slices:
                cp, nnz = tool.symbolic(s1 - s0, K_loc, B.N, a_dev[1][s0:s1 + 1], ac, bp, bc)
                ccol = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
                cval = torch.empty(max(nnz, 1), dtype=dt, device=dev)
                tool.numeric_into(a_dev[3], bv, ccol, cval)
                total_local += nnz
                last = (cp, ccol[:nnz], cval[:nnz])
            off, total = slice_offsets(total_local, rank, world, dev)
            return last[0], last[1], last[2], off, total

    # the very first call is timed too: cold workspace (allocations), cold caches, module load
    cold0, cold1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cold0.record(stream)
    out = one_step()
    cold1.record(stream)
    torch.cuda.synchronize()
    cold_ms = cold0.elapsed_time(cold1)
    for _ in range(max(args.warmup - 1, 0)):
        out = one_step()
    nnzC_total = out[4].offsets()[1] if hasattr(out[4], "offsets") else int(out[4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    num_ms, launches = [], 0
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict L2 between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        one_step()
        ev[k][1].record(stream)
        num_ms.append(tool.timing.Numeric)
        launches += tool.stats["gpu_launches"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)  # max over ranks, per step
    step_ms = step_ms.cpu().numpy()
    ms = float(step_ms.mean())
    timing = tool.timing.as_dict()
    stats = tool.stats

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H inside) ----
    if strong:
        # the per-rank product can exceed the int32 contract of one host call; the device path
        # above is the measurement for this workload
        e2e_mean, h2d, d2h = None, 0, 0
    else:
        PA = tool.pin(Ablk)
        PB = PA if world == 1 else tool.pin(B)
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
        h2d = 4 * (Ablk.M + 1) + 12 * Ablk.nnz + (0 if world == 1 else 4 * (B.M + 1) + 12 * B.nnz)
        d2h = 4 * (Ablk.M + 1) + 12 * Ch.nnz

    if rank == 0:
        ba = bytes_alg(A, B, nnzC_total)
        kern_ms = float(np.mean(num_ms))
        # per-rank share of the compulsory traffic for the kernel's roofline (rank 0's slice)
        ba_rank = bytes_alg(Ablk, B, int(out[1].numel()))
        achieved = ba_rank / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(2.0 * intprod / ms / 1e6, 3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(cfg, intprod=intprod, nnzC=nnzC_total, l2="flushed between timed steps (256 MiB write)",
                           parallelism=("single GPU" if world == 1 else
                                        f"A row-sharded x{world} by product count; " +
                                        ("B row-sharded like A, referenced row range gathered by NCCL send/recv "
                                         "each step" if use_range else "B NCCL-broadcast from rank 0 each step")),
                           exchange_bytes_received_rank0=exch_bytes),
            "roofline": {"bound": "hbm", "kernel": "numeric (k_num_compact_rowtwins<double>)" if args.workload == "F"
                         else "numeric (all bins)", "achieved": round(achieved, 2), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4),
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload from
                         # one `ncu --set full` capture (profiles/r1h_numcompact_final.md); other workloads: null
                         "traffic": 243513856 if (args.workload == "F" and world == 1) else None,
                         "bound_on_chip": "LSU data pipe 75 % (shared-memory accumulator traffic), issue 40 %, "
                                          "14 warps/SM (profiles/r1h_numcompact_final.md)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ba_rank,
                         "kernel_ms": round(kern_ms, 4),
                         "pipeline_frac": round(ba / (ms * 1e-3) / 1e9 / peak / world, 4)},
            "e2e": (None if e2e_mean is None else
                    {"value": round(2.0 * intprod / e2e_mean / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_mean, 3)}),
            "gpu_launches": launches, "clocks": clocks,
            "stage_ms": {k: round(v, 4) for k, v in timing.items()},
            # the reference's own total leaves the mask build out (src/Timing.cpp:39-42)
            "ms_per_step_reference_convention": round(ms - timing.get("Form_mask_matrix_B", 0.0), 4),
            "cold_first_call_ms": round(cold_ms, 3),
            "bins": {"sym": {k: v for k, v in stats["sym_bins"].items() if v},
                     "num": {k: v for k, v in stats["num_bins"].items() if v}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(A, B, intprod)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
