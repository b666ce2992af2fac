"""Turn ncu captures brought back in gpurun_out/ into the committed summaries of profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<name>.md
  python scripts/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  > profiles/<name>.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__t_set_accesses_pipe_lsu_mem_shared_op_atom.sum", "smsp__inst_executed_op_global_ld.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg.setdefault((r[ki].split("(")[0][:70], r[gi], r[bi]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | grid | block | launches | avg us | total us | share |\n|---|---|---|---|---|---|---|")
    for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {g} | {b} | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |")
    print(f"\ntotal device time of the listed launches: {tot:.1f} us "
          "(ncu serialises launches and runs them cold-cache: compare shares, not absolutes)")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"### `{name[:100]}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {vals[hdr.index(k)]} | {units[hdr.index(k)]} |")
        st = [(float(vals[i]), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")
              and "not_issued" not in h and vals[i]]
        if st:
            print("\nwarp stall reasons (cycles per issued instruction, top 6):\n")
            for v, h in sorted(st, reverse=True)[:6]:
                print(f"* {h.split('issue_stalled_')[1].split('_per_')[0]}: {v:.2f}")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
