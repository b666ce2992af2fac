"""profiles/r2_scaling.md from the bench lines of the scaling runs.
  python scripts/r2_scaling_table.py <title> file.json [file.json ...]   (one bench line per file)"""
import json
import sys

rows = []
for f in sys.argv[1:]:
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            rows.append((f, d))
base = {}
print("| file | workload | exchange / contract | N | ms / step | GFLOPS | vs N=1 (weak: / N) | per-rank compute (stage total) ms | e2e ms | parity |")
print("|---|---|---|---|---|---|---|---|---|---|")
for f, d in rows:
    wl = d["config"]["workload"].split(",")[0][:40]
    key = (wl.split(" ")[0], d["scaling"])
    par = d.get("parallelism", "")
    ex = "peer" if "CUDA-IPC" in par else ("broadcast" if "ncclBroadcast" in par else ("sendrecv" if "send/recv" in par else "single"))
    ex += " / " + ("fused" if str(d.get("call", "")).startswith("mhb_spgemm_into") else "two-phase")
    n = d["n_gpus"]
    if n == 1:
        base[key] = d["value"]
    eff = ""
    if key in base:
        eff = f"{d['value'] / base[key] / (n if d['scaling'] == 'weak' else 1):.2f}" + ("" if d["scaling"] == "weak" else "x")
    e2e = (d.get("e2e") or {}).get("ms_per_step", "")
    p = d.get("parity") or {}
    print(f"| `{f.split('gpurun_out/')[-1]}` | {wl} | {ex} | {n} | {d['ms_per_step']} | {d['value']} | {eff} | "
          f"{d['stage_ms']['total']} | {e2e} | {p.get('structure', '')} |")
