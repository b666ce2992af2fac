"""profiles/r2_launches_<W>.md: the kernels of ONE bench step (ncu --metrics gpu__time_duration.sum
--clock-control none, serialised, cold caches) at the start and at the end of round 2.
  python scripts/r2_launch_tables.py F gpurun_out/r2h/launches_F.csv gpurun_out/r2o/launches_F.csv"""
import collections
import csv
import sys


def last_fused_step(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").replace("mhb::", "")})[r[mi]] = \
                float(r[vi].replace(",", ""))
    L = list(per.values())
    starts = [i for i, p in enumerate(L) if p["name"] == "k_zero16"] or [i for i, p in enumerate(L) if "k_mask" in p["name"]][:1]
    # steps of the device-resident loop carry no memcpy-only host call: take the last step that is
    # followed by another step or by the end, preferring one with the gate kernel (the fused call)
    steps = [L[a:b] for a, b in zip(starts, starts[1:] + [len(L)])]
    steps = [s for s in steps if not any(p["name"].startswith("at::") for p in s[:-3])] or steps
    fused = [s for s in steps if any(p["name"] == "k_fused_gate" for p in s)]
    st = (fused or steps)[-1]
    return [p for p in st if not (p["name"].startswith("at::") or "native::" in p["name"])]


def table(step):
    agg = collections.OrderedDict()
    for p in step:
        a = agg.setdefault(p["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += p.get("gpu__time_duration.sum", 0.0) / 1000.0
        a[2] += p.get("smsp__inst_executed.sum", 0.0) / 1e6
    return agg


w, before, after = sys.argv[1], sys.argv[2], sys.argv[3]
A, B = table(last_fused_step(before)), table(last_fused_step(after))
names = list(B) + [n for n in A if n not in B]
print(f"# r2 -- kernels of one bench step, workload {w}: start of round 2 vs end of round 2\n")
print(f"`ncu --metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --clock-control none` over `python bench.py --steps 1 --warmup 3 "
      f"--workload {w}` (serialised, cold caches: the SHARES are what compares with the bench's stage times, not the absolute "
      f"sum).  before = `{before}`, after = `{after}`.\n")
print("| kernel | launches before | us before | launches after | us after | warp inst after (M) |")
print("|---|---|---|---|---|---|")
for n in names:
    a, b = A.get(n, [0, 0.0, 0.0]), B.get(n, [0, 0.0, 0.0])
    print(f"| `{n}` | {a[0]} | {a[1]:.1f} | {b[0]} | {b[1]:.1f} | {b[2]:.1f} |" if b[2] else
          f"| `{n}` | {a[0]} | {a[1]:.1f} | {b[0]} | {b[1]:.1f} | |")
print(f"| **sum** | {sum(a[0] for a in A.values())} | {sum(a[1] for a in A.values()):.1f} | {sum(b[0] for b in B.values())} | "
      f"{sum(b[1] for b in B.values()):.1f} | {sum(b[2] for b in B.values()):.1f} |")
