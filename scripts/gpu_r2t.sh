set -x
O=gpurun_out/r2t
mkdir -p $O
MHB_TRACE_HOST=1 timeout 300 python bench.py --workload F --steps 3 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed --no-parity > $O/bench_F.json 2> $O/bench_F.err
python - > $O/pcie.txt 2>&1 <<'PY'
import torch, time
d = torch.empty(202 << 20, dtype=torch.uint8, device="cuda")
h = torch.empty(202 << 20, dtype=torch.uint8).pin_memory()
for n in (202 << 20, 50 << 20, 17 << 20):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h[:n].copy_(d[:n], non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("D2H", n >> 20, "MiB", round((t1 - t0) * 1e3, 3), "ms", round(n / (t1 - t0) / 1e9, 1), "GB/s")
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d[:n].copy_(h[:n], non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("H2D", n >> 20, "MiB", round((t1 - t0) * 1e3, 3), "ms", round(n / (t1 - t0) / 1e9, 1), "GB/s")
PY
