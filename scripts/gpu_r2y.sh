set -x
O=gpurun_out/r2y
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "mask_builder or outstanding or fixture or fused or stage_outputs or fuzz or host_path" 2>&1 | tail -8 > $O/tests.log
for w in F P R; do
  for m in 1 2; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed --opt mask_onepass=$m > $O/bench_${w}_mask$m.json 2> $O/bench_${w}_mask$m.err
  done
done
MHB_RMAT_SCALE=20 timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline --opt mask_onepass=2 > $O/bench_G20_mask2.json 2> $O/bench_G20_mask2.err
