set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
for w in F R P; do
  for c in fused two-phase; do
    timeout 300 python bench.py --workload $w --contract $c --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_${w}_$c.json 2> $O/bench_${w}_$c.err
  done
done
