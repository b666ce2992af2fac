set -x
O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
for w in F R P; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 300 python bench.py --workload F --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed --opt row_chunks=1 > $O/bench_F_nochunk.json 2> $O/bench_F_nochunk.err
timeout 300 python bench.py --workload F --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed --opt row_chunks=8 > $O/bench_F_chunk8.json 2> $O/bench_F_chunk8.err
