set -x
mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2c/tests.log
for w in F R P; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > gpurun_out/r2c/bench_$w.json 2> gpurun_out/r2c/bench_$w.err; done
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2c/bench_F_ref.json 2> gpurun_out/r2c/bench_F_ref.err
