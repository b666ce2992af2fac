set -x
O=gpurun_out/r2aa
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/tests.log
timeout 300 python bench.py --workload R --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_R.json 2> $O/bench_R.err
MHB_RMAT_SCALE=20 timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_G20.json 2> $O/bench_G20.err
timeout 900 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
