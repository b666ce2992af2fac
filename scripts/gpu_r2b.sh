set -x
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q -x -k "multi_rank or suite_analog or transpose or aat or cli" 2>&1 | tail -30 > gpurun_out/r2b/tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b/bench_F.json 2> gpurun_out/r2b/bench_F.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2b/bench_F_ref.json 2> gpurun_out/r2b/bench_F_ref.err
timeout 300 python bench.py --workload R --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench_R.json 2> gpurun_out/r2b/bench_R.err
