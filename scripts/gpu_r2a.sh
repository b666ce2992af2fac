set -x
mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpus.txt
timeout 900 python -m pytest tests -m gpu -q -k "not suite_analog" 2>&1 | tail -25 > gpurun_out/r2a/tests.log
timeout 600 python tests/golden/make_golden.py gpurun_out/r2a/golden --suite > gpurun_out/r2a/golden.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/memcheck_smoke.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/racecheck_smoke.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python -m pytest tests/test_gpu_parity.py -q -x -k "window_kernel_variants or device_transpose or hash_probe" > gpurun_out/r2a/racecheck_twins.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -q -x -k "window_kernel_variants or device_transpose or hash_probe or (spgemm_matches_oracle and rmat14)" > gpurun_out/r2a/memcheck_twins.log 2>&1
for w in F R P; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_$w.json 2> gpurun_out/r2a/bench_$w.err; done
