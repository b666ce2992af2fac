set -x
mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2g/tests.log
for w in F R P; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > gpurun_out/r2g/bench_$w.json 2> gpurun_out/r2g/bench_$w.err; done
MHB_RMAT_SCALE=20 timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g/bench_G20.json 2> gpurun_out/r2g/bench_G20.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g/launches_F.csv $B --workload F > gpurun_out/r2g/ncu_F.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g/launches_R.csv $B --workload R > gpurun_out/r2g/ncu_R.log 2>&1
MHB_RMAT_SCALE=20 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g/launches_G20.csv $B --workload G > gpurun_out/r2g/ncu_G.log 2>&1
