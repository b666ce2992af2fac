set -x
O=gpurun_out/r2p
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
for w in F P R; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 300 python bench.py --workload FP --steps 10 --warmup 3 --no-cpu-baseline --no-suite --opt compact_rows=0 > $O/bench_FP_nocompact.json 2> $O/bench_FP_nocompact.err
timeout 300 python bench.py --workload FP --steps 10 --warmup 3 --no-cpu-baseline --no-suite --opt compact_rows=0 --opt row_twins=1 > $O/bench_FP_rowtwins.json 2> $O/bench_FP_rowtwins.err
MHB_RMAT_SCALE=20 timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_G20.json 2> $O/bench_G20.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity --no-perturbed"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 600 --csv --log-file $O/launches_R.csv $B --workload R > $O/ncu_R.log 2>&1
