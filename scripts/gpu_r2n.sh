set -x
O=gpurun_out/r2n
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
for w in F P R; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 300 python bench.py --workload F --contract two-phase --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_F_twophase.json 2> $O/bench_F_twophase.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 600 --csv --log-file $O/launches_R.csv $B --workload R > $O/ncu_R.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_F.csv $B --workload F > $O/ncu_F.log 2>&1
