# usage: bash scripts/gpu_r2o.sh N
N=$1
set -x
O=gpurun_out/r2o
mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
  timeout 300 python bench.py --workload F --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_F.json 2> $O/bench_F.err
  for w in P R FP; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_$w.json 2> $O/bench_$w.err
  done
  timeout 300 python bench.py --workload F --contract two-phase --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed > $O/bench_F_twophase.json 2> $O/bench_F_twophase.err
  B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity --no-perturbed"
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 600 --csv --log-file $O/launches_R.csv $B --workload R > $O/ncu_R.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_F.csv $B --workload F > $O/ncu_F.log 2>&1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline > $O/F_n${N}_peer.json 2> $O/F_n${N}_peer.err
  MHB_BENCH_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline > $O/F_n${N}_peer_trace.json 2> $O/F_n${N}_peer_trace.err
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline --contract two-phase > $O/F_n${N}_peer_twophase.json 2> $O/F_n${N}_peer_twophase.err
  timeout 600 python -m pytest tests -m gpu -q -x -k "multi_rank" 2>&1 | tail -5 > $O/tests_n$N.log
fi
