# usage: bash scripts/gpu_r2m.sh N
N=$1
set -x
O=gpurun_out/r2m
mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests -m gpu -q -x -k "fixture or fused or specul or stage or suite_analog" 2>&1 | tail -5 > $O/tests.log
  for w in F P R; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_$w.json 2> $O/bench_$w.err
  done
  B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_F.csv $B --workload F > $O/ncu_F.log 2>&1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline > $O/F_n${N}_peer.json 2> $O/F_n${N}_peer.err
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline --contract two-phase > $O/F_n${N}_peer_twophase.json 2> $O/F_n${N}_peer_twophase.err
fi
