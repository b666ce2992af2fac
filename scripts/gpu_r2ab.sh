# usage: bash scripts/gpu_r2ab.sh N
N=$1
set -x
O=gpurun_out/r2ab
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline > $O/F_n${N}_peer.json 2> $O/F_n${N}_peer.err
