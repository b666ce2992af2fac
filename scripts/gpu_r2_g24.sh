# usage: bash scripts/gpu_r2_g24.sh N EXCHANGE   (under gpurun --gpus N): BASELINE configs[4] at full size, R-MAT scale 24
N=$1
X=${2:-peer}
set -x
O=gpurun_out/r2x
mkdir -p $O
export MHB_RMAT_SCALE=24
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 1200 $TR bench.py --gpus $N --workload G --steps 2 --warmup 3 --no-cpu-baseline --exchange $X > $O/G24_n${N}_$X.json 2> $O/G24_n${N}_$X.err
tail -5 $O/G24_n${N}_$X.err
nvidia-smi --query-gpu=memory.used --format=csv > $O/mem_n${N}_$X.txt
