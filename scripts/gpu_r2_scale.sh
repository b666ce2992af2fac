# usage: bash scripts/gpu_r2_scale.sh N   (run under gpurun --gpus N)
N=$1
set -x
mkdir -p gpurun_out/r2s
nvidia-smi topo -m > gpurun_out/r2s/topo_n$N.txt 2>&1
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite --no-cpu-baseline > gpurun_out/r2s/F_n${N}_peer.json 2> gpurun_out/r2s/F_n${N}_peer.err
if [ "$N" != "1" ]; then
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --exchange broadcast > gpurun_out/r2s/F_n${N}_broadcast.json 2> gpurun_out/r2s/F_n${N}_broadcast.err
fi
export MHB_RMAT_SCALE=22
timeout 900 $TR bench.py --gpus $N --workload G --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s/G22_n${N}_peer.json 2> gpurun_out/r2s/G22_n${N}_peer.err
if [ "$N" != "1" ]; then
timeout 900 $TR bench.py --gpus $N --workload G --steps 3 --warmup 3 --exchange broadcast > gpurun_out/r2s/G22_n${N}_broadcast.json 2> gpurun_out/r2s/G22_n${N}_broadcast.err
fi
rm -f /dev/shm/mhb_bench_G_*.npz
