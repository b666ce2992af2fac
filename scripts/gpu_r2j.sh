set -x
mkdir -p gpurun_out/r2j
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
MHB_BENCH_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite > gpurun_out/r2j/F_n${N}_peer_trace.json 2> gpurun_out/r2j/F_n${N}_peer_trace.err
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-suite > gpurun_out/r2j/F_n${N}_peer.json 2> gpurun_out/r2j/F_n${N}_peer.err
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --exchange broadcast > gpurun_out/r2j/F_n${N}_broadcast.json 2> gpurun_out/r2j/F_n${N}_broadcast.err
export MHB_RMAT_SCALE=22
timeout 900 $TR bench.py --gpus $N --workload G --steps 3 --warmup 3 > gpurun_out/r2j/G22_n${N}_peer.json 2> gpurun_out/r2j/G22_n${N}_peer.err
timeout 900 $TR bench.py --gpus $N --workload G --steps 3 --warmup 3 --exchange broadcast > gpurun_out/r2j/G22_n${N}_broadcast.json 2> gpurun_out/r2j/G22_n${N}_broadcast.err
nproc > gpurun_out/r2j/nproc.txt; cat /proc/cpuinfo | grep "model name" | sort | uniq -c >> gpurun_out/r2j/nproc.txt
rm -f /dev/shm/mhb_bench_G_*.npz
