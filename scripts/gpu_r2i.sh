set -x
mkdir -p gpurun_out/r2i
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
MHB_BENCH_TRACE=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2i/bench_n2_peer.json 2> gpurun_out/r2i/bench_n2_peer.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-suite --no-cpu-baseline > gpurun_out/r2i/bench_n1.json 2> gpurun_out/r2i/bench_n1.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange broadcast > gpurun_out/r2i/bench_n2_broadcast.json 2> gpurun_out/r2i/bench_n2_broadcast.err
export MHB_RMAT_SCALE=20
timeout 300 $TR bench.py --gpus 2 --workload G --steps 3 --warmup 3 > gpurun_out/r2i/bench_G20_n2_peer.json 2> gpurun_out/r2i/bench_G20_n2_peer.err
timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i/bench_G20_n1.json 2> gpurun_out/r2i/bench_G20_n1.err
