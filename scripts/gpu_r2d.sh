set -x
mkdir -p gpurun_out/r2d
nvidia-smi -L > gpurun_out/r2d/gpus.txt
nvidia-smi topo -m > gpurun_out/r2d/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --steps 20 --warmup 3 --no-suite --no-cpu-baseline > gpurun_out/r2d/bench_n1.json 2> gpurun_out/r2d/bench_n1.err
for ex in peer broadcast sendrecv; do
  timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex > gpurun_out/r2d/bench_n2_$ex.json 2> gpurun_out/r2d/bench_n2_$ex.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "multi_rank" 2>&1 | tail -15 > gpurun_out/r2d/tests.log
export MHB_RMAT_SCALE=20
timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d/bench_G20_n1.json 2> gpurun_out/r2d/bench_G20_n1.err
for ex in peer broadcast; do
  timeout 300 $TR bench.py --gpus 2 --workload G --steps 3 --warmup 3 --exchange $ex > gpurun_out/r2d/bench_G20_n2_$ex.json 2> gpurun_out/r2d/bench_G20_n2_$ex.err
done
