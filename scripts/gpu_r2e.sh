set -x
mkdir -p gpurun_out/r2e
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ex in peer sendrecv; do
  timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex > gpurun_out/r2e/bench_n2_$ex.json 2> gpurun_out/r2e/bench_n2_$ex.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-suite --no-cpu-baseline > gpurun_out/r2e/bench_n1.json 2> gpurun_out/r2e/bench_n1.err
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-suite --no-cpu-baseline > gpurun_out/r2e/bench_n1_gpu1.json 2> gpurun_out/r2e/bench_n1_gpu1.err
