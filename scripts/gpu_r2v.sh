set -x
O=gpurun_out/r2v
mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
MHB_RMAT_SCALE=16 MHB_RMAT_GEN=device timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_G16_dev.json 2> $O/bench_G16_dev.err
timeout 300 python bench.py --workload R --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_R.json 2> $O/bench_R.err
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
