set -x
O=gpurun_out/r2z
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/tests.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
for w in P R FP; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_$w.json 2> $O/bench_$w.err
done
MHB_RMAT_SCALE=20 timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_G20.json 2> $O/bench_G20.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity --no-perturbed"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 600 --csv --log-file $O/launches_R.csv $B --workload R > $O/ncu_R.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file $O/launches_F.csv $B --workload F > $O/ncu_F.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_num_hash_list|k_sym_hash_group|k_num_hash_block|k_sym_hash_block|k_num_tiny|k_sym_tiny" -s 45 -c 15 -o $O/prof_hash_R_after $B --workload R > $O/ncu_R_full.log 2>&1
ls -la $O
