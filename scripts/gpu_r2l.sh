set -x
O=gpurun_out/r2l
mkdir -p $O
for w in F P; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite > $O/bench_$w.json 2> $O/bench_$w.err
done
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-suite --no-parity"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_num_hash_list|k_sym_hash_group|k_num_hash_block|k_sym_hash_block|k_num_tiny|k_sym_tiny" -s 45 -c 15 -o $O/prof_hash_R $B --workload R > $O/ncu_R.log 2>&1
ls -la $O
