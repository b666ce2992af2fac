set -x
O=gpurun_out/r2q
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
for w in R F; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-suite --no-perturbed > $O/bench_$w.json 2> $O/bench_$w.err
done
MHB_RMAT_SCALE=20 timeout 300 python bench.py --workload G --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_G20.json 2> $O/bench_G20.err
timeout 120 scripts/ab/bload_ab > $O/bload_ab.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $O/bload_ab_ncu.csv scripts/ab/bload_ab > $O/bload_ab_ncu.log 2>&1
