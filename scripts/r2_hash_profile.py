"""profiles/r2l_hash_kernels_R.md from the ncu --set full capture of the hash / thread-per-row kernels of the
webbase-like input (gpurun call r2l):  python scripts/r2_hash_profile.py gpurun_out/r2l/prof_hash_R.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
K = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem KB"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("smsp__inst_executed.sum", "warp inst M"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU %"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts M"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "conflict M"),
     ("smsp__inst_executed_op_shared_atom.sum", "smem atom M")]
print("# r2l -- `ncu --set full` of the hash and thread-per-row kernels, webbase-like R-MAT (configs[2]), BEFORE the round-2 changes to them\n")
print("Command: `ncu --set full --clock-control none --import-source on -k regex:\"k_num_hash_list|k_sym_hash_group|k_num_hash_block|"
      "k_sym_hash_block|k_num_tiny|k_sym_tiny\" -s 45 -c 15 python bench.py --steps 1 --warmup 3 --workload R ...` (gpurun call r2l, one B200).\n")
print("| kernel | " + " | ".join(n for _, n in K) + " |")
print("|---|" + "---|" * len(K))
ni = h.index("Kernel Name")
for r in rows[2:]:
    name = r[ni].split("(")[0].replace("void ", "").replace("mhb::", "")
    vals = []
    for k, n in K:
        v = float(r[h.index(k)].replace(",", "")) if k in h and r[h.index(k)] else 0.0
        if n.endswith(" M"):
            v /= 1e6
        vals.append(f"{v:.1f}" if v < 1000 else f"{v:.0f}")
    print(f"| `{name}` | " + " | ".join(vals) + " |")
print("""
Reading (57.3 M products, 57.0 M entries of C: compression 1.004):

* every kernel is instruction-bound at low occupancy (issue 33-71 %, LSU 30-58 %): the numeric kernels execute
  1 005 M warp instructions for 57 M products -- 17 per product -- the symbolic ones 559 M.
* `k_num_tiny` / `k_sym_tiny` run **4 of 32 lanes** (thr/inst 4.3 / 4.0): one thread per row, nested loops (nonzeros of
  A, entries of the B row) that reconverge at the end of every B row, and rows whose cost differs 100x side by side.
  302 K rows with 2.8 M products cost 92.6 M warp instructions.
* source page of `k_num_hash_list` (1 024-slot bin): 18 % of the instructions are the rank-inside-bucket loop of
  the sort (16-24 lanes active: R-MAT columns cluster, so linear buckets are uneven), 20 % the lockstep find-or-claim
  loop, 5 % the fp64 shared-memory atomicAdd, 10 % the flat expansion.
* 94 % of the rows of these bins (57 % of all products) have nnz == products: no two products share a column.

Consequences (this round): duplicate-free rows skip the hash table (products stored in expansion order, then the same
bucket sort), the thread-per-row kernels walk a row's products in ONE loop over rows sorted into three cost classes.
After (gpurun call r2n/r2p, `profiles/r2_launches_R.md`): numeric warp instructions 1 005 M -> 790 M, numeric phase
1.63 -> 1.39 ms, step 2.67 -> 2.38 ms.""")
