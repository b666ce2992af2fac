"""Where a multi-GPU bench step spends its time (rank 0, wall clock with syncs between phases).
torchrun --nproc-per-node N scripts/step_profile.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mh_spgemm_b200
from mh_spgemm_b200 import api, generators as G
from mh_spgemm_b200.distributed import RangeExchange, SliceSizes, column_range, partition_rows, row_work, slice_offsets

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
A = G.fem3d(8, 8, 325 * world, 3, seed=1)
work = row_work(A, A); b = partition_rows(work, world)
blk = A.rows(int(b[rank]), int(b[rank + 1]))
tool = api.Tool(local); tool.set_stream(torch.cuda.current_stream().cuda_stream)
ap, ac, av = (torch.from_numpy(x).to(dev) for x in (blk.ptr, blk.col, blk.val))
kr = [column_range(A.rows(int(b[r]), int(b[r + 1]))) for r in range(world)]
plan = RangeExchange(rank, world, b, kr, A.ptr, torch.float64, dev)
acs = ac - plan.k0
oc, ov = plan.own_views()
if oc.numel() == ac.numel():
    oc.copy_(ac); ov.copy_(av)
else:
    oc, ov = ac, av
sizes = SliceSizes(rank, world, dev)
T = {k: [] for k in ("exchange", "symbolic", "alloc", "numeric", "offsets", "offsets_old")}
for it in range(13):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); bp, bc, bv = plan.run(oc, ov); torch.cuda.synchronize(); t1 = time.perf_counter()
    cp, nnz = tool.symbolic(blk.M, plan.K_local, A.N, ap, acs, bp, bc[:plan.nnz_local]); t2 = time.perf_counter()
    cc = torch.empty(nnz, dtype=torch.int32, device=dev); cv = torch.empty(nnz, dtype=torch.float64, device=dev); t3 = time.perf_counter()
    tool.numeric_into(av, bv, cc, cv); t4 = time.perf_counter()
    sizes.gather(nnz); torch.cuda.synchronize(); t5 = time.perf_counter()
    slice_offsets(nnz, rank, world, dev, on_host=False); torch.cuda.synchronize(); t6 = time.perf_counter()
    if it >= 3:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            T[k].append(v * 1e3)
if rank == 0:
    print({k: round(float(np.median(v)), 3) for k, v in T.items()}, "sum", round(sum(float(np.median(v)) for v in T.values()), 3))
dist.destroy_process_group()
