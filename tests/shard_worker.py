"""One rank of the row-sharded SpGEMM parity check (launched by tests/test_gpu_parity.py and
bench-independent):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/shard_worker.py [case ...]

torch.distributed (gloo) is ONLY the transport of the two 128-byte set-up blobs; the data path
is the C ABI (mhb_shard_*): peer-mapped B windows, one-sided halo pull, one-sided slice sizes.
With fewer GPUs than ranks the ranks share a device (CUDA IPC works between processes on one
device), so the multi-rank path is exercised on a single-GPU box too.  Every rank compares ITS
slice of C with the host oracle on the same rows: structure bit-exact, values to 1e-12.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mh_spgemm_b200  # noqa: E402,F401
from mh_spgemm_b200 import api, generators as G  # noqa: E402
from mh_spgemm_b200.csr import CSR  # noqa: E402
from mh_spgemm_b200.distributed import Shard, partition_rows, row_work  # noqa: E402

CASES = {
    "fem": lambda: G.fem3d(4, 4, 40, 3, seed=51),                        # banded: a halo of a few rows
    "rmat": lambda: G.rmat(13, 8000, 40000, seed=52),                    # every block references all of B
    "poisson": lambda: G.poisson2d(40),
    "f32": lambda: G.fem3d(3, 3, 30, 2, seed=53).astype(np.float32),
    "empty_tail": lambda: CSR.from_coo(600, 600, np.arange(300), (np.arange(300) * 7) % 300,   # rows >= 300 empty
                                       rng=np.random.default_rng(54)),
}


def main():
    import torch
    import torch.distributed as dist
    from oracle import Oracle
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    tool = api.Tool(dev)
    orc = Oracle()
    # Ranks that share a GPU must never have a kernel of one process spin on a flag that a kernel
    # of another process writes (the two are not guaranteed to run at the same time; the pool's
    # profiling guide reports Xid 109 for exactly that).  With fewer GPUs than ranks every wait
    # is therefore satisfied BEFORE its kernel is launched: stream sync + host barrier.
    shared = torch.cuda.device_count() < world

    def host_fence():
        torch.cuda.synchronize()
        dist.barrier()

    def exchange(sh):
        if shared:
            sh.publish()
            host_fence()
            sh.pull()
        else:
            sh.exchange()

    def barrier(sh):
        if shared:
            host_fence()
        else:
            sh.barrier()
    names = sys.argv[1:] or list(CASES)
    for name in names:
        A = CASES[name]()
        B = A
        rtol = 1e-12 if A.val.dtype == np.float64 else 1e-5
        work = row_work(A, B)
        bounds = partition_rows(work, world)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        Ablk, Bown = A.rows(r0, r1), B.rows(r0, r1)
        dAp, dAc, dAv = api.DeviceArray(Ablk.ptr), api.DeviceArray(Ablk.col), api.DeviceArray(Ablk.val)
        dBp = api.DeviceArray(Bown.ptr)
        sh = Shard(tool, rank, world, B.M, B.N, A.val.dtype, bounds).build(Ablk.M, dAp, dAc, dBp)
        k0, k1, nimg, halo = sh.image()
        assert k0 <= r0 and k1 >= r1 and nimg >= Bown.nnz
        col_own, val_own = sh.own_B()
        assert col_own.numel() == Bown.nnz
        col_own.upload(Bown.col)
        Cp, Cc, Cv = orc.spgemm(Ablk, B)
        gp = orc.symbolic(A, B)
        for step in range(3):
            scale = 1.0 + step  # B's values change every step: the exchange must deliver the new ones
            val_own.upload(Bown.val * A.val.dtype.type(scale))
            exchange(sh)
            dCp = api.DeviceArray(count=Ablk.M + 1, dtype=np.int32)
            nnz = sh.symbolic(0, Ablk.M, dCp)
            dCc = api.DeviceArray(count=max(nnz, 1), dtype=np.int32)
            dCv = api.DeviceArray(count=max(nnz, 1), dtype=A.val.dtype)
            sh.numeric_into(dAv, dCc, dCv)
            sh.post_size(nnz)
            if shared:
                host_fence()
            off, tot, sizes = sh.offsets()
            cp, cc, cv = dCp.numpy(), dCc.numpy()[:nnz], dCv.numpy()[:nnz]
            assert np.array_equal(cp.astype(np.int64), Cp), f"{name} rank {rank}: row_ptr differs"
            assert np.array_equal(cc, Cc), f"{name} rank {rank}: col_idx differs"
            bad, first = orc.compare(Ablk.M, (cp, cc, cv), (Cp, Cc, Cv * A.val.dtype.type(scale)), rtol)
            assert bad == 0, f"{name} rank {rank} step {step}: {bad} values out of tolerance (first {first})"
            # slice offsets == the oracle's global row_ptr at the block boundaries
            assert off == int(gp[r0]) and tot == int(gp[-1]), (off, tot, int(gp[r0]), int(gp[-1]))
            assert sizes == [int(gp[int(bounds[r + 1])] - gp[int(bounds[r])]) for r in range(world)]
            barrier(sh)  # nobody may still be pulling when the next step rewrites the shard
            for d in (dCp, dCc, dCv):
                d.free()
        # the fused call through the shard layer (mhb_shard_spgemm_into_begin / _end): caller-owned C
        # arrays, slice size posted from the DEVICE behind the kernels and confirmed after the host
        # read; the second step runs with a single host synchronisation
        cap = max(int(Cp[-1]), 1)
        dCp = api.DeviceArray(count=Ablk.M + 1, dtype=np.int32)
        dCc, dCv = api.DeviceArray(count=cap, dtype=np.int32), api.DeviceArray(count=cap, dtype=A.val.dtype)
        fused0 = tool.stats["fused_calls"]
        for step in range(2):
            exchange(sh)
            sh.spgemm_into_begin(0, Ablk.M, dAv, dCp, dCc, dCv)
            sh.post_size(-1)
            nnz = sh.spgemm_into_end()
            sh.repost_size(nnz)
            if shared:
                host_fence()
            off, tot, sizes = sh.offsets()
            cp, cc, cv = dCp.numpy(), dCc.numpy()[:nnz], dCv.numpy()[:nnz]
            assert nnz == int(Cp[-1]) and np.array_equal(cp.astype(np.int64), Cp) and np.array_equal(cc, Cc), \
                f"{name} rank {rank}: fused structure differs"
            bad, first = orc.compare(Ablk.M, (cp, cc, cv), (Cp, Cc, Cv * A.val.dtype.type(3.0)), rtol)
            assert bad == 0, f"{name} rank {rank} fused step {step}: {bad} values out of tolerance (first {first})"
            assert off == int(gp[r0]) and tot == int(gp[-1]), (off, tot)
            barrier(sh)
        assert tool.stats["fused_calls"] >= fused0 + 1, tool.stats
        for d in (dCp, dCc, dCv):
            d.free()
        # the rows of a rank cut into two slices after ONE exchange (the int32-overflow path)
        if Ablk.M >= 2:
            exchange(sh)
            mid = Ablk.M // 2
            got = []
            for lo, hi in ((0, mid), (mid, Ablk.M)):
                dCp = api.DeviceArray(count=hi - lo + 1, dtype=np.int32)
                nnz = sh.symbolic(lo, hi, dCp)
                dCc = api.DeviceArray(count=max(nnz, 1), dtype=np.int32)
                dCv = api.DeviceArray(count=max(nnz, 1), dtype=A.val.dtype)
                sh.numeric_into(dAv, dCc, dCv)
                got.append((dCp.numpy(), dCc.numpy()[:nnz], dCv.numpy()[:nnz]))
            from mh_spgemm_b200.distributed import concat_slices
            p2, c2, v2 = concat_slices(got)
            assert np.array_equal(p2, Cp) and np.array_equal(c2, Cc)
            barrier(sh)
        torch.cuda.synchronize()
        sh.close()
        print(f"SHARD-OK {name} rank {rank}/{world} dev {dev} rows [{r0},{r1}) image [{k0},{k1}) halo {halo} B nnzC {int(Cp[-1])}",
              flush=True)
        dist.barrier()
    tool.release()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
