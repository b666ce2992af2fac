import sys, numpy as np
sys.path.insert(0,'/root/repo')
import mh_spgemm_b200
from mh_spgemm_b200 import generators as G
import scipy.sparse as sp
ladder=[(24,'G8/tiny',32),(80,'XS',128),(160,'S',256),(320,'M',512),(640,'L',1024),(2560,'BLOCK_S',4096),(10240,'BLOCK_L',16384),(1<<30,'GLOBAL',0)]
def smem(S): return S*12 + (S//4+4)*4 + (S//8)*5*8
for name in ("offshore","cage12","cop20k_A","webbase-1M","scircuit","mac_econ_fwd500"):
    A=G.suite(name)
    S=A.to_scipy().astype(bool).astype(np.float32)
    C=(S@S).tocsr()
    n=np.diff(C.indptr)
    ip=np.asarray(S@np.diff(S.indptr).astype(np.float64)).ravel() if False else None
    rows=len(n)
    line=f"{name}: rows {rows}, nnzC {C.nnz}, mean n {n[n>0].mean():.0f}, max {n.max()}; "
    lo=0; parts=[]
    for hi,lab,slots in ladder:
        m=(n>lo)&(n<=hi)
        if m.sum():
            fill=n[m].mean()/slots if slots else 0
            parts.append(f"{lab}: {100*m.sum()/rows:.1f}% rows / {100*n[m].sum()/C.nnz:.1f}% nnz, mean fill {fill:.2f}")
        lo=hi
    print(line+" | ".join(parts))
    # finer ladder with 1.6x tables (non power of two): occupancy estimate by shared memory
    fine=[48,80,120,160,240,320,480,640]
    lo=24; tot_now=0; tot_fine=0
    for hi in fine:
        m=(n>lo)&(n<=hi)
        if m.sum():
            S_now=[s for h,l,s in ladder if hi<=h][0]
            S_fine=int(np.ceil(hi*1.6/8)*8)
            w_now=min(32 if S_now<=256 else 32, 227*1024//(smem(S_now)+1024)) if S_now>256 else min(40,227*1024//(smem(S_now)+256))
            w_fine=min(40,227*1024//(smem(S_fine)+256))
            tot_now+=n[m].sum()/w_now; tot_fine+=n[m].sum()/w_fine
        lo=hi
    if tot_now: print(f"   relative latency-bound time with 1.6x tables vs now (warps/SM by smem, cap 40): {tot_fine/tot_now:.2f}")
