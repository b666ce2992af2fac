import sys, numpy as np
sys.path.insert(0,'/root/repo'); import mh_spgemm_b200
from mh_spgemm_b200 import generators as G
A = G.fem3d(8,8,12,3,seed=1)
ptr, col = A.ptr, A.col
def wavefronts(slots):
    # slots: array of slot per lane for one chunk (<=32 lanes); 64-bit access: per half warp, max multiplicity per bank pair
    w = 0
    for h in (slots[:16], slots[16:]):
        if len(h)==0: continue
        u = np.unique(h)  # same address -> broadcast
        w += np.bincount(u % 16, minlength=16).max()
    return w
def layouts(ccols):
    ccols = np.asarray(ccols); cmin = ccols.min()
    out = {}
    rank = {c:i for i,c in enumerate(ccols)}
    out['rank'] = lambda c: rank[c]
    out['dense'] = lambda c: c - cmin
    hws = np.unique(ccols >> 4); hwi = {h:i for i,h in enumerate(hws)}
    out['half16'] = lambda c: hwi[c>>4]*16 + (c & 15)
    ws = np.unique(ccols >> 5); wi = {h:i for i,h in enumerate(ws)}
    out['word32'] = lambda c: wi[c>>5]*32 + (c & 31)
    # rank + pad per C run (gap in columns)
    runs = np.concatenate([[0], np.cumsum(np.diff(ccols) > 1)])
    runid = {c:r for c,r in zip(ccols, runs)}
    for pad in (1,2,9,10):
        out[f'rank+{pad}/run'] = (lambda pad: (lambda c: rank[c] + pad*runid[c]))(pad)
    # rank with word-indexed xor swizzle
    out['rank^word'] = lambda c: rank[c] ^ ((c>>5) & 15)
    out['rank+word'] = lambda c: rank[c] + wi[c>>5]
    out['rank+3word'] = lambda c: rank[c] + 3*wi[c>>5]
    return out, {k: None for k in out}
tot = {}
ideal = 0
rows = range(3*(8*8*5+8*3+3), 3*(8*8*5+8*3+3)+3*4, 3)  # a few interior rows
rows = list(range(A.M//2, A.M//2 + 96, 3))
for r in rows:
    ak = col[ptr[r]:ptr[r+1]]
    ccols = np.unique(np.concatenate([col[ptr[k]:ptr[k+1]] for k in ak]))
    L, _ = layouts(ccols)
    size = {}
    for name, f in L.items():
        size[name] = max(f(c) for c in ccols)+1
    # steps: B rows grouped by 3 twins (fold) -> use first of each node
    for k in ak[::3]:
        bc = col[ptr[k]:ptr[k+1]]
        for t in range(0, len(bc), 32):
            ch = bc[t:t+32]
            ideal += (1 if len(ch)<=16 else 2)
            for name, f in L.items():
                tot[name] = tot.get(name,0) + wavefronts(np.array([f(c) for c in ch]))
print('ideal', ideal)
for k,v in sorted(tot.items(), key=lambda kv: kv[1]): print(f'{k:14s} wavefronts {v:7d}  x{v/ideal:.2f}  slots {size[k]}')
