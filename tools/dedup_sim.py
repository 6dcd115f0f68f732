#!/usr/bin/env python
"""How many grad_value reduction rows remain when the reductions of G consecutive queries of one head that hit the same value row
are merged: the bench's "local" location distribution on the R50_ovis_360 pyramid, per level, for G = 1 (what the kernel
does today: merge inside one query) .. 64.  Pure numpy, no GPU."""
import numpy as np
rng=np.random.default_rng(0)
pyr=[(48,80),(24,40),(12,20),(6,10)]
# reference points: pixel centres of every level
refs=[]; qlvl=[]; qxy=[]
for l,(H,W) in enumerate(pyr):
    ys,xs=np.meshgrid(np.arange(H)+0.5,np.arange(W)+0.5,indexing='ij')
    refs.append(np.stack([xs.ravel()/W,ys.ravel()/H],-1)); qlvl+= [l]*(H*W)
ref=np.concatenate(refs); Lq=len(ref); qlvl=np.array(qlvl)
M=2  # heads are iid; 2 is enough for stats
loc=ref[:,None,None,None,:]+0.05*rng.standard_normal((Lq,M,4,4,2))
loc=np.clip(loc,-0.1,1.1)
tot_unmerged=0
def rows_for_level(l):
    H,W=pyr[l]
    x=loc[:,:,l,:,0]*W-0.5; y=loc[:,:,l,:,1]*H-0.5
    x0=np.floor(x).astype(int); y0=np.floor(y).astype(int)
    out=[]
    for dy in (0,1):
        for dx in (0,1):
            xx=x0+dx; yy=y0+dy
            ok=(xx>=0)&(xx<W)&(yy>=0)&(yy<H)
            out.append(np.where(ok,yy*W+xx,-1))
    return np.stack(out,-1).reshape(Lq,M,16)   # [q, m, 16 corner rows]
def count(group_ids, r):
    # distinct (group, m, row)
    n=0
    G=group_ids.max()+1
    for m in range(M):
        key=group_ids[:,None]*100000+r[:,m,:]
        key=key[r[:,m,:]>=0]
        n+=len(np.unique(key))
    return n/M
for l in range(4):
    r=rows_for_level(l)
    valid=(r>=0).sum()/M
    res={}
    q=np.arange(Lq)
    for g in (1,2,4,8,16,32,64):
        res[g]=count(q//g,r)
    print("level",l,"valid corner rows/head",valid, {g:round(v/valid,3) for g,v in res.items()})
