"""The box rules of csrc/trilinear_staged.cu (corner-ray bounds, margins, clipping, 16-byte alignment, capacity)
mirrored in numpy on config-2 geometry: how many slabs fit the staging buffer, how many samples fall inside their box,
how large the boxes are and how many bytes a launch of 116 poses would stage.
    K=4 CAP=12288 python scripts/sim_staged_box_rules.py        (K = layers per slab, CAP = floats per buffer)
Round 1:  K=2: all fit, boxes <= 7 056 floats, 100 GB staged;  K=4: all fit, <= 12 000, 86 GB (the shipped default);
          K=8: 55 % fit 12 288 floats, all fit 24 576 (<= 22 464), 82 GB."""
import sys, os
sys.argv=['x']
_here=os.path.dirname(os.path.abspath(__file__))
src=open(os.path.join(_here,'sim_brick_staging.py')).read().split('if __name__ == "__main__":')[0]
g={"__file__":os.path.join(_here,'sim_brick_staging.py')}
exec(compile(src,"sim","exec"),g)
import numpy as np
src_, tgt, amin, amax, P, DET, NP, N = g["src"], g["tgt"], g["amin"], g["amax"], g["P"], g["DET"], g["NP"], g["N"]
tile_samples=g["tile_samples"]
import os
K=int(os.environ.get("K","4")); CAP=int(os.environ.get("CAP","12288")); T=16
rng=np.random.default_rng(0)
tot=inbox=0; slabs=fit=0; elems=[]; axes=[0,0,0]; fill=0.0; ntiles=0
for b in range(P):
    for _ in range(12):
        i0=int(rng.integers(0,DET//T))*T; j0=int(rng.integers(0,DET//T))*T
        x=tile_samples(b,i0,j0,T,T)
        ok=~np.isnan(x[...,0])
        if ok.sum()==0: continue
        ntiles+=1
        dc=tgt[b,i0+T//2,j0+T//2]-src_[b]
        A=int(np.abs(dc).argmax()); fwd=dc[A]>0; axes[A]+=1
        O1=1 if A==0 else 0; O2=1 if A==2 else 2
        corners=[(i0,j0),(i0+T-1,j0),(i0,j0+T-1),(i0+T-1,j0+T-1)]
        ix=np.floor(x).astype(np.int64)
        slab=np.where(ok,(np.maximum(ix[...,A],-K)+K)//K,-1)
        for s in np.unique(slab[slab>=0]):
            m=slab==s
            loA=(s-1)*K
            pts=[]
            for (ci,cj) in corners:
                d=tgt[b,ci,cj]-src_[b]
                for pl in (loA,loA+K):
                    al=(pl-src_[b][A])/d[A]
                    pts.append(src_[b]+al*d)
            pts=np.array(pts)
            lo=[0,0,0]; hi=[0,0,0]
            lo[A]=loA; hi[A]=loA+K
            for o in (O1,O2):
                lo[o]=int(np.floor(pts[:,o].min()))-1; hi[o]=int(np.floor(pts[:,o].max()))+2
            for a in range(3):
                lo[a]=max(lo[a],-1); hi[a]=min(hi[a],N)
            lo[2]&=~3; hi[2]=((hi[2]+4)&~3)-1
            E=[hi[a]-lo[a]+1 for a in range(3)]
            n=E[0]*E[1]*E[2]; elems.append(n)
            slabs+=1
            cnt=m.sum(); tot+=cnt
            if n<=CAP:
                fit+=1; fill+=n*4
                q=ix[m]
                inb=np.ones(len(q),bool)
                for a in range(3):
                    l=q[:,a]-lo[a]
                    inb&=(l>=0)&(l<E[a]-1)
                inbox+=inb.sum()
print("tiles",ntiles,"axis histogram",axes)
print("slabs fitting CAP: %.3f"%(fit/slabs),"samples served from the box: %.4f"%(inbox/tot))
e=np.array(elems); print("box elems: median %d p90 %d max %d"%(np.median(e),np.percentile(e,90),e.max()))
print("fill GB/launch (fitting slabs): %.1f"%(fill/ntiles*(DET//T)**2*116/1e9))
