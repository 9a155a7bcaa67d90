"""Per-ray CPU emulation of the Siddon integer walk (csrc/siddon.cu, walk_voxel): crossings merged as the kernel
merges them (ties to the lower axis), the walk started at the first segment the three-axis certificate accepts and
advanced by +-stride at every opening crossing, the one-compare certificate min|d_a| * len / 2 > tol deciding whether
the walk's cell is used -- compared with the oracle's reconstruction of the reference's indices on every segment of
positive length.  Round 1: 360 rays, 52 439 segments, 99.6 % served by the walk, 0 wrong.
    PYTHONPATH=. python scripts/emulate_siddon_walk.py
"""
import torch, numpy as np, oracle
from tests._scene import pixel_size
from bench import pose_batch
N=128; H=W=48; SDD=1020.0
sp=256.0/N
aff=torch.diag(torch.tensor([sp,sp,sp,1.0])); aff[:3,3]=-sp*(N-1)/2
affinv=torch.linalg.inv(aff)[None]
rot,xyz=pose_batch(116,0)
P=3
pose=oracle.pose_from_params(rot[:P],xyz[:P],"euler_angles","ZXY")
s,t=oracle.detector_rays(pose,oracle.REORIENT["AP"],H,W,pixel_size(H),pixel_size(H),0.,0.,SDD,False)
s,t=oracle.apply(affinv,s),oracle.apply(affinv,t)
shift=0.5; eps=1e-8
shape=(N,N,N)
ref_idx,ref_seg=oracle.siddon_segments(shape,s,t,voxel_shift=shift)
f32=np.float32
S=s.numpy(); T=t.numpy()
tot=cheap=bad=0; init_late=0
rng=np.random.default_rng(0)
rays=[(b,int(r)) for b in range(P) for r in rng.choice(H*W,120,replace=False)]
for b,r in rays:
    src=S[b,0].astype(f32); d=((T[b,r]-src).astype(f32)+f32(eps)).astype(f32)
    # per-axis crossings exactly as the reference computes them (fp32), then the kernel's merge (ties -> lower axis)
    lo=(np.zeros(3,f32)-f32(shift)); hi=(np.array(shape,f32)-f32(shift))
    a0=((lo-src)/d).astype(f32); a1=((hi-src)/d).astype(f32)
    amin=max(np.minimum(a0,a1).max(),f32(0)); amax=min(np.maximum(a0,a1).min(),f32(1))
    if not amin<amax: continue
    ev=[]
    for a in range(3):
        planes=(np.arange(shape[a]+1,dtype=f32)-f32(shift))
        al=((planes-src[a])/d[a]).astype(f32)
        ok=(al>=amin)&(al<=amax)
        for v in al[ok]: ev.append((float(v),a))
    ev.sort(key=lambda e:(e[0],e[1]))
    mag=np.abs(d).max(); dmin_half=f32(0.5)*np.abs(d).min()
    tol=f32(1.5*5.9604645e-8)*(mag+f32(7.0)*f32(N))
    step=[1 if d[a]>0 else -1 for a in range(3)]
    M=len(ev)-1
    refi=ref_idx[b,r].numpy(); refs=ref_seg[b,r].numpy()
    walk=None
    for m in range(M):
        prev,aprev=ev[m]; nxt,_=ev[m+1]
        segl=f32(nxt)-f32(prev)
        mid=f32((f32(prev)+f32(nxt))*f32(0.5))
        if walk is not None:
            walk[aprev]+=step[aprev]
        # exact-path certificate (kernel's midpoint_voxel_checked): distance of u from rounding boundary on all axes
        u=(mid*d+src).astype(f32)+f32(shift-0.5)
        rr=np.rint(u); certain=(np.abs(u-rr).max()<0.5-tol)
        if walk is None and certain:
            walk=[int(v) for v in rr]; init_late+= (m>0)
        use_walk = walk is not None and (segl*dmin_half>tol)
        if segl>0:
            tot+=1
            if use_walk:
                cheap+=1
                flat=(walk[0]*N+walk[1])*N+walk[2]
                if not (0<=min(walk) and max(walk)<N) or flat!=refi[m]: bad+=1
print("rays",len(rays),"segments with length",tot,"served by the walk %.4f"%(cheap/tot),"wrong",bad,"rays initialised after their first segment",init_late)
