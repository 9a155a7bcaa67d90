"""CPU emulation of volume_grad_kernel's per-voxel logic (csrc/volgrad.cu: pixel window -> per-ray alpha window ->
sample range -> trilinear hat weights) against a brute force over every ray and sample, on one pose of the
edge-pose set (default: pose 3, the source inside the volume).

    PYTHONPATH=. python scripts/emulate_gather_kernel.py [pose 0..4] 

`kernel(pv, thr)` uses the projected-corner window for voxels nearer than thr support-depths to the source plane and
the old linearised window beyond; the shipped kernel is thr = infinity (corner window everywhere): 0 of 475 voxels
differ on every pose.  thr = 8 (the partial fix that measured 3.0e-4 on the B200) leaves 7 voxels of pose 3 wrong.
"""
import numpy as np, torch, oracle, sys
from tests._scene import pixel_size
from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
N=64; H=W=32; SDD=1020.0; NP=500
sp=256.0/N
aff=torch.diag(torch.tensor([sp,sp,sp,1.0])); aff[:3,3]=-sp*(N-1)/2
affinv=torch.linalg.inv(aff)
pose=oracle.pose_from_params(torch.tensor(EDGE_ROT),torch.tensor(EDGE_XYZ),"euler_angles","ZXY")
reo=oracle.REORIENT["AP"]
cam2world=oracle.compose(reo[None],pose)
cam2vox=(affinv[None]@cam2world).double().numpy()
vox2cam=np.linalg.inv(cam2vox)
delx=pixel_size(H)
grid=oracle.detector_grid(H,W,delx,delx,0.0,0.0,SDD,False).view(H,W,3).double().numpy()
o=grid[0,0]; u=grid[1,0]-grid[0,0]; v=grid[0,1]-grid[0,0]
s_,t_=oracle.detector_rays(pose,reo,H,W,delx,delx,0.0,0.0,SDD,False)
s_=oracle.apply(affinv[None],s_).double().numpy(); t_=oracle.apply(affinv[None],t_).double().numpy()
amin,amax=oracle.alpha_minmax(torch.as_tensor(s_,dtype=torch.float32),torch.as_tensor(t_,dtype=torch.float32),torch.tensor([N-1.0]*3),1e-8)
amin=amin.double().numpy()[...,0]; amax=amax.double().numpy()[...,0]
uu=np.linspace(0,1,NP)
b=int(sys.argv[1]) if len(sys.argv)>1 else 3
Gi=vox2cam[b]; src=s_[b,0]; d=t_[b]-src           # (Nrays,3)
span=amax[b]-amin[b]
# rays contribute iff they pass the padded box (c != 0); emulate with span>0 or grazing: use brute force on positions
pos=src+(amin[b][:,None]+uu[None]*span[:,None])[...,None]*d[:,None,:]     # (Nrays,NP,3)
def brute(pv):
    w=np.clip(1-np.abs(pos-pv),0,None).prod(-1)     # (Nrays,NP)
    return w.sum(1)                                  # per ray weight sum
def kernel(pv, thr=1e30):
    row=Gi[:3,:3]; q=row@pv+Gi[:3,3]
    rz=np.abs(row[2]).sum()
    j0,j1,i0,i1=0,W-1,0,H-1
    if q[2]>thr*rz:
        m=SDD/q[2]; cj=(q[0]*m-o[0])/v[0]; ci=(q[1]*m-o[1])/u[1]
        rj=sum(abs(row[0][a]-(q[0]/q[2])*row[2][a]) for a in range(3))*m*abs(1/v[0])*1.01+1e-3
        ri=sum(abs(row[1][a]-(q[1]/q[2])*row[2][a]) for a in range(3))*m*abs(1/u[1])*1.01+1e-3
        j0=max(0,int(np.ceil(cj-rj))); j1=min(W-1,int(np.floor(cj+rj))); i0=max(0,int(np.ceil(ci-ri))); i1=min(H-1,int(np.floor(ci+ri)))
    else:
        js=[];is_=[];inf=0
        for cidx in range(8):
            sg=np.array([1 if cidx&1 else -1, 1 if cidx&2 else -1, 1 if cidx&4 else -1],float)
            qc=q+row@sg
            if qc[2]>1e-3*SDD:
                m=SDD/qc[2]; js.append((qc[0]*m-o[0])/v[0]); is_.append((qc[1]*m-o[1])/u[1]); inf+=1
        if inf==0: return np.zeros(H*W)
        if inf==8:
            j0=max(0,int(np.ceil(min(js)-1e-3))); j1=min(W-1,int(np.floor(max(js)+1e-3))); i0=max(0,int(np.ceil(min(is_)-1e-3))); i1=min(H-1,int(np.floor(max(is_)+1e-3)))
    out=np.zeros(H*W)
    for i in range(i0,i1+1):
        for j in range(j0,j1+1):
            n=i*W+j
            if not span[n]!=0: continue
            r=1.0/d[n]
            e=pv-src
            x0=(e-1)*r; x1=(e+1)*r
            alo=np.minimum(x0,x1).max(); ahi=np.maximum(x0,x1).min()
            if not alo<ahi: continue
            sc=(NP-1)/span[n]
            k0=(alo-amin[b][n])*sc; k1=(ahi-amin[b][n])*sc
            klo=max(min(k0,k1)-0.02,0.0); khi=min(max(k0,k1)+0.02,NP-1.0)
            if not klo<=khi: continue
            ks=np.arange(int(np.ceil(klo)),int(np.floor(khi))+1)
            if len(ks)==0: continue
            p_=pos[n,ks]
            dd=np.abs(p_-pv)
            ok=(dd<1).all(1)
            out[n]=((1-dd[ok]).prod(1)).sum()
    return out
rng=np.random.default_rng(1)
c=np.clip(np.round(src).astype(int),0,N-1)
vox=[np.clip(c+np.array([dx,dy,dz]),0,N-1) for dx in range(-4,5,2) for dy in range(-10,11,2) for dz in range(-4,5,2)]
vox+=[rng.integers(0,N,3) for _ in range(200)]
worst=[]
tot_ref=0; tot_err=0
for pv in vox:
    pv=pv.astype(float)
    ref=brute(pv); got=kernel(pv)
    e=np.abs(got-ref).sum(); tot_err+=e; tot_ref+=ref.sum()
    if e>1e-6*max(1,ref.sum()): worst.append((e,ref.sum(),pv.tolist(),float((Gi[:3,:3]@pv+Gi[:3,3])[2]),int((np.abs(got-ref)>1e-9).sum())))
print("pose",b,"src",src,"voxels",len(vox),"sum ref %.3f sum |err| %.5f"%(tot_ref,tot_err),"bad voxels",len(worst))
for w in sorted(worst,reverse=True)[:12]: print("err %.4f of %.3f voxel %s depth %.2f rays wrong %d"%w)
