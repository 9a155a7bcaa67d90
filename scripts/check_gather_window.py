"""CPU brute-force check of the pixel window the gather kernel (csrc/volgrad.cu, volume_grad_kernel) gives a voxel,
on the edge-pose set (rays missing / grazing the volume, source inside it): every ray that has a sample inside the
voxel's +-1 support must lie inside the window.

Result at the end of round 1 (1 500 voxels, 219 of them within 8 support-depths of the source plane): 0 rays outside
their window with the current rule.  With the rule the kernel had when it measured 6 % off on the B200 -- voxels at
or behind the source plane skipped outright -- 1 024 ray hits are lost: with the source inside the volume every ray's
first sample (alpha = 0, the source itself) lands in cells half of whose corners lie behind the source plane.
    PYTHONPATH=. python scripts/check_gather_window.py
"""
import numpy as np, torch, oracle
from tests._scene import pixel_size
from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
N=64; H=W=32; SDD=1020.0; NP=500
sp=256.0/N
aff=torch.diag(torch.tensor([sp,sp,sp,1.0])); aff[:3,3]=-sp*(N-1)/2
affinv=torch.linalg.inv(aff)
pose=oracle.pose_from_params(torch.tensor(EDGE_ROT),torch.tensor(EDGE_XYZ),"euler_angles","ZXY")
reo=oracle.REORIENT["AP"]
cam2world=oracle.compose(reo[None],pose)           # (B,4,4)
cam2vox=(affinv[None]@cam2world).double().numpy()
vox2cam=np.linalg.inv(cam2vox)
delx=pixel_size(H)
# detector basis as the product builds it: point(i,j) = o + i*u + j*v in camera frame
grid=oracle.detector_grid(H,W,delx,delx,0.0,0.0,SDD,False).view(H,W,3).double().numpy()
o=grid[0,0]; u=grid[1,0]-grid[0,0]; v=grid[0,1]-grid[0,0]
print("o,u,v",o,u,v)
s_,t_=oracle.detector_rays(pose,reo,H,W,delx,delx,0.0,0.0,SDD,False)
s_=oracle.apply(affinv[None],s_).double().numpy(); t_=oracle.apply(affinv[None],t_).double().numpy()
amin,amax=oracle.alpha_minmax(torch.as_tensor(s_,dtype=torch.float32),torch.as_tensor(t_,dtype=torch.float32),torch.tensor([N-1.0]*3),1e-8)
amin=amin.double().numpy()[...,0]; amax=amax.double().numpy()[...,0]
uu=np.linspace(0,1,NP)
rng=np.random.default_rng(0)
bad=0; tested=0; near_cnt=0; full_cnt=0
for b in range(len(EDGE_ROT)):
    Gi=vox2cam[b]
    src=s_[b,0]
    pos=src+(amin[b][:,None]+uu[None]*(amax[b]-amin[b])[:,None])[...,None]*(t_[b]-src)[:,None,:]   # (N,NP,3)
    valid=(amax[b]>amin[b])
    # voxels: random + those near the source
    vox=[rng.integers(0,N,3) for _ in range(150)]
    c=np.clip(np.round(src).astype(int),0,N-1)
    vox+= [np.clip(c+rng.integers(-3,4,3),0,N-1) for _ in range(150)]
    for pv in vox:
        pv=pv.astype(float)
        row=Gi[:3,:3]; q=row@pv+Gi[:3,3]
        rz=np.abs(row[2]).sum()
        j0,j1,i0,i1=0,W-1,0,H-1
        skip=False
        if False:  # the linearised window was dropped: the projected-corner box is used for every voxel
            m=SDD/q[2]; cj=(q[0]*m-o[0])/v[0]; ci=(q[1]*m-o[1])/u[1]
            rj=sum(abs(row[0][a]-(q[0]/q[2])*row[2][a]) for a in range(3))*m*abs(1/v[0])*1.01+1e-3
            ri=sum(abs(row[1][a]-(q[1]/q[2])*row[2][a]) for a in range(3))*m*abs(1/u[1])*1.01+1e-3
            j0=max(0,int(np.ceil(cj-rj))); j1=min(W-1,int(np.floor(cj+rj))); i0=max(0,int(np.ceil(ci-ri))); i1=min(H-1,int(np.floor(ci+ri)))
        else:
            near_cnt+=1
            js=[];is_=[];inf=0
            for cidx in range(8):
                sg=np.array([1 if cidx&1 else -1, 1 if cidx&2 else -1, 1 if cidx&4 else -1],float)
                qc=q+row@sg
                if qc[2]>1e-3*SDD:
                    m=SDD/qc[2]; js.append((qc[0]*m-o[0])/v[0]); is_.append((qc[1]*m-o[1])/u[1]); inf+=1
            if inf==0: skip=True
            elif inf==8:
                j0=max(0,int(np.ceil(min(js)-1e-3))); j1=min(W-1,int(np.floor(max(js)+1e-3))); i0=max(0,int(np.ceil(min(is_)-1e-3))); i1=min(H-1,int(np.floor(max(is_)+1e-3)))
            else: full_cnt+=1
        # brute force: rays with a sample strictly within the +-1 support
        dist=np.abs(pos-pv).max(-1)    # (N,NP)
        hit=((dist<1.0).any(1))&valid
        idx=np.where(hit)[0]
        tested+=1
        for n in idx:
            i,j=divmod(n,W)
            if skip or not (i0<=i<=i1 and j0<=j<=j1):
                bad+=1
print("voxels tested",tested,"near-branch",near_cnt,"full-window",full_cnt,"rays outside their window",bad)
