import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import oracle as orc
from uammd_b200 import synthetic as syn
from uammd_b200.fcm import IBM, Peskin3
cells=(32,32,32); L=(32.0,)*3; h=1.0
N=5000
pos=np.zeros((N,4)); pos[:,:3]=syn.uniform_cloud(N,L,seed=3)[:,:3].astype(np.float64)
pos[::9,:3]*=2.3
val=syn.gaussian_forces(N,seed=4)
nxPad=34
g=orc.make_grid_d(L,cells)
ref=orc.ibm_spread(g,orc.peskin3(h),pos,val,nxPad)
ibm=IBM(Peskin3(h),L,cells,nxPad)
cuda=torch.device('cuda:0')
grid=torch.full((32,32,nxPad,3),7.0,dtype=torch.float64,device=cuda)
ibm.spread(torch.from_numpy(pos).to(cuda),torch.from_numpy(val).to(cuda),grid,overwrite=True)
sp=grid.cpu().numpy()
d=np.abs(sp[:,:,:32]-ref[:,:,:32]).max(axis=3)
bad=np.argwhere(d>1e-10)
print("bad nodes",len(bad),"of",32**3)
print("z hist",np.bincount(bad[:,0],minlength=32))
print("y hist",np.bincount(bad[:,1],minlength=32))
print("x hist",np.bincount(bad[:,2],minlength=32))
# single-particle test
for p in [(0.3,0.2,0.1),(-15.9,3.3,4.4),(15.9,-15.9,15.9),(5.5,15.7,-15.8)]:
    pos1=np.zeros((1,4)); pos1[0,:3]=p
    v1=np.ones((1,3))
    r1=orc.ibm_spread(g,orc.peskin3(h),pos1,v1,nxPad)
    g1=torch.full((32,32,nxPad,3),7.0,dtype=torch.float64,device=cuda)
    ibm.spread(torch.from_numpy(pos1).to(cuda),torch.from_numpy(v1).to(cuda),g1,overwrite=True)
    dd=np.abs(g1.cpu().numpy()[:,:,:32]-r1[:,:,:32]).max(axis=3)
    print(p,"bad",np.argwhere(dd>1e-12).tolist()[:10], "ref nz", np.argwhere(np.abs(r1[:,:,:32]).max(axis=3)>0).tolist()[:3])
