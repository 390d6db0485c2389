import sys; sys.path.insert(0, '.')
import numpy as np, torch
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, PairForces
from oracle import oracle as orc
N=20000; Lb=32.0
def run(name, pos, pot):
    box=Box(Lb); pf=PairForces(pot, box)
    dpos=torch.from_numpy(pos).cuda(); force=torch.zeros(N,4,device='cuda')
    pf.sum(dpos, force=force); torch.cuda.synchronize()
    g = orc.make_grid_f(box.boxSize, orc.neighbour_celldim(box.boxSize, pot.getCutOff()))
    cl = orc.celllist_build(g, pos)
    f64,e64,v64,a = orc.lj_f64(g, cl, pot.table(), pot.ntypes, N)
    F=force.cpu().numpy()
    err=np.abs(F[:,:3]-f64).max(axis=1)/np.maximum(a,1e-30)
    bad=np.nonzero(err>2e-4)[0]
    cells=orc.get_cells(g, pos[bad]) if len(bad) else np.zeros((0,3))
    print(name, "nbad",len(bad),"maxerr",err.max(), "cells of bad:", [tuple(c) for c in cells[:30]])
    return bad
pos = syn.uniform_cloud(N, Lb, seed=6, ntypes=3)
def potA(fn):
    pot = LJ()
    for a in range(3):
        for b in range(a, 3):
            pot.setPotParameters(a, b, **fn(a,b))
    return pot
run("orig", pos, potA(lambda a,b: dict(cutOff=2.0 + 0.25 * (a + b), sigma=0.9 + 0.1 * a + 0.05 * b, epsilon=1.0 + 0.5 * a * b, shift=(a == b))))
run("sameparams", pos, potA(lambda a,b: dict(cutOff=2.5, sigma=1.0, epsilon=1.0)))
run("samecut3", pos, potA(lambda a,b: dict(cutOff=3.0, sigma=0.9 + 0.1 * a + 0.05 * b, epsilon=1.0 + 0.5 * a * b)))
run("diffcut_only", pos, potA(lambda a,b: dict(cutOff=2.0 + 0.25 * (a + b), sigma=1.0, epsilon=1.0)))
p1=pos.copy(); p1[:,3]=0
pot1=LJ(); pot1.setPotParameters(0,0,cutOff=3.0)
run("single", p1, pot1)
p2=pos.copy(); p2[:,[0,1]]=p2[:,[1,0]]
run("orig_swapxy", p2, potA(lambda a,b: dict(cutOff=2.0 + 0.25 * (a + b), sigma=0.9 + 0.1 * a + 0.05 * b, epsilon=1.0 + 0.5 * a * b, shift=(a == b))))
