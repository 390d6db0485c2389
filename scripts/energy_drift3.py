import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, CellList, LJ, LJMD, PairForces
dev = torch.device("cuda:0")
n, dt = 24, 0.005
N = 4 * n ** 3
Lb = syn.lj_box_length(N, 0.8)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
box = Box(Lb)
pf = PairForces(pot, box, nl=CellList())
def energy(p, v):
    e = torch.zeros(N, device=dev); pf.sum(p, energy=e); torch.cuda.synchronize()
    ke = 0.5 * float((v.double() ** 2).sum()) / N
    return round(ke + float(e.double().sum()) / N, 5), round(2 * ke / 3, 4)
# (a) stable liquid: hotter start
for name, env in (("column", {}), ("cell", {"UB200_LJ_ENGINE": "cell"})):
    os.environ.pop("UB200_LJ_ENGINE", None); os.environ.update(env)
    pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 2.2, seed=7)
    md = LJMD(box, pot, dt)
    p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
    rows = []
    for c in range(11):
        rows.append((c * 2000,) + energy(p, v))
        if c < 10: md.run(p, v, f, 2000)
    print("hot", name, json.dumps(rows))
# (b) cold start, cell engine for 30000 steps: does it make the same transition later?
os.environ["UB200_LJ_ENGINE"] = "cell"
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
md = LJMD(box, pot, dt)
p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
rows = []
for c in range(13):
    rows.append((c * 2500,) + energy(p, v))
    if c < 12: md.run(p, v, f, 2500)
print("cold cell 30000", json.dumps(rows))
# (c) cold start, column engine: forces against the cell traversal at EVERY step from 5000 to 8000
os.environ.pop("UB200_LJ_ENGINE", None)
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
md = LJMD(box, pot, dt)
p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
md.run(p, v, f, 5000)
worst, worst_step = 0.0, -1
for s in range(3000):
    md.run(p, v, f, 1)
    fc = torch.zeros(N, 4, device=dev); pf.sum(p, force=fc)
    d = float((f[:, :3] - fc[:, :3]).abs().max() / fc[:, :3].abs().max())
    if d > worst: worst, worst_step = d, 5001 + s
print("cold column every step 5000..8000: worst rel force diff", worst, "at", worst_step, "energy now", energy(p, v))
