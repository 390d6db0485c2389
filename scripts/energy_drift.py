"""Diagnostic: total energy and momentum of a long NVE run (N = 55 296 LJ liquid from FCC) for engine variants."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, CellList, LJ, LJMD, PairForces

dev = torch.device("cuda:0")
n, dt = 24, 0.005
N = 4 * n ** 3
Lb = syn.lj_box_length(N, 0.8)
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
variants = {"column": {}, "column_ldg": {"UB200_LJ_STAGE": "ldg"}, "column_nowiden": {"UB200_LJ_WIDEN": "0"}, "cell": {"UB200_LJ_ENGINE": "cell"}}
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for name, env in variants.items():
    for k in ("UB200_LJ_STAGE", "UB200_LJ_WIDEN", "UB200_LJ_ENGINE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    md = LJMD(Box(Lb), pot, dt)
    p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
    pf = PairForces(pot, Box(Lb), nl=CellList())
    rows = []
    for c in range(10000 // chunk + 1):
        e = torch.zeros(N, device=dev)
        pf.sum(p, energy=e)
        torch.cuda.synchronize()
        ke = 0.5 * float((v.double() ** 2).sum()) / N
        pe = float(e.double().sum()) / N
        mom = float(v.double().sum(0).abs().max()) / N
        fsum = float(f[:, :3].double().sum(0).abs().max())
        rows.append([c * chunk, round(ke + pe, 5), round(2 * ke / 3, 4), f"{mom:.1e}", f"{fsum:.1e}"])
        if c < 10000 // chunk:
            md.run(p, v, f, chunk)
    print(name, json.dumps(rows[::max(1, len(rows) // 11)]))
