import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, CellList, LJ, LJMD, LJEngine, PairForces
dev = torch.device("cuda:0")
n, dt = 24, 0.005
N = 4 * n ** 3
Lb = syn.lj_box_length(N, 0.8)
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
box = Box(Lb)
md = LJMD(box, pot, dt)
p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
pfc = PairForces(pot, box, nl=CellList())
for chunk in range(20):
    md.run(p, v, f, 500)
    fc = torch.zeros(N, 4, device=dev); pfc.sum(p, force=fc)
    for stage in ("tma", "ldg"):
        os.environ["UB200_LJ_STAGE"] = stage
        eng = LJEngine()
        fe = torch.zeros(N, 4, device=dev)
        eng.sum(p, box, pot.table(), 1, force=fe)
        torch.cuda.synchronize()
        d = (fe[:, :3] - fc[:, :3]).abs().max(dim=1).values
        fmax = float(fc[:, :3].abs().max())
        bad = torch.nonzero(d > 1e-3 * fmax).flatten()
        print(json.dumps({"step": (chunk + 1) * 500, "stage": stage, "fmax": fmax, "maxdiff": float(d.max()), "nbad": int(bad.numel()),
                          "err": eng.errorFlag(), "pos_absmax": float(p[:, :3].abs().max()),
                          "bad_pos": p[bad[:3]].cpu().numpy().tolist(), "bad_diff": d[bad[:3]].cpu().numpy().tolist(),
                          "f_engine_md_maxdiff": float((f[:, :3] - fe[:, :3]).abs().max())}))
    del os.environ["UB200_LJ_STAGE"]
