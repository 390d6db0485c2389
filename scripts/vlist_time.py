"""Verlet-list path at the BASELINE config-1 shape (N = 1e6 liquid): time of a rebuild, of the per-step position refresh +
traversal, and of the whole MD step (VerletNVE + PairForces<LJ, VerletList>), for the row list and for the reference-layout
list (UB200_VERLET_FAST=0)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, LJMD, PairForces, VerletList
dev = torch.device("cuda:0")
N = 1_000_000
N = 4 * round((N / 4) ** (1 / 3)) ** 3
Lb = syn.lj_box_length(N, 0.8)
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
p0, v0, f0 = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
LJMD(Box(Lb), pot, 0.005).run(p0, v0, f0, 300)
fref = torch.zeros(N, 4, device=dev)
PairForces(pot, Box(Lb)).sum(p0, fref)


def timed(fn, n=10):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {"N": N}
for fast in ("1", "0"):
    os.environ["UB200_VERLET_FAST"] = fast
    p, v, f = p0.clone(), v0.clone(), torch.zeros(N, 4, device=dev)
    nl = VerletList()
    pf = PairForces(pot, Box(Lb), nl=nl)
    pf.sum(p, f)
    err = (f - fref)[:, :3].abs().max().item() / fref[:, :3].abs().max().item()

    def rebuild():
        nl.forceNextUpdate = True
        nl.update(p, Box(Lb), 2.5)
    r = {"force_err_vs_engine": err, "rebuild_ms": timed(rebuild), "update_no_rebuild_ms": timed(lambda: nl.update(p, Box(Lb), 2.5)),
         "traversal_ms": timed(lambda: pf.sumWithCurrentList(f), 20)}
    md = LJMD(Box(Lb), pot, 0.005)
    md.runVerlet(nl, p, v, f, 50)
    torch.cuda.synchronize()
    reb0 = nl.rebuilds()
    steps = 200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); md.runVerlet(nl, p, v, f, steps, forcesAreCurrent=True); b.record(); torch.cuda.synchronize()
    r["md_ms_per_step_back_to_back"] = a.elapsed_time(b) / steps
    r["rebuilds_per_step"] = (nl.rebuilds() - reb0) / steps
    if fast == "1":
        rows = nl.getRowList()
        r["stride"] = rows["stride"]; r["mean_neighbours"] = float(rows["count"].float().mean()); r["max_neighbours"] = int(rows["count"].max())
    out["rows" if fast == "1" else "reference_layout"] = r
print(json.dumps(out))
