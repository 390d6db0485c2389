"""A/B on one B200: LJ traversal kernels at the BASELINE config-1 shape (N = 1e6 liquid).
column traversal with TMA staging / with per-lane row copies / the cell traversal over the reference-layout list;
plus the whole fused MD step. Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn  # noqa: E402
from uammd_b200.md import Box, CellList, LJ, LJEngine, LJMD, PairForces  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
Lb = syn.lj_box_length(N, 0.8)
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
box = Box(Lb)
p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
md = LJMD(box, pot, 0.005)
md.run(p, v, f, 300)
torch.cuda.synchronize()
scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        scrub.fill_(1)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


out = {"N": N}
ref = torch.zeros(N, 4, device=dev)
pfc = PairForces(pot, box, nl=CellList())
pfc.nl.update(p, box, 2.5)
pfc.sumWithCurrentList(force=ref)
out["cell_traversal_ms"] = timeit(lambda: pfc.sumWithCurrentList(force=ref))
out["cell_build_ms"] = timeit(lambda: pfc.nl.update(p, box, 2.5))
ref.zero_(); pfc.sumWithCurrentList(force=ref)
for stage in ("tma", "ldg"):
    os.environ["UB200_LJ_STAGE"] = stage
    eng = LJEngine()
    g = torch.zeros(N, 4, device=dev)
    eng.sum(p, box, pot.table(), 1, force=g, accumulate=False)
    torch.cuda.synchronize()
    out[f"column_{stage}_path"] = eng.lastPath()
    out[f"column_{stage}_err"] = eng.errorFlag()
    out[f"column_{stage}_maxdiff_rel"] = float((g[:, :3] - ref[:, :3]).abs().max() / ref[:, :3].abs().max())
    out[f"column_{stage}_traversal_ms"] = timeit(lambda: eng.traverse(g, accumulate=False))
    out[f"column_{stage}_build_plus_traversal_ms"] = timeit(lambda: eng.sum(p, box, pot.table(), 1, force=g, accumulate=False))
    out[f"md_step_{stage}_ms"] = timeit(lambda: md.run(p, v, f, 1))
os.environ["UB200_LJ_ENGINE"] = "cell"
out["md_step_cell_ms"] = timeit(lambda: md.run(p, v, f, 1))
print(json.dumps(out))
