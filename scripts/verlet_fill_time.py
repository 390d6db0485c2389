"""Time of one Verlet-list rebuild (cell list over rc x 1.08 + verletFill) at the BASELINE config-1 shape, N = 1e6 liquid."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, LJMD, VerletList
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Lb = syn.lj_box_length(N, 0.8)
N = 4 * round((N / 4) ** (1 / 3)) ** 3
Lb = syn.lj_box_length(N, 0.8)
pos, vel = syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
LJMD(Box(Lb), pot, 0.005).run(p, v, f, 300)
nl = VerletList()
nl.update(p, Box(Lb), 2.5)
ts = []
for _ in range(10):
    nl.forceNextUpdate = True
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); nl.update(p, Box(Lb), 2.5); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
d = nl.getVerletList()
print(json.dumps({"N": N, "rebuild_ms_median": float(np.median(ts)), "maxNeighboursPerParticle": d["maxNeighboursPerParticle"],
                  "mean_neighbours": float(d["numberNeighbours"].float().mean())}))
