"""Run a few LJ MD steps of BASELINE config 1 (used under ncu)."""
import sys; sys.path.insert(0, '.')
import torch
from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, LJMD
N = 1_000_000
Lb = syn.lj_box_length(N)
dev = torch.device('cuda:0')
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
p = torch.from_numpy(syn.fcc_lattice(N, Lb)).to(dev); v = torch.from_numpy(syn.maxwell_velocities(N, 1.0, seed=7)).to(dev)
f = torch.zeros(N, 4, device=dev)
md = LJMD(Box(Lb), pot, 0.005)
md.run(p, v, f, int(sys.argv[1]) if len(sys.argv) > 1 else 10)
torch.cuda.synchronize()
print("ok")
