"""TEST INFRASTRUCTURE: CPU checker engines (oracle arithmetic) injected into the host-side multi-GPU logic of
uammd_b200/multigpu.py and uammd_b200/domain.py so that the decompositions are covered by gloo / lock-step tests
without a GPU. Nothing under uammd_b200/ imports this module."""
import numpy as np
import torch

from oracle import oracle as orc


class OracleLJEngine:
    """Engine interface of multigpu.DistributedLJMD (replicated positions, block-owned particles)."""

    def __init__(self, box, pot, dt):
        self.box, self.pot, self.dt = box, pot, float(dt)

    def half(self, step, pos_blk, vel_blk, force_blk):
        orc.nve_half(pos_blk.numpy(), vel_blk.numpy(), force_blk.numpy(), self.dt, 1.0, step)

    def kick_kick_drift(self, pos_blk, vel_blk, force_blk):
        self.half(2, pos_blk, vel_blk, force_blk)
        self.half(1, pos_blk, vel_blk, force_blk)

    def forces_owned(self, pos, force, lo, hi):
        L = self.box.boxSize
        g = orc.make_grid_f(L, orc.neighbour_celldim(L, self.pot.getCutOff()))
        cl = orc.celllist_build(g, pos.numpy())
        f, _, _ = orc.lj_f32(g, cl, self.pot.table(), self.pot.ntypes, pos.shape[0])
        force.numpy()[lo:hi] = f[lo:hi]


class OracleDPDEngine:
    """Engine interface of multigpu.DistributedDPDMD."""

    def __init__(self, box, pot, dt, N):
        self.box, self.pot, self.dt, self.N = box, pot, float(dt), N

    def forces_owned(self, pos, vel, force, lo, hi):
        L, p = self.box.boxSize, self.pot
        g = orc.make_grid_f(L, orc.neighbour_celldim(L, p.getCutOff()))
        cl = orc.celllist_build(g, pos.numpy())
        f32, _ = orc.dpd_f32(g, cl, vel.numpy(), p.A, p.gamma, p.sigma, p.rcut, p.seed, p.step & 0xFFFFFFFF, self.N)
        force.numpy()[lo:hi] = f32[lo:hi]

    def half(self, step, pb, vb, fb):
        orc.nve_half(pb.numpy(), vb.numpy(), fb.numpy(), self.dt, 1.0, step)


class _OracleBrickEngine:
    """Engine interface of domain.DomainDecomposedMD: classify / half / forces on CPU tensors."""

    def __init__(self, box, cutOff, dt):
        self.box, self.cutOff, self.dt = box, float(cutOff), float(dt)
        self.cellDim = orc.neighbour_celldim(box.boxSize, cutOff)
        self.grid = orc.make_grid_f(box.boxSize, self.cellDim, [int(p) for p in box.periodic])

    def classify(self, pos, rankGrid):
        _, owner, mask = orc.brick_classify(self.grid, pos.numpy(), rankGrid)
        return torch.from_numpy(owner), torch.from_numpy(mask.view(np.int32))

    def half(self, step, pos, vel, force):
        if pos.shape[0]:
            orc.nve_half(pos.numpy(), vel.numpy(), force.numpy(), self.dt, 1.0, step)


class OracleLJBrickEngine(_OracleBrickEngine):
    ghostNeedsVelocity = False

    def __init__(self, box, pot, dt):
        super().__init__(box, pot.getCutOff(), dt)
        self.pot = pot

    def forces(self, md):
        if md.nOwned == 0:
            return
        cl = orc.celllist_build(self.grid, md.pos.numpy())
        f, _, _ = orc.lj_f32(self.grid, cl, self.pot.table(), self.pot.ntypes, md.pos.shape[0])
        md.force.numpy()[:md.nOwned] = f[:md.nOwned]


class OracleDPDBrickEngine(_OracleBrickEngine):
    ghostNeedsVelocity = True

    def __init__(self, box, pot, dt):
        super().__init__(box, pot.getCutOff(), dt)
        self.pot = pot

    def forces(self, md):
        p = self.pot
        p.step += 1
        if md.nOwned == 0:
            return
        cl = orc.celllist_build(self.grid, md.pos.numpy())
        f = orc.dpd_ids_f32(self.grid, cl, md.vel.numpy(), p.A, p.gamma, p.sigma, p.rcut, p.seed, p.step & 0xFFFFFFFF,
                            md.pos.shape[0], md.gid.numpy(), md.N)
        md.force.numpy()[:md.nOwned] = f[:md.nOwned]


def lockstep(ranks, method="forwardTimeSteps"):
    """Drive the step generators of several virtual ranks of one process in lock step, doing the all-to-all by hand:
    the same code path as DomainDecomposedMD.forwardTime with a TorchFabric, without torch.distributed."""
    world = len(ranks)
    gens = [getattr(r, method)() for r in ranks]
    msgs = [next(g) for g in gens]
    while True:
        # msgs[s] = (rows grouped by destination, counts per destination)
        offs = [np.concatenate([[0], np.cumsum(m[1])]).astype(int) for m in msgs]
        recv = [torch.cat([msgs[s][0][offs[s][d]:offs[s][d + 1]] for s in range(world)]) for d in range(world)]
        nxt, done = [], 0
        for d in range(world):
            try:
                nxt.append(gens[d].send(recv[d]))
            except StopIteration:
                done += 1
        if done:
            assert done == world, "virtual ranks fell out of lock step"
            return
        msgs = nxt


def gather_lockstep(ranks, N):
    """Global pos / vel in id order from the owned blocks of all virtual ranks."""
    pos = torch.zeros(N, 4)
    vel = torch.zeros(N, 3)
    seen = torch.zeros(N, dtype=torch.int32)
    for r in ranks:
        p, v, g, _ = r.owned()
        idx = g.to(torch.int64).cpu()
        pos[idx], vel[idx] = p.cpu(), v.cpu()
        seen[idx] += 1
    assert bool((seen == 1).all()), "every particle must be owned by exactly one rank"
    return pos.numpy(), vel.numpy()
