"""Brick domain decomposition of the short-range pair path over the GPUs of one box: owned particles + ghost cells,
halo exchange every step (BASELINE config 4: "DPD fluid, ghost-cell halo exchange, domain-decomposed over 8 GPUs";
SURVEY 8(e)). The reference is single-GPU; the oracle of this module is the single-GPU result on the same input.

Decomposition (ub200_brick_classify_f32, uammd_b200/csrc/domain.cu): the reference's neighbour grid (cell size >= cut-off,
CellList.cuh:100-126) is cut into px x py x pz bricks of whole cells. A rank owns the particles whose cell lies in its
brick and receives, as ghosts, the particles of the cells adjacent to its brick. Forces are computed for owned particles
only with the full-neighbour scheme, so nothing is sent back. Every step:
  1. migration   particles whose cell left the brick move to their new owner      {pos real4, vel real3, id}   32 B
  2. halo        particles of boundary cells are copied to the ranks that need them  LJ {pos, id} 20 B, DPD 32 B
  3. forces      cell list over [owned | ghosts] on the GLOBAL grid, traversal restricted to the owned block
                 (ub200_lj_sum_owned_f32 / ub200_dpd_sum_owned_ids_f32), velocity Verlet on the owned block.
Both exchanges are one variable-size all-to-all each (NVSwitch: uniform bandwidth to every peer, so no staged
x/y/z forwarding). The owned block is kept sorted by global id and the ghosts of one cell always come from one
rank in id order, so the particles of every cell appear in the single-GPU order and the forces - hence the whole
trajectory - are BIT-IDENTICAL to the single-GPU run (tests/test_domain_*.py); the DPD noise is keyed on global ids.

The host logic is written against two small interfaces so that it is covered without GPUs:
  engine   classify(pos) -> (owner, ghostMask); forces(md); half(step, pos, vel, force)      [CUDA: C ABI below]
  fabric   exchange(rows, counts) -> (rows, counts)                                          [torch.distributed]
The step is a generator (`forwardTimeSteps`) that yields at every exchange; `forwardTime` drives it with the fabric.
"""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import UB200Error


def choose_rank_grid(world, cellDim):
    """px x py x pz = world, as cubic as possible, larger factors on the dimensions with more cells."""
    best = None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // (px * py)
            g = (px, py, pz)
            if any(g[d] > cellDim[d] for d in range(3)):
                continue
            # surface of one brick in cells (communication volume); ties -> lexicographically smallest grid
            b = [cellDim[d] / g[d] for d in range(3)]
            cost = b[0] * b[1] + b[1] * b[2] + b[0] * b[2]
            if best is None or cost < best[0] - 1e-9:
                best = (cost, g)
    if best is None:
        raise UB200Error(f"cannot split a {cellDim} cell grid over {world} ranks with at least one cell per brick")
    return best[1]


class BrickDecomposition:
    """Which cells a rank owns. rank = kx + px (ky + py kz); brick k of a dimension with n cells and p bricks holds the
    cells [floor(k n / p), floor((k + 1) n / p))."""

    def __init__(self, cellDim, rankGrid, rank):
        self.cellDim = tuple(int(c) for c in cellDim)
        self.rankGrid = tuple(int(p) for p in rankGrid)
        self.world = self.rankGrid[0] * self.rankGrid[1] * self.rankGrid[2]
        if self.world > 32:
            raise UB200Error("brick decomposition: at most 32 ranks (one ghost-mask bit per rank)")
        if any(p < 1 or p > n for p, n in zip(self.rankGrid, self.cellDim)):
            raise UB200Error(f"brick decomposition: rank grid {self.rankGrid} needs at least one cell of {self.cellDim} per brick")
        if not 0 <= rank < self.world:
            raise UB200Error(f"rank {rank} outside the rank grid {self.rankGrid}")
        self.rank = rank
        px, py = self.rankGrid[0], self.rankGrid[1]
        self.coords = (rank % px, (rank // px) % py, rank // (px * py))

    @staticmethod
    def brickOfCell(c, n, p):
        return ((c + 1) * p - 1) // n

    def cellRange(self, rank=None):
        r = self.rank if rank is None else rank
        px, py = self.rankGrid[0], self.rankGrid[1]
        k = (r % px, (r // px) % py, r // (px * py))
        return tuple(((k[d] * self.cellDim[d]) // self.rankGrid[d], ((k[d] + 1) * self.cellDim[d]) // self.rankGrid[d])
                     for d in range(3))


class TorchFabric:
    """Variable-size all-to-all of int32 rows over torch.distributed (NCCL on GPUs, gloo in the CPU tests). The counts
    travel first (one tiny all-to-all + a host read), then the rows."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows_sent = 0

    def exchange(self, rows, counts):
        cnt = torch.tensor(counts, dtype=torch.int64, device=rows.device)
        rcnt = torch.empty_like(cnt)
        dist.all_to_all_single(rcnt, cnt, group=self.group)
        rcounts = rcnt.tolist()
        out = torch.empty((sum(rcounts), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out, rows.contiguous(), output_split_sizes=rcounts, input_split_sizes=list(counts),
                               group=self.group)
        self.rows_sent += int(rows.shape[0])
        return out, rcounts


class _CudaBrickEngine:
    """classify / half through the C ABI; subclasses add the force evaluation."""

    def __init__(self, box, cutOff, dt):
        from . import _lib
        from .md import CellList, _ptr, _stream_ptr
        self.lib, self.check, self._ptr, self._stream = _lib.lib(), _lib.check, _ptr, _stream_ptr
        self.f3, self.i3 = _lib.f3, _lib.i3
        self.box, self.cutOff, self.dt = box, float(cutOff), float(dt)
        self.cellDim = CellList.gridFor(box, cutOff)
        self.nl = CellList()
        vp, i = C.c_void_p, C.c_int
        self.lib.ub200_brick_classify_f32.restype = i
        self.lib.ub200_brick_classify_f32.argtypes = [vp, i, C.c_float * 3, C.c_int * 3, C.c_int * 3, C.c_int * 3, vp, vp, vp, vp]

    def classify(self, pos, rankGrid):
        n = pos.shape[0]
        owner = torch.empty(n, dtype=torch.int32, device=pos.device)
        mask = torch.empty(n, dtype=torch.int32, device=pos.device)
        self.check(self.lib.ub200_brick_classify_f32(self._ptr(pos), n, self.f3(self.box.boxSize),
                                                     self.i3([int(p) for p in self.box.periodic]), self.i3(self.cellDim),
                                                     self.i3(rankGrid), None, self._ptr(owner), self._ptr(mask), self._stream()))
        return owner, mask

    def half(self, step, pos, vel, force):
        if pos.shape[0]:
            self.check(self.lib.ub200_nve_half_step_f32(self._ptr(pos), self._ptr(vel), self._ptr(force), C.c_void_p(0), 1.0,
                                                        C.c_void_p(0), pos.shape[0], self.dt, 0, step, self._stream()))


class CudaLJBrickEngine(_CudaBrickEngine):
    ghostNeedsVelocity = False

    def __init__(self, box, pot, dt):
        super().__init__(box, pot.getCutOff(), dt)
        self.pot = pot
        self.tab = pot.table()
        self.tabp = self.tab.ctypes.data_as(C.POINTER(C.c_float))

    def forces(self, md):
        if md.nOwned == 0:
            return
        self.nl.update(md.pos, self.box, self.cutOff, cellDim=self.cellDim)
        self.check(self.lib.ub200_lj_sum_owned_f32(self.nl._h, self.tabp, self.pot.ntypes, self._ptr(md.force), 0, md.nOwned, 0,
                                                   self._stream()))


class CudaDPDBrickEngine(_CudaBrickEngine):
    ghostNeedsVelocity = True

    def __init__(self, box, pot, dt):
        super().__init__(box, pot.getCutOff(), dt)
        self.pot = pot
        vp, i, f, u32 = C.c_void_p, C.c_int, C.c_float, C.c_uint32
        self.lib.ub200_dpd_sum_owned_ids_f32.restype = i
        self.lib.ub200_dpd_sum_owned_ids_f32.argtypes = [vp, vp, f, f, f, f, u32, u32, i, vp, i, i, i, vp, vp]

    def forces(self, md):
        p = self.pot
        p.step += 1  # DPD_impl::getForceTransverser increments the step before every evaluation (DPD.cuh:161-170)
        if md.nOwned == 0:
            return
        self.nl.update(md.pos, self.box, self.cutOff, cellDim=self.cellDim)
        self.check(self.lib.ub200_dpd_sum_owned_ids_f32(self.nl._h, self._ptr(md.vel), p.A, p.gamma, p.sigma, p.rcut, p.seed,
                                                        p.step & 0xFFFFFFFF, md.N, self._ptr(md.force), 0, md.nOwned, 0,
                                                        self._ptr(md.gid), self._stream()))


def _rows(pos, vel, gid):
    """{pos real4, vel real3 | nothing, id} as int32 rows (bit copies; 8 or 5 words)."""
    parts = [pos.view(torch.int32)]
    if vel is not None:
        parts.append(vel.view(torch.int32))
    parts.append(gid.view(-1, 1))
    return torch.cat(parts, dim=1)


class DomainDecomposedMD:
    """VerletNVE + PairForces<LJ | DPD, CellList> (VerletNVE.cu:174-188, PairForces.cu:43-78) on one brick per rank.

    State of a rank: pos [nLocal,4], vel [nLocal,3], gid [nLocal] with the nOwned owned particles first (sorted by global
    id) and the ghosts after them; force [nLocal,4] is meaningful on the owned block."""

    def __init__(self, engine, N, rank, world, rankGrid=None, fabric=None):
        self.eng, self.N = engine, int(N)
        grid = tuple(rankGrid) if rankGrid is not None else choose_rank_grid(world, engine.cellDim)
        self.dec = BrickDecomposition(engine.cellDim, grid, rank)
        if self.dec.world != world:
            raise UB200Error(f"rank grid {grid} does not match the number of ranks ({world})")
        self.fabric = fabric
        self.steps = 0
        self.nOwned = 0
        self.pos = self.vel = self.gid = self.force = None
        self.stats = {"migrated": 0, "ghosts": 0}

    # -- state ------------------------------------------------------------------------------------------------------
    def setGlobalState(self, pos, vel):
        """Every rank passes the same full arrays (replicated initial condition) and keeps the particles it owns."""
        owner, _ = self.eng.classify(pos, self.dec.rankGrid)
        mine = torch.nonzero(owner == self.dec.rank).flatten()
        self.pos, self.vel = pos[mine].contiguous(), vel[mine].contiguous()
        self.gid = mine.to(torch.int32)
        self.nOwned = int(mine.shape[0])
        self.force = torch.zeros_like(self.pos)
        self.steps = 0

    def owned(self):
        n = self.nOwned
        return self.pos[:n], self.vel[:n], self.gid[:n], self.force[:n]

    # -- one refresh of ownership and ghosts: two exchanges ----------------------------------------------------------
    def _packMigrants(self):
        pos, vel, gid, _ = self.owned()
        owner, mask = self.eng.classify(pos, self.dec.rankGrid)
        leave = owner != self.dec.rank
        nleave = int(leave.sum())
        world = self.dec.world
        if nleave == 0:
            self._kept = (pos, vel, gid, mask)
            return torch.empty((0, 8), dtype=torch.int32, device=pos.device), [0] * world
        li = torch.nonzero(leave).flatten()
        dest = owner[li].to(torch.int64)
        order = torch.argsort(dest, stable=True)
        li = li[order]
        counts = torch.bincount(dest, minlength=world).tolist()
        stay = torch.nonzero(~leave).flatten()
        self._kept = (pos[stay], vel[stay], gid[stay], None)
        self.stats["migrated"] += nleave
        return _rows(pos[li], vel[li], gid[li]), counts

    def _absorbMigrants(self, rows):
        pos, vel, gid, mask = self._kept
        if rows.shape[0]:
            pos = torch.cat([pos, rows[:, 0:4].view(torch.float32)])
            vel = torch.cat([vel, rows[:, 4:7].view(torch.float32)])
            gid = torch.cat([gid, rows[:, 7]])
            mask = None
        if mask is None:  # the owned set changed: restore the id order, classify again for the ghost masks
            order = torch.argsort(gid.to(torch.int64))
            pos, vel, gid = pos[order].contiguous(), vel[order].contiguous(), gid[order].contiguous()
            _, mask = self.eng.classify(pos, self.dec.rankGrid)
        self._kept = (pos, vel, gid, mask)

    def _packGhosts(self):
        pos, vel, gid, mask = self._kept
        world = self.dec.world
        cand = torch.nonzero(mask != 0).flatten()
        if cand.shape[0] == 0:
            w = 8 if self.eng.ghostNeedsVelocity else 5
            return torch.empty((0, w), dtype=torch.int32, device=pos.device), [0] * world
        bits = (mask[cand].view(-1, 1) >> torch.arange(world, dtype=torch.int32, device=pos.device).view(1, -1)) & 1
        dr, di = torch.nonzero(bits.t().contiguous(), as_tuple=True)  # destination-major, then ascending id
        counts = torch.bincount(dr, minlength=world).tolist()
        src = cand[di]
        self.stats["ghosts"] += int(src.shape[0])
        return _rows(pos[src], vel[src] if self.eng.ghostNeedsVelocity else None, gid[src]), counts

    def _absorbGhosts(self, rows):
        pos, vel, gid, _ = self._kept
        self._kept = None
        self.nOwned = int(pos.shape[0])
        ng = int(rows.shape[0])
        if ng:
            gpos = rows[:, 0:4].view(torch.float32)
            if self.eng.ghostNeedsVelocity:
                gvel, ggid = rows[:, 4:7].view(torch.float32), rows[:, 7]
            else:
                gvel, ggid = torch.zeros((ng, 3), dtype=torch.float32, device=pos.device), rows[:, 4]
            pos, vel, gid = torch.cat([pos, gpos]), torch.cat([vel, gvel]), torch.cat([gid, ggid])
        self.pos, self.vel, self.gid = pos.contiguous(), vel.contiguous(), gid.contiguous()
        self.force = torch.zeros_like(self.pos)

    def _refreshSteps(self):
        rows = yield self._packMigrants()
        self._absorbMigrants(rows)
        rows = yield self._packGhosts()
        self._absorbGhosts(rows)

    # -- the integrator ----------------------------------------------------------------------------------------------
    def forwardTimeSteps(self):
        """VerletNVE::forwardTime as a generator: yields (rows, counts) at every exchange and expects the received rows."""
        self.steps += 1
        if self.steps == 1:
            yield from self._refreshSteps()
            self.eng.forces(self)
        p, v, _, f = self.owned()
        self.eng.half(1, p, v, f)
        yield from self._refreshSteps()
        self.eng.forces(self)
        p, v, _, f = self.owned()
        self.eng.half(2, p, v, f)

    def forwardTime(self):
        if self.fabric is None:
            raise UB200Error("DomainDecomposedMD.forwardTime needs a fabric (TorchFabric)")
        gen = self.forwardTimeSteps()
        try:
            msg = next(gen)
            while True:
                rows, _ = self.fabric.exchange(*msg)
                msg = gen.send(rows)
        except StopIteration:
            pass

    def gatherGlobalState(self):
        """Full pos [N,4] / vel [N,3] in global-id order on every rank (diagnostics and tests; not on the step path)."""
        pos, vel, gid, _ = self.owned()
        rows = _rows(pos, vel, gid)
        world = self.dec.world
        out, _ = self.fabric.exchange(rows.repeat(world, 1), [rows.shape[0]] * world)
        order = torch.argsort(out[:, 7].to(torch.int64))
        out = out[order]
        return out[:, 0:4].contiguous().view(torch.float32), out[:, 4:7].contiguous().view(torch.float32)


def make_lj(box, pot, dt, N, rankGrid=None, group=None):
    """One rank of the brick-decomposed LJ MD (CUDA engine, torch.distributed fabric)."""
    return DomainDecomposedMD(CudaLJBrickEngine(box, pot, dt), N, dist.get_rank(group), dist.get_world_size(group), rankGrid,
                              TorchFabric(group))


def make_dpd(box, pot, dt, N, rankGrid=None, group=None):
    """One rank of the brick-decomposed DPD fluid (BASELINE config 4 shape)."""
    return DomainDecomposedMD(CudaDPDBrickEngine(box, pot, dt), N, dist.get_rank(group), dist.get_world_size(group), rankGrid,
                              TorchFabric(group))
