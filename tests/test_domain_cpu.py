"""CPU: the brick domain decomposition with ghost-cell halo exchange (uammd_b200/domain.py) driven by the checker
engines of tests/_checker_engines.py. Virtual ranks stepped in lock step in one process, and two real ranks over
gloo, must reproduce the single-process trajectory BIT FOR BIT (each cell is wholly owned or wholly ghost and keeps
the single-process order)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, "tests")) if p not in sys.path]  # also in the spawned workers


def _lj_setup(N):
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, LJ
    Lb = syn.lj_box_length(N)
    pos = syn.fcc_lattice(N, Lb)
    pos[:, :3] += np.random.default_rng(5).normal(0, 0.05, (N, 3)).astype(np.float32)
    vel = syn.maxwell_velocities(N, 1.5)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    return Box(Lb), pot, pos, vel


def _dpd_setup(N):
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, DPD
    L = (N / 3.0) ** (1.0 / 3.0)
    return (Box(L), DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=99),
            syn.uniform_cloud(N, L, seed=21), syn.maxwell_velocities(N, 1.0, seed=22))


def test_brick_ranges_and_rank_grid():
    from uammd_b200.domain import BrickDecomposition, choose_rank_grid
    assert choose_rank_grid(8, (43, 43, 43)) == (2, 2, 2)
    assert choose_rank_grid(2, (43, 43, 43)) in ((1, 1, 2), (1, 2, 1), (2, 1, 1))
    assert choose_rank_grid(4, (10, 10, 3))[2] == 1           # cut the long dimensions
    d = BrickDecomposition((43, 43, 7), (2, 2, 1), 3)
    assert d.coords == (1, 1, 0) and d.cellRange() == ((21, 43), (21, 43), (0, 7))
    for n, p in ((43, 2), (43, 3), (7, 7), (110, 8), (5, 2)):
        owners = [BrickDecomposition.brickOfCell(c, n, p) for c in range(n)]
        for k in range(p):   # brick k holds exactly [floor(k n/p), floor((k+1) n/p))
            assert [c for c in range(n) if owners[c] == k] == list(range((k * n) // p, ((k + 1) * n) // p))
    with pytest.raises(Exception):
        BrickDecomposition((3, 3, 3), (4, 1, 1), 0)          # a brick without a cell
    with pytest.raises(Exception):
        choose_rank_grid(64, (3, 3, 3))


def test_classify_owner_and_ghost_mask_properties(orc):
    """Every particle has one owner; rank r's ghost set is exactly the particles of the cells adjacent to r's cells."""
    from uammd_b200.domain import BrickDecomposition
    box, pot, pos, _ = _dpd_setup(4000)
    cd = orc.neighbour_celldim(box.boxSize, 1.0)
    g = orc.make_grid_f(box.boxSize, cd)
    rg = (2, 2, 2)
    cell, owner, mask = orc.brick_classify(g, pos, rg)
    cells = orc.get_cells(g, pos)
    assert np.array_equal(cell, cells[:, 0] + cd[0] * (cells[:, 1] + cd[1] * cells[:, 2]))
    for r in range(8):
        rng = BrickDecomposition(cd, rg, r).cellRange()
        inside = np.ones(len(pos), bool)
        near = np.ones(len(pos), bool)
        for d in range(3):
            lo, hi = rng[d]
            c = cells[:, d]
            inside &= (c >= lo) & (c < hi)
            nd = np.zeros(len(pos), bool)
            for o in (-1, 0, 1):
                cc = (c + o) % cd[d]
                nd |= (cc >= lo) & (cc < hi)
            near &= nd
        assert np.array_equal(owner == r, inside)
        assert np.array_equal((mask >> r) & 1 == 1, near & ~inside)


@pytest.mark.parametrize("rankGrid", [(2, 1, 1), (2, 2, 2), (1, 3, 2)])
def test_lj_virtual_ranks_reproduce_single_process(orc, rankGrid):
    from _checker_engines import OracleLJBrickEngine, gather_lockstep, lockstep
    from uammd_b200 import synthetic as syn
    from uammd_b200.domain import DomainDecomposedMD
    N, steps, dt = 4 * 10 ** 3, 5, 0.004
    box, pot, pos, vel = _lj_setup(N)
    world = rankGrid[0] * rankGrid[1] * rankGrid[2]
    ranks = [DomainDecomposedMD(OracleLJBrickEngine(box, pot, dt), N, r, world, rankGrid) for r in range(world)]
    for r in ranks:
        r.setGlobalState(torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy()))
    assert sum(r.nOwned for r in ranks) == N
    for _ in range(steps):
        lockstep(ranks)
    gp, gv = gather_lockstep(ranks, N)
    ref = orc.MDOracle(box.boxSize, 2.5, syn.lj_params(), dt, pos, vel)
    ref.step(steps)
    assert np.array_equal(gp.view(np.uint32), ref.pos.view(np.uint32))
    assert np.array_equal(gv.view(np.uint32), ref.vel.view(np.uint32))
    assert all(r.stats["ghosts"] > 0 for r in ranks)
    assert all(r.pos.shape[0] < N for r in ranks), "a rank must not hold the whole system"


def test_dpd_virtual_ranks_reproduce_single_process(orc):
    """DPD with the noise keyed on global ids; enough steps at rho = 3 for particles to migrate between bricks."""
    from _checker_engines import OracleDPDBrickEngine, OracleDPDEngine, gather_lockstep, lockstep
    from uammd_b200.domain import DomainDecomposedMD
    from uammd_b200.multigpu import DistributedDPDMD
    N, steps, dt = 3000, 12, 0.01
    rankGrid = (2, 2, 1)
    box, pot, pos, vel = _dpd_setup(N)
    pots = [_dpd_setup(N)[1] for _ in range(4)]
    ranks = [DomainDecomposedMD(OracleDPDBrickEngine(box, pots[r], dt), N, r, 4, rankGrid) for r in range(4)]
    for r in ranks:
        r.setGlobalState(torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy()))
    for _ in range(steps):
        lockstep(ranks)
    gp, gv = gather_lockstep(ranks, N)
    p, v, f = torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy()), torch.zeros(N, 4)
    single = DistributedDPDMD(box, pot, dt, N, engine=OracleDPDEngine(box, pot, dt, N))
    for _ in range(steps):
        single.forwardTime(p, v, f)
    assert np.array_equal(gp.view(np.uint32), p.numpy().view(np.uint32))
    assert np.array_equal(gv.view(np.uint32), v.numpy().view(np.uint32))
    assert sum(r.stats["migrated"] for r in ranks) > 0, "the test must exercise migration"


def _gloo_worker(rank, world, port, N, steps, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _checker_engines import OracleDPDBrickEngine
    from uammd_b200.domain import DomainDecomposedMD, TorchFabric
    box, pot, pos, vel = _dpd_setup(N)
    md = DomainDecomposedMD(OracleDPDBrickEngine(box, pot, 0.01), N, rank, world, fabric=TorchFabric())
    md.setGlobalState(torch.from_numpy(pos), torch.from_numpy(vel))
    for _ in range(steps):
        md.forwardTime()
    gp, gv = md.gatherGlobalState()
    if rank == 1:
        np.save(out, np.concatenate([gp.numpy().ravel(), gv.numpy().ravel()]))
    dist.destroy_process_group()


def test_dpd_two_gloo_ranks_with_halo_exchange(tmp_path, orc):
    from _checker_engines import OracleDPDEngine
    from uammd_b200.multigpu import DistributedDPDMD
    N, steps = 3000, 4
    out = str(tmp_path / "brick.npy")
    mp.spawn(_gloo_worker, args=(2, 29527, N, steps, out), nprocs=2, join=True)
    got = np.load(out)
    box, pot, pos, vel = _dpd_setup(N)
    p, v, f = torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy()), torch.zeros(N, 4)
    single = DistributedDPDMD(box, pot, 0.01, N, engine=OracleDPDEngine(box, pot, 0.01, N))
    for _ in range(steps):
        single.forwardTime(p, v, f)
    assert np.array_equal(got[:4 * N].view(np.uint32), p.numpy().ravel().view(np.uint32))
    assert np.array_equal(got[4 * N:].view(np.uint32), v.numpy().ravel().view(np.uint32))


@pytest.mark.parametrize("periodic,L,rankGrid", [((1, 1, 0), (12.0, 12.0, 9.0), (2, 2, 1)),     # open z: no wrap across it
                                                  ((1, 1, 1), (12.0, 6.0, 2.9), (2, 2, 1)),      # a one-cell dimension
                                                  ((0, 0, 0), (10.0, 10.0, 10.0), (2, 1, 2))])
def test_classify_against_brute_force_neighbourhoods(orc, periodic, L, rankGrid):
    """Independent statement of the ghost rule: rank r needs particle i iff some particle position in r's bricks could sit in
    a cell the traversal visits from i's cell, i.e. iff a cell of the reference's neighbour stencil around cell(i) (wrap only in
    periodic dimensions, Grid::pbc_cell) belongs to r."""
    from uammd_b200.domain import BrickDecomposition
    rng = np.random.default_rng(4)
    N = 1500
    pos = np.zeros((N, 4), np.float32)
    pos[:, :3] = (rng.random((N, 3)) - 0.5) * np.array(L)
    cd = orc.neighbour_celldim(L, 2.5)
    g = orc.make_grid_f(L, cd, periodic)
    _, owner, mask = orc.brick_classify(g, pos, rankGrid)
    cells = orc.get_cells(g, pos)
    world = rankGrid[0] * rankGrid[1] * rankGrid[2]
    own_of = lambda c: sum(BrickDecomposition.brickOfCell(int(c[d]), cd[d], rankGrid[d]) * m
                           for d, m in enumerate((1, rankGrid[0], rankGrid[0] * rankGrid[1])))
    for i in range(0, N, 7):
        want = set()
        for o in np.ndindex(3, 3, 3):
            c = cells[i] + np.array(o) - 1
            ok = True
            for d in range(3):
                if c[d] < 0 or c[d] >= cd[d]:
                    if periodic[d]:
                        c[d] %= cd[d]
                    else:
                        ok = False
            if ok:
                want.add(own_of(c))
        assert owner[i] == own_of(cells[i])
        want.discard(int(owner[i]))
        assert {r for r in range(world) if (int(mask[i]) >> r) & 1} == want


def _solo_group_worker(rank, world, port, N, steps, out):
    """The flow of scripts/brick_dpd.py: brick run on all ranks, then a single-rank reference run on rank 0 INSIDE the same
    process group - which must use a one-rank subgroup, or its all-gathers wait for ranks that never join."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    solo = dist.new_group(ranks=[0])
    from _checker_engines import OracleDPDBrickEngine, OracleDPDEngine
    from uammd_b200.domain import DomainDecomposedMD, TorchFabric
    from uammd_b200.multigpu import DistributedDPDMD
    box, pot, pos, vel = _dpd_setup(N)
    md = DomainDecomposedMD(OracleDPDBrickEngine(box, pot, 0.01), N, rank, world, fabric=TorchFabric())
    md.setGlobalState(torch.from_numpy(pos), torch.from_numpy(vel))
    for _ in range(steps):
        md.forwardTime()
    gp, gv = md.gatherGlobalState()
    if rank == 0:
        box, pot, pos, vel = _dpd_setup(N)
        p, v, f = torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy()), torch.zeros(N, 4)
        single = DistributedDPDMD(box, pot, 0.01, N, engine=OracleDPDEngine(box, pot, 0.01, N), group=solo)
        assert single.world == 1
        for _ in range(steps):
            single.forwardTime(p, v, f)
        same = torch.equal(gp.view(torch.int32), p.view(torch.int32)) and torch.equal(gv.view(torch.int32), v.view(torch.int32))
        open(out, "w").write("same" if same else "different")
    dist.destroy_process_group()


def test_single_rank_check_inside_a_two_rank_group_does_not_hang(tmp_path):
    out = str(tmp_path / "solo.txt")
    mp.spawn(_solo_group_worker, args=(2, 29529, 2000, 2, out), nprocs=2, join=True)
    assert open(out).read() == "same"
