"""GPU: cell list built by the CUDA path through the C ABI must be BIT-EXACT against the oracle
(sortPos, groupIndex = stable Morton order, cellStart/cellEnd)."""
import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, CellList

pytestmark = pytest.mark.gpu


def _check_against_oracle(orc, cuda, pos, L, cellDim=None, rc=2.5, periodic=(1, 1, 1), rebuilds=1):
    box = Box(L)
    box.setPeriodicity(*periodic)
    cd = cellDim if cellDim is not None else orc.neighbour_celldim(box.boxSize, rc)
    g = orc.make_grid_f(box.boxSize, cd, periodic)
    ref = orc.celllist_build(g, pos)
    cl = CellList()
    dpos = torch.from_numpy(pos).to(cuda)
    for _ in range(rebuilds):
        cl.update(dpos, box, rc, cellDim=cd)
    torch.cuda.synchronize()
    d = cl.getCellList()
    assert np.array_equal(d["groupIndex"].cpu().numpy(), ref["index"])
    assert np.array_equal(d["sortPos"].cpu().numpy().view(np.uint32), ref["sortPos"].view(np.uint32))
    cs, ce = cl.normalizedCells()
    assert np.array_equal(cs, ref["cellStart"]) and np.array_equal(ce, ref["cellEnd"])
    assert cl.errorFlag() == ref["error"]
    return cl


@pytest.mark.parametrize("N,rho", [(1, 0.01), (33, 0.3), (1000, 0.8), (20000, 0.8), (200000, 0.8)])
def test_uniform_cloud_bit_exact(orc, cuda, N, rho):
    Lb = max(syn.lj_box_length(N, rho), 10.5)
    _check_against_oracle(orc, cuda, syn.uniform_cloud(N, Lb, seed=N), (Lb,) * 3)


def test_non_cubic_box_and_explicit_grid(orc, cuda):
    L = (64.0, 32.0, 7.0)
    _check_against_oracle(orc, cuda, syn.uniform_cloud(50000, L, seed=5), L, cellDim=(64, 32, 7))


@pytest.mark.parametrize("cells,N", [((520, 4, 4), 20000), ((4, 300, 5), 10000), ((129, 129, 129), 100000), ((1, 700, 1), 3000)])
def test_grids_beyond_256_cells_per_dimension_and_awkward_shapes(orc, cuda, cells, N):
    """One bin per cell in Morton order (rank table): more than 256 cells in a dimension (the reference's 10-bit Morton fields
    allow 1024), shapes whose largest Morton code has far more bits than the grid has cells, collapsed dimensions. The lists
    that hang off the cell list must agree too: LJ forces over it against the half-cell engine."""
    from uammd_b200.md import LJ, PairForces
    L = tuple(2.5 * c for c in cells)
    pos = syn.uniform_cloud(N, L, seed=cells[0])
    cl = _check_against_oracle(orc, cuda, pos, L, cellDim=cells, rebuilds=2)
    if min(cells) >= 4:
        pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
        dpos = torch.from_numpy(pos).to(cuda)
        f0, f1 = torch.zeros(N, 4, device=cuda), torch.zeros(N, 4, device=cuda)
        PairForces(pot, Box(L)).sum(dpos, f0)
        nl = CellList()
        PairForces(pot, Box(L), nl=nl).sum(dpos, f1)
        scale = max(1.0, f0[:, :3].abs().max().item())
        assert (f0 - f1)[:, :3].abs().max().item() < 2e-5 * scale


def test_particles_outside_primary_box_are_folded(orc, cuda):
    L = (30.0, 22.0, 41.0)
    pos = syn.uniform_cloud(30000, L, seed=8)
    pos[::5, :3] *= 3.7
    _check_against_oracle(orc, cuda, pos, L)


def test_particles_on_cell_faces(orc, cuda):
    L = (40.0, 40.0, 40.0)
    pos = syn.uniform_cloud(4096, L, seed=9)
    k = np.arange(4096)
    pos[:, 0] = (k % 16) * 2.5 - 20.0          # exactly on x faces, including -L/2
    pos[::3, 1] = 20.0                          # +L/2 face (folds to cell 0)
    _check_against_oracle(orc, cuda, pos, L, cellDim=(16, 16, 16))


def test_fcc_lattice_and_repeated_rebuild_epochs(orc, cuda):
    N = 32000
    Lb = syn.lj_box_length(N)
    cl = _check_against_oracle(orc, cuda, syn.fcc_lattice(N, Lb), (Lb,) * 3, rebuilds=5)
    assert cl.getCellList()["VALID_CELL"] == 5 * N  # CellListBase.cuh:210-230


def test_collapsed_dimension_and_nonperiodic(orc, cuda):
    L = (30.0, 7.0, 30.0)  # y: 2 cells -> collapses to 1
    pos = syn.uniform_cloud(20000, L, seed=10)
    _check_against_oracle(orc, cuda, pos, L)
    _check_against_oracle(orc, cuda, pos, L, periodic=(1, 1, 0))


def test_out_of_range_nonperiodic_flags_error(orc, cuda):
    L = (30.0, 30.0, 30.0)
    pos = syn.uniform_cloud(1000, L, seed=11)
    pos[17, 2] = 40.0
    box = Box(L); box.setPeriodicity(1, 1, 0)
    cl = CellList()
    cl.update(torch.from_numpy(pos).to(cuda), box, 2.5)
    assert cl.errorFlag() == 1


def test_group_index_indirection(orc, cuda):
    L = (30.0, 30.0, 30.0)
    pos = syn.uniform_cloud(10000, L, seed=12)
    sel = np.sort(np.random.default_rng(1).choice(10000, 4000, replace=False)).astype(np.int32)
    g = orc.make_grid_f(L, orc.neighbour_celldim(L, 2.5))
    ref = orc.celllist_build(g, pos[sel])
    cl = CellList()
    cl.update(torch.from_numpy(pos).to(cuda), Box(L), 2.5, groupIndex=torch.from_numpy(sel).to(cuda))
    d = cl.getCellList()
    assert np.array_equal(d["groupIndex"].cpu().numpy(), ref["index"])
    assert np.array_equal(d["sortPos"].cpu().numpy(), ref["sortPos"])


def test_full_size_1e6_properties(orc, cuda):
    """BASELINE config 2 size: size-independent properties + bit-exact against the oracle."""
    N = 1_000_000
    Lb = syn.lj_box_length(N)
    pos = syn.uniform_cloud(N, Lb, seed=2024)
    cl = _check_against_oracle(orc, cuda, pos, (Lb,) * 3)
    d = cl.getCellList()
    idx = d["groupIndex"].long()
    assert torch.equal(torch.sort(idx).values, torch.arange(N, device=cuda))   # a permutation
    bs = d["binStart"].long()
    assert int(bs[-1]) == N and bool(torch.all(bs[1:] >= bs[:-1]))               # checksum of bin counts
    assert tuple(d["cellDim"]) == (43, 43, 43)


def test_sort_particles_is_the_reference_permutation(orc, cuda):
    """ParticleData::sortParticles: the stable Morton-hash order (the oracle's ParticleSorter restatement, pinned by the golden
    morton.bin and by the bit-exact cell lists), applied to every property; the forces of the reordered system are the
    reordered forces."""
    from uammd_b200.md import LJ, PairForces, sortParticles
    N = 50000
    Lb = syn.lj_box_length(N)
    rng = np.random.default_rng(2)
    pos = syn.uniform_cloud(N, Lb, seed=12)
    pos = pos[rng.permutation(N)]
    vel = rng.normal(size=(N, 3)).astype(np.float32)
    ids = np.arange(N, dtype=np.int32)
    box = Box(Lb)
    hashCutOff = 1.25
    cd = tuple(int(np.float32(Lb) / np.float32(hashCutOff)) for _ in range(3))
    ref = orc.celllist_build(orc.make_grid_f(box.boxSize, cd, (1, 1, 1)), pos)
    dpos, dvel, dids = (torch.from_numpy(a).to(cuda) for a in (pos, vel, ids))
    order, spos, svel, sids = sortParticles(box, hashCutOff, dpos, dvel, dids)
    assert np.array_equal(order.cpu().numpy(), ref["index"])
    assert np.array_equal(spos.cpu().numpy().view(np.uint32), pos[ref["index"]].view(np.uint32))
    assert np.array_equal(svel.cpu().numpy(), vel[ref["index"]]) and np.array_equal(sids.cpu().numpy(), ref["index"])
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    f0, f1 = torch.zeros(N, 4, device=cuda), torch.zeros(N, 4, device=cuda)
    PairForces(pot, box).sum(dpos, f0)
    PairForces(pot, box).sum(spos, f1)
    # same pairs; the order inside a half cell follows the particle indices, which the sort changes: fp32 summation order only
    assert (f1 - f0[order.long()]).abs().max().item() < 1e-5 * f0.abs().max().item()
