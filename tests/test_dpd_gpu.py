"""GPU: DPD pair forces through the C ABI vs the oracle (Saru stream bit-exact -> forces within fp32 tolerance)."""
import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, DPD, PairForcesDPD

pytestmark = pytest.mark.gpu


def _run(orc, cuda, N, L, seed=1234, steps=2, periodic=(1, 1, 1)):
    pos = syn.uniform_cloud(N, L, seed=21)
    vel = syn.maxwell_velocities(N, 1.0, seed=22)
    pot = DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=seed)
    box = Box(L); box.setPeriodicity(*periodic)
    pf = PairForcesDPD(pot, box)
    dpos, dvel = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda)
    g = orc.make_grid_f(box.boxSize, orc.neighbour_celldim(box.boxSize, 1.0), periodic)
    cl = orc.celllist_build(g, pos)
    out = None
    for s in range(1, steps + 1):
        force = torch.zeros(N, 4, device=cuda)
        pf.sum(dpos, dvel, force)
        torch.cuda.synchronize()
        f32, f64 = orc.dpd_f32(g, cl, vel, 25.0, 4.5, pot.sigma, 1.0, seed, s, N)
        F = force.cpu().numpy()
        scale = np.abs(f64).max()
        # libdevice vs glibc logf/sinf differ by an ulp; everything else is the same arithmetic
        assert np.abs(F[:, :3] - f64).max() < 2e-4 * scale, np.abs(F[:, :3] - f64).max() / scale
        assert np.all(F[:, 3] == 0)
        tot = F[:, :3].astype(np.float64).sum(0)
        assert np.abs(tot).max() < 1e-5 * np.abs(F[:, :3]).sum()  # pairwise antisymmetric noise
        assert out is None or not np.allclose(out, F)            # new noise every step
        out = F
    return out


@pytest.mark.parametrize("N,L", [(3000, (10.0, 10.0, 10.0)), (81000, (30.0, 30.0, 30.0))])
def test_dpd_forces_match_oracle(orc, cuda, N, L):
    _run(orc, cuda, N, L)


def test_dpd_collapsed_dimension(orc, cuda):
    _run(orc, cuda, 12000, (20.0, 3.5, 20.0))


def test_dpd_id_product_wraps_like_reference(orc, cuda):
    # N > 46340 -> i + N*j overflows int32 exactly as in DPD.cuh:128
    _run(orc, cuda, 120000, (34.2, 34.2, 34.2), steps=1)


def test_dpd_full_size_momentum(cuda):
    """BASELINE config 5 size per GPU share (4e6/8): total momentum change vanishes."""
    N = 500000
    L = (55.032,) * 3
    pos = syn.uniform_cloud(N, L, seed=21)
    vel = syn.maxwell_velocities(N, 1.0, seed=22)
    pf = PairForcesDPD(DPD(1.0, 0.01, 4.5, 1.0, 25.0, seed=7), Box(L))
    force = torch.zeros(N, 4, device=cuda)
    pf.sum(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), force)
    F = force.double()
    assert float(F[:, :3].sum(0).abs().max()) < 1e-5 * float(F[:, :3].abs().sum())


def _dpd_forces(cuda, pos, vel, L, tile, seed=5, step=3):
    """One DPD force evaluation with the traversal kernel forced by UB200_DPD_TILE (read by the library at every call)."""
    import os
    N = pos.shape[0]
    pot = DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=seed)
    pot.step = step - 1
    force = torch.zeros(N, 4, device=cuda)
    old = os.environ.get("UB200_DPD_TILE")
    os.environ["UB200_DPD_TILE"] = "1" if tile else "0"
    try:
        PairForcesDPD(pot, Box(L)).sum(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), force)
        torch.cuda.synchronize()
    finally:
        if old is None:
            del os.environ["UB200_DPD_TILE"]
        else:
            os.environ["UB200_DPD_TILE"] = old
    return force.cpu().numpy(), pot


def test_dpd_block_kernel_equals_cell_kernel_bitwise(orc, cuda):
    """dpdTileTraversal (CTA per 4x4x4 block of cells, the default at DPD densities) visits the candidates of a home
    particle in the flattened order of dpdCellTraversal with the same lanes: identical bits. 30 cells per dimension are not
    a multiple of the block size, so partial blocks and the periodic wrap of the halo are exercised."""
    N, L = 81000, 30.0
    pos, vel = syn.uniform_cloud(N, L, seed=31), syn.maxwell_velocities(N, 1.0, seed=32)
    f_cell, _ = _dpd_forces(cuda, pos, vel, L, tile=False)
    f_tile, _ = _dpd_forces(cuda, pos, vel, L, tile=True)
    assert np.abs(f_cell[:, :3]).max() > 10.0
    assert np.array_equal(f_cell.view(np.uint32), f_tile.view(np.uint32))


def test_dpd_dense_cluster_takes_the_fallbacks(orc, cuda):
    """A dilute fluid with one dense cluster: the cluster's neighbourhoods exceed the staging areas, so the cell kernel walks
    global memory and the block kernel falls back to the per-cell algorithm for the blocks around it. Both must agree bit
    for bit with each other and, within the usual tolerance, with the oracle."""
    L = 12.0
    rng = np.random.default_rng(9)
    dilute = syn.uniform_cloud(3000, L, seed=41)
    blob = np.zeros((1500, 4), np.float32)
    blob[:, :3] = (rng.random((1500, 3)) * 2.0 + 1.0).astype(np.float32)      # 1500 particles in a 2x2x2 region
    pos = np.concatenate([dilute, blob]).astype(np.float32)
    N = pos.shape[0]
    vel = syn.maxwell_velocities(N, 1.0, seed=42)
    f_cell, pot = _dpd_forces(cuda, pos, vel, L, tile=False)
    f_tile, _ = _dpd_forces(cuda, pos, vel, L, tile=True)
    assert np.array_equal(f_cell.view(np.uint32), f_tile.view(np.uint32))
    g = orc.make_grid_f((L,) * 3, orc.neighbour_celldim((L,) * 3, 1.0))
    cl = orc.celllist_build(g, pos)
    assert (cl["cellEnd"] - cl["cellStart"]).max() > 100                      # the cluster really is dense
    _, f64 = orc.dpd_f32(g, cl, vel, 25.0, 4.5, pot.sigma, 1.0, pot.seed, pot.step, N)
    # per particle: error relative to the sum of its pair force magnitudes would be the sharp bound; the largest force is enough
    assert np.abs(f_tile[:, :3] - f64).max() < 2e-4 * np.abs(f64).max()
