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
