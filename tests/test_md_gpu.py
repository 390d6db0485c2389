"""GPU: velocity Verlet half steps (bit-exact vs oracle), the composed VerletNVE + PairForces step and the
fused LJMD engine (trajectory parity within fp32 tolerance, energy conservation, host-buffer entry point)."""
import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, LJMD, PairForces, VerletNVE

pytestmark = pytest.mark.gpu


def _pot():
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=2.5, sigma=1.0, epsilon=1.0)
    return pot


def test_half_steps_bit_exact(orc, cuda):
    N = 100003
    rng = np.random.default_rng(3)
    pos = rng.normal(0, 20, (N, 4)).astype(np.float32)
    vel = rng.normal(0, 1, (N, 3)).astype(np.float32)
    force = rng.normal(0, 30, (N, 4)).astype(np.float32)
    for step in (1, 2):
        p, v = pos.copy(), vel.copy()
        orc.nve_half(p, v, force, 0.005, 1.0, step)
        nve = VerletNVE(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), 0.005)
        nve.force = torch.from_numpy(force).to(cuda)
        nve._half(step)
        torch.cuda.synchronize()
        assert np.array_equal(nve.vel.cpu().numpy().view(np.uint32), v.view(np.uint32))
        assert np.array_equal(nve.pos.cpu().numpy().view(np.uint32), p.view(np.uint32))


def _setup(N, T=0.7, rho=0.8):
    Lb = syn.lj_box_length(N, rho)
    return Lb, syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, T)


def test_composed_step_matches_oracle_trajectory(orc, cuda):
    N = 4 * 12 ** 3
    Lb, pos, vel = _setup(N)
    par = syn.lj_params()
    ref = orc.MDOracle((Lb,) * 3, 2.5, par, 0.004, pos, vel)
    nve = VerletNVE(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), 0.004)
    nve.addInteractor(PairForces(_pot(), Box(Lb)))
    for _ in range(10):
        nve.forwardTime()
    ref.step(10)
    torch.cuda.synchronize()
    assert np.abs(nve.pos.cpu().numpy() - ref.pos).max() < 2e-4
    assert np.abs(nve.vel.cpu().numpy() - ref.vel).max() < 2e-3


def test_fused_engine_equals_composed_step(cuda):
    N = 4 * 16 ** 3
    Lb, pos, vel = _setup(N)
    nve = VerletNVE(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), 0.005)
    nve.addInteractor(PairForces(_pot(), Box(Lb)))
    for _ in range(7):
        nve.forwardTime()
    p = torch.from_numpy(pos).to(cuda); v = torch.from_numpy(vel).to(cuda); f = torch.zeros(N, 4, device=cuda)
    md = LJMD(Box(Lb), _pot(), 0.005)
    md.run(p, v, f, 3)
    md.run(p, v, f, 4)
    torch.cuda.synchronize()
    # identical arithmetic per particle (same kernels, same summation order): bit-exact
    assert torch.equal(p, nve.pos) and torch.equal(v, nve.vel)


def test_energy_conservation_200_steps(orc, cuda):
    N = 4 * 24 ** 3  # 55296
    Lb, pos, vel = _setup(N, T=1.0)
    pot = _pot()
    p = torch.from_numpy(pos).to(cuda); v = torch.from_numpy(vel).to(cuda); f = torch.zeros(N, 4, device=cuda)
    md = LJMD(Box(Lb), pot, 0.004)
    shifted = LJ(); shifted.setPotParameters(0, 0, cutOff=2.5, shift=True)
    pf = PairForces(shifted, Box(Lb))

    def total_energy():
        e = torch.zeros(N, device=cuda)
        pf.sum(p, energy=e)
        return float(e.double().sum() + 0.5 * (v.double() ** 2).sum())

    md.run(p, v, f, 100)   # melt the lattice first
    e0 = total_energy()
    md.run(p, v, f, 200)
    e1 = total_energy()
    assert abs(e1 - e0) / N < 2e-3, (e0 / N, e1 / N)
    assert abs(float(v.double().sum(0).abs().max())) < 1e-2 * N ** 0.5  # momentum stays ~0


def test_host_buffer_entry_point(cuda):
    N = 4 * 10 ** 3
    Lb, pos, vel = _setup(N)
    md = LJMD(Box(Lb), _pot(), 0.005)
    p = torch.from_numpy(pos).to(cuda); v = torch.from_numpy(vel).to(cuda); f = torch.zeros(N, 4, device=cuda)
    md.run(p, v, f, 5)
    hp = torch.from_numpy(pos.copy()).pin_memory(); hv = torch.from_numpy(vel.copy()).pin_memory()
    hf = torch.zeros(N, 4).pin_memory()
    md2 = LJMD(Box(Lb), _pot(), 0.005)
    md2.runHost(hp, hv, hf, 5)
    assert torch.equal(hp, p.cpu()) and torch.equal(hv, v.cpu()) and torch.equal(hf, f.cpu())
