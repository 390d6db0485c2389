"""GPU: VerletNVT::GronbechJensen (uammd_b200/nvt.py -> ub200_nvt_gj_half_step_f32 / ub200_nvt_initial_velocities_f32)
against the compiled, unmodified reference (oracle/_ref/ref_nvt) and the C oracle.

Without interactors the trajectory depends only on the integrator's arithmetic and its Saru stream, so the reference must
be reproduced BIT FOR BIT (the kernel spells out the roundings nvcc applies to the reference kernel). With the LJ
interactor the two force kernels sum in different orders (fp32 rounding), so the comparison is a short trajectory
within a stated tolerance."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from uammd_b200.bd import System
from uammd_b200.md import Box, LJ, PairForces, VerletList
from uammd_b200.nvt import Basic, GronbechJensen, Parameters

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref(tmp_path, N, L, steps, T, friction, dt, sysseed, lj, initVel, scheme="gj"):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_nvt")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_nvt not built (needs the reference tree at build time)")
    out = str(tmp_path / "nvt")
    r = subprocess.run([exe, str(N), repr(L), str(steps), repr(T), repr(friction), repr(dt), str(sysseed), str(int(lj)),
                        str(int(initVel)), out], check=True, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, REF_NVT_SCHEME="basic" if scheme == "basic" else "gj")).stdout
    info = json.loads([l for l in r.splitlines() if l.startswith("{")][-1])
    rd = lambda name, w: np.fromfile(f"{out}.{name}.bin", dtype=np.float32).reshape(N, w)
    return info, rd("pos0", 4), rd("vel0", 3), rd("pos", 4), rd("vel", 3)


def _record(name, payload):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, name), "w") as f:
            json.dump(payload, f)


@pytest.mark.parametrize("scheme", ["gj", "basic"])
def test_ideal_langevin_gas_bit_identical_to_reference(cuda, tmp_path, scheme):
    """scheme "basic": VerletNVT::Basic (Basic.cu:87-172) through a derived class in the harness - the reference never defines
    its public constructor."""
    N, L, steps, T, friction, dt, sysseed = 4096, 32.0, 25, 1.3, 0.7, 0.01, 1234
    info, pos0, vel0, rpos, rvel = _ref(tmp_path, N, L, steps, T, friction, dt, sysseed, lj=False, initVel=True, scheme=scheme)
    assert info["mode"] == ("nvt_basic" if scheme == "basic" else "nvt_gj")
    GronbechJensen = Basic if scheme == "basic" else globals()["GronbechJensen"]
    pos = torch.from_numpy(pos0.copy()).to(cuda)
    vel = torch.zeros(N, 3, device=cuda)
    sys_ = System(sysseed)
    nvt = GronbechJensen(pos, vel, Parameters(temperature=T, dt=dt, friction=friction, initVelocities=True), sys=sys_)
    assert nvt.seed == info["seed"]
    torch.cuda.synchronize()
    v0 = vel.cpu().numpy()
    mism_v0 = int((v0.view(np.uint32) != vel0.view(np.uint32)).sum())
    for _ in range(steps):
        nvt.forwardTime()
    torch.cuda.synchronize()
    p, v = pos.cpu().numpy(), vel.cpu().numpy()
    mism_p = int((p.view(np.uint32) != rpos.view(np.uint32)).sum())
    mism_v = int((v.view(np.uint32) != rvel.view(np.uint32)).sum())
    _record("nvt_parity.json" if scheme == "gj" else "nvt_basic_parity.json", {"N": N, "steps": steps, "mismatch_words": {"vel0": mism_v0, "pos": mism_p, "vel": mism_v},
                                "max_abs": {"vel0": float(np.abs(v0 - vel0).max()), "pos": float(np.abs(p - rpos).max()),
                                            "vel": float(np.abs(v - rvel).max())}})
    g = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(g):   # golden vectors for the CPU oracle test (tests/golden/nvt_gj_ref.npz is made from these)
        np.savez_compressed(os.path.join(g, f"nvt_{scheme}_ref.npz"), pos0=pos0[:512], vel0=vel0[:512], pos=rpos[:512], vel=rvel[:512],
                            seed=info["seed"], vel_seed=info["vel_seed"], meta=np.array([N, L, steps, T, friction, dt, sysseed]))
    assert np.abs(vel0).max() > 1.0 and np.abs(rpos - pos0).max() > 1e-2          # the reference did move
    assert mism_v0 == 0, f"initial velocities differ from the reference in {mism_v0} words"
    assert mism_p == 0 and mism_v == 0, f"trajectory differs from the reference: {mism_p} pos / {mism_v} vel words"


@pytest.mark.parametrize("scheme", ["gj", "basic"])
def test_half_steps_match_oracle_with_forces_masses_and_2d(orc, cuda, scheme):
    """Random forces, per-particle masses, 2-D mode, a group index list: against the C restatement. The host libm's
    logf/sinf/cosf differ from the device's in the last ulp: tolerance 2e-6 of the noise amplitude scale."""
    N = 5000
    rng = np.random.default_rng(3)
    for is2D, use_mass in ((False, False), (True, True)):
        pos = rng.normal(0, 5, (N, 4)).astype(np.float32)
        vel = rng.normal(0, 1, (N, 3)).astype(np.float32)
        force = rng.normal(0, 20, (N, 4)).astype(np.float32)
        mass = rng.uniform(0.5, 3.0, N).astype(np.float32) if use_mass else None
        par = Parameters(temperature=0.9, dt=0.005, friction=2.0, is2D=is2D, initVelocities=False)
        dp, dv = torch.from_numpy(pos.copy()).to(cuda), torch.from_numpy(vel.copy()).to(cuda)
        cls, half = (Basic, orc.nvt_basic_half) if scheme == "basic" else (GronbechJensen, orc.nvt_gj_half)
        nvt = cls(dp, dv, par, sys=System(77), mass=torch.from_numpy(mass).to(cuda) if use_mass else None)
        nvt.force.copy_(torch.from_numpy(force))
        nvt.steps = 5
        nvt._half(1)
        torch.cuda.synchronize()
        assert float(nvt.force.abs().max()) == 0.0                       # step 1 resets the forces
        nvt.force.copy_(torch.from_numpy(force))
        nvt._half(2)
        torch.cuda.synchronize()
        op, ov, of = pos.copy(), vel.copy(), force.copy()
        kw = dict(defaultMass=0.0 if use_mass else 1.0, mass=mass, is2D=is2D)
        half(op, ov, of, par.dt, par.friction, nvt.noiseAmplitude, 5, nvt.seed, 1, **kw)
        assert np.all(of == 0)
        half(op, ov, force.copy(), par.dt, par.friction, nvt.noiseAmplitude, 5, nvt.seed, 2, **kw)
        assert np.abs(dp.cpu().numpy() - op).max() <= 2e-6 * (1.0 + np.abs(op).max())   # an ulp of the largest coordinate
        assert np.abs(dv.cpu().numpy() - ov).max() <= 2e-6 * (1.0 + np.abs(ov).max())
        if is2D:
            assert np.all(dv.cpu().numpy()[:, 2] == 0)


def test_lj_langevin_run_next_to_reference(cuda, tmp_path):
    """benchmark.cu's configuration: VerletNVT::GronbechJensen + PairForces<LJ, VerletList> (here 20 steps of a liquid)."""
    N, steps, T, friction, dt, sysseed = 32768, 20, 1.0, 1.0, 0.002, 99
    L = float(np.float32((N / 0.6) ** (1 / 3)))
    info, pos0, vel0, rpos, rvel = _ref(tmp_path, N, L, steps, T, friction, dt, sysseed, lj=True, initVel=True)
    pos = torch.from_numpy(pos0.copy()).to(cuda)
    vel = torch.zeros(N, 3, device=cuda)
    nvt = GronbechJensen(pos, vel, Parameters(temperature=T, dt=dt, friction=friction, initVelocities=True), sys=System(sysseed))
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    nvt.addInteractor(PairForces(pot, Box(L), nl=VerletList()))
    for _ in range(steps):
        nvt.forwardTime()
    torch.cuda.synchronize()
    dp = np.abs(pos.cpu().numpy() - rpos).max()
    dv = np.abs(vel.cpu().numpy() - rvel).max()
    _record("nvt_lj_parity.json", {"N": N, "steps": steps, "max_dpos": float(dp), "max_dvel": float(dv),
                                   "ref_ms_per_step": info["ms_per_step"]})
    assert np.abs(rpos - pos0).max() > 1e-2
    assert dp < 1e-4 and dv < 1e-2, (dp, dv)
