"""CPU: the C restatement of VerletNVT::GronbechJensen (oracle/oracle_md.c) against golden vectors produced by the
compiled, unmodified reference on a B200 (tests/golden/nvt_gj_ref.npz, written by
tests/test_nvt_gpu.py::test_ideal_langevin_gas_bit_identical_to_reference from oracle/_ref/ref_nvt: the first 512 particles of an
ideal Langevin gas, N = 4096, 25 steps). The host libm's logf/sinf/cosf differ from the device's in the last ulp, so the
comparison is to 1e-6 of the noise scale, not bitwise - bit parity with the reference is asserted on the GPU."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvt_gj_ref.npz")


def test_xorshift_seeds_match_the_reference_draws():
    from uammd_b200.synthetic import Xorshift128plus
    g = np.load(GOLD)
    N, L, steps, T, friction, dt, sysseed = g["meta"]
    rng = Xorshift128plus(int(sysseed))
    rng.next32(); rng.next32()
    assert rng.next32() == int(g["seed"])          # the integrator's Saru seed: third draw (Basic.cu:36-38)
    assert rng.next32() == int(g["vel_seed"])      # initVelocities: the fourth (Basic.cu:74-76)


def test_initial_velocities_match_reference(orc):
    g = np.load(GOLD)
    T = np.float32(g["meta"][3])
    vamp = np.float32(np.sqrt(3.0 * float(T)))     # real velAmplitude = sqrt(3.0 * temperature) (Basic.cu:64)
    v = orc.nvt_initial_velocities(512, vamp, int(g["vel_seed"]))
    assert np.abs(v - g["vel0"]).max() <= 2e-6 * np.abs(g["vel0"]).max()
    assert abs(g["vel0"].var() / (3.0 * float(T)) - 1.0) < 0.15      # the reference's amplitude really is sqrt(3 T)


def test_ideal_gas_trajectory_matches_reference(orc):
    g = np.load(GOLD)
    N, L, steps, T, friction, dt, sysseed = g["meta"]
    f = np.float32
    amp = float(np.sqrt(f(f(f(f(2.0) * f(dt)) * f(friction)) * f(T))))   # noiseAmplitude in `real` (Basic.cu:45)
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    force = np.zeros_like(pos)
    for s in range(1, int(steps) + 1):
        orc.nvt_gj_half(pos, vel, force, float(f(dt)), float(f(friction)), amp, s, int(g["seed"]), 1)
        orc.nvt_gj_half(pos, vel, force, float(f(dt)), float(f(friction)), amp, s, int(g["seed"]), 2)
    assert np.abs(g["pos"] - g["pos0"]).max() > 1e-2
    assert np.abs(pos - g["pos"]).max() <= 1e-5
    assert np.abs(vel - g["vel"]).max() <= 1e-5


def test_basic_scheme_trajectory_matches_reference(orc):
    """VerletNVT::Basic (Basic.cu:87-172): golden vector from the compiled reference on a B200 (tests/golden/nvt_basic_ref.npz,
    same generator as above with REF_NVT_SCHEME=basic). Both half steps draw noise; the second one's Saru index is offset by
    the group size."""
    import pytest
    path = GOLD.replace("nvt_gj_ref", "nvt_basic_ref")
    if not os.path.exists(path):
        pytest.skip("tests/golden/nvt_basic_ref.npz not generated yet")
    g = np.load(path)
    N, L, steps, T, friction, dt, sysseed = g["meta"]
    f = np.float32
    amp = float(np.sqrt(f(f(f(f(2.0) * f(dt)) * f(friction)) * f(T))))
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    force = np.zeros_like(pos)
    for s in range(1, int(steps) + 1):
        orc.nvt_basic_half(pos, vel, force, float(f(dt)), float(f(friction)), amp, s, int(g["seed"]), 1, Ngroup=int(N))
        orc.nvt_basic_half(pos, vel, force, float(f(dt)), float(f(friction)), amp, s, int(g["seed"]), 2, Ngroup=int(N))
    assert np.abs(g["pos"] - g["pos0"]).max() > 1e-2
    assert np.abs(pos - g["pos"]).max() <= 1e-5
    assert np.abs(vel - g["vel"]).max() <= 1e-5
