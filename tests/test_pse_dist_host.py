"""CPU: host logic of uammd_b200.multigpu.DistributedPSE with stand-ins for the two engines - the random draws must come in
the order of the single-GPU PSE (BDHI_PSE.cuh:141-158 and the constructors: near seed, far seed, then per noisy call the
near-field seed2 before the far-field seed2; no draw at T = 0), and the pieces must be called in the documented order."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT,) if p not in sys.path]


def _make(monkeypatch, seed):
    from uammd_b200 import bd, multigpu
    log = []

    class Near:
        def __init__(self, pos, par, sys=None, group=None, rank=None, world=None):
            rng = sys.rng()
            self.pse = types.SimpleNamespace(seedNear=rng.next32(), seedFar=rng.next32())   # what pse.PSE.__init__ draws
            log.append(("near.ctor", self.pse.seedNear, self.pse.seedFar))

        def prepare(self, stream=None):
            log.append(("prepare",))

        def Mdot(self, force, MF, stream=None):
            log.append(("mdot", force is not None))

        def noiseAdd(self, out, temperature, prefactor, seed2, stream=None):
            log.append(("noise", temperature, prefactor, seed2))
            return 7

        def errorFlag(self, stream=None):
            return 0

    class Far:
        def __init__(self, par, maxParticles, seedFar, dtype=None, group=None):
            log.append(("far.ctor", maxParticles, seedFar))
            self.fcm = types.SimpleNamespace(errorFlag=lambda: 0)

        def computeHydrodynamicDisplacements(self, pos, force, MF, temperature=0.0, prefactor=0.0, seed2=0, stream=None):
            log.append(("far", force is not None, temperature, prefactor, seed2))

        def close(self):
            log.append(("close",))

    monkeypatch.setattr(multigpu, "DistributedPSENearField", Near)
    monkeypatch.setattr(multigpu, "DistributedPSEFarField", Far)
    pos = torch.zeros(10, 4)
    m = multigpu.DistributedPSE(pos, par=None, sys=bd.System(seed))
    return m, log, bd.System(seed).rng()


def test_seed_draws_follow_the_single_gpu_order(monkeypatch):
    m, log, ref = _make(monkeypatch, 42)
    d = [ref.next32() for _ in range(6)]
    assert log[0] == ("near.ctor", d[0], d[1]) and log[1] == ("far.ctor", 10, d[1])
    MF, F = torch.ones(10, 3), torch.zeros(10, 4)
    it = m.computeHydrodynamicDisplacements(F, MF, 0.5, 2.0)
    assert it == 7 and float(MF.abs().sum()) == 0.0          # MF is zeroed first, the pieces accumulate
    assert log[2:] == [("prepare",), ("mdot", True), ("noise", 0.5, 2.0, d[2]), ("far", True, 0.5, 2.0, d[3])]
    del log[:]
    m.computeHydrodynamicDisplacements(None, MF, 0.5, 2.0)   # noise only: still both draws, near first
    assert log == [("prepare",), ("mdot", False), ("noise", 0.5, 2.0, d[4]), ("far", False, 0.5, 2.0, d[5])]


def test_no_draw_and_no_list_without_temperature_or_force(monkeypatch):
    m, log, ref = _make(monkeypatch, 7)
    del log[:]
    MF = torch.zeros(10, 3)
    assert m.computeHydrodynamicDisplacements(torch.zeros(10, 4), MF, 0.0, 1.0) == 0
    assert log == [("prepare",), ("mdot", True), ("far", True, 0.0, 1.0, 0)]
    del log[:]
    m.computeHydrodynamicDisplacements(None, MF, 0.0, 1.0)   # nothing to do in the near field: no list build either
    assert log == [("mdot", False), ("far", False, 0.0, 1.0, 0)]
    ref.next32(); ref.next32()
    assert m.sys.rng().next32() == ref.next32()              # the generator was not touched by the T = 0 calls
    assert m.errorFlag() == 0
    m.close()
    assert log[-1] == ("close",)
