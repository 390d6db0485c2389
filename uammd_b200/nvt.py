"""Host-side mirror of the reference's VerletNVT::GronbechJensen (Integrator/VerletNVT.cuh:59-117, VerletNVT/Basic.cu:31-77,
VerletNVT/GronbechJensen.cu:68-127) over the C ABI (ub200_nvt_gj_half_step_f32, ub200_nvt_initial_velocities_f32): the
Langevin integrator that generic_md and examples/misc/benchmark.cu instantiate.
torch only owns device memory and streams; there is no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from ._lib import UB200Error, check, lib
from .bd import System
from .md import _ptr, _stream_ptr


def _declare():
    l = lib()
    vp, i, f, u32 = C.c_void_p, C.c_int, C.c_float, C.c_uint32
    l.ub200_nvt_gj_half_step_f32.restype = i
    l.ub200_nvt_gj_half_step_f32.argtypes = [vp, vp, vp, vp, f, vp, i, f, f, i, f, u32, u32, i, vp]
    l.ub200_nvt_basic_half_step_f32.restype = i
    l.ub200_nvt_basic_half_step_f32.argtypes = [vp, vp, vp, vp, f, vp, i, f, f, i, f, u32, u32, i, vp]
    l.ub200_nvt_initial_velocities_f32.restype = i
    l.ub200_nvt_initial_velocities_f32.argtypes = [vp, vp, i, f, i, u32, vp]
    return l


class Parameters:
    """VerletNVT::Basic::Parameters (VerletNVT.cuh:61-68)."""

    def __init__(self, temperature=0.0, dt=0.0, friction=1.0, is2D=False, initVelocities=True, mass=-1.0):
        self.temperature, self.dt, self.friction = temperature, dt, friction
        self.is2D, self.initVelocities, self.mass = is2D, initVelocities, mass


class GronbechJensen:
    """VerletNVT::GronbechJensen(pd, par). pos real4[N], vel real3[N] (float32 CUDA tensors, updated in place); mass: optional
    per-particle masses (pd->getMass), used when par.mass < 0. Interactors are objects with sum(pos, force=) that ACCUMULATE
    into the real4 force tensor. The constructor draws the Saru seed as the third next32() of the system generator and, with
    par.initVelocities, fills the velocities with the fourth as seed - like VerletNVT::Basic (Basic.cu:36-52,56-77)."""

    def __init__(self, pos, vel, par, sys=None, mass=None, groupIndex=None):
        self.l = _declare()
        for t, w in ((pos, 4), (vel, 3)):
            if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != w or not t.is_cuda or not t.is_contiguous():
                raise UB200Error("GronbechJensen: pos / vel must be contiguous CUDA float32 [N,4] / [N,3] tensors")
        self.pos, self.vel, self.par, self.mass, self.groupIndex = pos, vel, par, mass, groupIndex
        self.N = pos.shape[0] if groupIndex is None else groupIndex.numel()
        self.force = torch.zeros_like(pos)
        self.sys = sys if sys is not None else System()
        self.sys.rng().next32(); self.sys.rng().next32()
        self.seed = self.sys.rng().next32()
        f = np.float32
        self.dt, self.friction, self.temperature = f(par.dt), f(par.friction), f(par.temperature)
        # this->noiseAmplitude = sqrt(2 * dt * friction * temperature) evaluated in `real` (Basic.cu:45)
        self.noiseAmplitude = float(np.sqrt(f(f(f(f(2.0) * self.dt) * self.friction) * self.temperature)))
        self.defaultMass = float(par.mass)
        if mass is None and self.defaultMass < 0:
            self.defaultMass = 1.0
        self.is2D = bool(par.is2D)
        self.interactors = []
        self.steps = 0
        if par.initVelocities:
            self.initVelocities()

    def addInteractor(self, it):
        self.interactors.append(it)

    def initVelocities(self):
        velAmplitude = float(f32(np.sqrt(3.0 * float(self.temperature))))
        check(self.l.ub200_nvt_initial_velocities_f32(_ptr(self.vel), _ptr(self.groupIndex), self.N, velAmplitude,
                                                      int(self.is2D), self.sys.rng().next32(), _stream_ptr()))

    _half_step_symbol = "ub200_nvt_gj_half_step_f32"

    def _half(self, step):
        mass = None if self.defaultMass > 0 else self.mass
        check(getattr(self.l, self._half_step_symbol)(_ptr(self.pos), _ptr(self.vel), _ptr(self.force), _ptr(mass),
                                                self.defaultMass if self.defaultMass > 0 else 0.0, _ptr(self.groupIndex),
                                                self.N, float(self.dt), float(self.friction), int(self.is2D),
                                                self.noiseAmplitude, self.steps & 0xFFFFFFFF, self.seed, step, _stream_ptr()))

    def _sumForces(self):
        for it in self.interactors:
            it.sum(self.pos, force=self.force)

    def forwardTime(self):
        """GronbechJensen::forwardTime (GronbechJensen.cu:96-127): step 1 zeroes the forces itself."""
        self.steps += 1
        if self.steps == 1:
            self.force.zero_()
            self._sumForces()
        self._half(1)
        self._sumForces()
        self._half(2)


class Basic(GronbechJensen):
    """VerletNVT::Basic(pd, par) (Integrator/VerletNVT/Basic.cu:31-52,87-172): the plain Langevin velocity Verlet - friction and a
    fresh noise draw in both half kicks. Same constructor, seed draws, initial velocities and forwardTime sequence as
    GronbechJensen (which derives from it in the reference, VerletNVT.cuh:59-117)."""
    _half_step_symbol = "ub200_nvt_basic_half_step_f32"


def f32(x):
    return np.float32(x)
