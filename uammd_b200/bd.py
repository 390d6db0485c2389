"""Host-side mirror of the reference's BD::EulerMaruyama (Integrator/BrownianDynamics.cuh:111-126, .cu:10-173) over
the C ABI (ub200_bd_euler_maruyama_step). BASELINE config 0: ideal Brownian particles, fp64 - the README example.
torch only owns device memory and streams; there is no CPU fallback."""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import check, lib


def _declare():
    l = lib()
    vp, i, d, u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint32
    l.ub200_bd_euler_maruyama_step.restype = i
    l.ub200_bd_euler_maruyama_step.argtypes = [i, vp, vp, vp, C.POINTER(d), d, vp, d, i, d, i, u32, u32, vp]
    return l


class System:
    """The part of uammd::System the integrators use: the host Xorshift128plus stream (System/System.h:199-275)."""

    def __init__(self, seed=None):
        from .synthetic import Xorshift128plus
        self._rng = Xorshift128plus(seed)

    def rng(self):
        return self._rng


class Parameters:
    """BD::Parameters (BrownianDynamics.cuh:47-62)."""

    def __init__(self, temperature=0.0, viscosity=1.0, hydrodynamicRadius=-1.0, dt=0.0, is2D=False, K=None):
        self.temperature, self.viscosity, self.hydrodynamicRadius, self.dt = temperature, viscosity, hydrodynamicRadius, dt
        self.is2D, self.K = is2D, K


class EulerMaruyama:
    """BD::EulerMaruyama(pd, par): forwardTime() = steps++, forces from the interactors, position update.

    pos: device real4[N] tensor (float32 or float64 = the reference's `real`), updated in place. interactors are
    callables `f(force_tensor)` that ACCUMULATE into the real4 force tensor (Interactor::sum semantics). The Saru
    seed is drawn like BaseBrownianIntegrator's constructor does (BrownianDynamics.cu:14-16): the third next32()."""

    def __init__(self, pos, par, sys=None, radius=None, groupIndex=None):
        self.l = _declare()
        assert pos.is_cuda and pos.dim() == 2 and pos.shape[1] == 4 and pos.is_contiguous()
        assert pos.dtype in (torch.float32, torch.float64)
        self.pos, self.par, self.radius, self.groupIndex = pos, par, radius, groupIndex
        self.N = pos.shape[0] if groupIndex is None else groupIndex.numel()
        self.sys = sys if sys is not None else System()
        self.sys.rng().next32(); self.sys.rng().next32()
        self.seed = self.sys.rng().next32()
        real = np.float32 if pos.dtype == torch.float32 else np.float64
        m = real(1.0 / (6.0 * math.pi * par.viscosity))
        if par.hydrodynamicRadius != -1.0:
            m = real(m / real(par.hydrodynamicRadius))
        self.selfMobility = float(m)
        self.steps = 0
        self.interactors = []
        self.force = None
        self._K = (C.c_double * 9)(*[float(x) for row in par.K for x in row]) if par.K is not None else None

    def addInteractor(self, fn):
        self.interactors.append(fn)
        if self.force is None:
            self.force = torch.zeros_like(self.pos)

    def forwardTime(self, stream=None):
        self.steps += 1
        if self.interactors:
            self.force.zero_()
            for fn in self.interactors:
                fn(self.force)
        st = stream if stream is not None else torch.cuda.current_stream(self.pos.device).cuda_stream
        check(self.l.ub200_bd_euler_maruyama_step(
            8 if self.pos.dtype == torch.float64 else 4, self.pos.data_ptr(),
            self.groupIndex.data_ptr() if self.groupIndex is not None else None,
            self.force.data_ptr() if self.force is not None else None, self._K, self.selfMobility,
            self.radius.data_ptr() if self.radius is not None else None, self.par.dt, int(self.par.is2D),
            self.par.temperature, self.N, self.steps, self.seed, st))
