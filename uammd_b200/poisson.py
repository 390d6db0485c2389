"""Host-side mirror of the reference's Poisson interactor (Interactor/SpectralEwaldPoisson.cuh:84-184) over the C ABI
(ub200_poisson_*): same Parameters, sum(force=, energy=) accumulates like Interactor::sum, computeFieldPotentialAtParticles()
returns (Ex, Ey, Ez, phi) per particle. torch only owns device memory and streams; there is no CPU fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import UB200Error, check
from .md import _ptr, _stream_ptr


class PoissonParams(C.Structure):
    _fields_ = [("L", C.c_double * 3), ("epsilon", C.c_double), ("tolerance", C.c_double), ("gw", C.c_double),
                ("split", C.c_double), ("upsampling", C.c_double)]


class PoissonInfo(C.Structure):
    _fields_ = [("cells", C.c_int * 3), ("support", C.c_int), ("nTable", C.c_int), ("h", C.c_double),
                ("farFieldGaussianWidth", C.c_double), ("nearFieldCutOff", C.c_double)]


def _declare():
    lib = _lib.lib()
    if getattr(lib, "_poisson_declared", False):
        return lib
    vp, i = C.c_void_p, C.c_int
    lib.ub200_poisson_create.restype = i
    lib.ub200_poisson_create.argtypes = [C.POINTER(vp), i, C.POINTER(PoissonParams)]
    lib.ub200_poisson_destroy.argtypes = [vp]
    lib.ub200_poisson_info.argtypes = [vp, C.POINTER(PoissonInfo)]
    lib.ub200_poisson_sum.restype = i
    lib.ub200_poisson_sum.argtypes = [vp, vp, vp, i, vp, vp, vp]
    lib.ub200_poisson_field_potential.restype = i
    lib.ub200_poisson_field_potential.argtypes = [vp, vp, vp, i, vp, vp]
    lib._poisson_declared = True
    return lib


class Parameters:
    """Poisson::Parameters (SpectralEwaldPoisson.cuh:93-104); cells / support are accepted and, like in the reference's
    constructor, not used."""

    def __init__(self, box, epsilon=-1.0, tolerance=1e-5, gw=-1.0, split=-1.0, upsampling=-1.0, cells=None, support=-1):
        self.box = (box, box, box) if isinstance(box, (int, float)) else tuple(box)
        self.epsilon, self.tolerance, self.gw, self.split, self.upsampling = epsilon, tolerance, gw, split, upsampling
        self.cells, self.support = cells, support


class Poisson:
    """Poisson(pd, par): pos real4 [N,4], charge real [N] (float32 or float64 CUDA tensors = the reference's `real`)."""

    def __init__(self, pos, charge, par):
        self.lib = _declare()
        if not (pos.is_cuda and pos.dim() == 2 and pos.shape[1] == 4 and pos.is_contiguous() and charge.is_cuda and
                charge.dtype == pos.dtype and charge.shape == (pos.shape[0],) and pos.dtype in (torch.float32, torch.float64)):
            raise UB200Error("Poisson: pos must be a contiguous CUDA [N,4] tensor and charge a [N] tensor of the same dtype")
        self.pos, self.charge, self.par, self.N = pos, charge, par, pos.shape[0]
        p = PoissonParams()
        p.L[:] = [float(x) for x in par.box]
        p.epsilon, p.tolerance, p.gw, p.split, p.upsampling = par.epsilon, par.tolerance, par.gw, par.split, par.upsampling
        self._h = C.c_void_p()
        rc = self.lib.ub200_poisson_create(C.byref(self._h), 8 if pos.dtype == torch.float64 else 4, C.byref(p))
        if rc == -6:
            raise UB200Error("[Poisson] Kernel support is too large for this configuration, try increasing the splitting "
                             "parameter or decreasing the tolerance")
        if rc == -1:
            raise UB200Error("[Poisson] invalid parameters (near field cut off larger than half the box? increase the splitting parameter)")
        check(rc)

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_poisson_destroy(self._h)
        except Exception:
            pass

    def info(self):
        inf = PoissonInfo()
        check(self.lib.ub200_poisson_info(self._h, C.byref(inf)))
        return inf

    def sum(self, pos=None, force=None, energy=None, virial=None, stream=None):
        if virial is not None:
            raise UB200Error("[Poisson] Virial functionality not implemented.")   # SpectralEwaldPoisson.cuh:117-121
        pos = self.pos if pos is None else pos
        check(self.lib.ub200_poisson_sum(self._h, _ptr(pos), _ptr(self.charge), self.N, _ptr(force), _ptr(energy),
                                         _stream_ptr(stream)))

    def computeFieldPotentialAtParticles(self, stream=None):
        out = torch.zeros(self.N, 4, dtype=self.pos.dtype, device=self.pos.device)
        check(self.lib.ub200_poisson_field_potential(self._h, _ptr(self.pos), _ptr(self.charge), self.N, _ptr(out),
                                                     _stream_ptr(stream)))
        return out
