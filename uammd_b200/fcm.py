"""Host-side mirror of the reference's operator interface for path 2 (FFT-based hydrodynamics).

  kernels: Peskin3 / Peskin4 / Gaussian        Integrator/BDHI/FCM/FCM_kernels.cuh:22-58,159-196
  FFT3D                                         cuFFT plans of FCM_impl.cuh:179-234
  IBM                                           misc/IBM.cuh:99-203  (spread / gather)
  FCM_impl                                      Integrator/BDHI/FCM/FCM_impl.cuh:36-129,652-693
  FCM (BDHI Method) + EulerMaruyama             BDHI_FCM.cuh:85-153, BDHI_EulerMaruyama.cu:82-166
Everything computes through the C ABI (include/uammd_b200.h); torch only owns device memory/streams.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import UB200Error, check, d3, i3
from .md import _ptr, _stream_ptr


class IBMKernelStruct(C.Structure):
    _fields_ = [("kind", C.c_int), ("support", C.c_int), ("h", C.c_double), ("prefactor", C.c_double),
                ("tau", C.c_double), ("rmax", C.c_double)]


def _declare():
    lib = _lib.lib()
    if getattr(lib, "_fcm_declared", False):
        return lib
    vp, i, d, u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint32
    D3, I3 = C.c_double * 3, C.c_int * 3
    K = C.POINTER(IBMKernelStruct)
    sig = {
        "ub200_fft3d_create": (i, [C.POINTER(vp), i, i, i, i]),
        "ub200_fft3d_destroy": (i, [vp]),
        "ub200_fft3d_exec": (i, [vp, vp, i, vp]),
        "ub200_ibm_create": (i, [C.POINTER(vp), i, D3, I3, I3, K, i]),
        "ub200_ibm_destroy": (i, [vp]),
        "ub200_ibm_spread": (i, [vp, vp, vp, i, i, vp, vp]),
        "ub200_ibm_spread_overwrite": (i, [vp, vp, vp, i, i, vp, vp]),
        "ub200_ibm_gather": (i, [vp, vp, i, vp, vp, vp]),
        "ub200_fcm_create": (i, [C.POINTER(vp), i, D3, I3, K, d, u32]),
        "ub200_fcm_destroy": (i, [vp]),
        "ub200_fcm_mdot": (i, [vp, vp, vp, i, d, d, vp, vp]),
        "ub200_fcm_grid_info": (i, [vp, I3, C.POINTER(i), C.POINTER(vp)]),
        "ub200_fcm_set_torque_kernel": (i, [vp, K]),
        "ub200_fcm_mdot_torque": (i, [vp, vp, vp, vp, i, d, d, vp, vp, vp]),
        "ub200_bdhi_euler_update": (i, [i, vp, vp, vp, vp, vp, i, d, d, i, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._fcm_declared = True
    return lib


# ---------------- spreading kernels (host-side parameter resolution, like the reference constructors) ----
class Peskin3:
    """FCM_ns::Kernels::Peskin::threePoint (FCM_kernels.cuh:159-176): support 3, hydrodynamic radius = h."""
    support = 3

    def __init__(self, h, tolerance=None):
        self.h = float(h)

    @staticmethod
    def adviseGridSize(hydrodynamicRadius, tolerance=None):
        return hydrodynamicRadius

    def fixHydrodynamicRadius(self, hydrodynamicRadius, h):
        return h

    def struct(self):
        return IBMKernelStruct(0, 3, self.h, 0.0, 0.0, 0.0)


class Peskin4:
    """FCM_ns::Kernels::Peskin::fourPoint (FCM_kernels.cuh:178-196): support 4, a = 1.31 h."""
    support = 4
    fac = 1.31

    def __init__(self, h, tolerance=None):
        self.h = float(h)

    @staticmethod
    def adviseGridSize(hydrodynamicRadius, tolerance=None):
        return hydrodynamicRadius / Peskin4.fac

    def fixHydrodynamicRadius(self, hydrodynamicRadius, h):
        return h * Peskin4.fac

    def struct(self):
        return IBMKernelStruct(1, 4, self.h, 0.0, 0.0, 0.0)


class Gaussian:
    """FCM_ns::Kernels::Gaussian (FCM_kernels.cuh:22-58): width = h * upsampling(tolerance); the support is
    found by marching r in steps of h/2 until phi(r) <= tolerance."""

    @staticmethod
    def computeUpsampling(tolerance):
        amin, amax = 0.55, 1.65
        x = -math.log10(3 * tolerance) / 10.0
        return min(amin + x * (amax - amin), amax)

    def __init__(self, h, tolerance):
        self.h = float(h)
        ups = self.computeUpsampling(tolerance)
        width = h * ups
        self.prefactor = (2.0 * math.pi * width * width) ** -0.5
        self.tau = -0.5 / (width * width)
        dr = 0.5 * h
        r = dr
        while self.prefactor * math.exp(self.tau * r * r) > tolerance:
            r += dr
        self.support = max(3, int(2 * r / h + 0.5))
        self.rmax = self.support * h
        self.a = h * ups * math.sqrt(math.pi)

    @staticmethod
    def adviseGridSize(hydrodynamicRadius, tolerance):
        return hydrodynamicRadius / (math.sqrt(math.pi) * Gaussian.computeUpsampling(tolerance))

    def fixHydrodynamicRadius(self, hydrodynamicRadius, h):
        return self.a

    def struct(self):
        return IBMKernelStruct(2, self.support, self.h, self.prefactor, self.tau, self.rmax)


class GaussianTorque:
    """FCM_ns::Kernels::GaussianTorque (FCM_kernels.cuh:60-80): IBM_kernels::Gaussian(width) truncated where it drops below
    the tolerance. forHydrodynamicRadius builds it like detail::initializeKernelTorque (BDHI_FCM.cuh:69-80)."""

    def __init__(self, width, h, tolerance):
        self.width, self.h = width, h
        self.prefactor = (2.0 * math.pi * width * width) ** -0.5
        self.tau = -0.5 / (width * width)
        dr = 0.5 * h
        r = dr
        while self.prefactor * math.exp(self.tau * r * r) > tolerance:
            r += dr
        self.support = max(3, int(2 * r / h + 0.5))
        self.rmax = self.support * h

    @staticmethod
    def forHydrodynamicRadius(a, h, tolerance):
        return GaussianTorque(a / (6 * math.sqrt(math.pi)) ** (1 / 3.0), h, tolerance)

    def struct(self):
        return IBMKernelStruct(2, self.support, self.h, self.prefactor, self.tau, self.rmax)


class BarnettMagland:
    """IBM_kernels::BarnettMagland (misc/IBM_kernels.cuh:83-113), the "exponential of a semicircle" window of Barnett, Magland
    and af Klinteberg: phi(r) = exp(beta (sqrt(1 - (r/alpha)^2) - 1)) / norm for |r| <= alpha. alpha = half the support width,
    `support` = grid points per dimension (the reference leaves it to the wrapping kernel). The norm follows computeNorm
    (:93-97): 2 x composite Simpson over [0, alpha] with 20000 intervals, compensated sums."""

    def __init__(self, alpha, beta, support):
        self.alpha, self.beta, self.support = float(alpha), float(beta), int(support)
        n = 20000
        dx = self.alpha / n
        terms = []
        for i in range(n + 1):
            z = (i * dx) / self.alpha
            dz2 = 1.0 - z * z
            w = 1.0 if i in (0, n) else (4.0 if i % 2 else 2.0)
            terms.append(w * (0.0 if dz2 < 0.0 else math.exp(self.beta * (math.sqrt(dz2) - 1.0))))
        self.norm = 2.0 * (dx / 3.0 * math.fsum(terms))
        self.h = 2.0 * self.alpha / self.support

    def phi(self, r):
        z = r / self.alpha
        dz2 = 1.0 - z * z
        return 0.0 if dz2 < 0.0 else math.exp(self.beta * (math.sqrt(dz2) - 1.0)) / self.norm

    def struct(self):
        return IBMKernelStruct(3, self.support, self.h, 1.0 / self.norm, self.beta, self.alpha)


class SixPoint:
    """IBM_kernels::GaussianFlexible::sixPoint (misc/IBM_kernels.cuh:163-237): the C3 six-point kernel of Bao, Kaye and
    Peskin; support 6."""
    support = 6

    def __init__(self, h, tolerance=None):
        self.h = float(h)

    def struct(self):
        return IBMKernelStruct(4, 6, self.h, 0.0, 0.0, 0.0)


def _prec(dtype):
    if dtype == torch.float64:
        return 8
    if dtype == torch.float32:
        return 4
    raise UB200Error("precision must be torch.float32 or torch.float64")


class FFT3D:
    """In-place batched-3 real FFT on the padded real3 grid [nz, ny, 2(nx/2+1), 3]."""

    def __init__(self, nx, ny, nz, dtype=torch.float64):
        self.lib = _declare()
        self.n = (nx, ny, nz)
        self.dtype = dtype
        self._h = C.c_void_p()
        check(self.lib.ub200_fft3d_create(C.byref(self._h), _prec(dtype), nx, ny, nz))

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_fft3d_destroy(self._h)
        except Exception:
            pass

    def _check(self, grid):
        nx, ny, nz = self.n
        if tuple(grid.shape) != (nz, ny, 2 * (nx // 2 + 1), 3) or grid.dtype != self.dtype or not grid.is_contiguous():
            raise UB200Error("FFT3D: grid must be a contiguous [nz, ny, 2(nx/2+1), 3] tensor of the plan precision")

    def forward(self, grid, stream=None):
        """real3 -> complex3 in place; returns a complex view [nz, ny, nx/2+1, 3]."""
        self._check(grid)
        check(self.lib.ub200_fft3d_exec(self._h, _ptr(grid), -1, _stream_ptr(stream)))
        nx, ny, nz = self.n
        return torch.view_as_complex(grid.view(nz, ny, nx // 2 + 1, 3, 2))

    def inverse(self, grid, stream=None):
        self._check(grid)
        check(self.lib.ub200_fft3d_exec(self._h, _ptr(grid), 1, _stream_ptr(stream)))
        return grid


class IBM:
    """IBM<Kernel>(kernel, grid, LinearIndex3D(nxPad, ny, nz)): spread / gather accumulate like the reference."""

    def __init__(self, kernel, L, cells, nxPad=None, periodic=(1, 1, 1), dtype=torch.float64):
        self.lib = _declare()
        self.kernel, self.cells, self.dtype = kernel, tuple(int(c) for c in cells), dtype
        self.nxPad = int(nxPad) if nxPad is not None else self.cells[0]
        L = (L, L, L) if np.isscalar(L) else L
        self._h = C.c_void_p()
        ks = kernel.struct()
        check(self.lib.ub200_ibm_create(C.byref(self._h), _prec(dtype), d3(L), i3([int(p) for p in periodic]),
                                        i3(self.cells), C.byref(ks), self.nxPad))

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_ibm_destroy(self._h)
        except Exception:
            pass

    def newGrid(self, device):
        nx, ny, nz = self.cells
        return torch.zeros(nz, ny, self.nxPad, 3, dtype=self.dtype, device=device)

    def spread(self, pos, values, grid, stream=None, overwrite=False):
        fn = self.lib.ub200_ibm_spread_overwrite if overwrite else self.lib.ub200_ibm_spread
        check(fn(self._h, _ptr(pos), _ptr(values), values.shape[1], pos.shape[0], _ptr(grid), _stream_ptr(stream)))

    def gather(self, pos, grid, out, stream=None):
        check(self.lib.ub200_ibm_gather(self._h, _ptr(pos), pos.shape[0], _ptr(grid), _ptr(out), _stream_ptr(stream)))


def hasimotoSelfMobility(hydrodynamicRadius, viscosity, L):
    """FCM_impl::getSelfMobility (FCM_impl.cuh:102-119): periodic correction to O(a^6)."""
    a = hydrodynamicRadius / L
    a3 = a ** 3
    c = 2.83729747948061947666591710460773907
    b = 0.19457
    a6pref = 16.0 * math.pi ** 2 / 45.0 + 630.0 * b * b
    return 1.0 / (6.0 * math.pi * viscosity * hydrodynamicRadius) * (1.0 - c * a + (4.0 / 3.0) * math.pi * a3 - a6pref * a3 * a3)


class FCM_impl:
    """FCM_impl<Kernel, KernelTorque> without torques. Parameters mirror FCM_impl::Parameters."""

    def __init__(self, box, cells, kernel, viscosity, hydrodynamicRadius=None, seed=0, dtype=torch.float64, kernelTorque=None):
        self.lib = _declare()
        L = (box, box, box) if np.isscalar(box) else tuple(box)
        if L[0] <= 0:
            raise UB200Error("FCM_impl requires a valid box")
        if cells[0] <= 0:
            raise UB200Error("FCM_impl requires a valid grid dimension")
        if kernel is None:
            raise UB200Error("FCM_impl requires instances of the spreading kernels")
        self.L, self.cells, self.kernel, self.viscosity, self.dtype = L, tuple(int(c) for c in cells), kernel, viscosity, dtype
        h = min(L[d] / cells[d] for d in range(3))
        self.hydrodynamicRadius = hydrodynamicRadius if hydrodynamicRadius is not None else kernel.fixHydrodynamicRadius(0, h)
        if seed == 0:
            seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFF) or 1
        self.seed = seed & 0xFFFFFFFF
        self._h = C.c_void_p()
        ks = kernel.struct()
        check(self.lib.ub200_fcm_create(C.byref(self._h), _prec(dtype), d3(L), i3(self.cells), C.byref(ks),
                                        float(viscosity), self.seed))
        self.kernelTorque = kernelTorque
        if kernelTorque is not None:
            kt = kernelTorque.struct()
            check(self.lib.ub200_fcm_set_torque_kernel(self._h, C.byref(kt)))

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_fcm_destroy(self._h)
        except Exception:
            pass

    def getHydrodynamicRadius(self):
        return self.hydrodynamicRadius

    def getSelfMobility(self):
        return hasimotoSelfMobility(self.hydrodynamicRadius, self.viscosity, self.L[0])

    def computeHydrodynamicDisplacements(self, pos, force, N=None, temperature=0.0, prefactor=0.0, out=None, stream=None,
                                         torque=None):
        """pos, force, torque: real4 [N,4]; returns the real3 [N,3] linear displacements, or the pair (linear, angular)
        when torques are given (FCM_impl::computeHydrodynamicDisplacements returns the same pair)."""
        N = pos.shape[0] if N is None else N
        if out is None:
            out = torch.empty(N, 3, dtype=self.dtype, device=pos.device)
        if torque is not None:
            if self.kernelTorque is None:
                raise UB200Error("FCM_impl requires a torque kernel to compute angular displacements")
            ang = torch.empty(N, 3, dtype=self.dtype, device=pos.device)
            check(self.lib.ub200_fcm_mdot_torque(self._h, _ptr(pos), _ptr(force), _ptr(torque), N, float(temperature),
                                                 float(prefactor), _ptr(out), _ptr(ang), _stream_ptr(stream)))
            return out, ang
        check(self.lib.ub200_fcm_mdot(self._h, _ptr(pos), _ptr(force), N, float(temperature), float(prefactor),
                                      _ptr(out), _stream_ptr(stream)))
        return out

    def velocityGrid(self):
        """Copy of the internal grid buffer (real-space velocities after the last call)."""
        cells, nxPad, ptr = i3((0, 0, 0)), C.c_int(0), C.c_void_p()
        check(self.lib.ub200_fcm_grid_info(self._h, cells, C.byref(nxPad), C.byref(ptr)))
        from .md import _device_copy
        return _device_copy(ptr.value, (cells[2], cells[1], nxPad.value, 3), self.dtype, torch.device("cuda"))


class FCM:
    """BDHI Method (BDHI_FCM.cuh:85-153): computeMF = computeHydrodynamicDisplacements(T, prefactor = 1/sqrt(dt));
    computeBdW is a no-op (noise is already in MF)."""

    def __init__(self, box, cells, kernel, viscosity, temperature, dt, seed=0, dtype=torch.float64):
        self.impl = FCM_impl(box, cells, kernel, viscosity, seed=seed, dtype=dtype)
        self.temperature, self.dt = temperature, dt

    def computeMF(self, pos, force, MF, stream=None):
        return self.impl.computeHydrodynamicDisplacements(pos, force, temperature=self.temperature,
                                                          prefactor=1.0 / math.sqrt(self.dt), out=MF, stream=stream)

    def computeBdW(self, BdW, stream=None):
        return None

    def getHydrodynamicRadius(self):
        return self.impl.getHydrodynamicRadius()

    def getSelfMobility(self):
        return self.impl.getSelfMobility()


class EulerMaruyama:
    """BDHI::EulerMaruyama<Method>::forwardTime (BDHI_EulerMaruyama.cu:125-166) with external forces supplied by a
    callable force(pos) -> real4 tensor (the interactors); positions updated as x += MF dt (noise lives in MF for
    FCM, integrateGPUD :82-113 with BdW = nullptr semantics)."""

    def __init__(self, method, pos, dt, forceFn=None):
        self.method, self.pos, self.dt, self.forceFn = method, pos, float(dt), forceFn
        self.MF = torch.zeros(pos.shape[0], 3, dtype=pos.dtype, device=pos.device)
        self.force = torch.zeros_like(pos)
        self.steps = 0

    def forwardTime(self):
        self.steps += 1
        if self.forceFn is not None:
            self.force = self.forceFn(self.pos)
        self.method.computeMF(self.pos, self.force, self.MF)
        T = getattr(self.method, "temperature", 0.0)
        check(_declare().ub200_bdhi_euler_update(_prec(self.pos.dtype), _ptr(self.pos), None, _ptr(self.MF), None, None,
                                                 self.pos.shape[0], math.sqrt(2 * self.dt * T), self.dt, 0,
                                                 _stream_ptr()))
