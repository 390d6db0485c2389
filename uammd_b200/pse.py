"""Host-side mirror of the reference's BDHI::PSE (Integrator/BDHI/BDHI_PSE.cuh:82-176) over the C ABI
(ub200_pse_*): same method names, argument meaning and seed-draw order, so the parity tests read like the
reference's own (test/BDHI/PSE/pse_test.cu). torch only owns device memory and streams; there is no CPU fallback.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import UB200Error, check
from .bd import System
from .md import _ptr, _stream_ptr


class PSEParams(C.Structure):
    _fields_ = [("L", C.c_double * 3), ("viscosity", C.c_double), ("hydrodynamicRadius", C.c_double),
                ("tolerance", C.c_double), ("psi", C.c_double), ("shearStrain", C.c_double),
                ("cellsOverride", C.c_int * 3)]


class IBMKernelStruct(C.Structure):
    _fields_ = [("kind", C.c_int), ("support", C.c_int), ("h", C.c_double), ("prefactor", C.c_double),
                ("tau", C.c_double), ("rmax", C.c_double)]


class PSEInfo(C.Structure):
    _fields_ = [("cells", C.c_int * 3), ("support", C.c_int), ("nTable", C.c_int), ("lastLanczosIterations", C.c_int),
                ("eta", C.c_double), ("rcut", C.c_double), ("d_table", C.c_void_p), ("d_grid", C.c_void_p),
                ("kernel", IBMKernelStruct)]


def _declare():
    lib = _lib.lib()
    if getattr(lib, "_pse_declared", False):
        return lib
    vp, i, d, u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint32
    sig = {
        "ub200_pse_create": (i, [C.POINTER(vp), i, C.POINTER(PSEParams), u32, u32]),
        "ub200_pse_destroy": (i, [vp]),
        "ub200_pse_info": (i, [vp, C.POINTER(PSEInfo)]),
        "ub200_pse_set_shear_strain": (i, [vp, d]),
        "ub200_pse_far_mdot": (i, [vp, vp, vp, i, d, d, u32, vp, vp]),
        "ub200_pse_near_mdot": (i, [vp, vp, vp, i, i, vp, vp]),
        "ub200_pse_near_noise": (i, [vp, vp, i, d, d, u32, vp, C.POINTER(i), vp]),
        "ub200_pse_near_noise_add": (i, [vp, vp, i, d, d, u32, vp, C.POINTER(i), vp]),
        "ub200_pse_near_mdot_list": (i, [vp, vp, vp, i, i, vp, vp]),
        "ub200_pse_near_noise_reuse": (i, [vp, vp, i, d, d, u32, vp, C.POINTER(i), vp]),
        "ub200_bdhi_euler_update": (i, [i, vp, vp, vp, vp, vp, i, d, d, i, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._pse_declared = True
    return lib


class Parameters:
    """pse_ns::Parameters (PSE/utils.cuh:17-24) on top of BDHI::Parameters (BDHI/BDHI.cuh:13-24)."""

    def __init__(self, box, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=1e-3, psi=0.5, shearStrain=0.0,
                 temperature=0.0, dt=1.0, cells=None):
        self.box = (box, box, box) if isinstance(box, (int, float)) else tuple(box)
        self.viscosity, self.hydrodynamicRadius, self.tolerance = viscosity, hydrodynamicRadius, tolerance
        self.psi, self.shearStrain, self.temperature, self.dt, self.cells = psi, shearStrain, temperature, dt, cells


def computeSelfMobility(par):
    """pse_ns::computeSelfMobility (PSE/initialization.cu:31-49): Hasimoto's periodic correction to O(a^8)."""
    rh, L = par.hydrodynamicRadius, par.box[0]
    a = rh / L
    a3 = a ** 3
    c, b = 2.83729747948061947666591710460773907, 0.19457
    a6pref = 16.0 * math.pi ** 2 / 45.0 + 630.0 * b * b
    return 1.0 / (6.0 * math.pi * par.viscosity * rh) * (1.0 - c * a + (4.0 / 3.0) * math.pi * a3 - a6pref * a3 * a3)


class PSE:
    """BDHI::PSE(pd, par). `pos` is the particle position tensor (real4 [N,4], float32 or float64 = the reference's
    `real`), `forceFn()` returns the current real4 force tensor (pd->getForce) or None."""

    def __init__(self, pos, par, sys=None, force=None):
        self.lib = _declare()
        assert pos.is_cuda and pos.dim() == 2 and pos.shape[1] == 4 and pos.is_contiguous()
        self.pos, self.par, self.force = pos, par, force
        self.N = pos.shape[0]
        self.sys = sys if sys is not None else System()
        if par.box[0] == 0 and par.box[1] == 0 and par.box[2] == 0:
            raise ValueError("Box of size zero detected")          # checkInputValidity (initialization.cu:11-29)
        if par.tolerance > 0.1:
            raise ValueError("Tolerance too high")
        self.M0 = computeSelfMobility(par)
        self.temperature, self.dt = par.temperature, par.dt
        # NearField draws its seed first, then FarField (initialization.cu:57-59, NearField.cuh:36, FarField.cuh:325)
        self.seedNear = self.sys.rng().next32()
        self.seedFar = self.sys.rng().next32()
        p = PSEParams()
        p.L[:] = [float(x) for x in par.box]
        p.viscosity, p.hydrodynamicRadius, p.tolerance = par.viscosity, par.hydrodynamicRadius, par.tolerance
        p.psi, p.shearStrain = par.psi, par.shearStrain
        p.cellsOverride[:] = list(par.cells) if par.cells is not None else [0, 0, 0]
        self._h = C.c_void_p()
        rc = self.lib.ub200_pse_create(C.byref(self._h), 8 if pos.dtype == torch.float64 else 4, C.byref(p),
                                       self.seedNear, self.seedFar)
        if rc == -1:
            raise UB200Error("[BDHI::PSE] invalid parameters (cut off larger than half the box? try increasing psi)")
        check(rc)

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_pse_destroy(self._h)
        except Exception:
            pass

    def info(self):
        inf = PSEInfo()
        check(self.lib.ub200_pse_info(self._h, C.byref(inf)))
        return inf

    # ---- BDHI Method concept ----
    def setup_step(self, stream=None):
        pass

    def finish_step(self, stream=None):
        pass

    def computeMF(self, MF, stream=None, listForNoise=False):
        """listForNoise: the near-field product runs over the Verlet list and leaves it for the computeBdW(reuseList=True)
        that follows on the same, unchanged positions (EulerMaruyama.forwardTime with T > 0)."""
        MF.zero_()
        self.computeMFFarField(MF, stream)
        self.computeMFNearField(MF, stream, listForNoise)

    def computeMFNearField(self, MF, stream=None, listForNoise=False):
        fn = self.lib.ub200_pse_near_mdot_list if listForNoise else self.lib.ub200_pse_near_mdot
        check(fn(self._h, _ptr(self.pos), _ptr(self.force) if self.force is not None else None,
                                           4, self.N, _ptr(MF), _stream_ptr(stream)))

    def computeMFFarField(self, MF, stream=None):
        seed2 = self.sys.rng().next32() if self.temperature > 0 else 0      # FarField::addBrownianNoise :478
        check(self.lib.ub200_pse_far_mdot(self._h, _ptr(self.pos), _ptr(self.force) if self.force is not None else None,
                                          self.N, float(self.temperature), 1.0 / math.sqrt(self.dt), seed2, _ptr(MF),
                                          _stream_ptr(stream)))

    def computeBdW(self, BdW, stream=None, reuseList=False):
        self._nearNoise(BdW, self.temperature, 1.0, stream, reuse=reuseList)

    def _nearNoise(self, out, temperature, prefactor, stream, add=False, reuse=False):
        if temperature == 0:
            return 0
        seed2 = self.sys.rng().next32()                                     # NearField.cuh:274
        it = C.c_int(0)
        fn = self.lib.ub200_pse_near_noise_add if add else (self.lib.ub200_pse_near_noise_reuse if reuse else self.lib.ub200_pse_near_noise)
        check(fn(self._h, _ptr(self.pos), self.N, float(temperature), float(prefactor), seed2, _ptr(out), C.byref(it),
                 _stream_ptr(stream)))
        return it.value

    def computeHydrodynamicDisplacements(self, force, MF, temperature, noise_prefactor, stream=None):
        """MF = Mobility force + noise_prefactor sqrt(2 T M) dW (BDHI_PSE.cuh:141-158, same call order and random draws).
        Deviation from the reference, on purpose: its Lanczos solver OVERWRITES the vector it is given (final gemv with
        beta = 0, LanczosAlgorithm.cu:163-172), so with a force and T > 0 the reference drops the near-field M F it has just
        computed; here the near-field noise is added to it, which is what the method documents."""
        MF.zero_()
        if force is not None:
            check(self.lib.ub200_pse_near_mdot(self._h, _ptr(self.pos), _ptr(force), 4, self.N, _ptr(MF), _stream_ptr(stream)))
        self._nearNoise(MF, temperature, noise_prefactor, stream, add=True)
        seed2 = self.sys.rng().next32() if temperature > 0 else 0
        check(self.lib.ub200_pse_far_mdot(self._h, _ptr(self.pos), _ptr(force) if force is not None else None, self.N,
                                          float(temperature), float(noise_prefactor), seed2, _ptr(MF), _stream_ptr(stream)))

    def setShearStrain(self, strain):
        check(self.lib.ub200_pse_set_shear_strain(self._h, float(strain)))

    def getHydrodynamicRadius(self):
        return self.par.hydrodynamicRadius

    def getSelfMobility(self):
        return self.M0


class EulerMaruyama:
    """BDHI::EulerMaruyama<PSE>::forwardTime (BDHI_EulerMaruyama.cu:125-166): dR = dt MF + sqrt(2 T dt) BdW."""

    def __init__(self, method, dt, temperature):
        self.method, self.dt, self.temperature = method, float(dt), float(temperature)
        pos = method.pos
        self.MF = torch.zeros(pos.shape[0], 3, dtype=pos.dtype, device=pos.device)
        self.BdW = torch.zeros_like(self.MF)
        self.steps = 0

    def forwardTime(self, stream=None):
        self.steps += 1
        m = self.method
        noisy = self.temperature > 0 and m.force is not None
        m.computeMF(self.MF, stream, listForNoise=noisy)
        if self.temperature > 0:
            m.computeBdW(self.BdW, stream, reuseList=noisy)
        check(m.lib.ub200_bdhi_euler_update(8 if m.pos.dtype == torch.float64 else 4, _ptr(m.pos), None, _ptr(self.MF),
                                            _ptr(self.BdW) if self.temperature > 0 else None, None, m.N,
                                            math.sqrt(2 * self.dt * self.temperature), self.dt, 0, _stream_ptr(stream)))
