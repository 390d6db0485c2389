"""Host-side mirror of the reference's operator interface for path 1 (short-range pair forces).

Same names, argument meaning and error behaviour as the reference classes they stand in for:
  Box                  utils/Box.cuh:15-98
  LJ                   Interactor/Potential/Potential.cuh:25-86 (Potential::LJ = Radial<LJFunctor>)
  CellList             Interactor/NeighbourList/CellList.cuh:66-208 (+ CellListBase::CellListData)
  PairForces           Interactor/PairForces.cuh:23-68, PairForces.cu:43-78 (Interactor::sum)
  VerletNVE            Integrator/VerletNVE.cu:174-188 (Integrator::forwardTime)
Everything computes through the C ABI (include/uammd_b200.h); torch only owns device memory/streams.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import UB200Error, check, f3, i3


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_cudart = None


def _memcpy_d2d(dst_ptr, src_ptr, nbytes):
    global _cudart
    if _cudart is None:
        _cudart = C.CDLL("libcudart.so.12")
        _cudart.cudaMemcpy.restype = C.c_int
        _cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    err = _cudart.cudaMemcpy(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), nbytes, 3)
    if err != 0:
        raise UB200Error(f"cudaMemcpy D2D failed with cudaError {err}")


def _device_copy(ptr, shape, dtype, device):
    out = torch.empty(shape, dtype=dtype, device=device)
    if out.numel():
        torch.cuda.current_stream().synchronize()
        _memcpy_d2d(out.data_ptr(), ptr, out.numel() * out.element_size())
    return out


class Box:
    """utils/Box.cuh: box sizes and per-dimension periodicity."""

    def __init__(self, L):
        if np.isscalar(L):
            L = (L, L, L)
        self.boxSize = tuple(float(np.float32(x)) for x in L)
        self.periodic = [not (x == 0.0 or math.isinf(x)) for x in self.boxSize]

    def setPeriodicity(self, x, y, z):
        self.periodic = [p and bool(f) for p, f in zip(self.periodic, (x, y, z))]

    def __eq__(self, other):
        return isinstance(other, Box) and self.boxSize == other.boxSize and self.periodic == other.periodic


class LJ:
    """Potential::LJ. setPotParameters(ti, tj, cutOff=, sigma=, epsilon=, shift=) like
    Radial::setPotParameters (RadialPotential.cuh:75-81); the table rows are LJFunctor::PairParameters
    {cutOff2, sigma2, epsilonDivSigma2, shift} computed as processPairParameters does (Potential.cuh:67-82)."""

    def __init__(self):
        self._pairs = {}
        self.ntypes = 0

    def setPotParameters(self, ti, tj, cutOff, sigma=1.0, epsilon=1.0, shift=False):
        f = np.float32
        cutOff, sigma, epsilon = f(cutOff), f(sigma), f(epsilon)
        cutOff2 = f(cutOff * cutOff)
        sigma2 = f(sigma * sigma)
        eds2 = f(epsilon / sigma2)
        sh = f(0.0)
        if shift:
            ic2 = f(sigma2 / cutOff2)
            ic6 = f(f(ic2 * ic2) * ic2)
            sh = f(f(f(epsilon * f(4.0)) * ic6) * f(ic6 - f(1.0)))
        row = (cutOff2, sigma2, eds2, sh, cutOff)
        self._pairs[(ti, tj)] = row
        self._pairs[(tj, ti)] = row
        self.ntypes = max(self.ntypes, ti + 1, tj + 1)

    def getCutOff(self):
        if not self._pairs:
            raise UB200Error("LJ: no pair parameters set")
        return float(max(r[4] for r in self._pairs.values()))

    def table(self):
        n = self.ntypes
        t = np.zeros((n, n, 4), dtype=np.float32)
        for (a, b), r in self._pairs.items():
            t[a, b] = r[:4]
        return np.ascontiguousarray(t.reshape(-1))


class CellList:
    """NeighbourList concept: update(pos, box, cutOff) rebuilds the list on every call (the reference's
    force_next_update is never cleared), getCellList() exposes CellListData."""

    def __init__(self):
        self._h = C.c_void_p()
        check(_lib.lib().ub200_celllist_create(C.byref(self._h)))
        self.N = 0
        self.device = None

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ub200_celllist_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def gridFor(box, cutOff):
        cd = i3((0, 0, 0))
        check(_lib.lib().ub200_neighbour_celldim_f32(f3(box.boxSize), float(cutOff), cd))
        return tuple(cd)

    def update(self, pos, box, cutOff, stream=None, cellDim=None, groupIndex=None):
        if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 4 or not pos.is_cuda:
            raise UB200Error("CellList.update: pos must be a CUDA float32 [N,4] tensor (real4)")
        if not pos.is_contiguous():
            raise UB200Error("CellList.update: pos must be contiguous")
        N = pos.shape[0] if groupIndex is None else groupIndex.shape[0]
        cd = cellDim if cellDim is not None else self.gridFor(box, cutOff)
        check(_lib.lib().ub200_celllist_build_f32(self._h, _ptr(pos), _ptr(groupIndex), N, f3(box.boxSize),
                                                  i3([int(p) for p in box.periodic]), i3(cd), _stream_ptr(stream)))
        self.N, self.device, self.cellDim = N, pos.device, tuple(cd)

    def view(self):
        v = _lib.CellListView()
        check(_lib.lib().ub200_celllist_view_get(self._h, C.byref(v)))
        return v

    def errorFlag(self):
        flag = C.c_int(0)
        check(_lib.lib().ub200_celllist_error_flag(self._h, _stream_ptr(), C.byref(flag)))
        return flag.value

    def getCellList(self):
        """Copies of the CellListData arrays (CellListBase.cuh:145-160) as torch tensors."""
        v = self.view()
        ncells = v.cellDim[0] * v.cellDim[1] * v.cellDim[2]
        dev = self.device
        return {
            "cellStart": _device_copy(v.d_cellStart, (ncells,), torch.int32, dev),  # uint32 bit pattern
            "cellEnd": _device_copy(v.d_cellEnd, (ncells,), torch.int32, dev),
            "sortPos": _device_copy(v.d_sortPos, (v.numberParticles, 4), torch.float32, dev),
            "groupIndex": _device_copy(v.d_groupIndex, (v.numberParticles,), torch.int32, dev),
            "VALID_CELL": int(v.VALID_CELL),
            "cellDim": tuple(v.cellDim),
            "binStart": _device_copy(v.d_binStart, (v.nbins + 1,), torch.int32, dev),
        }

    def normalizedCells(self):
        """cellStart/cellEnd with the VALID_CELL epoch removed: -1 for empty cells (parity helper)."""
        d = self.getCellList()
        cs = d["cellStart"].cpu().numpy().view(np.uint32).astype(np.int64)
        ce = d["cellEnd"].cpu().numpy().astype(np.int64)
        empty = cs < d["VALID_CELL"]
        cs = np.where(empty, -1, cs - d["VALID_CELL"])
        ce = np.where(empty, -1, ce)
        return cs, ce


class VerletListView(C.Structure):
    _fields_ = [("d_neighbourList", C.c_void_p), ("d_numberNeighbours", C.c_void_p), ("d_sortPos", C.c_void_p),
                ("d_groupIndex", C.c_void_p), ("particleStride", C.c_int), ("numberParticles", C.c_int),
                ("maxNeighboursPerParticle", C.c_int), ("stepsSinceLastUpdate", C.c_int), ("rebuilds", C.c_int)]


class VerletListRows(C.Structure):
    _fields_ = [("d_list", C.c_void_p), ("d_count", C.c_void_p), ("d_pos", C.c_void_p), ("d_index", C.c_void_p),
                ("stride", C.c_int), ("numberParticles", C.c_int), ("indexBits", C.c_int)]


def _declare_verlet():
    lib = _lib.lib()
    if getattr(lib, "_verlet_declared", False):
        return lib
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.ub200_verletlist_create.argtypes = [C.POINTER(vp)]
    lib.ub200_verletlist_destroy.argtypes = [vp]
    lib.ub200_verletlist_set_cutoff_multiplier.argtypes = [vp, f]
    lib.ub200_verletlist_update_f32.argtypes = [vp, vp, vp, i, C.c_float * 3, C.c_int * 3, f, i, C.POINTER(i), vp]
    lib.ub200_verletlist_view_get.argtypes = [vp, C.POINTER(VerletListView)]
    lib.ub200_lj_sum_verlet_f32.argtypes = [vp, C.POINTER(C.c_float), i, vp, vp, vp, vp, vp]
    lib.ub200_verletlist_stats.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    lib.ub200_verletlist_rows_get.argtypes = [vp, C.POINTER(VerletListRows)]
    lib._verlet_declared = True
    return lib


class VerletList:
    """NeighbourList concept with a skin (Interactor/NeighbourList/VerletList.cuh:83-201): update(pos, box, cutOff)
    rebuilds only when some particle moved (multiplier - 1) cutOff / 2 since the last rebuild; getVerletList() exposes
    VerletListData (neighbourList [k*N + i] over sorted indices, numberNeighbours, sortPos, groupIndex)."""

    def __init__(self):
        self.lib = _declare_verlet()
        self._h = C.c_void_p()
        check(self.lib.ub200_verletlist_create(C.byref(self._h)))
        self.device = None
        self.forceNextUpdate = True

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_verletlist_destroy(self._h)
        except Exception:
            pass

    def setCutOffMultiplier(self, m):
        check(self.lib.ub200_verletlist_set_cutoff_multiplier(self._h, float(m)))

    def update(self, pos, box, cutOff, stream=None, groupIndex=None):
        if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 4 or not pos.is_cuda or not pos.is_contiguous():
            raise UB200Error("VerletList.update: pos must be a contiguous CUDA float32 [N,4] tensor (real4)")
        N = pos.shape[0] if groupIndex is None else groupIndex.shape[0]
        rebuilt = C.c_int(0)
        check(self.lib.ub200_verletlist_update_f32(self._h, _ptr(pos), _ptr(groupIndex), N, f3(box.boxSize),
                                                   i3([int(p) for p in box.periodic]), float(cutOff),
                                                   int(self.forceNextUpdate), C.byref(rebuilt), _stream_ptr(stream)))
        self.forceNextUpdate = False
        self.device = pos.device
        return bool(rebuilt.value)

    def view(self):
        v = VerletListView()
        check(self.lib.ub200_verletlist_view_get(self._h, C.byref(v)))
        return v

    def getNumberOfStepsSinceLastUpdate(self):
        steps = C.c_int(0)
        check(self.lib.ub200_verletlist_stats(self._h, C.byref(steps), None))
        return steps.value

    def rebuilds(self):
        n = C.c_int(0)
        check(self.lib.ub200_verletlist_stats(self._h, None, C.byref(n)))
        return n.value

    def getRowList(self):
        """The row list of the built-in LJ traversal (ub200_verletlist_rows_get): list [N, stride] over half-cell slots
        (entry = slot | image << indexBits), count [N], pos [N, 4] in half-cell order, index [N] slot -> group index."""
        r = VerletListRows()
        check(self.lib.ub200_verletlist_rows_get(self._h, C.byref(r)))
        N, dev = r.numberParticles, self.device
        return {"list": _device_copy(r.d_list, (N, r.stride), torch.int32, dev),
                "count": _device_copy(r.d_count, (N,), torch.int32, dev),
                "pos": _device_copy(r.d_pos, (N, 4), torch.float32, dev),
                "index": _device_copy(r.d_index, (N,), torch.int32, dev), "stride": r.stride, "indexBits": r.indexBits}

    def getVerletList(self):
        v = self.view()
        N, dev = v.numberParticles, self.device
        return {
            "neighbourList": _device_copy(v.d_neighbourList, (v.maxNeighboursPerParticle + 1, N), torch.int32, dev),
            "numberNeighbours": _device_copy(v.d_numberNeighbours, (N,), torch.int32, dev),
            "sortPos": _device_copy(v.d_sortPos, (N, 4), torch.float32, dev),
            "groupIndex": _device_copy(v.d_groupIndex, (N,), torch.int32, dev),
            "particleStride": v.particleStride, "maxNeighboursPerParticle": v.maxNeighboursPerParticle,
            "rebuilds": v.rebuilds,
        }


def sortParticles(box, hashCutOff, pos, *properties, stream=None):
    """ParticleData::sortParticles (ParticleData/ParticleData.cuh:492-522): reorders the particles by the Morton hash of their cell
    (cells of box / hashCutOff) for memory locality. Returns (order, pos_sorted, *properties_sorted) where order[k] is the old
    index of the particle now at k (apply it to `id`-like arrays to keep track, as the reference's originalOrderIndex does).
    Properties: contiguous CUDA tensors with one row per particle and 4-byte elements."""
    lib = _lib.lib()
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.ub200_particles_sort_order_f32.argtypes = [vp, vp, i, C.c_float * 3, C.c_int * 3, f, vp, vp]
    lib.ub200_apply_order.argtypes = [vp, vp, vp, i, i, vp]
    if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 4 or not pos.is_cuda or not pos.is_contiguous():
        raise UB200Error("sortParticles: pos must be a contiguous CUDA float32 [N,4] tensor (real4)")
    N = pos.shape[0]
    scratch = CellList()
    order = torch.empty(N, dtype=torch.int32, device=pos.device)
    check(lib.ub200_particles_sort_order_f32(scratch._h, _ptr(pos), N, f3(box.boxSize), i3([int(p) for p in box.periodic]),
                                             float(hashCutOff), _ptr(order), _stream_ptr(stream)))
    out = []
    for t in (pos,) + tuple(properties):
        if not (t.is_cuda and t.is_contiguous() and t.shape[0] == N and t.element_size() == 4):
            raise UB200Error("sortParticles: properties must be contiguous CUDA tensors of 4-byte elements with N rows")
        o = torch.empty_like(t)
        check(lib.ub200_apply_order(_ptr(t), _ptr(o), _ptr(order), N, t.element_size() * (t.numel() // N), _stream_ptr(stream)))
        out.append(o)
    return (order,) + tuple(out)


class PairForcesLJ64:
    """PairForces<Potential::LJ, CellList>::sum of a `real = double` build of the reference (ub200_lj_sum_f64): positions
    float64 [N,4] (w = type), forces float64 [N,4] / energies / virials float64 [N] accumulated like Transverser::set.
    `pot` is the same LJ object as for the single precision path; its table rows are widened to double (the reference's
    processPairParameters would compute them in double from the same inputs: identical whenever cutOff, sigma and epsilon
    are exact in single precision, as in every BASELINE configuration)."""

    def __init__(self, box, pot):
        self.box, self.pot = box, pot
        self._h = C.c_void_p()
        l = _lib.lib()
        l.ub200_lj64_create.restype = C.c_int
        l.ub200_lj64_create.argtypes = [C.POINTER(C.c_void_p)]
        l.ub200_lj64_destroy.argtypes = [C.c_void_p]
        l.ub200_lj_sum_f64.restype = C.c_int
        l.ub200_lj_sum_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double * 3, C.c_int * 3, C.c_double,
                                       C.POINTER(C.c_double), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        check(l.ub200_lj64_create(C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ub200_lj64_destroy(self._h)
        except Exception:
            pass

    def sum(self, pos, force=None, energy=None, virial=None, stream=None):
        if pos.dtype != torch.float64 or pos.dim() != 2 or pos.shape[1] != 4 or not pos.is_cuda or not pos.is_contiguous():
            raise UB200Error("PairForcesLJ64.sum: pos must be a contiguous CUDA float64 [N,4] tensor (real4)")
        table = np.ascontiguousarray(self.pot.table().astype(np.float64))
        L = (C.c_double * 3)(*[float(x) for x in self.box.boxSize])
        check(_lib.lib().ub200_lj_sum_f64(self._h, _ptr(pos), pos.shape[0], L, i3([int(p) for p in self.box.periodic]),
                                          float(self.pot.getCutOff()), table.ctypes.data_as(C.POINTER(C.c_double)),
                                          self.pot.ntypes, _ptr(force), _ptr(energy), _ptr(virial), _stream_ptr(stream)))


class LJEngine:
    """ub200_ljengine: PairForces<Potential::LJ, CellList>::sum in one call (Interactor/PairForces.cu:43-78) over the
    engine's private half-cell list (column traversal, uammd_b200/csrc/lj_column.cu)."""

    PATHS = {0: "column", 1: "cell", 2: "nbody", -1: "none"}

    def __init__(self):
        self._h = C.c_void_p()
        check(_lib.lib().ub200_ljengine_create(C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ub200_ljengine_destroy(self._h)
        except Exception:
            pass

    def sum(self, pos, box, table, ntypes, force=None, energy=None, virial=None, stream=None, groupIndex=None,
            globalIndex=None, accumulate=True, owner=(0, 0x7fffffff)):
        if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 4 or not pos.is_cuda or not pos.is_contiguous():
            raise UB200Error("PairForces.sum: pos must be a contiguous CUDA float32 [N,4] tensor (real4)")
        N = pos.shape[0] if groupIndex is None else groupIndex.shape[0]
        check(_lib.lib().ub200_ljengine_sum_f32(self._h, _ptr(pos), _ptr(groupIndex), N, f3(box.boxSize),
                                                i3([int(p) for p in box.periodic]), table.ctypes.data_as(C.POINTER(C.c_float)),
                                                ntypes, _ptr(force), _ptr(energy), _ptr(virial), _ptr(globalIndex),
                                                int(accumulate), int(owner[0]), int(owner[1]), _stream_ptr(stream)))

    def traverse(self, force, accumulate=True, stream=None):
        """The traversal kernel alone over the list of the last sum() (kernel timing)."""
        check(_lib.lib().ub200_ljengine_traverse_f32(self._h, _ptr(force), int(accumulate), _stream_ptr(stream)))

    def lastPath(self):
        return self.PATHS[_lib.lib().ub200_ljengine_last_path(self._h)]

    def grid(self):
        cd = i3((0, 0, 0))
        check(_lib.lib().ub200_ljengine_grid(self._h, cd))
        return tuple(cd)

    def errorFlag(self):
        flag = C.c_int(0)
        check(_lib.lib().ub200_ljengine_error_flag(self._h, _stream_ptr(), C.byref(flag)))
        return flag.value


class PairForces:
    """Interactor: sum(pos, force=, energy=, virial=) accumulates (+=) like Transverser::set. nl: the second template
    argument of the reference's PairForces - None (default) lets the engine search neighbours its own way (LJEngine: the
    forces are the reference's, the list is private); a CellList or VerletList instance keeps the reference-layout list,
    which getCellList() / getVerletList() expose. With the default, `nl` is a CellList built on demand from the positions
    of the last sum()."""

    def __init__(self, potential, box, nl=None):
        self.pot, self.box = potential, box
        self._nl = nl
        self._engine = LJEngine() if nl is None else None
        self._last = None

    @property
    def nl(self):
        if self._nl is None:
            self._nl = CellList()
        if self._engine is not None and self._last is not None:
            pos, stream = self._last
            self._nl.update(pos, self.box, self.pot.getCutOff(), stream)
            self._last = None
        return self._nl

    def updateBox(self, box):
        self.box = box

    def sum(self, pos, force=None, energy=None, virial=None, stream=None, globalIndex=None):
        if self._engine is not None:
            self._engine.sum(pos, self.box, self.pot.table(), self.pot.ntypes, force, energy, virial, stream,
                             globalIndex=globalIndex)
            self._last = (pos, stream)
            return
        rcut = self.pot.getCutOff()
        L = self.box.boxSize
        if all(l <= 3 * rcut for l in L):
            # PairForces.cu:49-53,61-66: no neighbour list for such a box, NBody::transverse over all pairs
            self.sumNBody(pos, force, energy, virial, stream, globalIndex)
            return
        self.nl.update(pos, self.box, rcut, stream)
        self.sumWithCurrentList(force, energy, virial, stream, globalIndex)

    def sumNBody(self, pos, force=None, energy=None, virial=None, stream=None, globalIndex=None):
        """NBody::transverse (Interactor/NBodyBase.cuh:46-116) with the LJ transverser: all pairs, minimum image."""
        lib = _lib.lib()
        lib.ub200_lj_nbody_f32.restype = C.c_int
        lib.ub200_lj_nbody_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float * 3, C.c_int * 3, C.POINTER(C.c_float),
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 4 or not pos.is_cuda or not pos.is_contiguous():
            raise UB200Error("PairForces.sumNBody: pos must be a contiguous CUDA float32 [N,4] tensor (real4)")
        tab = self.pot.table()
        N = pos.shape[0] if globalIndex is None else globalIndex.shape[0]
        check(lib.ub200_lj_nbody_f32(_ptr(pos), _ptr(globalIndex), N, f3(self.box.boxSize),
                                     i3([int(p) for p in self.box.periodic]), tab.ctypes.data_as(C.POINTER(C.c_float)),
                                     self.pot.ntypes, _ptr(force), _ptr(energy), _ptr(virial), _stream_ptr(stream)))

    def sumWithCurrentList(self, force=None, energy=None, virial=None, stream=None, globalIndex=None):
        tab = self.pot.table()
        if isinstance(self.nl, VerletList):
            check(self.nl.lib.ub200_lj_sum_verlet_f32(self.nl._h, tab.ctypes.data_as(C.POINTER(C.c_float)), self.pot.ntypes,
                                                      _ptr(force), _ptr(energy), _ptr(virial), _ptr(globalIndex),
                                                      _stream_ptr(stream)))
            return
        check(_lib.lib().ub200_lj_sum_f32(self.nl._h, tab.ctypes.data_as(C.POINTER(C.c_float)), self.pot.ntypes,
                                          _ptr(force), _ptr(energy), _ptr(virial), _ptr(globalIndex),
                                          _stream_ptr(stream)))


class DPD:
    """Potential::DPD parameters (Interactor/Potential/DPD.cuh:51-92). sigma = sqrt(2 T)/sqrt(dt)."""

    def __init__(self, cutOff=1.0, dt=0.01, gamma=1.0, temperature=1.0, A=1.0, seed=0):
        self.rcut, self.dt, self.gamma, self.temperature, self.A = cutOff, dt, gamma, temperature, A
        self.seed, self.step = int(seed) & 0xFFFFFFFF, 0
        self._updateSigma()

    def _updateSigma(self):
        # sigma = sqrt(2.0 * temperature) / sqrt(dt) with `real` members (DPD.cuh:66): a double square root over a float one
        f = np.float32
        self.sigma = float(f(np.sqrt(2.0 * float(f(self.temperature))) / float(np.sqrt(f(self.dt)))))

    def getCutOff(self):
        return self.rcut

    def updateTemperature(self, T):
        self.temperature = T
        self._updateSigma()

    def updateTimeStep(self, dt):
        self.dt = dt
        self._updateSigma()


class PairForcesDPD:
    """PairForces<Potential::DPD> through a getTransverser adaptor (the stock Potential::DPD only exposes the
    pre-v2 getForceTransverser and is silently skipped by PairForces). step is incremented before
    every force evaluation like DPD_impl::getForceTransverser (DPD.cuh:161-170)."""

    def __init__(self, potential, box, nl=None):
        self.pot, self.box = potential, box
        self.nl = nl if nl is not None else CellList()

    def sum(self, pos, vel, force, stream=None, globalIndex=None):
        self.nl.update(pos, self.box, self.pot.getCutOff(), stream)
        self.pot.step += 1
        p = self.pot
        check(_lib.lib().ub200_dpd_sum_f32(self.nl._h, _ptr(vel), p.A, p.gamma, p.sigma, p.rcut, p.seed,
                                           p.step & 0xFFFFFFFF, pos.shape[0], _ptr(force), _ptr(globalIndex),
                                           _stream_ptr(stream)))


class VerletNVE:
    """Integrator: forwardTime() = kick+drift, reset forces, sum interactors, kick (VerletNVE.cu:174-188)."""

    def __init__(self, pos, vel, dt, mass=1.0, is2D=False):
        self.pos, self.vel, self.dt, self.mass, self.is2D = pos, vel, float(dt), float(mass), bool(is2D)
        self.force = torch.zeros_like(pos)
        self.interactors = []
        self.steps = 0

    def addInteractor(self, it):
        self.interactors.append(it)

    def _half(self, step):
        N = self.pos.shape[0]
        check(_lib.lib().ub200_nve_half_step_f32(_ptr(self.pos), _ptr(self.vel), _ptr(self.force), C.c_void_p(0),
                                                 self.mass, C.c_void_p(0), N, self.dt, int(self.is2D), step,
                                                 _stream_ptr()))

    def _sumForces(self):
        self.force.zero_()
        for it in self.interactors:
            it.sum(self.pos, force=self.force)

    def forwardTime(self):
        self.steps += 1
        if self.steps == 1:
            self._sumForces()
        self._half(1)
        self._sumForces()
        self._half(2)


class LJMD:
    """Fused engine for VerletNVE + PairForces<LJ, CellList> (ub200_md_* in include/uammd_b200.h)."""

    def __init__(self, box, potential, dt):
        self.box, self.pot, self.dt = box, potential, float(dt)
        self._h = C.c_void_p()
        check(_lib.lib().ub200_md_create(C.byref(self._h)))
        self._tab = potential.table()
        self._tabp = self._tab.ctypes.data_as(C.POINTER(C.c_float))
        self.prepared = False

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ub200_md_destroy(self._h)
        except Exception:
            pass

    def prepare(self, pos, force, stream=None):
        check(_lib.lib().ub200_md_lj_nve_prepare_f32(self._h, _ptr(pos), _ptr(force), pos.shape[0],
                                                     f3(self.box.boxSize), self.pot.getCutOff(), self._tabp,
                                                     self.pot.ntypes, _stream_ptr(stream)))
        self.prepared = True

    def run(self, pos, vel, force, nsteps, stream=None):
        if not self.prepared:
            self.prepare(pos, force, stream)
        check(_lib.lib().ub200_md_lj_nve_run_f32(self._h, _ptr(pos), _ptr(vel), _ptr(force), pos.shape[0],
                                                 f3(self.box.boxSize), self.pot.getCutOff(), self._tabp,
                                                 self.pot.ntypes, self.dt, int(nsteps), _stream_ptr(stream)))

    def kineticEnergy(self, vel, out, index=0, stream=None):
        """sum v^2 / 2 of vel [N,3] -> out[index] (a pinned float64 host tensor), asynchronously on the stream."""
        check(_lib.lib().ub200_md_kinetic_energy_f32(self._h, _ptr(vel), vel.shape[0], C.c_void_p(out.data_ptr() + 8 * index),
                                                     _stream_ptr(stream)))

    def runVerlet(self, nl, pos, vel, force, nsteps, forcesAreCurrent=False, stream=None):
        """The same loop over PairForces<LJ, VerletList> (nl: VerletList)."""
        lib = _declare_verlet()
        lib.ub200_md_lj_nve_verlet_run_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                       C.c_float * 3, C.c_float, C.POINTER(C.c_float), C.c_int, C.c_float,
                                                       C.c_int, C.c_int, C.c_void_p]
        check(lib.ub200_md_lj_nve_verlet_run_f32(self._h, nl._h, _ptr(pos), _ptr(vel), _ptr(force), pos.shape[0],
                                                 f3(self.box.boxSize), self.pot.getCutOff(), self._tabp, self.pot.ntypes,
                                                 self.dt, int(nsteps), int(forcesAreCurrent), _stream_ptr(stream)))
        nl.device = pos.device

    def runHost(self, h_pos, h_vel, h_force, nsteps, stream=None):
        """Host (pinned) buffers in/out: H2D, prepare, nsteps, D2H, synchronise."""
        check(_lib.lib().ub200_md_lj_nve_run_host_f32(self._h, _ptr(h_pos), _ptr(h_vel), _ptr(h_force),
                                                      h_pos.shape[0], f3(self.box.boxSize), self.pot.getCutOff(),
                                                      self._tabp, self.pot.ntypes, self.dt, int(nsteps),
                                                      _stream_ptr(stream)))
