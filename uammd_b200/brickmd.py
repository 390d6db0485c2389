"""Host-side mirror of the brick-decomposed pair path (ub200_brick_* / ub200_halo_exchange_* in include/uammd_b200.h):
VerletNVE + PairForces<LJ, CellList> (Integrator/VerletNVE.cu:174-188, Interactor/PairForces.cu:43-78) with the particles
of one system spread over the GPUs of one box. The reference is single-GPU; the oracle is the single-GPU engine
(uammd_b200.md.LJMD), whose trajectory the bricks reproduce bit for bit.

Everything on the step path is CUDA behind the C ABI (kick + drift + classification + peer-to-peer push, unpack, list
build on the rank's window, traversal, kick); torch.distributed is used ONCE, to move the 64-byte CUDA IPC handles.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import UB200Error, check, f3, i3
from .domain import choose_rank_grid
from .md import _device_copy, _ptr, _stream_ptr


class BrickInfo(C.Structure):
    _fields_ = [("d_pos", C.c_void_p), ("d_vel", C.c_void_p), ("d_gid", C.c_void_p), ("d_force", C.c_void_p),
                ("d_counts", C.c_void_p), ("capacity", C.c_int), ("rank", C.c_int), ("world", C.c_int),
                ("halfCells", C.c_int * 3), ("window", C.c_int * 3), ("windowOrigin", C.c_int * 3)]


def _declare():
    lib = _lib.lib()
    if getattr(lib, "_brick_declared", False):
        return lib
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    sig = {
        "ub200_brick_create": [C.POINTER(vp), i, C.c_int * 3, C.c_float * 3, C.c_int * 3, f, i, i],
        "ub200_brick_destroy": [vp],
        "ub200_comm_ipc_size": [],
        "ub200_brick_ipc_export": [vp, vp],
        "ub200_brick_ipc_import": [vp, vp],
        "ub200_brick_arena": [vp, C.POINTER(vp)],
        "ub200_brick_attach_local": [vp, C.POINTER(vp)],
        "ub200_brick_set_global_state_f32": [vp, vp, vp, i, vp],
        "ub200_halo_exchange_f32": [vp, vp],
        "ub200_halo_exchange_phase_f32": [vp, i, vp],
        "ub200_brick_lj_forces_f32": [vp, fp, i, vp],
        "ub200_brick_lj_nve_run_f32": [vp, fp, i, f, i, vp],
        "ub200_brick_lj_nve_phase_f32": [vp, i, fp, i, f, i, vp],
        "ub200_brick_dpd_nve_run_f32": [vp, f, f, f, f, C.c_uint32, f, i, vp],
        "ub200_brick_dpd_nve_phase_f32": [vp, i, f, f, f, f, C.c_uint32, f, i, vp],
        "ub200_brick_info": [vp, C.POINTER(BrickInfo)],
        "ub200_brick_counts": [vp, vp, ip, ip, ip],
        "ub200_brick_profile": [vp, C.POINTER(C.c_double)],
        "ub200_brick_download_owned_f32": [vp, vp, vp, vp, i, vp],
        "ub200_brick_upload_owned_f32": [vp, vp, vp, vp, i, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = C.c_int, args
    lib._brick_declared = True
    return lib


def half_cells(box, cutOff):
    """Half cells per dimension of the engine's grid (colCellsFor in include/uammd_b200/colgeom.h)."""
    return tuple(1 if l == 0.0 else max(1, int(2.0 * l / (cutOff * 1.00001))) for l in box.boxSize)


class BrickLJMD:
    """One rank of the brick-decomposed LJ MD. State lives in the handle on the device: owned particles first, ghosts
    behind them, in no particular order (global ids travel with the particles)."""

    def __init__(self, box, pot, dt, N, rank, world, rankGrid=None, capacity=0):
        self.lib = _declare()
        self.box, self.pot, self.dt, self.N, self.rank, self.world = box, pot, float(dt), int(N), int(rank), int(world)
        self.rankGrid = tuple(rankGrid) if rankGrid is not None else choose_rank_grid(world, half_cells(box, pot.getCutOff()))
        if self.rankGrid[0] * self.rankGrid[1] * self.rankGrid[2] != world:
            raise UB200Error(f"rank grid {self.rankGrid} does not match the number of ranks ({world})")
        self._tab = pot.table()
        self._tabp = self._tab.ctypes.data_as(C.POINTER(C.c_float))
        self._h = C.c_void_p()
        check(self.lib.ub200_brick_create(C.byref(self._h), self.rank, i3(self.rankGrid), f3(box.boxSize),
                                          i3([int(p) for p in box.periodic]), float(pot.getCutOff()), self.N, int(capacity)))
        self.device = None

    def __del__(self):
        try:
            if self._h:
                self.lib.ub200_brick_destroy(self._h)
        except Exception:
            pass

    # -- wiring -----------------------------------------------------------------------------------------------------
    def connect(self, group=None):
        """Exchange the CUDA IPC handles of the arenas over torch.distributed (one process per GPU)."""
        import torch.distributed as dist
        if self.world == 1:
            return
        blob = C.create_string_buffer(self.lib.ub200_comm_ipc_size())
        check(self.lib.ub200_brick_ipc_export(self._h, blob))
        blobs = [None] * self.world
        dist.all_gather_object(blobs, blob.raw, group=group)
        self._blobs = C.create_string_buffer(b"".join(blobs))
        check(self.lib.ub200_brick_ipc_import(self._h, self._blobs))

    @staticmethod
    def connectLocal(ranks):
        """Virtual ranks inside one process (tests on one GPU): hand every rank the arenas of all."""
        arenas = (C.c_void_p * len(ranks))()
        for k, r in enumerate(ranks):
            a = C.c_void_p()
            check(r.lib.ub200_brick_arena(r._h, C.byref(a)))
            arenas[k] = a
        for r in ranks:
            check(r.lib.ub200_brick_attach_local(r._h, arenas))

    # -- state ------------------------------------------------------------------------------------------------------
    def setGlobalState(self, pos, vel, stream=None):
        if pos.dtype != torch.float32 or pos.shape != (self.N, 4) or vel.shape != (self.N, 3) or not pos.is_cuda:
            raise UB200Error("BrickLJMD.setGlobalState: pos [N,4] / vel [N,3] float32 CUDA tensors")
        check(self.lib.ub200_brick_set_global_state_f32(self._h, _ptr(pos.contiguous()), _ptr(vel.contiguous()), self.N,
                                                        _stream_ptr(stream)))
        self.device = pos.device

    def info(self):
        v = BrickInfo()
        check(self.lib.ub200_brick_info(self._h, C.byref(v)))
        return v

    def counts(self):
        """(nOwned, nLocal, errorFlag); synchronises."""
        a, b, e = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.lib.ub200_brick_counts(self._h, _stream_ptr(), C.byref(a), C.byref(b), C.byref(e)))
        return a.value, b.value, e.value

    def owned(self):
        """Copies of the owned block: pos [n,4], vel [n,3], gid [n], force [n,4]."""
        n, _, err = self.counts()
        if err:
            raise UB200Error(f"brick decomposition: device error flag {err}")
        v = self.info()
        dev = self.device
        return (_device_copy(v.d_pos, (n, 4), torch.float32, dev), _device_copy(v.d_vel, (n, 3), torch.float32, dev),
                _device_copy(v.d_gid, (n,), torch.int32, dev), _device_copy(v.d_force, (n, 4), torch.float32, dev))

    def profile(self):
        """Mean ms per step of (push, unpack + wait, list build, traversal, kick); needs UB200_BRICK_PROFILE=1."""
        ph = (C.c_double * 5)()
        check(self.lib.ub200_brick_profile(self._h, ph))
        return dict(zip(("push", "unpack_wait", "build", "traversal", "kick"), ph))

    def downloadOwned(self, h_pos, h_vel, n, h_gid=None, stream=None):
        """Owned block -> pinned host tensors (asynchronous)."""
        check(self.lib.ub200_brick_download_owned_f32(self._h, _ptr(h_pos), _ptr(h_vel), _ptr(h_gid), int(n), _stream_ptr(stream)))

    def uploadOwned(self, h_pos, h_vel, n, h_gid=None, stream=None):
        """Pinned host tensors -> owned block, in place (asynchronous)."""
        check(self.lib.ub200_brick_upload_owned_f32(self._h, _ptr(h_pos), _ptr(h_vel), _ptr(h_gid), int(n), _stream_ptr(stream)))

    # -- stepping ---------------------------------------------------------------------------------------------------
    def run(self, nsteps, stream=None):
        check(self.lib.ub200_brick_lj_nve_run_f32(self._h, self._tabp, self.pot.ntypes, self.dt, int(nsteps), _stream_ptr(stream)))

    def phase(self, phase, doKick=True, stream=None):
        check(self.lib.ub200_brick_lj_nve_phase_f32(self._h, int(phase), self._tabp, self.pot.ntypes, self.dt, int(doKick),
                                                    _stream_ptr(stream)))

    @staticmethod
    def runLocal(ranks, nsteps, prepared=False):
        """nsteps of virtual ranks sharing one process and stream: phase 0 of every rank, then phase 1 of every rank."""
        if not prepared:
            for ph in (0, 1):
                for r in ranks:
                    r.phase(ph, doKick=False)
        for _ in range(nsteps):
            for ph in (0, 1):
                for r in ranks:
                    r.phase(ph, doKick=True)

    def gatherGlobalState(self, group=None):
        """Full pos [N,4] / vel [N,3] / force [N,4] in global-id order on every rank (diagnostics and tests)."""
        import torch.distributed as dist
        pos, vel, gid, force = self.owned()
        if self.world == 1:
            parts = [(pos, vel, gid, force)]
        else:
            parts = [None] * self.world
            dist.all_gather_object(parts, tuple(t.cpu() for t in (pos, vel, gid, force)), group=group)
        return assemble(parts, self.N)


class BrickDPDMD(BrickLJMD):
    """One rank of the brick-decomposed DPD fluid (BASELINE config 4). pot: uammd_b200.md.DPD (cutOff, dt, gamma,
    temperature, A, seed); the oracle is VerletNVE + PairForcesDPD on one GPU, reproduced bit for bit."""

    def __init__(self, box, pot, dt, N, rank, world, rankGrid=None, capacity=0):
        class _Cut:  # the brick geometry only needs the cut-off
            ntypes = 1

            def getCutOff(self_inner):
                return pot.getCutOff()

            def table(self_inner):
                return np.zeros(4, np.float32)
        super().__init__(box, _Cut(), dt, N, rank, world, rankGrid, capacity)
        self.dpd = pot

    def _args(self):
        p = self.dpd
        return (float(p.A), float(p.gamma), float(p.sigma), float(p.rcut), int(p.seed) & 0xFFFFFFFF, self.dt)

    def run(self, nsteps, stream=None):
        check(self.lib.ub200_brick_dpd_nve_run_f32(self._h, *self._args(), int(nsteps), _stream_ptr(stream)))

    def phase(self, phase, doKick=True, stream=None):
        a = self._args()
        check(self.lib.ub200_brick_dpd_nve_phase_f32(self._h, int(phase), *a, int(doKick), _stream_ptr(stream)))


def assemble(parts, N):
    """[(pos, vel, gid, force) per rank] -> arrays in global-id order; every id must appear exactly once."""
    gid = np.concatenate([p[2].cpu().numpy() for p in parts])
    if gid.shape[0] != N or not np.array_equal(np.sort(gid), np.arange(N)):
        raise UB200Error(f"brick decomposition lost or duplicated particles: {gid.shape[0]} owned rows for N = {N}")
    order = np.argsort(gid)
    return tuple(np.concatenate([p[k].cpu().numpy() for p in parts])[order] for k in (0, 1, 3))
