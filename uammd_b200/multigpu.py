"""Multi-GPU LJ molecular dynamics by particle decomposition (one process per GPU, torch.distributed).

The reference is single-GPU: this is new functionality whose oracle is the single-GPU result on the
same global input. Scheme ("replicated positions, block-owned particles"):
  * particle i is owned by rank i // (N / world); a rank keeps the velocities and integrates only its block;
  * every step the ranks all-gather their position blocks (NCCL over NVLink; 16 B per particle) so that each rank
    holds all positions, builds the full cell list (cheap, replicated) and computes forces ONLY for its block
    (ub200_lj_sum_owned_f32 skips cells without owned home particles);
  * each particle's force is computed by exactly one rank with the same kernel, list and summation order as on one
    GPU, so the trajectory is bit-identical to the single-GPU trajectory.
The host logic is written against a small engine interface (half / kick_kick_drift / forces_owned); the product engine
below drives the C ABI and raises when the CUDA library is missing. The world_size-2 gloo tests on CPU inject their own
checker engine from tests/ - nothing in this package computes on the CPU.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class BlockDecomposition:
    """Contiguous index blocks of equal size (all_gather_into_tensor needs equal contributions)."""

    def __init__(self, N, world, rank):
        if N % world != 0:
            raise ValueError(f"particle decomposition needs N ({N}) divisible by the number of ranks ({world})")
        self.N, self.world, self.rank = N, world, rank
        self.block = N // world
        self.lo, self.hi = rank * self.block, (rank + 1) * self.block


class _CudaEngine:
    def __init__(self, box, pot, dt):
        from . import _lib
        from .md import CellList, _ptr, _stream_ptr
        self.lib, self._ptr, self._stream = _lib.lib(), _ptr, _stream_ptr
        self.check = _lib.check
        self.box, self.pot, self.dt = box, pot, float(dt)
        self.nl = CellList()
        self.tab = pot.table()
        self.tabp = self.tab.ctypes.data_as(C.POINTER(C.c_float))

    def half(self, step, pos_blk, vel_blk, force_blk):
        self.check(self.lib.ub200_nve_half_step_f32(self._ptr(pos_blk), self._ptr(vel_blk), self._ptr(force_blk),
                                                    C.c_void_p(0), 1.0, C.c_void_p(0), pos_blk.shape[0], self.dt, 0,
                                                    step, self._stream()))

    def kick_kick_drift(self, pos_blk, vel_blk, force_blk):
        self.check(self.lib.ub200_nve_kick_kick_drift_f32(self._ptr(pos_blk), self._ptr(vel_blk), self._ptr(force_blk),
                                                          pos_blk.shape[0], self.dt, self._stream()))

    def forces_owned(self, pos, force, lo, hi):
        self.nl.update(pos, self.box, self.pot.getCutOff())
        self.check(self.lib.ub200_lj_sum_owned_f32(self.nl._h, self.tabp, self.pot.ntypes, self._ptr(force), lo, hi, 0,
                                                   self._stream()))


class DistributedLJMD:
    """VerletNVE + PairForces<LJ, CellList> over `world` ranks. pos/force are full-size [N,4] tensors replicated on
    every rank (only the owned block of force is meaningful), vel is the rank's own [N/world,3] block."""

    def __init__(self, box, pot, dt, N, engine=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dec = BlockDecomposition(N, self.world, self.rank)
        self.eng = engine if engine is not None else _CudaEngine(box, pot, dt)
        self.prepared = False

    def _gather(self, pos):
        if self.world > 1:
            dist.all_gather_into_tensor(pos, pos[self.dec.lo:self.dec.hi].clone() if pos.device.type == "cpu"
                                        else pos[self.dec.lo:self.dec.hi], group=self.group)

    def prepare(self, pos, force):
        self.eng.forces_owned(pos, force, self.dec.lo, self.dec.hi)
        self.prepared = True

    def run(self, pos, vel_blk, force, nsteps):
        lo, hi = self.dec.lo, self.dec.hi
        if not self.prepared:
            self.prepare(pos, force)
        pb, fb = pos[lo:hi], force[lo:hi]
        for s in range(nsteps):
            if s == 0:
                self.eng.half(1, pb, vel_blk, fb)
            self._gather(pos)
            self.eng.forces_owned(pos, force, lo, hi)
            if s == nsteps - 1:
                self.eng.half(2, pb, vel_blk, fb)
            else:
                self.eng.kick_kick_drift(pb, vel_blk, fb)


class _CudaDPDEngine:
    def __init__(self, box, pot, dt, N):
        from . import _lib
        from .md import CellList, _ptr, _stream_ptr
        self.lib, self.check, self._ptr, self._stream = _lib.lib(), _lib.check, _ptr, _stream_ptr
        self.box, self.pot, self.dt, self.N = box, pot, float(dt), N
        self.nl = CellList()

    def forces_owned(self, pos, vel, force, lo, hi):
        p = self.pot
        self.nl.update(pos, self.box, p.getCutOff())
        self.check(self.lib.ub200_dpd_sum_owned_f32(self.nl._h, self._ptr(vel), p.A, p.gamma, p.sigma, p.rcut, p.seed,
                                                    p.step & 0xFFFFFFFF, self.N, self._ptr(force), lo, hi, 0, self._stream()))

    def half(self, step, pb, vb, fb):
        self.check(self.lib.ub200_nve_half_step_f32(self._ptr(pb), self._ptr(vb), self._ptr(fb), C.c_void_p(0), 1.0,
                                                    C.c_void_p(0), pb.shape[0], self.dt, 0, step, self._stream()))


class DistributedDPDMD:
    """VerletNVE + PairForces<DPD> over `world` ranks by the same particle decomposition (BASELINE config 4 shape): the
    DPD force depends on the velocities, so positions AND velocities are all-gathered every step (28 B per particle);
    the pairwise noise is keyed on global particle indices, hence independent of the number of ranks. pos [N,4] and vel
    [N,3] are full-size tensors replicated on every rank; force [N,4] is meaningful on the owned block.
    (uammd_b200/domain.py holds the brick decomposition with a ghost-cell halo exchange instead of the all-gather.)"""

    def __init__(self, box, pot, dt, N, engine=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dec = BlockDecomposition(N, self.world, self.rank)
        self.box, self.pot, self.dt, self.N = box, pot, float(dt), N
        self.eng = engine if engine is not None else _CudaDPDEngine(box, pot, dt, N)
        self.steps = 0

    def _gather(self, t):
        if self.world > 1:
            lo, hi = self.dec.lo, self.dec.hi
            dist.all_gather_into_tensor(t, t[lo:hi].clone() if t.device.type == "cpu" else t[lo:hi], group=self.group)

    def _forces(self, pos, vel, force):
        self.pot.step += 1  # DPD_impl::getForceTransverser increments the step before every evaluation (DPD.cuh:161-170)
        self.eng.forces_owned(pos, vel, force, self.dec.lo, self.dec.hi)

    def _half(self, step, pos, vel, force):
        lo, hi = self.dec.lo, self.dec.hi
        self.eng.half(step, pos[lo:hi], vel[lo:hi], force[lo:hi])

    def forwardTime(self, pos, vel, force):
        """VerletNVE::forwardTime (VerletNVE.cu:174-188) with the DPD interactor."""
        self.steps += 1
        if self.steps == 1:
            self._forces(pos, vel, force)
        self._half(1, pos, vel, force)
        self._gather(pos); self._gather(vel)
        self._forces(pos, vel, force)
        self._half(2, pos, vel, force)

    def gatherState(self, pos, vel):
        """Make the replicated copies consistent (the other ranks' blocks lag by the last kick otherwise)."""
        self._gather(pos); self._gather(vel)


# ---------------------------------------------------------------------------------------------------------------
# Slab-decomposed FCM (ub200_fcm_dist_*): z slabs of the grid, NVLink peer stores fused into the FFT passes.
# ---------------------------------------------------------------------------------------------------------------
def slab_ranges(n, world):
    """[lo, hi) plane range of every rank (equal slabs; the C ABI requires divisibility)."""
    if n % world:
        raise ValueError(f"slab decomposition needs the grid dimension ({n}) divisible by the number of ranks ({world})")
    b = n // world
    return [(r * b, (r + 1) * b) for r in range(world)]


def exchange_blobs(blob, group=None):
    """All ranks' opaque set-up blobs in rank order (host-side plumbing only; works on gloo and nccl)."""
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(blob), group=group)
    return b"".join(out)


class DistributedFCM:
    """FCM_impl::computeHydrodynamicDisplacements over `world` GPUs. pos / force: replicated real4 [N,4] tensors
    (every rank passes the same data); returns the full real3 [N,3] result on every rank."""

    def __init__(self, box, cells, kernel, viscosity, maxParticles, seed=1, dtype=torch.float64, group=None):
        from . import _lib
        from .fcm import _declare as _fcm_declare, _prec
        from ._lib import check, d3, i3
        self.lib, self.check = _lib.lib(), check
        _fcm_declare()
        vp, i, d, u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint32
        l = self.lib
        l.ub200_fcm_dist_create.restype = i
        l.ub200_fcm_dist_create.argtypes = [C.POINTER(vp), i, C.c_double * 3, C.c_int * 3, vp, d, u32, i, i, i]
        l.ub200_fcm_dist_destroy.argtypes = [vp]
        l.ub200_fcm_dist_ipc_size.restype = i
        l.ub200_fcm_dist_ipc_export.argtypes = [vp, vp]
        l.ub200_fcm_dist_ipc_import.argtypes = [vp, vp]
        l.ub200_fcm_dist_mdot.restype = i
        l.ub200_fcm_dist_mdot.argtypes = [vp, vp, vp, i, d, d, vp, vp]
        l.ub200_fcm_dist_error_flag.argtypes = [vp, vp, C.POINTER(i)]
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        L = (box, box, box) if np.isscalar(box) else tuple(box)
        self.cells, self.dtype = tuple(int(c) for c in cells), dtype
        self.zrange = slab_ranges(self.cells[2], self.world)[self.rank]
        slab_ranges(self.cells[1], self.world)
        self._h = vp()
        ks = kernel.struct()
        check(l.ub200_fcm_dist_create(C.byref(self._h), _prec(dtype), d3(L), i3(self.cells), C.cast(C.byref(ks), vp),
                                      float(viscosity), seed & 0xFFFFFFFF, self.rank, self.world, int(maxParticles)))
        blob = C.create_string_buffer(l.ub200_fcm_dist_ipc_size())
        check(l.ub200_fcm_dist_ipc_export(self._h, blob))
        allb = exchange_blobs(blob.raw, group)
        self._blobs = C.create_string_buffer(allb, len(allb))
        check(l.ub200_fcm_dist_ipc_import(self._h, self._blobs))
        dist.barrier(group)

    def computeHydrodynamicDisplacements(self, pos, force, temperature=0.0, prefactor=0.0, out=None, stream=None):
        from .md import _ptr, _stream_ptr
        N = pos.shape[0]
        if out is None:
            out = torch.empty(N, 3, dtype=self.dtype, device=pos.device)
        self.check(self.lib.ub200_fcm_dist_mdot(self._h, _ptr(pos), _ptr(force) if force is not None else None, N,
                                                float(temperature), float(prefactor), _ptr(out), _stream_ptr(stream)))
        return out

    def errorFlag(self):
        from .md import _stream_ptr
        f = C.c_int(0)
        self.check(self.lib.ub200_fcm_dist_error_flag(self._h, _stream_ptr(None), C.byref(f)))
        return f.value

    def close(self):
        if self._h:
            self.lib.ub200_fcm_dist_destroy(self._h)
            self._h = None


class DistributedPSEFarField:
    """pse_ns::FarField::computeHydrodynamicDisplacements (PSE/FarField.cuh:535-553) over `world` GPUs: the slab machinery of
    DistributedFCM with the PSE Gaussian window and the Hasimoto-split RPY Green's function (BASELINE config 3:
    "slab-decomposed FFT over 8 GPUs"). Grid, support and eta are resolved by the single-GPU ub200_pse_create, so the two
    paths cannot drift apart; the grids are bit-identical to the single-GPU ones, the interpolation differs in fp32 summation order."""

    def __init__(self, par, maxParticles, seedFar, dtype=torch.float32, group=None):
        from . import pse as P
        from .bd import System
        # a throw-away single-GPU handle only to resolve the derived parameters like the reference constructors do
        probe = P.PSE(torch.zeros(1, 4, dtype=dtype, device="cuda"), par, sys=System(1))
        inf = probe.info()
        self.cells, self.support, self.eta = tuple(inf.cells), inf.support, inf.eta

        class _K:
            def struct(_self):
                from .fcm import IBMKernelStruct
                k = inf.kernel
                return IBMKernelStruct(k.kind, k.support, k.h, k.prefactor, k.tau, k.rmax)
        self.fcm = DistributedFCM(par.box, self.cells, _K(), par.viscosity, maxParticles, seed=seedFar, dtype=dtype, group=group)
        l = self.fcm.lib
        l.ub200_fcm_dist_set_pse_operator.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        l.ub200_fcm_dist_set_noise_seed2.argtypes = [C.c_void_p, C.c_uint32]
        self.fcm.check(l.ub200_fcm_dist_set_pse_operator(self.fcm._h, par.hydrodynamicRadius, par.psi, self.eta, par.shearStrain))
        del probe

    def computeHydrodynamicDisplacements(self, pos, force, MF, temperature=0.0, prefactor=0.0, seed2=0, stream=None):
        """MF += Mw F + prefactor sqrt(2T) Mw^1/2 dW (the reference's gather accumulates)."""
        self.fcm.check(self.fcm.lib.ub200_fcm_dist_set_noise_seed2(self.fcm._h, seed2 & 0xFFFFFFFF))
        out = self.fcm.computeHydrodynamicDisplacements(pos, force, temperature=temperature, prefactor=prefactor, stream=stream)
        MF += out
        return MF

    def close(self):
        self.fcm.close()


class DistributedPSENearField:
    """pse_ns::NearField (PSE/NearField.cuh:236-282: Mdot and the Lanczos noise) over `world` GPUs (ub200_pse_dist_*):
    replicated positions, every rank evaluates the rows of its share of the cell-sorted order and stores the next Krylov
    vector into the records of all ranks over NVLink; the scalars of the iteration are summed by a one-warp kernel that is
    also the barrier. Results are complete on every rank and equal the single-GPU ones up to the summation order of the
    dot products (the noise of a particle is keyed by its index).

    group = a torch.distributed group (one rank per process, CUDA IPC), or None with rank / world given and
    `attachLocal` called on the list of all virtual ranks (one process, one stream and one host thread per rank: tests)."""

    def __init__(self, pos, par, sys=None, group=None, rank=None, world=None):
        from . import pse as P
        from ._lib import check
        self.pse = P.PSE(pos, par, sys=sys)
        self.lib, self.check, self.pos, self.N = self.pse.lib, check, pos, pos.shape[0]
        vp, i, d, u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint32
        l = self.lib
        for name, args in {"ub200_pse_dist_create": [vp, i, i, i], "ub200_pse_dist_ipc_export": [vp, vp],
                           "ub200_pse_dist_ipc_import": [vp, vp], "ub200_pse_dist_arena": [vp, C.POINTER(vp)],
                           "ub200_pse_dist_attach_local": [vp, C.POINTER(vp)],
                           "ub200_pse_dist_near_prepare": [vp, vp, i, vp],
                           "ub200_pse_dist_near_mdot": [vp, vp, i, i, vp, vp],
                           "ub200_pse_dist_near_noise_add": [vp, i, d, d, u32, vp, C.POINTER(i), vp],
                           "ub200_pse_dist_error_flag": [vp, vp, C.POINTER(i)]}.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = i, args
        l.ub200_comm_ipc_size.restype = i
        local = rank is not None
        self.rank = rank if local else dist.get_rank(group)
        self.world = world if local else dist.get_world_size(group)
        check(l.ub200_pse_dist_create(self.pse._h, self.rank, self.world, self.N))
        if not local:
            blob = C.create_string_buffer(l.ub200_comm_ipc_size())
            check(l.ub200_pse_dist_ipc_export(self.pse._h, blob))
            allb = exchange_blobs(blob.raw, group)
            self._blobs = C.create_string_buffer(allb, len(allb))
            check(l.ub200_pse_dist_ipc_import(self.pse._h, self._blobs))
            dist.barrier(group)

    @staticmethod
    def attachLocal(ranks):
        """virtual ranks of one process: hand every rank the arenas of all of them"""
        arenas = (C.c_void_p * len(ranks))()
        for r in ranks:
            a = C.c_void_p()
            r.check(r.lib.ub200_pse_dist_arena(r.pse._h, C.byref(a)))
            arenas[r.rank] = a.value
        for r in ranks:
            r.check(r.lib.ub200_pse_dist_attach_local(r.pse._h, arenas))

    def prepare(self, stream=None):
        """NearField::updateNeighbourList for the current (replicated) positions; no communication"""
        from .md import _ptr, _stream_ptr
        self.check(self.lib.ub200_pse_dist_near_prepare(self.pse._h, _ptr(self.pos), self.N, _stream_ptr(stream)))

    def Mdot(self, force, MF, stream=None):
        """MF += M_near F (force: replicated real4 or real3 tensor)"""
        from .md import _ptr, _stream_ptr
        if force is None:
            return
        self.check(self.lib.ub200_pse_dist_near_mdot(self.pse._h, _ptr(force), force.shape[1], self.N, _ptr(MF), _stream_ptr(stream)))

    def noiseAdd(self, out, temperature, prefactor, seed2, stream=None):
        """out += prefactor sqrt(2 T) M_near^1/2 dW; returns the Lanczos iterations"""
        from .md import _ptr, _stream_ptr
        it = C.c_int(0)
        self.check(self.lib.ub200_pse_dist_near_noise_add(self.pse._h, self.N, float(temperature), float(prefactor),
                                                          seed2 & 0xFFFFFFFF, _ptr(out), C.byref(it), _stream_ptr(stream)))
        return it.value

    def errorFlag(self, stream=None):
        from .md import _stream_ptr
        f = C.c_int(0)
        self.check(self.lib.ub200_pse_dist_error_flag(self.pse._h, _stream_ptr(stream), C.byref(f)))
        return f.value


class DistributedPSE:
    """BDHI::PSE::computeHydrodynamicDisplacements (BDHI_PSE.cuh:141-158) over `world` GPUs (BASELINE config 3 end to end):
    near field and Lanczos noise by rows (DistributedPSENearField), far field by FFT slabs (DistributedPSEFarField).
    Same seed draws, in the same order, as the single-GPU uammd_b200.pse.PSE."""

    def __init__(self, pos, par, sys=None, group=None):
        from .bd import System
        self.sys = sys if sys is not None else System()
        self.near = DistributedPSENearField(pos, par, sys=self.sys, group=group)
        self.far = DistributedPSEFarField(par, pos.shape[0], self.near.pse.seedFar, dtype=pos.dtype, group=group)
        self.pos = pos

    def computeHydrodynamicDisplacements(self, force, MF, temperature, noise_prefactor, stream=None):
        MF.zero_()
        rng = self.sys.rng()
        if force is not None or temperature > 0:
            self.near.prepare(stream)
        self.near.Mdot(force, MF, stream)
        iterations = 0
        if temperature > 0:
            iterations = self.near.noiseAdd(MF, temperature, noise_prefactor, rng.next32(), stream)
        seed2 = rng.next32() if temperature > 0 else 0
        self.far.computeHydrodynamicDisplacements(self.pos, force, MF, temperature=temperature, prefactor=noise_prefactor,
                                                  seed2=seed2, stream=stream)
        return iterations

    def errorFlag(self):
        return self.near.errorFlag() | self.far.fcm.errorFlag()

    def close(self):
        self.far.close()
