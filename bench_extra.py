"""Secondary legs of bench.py (same timing protocol as the headline: CUDA events on the launching stream, L2 evicted
before every step, >= 3 warm-up steps):
  verlet : VerletNVE + PairForces<LJ, VerletList> at the headline workload (what generic_md instantiates)
  pse    : BDHI::EulerMaruyama<PSE> at BASELINE config 3's single-GPU shape (N = 1e6, L = 256 -> 256^3, fp32, T = 1)
  bd     : BD::EulerMaruyama ideal particles, N = 1e5 fp64 (config 0)
each next to the unmodified reference's own CUDA path (oracle/_ref binaries) in the reference arm."""
import json
import math
import os
import sys
import subprocess
import tempfile

import numpy as np
import torch

PSE_N, PSE_L, PSE_TOL, PSE_PSI, PSE_T, PSE_DT = 1_000_000, 256.0, 1e-3, 0.593, 1.0, 0.01


def _timed(dev, step, steps, warmup):
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        scrub.fill_(5)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def verlet(dev, N, Lb, pos, vel, rc, dt, steps=100, warmup=20, equil=300):
    from uammd_b200.md import Box, LJ, LJMD, VerletList
    from uammd_b200 import lib
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=rc)
    p, v = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev)
    f = torch.zeros(N, 4, device=dev)
    nl = VerletList()
    md = LJMD(Box(Lb), pot, dt)
    md.runVerlet(nl, p, v, f, equil)
    r0 = nl.rebuilds()
    l0 = lib().ub200_launch_count()
    ms = _timed(dev, lambda: md.runVerlet(nl, p, v, f, 1, forcesAreCurrent=True), steps, warmup)
    return {"metric": "MD steps/s @1e6 LJ particles (VerletList)", "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms,
            "rebuilds_per_step": (nl.rebuilds() - r0) / float(steps + warmup), "gpu_launches": int(lib().ub200_launch_count() - l0),
            "what": "VerletNVE + PairForces<LJ, VerletList> (skin 1.08), drift check read back every step like the reference"}


def verlet_reference(root, N, Lb, pos, vel, rc, dt, steps=100, warmup=20, equil=300):
    exe = os.path.join(root, "oracle", "_ref", "ref_lj_verlet")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_lj_verlet was not built"}
    with tempfile.TemporaryDirectory() as td:
        pos.tofile(os.path.join(td, "p.bin")); vel.tofile(os.path.join(td, "v.bin"))
        out = subprocess.run([exe, "md", str(N), str(Lb), str(Lb), str(Lb), str(rc), "1", "1", str(dt), str(warmup + equil), str(steps),
                              "1", os.path.join(td, "p.bin"), os.path.join(td, "v.bin"), "-"], check=True, capture_output=True,
                             text=True, timeout=1200).stdout
    r = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
    return {"metric": "MD steps/s @1e6 LJ particles (VerletList)", "value": 1000.0 / r["ms_per_step_mean"], "unit": "steps/s",
            "ms_per_step": r["ms_per_step_mean"], "rebuilds_per_step": r["rebuilds"] / float(steps),
            "what": "unmodified reference PairForces<LJ, VerletList> + VerletNVE on the same B200"}


def dpd(dev, steps=30, warmup=5, equil=300, world=1, rank=0):
    """BASELINE config 4: DPD fluid, N = 4e6, rho = 3, rc = 1, A = 25, gamma = 4.5, T = 1, dt = 0.01; VerletNVE +
    PairForces<DPD, CellList> on bricks with the device-side halo exchange (uammd_b200.brickmd.BrickDPDMD; one rank is a
    1 x 1 x 1 brick, world > 1 runs under torchrun). The reference's Potential::DPD is a silent no-op through PairForces at
    this commit, so there is no reference arm for this leg. Returns None except on rank 0."""
    from uammd_b200 import lib, synthetic as syn
    from uammd_b200.brickmd import BrickDPDMD
    from uammd_b200.md import Box, DPD
    N = 4_000_000
    L = (N / 3.0) ** (1.0 / 3.0)
    md = BrickDPDMD(Box(L), DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=99), 0.01, N, rank, world)
    md.connect()
    md.setGlobalState(torch.from_numpy(syn.uniform_cloud(N, L, seed=21)).to(dev),
                      torch.from_numpy(syn.maxwell_velocities(N, 1.0, seed=22)).to(dev))
    md.run(equil)
    l0 = lib().ub200_launch_count()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize(); dist.barrier()
    ms = _timed(dev, lambda: md.run(1), steps, warmup)
    no, nl, err = md.counts()
    _, v, _, _ = md.owned()
    stats = torch.tensor([ms, float((v.double() ** 2).sum()), float(no), float(nl), float(err)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = stats[:1].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        stats[0] = mx[0]
    if rank != 0:
        return None
    ms = float(stats[0])
    return {"metric": "DPD MD steps/s @4e6 particles", "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "n_gpus": world,
            "gpu_launches": int(lib().ub200_launch_count() - l0), "kT_from_velocities": float(stats[1]) / (3.0 * N),
            "rank_grid": list(md.rankGrid), "owned_total": int(stats[2]), "local_total_with_ghosts": int(stats[3]),
            "error_flags": int(stats[4]),
            "what": "VerletNVE + PairForces<DPD, CellList>, N = 4e6, rho = 3, rc = 1, A = 25, gamma = 4.5: bricks with ghost half "
                    "cells, one peer-to-peer halo exchange per step (ghosts carry velocities, noise keyed on global ids)"}


LANGEVIN_N, LANGEVIN_L = 1 << 20, 128.0


def langevin(dev, steps=100, warmup=20, equil=300):
    """The reference's own published benchmark (examples/misc/benchmark.cu:8,69-108,172-181: "~90 FPS on a GTX 980"):
    N = 1 048 576 LJ particles in a 128^3 box (rho = 0.5) from an FCC lattice, VerletNVT::GronbechJensen (T = 1, friction = 1,
    dt = 0.01) + PairForces<LJ, VerletList> with a 1.2 cut-off multiplier."""
    from uammd_b200 import lib, synthetic as syn
    from uammd_b200.bd import System
    from uammd_b200.md import Box, LJ, PairForces, VerletList
    from uammd_b200.nvt import GronbechJensen, Parameters
    N, L = LANGEVIN_N, LANGEVIN_L
    p = torch.from_numpy(syn.fcc_lattice(N, L)).to(dev)
    v = torch.zeros(N, 3, device=dev)
    nvt = GronbechJensen(p, v, Parameters(temperature=1.0, dt=0.01, friction=1.0, initVelocities=True), sys=System(1234))
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    nl = VerletList(); nl.setCutOffMultiplier(1.2)
    nvt.addInteractor(PairForces(pot, Box(L), nl=nl))
    for _ in range(equil):
        nvt.forwardTime()
    r0, l0 = nl.rebuilds(), lib().ub200_launch_count()
    ms = _timed(dev, nvt.forwardTime, steps, warmup)
    ke = float((v * v).sum().item()) / (3.0 * N)
    return {"metric": "Langevin MD steps/s @2^20 LJ particles (benchmark.cu)", "value": 1000.0 / ms, "unit": "steps/s",
            "ms_per_step": ms, "rebuilds_per_step": (nl.rebuilds() - r0) / float(steps + warmup),
            "gpu_launches": int(lib().ub200_launch_count() - l0), "kT_from_velocities": ke, "published_gtx980_steps_per_s": 90,
            "what": "VerletNVT::GronbechJensen + PairForces<LJ, VerletList> (skin 1.2), N = 1048576, box 128^3, dt 0.01, T 1"}


def langevin_reference(root, steps=100, warmup=20, equil=300):
    from uammd_b200 import synthetic as syn
    exe = os.path.join(root, "oracle", "_ref", "ref_nvt")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_nvt was not built"}
    N, L = LANGEVIN_N, LANGEVIN_L
    with tempfile.TemporaryDirectory() as td:
        syn.fcc_lattice(N, L).tofile(os.path.join(td, "p.bin")); np.zeros((N, 3), np.float32).tofile(os.path.join(td, "v.bin"))
        out = subprocess.run([exe, str(N), str(L), str(steps), "1.0", "1.0", "0.01", "1234", "1", "1", os.path.join(td, "o"),
                              os.path.join(td, "p.bin"), os.path.join(td, "v.bin"), "1.2", str(warmup + equil)], check=True,
                             capture_output=True, text=True, timeout=1200).stdout
    r = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
    return {"metric": "Langevin MD steps/s @2^20 LJ particles (benchmark.cu)", "value": 1000.0 / r["ms_per_step"], "unit": "steps/s",
            "ms_per_step": r["ms_per_step"], "published_gtx980_steps_per_s": 90,
            "what": "unmodified reference VerletNVT::GronbechJensen + PairForces<LJ, VerletList> (skin 1.2) on the same B200; "
                    "back-to-back steps (no L2 flush between steps)"}


def _pse_inputs():
    from uammd_b200 import synthetic as syn
    pos = np.zeros((PSE_N, 4), np.float32); pos[:, :3] = syn.uniform_cloud(PSE_N, PSE_L, seed=31)[:, :3]
    force = np.zeros((PSE_N, 4), np.float32); force[:, :3] = syn.gaussian_forces(PSE_N, seed=32, dtype=np.float32)
    return pos, force


def pse(dev, steps=20, warmup=3):
    from uammd_b200 import bd, lib
    from uammd_b200 import pse as P
    pos, force = _pse_inputs()
    p, f = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    m = P.PSE(p, P.Parameters(PSE_L, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=PSE_TOL, psi=PSE_PSI, temperature=PSE_T,
                              dt=PSE_DT), sys=bd.System(1234), force=f)
    integ = P.EulerMaruyama(m, PSE_DT, PSE_T)
    l0 = lib().ub200_launch_count()
    ms = _timed(dev, integ.forwardTime, steps, warmup)
    inf = m.info()
    # far + near timed alone (deterministic part)
    MF = torch.zeros(PSE_N, 3, device=dev)
    m.temperature = 0.0
    far = _timed(dev, lambda: m.computeMFFarField(MF), 10, 3)
    near = _timed(dev, lambda: m.computeMFNearField(MF), 10, 3)
    return {"metric": "PSE steps/s @1e6 particles, 256^3", "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "dtype": "f32",
            "config": {"workload": f"BDHI::EulerMaruyama<PSE>, N={PSE_N}, L={PSE_L}, a=1, tol={PSE_TOL}, psi={PSE_PSI} -> grid "
                                   f"{tuple(inf.cells)}, support {inf.support}, rcut {inf.rcut:.3f}, T={PSE_T}, dt={PSE_DT}"},
            "lanczos_iterations": inf.lastLanczosIterations, "far_field_T0_ms": far, "near_field_T0_ms": near,
            "gpu_launches": int(lib().ub200_launch_count() - l0)}


def pse_far_distributed(dev, steps=20, warmup=3):
    """BASELINE config 3's far field ("slab-decomposed FFT over 8 GPUs"): pse_ns::FarField over all ranks
    (uammd_b200.multigpu.DistributedPSEFarField: z slabs, FFT transposes as NVLink peer stores), N = 1e6, 256^3 fp32, force and
    noise. Device time per call, max over ranks; rank 0 also times the single-GPU far field of the same inputs. Returns None
    except on rank 0."""
    import math
    import torch.distributed as dist
    from uammd_b200 import bd
    from uammd_b200 import pse as P
    from uammd_b200.multigpu import DistributedPSEFarField
    world, rank = dist.get_world_size(), dist.get_rank()
    pos, force = _pse_inputs()
    p, f = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    par = P.Parameters(PSE_L, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=PSE_TOL, psi=PSE_PSI, temperature=PSE_T, dt=PSE_DT)
    far = DistributedPSEFarField(par, PSE_N, seedFar=777)
    far_planes = far.cells[2]
    MF = torch.zeros(PSE_N, 3, device=dev)
    calls = [0]

    def step():
        calls[0] += 1
        far.computeHydrodynamicDisplacements(p, f, MF, temperature=PSE_T, prefactor=1.0 / math.sqrt(PSE_DT), seed2=calls[0])

    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(); dist.barrier()
    step(); step()  # untimed: re-aligns the ranks on the device after the host-side barrier (see bench.py)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        scrub.fill_(3)
        a.record(); step(); b.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in evs]))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = far.fcm.errorFlag()
    far.close()
    single = None
    if rank == 0:
        m = P.PSE(p, par, sys=bd.System(1234), force=f)
        single = _timed(dev, lambda: m.computeMFFarField(MF), 10, 3)
    dist.barrier()
    if rank != 0:
        return None
    return {"metric": "PSE far-field calls/s @1e6 particles, 256^3", "value": 1000.0 / float(t.item()), "unit": "calls/s",
            "ms_per_step": float(t.item()), "n_gpus": world, "scaling": "strong", "single_gpu_ms": single, "barrier_error_flag": err,
            "what": "pse_ns::FarField::computeHydrodynamicDisplacements (force + noise) over z slabs, %d planes per rank" % (far_planes // world)}


def run_bounded(limit_s, fn, on_timeout):
    """fn() under a watchdog: returns fn's result, or {"error": ...} if it raised. If fn has not returned after
    limit_s seconds, on_timeout() runs on the timer thread (it prints what there is to print) and the PROCESS exits with code 0:
    a leg that hangs on one rank (a lost peer, a collective nobody else reaches) must not take the lines measured before it
    down with it. Exactly one of the two ways out is taken."""
    import threading
    lock, finished = threading.Lock(), [False]

    def abandon():
        with lock:
            if finished[0]:
                return
            finished[0] = True
            try:
                on_timeout()
            finally:
                sys.stdout.flush()
                os._exit(0)
    timer = threading.Timer(limit_s, abandon)
    timer.daemon = True
    timer.start()
    try:
        result = fn()
    except Exception as e:  # a secondary leg must not take the headline down
        result = {"error": repr(e)[:300]}
    with lock:
        finished[0] = True
    timer.cancel()
    return result


class RankAgreement:
    """Keeps the ranks of a bench leg on ONE sequence of torch collectives whatever happens inside a rank's own steps (which use
    only the library's time-bounded peer barriers): `guarded(fn)` runs fn unless this rank has already failed and records an
    exception instead of raising it; `failed()` is a collective that tells every rank whether ANY rank failed (or reports a
    non-zero `error_flag()`, e.g. a timed-out peer barrier). A failure thus becomes an "error" entry of the JSON line, never a
    rank that left the others waiting in a collective. `failure[0]` holds this rank's own message (None if it was another's)."""

    def __init__(self, dev, error_flag=None):
        self.dev, self.error_flag, self.failure = dev, error_flag, [None]

    def guarded(self, fn):
        if self.failure[0] is None:
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                self.failure[0] = repr(e)[:200]

    def failed(self):
        import torch.distributed as dist
        if self.failure[0] is None and self.error_flag is not None:
            try:
                if self.error_flag() != 0:
                    self.failure[0] = "a peer barrier timed out"
            except Exception as e:  # noqa: BLE001
                self.failure[0] = repr(e)[:200]
        flag = torch.tensor([0.0 if self.failure[0] is None else 1.0], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return flag.item() != 0


def pse_near_distributed(dev, steps=10, warmup=2):
    """BASELINE config 3's near field over all ranks (uammd_b200.multigpu.DistributedPSENearField: rows of the cell-sorted order
    per rank, Krylov vectors exchanged as NVLink peer stores, scalars summed by the barrier kernel): list build (replicated) +
    M_near F + Lanczos noise at N = 1e6 fp32. Device time per call, max over ranks; rank 0 also times the same three pieces on
    one GPU. Returns None except on rank 0."""
    import math
    import torch.distributed as dist
    from uammd_b200 import bd
    from uammd_b200 import pse as P
    from uammd_b200.multigpu import DistributedPSENearField
    world, rank = dist.get_world_size(), dist.get_rank()
    pos, force = _pse_inputs()
    p, f = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    par = P.Parameters(PSE_L, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=PSE_TOL, psi=PSE_PSI, temperature=PSE_T, dt=PSE_DT)
    near = DistributedPSENearField(p, par, sys=bd.System(1234))
    MF = torch.zeros(PSE_N, 3, device=dev)
    calls, its = [0], [0]
    pref = 1.0 / math.sqrt(PSE_DT)

    def step():
        calls[0] += 1
        near.prepare()
        near.Mdot(f, MF)
        its[0] = near.noiseAdd(MF, PSE_T, pref, calls[0])

    agree = RankAgreement(dev, near.errorFlag)
    guarded, agreed_failure, failure = agree.guarded, agree.failed, agree.failure

    guarded(step)
    if agreed_failure():  # no number rather than a wrong one
        return {"error": failure[0] or "another rank failed in the first call"} if rank == 0 else None
    for _ in range(warmup):
        guarded(step)
    guarded(torch.cuda.synchronize); dist.barrier()
    guarded(step)  # untimed: re-aligns the ranks on the device after the host-side barrier (see bench.py)
    evs = []

    def timed():
        scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            scrub.fill_(3)
            a.record(); step(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
    guarded(timed)
    dist.barrier()
    if agreed_failure():
        return {"error": failure[0] or "another rank failed in the timed calls"} if rank == 0 else None
    t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in evs]))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = near.errorFlag()
    single, single_it = None, None
    if rank == 0:
        m = P.PSE(p, par, sys=bd.System(1234), force=f)

        def one():
            m.computeMFNearField(MF, listForNoise=True)
            m._nearNoise(MF, PSE_T, pref, None, add=False, reuse=True)
        try:
            single = _timed(dev, one, steps, warmup)
            single_it = m.info().lastLanczosIterations
        except Exception as e:  # noqa: BLE001 - the collective below must still be reached
            single = repr(e)[:200]
    dist.barrier()
    if rank != 0:
        return None
    return {"metric": "PSE near-field calls/s @1e6 particles", "value": 1000.0 / float(t.item()), "unit": "calls/s",
            "ms_per_step": float(t.item()), "n_gpus": world, "scaling": "strong", "single_gpu_ms": single,
            "lanczos_iterations": its[0], "single_gpu_lanczos_iterations": single_it, "barrier_error_flag": err,
            "what": "NearField::Mdot + computeStochasticDisplacements (Lanczos) by rows of the sorted order, %d rows per rank; "
                    "the neighbour list is built on every rank" % (PSE_N // world)}


def pse_reference(root, steps=20, warmup=3, reps=3):
    """the reference's step time moved between runs of the harness in round 1 (12 - 40 steps/s); `reps` runs of the same
    command, the median is reported and all of them are listed"""
    exe = os.path.join(root, "oracle", "_ref", "ref_pse_f32")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_pse_f32 was not built"}
    pos, force = _pse_inputs()
    runs = []
    with tempfile.TemporaryDirectory() as td:
        pos.tofile(os.path.join(td, "p.bin")); force.tofile(os.path.join(td, "f.bin"))
        for _ in range(reps):
            out = subprocess.run([exe, "time", str(PSE_N), repr(PSE_L), "1.0", "1.0", repr(PSE_TOL), repr(PSE_PSI), "0.0", repr(PSE_T),
                                  repr(PSE_DT), "1234", str(warmup), str(steps), "1", os.path.join(td, "p.bin"), os.path.join(td, "f.bin")],
                                 check=True, capture_output=True, text=True, timeout=1800).stdout
            runs.append([json.loads(l) for l in out.splitlines() if l.startswith("{")][-1])
    ms = sorted(r["ms_per_step"] for r in runs)
    med = ms[len(ms) // 2]
    return {"metric": "PSE steps/s @1e6 particles, 256^3", "value": 1000.0 / med, "unit": "steps/s", "ms_per_step": med,
            "ms_per_step_runs": [r["ms_per_step"] for r in runs], "dtype": "f32",
            "what": "unmodified reference BDHI::PSE (cuFFT + cuBLAS Lanczos) on the same B200, same protocol; median of %d runs" % reps}


POISSON_N, POISSON_RHO = 200_000, 0.1


def poisson(dev, steps=10, warmup=2):
    """Poisson::sum (forces + energies) on 2e5 unit charges at number density 0.1, gw = 0.25, split = 1,
    tolerance 1e-4, double precision - the workload examples/dropin_poisson.cu times for both implementations."""
    from uammd_b200.poisson import Parameters, Poisson
    N = POISSON_N
    L = (N / POISSON_RHO) ** (1.0 / 3.0)
    rng = np.random.default_rng(11)
    pos = np.zeros((N, 4)); pos[:, :3] = (rng.random((N, 3)) - 0.5) * L
    q = np.where(np.arange(N) % 2 == 1, 1.0, -1.0)
    p, c = torch.from_numpy(pos).to(dev), torch.from_numpy(q).to(dev)
    solver = Poisson(p, c, Parameters(L, epsilon=1.0, tolerance=1e-4, gw=0.25, split=1.0))
    force, energy = torch.zeros(N, 4, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
    ms = _timed(dev, lambda: solver.sum(force=force, energy=energy), steps, warmup)
    inf = solver.info()
    return {"metric": "Poisson sums/s @2e5 charges fp64", "value": 1000.0 / ms, "unit": "sums/s", "ms_per_step": ms,
            "cells": list(inf.cells), "support": inf.support, "near_field_cut_off": inf.nearFieldCutOff,
            "what": "Poisson::sum(force, energy), Ewald split 1.0, tolerance 1e-4, L2 flushed between calls"}


def poisson_reference(root):
    exe = os.path.join(root, "oracle", "_ref", "dropin_poisson")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/dropin_poisson was not built"}
    out = subprocess.run([exe, "2000", str(POISSON_N)], check=True, capture_output=True, text=True, timeout=600).stdout
    r = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
    return {"metric": "Poisson sums/s @2e5 charges fp64", "value": 1000.0 / r["sum_ms_reference"], "unit": "sums/s",
            "ms_per_step": r["sum_ms_reference"], "what": "unmodified reference Poisson::sum on the same B200, calls back to back; "
            "the same binary also ran b200::Poisson: %.3f ms" % r["sum_ms_ours"]}


def bd_ideal(dev, steps=200, warmup=10):
    from uammd_b200 import bd
    N = 100_000
    rng = np.random.default_rng(1)
    pos = np.zeros((N, 4)); pos[:, :3] = rng.random((N, 3)) - 0.5
    p = torch.from_numpy(pos).to(dev)
    integ = bd.EulerMaruyama(p, bd.Parameters(temperature=1.0, viscosity=1.0, hydrodynamicRadius=1.0, dt=0.1), sys=bd.System(1234))
    for _ in range(warmup):
        integ.forwardTime()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        integ.forwardTime()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"metric": "BD steps/s @1e5 ideal particles fp64", "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms,
            "what": "BD::EulerMaruyama README example, steps enqueued back to back (one 64 B/particle kernel per step)"}


def bd_reference(root, steps=200):
    exe = os.path.join(root, "oracle", "_ref", "ref_bd")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_bd was not built"}
    with tempfile.TemporaryDirectory() as td:
        out = subprocess.run([exe, "100000", str(steps), "1.0", "1.0", "1.0", "0.1", "1234", os.path.join(td, "bd")], check=True,
                             capture_output=True, text=True, timeout=600).stdout
    r = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
    return {"metric": "BD steps/s @1e5 ideal particles fp64", "value": 1000.0 / r["ms_per_step"], "unit": "steps/s",
            "ms_per_step": r["ms_per_step"], "what": "unmodified reference BD::EulerMaruyama on the same B200"}
