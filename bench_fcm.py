"""FCM leg of bench.py: BASELINE.json configs[2] - BDHI::FCM triply periodic, N = 5e5, 128^3 grid, Peskin 3pt,
fp64. A step is the body of BDHI::EulerMaruyama<FCM>::forwardTime with fixed external forces at T = 1:
spread + FFT + Stokes projector + Fourier noise + inverse FFT + gather + position update.
Metric: Mdof.steps/s with dof = 3*128^3 grid velocity unknowns."""
import json
import math
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

N, NGRID, L, ETA, TEMP, DT = 500_000, 128, 128.0, 1.0, 1.0, 0.01
DOF = 3 * NGRID ** 3
# compulsory bytes per step: 7 grid passes of 51.12 MB + 60 MB particle I/O
GR = 2 * (NGRID // 2 + 1) * NGRID * NGRID * 24
ALG_BYTES = 7 * GR + N * (32 * 2 + 32 + 24)


def inputs():
    from uammd_b200 import synthetic as syn
    pos = np.zeros((N, 4)); pos[:, :3] = syn.uniform_cloud(N, L, seed=11)[:, :3].astype(np.float64)
    force = np.zeros((N, 4)); force[:, :3] = syn.gaussian_forces(N, seed=12)
    return pos, force


def run(dev, hbm_peak, steps=100, warmup=10):
    from uammd_b200.fcm import EulerMaruyama, FCM, Peskin3
    from uammd_b200 import lib
    pos, force = inputs()
    dpos, dforce = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    method = FCM(L, (NGRID,) * 3, Peskin3(L / NGRID), ETA, TEMP, DT, seed=1234)
    integ = EulerMaruyama(method, dpos, DT)
    integ.force = dforce
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        integ.forwardTime()
    torch.cuda.synchronize()
    l0 = lib().ub200_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        scrub.fill_(3)
        a.record()
        integ.forwardTime()
        b.record()
    torch.cuda.synchronize()
    launches = lib().ub200_launch_count() - l0
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        integ.forwardTime()
    e1.record()
    torch.cuda.synchronize()
    ms_b2b = e0.elapsed_time(e1) / steps
    # e2e: host (pinned) positions + forces in, displacements out, every step
    hp, hf = torch.from_numpy(pos).pin_memory(), torch.from_numpy(force).pin_memory()
    hout = torch.zeros(N, 3, dtype=torch.float64).pin_memory()
    dout = torch.zeros(N, 3, dtype=torch.float64, device=dev)
    import time
    for it in range(13):
        if it == 3:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        dpos.copy_(hp, non_blocking=True); dforce.copy_(hf, non_blocking=True)
        method.computeMF(dpos, dforce, dout)
        hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 10
    gbs = ALG_BYTES / (ms * 1e-3) / 1e9
    return {
        "metric": "FCM Mdof.steps/s @128^3", "value": DOF / 1e6 * 1000.0 / ms, "unit": "Mdof.steps/s",
        "steps_per_s": 1000.0 / ms, "ms_per_step": ms, "value_back_to_back": DOF / 1e6 * 1000.0 / ms_b2b,
        "dtype": "f64", "steps": steps, "warmup": warmup, "gpu_launches": int(launches),
        "config": {"workload": f"BDHI::EulerMaruyama<FCM>, N={N}, {NGRID}^3 grid, Peskin 3pt, eta={ETA}, T={TEMP}, dt={DT}",
                   "l2": "flushed before every step (256 MiB write)"},
        "e2e": {"value": DOF / 1e6 * 1000.0 / e2e_ms, "unit": "Mdof.steps/s", "h2d_bytes_per_step": N * 64,
                "d2h_bytes_per_step": N * 24},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": 5.6e8, "traffic_source": "sum of dram bytes over the kernels of one step, profiles/r01b_fcm_raw.csv "
                                                         "(cold-cache replays: an upper bound, the 51 MB grid stays in L2 between passes)",
                     "algorithmic_bytes_per_step": ALG_BYTES,
                     "note": "whole step against the compulsory 7 grid passes + particle I/O (SURVEY 8(d))"},
    }


def run_distributed(dev, hbm_peak, steps=100, warmup=10):
    """Strong scaling of the SAME config over all ranks (z-slab FCM, uammd_b200.multigpu.DistributedFCM): every rank
    steps the replicated particle set; device time per step, max over ranks. Returns None except on rank 0."""
    import torch.distributed as dist
    from uammd_b200.fcm import Peskin3, _declare, _prec
    from uammd_b200.md import _ptr, _stream_ptr
    from uammd_b200.multigpu import DistributedFCM
    from uammd_b200._lib import check
    from uammd_b200 import lib
    world, rank = dist.get_world_size(), dist.get_rank()
    pos, force = inputs()
    dpos, dforce = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    fcm = DistributedFCM(L, (NGRID,) * 3, Peskin3(L / NGRID), ETA, N, seed=1234)
    MF = torch.zeros(N, 3, dtype=torch.float64, device=dev)
    l = _declare()
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        fcm.computeHydrodynamicDisplacements(dpos, dforce, temperature=TEMP, prefactor=1.0 / math.sqrt(DT), out=MF)
        check(l.ub200_bdhi_euler_update(_prec(dpos.dtype), _ptr(dpos), None, _ptr(MF), None, None, N,
                                        math.sqrt(2 * DT * TEMP), DT, 0, _stream_ptr()))

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(); dist.barrier()
    step(); step()  # untimed: re-aligns the ranks on the device after the host-side barrier (see bench.py)
    l0 = lib().ub200_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        scrub.fill_(3)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize(); dist.barrier()
    launches = lib().ub200_launch_count() - l0
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms_b2b = e0.elapsed_time(e1) / steps
    err = fcm.errorFlag()
    if os.environ.get("UB200_DIST_PROFILE"):
        import ctypes as C
        ph = (C.c_double * 12)()
        n = fcm.lib.ub200_fcm_dist_profile(fcm._h, ph)
        names = ["sort", "spread", "fft_x", "fft_y+transpose", "barrier1", "fused_z+transpose", "barrier2", "ifft_y+x", "gather",
                 "push", "barrier3", "scatter"]
        print(f"[rank {rank}] phases over {n} calls (us): " + ", ".join(f"{k}={1e3 * v:.1f}" for k, v in zip(names, ph)), file=sys.stderr)
    t = torch.tensor([ms, ms_b2b], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_b2b = (float(x) for x in t.cpu())
    fcm.close()
    if rank != 0:
        return None
    gbs = ALG_BYTES / (ms * 1e-3) / 1e9 / world
    return {
        "metric": "FCM Mdof.steps/s @128^3", "value": DOF / 1e6 * 1000.0 / ms, "unit": "Mdof.steps/s", "n_gpus": world,
        "scaling": "strong", "steps_per_s": 1000.0 / ms, "ms_per_step": ms, "value_back_to_back": DOF / 1e6 * 1000.0 / ms_b2b,
        "dtype": "f64", "steps": steps, "warmup": warmup, "gpu_launches": int(launches), "barrier_timeouts": err,
        "config": {"workload": f"BDHI::EulerMaruyama<FCM>, N={N}, {NGRID}^3 grid, Peskin 3pt, eta={ETA}, T={TEMP}, dt={DT}",
                   "l2": "flushed before every step (256 MiB write)",
                   "parallelism": f"z-slab decomposition over {world} GPUs: {NGRID // world} planes per rank, FFT transposes by "
                                  "NVLink peer stores fused into the y/z passes (halo planes included), 3 device-side barriers per step, no NCCL on the data path"},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                     "note": "compulsory bytes of the whole step (SURVEY 8(d)) per GPU"},
    }


def run_reference(root, steps=100, warmup=10):
    exe = os.path.join(root, "oracle", "_ref", "ref_fcm")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_fcm was not built"}
    pos, force = inputs()
    with tempfile.TemporaryDirectory() as td:
        pos.tofile(os.path.join(td, "p.bin")); force.tofile(os.path.join(td, "f.bin"))
        out = subprocess.run([exe, "time", "peskin3", str(N), str(L), str(NGRID), str(ETA), "1e-3", str(TEMP), str(DT),
                              str(warmup), str(steps), "1", os.path.join(td, "p.bin"), os.path.join(td, "f.bin")],
                             check=True, capture_output=True, text=True, timeout=1200).stdout
    r = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
    return {"metric": "FCM Mdof.steps/s @128^3", "value": DOF / 1e6 * r["steps_per_s"], "unit": "Mdof.steps/s",
            "steps_per_s": r["steps_per_s"], "ms_per_step": r["ms_per_step"], "dtype": "f64",
            "what": "unmodified reference FCM_impl<Peskin::threePoint> (cuFFT) on the same B200, same protocol"}
