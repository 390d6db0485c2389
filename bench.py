#!/usr/bin/env python
"""bench.py - headline benchmark of uammd_b200 (contract in the task statement).

Workload (BASELINE.json configs[1]): PairForces<LJ, CellList> + VerletNVE molecular dynamics, N = 1e6
particles, rho = 0.8, rc = 2.5 sigma, fp32, one B200. A "step" is one VerletNVE::forwardTime():
cell-list rebuild + LJ forces + velocity-Verlet update. Metric: MD steps/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles N]

Timing protocol: W untimed warm-up steps, then exactly K steps, each bracketed by CUDA events on the
launching stream with a 256 MiB L2-evicting write before every step (the working set, ~90 MB, would
otherwise sit in the 126 MB L2); ms_per_step = sum(step times)/K, max over ranks. `value_back_to_back`
reports the same K steps enqueued back to back (how an application runs them).

Multi GPU (N > 1, launched by torchrun): ONE 1e6-particle system over all ranks (strong scaling) by brick domain
decomposition with a device-side halo exchange (uammd_b200/brickmd.py, uammd_b200/csrc/brick_md.cu): owned particles +
ghost half cells, one peer-to-peer exchange per step. value = steps/s of that single system.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RHO, RC, DT, TEMP = 0.8, 2.5, 0.005, 1.0
ALG_BYTES_PAIR = 52          # per particle: R sortPos 16 + R groupIndex 4 + RMW force 32
ALG_BYTES_STEP = 216         # per particle and step: build 36 + traverse 52 + integrate 112 + zero 16
FLOP_PER_CANDIDATE = 25
NCU_TRAFFIC_PAIR = 22.61e6   # DRAM bytes per ljColumnTraversal launch at N = 1e6 (22.60 MB read + 0.01 MB written: the force writes
                             # stay in the 126 MB L2), profiles/r02_lj_raw.csv
FP32_PEAK_TFLOPS = 74.4      # 148 SM x 128 lanes x 2 x 1.965 GHz


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def workload(N):
    from uammd_b200 import synthetic as syn
    Lb = syn.lj_box_length(N, RHO)
    return Lb, syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, TEMP, seed=7)


def run_reference(args):
    """Reference arm: the UNMODIFIED reference (PairForces<Potential::LJ, CellList> + VerletNVE) compiled from
    /root/reference by oracle/Makefile into oracle/_ref/ref_lj. UAMMD has no CPU implementation - its only
    implementation of this path is CUDA - so this arm runs on the same B200, same workload, same protocol."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_lj")
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_lj was not built (no reference tree at build time)"}))
        return 0
    N = args.particles
    Lb, pos, vel = workload(N)
    with tempfile.TemporaryDirectory() as td:
        pos.tofile(os.path.join(td, "pos.bin"))
        vel.tofile(os.path.join(td, "vel.bin"))
        cmd = [exe, "md", str(N), str(Lb), str(Lb), str(Lb), str(RC), "1", "1", str(DT), str(args.warmup + args.equil),
               str(args.steps), "1", "1", os.path.join(td, "pos.bin"), os.path.join(td, "vel.bin"), "-"]
        t0 = time.time()
        out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=3000).stdout
        wall = time.time() - t0
    summ = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    s = [x for x in summ if x.get("mode") == "md_summary"][0]
    ms = s["ms_per_step_mean"]
    val = 1000.0 / ms
    line = {
        "impl": "reference", "metric": "MD steps/s @1e6 LJ particles", "value": val, "unit": "steps/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PairForces<LJ,CellList> + VerletNVE, N={N}, rho={RHO}, rc={RC}, dt={DT}, FCC start T={TEMP}",
                   "l2": "flushed before every step (256 MiB write)", "equilibration_steps": args.equil,
                   },
        "arm": "reference on the B200 GPU: the reference's only implementation of this path is CUDA (unmodified, sm_100a build)",
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": 1, "kind": "reference",
                         "sample": f"{args.steps} steps of the full workload; 1 host thread driving the reference's CUDA kernels"},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    if not args.no_fcm:
        import bench_fcm as fcm_bench
        line["fcm"] = fcm_bench.run_reference(ROOT, steps=args.fcm_steps)
    if not args.no_extra:
        import bench_extra as extra_bench
        for name, fn in (("verlet", lambda: extra_bench.verlet_reference(ROOT, N, Lb, pos, vel, RC, DT, equil=args.equil)),
                         ("pse", lambda: extra_bench.pse_reference(ROOT)), ("bd", lambda: extra_bench.bd_reference(ROOT)),
                         ("langevin", lambda: extra_bench.langevin_reference(ROOT)),
                         ("poisson", lambda: extra_bench.poisson_reference(ROOT))):
            try:
                line[name] = fn()
            except Exception as e:  # a secondary leg must not take the headline down
                line[name] = {"error": repr(e)[:300]}
    print(json.dumps(line))
    return 0


def cpu_baseline(N, nsteps=3):
    """The oracle's OpenMP restatement of the same MD step on the host cores (bounded sample)."""
    from oracle import oracle as orc
    from uammd_b200 import synthetic as syn
    Lb, pos, vel = workload(N)
    md = orc.MDOracle((Lb,) * 3, RC, syn.lj_params(), DT, pos, vel)
    md.step(1)
    t0 = time.time()
    md.step(nsteps)
    dt = time.time() - t0
    return {"value": nsteps / dt, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{nsteps} steps of the N={N} workload (OpenMP, all host cores) after 1 warm-up step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--equil", type=int, default=300, help="untimed equilibration steps (melts the FCC start)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fcm", action="store_true")
    ap.add_argument("--fcm-steps", type=int, default=100)
    ap.add_argument("--pse-near-limit", type=int, default=180, help="seconds the multi-GPU PSE near-field leg may take (N > 1)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary legs (VerletList MD, PSE, BD ideal)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import uammd_b200
    from uammd_b200.md import Box, LJ, LJMD, PairForces

    uammd_b200.lib()  # no CPU fallback: raises if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]  # both levels print NCCL's version banner on stdout; rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    N = args.particles
    Lb, pos, vel = workload(N)
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=RC, sigma=1.0, epsilon=1.0)
    box = Box(Lb)
    if world > 1:
        return main_distributed(args, dev, world, rank, local, N, Lb, pos, vel, pot, box)
    md = LJMD(box, pot, DT)
    p = torch.from_numpy(pos).to(dev)
    v = torch.from_numpy(vel).to(dev)
    f = torch.zeros(N, 4, device=dev)
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = uammd_b200.lib()

    md.run(p, v, f, args.equil + args.warmup)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, L2 evicted before each, CUDA events on the launching stream ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = lib.ub200_launch_count()
    barrier()
    with ClockSampler(local) as clk:
        for a, b in evs:
            scrub.fill_(1)
            a.record()
            md.run(p, v, f, 1)
            b.record()
        barrier()
    launches = lib.ub200_launch_count() - launches0
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = float(np.sum(ms_steps) / args.steps)

    # same K steps back to back (L2 warm, launch overheads overlapped)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    md.run(p, v, f, args.steps)
    e1.record()
    barrier()
    ms_b2b = e0.elapsed_time(e1) / args.steps

    # ---- roofline of the dominant kernel (LJ column traversal), timed alone on its stream ----
    from uammd_b200.md import CellList, LJEngine
    eng = LJEngine()
    ftmp = torch.zeros(N, 4, device=dev)
    eng.sum(p, box, pot.table(), pot.ntypes, force=ftmp, accumulate=False)
    assert eng.lastPath() == "column", eng.lastPath()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in kev:
        scrub.fill_(2)
        a.record()
        eng.traverse(ftmp, accumulate=False)
        b.record()
    torch.cuda.synchronize()
    t_pair_ms = float(np.median([a.elapsed_time(b) for a, b in kev]))
    peak, peak_src = measured_peaks()
    ncells = int(np.prod(CellList.gridFor(box, RC)))
    cand = 27.0 * N / ncells                       # candidates of the reference's 27-cell walk: the ALGORITHMIC work
    cand_exec = 125.0 * N / float(np.prod(eng.grid()))  # candidates the engine actually tests (5^3 half cells)
    pair_gbs = ALG_BYTES_PAIR * N / (t_pair_ms * 1e-3) / 1e9
    pair_tflops = cand * FLOP_PER_CANDIDATE * N / (t_pair_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "ljColumnTraversal", "bound": "hbm", "achieved": pair_gbs, "peak": peak, "unit": "GB/s",
        "frac": pair_gbs / peak, "traffic": NCU_TRAFFIC_PAIR if N == 1_000_000 else None, "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r02_lj_raw.csv",
        "peak_source": peak_src, "kernel_ms": t_pair_ms,
        "algorithmic_bytes_per_launch": ALG_BYTES_PAIR * N,
        "note": "the pair kernel is bound by instruction issue, not HBM (SURVEY 8(d)); see fp32 and pipeline",
        "fp32": {"achieved_tflops": pair_tflops, "peak_tflops": FP32_PEAK_TFLOPS, "frac": pair_tflops / FP32_PEAK_TFLOPS,
                 "candidates_per_particle": cand, "flop_per_candidate": FLOP_PER_CANDIDATE,
                 "what": "algorithmic flops of the reference's 27-cell walk (SURVEY 8(d): 339.6 candidates x 25 flop) per kernel time, "
                         "against the nominal non-tensor FP32 peak; the engine tests fewer candidates to get the same forces",
                 "candidates_tested_per_particle": cand_exec,
                 "executed_tflops": cand_exec * FLOP_PER_CANDIDATE * N / (t_pair_ms * 1e-3) / 1e12},
        "pipeline": {"bytes_per_step": ALG_BYTES_STEP * N, "achieved": ALG_BYTES_STEP * N / (ms_per_step * 1e-3) / 1e9,
                     "unit": "GB/s", "frac": ALG_BYTES_STEP * N / (ms_per_step * 1e-3) / 1e9 / peak},
    }

    # ---- e2e: the public host-buffer entry point, state round trip over PCIe every step ----
    hp, hv, hf = p.cpu().pin_memory(), v.cpu().pin_memory(), torch.zeros(N, 4).pin_memory()
    md2 = LJMD(box, pot, DT)
    for _ in range(3):
        md2.runHost(hp, hv, None, 1)
    n_e2e = max(10, min(args.steps, 50))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        md2.runHost(hp, hv, None, 1)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    # device-resident state through the same public API, one scalar (kinetic energy) read back per step
    ke_host = torch.zeros(n_e2e, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        md.run(p, v, f, 1)
        md.kineticEnergy(v, ke_host, k)  # the step's result (ub200_md_kinetic_energy_f32), copied down asynchronously
    torch.cuda.synchronize()
    e2e_res_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    ke = float(ke_host[-1])

    # ---- aggregate over ranks ----
    t = torch.tensor([ms_per_step, ms_b2b, e2e_ms, e2e_res_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, ms_b2b, e2e_ms, e2e_res_ms = (float(x) for x in t.cpu())
    if rank == 0:
        line = {
            "metric": "MD steps/s @1e6 LJ particles", "value": world * 1000.0 / ms_per_step, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PairForces<LJ,CellList> + VerletNVE, N={N}, rho={RHO}, rc={RC}, dt={DT}, FCC start T={TEMP}",
                       "l2": "flushed before every step (256 MiB write)", "equilibration_steps": args.equil,
                       },
            "arm": "uammd_b200, single GPU",
            "value_back_to_back": world * 1000.0 / ms_b2b,
            "clocks": clk.summary(),
            "e2e": {"value": world * 1000.0 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": N * 28,
                    "d2h_bytes_per_step": N * 28,
                    "what": "ub200_md_lj_nve_run_host_f32: every step the pinned host state (pos + vel) is uploaded, F(t) recomputed, one "
                            "step taken and pos + vel downloaded (transfers overlapped with the two force evaluations). The reference arm "
                            "keeps its state on the device (0 bytes/step): like for like is `value`, or e2e_resident"},
            "e2e_resident": {"value": world * 1000.0 / e2e_res_ms, "unit": "steps/s", "d2h_bytes_per_step": 4,
                             "what": "same API with device-resident state, kinetic energy copied to pinned host memory every step "
                                     "(asynchronously; one synchronisation after the last step)", "last_ke": ke},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(N)
        if not args.no_fcm and world == 1:
            try:
                import bench_fcm as fcm_bench
                line["fcm"] = fcm_bench.run(dev, peak, steps=args.fcm_steps)
            except ImportError:
                pass
        if not args.no_extra and world == 1:
            import bench_extra as extra_bench
            for name, fn in (("verlet", lambda: extra_bench.verlet(dev, N, Lb, pos, vel, RC, DT, equil=args.equil)),
                             ("pse", lambda: extra_bench.pse(dev)), ("bd", lambda: extra_bench.bd_ideal(dev)),
                             ("langevin", lambda: extra_bench.langevin(dev)), ("dpd", lambda: extra_bench.dpd(dev)),
                             ("poisson", lambda: extra_bench.poisson(dev))):
                try:
                    line[name] = fn()
                except Exception as e:  # a secondary leg must not take the headline down
                    line[name] = {"error": repr(e)[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main_distributed(args, dev, world, rank, local, N, Lb, pos, vel, pot, box):
    """N > 1: strong scaling of the single 1e6-particle system over bricks with a device-side halo exchange
    (uammd_b200.brickmd.BrickLJMD: ub200_brick_* / ub200_halo_exchange_* in the C ABI)."""
    import torch
    import torch.distributed as dist
    import uammd_b200
    from uammd_b200.brickmd import BrickLJMD
    lib = uammd_b200.lib()
    md = BrickLJMD(box, pot, DT, N, rank, world)
    md.connect()
    md.setGlobalState(torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev))
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    md.run(args.equil + args.warmup)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = lib.ub200_launch_count()
    barrier()
    with ClockSampler(local) as clk:
        # ranks leave the host-side barrier (and start their clock sampler) milliseconds apart; the steps themselves run in
        # lockstep through the device-side flags, so two untimed steps put every rank on the same device timeline - without
        # them the FIRST timed step of some rank measures that host skew (2 - 30 ms) instead of a step
        md.run(2)
        for a, b in evs:
            scrub.fill_(1)
            a.record()
            md.run(1)
            b.record()
        barrier()
    launches = lib.ub200_launch_count() - launches0
    per_step = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = float(np.sum(per_step) / args.steps)
    if os.environ.get("UB200_BENCH_DUMP"):  # diagnostics: the distribution of the per-step times of this rank
        ps = np.sort(per_step)
        print(f"[rank {rank}] per-step ms: min {ps[0]:.3f} median {ps[len(ps) // 2]:.3f} p90 {ps[int(0.9 * len(ps))]:.3f} "
              f"max {ps[-1]:.3f} first5 {[round(x, 3) for x in per_step[:5]]} argmax {int(np.argmax(per_step))}", file=sys.stderr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    md.run(args.steps)
    e1.record()
    barrier()
    ms_b2b = e0.elapsed_time(e1) / args.steps
    # e2e: every rank round-trips ITS owned block (pos + vel) through pinned host memory every step
    cap = md.info().capacity
    hp, hv = torch.zeros(cap, 4).pin_memory(), torch.zeros(cap, 3).pin_memory()
    n_e2e = max(10, min(args.steps, 50))
    moved = 0
    barrier()

    def e2e_step():
        md.run(1)
        no, _, _ = md.counts()          # synchronises: the host needs the size of its block
        md.downloadOwned(hp, hv, no)
        torch.cuda.synchronize()
        md.uploadOwned(hp, hv, no)
        return no

    e2e_step(); e2e_step()  # untimed: the ranks leave the barrier milliseconds apart, the steps run in lockstep
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        moved += e2e_step() * 28
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    no, nl, err = md.counts()
    t = torch.tensor([ms_per_step, ms_b2b, e2e_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, ms_b2b, e2e_ms = (float(x) for x in t.cpu())
    tot = torch.tensor([moved / n_e2e, float(no), float(nl), float(err)], device=dev, dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak, peak_src = measured_peaks()
        line = {
            "metric": "MD steps/s @1e6 LJ particles", "value": 1000.0 / ms_per_step, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PairForces<LJ,CellList> + VerletNVE, N={N}, rho={RHO}, rc={RC}, dt={DT}, FCC start T={TEMP}",
                       "l2": "flushed before every step (256 MiB write)", "equilibration_steps": args.equil,
                       },
            "arm": f"uammd_b200, {world} GPUs: brick decomposition {md.rankGrid} with ghost half cells, one peer-to-peer halo "
                   f"exchange per step (NVLink stores, no NCCL and no host round trip on the step path)",
            "bricks": {"rank_grid": list(md.rankGrid), "owned_total": int(tot[1]), "local_total_with_ghosts": int(tot[2]),
                       "error_flags": int(tot[3])},
            "value_back_to_back": 1000.0 / ms_b2b, "clocks": clk.summary(),
            "e2e": {"value": 1000.0 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": int(tot[0]), "d2h_bytes_per_step": int(tot[0]),
                    "what": "after every step each rank downloads its owned block (pos + vel) to pinned host memory and uploads it again"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": ALG_BYTES_STEP * N / (ms_per_step * 1e-3) / 1e9 / world, "peak": peak,
                         "unit": "GB/s", "frac": ALG_BYTES_STEP * N / (ms_per_step * 1e-3) / 1e9 / world / peak, "traffic": None,
                         "note": "whole-step pipeline bytes per GPU (216 B/particle/step over all ranks); see the N=1 line for the pair kernel"},
        }
        line_out = line
    dpd_line = None
    if not args.no_extra:
        import bench_extra as extra_bench
        try:
            dpd_line = extra_bench.dpd(dev, world=world, rank=rank)
        except Exception as e:  # a secondary leg must not take the headline down
            dpd_line = {"error": repr(e)[:300]}
    fcm_line = None
    if not args.no_fcm:
        import bench_fcm as fcm_bench
        peak, _ = measured_peaks()
        fcm_line = fcm_bench.run_distributed(dev, peak, steps=args.fcm_steps)
    pse_line = None
    if not args.no_extra:
        try:
            pse_line = extra_bench.pse_far_distributed(dev)
        except Exception as e:  # a secondary leg must not take the headline down
            pse_line = {"error": repr(e)[:300]}
    if rank == 0:
        if pse_line is not None:
            line_out["pse_far"] = pse_line
        if fcm_line is not None:
            line_out["fcm"] = fcm_line
        if dpd_line is not None:
            line_out["dpd"] = dpd_line
    # The PSE near-field leg is the newest multi-GPU path: it runs LAST, behind a watchdog. Whatever happens in it (an error
    # on one rank, a rank that never returns), rank 0 prints its ONE line with the legs measured so far and every rank leaves
    # within the limit.
    if not args.no_extra:
        def abandoned():
            if rank == 0:
                line_out["pse_near"] = {"error": "the leg did not finish within %d s and was abandoned" % args.pse_near_limit}
                print(json.dumps(line_out), flush=True)
        pse_near_line = extra_bench.run_bounded(args.pse_near_limit, lambda: extra_bench.pse_near_distributed(dev), abandoned)
        if rank == 0 and pse_near_line is not None:
            line_out["pse_near"] = pse_near_line
    if rank == 0:
        print(json.dumps(line_out), flush=True)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
