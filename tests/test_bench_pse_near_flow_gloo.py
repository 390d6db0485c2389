"""CPU, world_size 2 over gloo: the control flow of bench_extra.pse_near_distributed (the multi-GPU PSE near-field leg of
`bench.py --gpus N`) with stand-ins for the CUDA pieces - the engine classes, CUDA events and synchronisation are replaced,
the torch.distributed collectives, the rank agreement and the shape of the returned entry are the real ones. Guards the leg
against host-side slips (names, arguments, a rank that skips a collective) that only a multi-GPU box would otherwise show."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT,) if p not in sys.path]  # also in the spawned workers


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def _worker(rank, world, port, fail_rank, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench_extra
    from uammd_b200 import multigpu
    from uammd_b200 import pse as P
    N = 64
    bench_extra.PSE_N = N
    bench_extra._pse_inputs = lambda: (np.zeros((N, 4), np.float32), np.zeros((N, 4), np.float32))
    calls = {"prepare": 0, "mdot": 0, "noise": 0}

    class FakeNear:
        def __init__(self, pos, par, sys=None, group=None, rank=None, world=None):
            assert pos.shape == (N, 4) and par.temperature == bench_extra.PSE_T

        def prepare(self, stream=None):
            calls["prepare"] += 1

        def Mdot(self, force, MF, stream=None):
            calls["mdot"] += 1
            if rank == fail_rank and calls["mdot"] == 3:
                raise RuntimeError("stand-in failure")

        def noiseAdd(self, out, temperature, prefactor, seed2, stream=None):
            calls["noise"] += 1
            return 9

        def errorFlag(self, stream=None):
            return 0

    class FakePSE:
        def __init__(self, pos, par, sys=None, force=None):
            pass

        def computeMFNearField(self, MF, stream=None, listForNoise=False):
            assert listForNoise

        def _nearNoise(self, out, temperature, prefactor, stream, add=False, reuse=False):
            assert reuse and not add

        def info(self):
            return types.SimpleNamespace(lastLanczosIterations=9)

    multigpu.DistributedPSENearField = FakeNear
    P.PSE = FakePSE
    torch.cuda.Event = _Event
    torch.cuda.synchronize = lambda *a, **k: None
    bench_extra._timed = lambda dev, step, steps, warmup: (step(), 1.5)[1]
    line = bench_extra.pse_near_distributed(torch.device("cpu"), steps=3, warmup=1)
    torch.save({"line": line, "calls": calls}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, fail_rank, port):
    out = str(tmp_path / f"leg{fail_rank}")
    mp.spawn(_worker, args=(2, port, fail_rank, out), nprocs=2, join=True)
    return torch.load(out + ".0", weights_only=False), torch.load(out + ".1", weights_only=False)


def test_leg_returns_one_entry_on_rank_zero(tmp_path):
    r0, r1 = _run(tmp_path, -1, 29541)
    line = r0["line"]
    assert r1["line"] is None
    assert line["n_gpus"] == 2 and line["unit"] == "calls/s" and line["ms_per_step"] > 0 and line["value"] == 1000.0 / line["ms_per_step"]
    assert line["lanczos_iterations"] == 9 and line["single_gpu_lanczos_iterations"] == 9 and line["single_gpu_ms"] == 1.5
    assert line["barrier_error_flag"] == 0 and "32 rows per rank" in line["what"]
    # 1 first call + 1 warm-up + 1 re-alignment + 3 timed, the same on both ranks
    assert r0["calls"] == {"prepare": 6, "mdot": 6, "noise": 6} and r1["calls"] == r0["calls"]


def test_a_failing_rank_becomes_an_error_entry_not_a_hang(tmp_path):
    r0, r1 = _run(tmp_path, 1, 29543)
    assert r1["line"] is None
    assert "error" in r0["line"] and "another rank failed" in r0["line"]["error"]
    r0, r1 = _run(tmp_path, 0, 29545)
    assert "stand-in failure" in r0["line"]["error"] and r1["line"] is None
