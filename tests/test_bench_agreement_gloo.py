"""CPU, world_size 2 over gloo: bench_extra.RankAgreement, the host logic that keeps the ranks of a multi-GPU bench leg on one
sequence of collectives when only some of them fail (an exception in a rank's own step, a timed-out peer barrier): every
rank must learn about the failure from the same all-reduce, nobody may be left waiting."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT,) if p not in sys.path]  # also in the spawned workers


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench_extra import RankAgreement
    dev = torch.device("cpu")
    ran = []

    def ok():
        ran.append("ok")

    def boom():
        raise RuntimeError("rank 1 only")

    # 1. nobody fails
    a = RankAgreement(dev, lambda: 0)
    a.guarded(ok)
    r1 = a.failed()
    # 2. an exception on rank 1 only: both ranks agree, later steps of the failed rank are skipped, the other keeps stepping
    b = RankAgreement(dev, lambda: 0)
    b.guarded(boom if rank == 1 else ok)
    b.guarded(ok)
    r2 = b.failed()
    # 3. a barrier time-out flag on rank 0 only
    c = RankAgreement(dev, lambda: 7 if rank == 0 else 0)
    c.guarded(ok)
    r3 = c.failed()
    # 4. the flag reader itself throws on rank 1 (a dead context): still one collective per rank
    def dead():
        raise RuntimeError("no context")
    d = RankAgreement(dev, dead if rank == 1 else (lambda: 0))
    d.guarded(ok)
    r4 = d.failed()
    torch.save({"r": [r1, r2, r3, r4], "msg": [a.failure[0], b.failure[0], c.failure[0], d.failure[0]], "ran": len(ran)}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_ranks_agree_on_a_failure_of_one_of_them(tmp_path):
    out = str(tmp_path / "agree")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert r0["r"] == [False, True, True, True] and r1["r"] == r0["r"]
    assert r0["msg"][0] is None and r1["msg"][0] is None
    assert r0["msg"][1] is None and "rank 1 only" in r1["msg"][1]
    assert r0["msg"][2] == "a peer barrier timed out" and r1["msg"][2] is None
    assert r0["msg"][3] is None and "no context" in r1["msg"][3]
    assert r0["ran"] == 5 and r1["ran"] == 3   # rank 1 skipped the step after its failure in scenario 2


def _bounded(code, timeout=60):
    import subprocess
    import time
    t0 = time.time()
    r = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r)\nimport time\nfrom bench_extra import run_bounded\n%s" % (ROOT, code)],
                       capture_output=True, text=True, timeout=timeout)
    return r, time.time() - t0


def test_watchdog_lets_a_leg_that_returns_through():
    r, _ = _bounded("print(run_bounded(30, lambda: {'value': 1}, lambda: print('TIMEOUT')))\nprint('after')")
    assert r.returncode == 0 and r.stdout.splitlines()[-2:] == ["{'value': 1}", "after"] and "TIMEOUT" not in r.stdout


def test_watchdog_turns_an_exception_into_an_error_entry():
    r, _ = _bounded("def f():\n    raise RuntimeError('boom')\nprint(run_bounded(30, f, lambda: print('TIMEOUT')))")
    assert r.returncode == 0 and "'error'" in r.stdout and "boom" in r.stdout and "TIMEOUT" not in r.stdout


def test_watchdog_abandons_a_leg_that_hangs_and_exits_cleanly():
    """the line is printed once, by the time-out handler, and the process leaves with code 0 well before the leg would end"""
    r, dt = _bounded("print(run_bounded(1.0, lambda: time.sleep(40), lambda: print('LINE')))\nprint('never')")
    assert r.returncode == 0 and r.stdout.count("LINE") == 1 and "never" not in r.stdout
    assert dt < 30
