"""CPU: the index geometry of the column traversal (include/uammd_b200/colgeom.h), compiled for the host.

colgeom_check: every home half cell reaches exactly the 5 x 5 x 5 stencil (cells + image shifts, in staging order).
colpairs_check: emulation of the staged data path (canonical coordinates, row pieces, image shifts) finds exactly the
brute-force minimum-image pair set, including particles outside the primary box and at +-L/2."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["colgeom_check", "colpairs_check"])
def test_host_check(tmp_path, name):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-o", exe, os.path.join(ROOT, "tests", "host", name + ".cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
