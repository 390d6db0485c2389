"""CPU: the C-ABI library loads and exports every symbol include/uammd_b200.h declares; host-only entry
points behave (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            txt = open(os.path.join(inc, fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(ub200_\w+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import uammd_b200
    lib = C.CDLL(uammd_b200.LIB_PATH)
    missing = [n for n in _declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert len(_declared_symbols()) >= 15


def test_error_strings_and_version():
    import uammd_b200
    lib = uammd_b200.lib()
    assert lib.ub200_error_string(0) == b"ok"
    assert b"invalid" in lib.ub200_error_string(-1)
    assert b"sm_100a" in lib.ub200_version()


def test_neighbour_celldim_host_rule():
    from uammd_b200.md import Box, CellList
    assert CellList.gridFor(Box(107.7217), 2.5) == (43, 43, 43)
    assert CellList.gridFor(Box((10.0, 7.4, 100.0)), 2.5) == (4, 1, 40)


def test_invalid_arguments_are_reported_not_crashed():
    import uammd_b200
    from uammd_b200._lib import f3, i3
    lib = uammd_b200.lib()
    assert lib.ub200_celllist_build_f32(None, None, None, 0, f3((1, 1, 1)), i3((1, 1, 1)), i3((1, 1, 1)), None) == -1
    assert lib.ub200_neighbour_celldim_f32(f3((1, 1, 1)), 0.0, i3((0, 0, 0))) == -1
    with pytest.raises(uammd_b200.UB200Error):
        uammd_b200._lib.check(-4)
    # brick decomposition: a brick without a cell, more ranks than mask bits, missing outputs; empty input is fine
    L, per = f3((10, 10, 10)), i3((1, 1, 1))
    assert lib.ub200_brick_classify_f32(None, 0, L, per, i3((4, 4, 4)), i3((5, 1, 1)), None, None, None, None) == -1
    assert lib.ub200_brick_classify_f32(None, 0, L, per, i3((4, 4, 4)), i3((4, 4, 4)), None, None, None, None) == -6
    assert lib.ub200_brick_classify_f32(None, 8, L, per, i3((4, 4, 4)), i3((2, 2, 2)), None, None, None, None) == -1
    assert lib.ub200_brick_classify_f32(None, 0, L, per, i3((4, 4, 4)), i3((2, 2, 2)), None, None, None, None) == 0
    assert lib.ub200_lj_nbody_f32(None, None, 8, L, per, None, 1, None, None, None, None) == -1
    assert lib.ub200_dpd_sum_owned_ids_f32(None, None, C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(1), 0, 0, 8, None, 0,
                                           8, 0, None, None) == -1


def test_lj_parameter_table_matches_reference_rule():
    # LJFunctor::processPairParameters (Potential.cuh:67-82)
    from uammd_b200.md import LJ
    from uammd_b200 import synthetic as syn
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=2.5, sigma=1.0, epsilon=1.0)
    pot.setPotParameters(0, 1, cutOff=3.0, sigma=1.2, epsilon=0.5, shift=True)
    pot.setPotParameters(1, 1, cutOff=2.0, sigma=0.8, epsilon=2.0)
    t = pot.table().reshape(2, 2, 4)
    assert np.array_equal(t[0, 0], syn.lj_params())
    assert np.array_equal(t[0, 1], t[1, 0])
    assert np.array_equal(t[0, 1], syn.lj_params(1.2, 0.5, 3.0, True))
    assert pot.getCutOff() == 3.0
