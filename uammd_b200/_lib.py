"""ctypes loader for libuammd_b200.so (the C ABI declared in include/uammd_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import of any compute
entry point. torch is used only to own device memory and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuammd_b200.so")


class UB200Error(RuntimeError):
    pass


class CellListView(C.Structure):
    _fields_ = [
        ("d_cellStart", C.c_void_p),
        ("d_cellEnd", C.c_void_p),
        ("d_sortPos", C.c_void_p),
        ("d_groupIndex", C.c_void_p),
        ("VALID_CELL", C.c_uint32),
        ("cellDim", C.c_int * 3),
        ("numberParticles", C.c_int),
        ("d_binStart", C.c_void_p),
        ("nbins", C.c_int),
    ]


_lib = None

_F3 = C.c_float * 3
_I3 = C.c_int * 3
_D3 = C.c_double * 3


def _declare(lib):
    vp, i, f, d, u32 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_uint32
    fp = C.POINTER(C.c_float)
    sig = {
        "ub200_error_string": (C.c_char_p, [i]),
        "ub200_last_cuda_error": (i, []),
        "ub200_version": (C.c_char_p, []),
        "ub200_launch_count": (C.c_ulonglong, []),
        "ub200_celllist_create": (i, [C.POINTER(vp)]),
        "ub200_celllist_destroy": (i, [vp]),
        "ub200_neighbour_celldim_f32": (i, [_F3, f, _I3]),
        "ub200_celllist_build_f32": (i, [vp, vp, vp, i, _F3, _I3, _I3, vp]),
        "ub200_celllist_view_get": (i, [vp, C.POINTER(CellListView)]),
        "ub200_celllist_error_flag": (i, [vp, vp, C.POINTER(i)]),
        "ub200_lj_sum_f32": (i, [vp, fp, i, vp, vp, vp, vp, vp]),
        "ub200_lj_sum_devparams_f32": (i, [vp, vp, i, vp, vp, vp, vp, vp]),
        "ub200_lj_sum_owned_f32": (i, [vp, fp, i, vp, i, i, i, vp]),
        "ub200_nve_half_step_f32": (i, [vp, vp, vp, vp, f, vp, i, f, i, i, vp]),
        "ub200_nve_kick_kick_drift_f32": (i, [vp, vp, vp, i, f, vp]),
        "ub200_md_create": (i, [C.POINTER(vp)]),
        "ub200_md_destroy": (i, [vp]),
        "ub200_md_engine": (vp, [vp]),
        "ub200_md_kinetic_energy_f32": (i, [vp, vp, i, vp, vp]),
        "ub200_ljengine_create": (i, [C.POINTER(vp)]),
        "ub200_ljengine_destroy": (i, [vp]),
        "ub200_ljengine_sum_f32": (i, [vp, vp, vp, i, _F3, _I3, fp, i, vp, vp, vp, vp, i, i, i, vp]),
        "ub200_ljengine_traverse_f32": (i, [vp, vp, i, vp]),
        "ub200_ljengine_last_path": (i, [vp]),
        "ub200_ljengine_grid": (i, [vp, _I3]),
        "ub200_ljengine_error_flag": (i, [vp, vp, C.POINTER(i)]),
        "ub200_md_lj_nve_prepare_f32": (i, [vp, vp, vp, i, _F3, f, fp, i, vp]),
        "ub200_md_lj_nve_run_f32": (i, [vp, vp, vp, vp, i, _F3, f, fp, i, f, i, vp]),
        "ub200_md_lj_nve_run_host_f32": (i, [vp, vp, vp, vp, i, _F3, f, fp, i, f, i, vp]),
    }
    sig["ub200_dpd_sum_f32"] = (i, [vp, vp, f, f, f, f, u32, u32, i, vp, vp, vp])
    sig["ub200_dpd_sum_owned_f32"] = (i, [vp, vp, f, f, f, f, u32, u32, i, vp, i, i, i, vp])
    optional = {}
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    for name, (res, args) in optional.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args


def lib():
    """Load the shared library (once). Raises UB200Error when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(code):
    if code != 0:
        l = lib()
        msg = l.ub200_error_string(code).decode()
        if code == -2:
            msg += f" [cudaError {l.ub200_last_cuda_error()}]"
        raise UB200Error(f"uammd_b200: {msg} (code {code})")


def f3(v):
    return _F3(*[float(x) for x in v])


def i3(v):
    return _I3(*[int(x) for x in v])


def d3(v):
    return _D3(*[float(x) for x in v])
