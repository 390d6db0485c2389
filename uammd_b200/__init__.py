"""uammd_b200: B200-native (sm_100a) engine for UAMMD's two data-parallel hot paths.

Python here is the thin host-side mirror of the reference's operator interface used by tests and bench;
the product is the C-ABI CUDA library (include/uammd_b200.h, uammd_b200/csrc) plus the C++14 glue headers
in include/uammd_b200/ that drop into UAMMD programs.
"""
from ._lib import UB200Error, lib, LIB_PATH  # noqa: F401
from . import synthetic  # noqa: F401
