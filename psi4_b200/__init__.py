"""psi4_b200 -- B200-native density-fitted J/K engine behind psi4's MEM_DF JK interface.

Only what the hot path needs lives here: csrc/ (sm_100a kernels + the C ABI of
include/b200jk.h), lib.py (ctypes binding), dfhelper.py / jk.py (host mirrors of psi4's
DFHelper tables and JK/MemDFJK interface).  No CPU fallback: importing works anywhere, but
constructing an engine without libb200jk.so or without a CUDA device raises.
"""
from .dfhelper import DFHelper  # noqa: F401
from .jk import JK, MemDFJK, PsiException  # noqa: F401
from .lib import B200JKError, Engine  # noqa: F401

__all__ = ["DFHelper", "JK", "MemDFJK", "PsiException", "Engine", "B200JKError"]
