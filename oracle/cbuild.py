"""Builds the oracle's C SpMM (gcc + OpenMP) -- TEST INFRASTRUCTURE, see oracle/__init__.py."""
import ctypes
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "csrc" / "spmm.c"
LIB = HERE / "_build" / "liboracle_spmm.so"


def build(force=False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O3", "-fopenmp", "-mavx2", "-mfma", "-shared", "-fPIC", str(SRC), "-o", str(LIB)],
                   check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        _lib.oracle_spmm_sum_f32.argtypes = [vp, vp, vp, vp, i64, i64]
        _lib.oracle_spmm_sum_f32.restype = None
        _lib.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.oracle_get_max_threads.restype = ctypes.c_int
    return _lib


def set_threads(n: int):
    """Threads of the C SpMM (OpenMP) and of torch's CPU kernels."""
    import torch
    lib().oracle_set_threads(int(n))
    torch.set_num_threads(int(n))
