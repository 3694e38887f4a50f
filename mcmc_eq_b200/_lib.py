"""ctypes binding of libmcmceq_b200.so (include/mcmceq_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmcmceq_b200.so")
_lib = None

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)


class MqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmcmceq_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load the CUDA library.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `python -m mcmc_eq_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.mq_version.restype = C.c_char_p
        L.mq_last_error.restype = C.c_char_p
        L.mq_launch_count.restype = C.c_int64
        L.mq_time_2d.argtypes = [fp, fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
        L.mq_eikonal_batch.argtypes = [fp, ip, C.c_int, C.c_int, C.c_int, fp, ip, C.c_int]
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MqError(rc, lib().mq_last_error().decode())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def eikonal_batch(slow, src_iz, nxmod: int, device: int = 0, return_status: bool = False):
    """n solves: slow[n, nz] (h/v per depth cell), src_iz[n] -> t[n, nxmod, nz]."""
    slow = _f32(slow)
    n, nz = slow.shape
    iz = np.ascontiguousarray(src_iz, dtype=np.int32)
    assert iz.shape == (n,)
    t = np.empty((n, nxmod, nz), np.float32)
    st = np.zeros(n, np.int32)
    rc = lib().mq_eikonal_batch(slow.ctypes.data_as(fp), iz.ctypes.data_as(ip), n, nxmod, nz,
                                t.ctypes.data_as(fp), st.ctypes.data_as(ip), device)
    if return_status:
        return t, st, rc
    check(rc)
    return t


def time_2d(hs, xs: float, ys: float, eps_init: float = 0.001) -> np.ndarray:
    """Drop-in call of the reference's time_2d signature (hs[nx, ny] -> t[nx, ny])."""
    hs = _f32(hs)
    nx, ny = hs.shape
    t = np.zeros((nx, ny), np.float32)
    check(lib().mq_time_2d(hs.ctypes.data_as(fp), t.ctypes.data_as(fp), nx, ny, xs, ys, eps_init, 0))
    return t
