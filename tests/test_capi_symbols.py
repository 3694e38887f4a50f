"""The C-ABI library loads (no GPU needed) and exports every function include/*.h declares."""
import ctypes as C
import os
import re

from tests import util


def _declared(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mq(?:io)?_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mcmc_eq_b200 import build, _lib
    build.build_cuda()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared(os.path.join(util.ROOT, "include", "mcmceq_b200.h")) + \
        _declared(os.path.join(util.ROOT, "mcmc_eq_b200", "host", "mq_io.h"))
    assert len(names) > 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.mq_version.restype = C.c_char_p
    assert b"sm_100a" in L.mq_version()


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call and returns codes, never exits."""
    import numpy as np
    import mcmc_eq_b200 as mq
    L = mq.lib()
    hs = np.full((6, 5), 0.4, np.float32)
    t = np.zeros((6, 5), np.float32)
    fp = C.POINTER(C.c_float)
    assert L.mq_time_2d(hs.ctypes.data_as(fp), t.ctypes.data_as(fp), 6, 5, 1.0, 2.0, 0.001, 0) == -3   # xs != 0
    assert L.mq_time_2d(hs.ctypes.data_as(fp), t.ctypes.data_as(fp), 6, 5, 0.0, 2.5, 0.001, 0) == -3   # ys not a node
    hs[2, 1] = 0.5
    assert L.mq_time_2d(hs.ctypes.data_as(fp), t.ctypes.data_as(fp), 6, 5, 0.0, 2.0, 0.001, 0) == -3   # 2-D medium
    assert L.mq_time_2d(None, t.ctypes.data_as(fp), 6, 5, 0.0, 2.0, 0.001, 0) == -1
    assert b"mq_time_2d" in L.mq_last_error()
    assert L.mq_step(None, 1, None) == -1 and L.mq_forward(None, 3, None, None) == -1
