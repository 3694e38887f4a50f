"""mcmc_eq_b200 -- B200 implementation of mcmc_eq's forward-model / likelihood hot path.

The product is the C-ABI library `libmcmceq_b200.so` (include/mcmceq_b200.h) and the C
command-line front end in `host/`; this package is a thin ctypes binding used by the tests
and bench.py.  There is no CPU fallback: importing `lib()` fails if the library is missing.
"""
from ._lib import (lib, MqError, eikonal_batch, time_2d, read_config, Picks, Models, Sampler,  # noqa: F401
                   MqConfig, MqGrid)
