"""The two shipped example data sets as test inputs.

/root/reference does not exist on the GPU box, so tools/make_golden.py parses the example
files once (in the build container) and stores the parsed arrays under tests/golden/;
`materialise()` writes them back out as a config_eqx.dat / pick file pair in the reference's
formats (src/mcmc_eq.c:345-388, :1238-1294) for anything that needs real files."""
from __future__ import annotations

import json
import os

import numpy as np

from tests.util import GOLDEN

def load(name: str):
    """-> (config dict, arrays dict) of 'example' or 'example2'."""
    with open(os.path.join(GOLDEN, f"inputs_{name}.json")) as f:
        cfg = json.load(f)
    arr = dict(np.load(os.path.join(GOLDEN, f"inputs_{name}.npz")))
    return cfg, arr


from mcmc_eq_b200.io import write_config, write_picks  # noqa: E402,F401


def materialise(name: str, directory: str, **cfg_override):
    cfg, arr = load(name)
    os.makedirs(directory, exist_ok=True)
    cfgp, pkp = os.path.join(directory, "config_eqx.dat"), os.path.join(directory, "picks")
    write_config(cfg, cfgp, **cfg_override)
    write_picks(arr, pkp)
    return cfgp, pkp
