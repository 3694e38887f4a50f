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

CONFIG_LINES = [  # (fields, comment) per line of config_eqx.dat
    (["h"], "forward dx"), (["nx"], "forward NX"), (["ny"], "forward NY"), (["nz"], "forward NZ"),
    (["x0"], "model starts at X0"), (["y0"], "model starts at Y0"), (["z0"], "model starts at Z0"),
    (["max_dim"], "max # of cells/layers"), (["vpmin"], "minimum vel"), (["vpmax"], "maximum vel"),
    (["vpvsmin"], "minimum vpvs"), (["vpvsmax"], "maximum vpvs"), (["noise_min"], "minimum noise"),
    (["noise_max"], "maximum noise"), (["residual_min"], "min residual"), (["residual_max"], "max residual"),
    (["sdevx"], "sdev x (unused)"), (["sdevy"], "sdev y (unused)"), (["sdevz"], "sdev for z"), (["sdevvp"], "sdev for vel"),
    (["sdevvpvs"], "sdev for vp/vs"), (["sdevn"], "sdev for noise"), (["sdevxs", "epi_search"], "sdev x EQ, epicentre factor"),
    (["sdevys"], "sdev y EQ"), (["sdevzs"], "sdev z EQ"), (["sdevresidual"], "sdev residual"),
    (["inv_control"], "min layer thickness / LVZ switch"),
    (["reference_station", "scor_flag", "ref_statcor_P", "ref_statcor_S"], "reference station + flag"),
    (["tria"], "0 = Voronoi"), (["j_max_start", "j_max_main"], "number of models in chain"), (["deci"], "output every nth model"),
    (["true_random", "eikonal"], "seed (<=0 random), 1 = eikonal"), (["dstring_start", "dstring_main"], "proposal letters"),
    (["aflag", "inp_model_switch"], "0 mcmc, 1 prior only, 3 start from model.dat"), (None, "unused"),
    (["start_vp", "sdev_start_vp", "start_vp_grad"], "vp to start with"), (["start_vpvs", "sdev_start_vpvs"], "vp/vs to start with"),
    (["start_cell_number", "sdev_start_cell_number"], "cell number to start with"), (["start_noise"], "start noise"),
    (["start_delay", "sdev_start_delay"], "station delay to start with"), (["r_start_eqh", "r_start_eqv"], "start EQ region"),
]


def load(name: str):
    """-> (config dict, arrays dict) of 'example' or 'example2'."""
    with open(os.path.join(GOLDEN, f"inputs_{name}.json")) as f:
        cfg = json.load(f)
    arr = dict(np.load(os.path.join(GOLDEN, f"inputs_{name}.npz")))
    return cfg, arr


def write_config(cfg: dict, path: str, **override):
    c = dict(cfg)
    c.update(override)
    with open(path, "w") as f:
        for fields, comment in CONFIG_LINES:
            if fields is None:
                f.write("1 dummy 1\t# unused\n")
                continue
            f.write(" ".join(str(c[k]) for k in fields) + f"\t# {comment}\n")


def write_picks(arr: dict, path: str):
    ev_off, n_p = arr["ev_off"], arr["n_p"]
    with open(path, "w") as f:
        for e in range(len(n_p)):
            b, end = int(ev_off[e]), int(ev_off[e + 1])
            hdr = f"# {e} {int(n_p[e])} {end - b - int(n_p[e])} {arr['reftime'][e]:.6f}"
            fx = arr["fix"][e]
            if (fx != -9999.0).any():
                hdr += f" {fx[0]:.6f} {fx[1]:.6f} {fx[2]:.6f}"
            f.write(hdr + "\n")
            for j in range(b, end):
                ph = "P" if (j - b) < n_p[e] else "S"
                f.write(f"S{int(arr['st_id'][j]):03d} {int(arr['st_id'][j]):03d} {ph} {arr['x'][j]:8.3f} {arr['y'][j]:8.3f} "
                        f"{arr['z'][j]:8.3f} {arr['t64'][j]:8.3f} {int(arr['cls'][j])}\n")


def materialise(name: str, directory: str, **cfg_override):
    cfg, arr = load(name)
    os.makedirs(directory, exist_ok=True)
    cfgp, pkp = os.path.join(directory, "config_eqx.dat"), os.path.join(directory, "picks")
    write_config(cfg, cfgp, **cfg_override)
    write_picks(arr, pkp)
    return cfgp, pkp
