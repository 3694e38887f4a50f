"""Writers for the reference's input formats (config_eqx.dat: src/mcmc_eq.c:345-388; pick file:
src/mcmc_eq.c:1238-1294) and dict <-> MqConfig helpers.  Used by bench.py (to hand the synthetic
workload to the reference binary) and by the tests."""
from __future__ import annotations

import numpy as np

from ._lib import MqConfig, Picks

_LINES = [  # (fields, comment) per line of config_eqx.dat
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
_GRID = ("h", "nx", "ny", "nz", "x0", "y0", "z0")


def config_to_dict(c: MqConfig) -> dict:
    d = {}
    for name, _t in c._fields_:
        v = getattr(c, name)
        if name == "grid":
            for g in _GRID:
                d[g] = getattr(v, g)
        else:
            d[name] = v.decode() if isinstance(v, bytes) else v
    return {k: (float(np.format_float_positional(np.float32(v), unique=True)) if isinstance(v, float) else v)
            for k, v in d.items()}


def config_from_dict(d: dict) -> MqConfig:
    c = MqConfig()
    for k, v in d.items():
        if k in _GRID:
            setattr(c.grid, k, v)
        else:
            setattr(c, k, v.encode() if isinstance(v, str) else v)
    return c


def write_config(cfg, path: str, **override) -> None:
    d = dict(cfg) if isinstance(cfg, dict) else config_to_dict(cfg)
    d.update(override)
    with open(path, "w") as f:
        for fields, comment in _LINES:
            if fields is None:
                f.write("1 dummy 1\t# unused\n")
            else:
                f.write(" ".join(str(d[k]) for k in fields) + f"\t# {comment}\n")


def write_picks(pk, path: str, t64=None) -> None:
    """pk: Picks or a dict of arrays (ev_off, n_p, st_id, x, y, z, t/t64, cls, reftime, fix)."""
    a = pk if isinstance(pk, dict) else dict(ev_off=pk.ev_off, n_p=pk.n_p, st_id=pk.st_id, x=pk.x, y=pk.y, z=pk.z,
                                             t=pk.t, cls=pk.cls, reftime=pk.reftime, fix=pk.fix)
    t = t64 if t64 is not None else a.get("t64", a.get("t"))
    ev_off, n_p = a["ev_off"], a["n_p"]
    with open(path, "w") as f:
        for e in range(len(n_p)):
            b, end = int(ev_off[e]), int(ev_off[e + 1])
            hdr = f"# {e} {int(n_p[e])} {end - b - int(n_p[e])} {a['reftime'][e]:.6f}"
            fx = np.asarray(a["fix"]).reshape(-1, 3)[e]
            if (fx != -9999.0).any():
                hdr += f" {fx[0]:.6f} {fx[1]:.6f} {fx[2]:.6f}"
            f.write(hdr + "\n")
            for j in range(b, end):
                ph = "P" if (j - b) < n_p[e] else "S"
                f.write(f"S{int(a['st_id'][j]):03d} {int(a['st_id'][j]):03d} {ph} {a['x'][j]:8.3f} {a['y'][j]:8.3f} "
                        f"{a['z'][j]:8.3f} {t[j]:8.3f} {int(a['cls'][j])}\n")


def format_record(rec: dict, reftime, tag: str = "mod") -> str:
    """Text of one record as print_model_raw writes it (src/mcmc_eq.c:234-248; C twin: host/mq_io.c mqio_write_record).
    `rec` is a record dict of Sampler.drain() / Sampler.snapshot()."""
    code = {"mod": rec["code"] + ".", "sta": "ST", "bat": "BF"}[tag]
    n, rms = rec["noise"], float(np.float32(rec["rms"]))
    out = ["%3s %2s %8d %3d %f" % (tag, code, rec["number"], rec["dim"], rms) +
           "".join(" %f" % float(n[k]) for k in (0, 2, 4, 6, 1, 3, 5, 7)) +
           "".join(" %f %f %f" % (float(rec["z"][i]), float(rec["vp"][i]), float(rec["vpvs"][i])) for i in range(rec["dim"]))]
    for i, q in enumerate(rec["eq"]):
        out.append("EQ  %2s %8d %d %f %f %f %f %f %f" % (code, rec["number"], i, rms, float(q[0]), float(q[1]), float(q[2]),
                                                       float(reftime[i]), float(rec["origin"][i])))
    for i in range(len(rec["pres"])):
        out.append("RES %2s %8d %d %f %f %f" % (code, rec["number"], i, rms, float(rec["pres"][i]), float(rec["sres"][i])))
    return "\n".join(out) + "\n"
