"""Seeded synthetic workloads of the named shapes (BASELINE.json configs 3-5):
Example grid (h = 2 km, 200 x 200 x 62, z0 = -4), stations in a 120 km disc at elevations of 0-2 km, events in the
central 100 x 100 km at 0-60 km depth, every event picked at every station (P and S), classes uniform in 0..3,
noise scaled by class and phase the way Example/make_synthetics does (rms 0.10 * ((class+1+2.5*isS)/4)*2)."""
from __future__ import annotations

import numpy as np

from ._lib import MqConfig, Picks, Sampler
from .io import config_from_dict

# a layered crust over a mantle half space, of the same character as the reference's Example/synth_model
TRUTH_Z = np.array([-1.0, 5.0, 11.0, 17.0, 23.0, 29.0, 35.0, 41.0, 50.0, 80.0], np.float32)
TRUTH_VP = np.array([5.0, 5.2, 5.8, 6.2, 6.5, 6.8, 7.1, 7.26, 8.0, 8.0], np.float32)
TRUTH_VPVS = np.array([1.85, 1.85, 1.77, 1.73, 1.65, 1.65, 1.73, 1.73, 1.73, 1.73], np.float32)

EXAMPLE_LIKE = dict(
    h=2.0, nx=200, ny=200, nz=62, x0=-200.0, y0=-200.0, z0=-4.0, max_dim=22,
    vpmin=2.0, vpmax=12.0, vpvsmin=1.0, vpvsmax=3.0, noise_min=0.001, noise_max=10.0, residual_min=-5.0, residual_max=5.0,
    sdevx=10.0, sdevy=10.0, sdevz=5.0, sdevvp=0.05, sdevvpvs=0.02, sdevn=0.01, sdevxs=1.0, epi_search=2.0, sdevys=1.0,
    sdevzs=1.0, sdevresidual=0.02, inv_control=0.05, reference_station=1, scor_flag=0, ref_statcor_P=0.0, ref_statcor_S=0.0,
    tria=0, j_max_start=50000, j_max_main=250000, deci=2000, true_random=1, eikonal=1, dstring_start="QN",
    dstring_main="QVRPBDMN", aflag=0, inp_model_switch="VRN", start_vp=5.0, sdev_start_vp=0.5, start_vp_grad=0.03,
    start_vpvs=1.9, sdev_start_vpvs=0.2, start_cell_number=15, sdev_start_cell_number=5, start_noise=1.0,
    start_delay=0.0, sdev_start_delay=0.0, r_start_eqh=0.5, r_start_eqv=0.5,
)


# BASELINE.json configs[4]: 0.1 km grid to 200 km depth, 40 km aperture (eikonal plane 565 x 2001)
FINE_GRID = dict(h=0.1, nx=400, ny=400, nz=2001, x0=-20.0, y0=-20.0, z0=0.0, sdevxs=0.3, sdevys=0.3, sdevzs=0.5)


def config(**override) -> MqConfig:
    d = dict(EXAMPLE_LIKE)
    d.update(override)
    return config_from_dict(d)


def geometry(n_events: int, n_stations: int, seed: int = 33, aperture: float = 1.0, elev=(-2.0, 0.0), depth=(0.0, 60.0)):
    """Stations, events, classes, station corrections (before travel times exist)."""
    rng = np.random.default_rng(seed)
    r = 120.0 * aperture * np.sqrt(rng.uniform(0, 1, n_stations))
    th = rng.uniform(0, 2 * np.pi, n_stations)
    # 3 decimals, the precision of the pick-file format
    sx, sy = np.round(r * np.cos(th), 3), np.round(r * np.sin(th), 3)
    sz = np.round(rng.uniform(elev[0], elev[1], n_stations), 3)
    ev = np.stack([rng.uniform(-50, 50, n_events) * aperture, rng.uniform(-50, 50, n_events) * aperture,
                   rng.uniform(depth[0], depth[1], n_events)], axis=1).astype(np.float32)
    pcor = rng.normal(0, 0.3, n_stations); pcor -= pcor.mean()
    scor = rng.normal(0, 0.5, n_stations); scor -= scor.mean()
    cls = rng.integers(0, 4, (n_events, 2, n_stations))
    return dict(sx=sx, sy=sy, sz=sz, ev=ev, pcor=pcor.astype(np.float32), scor=scor.astype(np.float32), cls=cls, rng=rng)


def picks_from(geo, t=None) -> Picks:
    ne, ns = geo["ev"].shape[0], len(geo["sx"])
    per = 2 * ns
    ev_off = np.arange(ne + 1, dtype=np.int32) * per
    st = np.tile(np.arange(ns, dtype=np.int32), 2 * ne)
    x = np.tile(geo["sx"], 2 * ne).astype(np.float32)
    y = np.tile(geo["sy"], 2 * ne).astype(np.float32)
    z = np.tile(geo["sz"], 2 * ne).astype(np.float32)
    cls = geo["cls"].reshape(-1).astype(np.int32)
    tt = np.zeros(ne * per, np.float32) if t is None else t
    return Picks(ev_off, np.full(ne, ns, np.int32), st, x, y, z, tt, cls, np.arange(ne) * 1000.0, None, ns)


def workload(n_events: int = 200, n_stations: int = 50, seed: int = 33, device: int = 0, rms: float = 0.10, predictor=None,
             fine: bool = False, **cfg_override):
    """-> (config, picks, truth) with travel times predicted from the truth model + noise.  The prediction comes from the
    GPU library, or from `predictor(cfg, picks, truth_state) -> t_pred[n_picks]` when given (bench.py's reference arm
    passes the CPU oracle there so that nothing of this library runs in that arm)."""
    if fine:      # the fine-grid stress case: stations on the surface of a 40 km wide array, events down to 180 km
        cfg = config(**dict(FINE_GRID, **cfg_override))
        geo = geometry(n_events, n_stations, seed, aperture=1.0 / 6.0, elev=(0.0, 0.15), depth=(1.0, 180.0))
    else:
        cfg = config(**cfg_override)
        geo = geometry(n_events, n_stations, seed)
    pk0 = picks_from(geo)
    if predictor is not None:
        tpred = np.asarray(predictor(cfg, pk0, dict(z=TRUTH_Z, vp=TRUTH_VP, vpvs=TRUTH_VPVS, eq=geo["ev"], pres=geo["pcor"],
                                                   sres=geo["scor"])), np.float32)
    else:
        s = Sampler(cfg, pk0, 1, device, seed)
        m = s.new_models()
        d = len(TRUTH_Z)
        m.dim[0] = d
        m.z[0, :d], m.vp[0, :d], m.vpvs[0, :d] = TRUTH_Z, TRUTH_VP, TRUTH_VPVS
        m.eq[0] = geo["ev"]
        m.pres[0], m.sres[0] = geo["pcor"], geo["scor"]
        s.set_models(m)
        s.forward(3)
        _res, tpred = s.predictions(0)
        s.close()
    ns = n_stations
    is_s = np.tile(np.repeat([0.0, 1.0], ns), n_events)
    cls = geo["cls"].reshape(-1)
    sigma = rms * ((cls + 1 + 2.5 * is_s) / 4.0) * 2.0
    t = np.round(tpred.astype(np.float64) + geo["rng"].normal(0, 1, tpred.shape) * sigma, 3)   # pick files carry 3 decimals
    pk = picks_from(geo, t.astype(np.float32))
    truth = dict(z=TRUTH_Z, vp=TRUTH_VP, vpvs=TRUTH_VPVS, eq=geo["ev"], pres=geo["pcor"], sres=geo["scor"], t64=t,
                 tpred=np.asarray(tpred, np.float32))      # noise-free predictions the picks were made from
    return cfg, pk, truth
