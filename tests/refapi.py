"""ctypes mirror of the reference's structs (src/mc.h:61-134) so that the UNMODIFIED reference,
compiled into oracle/_ref/libmcmceq_ref.so, can be driven from Python.  Test infrastructure:
used by tools/make_golden.py (in the build container) and by the pin tests when _ref exists."""
from __future__ import annotations

import ctypes as C

import numpy as np

MD, MAX_OBS, MAX_STAT, MAX_NOQ = 1000, 1000, 1000, 3500
fp = C.POINTER(C.c_float)


class QUAKE(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Model(C.Structure):
    _fields_ = [("number", C.c_long), ("dimension", C.c_long), ("noq", C.c_long), ("nos", C.c_long),
                ("pres", C.c_float * MAX_STAT), ("sres", C.c_float * MAX_STAT), ("origin", C.c_float * MAX_NOQ),
                ("p_noise0", C.c_float), ("s_noise0", C.c_float), ("p_noise1", C.c_float), ("s_noise1", C.c_float),
                ("p_noise2", C.c_float), ("s_noise2", C.c_float), ("p_noise3", C.c_float), ("s_noise3", C.c_float),
                ("z", C.c_float * MD), ("vp", C.c_float * MD), ("vpvs", C.c_float * MD), ("eq", QUAKE * MAX_NOQ)]


class GRDHEAD(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("h", C.c_float),
                ("x0", C.c_float), ("y0", C.c_float), ("z0", C.c_float)]


class OBS(C.Structure):
    _fields_ = [("st_id", C.c_int), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("t", C.c_float),
                ("cl", C.c_int), ("layer", C.c_int), ("w1", C.c_float), ("w2", C.c_float)]


class DATA(C.Structure):
    _fields_ = [("eq_id", C.c_int), ("reftime", C.c_double), ("xfix", C.c_double), ("yfix", C.c_double),
                ("zfix", C.c_double), ("nobs_p", C.c_int), ("nobs_s", C.c_int),
                ("nobs_p0", C.c_int), ("nobs_s0", C.c_int), ("nobs_p1", C.c_int), ("nobs_s1", C.c_int),
                ("nobs_p2", C.c_int), ("nobs_s2", C.c_int), ("nobs_p3", C.c_int), ("nobs_s3", C.c_int),
                ("p_picks", OBS * MAX_OBS), ("s_picks", OBS * MAX_OBS)]


class RefForward:
    """cal_fit_newx / setup_table_new of the compiled reference on one pick file."""

    def __init__(self, ref: C.CDLL, grid: dict, picks_path: str):
        self.ref = ref
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        self.gh = GRDHEAD(grid["nx"], grid["ny"], grid["nz"], grid["h"], grid["x0"], grid["y0"], grid["z0"])
        self.data = (DATA * MAX_NOQ)()
        f = libc.fopen(picks_path.encode(), b"r")
        assert f, picks_path
        ref.read_mcmcdata.argtypes = [C.c_void_p, C.POINTER(DATA)]
        self.ne = ref.read_mcmcdata(f, self.data)
        libc.fclose(f)
        gh = self.gh
        for i in range(self.ne):   # receiver weights exactly as src/mcmc_eq.c:503-517 (float arithmetic)
            d = self.data[i]
            for picks, n in ((d.p_picks, d.nobs_p), (d.s_picks, d.nobs_s)):
                for j in range(n):
                    o = picks[j]
                    h, z0, z = np.float32(gh.h), np.float32(gh.z0), np.float32(o.z)
                    o.layer = int(np.float32(np.float32(z - z0) / h))
                    o.w2 = float(np.float32(-np.float32(np.float32(np.float32(np.float32(o.layer) * h) + z0) - z) / h))
                    o.w1 = float(np.float32(1.0 - float(np.float32(o.w2))))
        self.nxmod = int(np.sqrt(gh.nx * gh.nx + gh.ny * gh.ny))
        t3 = C.POINTER(C.POINTER(fp))
        ref.make_3d_array.restype = t3
        ref.make_3d_array.argtypes = [C.c_int] * 3
        self.tttp = ref.make_3d_array(gh.nz, gh.nz, self.nxmod)
        self.ttts = ref.make_3d_array(gh.nz, gh.nz, self.nxmod)
        ref.cal_fit_newx.restype = C.c_float
        ref.cal_fit_newx.argtypes = [C.POINTER(Model), C.POINTER(DATA), C.c_int, t3, t3, GRDHEAD, C.c_int] + [fp] * 8 + [C.c_int] * 3
        C.c_int.in_dll(ref, "TRIA").value = 0
        C.c_int.in_dll(ref, "aflag").value = 0
        self.nos = 1 + max(max([d.p_picks[j].st_id for j in range(d.nobs_p)] + [d.s_picks[j].st_id for j in range(d.nobs_s)])
                           for d in self.data[: self.ne])

    def forward(self, z, vp, vpvs, eq, pres, sres, calct=3, eikonal=1, tria=0):
        C.c_int.in_dll(self.ref, "TRIA").value = tria
        m = Model()
        m.dimension, m.noq, m.nos = len(z), self.ne, self.nos
        for i in range(len(z)):
            m.z[i], m.vp[i], m.vpvs[i] = float(z[i]), float(vp[i]), float(vpvs[i])
        for i in range(self.ne):
            m.eq[i].x, m.eq[i].y, m.eq[i].z = (float(v) for v in eq[i])
        for i in range(MAX_STAT):
            m.pres[i] = m.sres[i] = -99999.0
        for i in range(self.nos):
            m.pres[i], m.sres[i] = float(pres[i]), float(sres[i])
        mf = [C.c_float(0) for _ in range(8)]
        self.ref.cal_fit_newx(C.byref(m), self.data, self.ne, self.tttp, self.ttts, self.gh, calct,
                              *[C.byref(v) for v in mf], 0, eikonal, 0)
        # reference order of the out-parameters: mfp0,mfs0,mfp1,mfs1,... == index 2*class+phase
        return np.array([v.value for v in mf], np.float32), np.array(m.origin[: self.ne], np.float32)

    def table(self, phase):
        """ttt[j][iz][i] of the last forward, phase 1 = P, 2 = S."""
        nz = self.gh.nz
        t = np.zeros((nz, nz, self.nxmod), np.float32)
        src = self.tttp if phase == 1 else self.ttts
        for j in range(nz):
            for k in range(nz):
                t[j, k] = np.ctypeslib.as_array(src[j][k], shape=(self.nxmod,))
        return t
