import sys, time, numpy as np
sys.path.insert(0,'.')
from tests import util
import mcmc_eq_b200 as mq
rng=np.random.default_rng(0)
for gname,g in (('ex2',util.EXAMPLE2_GRID),('ex',util.EXAMPLE_GRID)):
    nx=util.nxmod_of(g); nz=g['nz']
    slows=[]; izs=[]
    for m in range(6):
        kind=['posterior','gradient','lvz','contrast'][m%4]
        z,vp,vpvs=util.voronoi_model(rng, int(rng.integers(1,21)), g['z0'], g['z0']+(nz-1)*g['h'], kind)
        s=util.rasterise_np(z,vp,vpvs,g['h'],g['z0'],nz,1+m%2)
        for iz in range(nz): slows.append(s); izs.append(iz)
    slows=np.array(slows); izs=np.array(izs)
    t0=time.time(); t=mq.eikonal_batch(slows,izs,nx); dt=time.time()-t0
    worst=0; bad=0
    for i in range(len(izs)):
        tr,rc=util.oracle_time_2d(slows[i],nx,izs[i])
        d=np.abs(t[i]-tr); worst=max(worst,d.max()); bad+=int((d>util.eikonal_tol(tr)).any())
    print(gname,'solves',len(izs),'gpu call s',round(dt,3),'max|dT|',worst,'bad',bad, flush=True)
# throughput probe
g=util.EXAMPLE_GRID; nx=util.nxmod_of(g); nz=g['nz']
slows=[]; izs=[]
for m in range(128):
    z,vp,vpvs=util.voronoi_model(rng, int(rng.integers(9,21)), g['z0'], g['z0']+(nz-1)*g['h'], 'posterior')
    s=util.rasterise_np(z,vp,vpvs,g['h'],g['z0'],nz,1)
    for iz in range(nz): slows.append(s); izs.append(iz)
slows=np.array(slows); izs=np.array(izs)
order=np.argsort(izs,kind='stable'); slows=slows[order]; izs=izs[order]
for rep in range(2):
    t0=time.time(); t=mq.eikonal_batch(slows,izs,nx); dt=time.time()-t0
    print('throughput probe: solves',len(izs),'wall',round(dt,3),'s (incl. H2D/D2H of full fields)')
print('launches',mq.lib().mq_launch_count())
