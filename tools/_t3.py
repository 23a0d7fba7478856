import sys, numpy as np, torch, ctypes as C
import os; R=os.environ.get('JROOT','/root/repo'); sys.path.insert(0,R); sys.path.insert(0,R+'/tests')
import test_gpu_parity as tp
from jaxpm_b200 import ops, _lib
from jaxpm_b200.cosmology import Planck15
from jaxpm_b200.ode import nbody_kick_drift, kick_drift_coefficients
from jaxpm_b200.pm import lpt
cuda=torch.device('cuda',0)
shape, box = (64,64,64),(256.,)*3
ic = tp._ic(shape, box)
cosmo = Planck15()
gdx,gp,_ = lpt(cosmo, tp.T(ic,cuda), particles=None, a=0.1, order=1)
ref=[]
nbody_kick_drift(cosmo, gdx.clone(), gp.clone(), 0.1,1.0,10, mesh_shape=shape, paint_absolute_pos=False, resident=False,
                 callback=lambda n,p,v: ref.append((p.clone(), v.clone())))
d,k = kick_drift_coefficients(cosmo, 0.1, 1.0, 10, "symplectic")
for kw in (dict(tile=16,margin=2), dict(tile=8,margin=0)):
    pos, vel = gdx.clone(), gp.clone()
    ops.axpby(1.0, pos, d[0], vel, out=pos)
    sim = ops.Sim(shape, shape, True, cuda, **kw)
    sim.load(pos, vel)
    prev = (0,0,0,0)
    for n in range(10):
        sim.step(k[n], d[n+1] if n+1<10 else 0.0)
        sim.store(pos, vel)
        fb = sim.fallback_counts()
        ep = (pos-ref[n][0]).abs(); ev=(vel-ref[n][1]).abs()
        print(kw, n, 'pos err %.3e (n>1e-3: %d)  vel err %.3e  max|dpos per step| %.2f' % (float(ep.max()), int((ep>1e-3).sum()), float(ev.max()),
              float((d[n+1] if n+1<10 else 0.0)*vel.abs().max())), 'fallbacks this step', tuple(a-b for a,b in zip(fb,prev)), flush=True)
        prev = fb

print("---- locate")
pos, vel = gdx.clone(), gp.clone()
ops.axpby(1.0, pos, d[0], vel, out=pos)
sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=2)
sim.load(pos, vel)
for n in range(7):
    before = (pos.clone(), vel.clone())
    sim.step(k[n], d[n+1]); sim.store(pos, vel)
ev = (vel-ref[6][1]).abs().amax(-1)
bad = torch.nonzero(ev > 1e-3)
print('bad particles', bad.tolist())
grid = torch.stack(torch.meshgrid(*[torch.arange(64, device=cuda)]*3, indexing='ij'), -1).float()
for b in bad.tolist():
    i,j,kk = b
    print(' ijk', b, 'disp before', before[0][i,j,kk].tolist(), 'pp', (grid[i,j,kk]+before[0][i,j,kk]).tolist(),
          'vel before', before[1][i,j,kk].tolist(), 'vel after', vel[i,j,kk].tolist(), 'ref after', ref[6][1][i,j,kk].tolist(),
          'ref before', ref[5][1][i,j,kk].tolist())
