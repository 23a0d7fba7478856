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
grid = torch.stack(torch.meshgrid(*[torch.arange(64, device=cuda)]*3, indexing='ij'), -1).float()
for rep in range(int(os.environ.get("REPS","3"))):
    junk = torch.full((1<<26,), float('nan'), device=cuda); del junk    # poison the allocator's free blocks
    pos, vel = gdx.clone(), gp.clone()
    ops.axpby(1.0, pos, d[0], vel, out=pos)
    sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=2)
    sim.load(pos, vel)
    for n in range(10):
        before=(pos.clone(), vel.clone())
        sim.step(k[n], d[n+1] if n+1<10 else 0.0)
        pos.fill_(float('nan')); vel.fill_(float('nan'))
        sim.store(pos, vel)
        lost = torch.isnan(pos).any(-1) | torch.isnan(vel).any(-1)
        ev = (vel-ref[n][1]).abs().amax(-1)
        bad = torch.nonzero((ev > 1e-3) | lost)
        print(rep, n, 'lost', int(lost.sum()), 'bad', len(bad), flush=True)
        for b in bad.tolist()[:6]:
            i,j,kk = b
            print('   ijk', b, 'pp before', [round(x,4) for x in (grid[i,j,kk]+before[0][i,j,kk]).tolist()],
                  'vel before', [round(x,4) for x in before[1][i,j,kk].tolist()], 'after', [round(x,4) for x in vel[i,j,kk].tolist()],
                  'ref', [round(x,4) for x in ref[n][1][i,j,kk].tolist()])
        if len(bad):
            plan = ops.get_plan(shape, cuda)
            rho = torch.zeros(shape, device=cuda); ops.cic_paint_dx_(rho, before[0])
            def padded(which):
                dims=(C.c_int32*3)()
                _lib.call("jpm_plan_padded_get_f32", plan.handle, _lib.stream(), which, None, dims)
                out=torch.empty(tuple(dims), device=cuda)
                _lib.call("jpm_plan_padded_get_f32", plan.handle, _lib.stream(), which, _lib.ptr(out), dims)
                return out
            G=4
            def fold(a):
                a=a.clone()
                for ax in range(3):
                    nn=a.shape[ax]-2*G
                    lo=a.narrow(ax,0,G); hi=a.narrow(ax,nn+G,G)
                    a.narrow(ax,nn,G).add_(lo); a.narrow(ax,G,G).add_(hi)
                    a=a.narrow(ax,G,nn)
                return a
            pd = padded(0)
            dd = fold(pd); e=(dd-rho).abs()
            print('   max density', float(rho.max()), 'paint err', float(e.max()), 'at', np.unravel_index(int(e.argmax()), shape), 'n>1e-3', int((e>1e-3).sum()))
            idx = torch.nonzero(e>1e-3)
            for t in idx.tolist()[:10]:
                print('     cell', t, 'got', float(dd[tuple(t)]), 'want', float(rho[tuple(t)]))
            f3 = ops.force_meshes_from_density(rho, plan)
            for c in range(3):
                f = padded(1+c)[G:-G,G:-G,G:-G]
                ef=(f-f3[c]).abs()
                print('   force', c, 'rel err vs direct', float(ef.max()/f3[c].abs().max()), 'at', np.unravel_index(int(ef.argmax()), shape))
            f3b = ops.force_meshes_from_density(dd.contiguous(), plan)
            for c in range(3):
                f = padded(1+c)[G:-G,G:-G,G:-G]
                print('   force', c, 'rel err vs forces-from-sim-density', float((f-f3b[c]).abs().max()/f3b[c].abs().max()))
            # ghost cells of the force meshes == periodic images of the interior?
            for c in range(3):
                pf = padded(1+c)
                inner = pf[G:-G,G:-G,G:-G]
                img = torch.nn.functional.pad(inner[None,None], (G,G,G,G,G,G), mode='circular')[0,0]
                eg = (pf-img).abs()
                print('   force', c, 'ghost-image err', float(eg.max()), 'at', np.unravel_index(int(eg.argmax()), tuple(pf.shape)))
            # expected kick of the first bad particle from the global force meshes
            i,j,kk = bad.tolist()[0]
            pp = (grid[i,j,kk]+before[0][i,j,kk]).double().cpu().numpy()
            i0 = np.floor(pp).astype(int); fr = pp-i0
            F = np.zeros(3)
            for c in range(3):
                fm = padded(1+c)[G:-G,G:-G,G:-G].double().cpu().numpy()
                for a in (0,1):
                    for b in (0,1):
                        for dd_ in (0,1):
                            w = (fr[0] if a else 1-fr[0])*(fr[1] if b else 1-fr[1])*(fr[2] if dd_ else 1-fr[2])
                            F[c] += w*fm[(i0[0]+a)%64,(i0[1]+b)%64,(i0[2]+dd_)%64]
            print('   expected dv', (k[n]*F).tolist(), 'sim dv', (vel[i,j,kk]-before[1][i,j,kk]).tolist(), 'ref dv', (ref[n][1][i,j,kk]-before[1][i,j,kk]).tolist())
            break
