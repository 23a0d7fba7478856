"""Step-by-step comparison of the fused pencil stepper (P ranks on one device) with the single-GPU stepper."""
import os, sys
import numpy as np, torch
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
from helpers import displaced
from jaxpm_b200 import ops
from jaxpm_b200.cosmology import Planck15
from jaxpm_b200.ode import kick_drift_coefficients
from jaxpm_b200.slab import SlabPlan, SlabStepper

cuda = torch.device("cuda", 0)
shape, pdims, gx, gy, tile = (32, 32, 32), (2, 2), 8, 8, 8
if len(sys.argv) > 1:
    pdims = tuple(int(v) for v in sys.argv[1].split("x"))
clip, K = float(os.environ.get("JPM_TEST_CLIP", 2.0)), int(os.environ.get("JPM_TEST_K", 7))
T = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(cuda)
_, disp = displaced(shape, 1.0)
disp = np.clip(disp, -clip, clip).astype(np.float32)
vel = (0.2 * np.random.default_rng(9).standard_normal(disp.shape)).astype(np.float32)
cosmo = Planck15()
d, k = kick_drift_coefficients(cosmo, 0.5, 0.8, K, "symplectic")
px, py = pdims
P = px * py
Lx, Ly = shape[0] // px, shape[1] // py
blocks = lambda a: [a[rx * Lx:(rx + 1) * Lx, ry * Ly:(ry + 1) * Ly] for rx in range(px) for ry in range(py)]
join = lambda l: torch.cat([torch.cat(l[rx * py:(rx + 1) * py], dim=1) for rx in range(px)], dim=0)
plans = [SlabPlan(shape, P, r, gx, cuda, pdims=pdims, gy=gy) for r in range(P)]
for p in plans:
    p.attach_local(plans)
streams = [torch.cuda.Stream(cuda) for _ in range(P)]
dl, vl = [T(b) for b in blocks(disp)], [T(b) for b in blocks(vel)]
rd, rv = T(disp), T(vel)
ops.axpby(1.0, rd, d[0], rv, out=rd)
for r in range(P):
    ops.axpby(1.0, dl[r], d[0], vl[r], out=dl[r])
ref = ops.Sim(shape, shape, True, cuda, tile=tile, margin=1)
ref.load(rd, rv)
torch.cuda.synchronize()
st = []
for r in range(P):
    with torch.cuda.stream(streams[r]):
        st.append(SlabStepper(dl[r], vl[r], gx, P, r, tile=tile, margin=1, plan=plans[r], pdims=pdims, gy=gy))
gridf = [torch.arange(m, device=cuda, dtype=torch.float32) for m in shape]
def near_edges(state, tag):
    for ax in range(3):
        view = [1, 1, 1]; view[ax] = -1
        x = gridf[ax].view(view) + state[..., ax]
        for edge in (0.0, -1.0, float(shape[ax]), float(shape[ax]) - 1.0):
            m = ((x - edge).abs() < 3e-5).nonzero()
            for row in m[:5].cpu().numpy().tolist():
                xv = float(x[tuple(row)])
                print(f"   {tag}: site {row} axis {ax} coordinate {xv!r} (edge {edge}, offset {xv - edge:.3e}) disp {state[tuple(row)].cpu().numpy()}")
for n in range(K):
    kk, dd = k[n], (d[n + 1] if n + 1 < K else 0.0)
    near_edges(rd, f"before step {n}")
    ref.step(kk, dd)
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st[r].step(kk, dd)
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st[r].store(dl[r], vl[r])
    torch.cuda.synchronize()
    ref.store(rd, rv)
    torch.cuda.synchronize()
    p, v = join(dl), join(vl)
    ev = (v - rv).abs().max(-1).values
    ep = (p - rd).abs().max(-1).values
    i = int(ev.argmax())
    ijk = np.unravel_index(i, shape)
    nbad = int((ev > 1e-4 * float(rv.abs().max())).sum())
    bad = (ev > 1e-4 * float(rv.abs().max())).nonzero()[:6].cpu().numpy().tolist()
    print(f"step {n}: max dvel {float(ev.max()):.3e} (|v|max {float(rv.abs().max()):.2f}) max dpos {float(ep.max()):.3e}; {nbad} bad; worst site {ijk} "
          f"disp {rd.reshape(-1, 3)[i].cpu().numpy()} pos {np.array(ijk) + rd.reshape(-1, 3)[i].cpu().numpy()}; ghost {plans[0].ghost_width()}; bad sites {bad}")
