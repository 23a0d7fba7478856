"""Forward mode and batching, mirroring /root/reference/tests/test_distributed_pm.py:263-330
(jacfwd == jacrev of pm_forces on an 8^3 mesh) and :333-409 (vmap over stacked inputs == loop)."""
import numpy as np
import pytest
import torch

from helpers import displaced, rel_err
from oracle import painting as OP
from oracle import pm as OPM

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("relative", [False, True])
def test_paint_jvp_is_the_transpose_of_readgrad_and_matches_fd(cuda, relative):
    from jaxpm_b200 import ops
    from jaxpm_b200.transforms import cic_paint_jvp, cic_read_jvp
    shape = (12, 16, 10)
    grid, disp = displaced(shape, 1.3, dtype=np.float64)
    rng = np.random.default_rng(5)
    x0 = disp if relative else grid + disp
    v = rng.standard_normal(x0.shape)
    m = rng.standard_normal(shape)
    paint = (lambda x: OP.cic_paint_dx(x)) if relative else (lambda x: OP.cic_paint(np.zeros(shape), x))
    eps = 1e-6
    fd = (paint(x0 + eps * v) - paint(x0 - eps * v)) / (2 * eps)
    got = cic_paint_jvp(T(x0.astype(np.float32), cuda), T(v.astype(np.float32), cuda), shape, relative).cpu().numpy()
    assert rel_err(got, fd) < 1e-4
    # <paint_jvp(v), m> == <v, readgrad(m)>
    _, g = ops.cic_readgrad(T(m.astype(np.float32), cuda), T(x0.astype(np.float32), cuda), relative, want_value=False)
    lhs = float((got.astype(np.float64) * m).sum())
    rhs = float((g.cpu().numpy().astype(np.float64) * v).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), abs(rhs))
    # read JVP against finite differences of the oracle read
    read = (lambda mm, x: OP.cic_read_dx(mm, x)) if relative else (lambda mm, x: OP.cic_read(mm, x))
    dm = rng.standard_normal(shape)
    fdr = (read(m + eps * dm, x0 + eps * v) - read(m - eps * dm, x0 - eps * v)) / (2 * eps)
    _, dr = cic_read_jvp(T(m.astype(np.float32), cuda), T(x0.astype(np.float32), cuda), T(dm.astype(np.float32), cuda),
                         T(v.astype(np.float32), cuda), relative)
    assert rel_err(dr.cpu().numpy(), fdr) < 1e-4


@pytest.mark.parametrize("relative", [False, True])
def test_pm_forces_jacfwd_equals_jacrev(cuda, relative):
    """Columns of the Jacobian from the forward-mode rule, rows from the reverse-mode rule (hand-written adjoint
    kernels), on the reference's own 8^3 case; and the JVP against central differences of the float64 oracle."""
    from jaxpm_b200.pm import pm_forces
    from jaxpm_b200.transforms import pm_forces_jvp
    shape = (8, 8, 8)
    grid, disp = displaced(shape, 0.6, dtype=np.float64)
    x0 = disp if relative else grid + disp
    n = x0.size
    rng = np.random.default_rng(8)
    cols = rng.choice(n, 24, replace=False)
    rows = rng.choice(n, 24, replace=False)
    xt = T(x0.astype(np.float32), cuda)
    Jf = np.zeros((n, len(cols)))
    for j, c in enumerate(cols):
        e = torch.zeros(n, device=cuda)
        e[c] = 1.0
        F, dF = pm_forces_jvp(xt, e.reshape(x0.shape), mesh_shape=shape, paint_absolute_pos=not relative)
        Jf[:, j] = dF.reshape(-1).cpu().numpy()
    Jr = np.zeros((len(rows), n))
    for i, r in enumerate(rows):
        x = xt.clone().requires_grad_(True)
        Fx = pm_forces(x, mesh_shape=shape, paint_absolute_pos=not relative)
        u = torch.zeros(n, device=cuda)
        u[r] = 1.0
        Fx.backward(u.reshape(x0.shape))
        Jr[i] = x.grad.reshape(-1).cpu().numpy()
    a, b = Jf[rows], Jr[:, cols]
    assert np.abs(a - b).max() < 1e-4 * max(np.abs(Jf).max(), np.abs(Jr).max())
    fwd = lambda xx: OPM.pm_forces(xx, mesh_shape=shape, paint_absolute_pos=not relative)
    assert rel_err(F.cpu().numpy(), fwd(x0)) < 1e-5
    v = rng.standard_normal(x0.shape)
    eps = 1e-5
    fd = (fwd(x0 + eps * v) - fwd(x0 - eps * v)) / (2 * eps)
    _, dF = pm_forces_jvp(xt, T(v.astype(np.float32), cuda), mesh_shape=shape, paint_absolute_pos=not relative)
    assert rel_err(dF.cpu().numpy(), fd) < 2e-3


@pytest.mark.parametrize("shape", [(16, 16, 16), (12, 10, 8)])
def test_batched_equals_loop(cuda, shape):
    """vmap semantics: a batched call over a leading axis == the stack of single calls (bit for bit: same kernels)."""
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.painting import cic_paint_dx, cic_read_dx
    from jaxpm_b200.pm import lpt, pm_forces
    from jaxpm_b200.transforms import cic_paint_dx_batched, cic_read_dx_batched, lpt_batched, pm_forces_batched
    rng = np.random.default_rng(2)
    B = 3
    disp = T((1.2 * rng.standard_normal((B, *shape, 3))).astype(np.float32), cuda)
    meshes = T(rng.standard_normal((B, *shape)).astype(np.float32), cuda)
    got = pm_forces_batched(disp, paint_absolute_pos=False)
    for b in range(B):
        ref = pm_forces(disp[b], mesh_shape=shape, paint_absolute_pos=False)
        assert got[b].shape == ref.shape and float((got[b] - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
    pb = cic_paint_dx_batched(disp)
    rb = cic_read_dx_batched(meshes, disp)
    for b in range(B):
        assert torch.equal(rb[b], cic_read_dx(meshes[b], disp[b]))
        assert float((pb[b] - cic_paint_dx(disp[b])).abs().max()) < 1e-5 * float(pb[b].abs().max())
    ics = T((0.1 * rng.standard_normal((B, *shape))).astype(np.float32), cuda)
    lb = lpt_batched(Planck15(), ics, a=0.1, order=2)
    for b in range(B):
        for got_f, ref_f in zip(lb, lpt(Planck15(), ics[b], a=0.1, order=2)):
            assert torch.equal(got_f[b], ref_f)
