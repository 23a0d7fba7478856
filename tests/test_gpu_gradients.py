"""Reverse mode through the WHOLE chain of BASELINE.json's config 5: initial conditions -> lpt
(1st / 2nd order) -> K drift-kick PM steps -> cic_paint_dx -> power_spectrum -> scalar.

Mirrors /root/reference/tests/test_gradients.py:30-80 (jax.grad of a scalar of the final field with respect to
`initial_conditions`, lpt order 1 and 2, a fixed-step solver with a checkpointing adjoint).  The CUDA gradient -
hand-written adjoint kernels (readgrad, weighted paint, transposed k-space passes, the 2LPT source VJP, the P(k)
adjoint) under a per-step-recompute driver - is checked against central differences of the float64 oracle run of
the same chain."""
import numpy as np
import pytest
import torch

from helpers import gaussian_ic
from oracle import cosmology as OC
from oracle import ode as OO
from oracle import painting as OP
from oracle import pm as OPM
from oracle import utils as OU

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


def _ic(shape, box, seed=0):
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    c = Planck15()
    return gaussian_ic(shape, box, lambda k: linear_matter_power(c, k), seed=seed)


@pytest.mark.parametrize("order", [1, 2])
def test_chained_gradient_of_final_pk_wrt_initial_conditions(cuda, order):
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import nbody_kick_drift_grad
    from jaxpm_b200.painting import cic_paint_dx
    from jaxpm_b200.pm import lpt
    from jaxpm_b200.utils import power_spectrum
    shape, box, K, a0, a1 = (16, 16, 16), (64., 64., 64.), 3, 0.1, 0.6
    ic0 = _ic(shape, box).astype(np.float64)
    cosmo, ocos = Planck15(), OC.Planck15()
    rng = np.random.default_rng(7)

    def oracle_pk(ic):
        dx, p, _ = OPM.lpt(ocos, ic, a=a0, order=order)
        drift, kick = OO.symplectic_ode(shape, ocos, paint_absolute_pos=False)
        pos, _ = OO.semi_implicit_euler(drift, kick, dx, p, a0, a1, K)
        return OU.power_spectrum(OP.cic_paint_dx(pos), box_shape=box, x64=False)[1]

    pk0 = oracle_pk(ic0)
    gw = rng.standard_normal(pk0.shape) / pk0          # weights of the scalar: sum_b gw_b P(k_b), every bin O(1)
    ic = torch.tensor(ic0.astype(np.float32), device=cuda, requires_grad=True)
    dx, p, _ = lpt(cosmo, ic, a=a0, order=order)
    pos, vel = nbody_kick_drift_grad(cosmo, dx, p, a0, a1, K, paint_absolute_pos=False)
    _, pk = power_spectrum(cic_paint_dx(pos), box_shape=box)
    assert np.abs(pk.detach().cpu().numpy() / pk0 - 1).max() < 1e-4          # the forward chain itself
    (pk * T(gw.astype(np.float32), cuda)).sum().backward()
    grad = ic.grad.cpu().numpy().astype(np.float64)
    assert np.isfinite(grad).all() and np.abs(grad).max() > 0
    loss = lambda x: float((oracle_pk(x) * gw).sum())
    for _ in range(3):
        v = rng.standard_normal(shape)
        eps = 1e-4 * ic0.std()
        fd = (loss(ic0 + eps * v) - loss(ic0 - eps * v)) / (2 * eps)
        an = float((grad * v).sum())
        assert abs(fd - an) < 5e-3 * max(abs(fd), abs(an)), (order, fd, an)


def test_recompute_driver_equals_plain_autograd(cuda):
    """The per-step-recompute driver gives the gradients of the plain autograd graph through make_ode_fn-style
    steps (drift, then kick with the differentiable pm_forces)."""
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import kick_drift_coefficients, nbody_kick_drift_grad
    from jaxpm_b200.pm import _Lincomb, pm_forces
    shape, K = (16, 16, 16), 2
    rng = np.random.default_rng(3)
    disp0 = (0.7 * rng.standard_normal((*shape, 3))).astype(np.float32)
    vel0 = (0.1 * rng.standard_normal((*shape, 3))).astype(np.float32)
    u = T(rng.standard_normal((*shape, 3)).astype(np.float32), cuda)
    cosmo = Planck15()
    grads = []
    for driver in ("recompute", "plain"):
        x = T(disp0, cuda).requires_grad_(True)
        v = T(vel0, cuda).requires_grad_(True)
        if driver == "recompute":
            p, w = nbody_kick_drift_grad(cosmo, x, v, 0.2, 0.5, K, paint_absolute_pos=False)
        else:
            d, k = kick_drift_coefficients(cosmo, 0.2, 0.5, K, "symplectic")
            p, w = x, v
            for n in range(K):
                p = _Lincomb.apply(1.0, p, float(d[n]), w)
                w = _Lincomb.apply(1.0, w, float(k[n]), pm_forces(p, mesh_shape=shape, paint_absolute_pos=False))
        ((p * u).sum() + (w * u).sum()).backward()
        grads.append((x.grad.cpu().numpy(), v.grad.cpu().numpy()))
    for a, b in zip(grads[0], grads[1]):
        assert np.abs(a - b).max() < 1e-5 * np.abs(b).max()


@pytest.mark.parametrize("relative", [True, False])
def test_fused_adjoint_passes(cuda, relative):
    """The fused reverse-mode passes of pm_forces (readgrad3, paint3, the real-space divergence + one transform pair on
    the potential chain) against the unfused adjoint (three gathers, three paints, cuFFT + the transposed k-space pass)
    and their building blocks against plain compositions."""
    from jaxpm_b200 import ops
    from jaxpm_b200 import pm as jpm_pm
    shape = (32, 16, 64)
    rng = np.random.default_rng(5)
    grid = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).astype(np.float32)
    disp = (1.5 * rng.standard_normal((*shape, 3))).astype(np.float32)
    x = T(disp if relative else grid + disp, cuda)
    u = T(rng.standard_normal((*shape, 3)).astype(np.float32), cuda)
    m3 = T(rng.standard_normal((3, *shape)).astype(np.float32), cuda)
    # readgrad3 == sum of three scaled readgrads; accumulate form of the one-mesh variant
    ref = torch.zeros_like(x).reshape(-1, 3)
    for d in range(3):
        _, gd = ops.cic_readgrad(m3[d], x, relative, want_value=False, grad_scale=u[..., d].contiguous().reshape(-1))
        ref += gd.reshape(-1, 3)
    got = ops.cic_readgrad3(m3, x, u, relative)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5
    _, g1 = ops.cic_readgrad(m3[0], x, relative, want_value=False)
    acc = got.clone()
    ops.cic_readgrad1_(acc, m3[0], x, relative, scale=-2.0)
    assert float((acc - (got - 2.0 * g1.reshape(-1, 3))).abs().max() / ref.abs().max()) < 1e-5
    # paint3 == three weighted paints
    G = torch.zeros((3, *shape), device=cuda)
    ops.cic_paint3_(G, x, u, relative)
    for d in range(3):
        one = torch.zeros(shape, device=cuda)
        w = u[..., d].contiguous()
        (ops.cic_paint_dx_ if relative else ops.cic_paint_)(one, x, w)
        assert float((G[d] - one).abs().max() / one.abs().max()) < 1e-5
    # divergence == the 4th-order central differences, periodic
    D = lambda f, ax: (8 * (torch.roll(f, -1, ax) - torch.roll(f, 1, ax)) - (torch.roll(f, -2, ax) - torch.roll(f, 2, ax))) / 12
    ref_div = D(m3[0], 0) + D(m3[1], 1) + D(m3[2], 2)
    assert float((ops.fd_divergence3(m3) - ref_div).abs().max() / ref_div.abs().max()) < 1e-5
    # the whole vector-Jacobian product, with and without the long-range split
    for r_split in (0.0, 1.3):
        outs = {}
        for fused in (True, False):
            jpm_pm._FUSED_VJP = fused
            try:
                xx = x.clone().requires_grad_(True)
                F = jpm_pm.pm_forces(xx, mesh_shape=shape, paint_absolute_pos=not relative, r_split=r_split)
                (g,) = torch.autograd.grad(F, xx, u)
                outs[fused] = (F.detach(), g)
            finally:
                jpm_pm._FUSED_VJP = True
        assert float((outs[True][0] - outs[False][0]).abs().max() / outs[False][0].abs().max()) < 1e-5
        err = float((outs[True][1] - outs[False][1]).abs().max() / outs[False][1].abs().max())
        assert err < 2e-5, (r_split, err)
