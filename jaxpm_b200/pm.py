"""Force composition, LPT and Gaussian ICs — same signatures as
/root/reference/jaxpm/pm.py (pm_forces :12-58, lpt :61-126, linear_field :129-144,
pgd_correction :147-172).

`pm_forces` runs paint -> R2C -> fused Green's x gradient pass -> batched C2R x3 ->
read3 as five launches (+cuFFT) instead of the reference's ~20 XLA ops, and is
differentiable (reverse mode) through hand-written adjoints:
    F = R3(x) L P(x),  L_d = IFFT diag(i a_d / k^2) FFT,  L_d^T = -L_d
    dF^T u = sum_d u_d dR(F_d)/dx  +  dP^T [ -sum_d L_d (P(x; weight=u_d)) ]
"""
import numpy as np
import torch

from . import cosmology as jc
from . import distributed, ops
from ._lib import as_f32
from .distributed import HalfSpectrum, _single, fft3d, ifft3d, normal_field
from .growth import dGf2a, dGfa, growth_factor, growth_factor_second, growth_rate, growth_rate_second


import os as _os

_FAST_API = _os.environ.get("JPM_FAST_API", "1") != "0"   # JPM_FAST_API=0: order-preserving kernels + cuFFT everywhere
_FUSED_VJP = _os.environ.get("JPM_FUSED_VJP", "1") != "0"  # 0: the unfused adjoint (3 gathers + 3 paints + cuFFT)


def _filter_key(ft):
    return None if ft is None else (ft[0], float(ft[1]))


def _pm_forces_vjp_fused(positions, u, mesh_shape, relative, r_split, filter_tab):
    """Cotangent u[..., 3] of the forces -> gradient with respect to the positions, on the fused passes (power-of-two
    meshes): recompute the force meshes through the fused FFT chain, ONE pass over the particles for the read adjoint
    (readgrad3), ONE for the three weighted paints (paint3), the real-space divergence of those meshes
    (sum_d L_d^T G_d = -Phi sum_d D_d G_d, D_d = the difference stencil the gradient kernel is the symbol of), ONE
    transform pair on the potential chain, and the paint adjoint accumulated in place.  Before: 4 gathers, 3 paints,
    6 axpys and 5 cuFFT transforms + a k-space pass per force evaluation."""
    plan = ops.get_plan(mesh_shape, positions.device)
    rho = torch.zeros(plan.shape, dtype=torch.float32, device=positions.device)
    if relative:
        ops.cic_paint_dx_(rho, positions)
    else:
        ops.cic_paint_(rho, positions)
    f3 = ops.force_meshes_from_density_fused(rho, plan, r_split, filter_tab)
    del rho
    g = ops.cic_readgrad3(f3, positions, u, relative)
    del f3
    G = torch.zeros((3, *plan.shape), dtype=torch.float32, device=positions.device)
    ops.cic_paint3_(G, positions, u, relative)
    S = ops.fd_divergence3(G)
    del G
    psi = ops.potential_from_density_fused(S, plan, r_split, filter_tab)
    del S
    ops.cic_readgrad1_(g, psi, positions, relative, scale=-1.0)
    return g.reshape(positions.shape)


class _PMForces(torch.autograd.Function):
    """forces[...,3] from positions (absolute) or displacements (relative), single device."""

    @staticmethod
    def forward(ctx, positions, mesh_shape, relative, r_split, filter_tab):
        if _FAST_API and _FUSED_VJP and ops.fast_path_shape(mesh_shape) and positions.numel() > 0:
            # value on the tile kernels; the backward pass recomputes what it needs (nothing but the positions is kept)
            ctx.save_for_backward(positions)
            ctx.cfg = (None, relative, r_split, filter_tab)
            ctx.mesh_shape = tuple(mesh_shape)
            return ops.pm_forces_tiles(positions, mesh_shape, relative, float(r_split), filter_tab)
        plan = ops.get_plan(mesh_shape, positions.device)
        rho = torch.zeros(plan.shape, dtype=torch.float32, device=positions.device)
        if relative:
            ops.cic_paint_dx_(rho, positions)
        else:
            ops.cic_paint_(rho, positions)
        f3 = ops.force_meshes_from_density(rho, plan, r_split, filter_tab)
        out = ops.cic_read3(f3, positions, 1.0, relative)
        ctx.save_for_backward(positions, f3)
        ctx.cfg = (plan, relative, r_split, filter_tab)
        return out

    @staticmethod
    def backward(ctx, u):
        if ctx.cfg[0] is None:
            (positions,) = ctx.saved_tensors
            _, relative, r_split, filter_tab = ctx.cfg
            return _pm_forces_vjp_fused(positions, u.contiguous(), ctx.mesh_shape, relative, r_split, filter_tab), \
                None, None, None, None
        positions, f3 = ctx.saved_tensors
        plan, relative, r_split, filter_tab = ctx.cfg
        u = u.contiguous().reshape(-1, 3)
        dev = positions.device
        # (1) through the read: sum_d u_d * d read(F_d)/dx
        g = torch.zeros_like(positions)
        ucols = u.t().contiguous()  # [3, np]
        for d in range(3):
            _, gd = ops.cic_readgrad(f3[d], positions, relative, want_value=False, grad_scale=ucols[d])
            ops.axpby(1.0, g, 1.0, gd, out=g)
        # (2) through the meshes: G_d = paint(weight=u_d); grho = -sum_d L_d G_d
        G = torch.zeros((3, *plan.shape), dtype=torch.float32, device=dev)
        for d in range(3):
            if relative:
                ops.cic_paint_dx_(G[d], positions, ucols[d])
            else:
                ops.cic_paint_(G[d], positions, ucols[d])
        Gk = torch.empty((3, *plan.spec_shape), dtype=torch.complex64, device=dev)
        for d in range(3):
            Gk[d] = ops.rfft3(G[d], plan)
        grho = ops.irfft3_(ops.greens_div(Gk, plan, None, r_split, filter_tab), plan, 1)
        # (3) through the paint (weight 1): d <grho, paint(x)> / dx = d read(grho)/dx
        _, gp = ops.cic_readgrad(grho, positions, relative, want_value=False)
        ops.axpby(1.0, g, 1.0, gp, out=g)
        return g, None, None, None, None


class _ForcesFromSpectrum(torch.autograd.Function):
    """forces from a given delta_k (lpt path, pm.py:41-47): differentiable wrt the real
    field the spectrum came from is handled by `_ForcesFromField`."""

    @staticmethod
    def forward(ctx, field, positions, relative, r_split, filter_tab):
        plan = ops.get_plan(field.shape, field.device)
        if ops.fast_path_shape(plan.shape) and _FAST_API:
            f3 = ops.force_meshes_from_density_fused(field, plan, r_split, filter_tab)
        else:
            f3 = ops.force_meshes_from_density(field, plan, r_split, filter_tab)
        out = ops.cic_read3(f3, positions, 1.0, relative)
        ctx.save_for_backward(positions, f3)
        ctx.cfg = (plan, relative, r_split, filter_tab)
        return out

    @staticmethod
    def backward(ctx, u):
        positions, f3 = ctx.saved_tensors
        plan, relative, r_split, filter_tab = ctx.cfg
        u = u.contiguous().reshape(-1, 3)
        ucols = u.t().contiguous()
        dev = positions.device
        gfield = gpos = None
        if ctx.needs_input_grad[0]:
            Gk = torch.empty((3, *plan.spec_shape), dtype=torch.complex64, device=dev)
            for d in range(3):
                Gd = torch.zeros(plan.shape, dtype=torch.float32, device=dev)
                if relative:
                    ops.cic_paint_dx_(Gd, positions, ucols[d])
                else:
                    ops.cic_paint_(Gd, positions, ucols[d])
                Gk[d] = ops.rfft3(Gd, plan)
            gfield = ops.irfft3_(ops.greens_div(Gk, plan, None, r_split, filter_tab), plan, 1)
        if ctx.needs_input_grad[1]:
            gpos = torch.zeros_like(positions)
            for d in range(3):
                _, gd = ops.cic_readgrad(f3[d], positions, relative, want_value=False,
                                         grad_scale=ucols[d])
                ops.axpby(1.0, gpos, 1.0, gd, out=gpos)
        return gfield, gpos, None, None, None


def _pm_forces_f64(positions, mesh_shape, r_split, paint_absolute_pos, sharding):
    """The x64 mode of pm_forces (jax_enable_x64, the mode of the reference's distributed tests,
    tests/test_distributed_pm.py:30): float64 positions -> float64 forces.  Paint and the three reads are the double
    kernels of csrc/f64.cu; the transforms in between are library FFTs in double (torch.fft = cuFFT D2Z / Z2D) with the
    k-space kernels of kernels.py:41-115 applied as written.  A parity / reference mode - single device, no gradient
    rules; the float32 path is the product."""
    from .painting import cic_paint, cic_paint_dx, cic_read, cic_read_dx
    if not _single(sharding):
        raise NotImplementedError("float64 pm_forces is single-device")
    relative = not paint_absolute_pos
    if relative:
        mesh_shape = tuple(positions.shape[:3])
        rho = cic_paint_dx(positions)
    else:
        mesh_shape = tuple(int(n) for n in mesh_shape)
        rho = cic_paint(torch.zeros(mesh_shape, dtype=torch.float64, device=positions.device), positions)
    dk = torch.fft.rfftn(rho)
    dev = rho.device
    w = [2 * np.pi * torch.fft.fftfreq(n, dtype=torch.float64, device=dev) for n in mesh_shape[:2]]
    w.append(2 * np.pi * torch.fft.rfftfreq(mesh_shape[2], dtype=torch.float64, device=dev))
    kx, ky, kz = w[0][:, None, None], w[1][None, :, None], w[2][None, None, :]
    kk = kx**2 + ky**2 + kz**2
    pot = dk * torch.where(kk == 0, torch.zeros_like(kk), -1.0 / torch.where(kk == 0, torch.ones_like(kk), kk))
    if r_split != 0:
        pot = pot * torch.exp(-kk * r_split**2)
    out = []
    for wd in (kx, ky, kz):
        grad = 1j * (8 * torch.sin(wd) - torch.sin(2 * wd)) / 6.0
        f = torch.fft.irfftn(-grad * pot, s=mesh_shape)
        out.append(cic_read_dx(f, positions) if relative else cic_read(f, positions))
    return torch.stack(out, dim=-1)


def pm_forces(positions, mesh_shape=None, delta=None, r_split=0, paint_absolute_pos=True, halo_size=0,
              sharding=None, filter_tab=None):
    """Computes gravitational forces on particles using a PM scheme (pm.py:12-58).

    `filter_tab=(table, kmax)` is the optional radial filter slot (PGD / neural correction,
    ode.py:194-196) multiplied into pot_k inside the fused k-space pass."""
    if mesh_shape is None:
        assert (delta is not None), "If mesh_shape is not provided, delta should be provided"
        mesh_shape = getattr(delta, "mesh_shape", None) or tuple(delta.shape)
    if isinstance(positions, torch.Tensor) and positions.dtype == torch.float64 and delta is None:
        return _pm_forces_f64(positions, mesh_shape, float(r_split), paint_absolute_pos, sharding)
    positions = as_f32(positions)
    relative = not paint_absolute_pos
    if not _single(sharding):
        from . import halo
        return halo.pm_forces(positions, mesh_shape, delta, r_split, relative, halo_size, sharding,
                              filter_tab)
    if delta is None:
        if relative:
            mesh_shape = tuple(positions.shape[:3])  # pm.py:36-39: mesh comes from the displacement shape
        if (ops.fast_path_shape(mesh_shape) and not (positions.requires_grad and torch.is_grad_enabled())
                and positions.numel() > 0 and _FAST_API):
            # no gradient requested: tile-binned paint, fused FFT chain, shared-memory gather
            return ops.pm_forces_tiles(positions, mesh_shape, relative, float(r_split), filter_tab)
        return _PMForces.apply(positions, tuple(mesh_shape), relative, float(r_split), filter_tab)
    if isinstance(delta, torch.Tensor) and delta.is_complex():
        # a spectrum from fft3d: go back to the real field once (cheap, keeps one code path)
        delta = ifft3d(delta)
    return _ForcesFromSpectrum.apply(as_f32(delta), positions, relative, float(r_split), filter_tab)


def lpt(cosmo, initial_conditions, particles=None, a=0.1, halo_size=0, sharding=None, order=1):
    """First and second order LPT displacement, momentum and force (pm.py:61-126).
    Returns (dx, p, f), each [nx, ny, nz, 3]."""
    ic = as_f32(initial_conditions)
    paint_absolute_pos = particles is not None
    if particles is None:
        particles = torch.zeros((*ic.shape, 3), dtype=torch.float32, device=ic.device)
    a = float(np.atleast_1d(a)[0])
    E = float(np.sqrt(jc.Esqr(cosmo, a)))
    D1, f1 = float(growth_factor(cosmo, a)), float(growth_rate(cosmo, a))
    force1 = pm_forces(particles, delta=ic, paint_absolute_pos=paint_absolute_pos,
                       halo_size=halo_size, sharding=sharding)
    c_dx, c_p, c_f = D1, a**2 * f1 * E * D1, a**2 * E * float(dGfa(cosmo, a))
    if order == 1:
        return _scale3(force1, c_dx, c_p, c_f)
    if not _single(sharding):
        from . import halo
        delta2 = halo.lpt2_source(ic, sharding)
    else:
        plan = ops.get_plan(ic.shape, ic.device)
        delta2 = _Lpt2Source.apply(ic, plan)
    force2 = pm_forces(particles, delta=delta2, paint_absolute_pos=paint_absolute_pos,
                       halo_size=halo_size, sharding=sharding)
    D2, f2 = float(growth_factor_second(cosmo, a)), float(growth_rate_second(cosmo, a))
    c2_dx = 3 / 7 * D2
    c2_p, c2_f = a**2 * f2 * E * c2_dx, a**2 * E * float(dGf2a(cosmo, a))
    return _scale3(force1, c_dx, c_p, c_f, force2, c2_dx, c2_p, c2_f)


class _Lincomb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, x, b, y):
        ctx.ab = (a, b, y is not None)
        return ops.axpby(a, x, b, y)

    @staticmethod
    def backward(ctx, g):
        a, b, has_y = ctx.ab
        g = g.contiguous()
        return None, ops.axpby(a, g), None, (ops.axpby(b, g) if has_y else None)


def _scale3(f1, a1, b1, c1, f2=None, a2=0.0, b2=0.0, c2=0.0):
    return tuple(_Lincomb.apply(k1, f1, k2, f2) if f2 is not None else _Lincomb.apply(k1, f1, 0.0, None)
                 for k1, k2 in ((a1, a2), (b1, b2), (c1, c2)))


class _Lpt2Source(torch.autograd.Function):
    """delta2 = sum_{i<j} phi,ii phi,jj - phi,ij^2 (pm.py:88-109) from the linear field."""

    @staticmethod
    def forward(ctx, ic, plan):
        dk = ops.rfft3(ic, plan)
        if not ctx.needs_input_grad[0]:
            return ops.lpt2_source(dk, plan)
        delta2, s6 = ops.lpt2_source(dk, plan, return_shear=True)
        ctx.save_for_backward(s6)
        ctx.plan = plan
        return delta2

    @staticmethod
    def backward(ctx, g):
        # delta2 = s00 s11 + s22 (s00 + s11) - s01^2 - s02^2 - s12^2 with s_q = L_q(ic), L_q self-adjoint
        (s6,) = ctx.saved_tensors
        return ops.lpt2_source_vjp(s6, g.contiguous(), ctx.plan), None


def linear_field(mesh_shape, box_size, pk, seed, sharding=None, white_noise=None, device="cuda"):
    """Gaussian initial conditions with power spectrum `pk` (pm.py:129-144).

    `pk` is a callable k[h/Mpc] -> P(k) (NumPy); it is tabulated on a log grid and applied in
    one fused k-space pass.  `white_noise` (optional) replaces the N(0,1) draw so that parity
    runs can share one IC array."""
    mesh_shape = tuple(int(s) for s in mesh_shape)
    field = white_noise if white_noise is not None else normal_field(seed, mesh_shape, sharding,
                                                                     device=device)
    field = as_f32(field)
    if not _single(sharding):
        from . import halo
        return halo.linear_field(field, mesh_shape, box_size, pk, sharding)
    plan = ops.get_plan(mesh_shape, field.device)
    kscale = [mesh_shape[i] / box_size[i] for i in range(3)]
    kmax = np.sqrt(sum((np.pi * s)**2 for s in kscale)) * 1.001
    kmin = 2 * np.pi / max(box_size) * 0.999
    lk = np.linspace(np.log10(kmin), np.log10(kmax), 16384)
    scale = np.prod(mesh_shape) / np.prod(box_size)
    amp = np.sqrt(np.asarray(pk(10.0**lk), dtype=np.float64) * scale)
    # the k = 0 mode is multiplied by sqrt(P(0) Nc / V) like every other mode (pm.py:141-143); 0 where P(0) is not finite
    with np.errstate(all="ignore"):
        p0 = np.asarray(pk(np.zeros(1)), dtype=np.float64).reshape(-1)[0]
    dc = float(np.sqrt(p0 * scale)) if np.isfinite(p0) and p0 >= 0 else 0.0
    if ops.fast_path_shape(mesh_shape) and _FAST_API:
        # three forward passes, the amplitude table inside the x pass, three inverse passes (csrc/pmfft.cu)
        return ops.linear_field_fused(field, plan, amp.astype(np.float32), lk[0], lk[-1], kscale, dc)
    spec = ops.rfft3(field, plan)
    spec = ops.kfilter_logtab(spec, plan, amp.astype(np.float32), lk[0], lk[-1], kscale, 1.0 / plan.ncell)
    out = ops.irfft3_(spec, plan, 1)
    # the table pass gives k = 0 the first table entry: put sqrt(P(0) Nc / V) there instead (a constant shift)
    mean = field.mean()
    return out + (dc - float(amp[0])) * mean


def pgd_correction(pos, mesh_shape, params):
    """Potential-gradient-descent displacement (pm.py:147-172, intended maths — the reference's
    function uses a forward FFT where the inverse is meant).  params=[alpha,kl,ks]."""
    from .kernels import pgd_filter_table
    alpha, kl, ks = (float(p) for p in params)
    forces = pm_forces(pos, mesh_shape=mesh_shape, filter_tab=pgd_filter_table(kl, ks))
    return _Lincomb.apply(alpha, forces, 0.0, None)
