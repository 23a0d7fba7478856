"""ODE right-hand sides with the signatures of /root/reference/jaxpm/ode.py
(symplectic_fpm_ode :13-82, symplectic_ode :85-119, make_ode_fn :122-147,
make_diffrax_ode :150-176, make_neural_ode_fn :179-215) and the fused fixed-step
drivers that replace diffrax for constant-step runs.

The RHS factories return differentiable tensors through `pm_forces`.  The
drivers (`nbody_kick_drift`, `nbody_leapfrog_midpoint`) keep the particle state
resident and run ONE fused read3+kick+drift kernel per force evaluation
(jpm_pm_step_f32 / jpm_cic_read3_kick_drift_f32).
"""
import numpy as np
import torch

from . import cosmology as jc
from . import ops
from ._lib import as_f32
from .distributed import _single
from .growth import E, Gf, dGfa, gp
from .growth import growth_factor as Gp
from .pm import _Lincomb, pm_forces


def _E(cosmo, a):
    return float(np.sqrt(jc.Esqr(cosmo, float(a))))


def fpm_factors(cosmo, a, dt0):
    """(drift, kick, first_kick) scalar coefficients of symplectic_fpm_ode (ode.py:19-80),
    each per unit dt0 exactly as the reference's closures return them."""
    a = float(a)
    t0, t1, t2 = a, a + dt0, a + 2 * dt0
    ac = (t0 * t1)**0.5
    drift_contr = (Gp(cosmo, t1) - Gp(cosmo, t0)) / gp(cosmo, ac)
    drift = 1 / (ac**3 * E(cosmo, ac)) * (drift_contr / dt0)
    t0t1, t1t2 = (t0 * t1)**0.5, (t1 * t2)**0.5
    k1 = (Gf(cosmo, t1) - Gf(cosmo, t0t1)) / dGfa(cosmo, t1)
    k2 = (Gf(cosmo, t1t2) - Gf(cosmo, t1)) / dGfa(cosmo, t1)
    kick = 1.0 / (t1**2 * E(cosmo, t1)) * ((k1 + k2) / dt0)
    fk = (Gf(cosmo, t0t1) - Gf(cosmo, t0)) / dGfa(cosmo, t0)
    first_kick = 1.0 / (a**2 * E(cosmo, a)) * (fk / dt0)
    return float(drift), float(kick), float(first_kick)


def symplectic_fpm_ode(mesh_shape, dt0, cosmo, paint_absolute_pos=True, halo_size=0, sharding=None):
    def F(pos):
        return pm_forces(pos, mesh_shape=mesh_shape, paint_absolute_pos=paint_absolute_pos,
                         halo_size=halo_size, sharding=sharding)

    def drift(a, vel, args):
        return _Lincomb.apply(fpm_factors(cosmo, a, dt0)[0], as_f32(vel), 0.0, None)

    def kick(a, pos, args):
        return _Lincomb.apply(fpm_factors(cosmo, a, dt0)[1] * 1.5 * cosmo.Omega_m, F(pos), 0.0, None)

    def first_kick(a, pos, args):
        c = args if args is not None else cosmo
        return _Lincomb.apply(fpm_factors(c, a, dt0)[2] * 1.5 * c.Omega_m, F(pos), 0.0, None)

    return drift, kick, first_kick


def symplectic_ode(mesh_shape, cosmo, paint_absolute_pos=True, halo_size=0, sharding=None):
    def drift(a, vel, args):
        return _Lincomb.apply(1 / (float(a)**3 * _E(cosmo, a)), as_f32(vel), 0.0, None)

    def kick(a, pos, args):
        forces = pm_forces(pos, mesh_shape=mesh_shape, paint_absolute_pos=paint_absolute_pos,
                           halo_size=halo_size, sharding=sharding)
        return _Lincomb.apply(1.5 * cosmo.Omega_m / (float(a)**2 * _E(cosmo, a)), forces, 0.0, None)

    return drift, kick


def make_ode_fn(mesh_shape, paint_absolute_pos=True, halo_size=0, sharding=None):
    def nbody_ode(state, a, cosmo):
        pos, vel = state
        forces = pm_forces(pos, mesh_shape=mesh_shape, paint_absolute_pos=paint_absolute_pos,
                           halo_size=halo_size, sharding=sharding)
        a_ = float(a)
        dpos = _Lincomb.apply(1. / (a_**3 * _E(cosmo, a_)), as_f32(vel), 0.0, None)
        dvel = _Lincomb.apply(1.5 * cosmo.Omega_m / (a_**2 * _E(cosmo, a_)), forces, 0.0, None)
        return dpos, dvel

    return nbody_ode


def make_diffrax_ode(mesh_shape, paint_absolute_pos=True, halo_size=0, sharding=None):
    f = make_ode_fn(mesh_shape, paint_absolute_pos, halo_size, sharding)

    def nbody_ode(a, state, args):
        dpos, dvel = f((state[0], state[1]), a, args)
        return torch.stack([dpos, dvel])

    return nbody_ode


def make_neural_ode_fn(model, mesh_shape):
    """ode.py:179-215 with the correction filter taken from `model(kk, a, params)`, a host
    callable of |k|/pi... (kk = sqrt(sum (k_i/pi)^2)) returning the multiplicative correction;
    it is tabulated and applied inside the fused k-space pass."""
    from .kernels import radial_filter_table

    def neural_nbody_ode(state, a, cosmo, params):
        pos, vel = state
        a_ = float(a)
        tab = radial_filter_table(lambda k: 1.0 + np.asarray(model(k / np.pi, a_, params)))
        forces = pm_forces(pos, mesh_shape=mesh_shape, filter_tab=tab)
        dpos = _Lincomb.apply(1. / (a_**3 * _E(cosmo, a_)), as_f32(vel), 0.0, None)
        dvel = _Lincomb.apply(1.5 * cosmo.Omega_m / (a_**2 * _E(cosmo, a_)), forces, 0.0, None)
        return dpos, dvel

    return neural_nbody_ode


# ---------------------------------------------------------------------------------------
# fused fixed-step drivers (no autograd; state updated in place on the device)
# ---------------------------------------------------------------------------------------
def kick_drift_coefficients(cosmo, a0, a1, nsteps, scheme="symplectic"):
    """Per-step scalars for the drift-kick sequence of diffrax.SemiImplicitEuler over
    symplectic_ode ("symplectic") or symplectic_fpm_ode ("fpm") terms:
        pos += d_n * vel ; vel += k_n * F(pos)          n = 0..nsteps-1
    returned as float64 arrays (d, k) with 1.5*Omega_m folded into k."""
    ts = np.linspace(a0, a1, nsteps + 1)
    d, k = np.empty(nsteps), np.empty(nsteps)
    for n in range(nsteps):
        a, dt = ts[n], ts[n + 1] - ts[n]
        if scheme == "symplectic":
            d[n] = dt / (a**3 * _E(cosmo, a))
            k[n] = dt * 1.5 * cosmo.Omega_m / (a**2 * _E(cosmo, a))
        elif scheme == "fpm":
            fd, fk, _ = fpm_factors(cosmo, a, dt)
            d[n] = dt * fd
            k[n] = dt * fk * 1.5 * cosmo.Omega_m
        else:
            raise ValueError(scheme)
    return d, k


def nbody_kick_drift(cosmo, pos, vel, a0, a1, nsteps, mesh_shape=None, paint_absolute_pos=True,
                     scheme="symplectic", halo_size=0, sharding=None, callback=None, resident=True,
                     tile=None, margin=None, force_mode="spectral", info=None):
    """Run `nsteps` drift-kick steps in place on (pos, vel) and return them.

    The first drift is a plain axpy; every following force evaluation is one fused
    paint -> FFT -> k-space -> 3x iFFT -> read3+kick+drift chain, with the drift of the NEXT
    step folded into the same kernel that applies the kick.  `resident=True` keeps the particles
    in the tile-sorted device state of jaxpm_b200/csrc/sim.cu between steps (fast for scattered,
    late-time distributions); `resident=False` runs the order-preserving kernels every step.

    `force_mode` (resident, power-of-two meshes): "spectral" = three inverse transforms of -gradient_kernel *
    pot_k as pm.py:54-56 writes them; "potential" = one inverse transform of pot_k and the 4th-order difference
    stencil that gradient_kernel (kernels.py:62-66) is the symbol of, applied in the read kernel; "auto" =
    per step, whichever the device-measured fp32 error bound allows (see include/jaxpm_b200.h).
    `info` (a dict) receives the force-path statistics of the run."""
    pos, vel = as_f32(pos), as_f32(vel)
    relative = not paint_absolute_pos
    mesh_shape = tuple(pos.shape[:3]) if (mesh_shape is None or relative) else tuple(mesh_shape)
    if not _single(sharding) and not relative:
        # the sharded steppers integrate displacements from Lagrangian sites (the only mode the reference supports
        # multi-device, painting.py:51-55); absolute positions would be silently wrong
        raise NotImplementedError("sharded nbody_kick_drift needs paint_absolute_pos=False (relative mode), "
                                  "like the reference's multi-device path")
    d, k = kick_drift_coefficients(cosmo, a0, a1, nsteps, scheme)
    ops.axpby(1.0, pos, d[0], vel, out=pos)
    if not _single(sharding):
        from . import halo
        return halo.nbody_kick_drift(pos, vel, d, k, mesh_shape, halo_size, sharding, callback, resident=resident,
                                     force_mode=force_mode)
    if not resident:
        plan = ops.get_plan(mesh_shape, pos.device)
        for n in range(nsteps):
            dn = d[n + 1] if n + 1 < nsteps else 0.0
            ops.pm_step_(plan, pos, vel, k[n], dn, relative)
            if callback is not None:
                callback(n, pos, vel)
        return pos, vel
    # resident tile-sorted state: load once, K fused steps, store back in the caller's order
    if margin is None:
        margin = 2 if force_mode == "spectral" else 1     # the potential path's psi box fills the 4-cell ghost zone
    sim = ops.Sim(mesh_shape, pos.shape[:3] if pos.dim() == 4 else (1, 1, pos.numel() // 3), relative,
                  pos.device, tile=tile, margin=margin)
    if force_mode != "spectral":
        sim.set_force_mode(force_mode)
    sim.load(pos, vel)
    for n in range(nsteps):
        sim.step(k[n], d[n + 1] if n + 1 < nsteps else 0.0)
        if callback is not None:
            sim.store(pos, vel)
            callback(n, pos, vel)
    sim.store(pos, vel)
    if info is not None:
        info.update(sim.force_info())
        info["fallbacks"] = sim.fallback_counts()
    return pos, vel


class _DriftKickStep(torch.autograd.Function):
    """One drift-kick step, pos' = pos + d vel ; vel' = vel + k F(pos'), differentiable with PER-STEP RECOMPUTE:
    only pos' (12 B per particle) is kept; the backward pass re-runs the force evaluation at pos' and pulls the
    cotangent through the hand-written adjoint kernels of pm_forces (paint^T = read, read^T = paint, the transposed
    k-space pass, the position gradients of the CIC weights).  What diffrax's RecursiveCheckpointAdjoint does for
    the reference's gradient runs (tests/test_gradients.py:27-28), at checkpoint granularity one step."""

    @staticmethod
    def forward(ctx, pos, vel, d, k, mesh_shape, relative):
        pos1 = ops.axpby(1.0, pos, d, vel)
        with torch.no_grad():
            F = pm_forces(pos1, mesh_shape=mesh_shape, paint_absolute_pos=not relative)
        vel1 = ops.axpby(1.0, vel, k, F)
        ctx.save_for_backward(pos1)
        ctx.cfg = (float(d), float(k), mesh_shape, relative)
        return pos1, vel1

    @staticmethod
    def backward(ctx, g_pos1, g_vel1):
        (pos1,) = ctx.saved_tensors
        d, k, mesh_shape, relative = ctx.cfg
        g_vel1 = g_vel1.contiguous()
        from . import pm as _pm
        if _pm._FAST_API and _pm._FUSED_VJP and ops.fast_path_shape(mesh_shape):
            # the force value is not needed here, only its vector-Jacobian product: straight to the fused adjoint passes
            gx = _pm._pm_forces_vjp_fused(pos1, ops.axpby(k, g_vel1), mesh_shape, relative, 0.0, None)
        else:
            with torch.enable_grad():
                x = pos1.detach().requires_grad_(True)
                F = pm_forces(x, mesh_shape=mesh_shape, paint_absolute_pos=not relative)
                (gx,) = torch.autograd.grad(F, x, ops.axpby(k, g_vel1))
        g_pos = ops.axpby(1.0, gx, 1.0, g_pos1.contiguous()) if g_pos1 is not None else gx
        g_vel = ops.axpby(1.0, g_vel1, d, g_pos)
        return g_pos, g_vel, None, None, None, None


def nbody_kick_drift_grad(cosmo, pos, vel, a0, a1, nsteps, mesh_shape=None, paint_absolute_pos=True,
                          scheme="symplectic"):
    """The fixed-step drift-kick run of `nbody_kick_drift` as a DIFFERENTIABLE function of (pos, vel): reverse mode
    through every step with per-step recompute (memory: one position array per step).  This is the driver of
    BASELINE.json's config 5 (gradient of the final power spectrum with respect to the initial conditions)."""
    pos, vel = as_f32(pos), as_f32(vel)
    relative = not paint_absolute_pos
    mesh_shape = tuple(pos.shape[:3]) if (mesh_shape is None or relative) else tuple(mesh_shape)
    d, k = kick_drift_coefficients(cosmo, a0, a1, nsteps, scheme)
    for n in range(nsteps):
        pos, vel = _DriftKickStep.apply(pos, vel, float(d[n]), float(k[n]), mesh_shape, relative)
    return pos, vel


def nbody_leapfrog_midpoint(cosmo, pos, vel, a0, a1, nsteps, mesh_shape=None, paint_absolute_pos=True):
    """diffrax.LeapfrogMidpoint + ConstantStepSize over make_diffrax_ode (the integrator of the
    reference's published runs, notebooks/05-MultiHost_PM.py:118-128):
        y_{n+1} = y_{n-1} + (t_{n+1} - t_{n-1}) f(t_n, y_n),  first step Euler.
    Two state copies ping-pong; one fused read3+update kernel per step."""
    pos, vel = as_f32(pos).clone(), as_f32(vel).clone()
    relative = not paint_absolute_pos
    mesh_shape = tuple(pos.shape[:3]) if (mesh_shape is None or relative) else tuple(mesh_shape)
    plan = ops.get_plan(mesh_shape, pos.device)
    ts = np.linspace(a0, a1, nsteps + 1)
    pm1, vm1, tm1 = pos.clone(), vel.clone(), ts[0]
    rho = torch.empty(plan.shape, dtype=torch.float32, device=pos.device)
    for n in range(nsteps):
        a, h = ts[n], ts[n + 1] - tm1
        rho.zero_()
        if relative:
            ops.cic_paint_dx_(rho, pos)
        else:
            ops.cic_paint_(rho, pos)
        f3 = ops.force_meshes_from_density(rho, plan)
        kick = h * 1.5 * cosmo.Omega_m / (a**2 * _E(cosmo, a))
        drift = h / (a**3 * _E(cosmo, a))
        # y_{n+1} written over y_{n-1}; then swap roles
        ops.read3_kick_drift_(f3, pos, vel, kick, drift, relative, pos_prev=pm1, vel_prev=vm1,
                              use_new_vel=False)
        tm1 = ts[n]
        pos, pm1 = pm1, pos
        vel, vm1 = vm1, vel
    return pos, vel
