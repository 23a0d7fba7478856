"""Oracle (test infrastructure): ODE right-hand sides and fixed-step drivers.

Restates /root/reference/jaxpm/ode.py:13-82 (symplectic_fpm_ode), :85-119
(symplectic_ode), :122-147 (make_ode_fn), :150-176 (make_diffrax_ode).

The reference leaves time stepping to diffrax [ext, absent].  Two fixed-step
schemes its notebooks/tests use are restated from diffrax's published update
rules: ``SemiImplicitEuler`` (drift with vel_n, then kick with the new
positions; the natural consumer of ``symplectic_ode``) and
``LeapfrogMidpoint`` (y_{n+1} = y_{n-1} + (t_{n+1}-t_{n-1}) f(t_n, y_n), Euler
first step; notebooks/04-MultiGPU_PM_Solvers.ipynb, 05-MultiHost_PM.py:118-128).
"""
import numpy as np

from . import cosmology as C
from .pm import pm_forces


def make_ode_fn(mesh_shape, paint_absolute_pos=True):
    def nbody_ode(state, a, cosmo):
        pos, vel = state
        forces = pm_forces(pos, mesh_shape=mesh_shape,
                           paint_absolute_pos=paint_absolute_pos) * 1.5 * cosmo.Omega_m
        E = np.sqrt(C.Esqr(cosmo, a))
        dpos = (1. / (a**3 * E) * vel).astype(pos.dtype)
        dvel = (1. / (a**2 * E) * forces).astype(pos.dtype)
        return dpos, dvel
    return nbody_ode


def make_diffrax_ode(mesh_shape, paint_absolute_pos=True):
    f = make_ode_fn(mesh_shape, paint_absolute_pos)

    def nbody_ode(a, state, args):
        return np.stack(f((state[0], state[1]), a, args))
    return nbody_ode


def symplectic_ode(mesh_shape, cosmo, paint_absolute_pos=True):
    def drift(a, vel, args):
        return (1 / (a**3 * C.E(cosmo, a)) * vel).astype(vel.dtype)

    def kick(a, pos, args):
        forces = pm_forces(pos, mesh_shape=mesh_shape,
                           paint_absolute_pos=paint_absolute_pos) * 1.5 * cosmo.Omega_m
        return (1.0 / (a**2 * C.E(cosmo, a)) * forces).astype(pos.dtype)
    return drift, kick


def fpm_factors(cosmo, a, dt0):
    """Scalar factors of symplectic_fpm_ode (ode.py:24-33, :39-58, :64-80):
    returns (drift_coef, kick_coef, first_kick_coef), each multiplying vel /
    (1.5 Om F) *per unit dt0* exactly as the reference's drift/kick/first_kick."""
    t0, t1, t2 = a, a + dt0, a + 2 * dt0
    ac = (t0 * t1)**0.5
    drift_contr = (C.growth_factor(cosmo, t1) - C.growth_factor(cosmo, t0)) / C.gp(cosmo, ac)
    drift = 1 / (ac**3 * C.E(cosmo, ac)) * (drift_contr / dt0)
    t0t1, t1t2 = (t0 * t1)**0.5, (t1 * t2)**0.5
    k1 = (C.Gf(cosmo, t1) - C.Gf(cosmo, t0t1)) / C.dGfa(cosmo, t1)
    k2 = (C.Gf(cosmo, t1t2) - C.Gf(cosmo, t1)) / C.dGfa(cosmo, t1)
    kick = 1.0 / (t1**2 * C.E(cosmo, t1)) * ((k1 + k2) / dt0)
    fk = (C.Gf(cosmo, t0t1) - C.Gf(cosmo, t0)) / C.dGfa(cosmo, t0)
    first_kick = 1.0 / (a**2 * C.E(cosmo, a)) * (fk / dt0)
    return float(drift), float(kick), float(first_kick)


def symplectic_fpm_ode(mesh_shape, dt0, cosmo, paint_absolute_pos=True):
    def F(pos):
        return pm_forces(pos, mesh_shape=mesh_shape,
                         paint_absolute_pos=paint_absolute_pos) * 1.5 * cosmo.Omega_m

    def drift(a, vel, args):
        return (fpm_factors(cosmo, a, dt0)[0] * vel).astype(vel.dtype)

    def kick(a, pos, args):
        return (fpm_factors(cosmo, a, dt0)[1] * F(pos)).astype(pos.dtype)

    def first_kick(a, pos, args):
        return (fpm_factors(cosmo, a, dt0)[2] * F(pos)).astype(pos.dtype)
    return drift, kick, first_kick


def semi_implicit_euler(drift, kick, pos, vel, a0, a1, nsteps, args=None):
    """diffrax.SemiImplicitEuler with ConstantStepSize over terms (drift, kick)."""
    ts = np.linspace(a0, a1, nsteps + 1)
    for n in range(nsteps):
        dt = ts[n + 1] - ts[n]
        pos = (pos + dt * drift(ts[n], vel, args)).astype(pos.dtype)
        vel = (vel + dt * kick(ts[n], pos, args)).astype(vel.dtype)
    return pos, vel


def leapfrog_midpoint(ode, y0, a0, a1, nsteps, args=None):
    """diffrax.LeapfrogMidpoint with ConstantStepSize; ``ode(a, y, args)`` in diffrax order."""
    ts = np.linspace(a0, a1, nsteps + 1)
    tm1, ym1 = ts[0], y0
    y = y0
    for n in range(nsteps):
        y1 = (ym1 + (ts[n + 1] - tm1) * ode(ts[n], y, args)).astype(y0.dtype)
        tm1, ym1 = ts[n], y
        y = y1
    return y
