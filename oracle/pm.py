"""Oracle (test infrastructure): force composition, LPT and Gaussian ICs.

Restates /root/reference/jaxpm/pm.py:12-58 (pm_forces), :61-126 (lpt),
:129-144 (linear_field, with the white-noise array passed in because JAX's
threefry stream is not reproducible here), :147-172 (pgd_correction, intended maths).
"""
import numpy as np

from . import cosmology as C
from . import kernels as K
from . import painting as P


def pm_forces(positions, mesh_shape=None, delta=None, r_split=0, paint_absolute_pos=True,
              kfilter=None):
    """pm.py:12-58.  ``kfilter`` is an optional real array multiplied into
    pot_k (the PGD / neural-filter slot, ode.py:194-196)."""
    positions = np.asarray(positions)
    dt = positions.dtype
    if mesh_shape is None:
        assert delta is not None
        mesh_shape = delta.shape
    if paint_absolute_pos:
        paint_fn = lambda pos: P.cic_paint(np.zeros(mesh_shape, dtype=dt), pos)
        read_fn = lambda m, pos: P.cic_read(m, pos)
    else:
        paint_fn = lambda d: P.cic_paint_dx(d)
        read_fn = lambda m, d: P.cic_read_dx(m, d)
    if delta is None:
        delta_k = K.fft3d(paint_fn(positions))
    elif np.isrealobj(delta):
        delta_k = K.fft3d(delta)
    else:
        delta_k = delta
    kvec = K.fftk(delta_k)
    pot_k = delta_k * K.invlaplace_kernel(kvec) * K.longrange_kernel(kvec, r_split)
    if kfilter is not None:
        pot_k = pot_k * kfilter
    forces = np.stack([
        read_fn(K.ifft3d(-K.gradient_kernel(kvec, i) * pot_k).astype(dt), positions)
        for i in range(3)
    ], axis=-1)
    return forces


def lpt(cosmo, initial_conditions, particles=None, a=0.1, order=1):
    """pm.py:61-126.  Returns (dx, p, f)."""
    ic = np.asarray(initial_conditions)
    dt = ic.dtype
    paint_absolute_pos = particles is not None
    if particles is None:
        particles = np.zeros((*ic.shape, 3), dtype=dt)
    a = float(a)
    E = C.E(cosmo, a)
    delta_k = K.fft3d(ic)
    f0 = pm_forces(particles, delta=delta_k, paint_absolute_pos=paint_absolute_pos)
    dx = (C.growth_factor(cosmo, a) * f0).astype(dt)
    p = (a**2 * C.growth_rate(cosmo, a) * E * dx).astype(dt)
    f = (a**2 * E * C.dGfa(cosmo, a) * f0).astype(dt)
    if order == 2:
        kvec = K.fftk(delta_k)
        pot_k = delta_k * K.invlaplace_kernel(kvec)
        delta2 = 0
        shear_acc = 0
        for i in range(3):
            shear_ii = K.ifft3d(K.gradient_kernel(kvec, i)**2 * pot_k)
            delta2 = delta2 + shear_ii * shear_acc
            shear_acc = shear_acc + shear_ii
            for j in range(i + 1, 3):
                nij = K.gradient_kernel(kvec, i) * K.gradient_kernel(kvec, j)
                delta2 = delta2 - K.ifft3d(nij * pot_k)**2
        f2 = pm_forces(particles, delta=K.fft3d(delta2.astype(dt)),
                       paint_absolute_pos=paint_absolute_pos)
        dx2 = (3 / 7 * C.growth_factor_second(cosmo, a) * f2).astype(dt)
        p2 = (a**2 * C.growth_rate_second(cosmo, a) * E * dx2).astype(dt)
        ff2 = (a**2 * E * C.dGf2a(cosmo, a) * f2).astype(dt)
        dx, p, f = dx + dx2, p + p2, f + ff2
    return dx, p, f


def linear_field(white_noise, box_size, pk):
    """pm.py:129-144 with the N(0,1) field supplied by the caller."""
    wn = np.asarray(white_noise)
    mesh_shape = wn.shape
    field = K.fft3d(wn)
    kvec = K.fftk(field)
    kmesh = sum((kk / box_size[i] * mesh_shape[i])**2 for i, kk in enumerate(kvec))**0.5
    pkmesh = pk(kmesh) * np.prod(mesh_shape) / np.prod(box_size)
    return K.ifft3d(field * np.sqrt(pkmesh)).astype(wn.dtype)


def pgd_correction(pos, mesh_shape, params):
    """pm.py:147-172 with the inverse transform the maths intends."""
    alpha, kl, ks = params
    kvec = K.fftk(mesh_shape)
    return alpha * pm_forces(pos, mesh_shape=mesh_shape, kfilter=K.PGD_kernel(kvec, kl, ks))
