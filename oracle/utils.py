"""Oracle (test infrastructure): the power-spectrum estimator (the parity metric).

Restates /root/reference/jaxpm/utils.py:14-73 (_initialize_pk) and :76-128
(power_spectrum), monopole / multipoles, auto and cross.
"""
import numpy as np
import scipy.fft as sfft
from scipy.special import legendre


def _initialize_pk(mesh_shape, box_shape, kedges, los, x64=False):
    mesh_shape = np.asarray(mesh_shape)
    box_shape = np.asarray(box_shape, dtype=np.float64)
    kmax = np.pi * np.min(mesh_shape / box_shape)
    if kedges is None or isinstance(kedges, (int, float)):
        if kedges is None:
            dk = 2 * np.pi / np.min(box_shape) * 2
        if isinstance(kedges, int):
            dk = kmax / (kedges + 1)
        elif isinstance(kedges, float):
            dk = kedges
        kedges = np.arange(dk, kmax, dk) + dk / 2
    kshapes = np.eye(len(mesh_shape), dtype=np.int32) * -2 + 1
    kvec = [(2 * np.pi * m / l) * np.fft.fftfreq(m).reshape(ks)
            for m, l, ks in zip(mesh_shape, box_shape, kshapes)]
    kmesh = np.sqrt(sum(ki**2 for ki in kvec))
    if not x64:   # utils.py:54 is a jnp.sqrt: float32 unless jax_enable_x64 (decides edge-on-bin modes)
        kmesh = kmesh.astype(np.float32)
    dig = np.digitize(kmesh.reshape(-1), kedges)
    kcount = np.bincount(dig, minlength=len(kedges) + 1)
    kavg = np.bincount(dig, weights=kmesh.reshape(-1), minlength=len(kedges) + 1) / kcount
    kavg = kavg[1:-1]
    if los is None:
        mumesh = 1.
    else:
        mumesh = sum(ki * li for ki, li in zip(kvec, los))
        knz = np.where(kmesh == 0, 1, kmesh)
        mumesh = np.where(kmesh == 0, 0, mumesh / knz)
    return dig, kcount, kavg, mumesh


def power_spectrum(mesh, mesh2=None, box_shape=None, kedges=None, multipoles=0, los=(0., 0., 1.), x64=False):
    mesh = np.asarray(mesh, dtype=np.float64)
    mesh_shape = np.array(mesh.shape)
    box_shape = mesh_shape if box_shape is None else np.asarray(box_shape)
    if np.ndim(multipoles) == 0 and multipoles == 0:
        los = None
    else:
        los = np.asarray(los, dtype=np.float64)
        los = los / np.linalg.norm(los)
    poles = np.atleast_1d(multipoles)
    dig, kcount, kavg, mumesh = _initialize_pk(mesh_shape, box_shape, kedges, los, x64)
    n_bins = len(kavg) + 2
    meshk = sfft.fftn(mesh, norm='ortho', workers=-1)
    if mesh2 is None:
        mmk = meshk.real**2 + meshk.imag**2
    else:
        mmk = meshk * sfft.fftn(np.asarray(mesh2, dtype=np.float64), norm='ortho', workers=-1).conj()
    pk = np.empty((len(poles), n_bins))
    for i, ell in enumerate(poles):
        wts = (mmk * (2 * ell + 1) * legendre(ell)(mumesh)).reshape(-1)
        if mesh2 is None:
            psum = np.bincount(dig, weights=wts, minlength=n_bins)
        else:
            psum = (np.bincount(dig, weights=wts.real, minlength=n_bins)**2 +
                    np.bincount(dig, weights=wts.imag, minlength=n_bins)**2)**.5
        pk[i] = psum
    pk = (pk / kcount)[:, 1:-1] * (box_shape / mesh_shape).prod()
    return (kavg, pk[0]) if np.ndim(multipoles) == 0 else (kavg, pk)


def MSE(x, y):
    return np.mean((np.asarray(x) - np.asarray(y))**2)


def MSRE(x, y):
    return np.mean(((np.asarray(x) - np.asarray(y)) / np.asarray(y))**2)
