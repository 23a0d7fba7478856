"""k-space kernels with the names/signatures of /root/reference/jaxpm/kernels.py
(fftk :10-23, gradient_kernel :41-66, invlaplace_kernel :69-92, longrange_kernel :95-115,
cic_compensation :118-136, PGD_kernel :139-165).

These are the *API-compatible* 1-D/broadcast forms for user code that composes
its own filters; the force loop itself never materialises them — it uses the
fused pass `jpm_greens_grad_c64` (jaxpm_b200/csrc/plan.cu).  They are tiny
(O(N) per axis), built on the host in float64 and uploaded, matching the layout
of `distributed.fft3d`'s output (R2C half-spectrum, last axis nz//2+1).
"""
import numpy as np
import torch


def fftk(k_array):
    """(kx, ky, kz) in radians per cell, broadcast-shaped against `k_array`
    (a `distributed.fft3d` output, which carries `.mesh_shape`) or a mesh shape tuple."""
    if isinstance(k_array, torch.Tensor):
        shape = getattr(k_array, "mesh_shape", None)
        if shape is None:
            raise ValueError("fftk needs the output of jaxpm_b200.distributed.fft3d (or a shape tuple)")
        dev = k_array.device
    else:
        shape, dev = tuple(k_array), None
    nx, ny, nz = shape
    kx = 2 * np.pi * np.fft.fftfreq(nx)
    ky = 2 * np.pi * np.fft.fftfreq(ny)
    kz = 2 * np.pi * np.fft.fftfreq(nz)[:nz // 2 + 1]
    out = [torch.as_tensor(k.astype(np.float32)).reshape(s)
           for k, s in ((kx, (-1, 1, 1)), (ky, (1, -1, 1)), (kz, (1, 1, -1)))]
    return tuple(o.to(dev) if dev is not None else o for o in out)


def gradient_kernel(kvec, direction, order=1):
    w = kvec[direction]
    if order == 0:
        # kernels.py:56-61 zeroes index len // 2 of a FULL-length frequency axis, i.e. the Nyquist mode |w| = pi.  The z
        # axis here is the R2C half axis (length nz // 2 + 1), where that mode is the LAST entry: zero by value
        wts = (1j * w).reshape(-1).clone()
        wts[w.reshape(-1).abs() >= np.pi * (1 - 1e-6)] = 0
        return wts.reshape(w.shape)
    a = 1 / 6.0 * (8 * torch.sin(w) - torch.sin(2 * w))
    return a * 1j


def invlaplace_kernel(kvec, fd=False):
    if fd:
        kk = sum((ki * torch.sinc(ki / (2 * np.pi)))**2 for ki in kvec)
    else:
        kk = sum(ki**2 for ki in kvec)
    kk_nz = torch.where(kk == 0, torch.ones_like(kk), kk)
    return -torch.where(kk == 0, torch.zeros_like(kk), 1 / kk_nz)


laplace_kernel = invlaplace_kernel  # the name BASELINE.json uses


def longrange_kernel(kvec, r_split):
    if r_split != 0:
        kk = sum(ki**2 for ki in kvec)
        return torch.exp(-kk * r_split**2)
    return 1.


def cic_compensation(kvec):
    kw = [torch.sinc(kvec[i] / (2 * np.pi)) for i in range(3)]
    return (kw[0] * kw[1] * kw[2])**(-2)


def PGD_kernel(kvec, kl, ks):
    """exp(-kl^2/k^2) exp(-k^4/ks^4), 0 at k=0 (the intended maths of kernels.py:157-165)."""
    kk = sum(ki**2 for ki in kvec)
    nz = kk != 0
    kk1 = torch.where(nz, kk, torch.ones_like(kk))
    return torch.exp(-kl**2 / kk1) * torch.exp(-kk1**2 / ks**4) * nz


def radial_filter_table(fn, n_tab=4096, kmax=None):
    """Tabulate a radial filter f(|k|) on [0, kmax] (default sqrt(3)*pi) for the fused pass."""
    kmax = np.sqrt(3.0) * np.pi * 1.0001 if kmax is None else kmax
    k = np.linspace(0.0, kmax, n_tab)
    return np.asarray(fn(k), dtype=np.float32), float(kmax)


def pgd_filter_table(kl, ks, n_tab=8192):
    def f(k):
        kk = k**2
        out = np.zeros_like(kk)
        nz = kk > 0
        out[nz] = np.exp(-kl**2 / kk[nz]) * np.exp(-kk[nz]**2 / ks**4)
        return out
    return radial_filter_table(f, n_tab)
