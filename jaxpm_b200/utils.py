"""Power-spectrum estimator on the device, with reverse-mode gradient (the parity metric of BASELINE.json and
the loss of its config 5).  Mirrors /root/reference/jaxpm/utils.py:14-73 (`_initialize_pk`) and :76-128
(`power_spectrum`): same arguments, same default `kedges` rule (dk = twice the fundamental of the shortest side),
same `np.digitize` bins, `kavg` = mean |k| of the modes of a bin, result in (box units)^3.

Differences, stated: the transform is the library's R2C half-spectrum (modes with 0 < kz < Nyquist count twice);
multipoles are limited to l in {0, 2, 4} (odd ones vanish for real fields); the gradient is implemented for the
auto spectrum.  `kavg` comes back as a NumPy float64 array like in the reference, `pk` as a float32 CUDA tensor.
"""
import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import as_f32, call, ptr, stream

_geom_cache = {}


def _edges(mesh_shape, box_shape, kedges):
    """utils.py:40-50."""
    kmax = np.pi * np.min(mesh_shape / box_shape)
    if kedges is None or isinstance(kedges, (int, float)):
        if kedges is None:
            dk = 2 * np.pi / np.min(box_shape) * 2
        if isinstance(kedges, int):
            dk = kmax / (kedges + 1)
        elif isinstance(kedges, float):
            dk = kedges
        kedges = np.arange(dk, kmax, dk) + dk / 2
    return np.ascontiguousarray(np.asarray(kedges, dtype=np.float64))


class _Geom:
    """Per-(shape, box, edges, device) tables on the device + the mode counts / mean |k| of every bin."""

    def __init__(self, mesh_shape, box_shape, kedges, device):
        self.shape = tuple(int(s) for s in mesh_shape)
        self.box = np.asarray(box_shape, dtype=np.float64)
        self.edges_np = kedges
        nx, ny, nz = self.shape
        # (2 pi m / l) fftfreq(m), utils.py:52-53; the half-spectrum keeps the first nz/2+1 entries along z
        kv = [(2 * np.pi * m / l) * np.fft.fftfreq(m) for m, l in zip(self.shape, self.box)]
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)
        self.kx, self.ky, self.kz = dev(kv[0]), dev(kv[1]), dev(kv[2][:nz // 2 + 1])
        self.edges = dev(kedges)
        self.nb = len(kedges) + 1
        self.device = device
        self.kcount = None      # torch float64 [nb]
        self.kavg = None        # numpy float64 [nb - 2]
        self.cellvol = float((self.box / np.asarray(self.shape)).prod())
        self.norm = 1.0 / float(nx * ny * nz)


def _geom(mesh_shape, box_shape, kedges, device):
    key = (tuple(mesh_shape), tuple(float(b) for b in box_shape), kedges.tobytes(), str(device))
    if key not in _geom_cache:
        _geom_cache[key] = _Geom(mesh_shape, box_shape, kedges, device)
    return _geom_cache[key]


def _bin(g, spec_a, spec_b, ells, los):
    nl = len(ells)
    out = torch.empty((2 * nl + 2, g.nb), dtype=torch.float64, device=g.device)
    ells_c = (C.c_int32 * nl)(*ells)
    los_c = (C.c_float * 3)(*([0.0, 0.0, 1.0] if los is None else [float(v) for v in los]))
    want = g.kcount is None
    nx, ny, nz = g.shape
    call("jpm_pk_bin_c64", stream(), ptr(spec_a), ptr(spec_b) if spec_b is not None else None, nx, ny, nz,
         ptr(g.kx), ptr(g.ky), ptr(g.kz), ptr(g.edges), g.nb - 1, ells_c, nl, los_c, g.norm, int(want), ptr(out))
    if want:
        g.kcount = out[2 * nl].clone()
        with np.errstate(divide="ignore", invalid="ignore"):
            g.kavg = (out[2 * nl + 1] / g.kcount).cpu().numpy()[1:-1]
    return out


class _AutoPk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mesh, g, ells, los):
        plan = ops.get_plan(g.shape, mesh.device)
        spec = ops.rfft3(mesh, plan)
        out = _bin(g, spec, None, ells, los)
        nl = len(ells)
        pk = (out[:nl] / g.kcount)[:, 1:-1] * g.cellvol
        ctx.save_for_backward(spec)
        ctx.g, ctx.ells, ctx.los, ctx.plan = g, ells, los, plan
        return pk.to(torch.float32)

    @staticmethod
    def backward(ctx, gpk):
        (spec,) = ctx.saved_tensors
        g, ells, nl = ctx.g, ctx.ells, len(ctx.ells)
        # W[l][b] = g_lb * cell volume / kcount[b]; the overflow bins (0 and nb-1) do not reach the output
        w = torch.zeros((nl, g.nb), dtype=torch.float64, device=spec.device)
        w[:, 1:-1] = gpk.to(torch.float64) * g.cellvol / g.kcount[1:-1]
        w = torch.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0).contiguous()
        tmp = torch.empty_like(spec)
        ells_c = (C.c_int32 * nl)(*ells)
        los_c = (C.c_float * 3)(*([0.0, 0.0, 1.0] if ctx.los is None else [float(v) for v in ctx.los]))
        nx, ny, nz = g.shape
        call("jpm_pk_weight_c64", stream(), ptr(spec), ptr(tmp), nx, ny, nz, ptr(g.kx), ptr(g.ky), ptr(g.kz),
             ptr(g.edges), g.nb - 1, ells_c, nl, los_c, ptr(w), g.norm)
        grad = ops.irfft3_(tmp, ctx.plan, batch=1)      # unnormalised C2R: sums over the full spectrum
        return grad, None, None, None


def power_spectrum(mesh, mesh2=None, box_shape=None, kedges=None, multipoles=0, los=(0., 0., 1.)):
    """Auto / cross power spectrum of 3-D fields with (even) multipoles — jaxpm/utils.py:76-128.

    Returns (kavg, pk): pk has shape [n_bins] for a scalar `multipoles`, else [len(multipoles), n_bins]."""
    mesh = as_f32(mesh)
    if mesh.dim() != 3:
        raise ValueError("power_spectrum expects a 3-D mesh")
    mesh_shape = np.array(mesh.shape)
    box = mesh_shape.astype(np.float64) if box_shape is None else np.asarray(box_shape, dtype=np.float64)
    scalar = np.ndim(multipoles) == 0
    ells = [int(e) for e in np.atleast_1d(multipoles)]
    if any(e not in (0, 2, 4) for e in ells) or not 1 <= len(ells) <= 3:
        raise NotImplementedError("multipoles: up to three of l = 0, 2, 4")
    if scalar and ells[0] == 0:
        los_n = None
    else:
        los_n = np.asarray(los, dtype=np.float64)
        los_n = los_n / np.linalg.norm(los_n)
    edges = _edges(mesh_shape, box, kedges)
    g = _geom(tuple(int(s) for s in mesh_shape), box, edges, mesh.device)
    if mesh2 is None:
        pk = _AutoPk.apply(mesh, g, ells, los_n)
    else:
        plan = ops.get_plan(g.shape, mesh.device)
        sa, sb = ops.rfft3(mesh, plan), ops.rfft3(as_f32(mesh2), plan)
        out = _bin(g, sa, sb, ells, los_n)
        nl = len(ells)
        psum = (out[:nl]**2 + out[nl:2 * nl]**2).sqrt()
        pk = ((psum / g.kcount)[:, 1:-1] * g.cellvol).to(torch.float32)
    return (g.kavg, pk[0]) if scalar else (g.kavg, pk)


def transfer(mesh0, mesh1, box_shape, kedges=None):
    """sqrt(P1 / P0) - jaxpm/utils.py:131-135."""
    ks, pk0 = power_spectrum(mesh0, box_shape=box_shape, kedges=kedges)
    ks, pk1 = power_spectrum(mesh1, box_shape=box_shape, kedges=kedges)
    return ks, (pk1 / pk0)**.5


def coherence(mesh0, mesh1, box_shape, kedges=None):
    """P01 / sqrt(P0 P1) - jaxpm/utils.py:138-143."""
    ks, pk01 = power_spectrum(mesh0, mesh1, box_shape=box_shape, kedges=kedges)
    ks, pk0 = power_spectrum(mesh0, box_shape=box_shape, kedges=kedges)
    ks, pk1 = power_spectrum(mesh1, box_shape=box_shape, kedges=kedges)
    return ks, pk01 / (pk0 * pk1)**.5


def pktranscoh(mesh0, mesh1, box_shape, kedges=None):
    """(k, P0, P1, transfer, coherence) - jaxpm/utils.py:146-151."""
    ks, pk01 = power_spectrum(mesh0, mesh1, box_shape=box_shape, kedges=kedges)
    ks, pk0 = power_spectrum(mesh0, box_shape=box_shape, kedges=kedges)
    ks, pk1 = power_spectrum(mesh1, box_shape=box_shape, kedges=kedges)
    return ks, pk0, pk1, (pk1 / pk0)**.5, pk01 / (pk0 * pk1)**.5


def gaussian_smoothing(im, sigma):
    """Gaussian smoothing of a 2-D image with scale `sigma` in pixels - jaxpm/utils.py:208-222: the image spectrum times
    norm.pdf(|k|, 0, 1 / (2 pi sigma)) / norm.pdf(0, ...) = exp(-2 pi^2 sigma^2 |k|^2), k in cycles per pixel, real
    part of the inverse transform.  A small library-FFT image operation (light-cone planes), not part of the force loop;
    works on CPU or CUDA tensors."""
    im = torch.as_tensor(im)
    kx = torch.fft.fftfreq(im.shape[0], device=im.device, dtype=torch.float64)
    ky = torch.fft.fftfreq(im.shape[1], device=im.device, dtype=torch.float64)
    k2 = kx[:, None]**2 + ky[None, :]**2
    filt = torch.exp(-2.0 * np.pi**2 * float(sigma)**2 * k2)
    out = torch.fft.ifft2(torch.fft.fft2(im.to(torch.float64)) * filt).real
    return out.to(im.dtype if im.dtype.is_floating_point else torch.float32)

