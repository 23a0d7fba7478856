"""Distributed 3-D real FFT over a (px, py) process grid — the [ext] jaxdecomp.pfft3d /
pifft3d the reference calls from /root/reference/jaxpm/distributed.py:37-42.

Layout (C = complex64, one process per GPU, rank = rx*py + ry):
  real      [lx = nx/px][ly = ny/py][nz]                       block (rx, ry)
  stage 1   R2C along z                  -> [lx][ly][nzh]
  transp A  all-to-all in the y-group (same rx): z-modes split py ways (uneven allowed), y gathered
                                         -> [lx][nzl][ny]       (y contiguous)
  stage 2   C2C along y
  transp B  all-to-all in the x-group (same ry): y split px ways, x gathered
                                         -> [ny2 = ny/px][nzl][nx]   (x contiguous)
  stage 3   C2C along x
The spectrum block therefore has array axes (y, z, x); `kspace_tables()` gives the matching slices
of the fftk / gradient tables and `axis_map = (2, 0, 1)` tells the fused k-space kernel where the
physical x, y, z directions live (the reference relies on exactly this freedom: fftk(delta_k) only
has to broadcast against delta_k, jaxpm/kernels.py:10-23).  With px == 1 or py == 1 one of the two
all-to-alls degenerates to a local repack (slab decomposition: one transpose per FFT).

Pack / unpack are the library's own transpose / strided-copy kernels; the collective is
`torch.distributed.all_to_all_single` (NCCL over NVLink).  A `backend` object supplies the local
compute so that the communication schedule can be exercised on CPU (gloo) in the tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def split_sizes(n, parts):
    """Chunk sizes of n items over `parts` ranks (first n % parts ranks get one more)."""
    q, r = divmod(n, parts)
    return [q + (1 if i < r else 0) for i in range(parts)]


def kspace_tables_1d(n, nh=None):
    """(w, a) of jaxpm/kernels.py:10-23 and :62-66 for one axis, as the CUDA plan builds them
    (jaxpm_b200/csrc/plan.cu build_tables): w fp32 from float64, a evaluated in float64 at the
    fp32 frequency, the self-conjugate Nyquist entry of the odd kernel set to exactly 0."""
    nh = n if nh is None else nh
    i = np.arange(nh)
    f = np.where(i < (n + 1) // 2, i, i - n)
    w = (2.0 * np.pi * f / n).astype(np.float32)
    wd = w.astype(np.float64)
    a = ((8.0 * np.sin(wd) - np.sin(2.0 * wd)) / 6.0).astype(np.float32)
    if n % 2 == 0 and n // 2 < nh:
        a[n // 2] = 0.0
    return w, a


class CudaBackend:
    """Local compute through the C ABI (cuFFT 1-D batched plans + the pack/unpack kernels)."""

    def __init__(self, device):
        self.device = device
        self._plans = {}

    def empty_c(self, n):
        return torch.empty(n, dtype=torch.complex64, device=self.device)

    def empty_r(self, shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def _plan(self, n, batch, kind):
        import ctypes as C
        from . import _lib
        key = (n, batch, kind)
        if key not in self._plans:
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.call("jpm_fft1d_create", C.byref(h), n, batch, kind)
            self._plans[key] = h
        return self._plans[key]

    def rfft_z(self, x, out):
        from . import _lib
        n = x.shape[-1]
        _lib.call("jpm_fft1d_exec", self._plan(n, x.numel() // n, 0), _lib.stream(), _lib.ptr(x), _lib.ptr(out), 0)

    def irfft_z(self, spec, out):
        from . import _lib
        n = out.shape[-1]
        _lib.call("jpm_fft1d_exec", self._plan(n, out.numel() // n, 1), _lib.stream(), _lib.ptr(spec), _lib.ptr(out), 0)

    def cfft(self, buf, n, inverse):
        from . import _lib
        _lib.call("jpm_fft1d_exec", self._plan(n, buf.numel() // n, 2), _lib.stream(), _lib.ptr(buf), _lib.ptr(buf),
                  int(inverse))

    def copy2d(self, dst, doff, src, soff, nrows, ncols, srs, drs):
        from . import _lib
        _lib.call("jpm_copy2d_c64", _lib.stream(), dst.data_ptr() + 8 * doff, src.data_ptr() + 8 * soff, nrows,
                  ncols, srs, drs)

    def transpose(self, dst, doff, src, soff, ni, nj, nb, ssi, ssb, dsj, dsb):
        from . import _lib
        _lib.call("jpm_transpose_c64", _lib.stream(), dst.data_ptr() + 8 * doff, src.data_ptr() + 8 * soff, ni, nj,
                  nb, ssi, ssb, dsj, dsb)

    def kspace(self, kind, inp, out, tabs, shape, axis_map, norm, r_split=0.0, filter_tab=None):
        from . import _lib, ops
        fp, nt, km, keep = ops._ftab(filter_tab, self.device)
        _lib.call("jpm_kspace_local_c64", _lib.stream(), kind, _lib.ptr(inp), _lib.ptr(out),
                  *[_lib.ptr(t) for t in tabs], *shape, *axis_map, float(norm), float(r_split), fp, nt, km)


def _a2a(out, inp, out_splits, in_splits, group):
    """all_to_all_single on complex buffers (sent as float32 pairs)."""
    if group is None or dist.get_world_size(group) == 1:
        out.copy_(inp)
        return
    dist.all_to_all_single(torch.view_as_real(out).reshape(-1), torch.view_as_real(inp).reshape(-1),
                           [2 * s for s in out_splits], [2 * s for s in in_splits], group=group)


class PencilFFT:
    def __init__(self, mesh_shape, sharding, backend=None, device=None):
        self.shape = tuple(int(s) for s in mesh_shape)
        self.sh = sharding
        nx, ny, nz = self.shape
        px, py = sharding.pdims
        if nx % px or ny % py or ny % px:
            raise ValueError(f"mesh {self.shape} not divisible by pdims {sharding.pdims} (need nx%px, ny%py, ny%px == 0)")
        self.lx, self.ly, self.nzh = nx // px, ny // py, nz // 2 + 1
        self.zsplit = split_sizes(self.nzh, py)
        self.nzl = self.zsplit[sharding.ry]
        self.zoff = sum(self.zsplit[:sharding.ry])
        self.ny2 = ny // px
        self.yoff = sharding.rx * self.ny2
        self.spec_shape = (self.ny2, self.nzl, nx)       # array axes (y, z, x)
        self.axis_map = (2, 0, 1)                        # physical x, y, z -> array axis
        self.ncell = nx * ny * nz
        self.backend = backend if backend is not None else CudaBackend(device)
        self.ygroup, self.xgroup = sharding.ygroup, sharding.xgroup
        self._tabs = None

    # ---- k-space tables for the local block --------------------------------------------------
    def kspace_tables(self, to_device):
        if self._tabs is None:
            nx, ny, nz = self.shape
            wx, ax = kspace_tables_1d(nx)
            wy, ay = kspace_tables_1d(ny)
            wz, az = kspace_tables_1d(nz, self.nzh)
            ys, zs = slice(self.yoff, self.yoff + self.ny2), slice(self.zoff, self.zoff + self.nzl)
            # order: w0, w1, w2, a0, a1, a2 for array axes (y, z, x)
            self._tabs = [to_device(np.ascontiguousarray(t)) for t in (wy[ys], wz[zs], wx, ay[ys], az[zs], ax)]
        return self._tabs

    # ---- forward: real [nb, lx, ly, nz] -> spectrum [nb, ny2, nzl, nx] -------------------------
    def forward(self, x):
        be = self.backend
        nx, ny, nz = self.shape
        lx, ly, nzh, nzl, ny2 = self.lx, self.ly, self.nzh, self.nzl, self.ny2
        px, py = self.sh.pdims
        nb = x.shape[0]
        a1 = be.empty_c(nb * lx * ly * nzh)
        be.rfft_z(x, a1)
        # transpose A (y-group): pack z-chunks per destination
        send = be.empty_c(nb * lx * ly * nzh)
        off, zo = 0, 0
        for r in range(py):
            zs = self.zsplit[r]
            be.copy2d(send, off, a1, zo, nb * lx * ly, zs, nzh, zs)
            off += nb * lx * ly * zs
            zo += zs
        recv = be.empty_c(py * nb * lx * ly * nzl)
        _a2a(recv, send, [nb * lx * ly * nzl] * py, [nb * lx * ly * zs for zs in self.zsplit], self.ygroup)
        a2 = be.empty_c(nb * lx * nzl * ny)
        for r in range(py):   # [nb*lx][ly][nzl] -> a2[nb*lx][nzl][r*ly + y]
            be.transpose(a2, r * ly, recv, r * nb * lx * ly * nzl, ly, nzl, nb * lx, nzl, ly * nzl, ny, nzl * ny)
        be.cfft(a2, ny, False)
        # transpose B (x-group): pack y-chunks per destination
        send = be.empty_c(nb * lx * nzl * ny)
        for r in range(px):
            be.copy2d(send, r * nb * lx * nzl * ny2, a2, r * ny2, nb * lx * nzl, ny2, ny, ny2)
        recv = be.empty_c(px * nb * lx * nzl * ny2)
        _a2a(recv, send, [nb * lx * nzl * ny2] * px, [nb * lx * nzl * ny2] * px, self.xgroup)
        a3 = be.empty_c(nb * ny2 * nzl * nx)
        for r in range(px):   # [nb][lx][nzl][ny2] -> a3[nb][ny2][nzl][r*lx + x]
            for b in range(nb):
                be.transpose(a3, b * ny2 * nzl * nx + r * lx, recv, (r * nb + b) * lx * nzl * ny2,
                             lx, ny2, nzl, nzl * ny2, ny2, nzl * nx, nx)
        be.cfft(a3, nx, False)
        return a3.reshape(nb, ny2, nzl, nx)

    # ---- inverse: spectrum [nb, ny2, nzl, nx] (destroyed) -> real [nb, lx, ly, nz], UNNORMALISED --
    def inverse(self, spec):
        be = self.backend
        nx, ny, nz = self.shape
        lx, ly, nzh, nzl, ny2 = self.lx, self.ly, self.nzh, self.nzl, self.ny2
        px, py = self.sh.pdims
        nb = spec.shape[0]
        a3 = spec.reshape(-1)
        be.cfft(a3, nx, True)
        send = be.empty_c(px * nb * lx * nzl * ny2)
        for r in range(px):   # a3[nb][ny2][nzl][r*lx + x] -> [nb][lx][nzl][ny2]
            for b in range(nb):
                be.transpose(send, (r * nb + b) * lx * nzl * ny2, a3, b * ny2 * nzl * nx + r * lx,
                             ny2, lx, nzl, nzl * nx, nx, nzl * ny2, ny2)
        recv = be.empty_c(px * nb * lx * nzl * ny2)
        _a2a(recv, send, [nb * lx * nzl * ny2] * px, [nb * lx * nzl * ny2] * px, self.xgroup)
        a2 = be.empty_c(nb * lx * nzl * ny)
        for s in range(px):
            be.copy2d(a2, s * ny2, recv, s * nb * lx * nzl * ny2, nb * lx * nzl, ny2, ny2, ny)
        be.cfft(a2, ny, True)
        send = be.empty_c(py * nb * lx * ly * nzl)
        for r in range(py):   # a2[nb*lx][nzl][r*ly + y] -> [nb*lx][ly][nzl]
            be.transpose(send, r * nb * lx * ly * nzl, a2, r * ly, nzl, ly, nb * lx, ny, nzl * ny, nzl, ly * nzl)
        recv = be.empty_c(nb * lx * ly * nzh)
        _a2a(recv, send, [nb * lx * ly * zs for zs in self.zsplit], [nb * lx * ly * nzl] * py, self.ygroup)
        a1 = be.empty_c(nb * lx * ly * nzh)
        off, zo = 0, 0
        for s in range(py):
            zs = self.zsplit[s]
            be.copy2d(a1, zo, recv, off, nb * lx * ly, zs, zs, nzh)
            off += nb * lx * ly * zs
            zo += zs
        out = be.empty_r((nb, lx, ly, nz))
        be.irfft_z(a1, out)
        return out


_ffts = {}


def get_pfft(mesh_shape, sharding, device):
    key = (tuple(mesh_shape), sharding.pdims, sharding.rank, torch.device(device).index or 0)
    if key not in _ffts:
        _ffts[key] = PencilFFT(mesh_shape, sharding, device=torch.device(device))
    return _ffts[key]


class ShardedSpectrum(torch.Tensor):
    """Local block of a distributed spectrum (array axes (y, z, x)) + the transform that made it."""

    @staticmethod
    def wrap(t, fft):
        out = t.as_subclass(ShardedSpectrum)
        out.fft = fft
        out.mesh_shape = fft.shape
        return out


def pfft3d(x, sharding):
    fft = get_pfft(tuple(sharding.global_shape(x.shape)), sharding, x.device)
    return ShardedSpectrum.wrap(fft.forward(x.unsqueeze(0))[0], fft)


def pifft3d(spec, sharding):
    from . import ops
    fft = spec.fft
    out = fft.inverse(spec.as_subclass(torch.Tensor).clone().unsqueeze(0))[0]
    return ops.axpby(1.0 / fft.ncell, out, out=out)
