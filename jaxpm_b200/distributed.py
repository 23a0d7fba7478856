"""Distribution shim with the names of /root/reference/jaxpm/distributed.py
(fft3d :37-38, ifft3d :41-42, get_halo_size :45-58, halo_exchange :61-65,
slice_pad :88-99, slice_unpad :102-113, get_local_shape :116-129,
uniform_particles :168-190, normal_field :193-223).

Execution model: ONE PROCESS PER GPU (`torch.distributed`, NCCL over NVLink).
A `Sharding` plays the role of `NamedSharding(mesh, P('x','y'))`: the (x, y)
axes of every mesh / particle array are split over a (px, py) process grid, z
stays local, and every array argument is the LOCAL block of the calling rank
(what `shard_map` hands to the per-shard function in the reference).
`sharding=None` means single device.
"""
import numpy as np
import torch

from . import ops
from ._lib import as_f32


class Sharding:
    """2-D domain decomposition over a (px, py) process grid.

    rank = rx * py + ry  (row-major, like jax.make_mesh(pdims)); block (rx, ry) owns
    mesh[rx*nx/px:(rx+1)*nx/px, ry*ny/py:(ry+1)*ny/py, :] and — particles never migrate
     — the particles whose Lagrangian cell lies in that block
    (distributed.py:176-184)."""

    def __init__(self, pdims, group=None, rank=None):
        import torch.distributed as dist
        self.pdims = (int(pdims[0]), int(pdims[1]))
        self.group = group
        self.size = self.pdims[0] * self.pdims[1]
        live = dist.is_available() and dist.is_initialized()
        if rank is None:
            rank = dist.get_rank(group) if live else 0
        self.rank = rank
        self.rx, self.ry = divmod(rank, self.pdims[1])
        self.ygroup = self.xgroup = None
        if live and self.size > 1:
            if dist.get_world_size(group) != self.size:
                raise ValueError(f"pdims {self.pdims} needs {self.size} ranks, group has {dist.get_world_size(group)}")
            px, py = self.pdims
            # every rank must create every sub-group, in the same order
            for rx in range(px):
                g = dist.new_group([self.global_rank(rx * py + j) for j in range(py)]) if py > 1 else None
                if rx == self.rx:
                    self.ygroup = g
            for ry in range(py):
                g = dist.new_group([self.global_rank(i * py + ry) for i in range(px)]) if px > 1 else None
                if ry == self.ry:
                    self.xgroup = g

    def global_rank(self, r):
        """Rank `r` of this sharding's group as a global (world) rank."""
        import torch.distributed as dist
        if self.group is None:
            return r
        return dist.get_global_rank(self.group, r)

    def global_shape(self, local_shape):
        return (local_shape[0] * self.pdims[0], local_shape[1] * self.pdims[1], *local_shape[2:])

    def neighbor(self, dx, dy):
        px, py = self.pdims
        return ((self.rx + dx) % px) * py + (self.ry + dy) % py

    def local_shape(self, mesh_shape):
        return get_local_shape(mesh_shape, self)

    def __repr__(self):
        return f"Sharding(pdims={self.pdims}, rank={self.rank})"


class HalfSpectrum(torch.Tensor):
    """complex64 R2C half-spectrum [nx, ny, nz//2+1] that remembers the real mesh shape,
    so that `fftk(delta_k)` / `ifft3d(delta_k)` work like in the reference."""

    @staticmethod
    def wrap(t, mesh_shape):
        out = t.as_subclass(HalfSpectrum)
        out.mesh_shape = tuple(mesh_shape)
        return out

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        res = super().__torch_function__(func, types, args, kwargs or {})
        if isinstance(res, HalfSpectrum) and not hasattr(res, "mesh_shape"):
            for a in args:
                if isinstance(a, HalfSpectrum) and hasattr(a, "mesh_shape") and a.shape == res.shape:
                    res.mesh_shape = a.mesh_shape
                    break
        return res


def _single(sharding):
    return sharding is None or sharding.size == 1


def fft3d(x, sharding=None):
    """Unnormalised forward 3-D FFT of a real field (R2C half-spectrum; the reference's
    C2C output carries the same information for real input)."""
    if isinstance(x, torch.Tensor) and x.is_complex():
        raise NotImplementedError("fft3d of a complex field is not on the force-loop path")
    if not _single(sharding):
        from . import pfft
        return pfft.pfft3d(as_f32(x), sharding)
    x = as_f32(x)
    return HalfSpectrum.wrap(ops.rfft3(x), x.shape)


def ifft3d(x, sharding=None):
    """Real part of the normalised inverse FFT (distributed.py:41-42)."""
    if not _single(sharding):
        from . import pfft
        return pfft.pifft3d(x, sharding)
    shape = getattr(x, "mesh_shape", None)
    if shape is None:
        raise ValueError("ifft3d needs the output of fft3d (a HalfSpectrum)")
    plan = ops.get_plan(shape, x.device)
    spec = torch.Tensor.contiguous(x.as_subclass(torch.Tensor)).clone()
    out = ops.irfft3_(spec, plan, 1)
    return ops.axpby(1.0 / plan.ncell, out, out=out)


def get_halo_size(halo_size, sharding):
    """((hx,hx),(hy,hy),(0,0)), (ex, ey) — distributed.py:45-58; accepts an int or a 2-tuple."""
    if _single(sharding):
        return ((0, 0), (0, 0), (0, 0)), (0, 0)
    pdims = sharding.pdims
    if np.isscalar(halo_size):
        halo_size = (int(halo_size), int(halo_size))
    hx = (0, 0) if pdims[0] == 1 else (halo_size[0],) * 2
    hy = (0, 0) if pdims[1] == 1 else (halo_size[1],) * 2
    ex = 0 if pdims[0] == 1 else halo_size[0] // 2
    ey = 0 if pdims[1] == 1 else halo_size[1] // 2
    return (hx, hy, (0, 0)), (ex, ey)


def get_local_shape(mesh_shape, sharding=None):
    if _single(sharding):
        return list(mesh_shape)
    px, py = sharding.pdims
    if mesh_shape[0] % px or mesh_shape[1] % py:
        raise ValueError(f"mesh {tuple(mesh_shape)} not divisible by pdims {sharding.pdims}")
    return [mesh_shape[0] // px, mesh_shape[1] // py, *mesh_shape[2:]]


def uniform_particles(mesh_shape, sharding=None, device="cuda"):
    """Lagrangian grid positions [nx,ny,nz,3] (local block when sharded), distributed.py:168-190."""
    loc = get_local_shape(mesh_shape, sharding)
    ox = 0 if _single(sharding) else sharding.rx * loc[0]
    oy = 0 if _single(sharding) else sharding.ry * loc[1]
    zeros = torch.zeros((*loc, 3), dtype=torch.float32, device=device)
    return ops.grid_plus_disp(zeros, (ox, oy))


def normal_field(seed, shape, sharding=None, dtype=torch.float32, device="cuda", stream_id=0):
    """N(0,1) field (local block when sharded) from the device generator jpm_normal_field_f32: counter-based
    (Philox4x32-10 keyed by `seed`, counter = global cell group), so the value of a cell depends only on
    (seed, global cell index).  Unlike the reference (distributed.py:204-215: one independent key per device) a
    sharded call draws exactly the single-device field.  Not JAX's threefry stream: the same seed does not
    reproduce a JAX run; parity runs share the IC array (`white_noise=` of linear_field)."""
    from ._lib import call, ptr, stream
    shape = tuple(int(n) for n in shape)
    if len(shape) != 3:
        raise NotImplementedError("normal_field: 3-D meshes (the force-loop path)")
    loc = get_local_shape(shape, sharding)
    ox = 0 if _single(sharding) else sharding.rx * loc[0]
    oy = 0 if _single(sharding) else sharding.ry * loc[1]
    out = torch.empty(tuple(loc), dtype=torch.float32, device=device)
    call("jpm_normal_field_f32", stream(), ptr(out), loc[0], loc[1], loc[2], ox, oy, shape[1],
         int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id))
    return out if dtype == torch.float32 else out.to(dtype)


# ---- halo protocol (multi-GPU; implemented in jaxpm_b200/halo.py) -----------------------
def halo_exchange(x, halo_extents, halo_periods=(True, True), sharding=None):
    if _single(sharding) or not (halo_extents[0] > 0 or halo_extents[1] > 0):
        return x
    from . import halo
    return halo.halo_exchange(x, halo_extents, sharding)


def slice_pad(x, pad_width, sharding):
    if _single(sharding) or not (pad_width[0][0] > 0 or pad_width[1][0] > 0):
        return x
    from . import halo
    return halo.pad(x, pad_width)


def slice_unpad(x, pad_width, sharding):
    if _single(sharding) or not (pad_width[0][0] > 0 or pad_width[1][0] > 0):
        return x
    from . import halo
    return halo.unpad_reduce(x, pad_width)


def sharded_cic_paint_dx(displacements, weight, halo_size, sharding):
    from . import halo
    return halo.cic_paint_dx(displacements, weight, halo_size, sharding)


def sharded_cic_read_dx(grid_mesh, disp, halo_size, sharding):
    from . import halo
    return halo.cic_read_dx(grid_mesh, disp, halo_size, sharding)


def sharded_cic_paint(grid_mesh, positions, weight, halo_size, sharding):
    raise NotImplementedError(
        "absolute-mode multi-device painting is 'not recommended' and self-inconsistent in the "
        "reference (painting.py:51-55, SURVEY.md §2.2); use cic_paint_dx")


def sharded_cic_read(grid_mesh, positions, halo_size, sharding):
    raise NotImplementedError(
        "absolute-mode multi-device read is self-inconsistent in the reference (SURVEY.md §2.2); "
        "use cic_read_dx")
