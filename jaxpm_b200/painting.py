"""CIC paint / read — same names, argument order and defaults as
/root/reference/jaxpm/painting.py (cic_paint :48-75, cic_read :109-128,
cic_paint_dx :192-215, cic_read_dx :239-260), backed by the sm_100a kernels.

Arrays are float32 CUDA tensors - or float64 ones for the four primitives below, which then compute in double like
the reference under jax_enable_x64 (no gradient rules, single device) - (torch is the device-memory provider here; in a
JAX deployment the same C-ABI entry points are bound through XLA FFI, see
INTEGRATION.md).  Gradients are provided the way `jax.grad` of the reference
would produce them (paint^T = read, read^T = paint, plus the position
gradients of the CIC weights with sign(0) = 0) via hand-written adjoint kernels.

`sharding` is a `jaxpm_b200.distributed.Sharding` (one process per GPU); with
`sharding=None` everything is single-device, exactly like the reference.
"""
import torch

from . import ops
from ._lib import as_f32


def _is_scalar(w):
    return not (isinstance(w, torch.Tensor) and w.numel() > 1)


class _CicPaint(torch.autograd.Function):
    """mesh + paint(positions, weight).  VJP: (g, w * d read(g)/d pos, read(g))."""

    @staticmethod
    def forward(ctx, grid_mesh, positions, weight):
        out = grid_mesh.clone()
        ops.cic_paint_(out, positions, weight)
        ctx.save_for_backward(positions, weight if isinstance(weight, torch.Tensor) else None)
        ctx.wscalar = None if isinstance(weight, torch.Tensor) and weight.numel() > 1 else float(weight)
        return out

    @staticmethod
    def backward(ctx, g):
        positions, weight = ctx.saved_tensors
        g = g.contiguous()
        need_pos, need_w = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gpos = gw = None
        if need_pos or need_w:
            val, gpos = ops.cic_readgrad(g, positions, want_value=need_w, want_grad=need_pos,
                                         grad_scale=ctx.wscalar if ctx.wscalar is not None else weight)
            if need_w:
                gw = val.reshape(weight.shape) if ctx.wscalar is None else val.sum().reshape(weight.shape)
        return g, gpos, gw


class _CicRead(torch.autograd.Function):
    """read(mesh, positions).  VJP: (paint(weight=u), u * d read(mesh)/d pos)."""

    @staticmethod
    def forward(ctx, grid_mesh, positions):
        ctx.save_for_backward(grid_mesh, positions)
        return ops.cic_read(grid_mesh, positions)

    @staticmethod
    def backward(ctx, u):
        mesh, positions = ctx.saved_tensors
        u = u.contiguous()
        gmesh = gpos = None
        if ctx.needs_input_grad[0]:
            gmesh = ops.cic_paint_(torch.zeros_like(mesh), positions, u)
        if ctx.needs_input_grad[1]:
            _, gpos = ops.cic_readgrad(mesh, positions, want_value=False, grad_scale=u)
        return gmesh, gpos


class _CicPaintDx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, displacements, weight, halo):
        nx, ny, nz = displacements.shape[:3]
        mesh = torch.zeros((nx + 2 * halo[0], ny + 2 * halo[1], nz), dtype=torch.float32,
                           device=displacements.device)
        ops.cic_paint_dx_(mesh, displacements, weight, halo)
        ctx.save_for_backward(displacements, weight if isinstance(weight, torch.Tensor) else None)
        ctx.wscalar = None if isinstance(weight, torch.Tensor) and weight.numel() > 1 else float(weight)
        ctx.halo = halo
        return mesh

    @staticmethod
    def backward(ctx, g):
        disp, weight = ctx.saved_tensors
        g = g.contiguous()
        need_d, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gd = gw = None
        if need_d or need_w:
            val, gd = ops.cic_readgrad(g, disp, relative=True, halo=ctx.halo, want_value=need_w,
                                       want_grad=need_d,
                                       grad_scale=ctx.wscalar if ctx.wscalar is not None else weight)
            if need_w:
                gw = val.reshape(weight.shape) if ctx.wscalar is None else val.sum().reshape(weight.shape)
        return gd, gw, None


class _CicReadDx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, padded_mesh, disp, halo):
        ctx.save_for_backward(padded_mesh, disp)
        ctx.halo = halo
        return ops.cic_read_dx(padded_mesh, disp, halo)

    @staticmethod
    def backward(ctx, u):
        mesh, disp = ctx.saved_tensors
        u = u.contiguous()
        gmesh = gd = None
        if ctx.needs_input_grad[0]:
            gmesh = ops.cic_paint_dx_(torch.zeros_like(mesh), disp, u, ctx.halo)
        if ctx.needs_input_grad[1]:
            _, gd = ops.cic_readgrad(mesh, disp, relative=True, halo=ctx.halo, want_value=False,
                                     grad_scale=u)
        return gmesh, gd, None


def _is_f64(t):
    return isinstance(t, torch.Tensor) and t.dtype == torch.float64


def _f64_weight(weight, n, device):
    """(per-particle float64 array or None, scalar) for the float64 entries."""
    if _is_scalar(weight):
        return None, float(weight)
    w = weight.to(device=device, dtype=torch.float64).contiguous().reshape(-1)
    if w.numel() != n:
        raise ValueError("Weight shape must match particle shape")
    return w, 1.0


def _f64_guard(*tensors, sharding=None):
    if sharding is not None and sharding.size > 1:
        raise NotImplementedError("float64 painting is single-device (the sharded paths compute in float32)")
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise NotImplementedError("float64 painting carries no gradient rules; use float32 inputs for reverse / forward mode")


def _prep_weight(weight, device):
    if _is_scalar(weight):
        return weight if isinstance(weight, torch.Tensor) else float(weight)
    return as_f32(weight, device)


def cic_paint(grid_mesh, positions, weight=1., halo_size=0, sharding=None):
    """Paints positions onto mesh (accumulating into a copy of `grid_mesh`).
    mesh: [nx, ny, nz]; positions: [..., 3] in cell units (reshaped to (nx,ny,nz,3), painting.py:57)."""
    if _is_f64(positions):
        # x64 mode of the reference (jax_enable_x64, tests/test_distributed_pm.py:30): rules and accumulation in double
        from ._lib import call, ptr, stream
        _f64_guard(grid_mesh, positions, weight, sharding=sharding)
        out = grid_mesh.to(device=positions.device, dtype=torch.float64).clone().contiguous()
        pos = positions.contiguous().reshape(-1, 3)
        w, ws = _f64_weight(weight, pos.shape[0], pos.device)
        call("jpm_cic_paint_f64", stream(), ptr(out), ptr(pos), ptr(w), ws, pos.shape[0], *out.shape)
        return out
    grid_mesh = as_f32(grid_mesh)
    positions = as_f32(positions, grid_mesh.device)
    if sharding is not None and sharding.size > 1:
        print("""
            WARNING : absolute painting is not recommended in multi-device mode.
            Please use relative painting instead.
            """)
        from . import distributed
        return distributed.sharded_cic_paint(grid_mesh, positions, weight, halo_size, sharding)
    positions = positions.reshape((*grid_mesh.shape, 3)) if positions.numel() == 3 * grid_mesh.numel() \
        else positions
    return _CicPaint.apply(grid_mesh, positions, _prep_weight(weight, grid_mesh.device))


def cic_read(grid_mesh, positions, halo_size=0, sharding=None):
    """Reads the mesh at `positions`; returns positions.shape[:-1] (painting.py:128)."""
    if _is_f64(positions):
        from ._lib import call, ptr, stream
        _f64_guard(grid_mesh, positions, sharding=sharding)
        mesh = grid_mesh.to(device=positions.device, dtype=torch.float64).contiguous()
        pos = positions.contiguous().reshape(-1, 3)
        out = torch.empty(pos.shape[0], dtype=torch.float64, device=pos.device)
        call("jpm_cic_read_f64", stream(), ptr(out), ptr(mesh), ptr(pos), pos.shape[0], *mesh.shape)
        return out.reshape(positions.shape[:-1])
    grid_mesh = as_f32(grid_mesh)
    positions = as_f32(positions, grid_mesh.device)
    if sharding is not None and sharding.size > 1:
        from . import distributed
        return distributed.sharded_cic_read(grid_mesh, positions, halo_size, sharding)
    return _CicRead.apply(grid_mesh, positions)


def cic_paint_dx(displacements, halo_size=0, sharding=None, weight=1.0, chunk_size=2**24):
    """Relative-mode paint: particle (i,j,k) sits at (i,j,k)+displacements[i,j,k].
    `chunk_size` is accepted for signature compatibility (the reference scans over
    chunks of 2**24 particles, painting_utils.py:115-131; the kernel needs no chunking)."""
    if _is_f64(displacements):
        from ._lib import call, ptr, stream
        _f64_guard(displacements, weight, sharding=sharding)
        d = displacements.contiguous()
        nx, ny, nz = d.shape[:3]
        if not _is_scalar(weight) and tuple(weight.shape) != tuple(d.shape[:-1]):
            raise ValueError("Weight shape must match particle shape")
        w, ws = _f64_weight(weight, nx * ny * nz, d.device)
        mesh = torch.zeros((nx, ny, nz), dtype=torch.float64, device=d.device)
        call("jpm_cic_paint_dx_f64", stream(), ptr(mesh), ptr(d), ptr(w), ws, nx, ny, nz, 0, 0)
        return mesh
    displacements = as_f32(displacements)
    weight = _prep_weight(weight, displacements.device)
    if not _is_scalar(weight) and tuple(weight.shape) != tuple(displacements.shape[:-1]):
        raise ValueError("Weight shape must match particle shape")
    if sharding is not None and sharding.size > 1:
        from . import distributed
        return distributed.sharded_cic_paint_dx(displacements, weight, halo_size, sharding)
    return _CicPaintDx.apply(displacements, weight, (0, 0))


def cic_read_dx(grid_mesh, disp, halo_size=0, sharding=None):
    if _is_f64(disp):
        from ._lib import call, ptr, stream
        _f64_guard(grid_mesh, disp, sharding=sharding)
        d = disp.contiguous()
        mesh = grid_mesh.to(device=d.device, dtype=torch.float64).contiguous()
        nx, ny, nz = d.shape[:3]
        out = torch.empty((nx, ny, nz), dtype=torch.float64, device=d.device)
        call("jpm_cic_read_dx_f64", stream(), ptr(out), ptr(mesh), ptr(d), nx, ny, nz, 0, 0)
        return out
    grid_mesh = as_f32(grid_mesh)
    disp = as_f32(disp, grid_mesh.device)
    if sharding is not None and sharding.size > 1:
        from . import distributed
        return distributed.sharded_cic_read_dx(grid_mesh, disp, halo_size, sharding)
    return _CicReadDx.apply(grid_mesh, disp, (0, 0))


class _CompensateCic(torch.autograd.Function):
    """R2C -> separable sinc^-2 filter -> C2R; a real symmetric linear operator, so it is its own adjoint."""

    @staticmethod
    def forward(ctx, field):
        import numpy as np
        from ._lib import call, ptr, stream
        plan = ops.get_plan(tuple(field.shape), field.device)
        nx, ny, nz = plan.shape
        key = ("cic_comp", plan.shape, field.device.index or 0)
        tabs = _tables.get(key)
        if tabs is None:
            # kernels.py:133-135: kwts_d = sinc(k_d / 2 pi) in float32 with k_d = 2 pi fftfreq(n_d); wts = (prod kwts)^-2
            mk = lambda n, nh: torch.as_tensor((np.sinc(np.fft.fftfreq(n)[:nh].astype(np.float32)).astype(np.float32))**-2.0,
                                               dtype=torch.float32).to(field.device)
            tabs = _tables[key] = (mk(nx, nx), mk(ny, ny), mk(nz, nz // 2 + 1))
        spec = ops.rfft3(field, plan)
        call("jpm_kseparable_c64", stream(), ptr(spec), ptr(spec), ptr(tabs[0]), ptr(tabs[1]), ptr(tabs[2]), nx, ny, nz,
             1.0 / plan.ncell)
        return ops.irfft3_(spec, plan, batch=1)

    @staticmethod
    def backward(ctx, g):
        return _CompensateCic.apply(g.contiguous())


_tables = {}


def compensate_cic(field):
    """Compensate for CIC painting (jaxpm/painting.py:263-275): divide the spectrum by the CIC window
    prod_d sinc(k_d / 2 pi)^2 (kernels.py:118-136) and transform back."""
    return _CompensateCic.apply(as_f32(field))


def cic_paint_2d(mesh, positions, weight):
    """Paints positions onto a 2-D mesh (jaxpm/painting.py:131-158): mesh [nx, ny], positions [npart, 2],
    weight [npart] or None.  Returns the updated mesh (a new tensor, like the functional reference).
    Forward only: the reference's lensing planes (lensing.py:11-44) are not differentiated through here."""
    from ._lib import call, ptr, stream
    out = as_f32(mesh).clone()
    pos = as_f32(positions, out.device).reshape(-1, 2)
    w = None if weight is None else as_f32(weight, out.device).reshape(-1)
    if w is not None and w.numel() != pos.shape[0]:
        raise ValueError("Weight shape must match particle shape")
    call("jpm_cic_paint_2d_f32", stream(), ptr(out), ptr(pos), ptr(w), pos.shape[0], out.shape[0], out.shape[1])
    return out
