"""Thin functional wrappers over the C ABI (no autograd, no sharding).

Every function takes/returns contiguous float32 (or complex64) CUDA tensors and
enqueues on torch's current stream.  Higher layers (``painting``, ``pm``,
``ode``) mirror the reference's Python API on top of these.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import as_f32, call, ptr, stream

_plans = {}


class Plan:
    """Opaque jpm_plan handle (cuFFT plans, k tables, scratch) for one mesh shape on one device."""

    def __init__(self, shape, device):
        self.shape = tuple(int(s) for s in shape)
        self.device = device
        h = C.c_void_p()
        with torch.cuda.device(device):
            call("jpm_plan_create", C.byref(h), *self.shape)
        self.handle = h
        self.nzh = self.shape[2] // 2 + 1
        self.ncell = self.shape[0] * self.shape[1] * self.shape[2]
        self.spec_shape = (self.shape[0], self.shape[1], self.nzh)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().jpm_plan_destroy(self.handle)
        except Exception:
            pass


def get_plan(shape, device):
    key = (tuple(int(s) for s in shape), torch.device(device).index or 0)
    if key not in _plans:
        _plans[key] = Plan(shape, torch.device("cuda", key[1]))
    return _plans[key]


def clear_plans():
    _force_sims.clear()
    _plans.clear()


_force_sims = {}


def fast_path_shape(mesh_shape):
    """True when the tile kernels + fused FFT chain serve this mesh (powers of two in [16, 1024])."""
    return all(16 <= int(n) <= 1024 and (int(n) & (int(n) - 1)) == 0 for n in mesh_shape)


def pm_forces_tiles(positions, mesh_shape, relative, r_split=0.0, filter_tab=None):
    """jaxpm/pm.py:12-58 on the fast kernels: tile-sort the particles (positions only), shared-memory paint with
    TMA reduce-add, fused FFT chain, shared-memory gather; forces come back in the caller's particle order.
    The positions-only resident state (16 B per particle) is cached per (mesh, particle count, mode, device)."""
    pos = as_f32(positions)
    mesh_shape = tuple(int(n) for n in mesh_shape)
    npart = pos.numel() // 3
    pshape = tuple(pos.shape[:3]) if (relative or pos.dim() == 4) else (1, 1, npart)
    key = (mesh_shape, pshape, bool(relative), pos.device.index or 0)
    sim = _force_sims.get(key)
    if sim is None:
        _force_sims.clear()          # one cached state at a time: it is 16 B per particle
        sim = _force_sims[key] = Sim(mesh_shape, pshape, relative, pos.device, tile=16 if min(mesh_shape) >= 64 else 8,
                                     margin=1, positions_only=True)
    sim.load(pos)
    out = torch.empty(pos.shape, dtype=torch.float32, device=pos.device)
    sim.forces(out, 1.0, r_split, filter_tab)
    return out


def _wargs(weight, n, device):
    """(weight_ptr, weight_scalar, keepalive) from a scalar or per-particle weight."""
    if isinstance(weight, torch.Tensor) and weight.numel() > 1:
        w = as_f32(weight, device).reshape(-1)
        if w.numel() != n:
            raise _lib.JpmError("Weight shape must match particle shape")
        return ptr(w), 1.0, w
    return None, float(weight), None


def cic_paint_(mesh, positions, weight=1.0):
    """In place: mesh += paint(positions).  jaxpm/painting.py:15-45."""
    pos = as_f32(positions, mesh.device).reshape(-1, 3)
    n = pos.shape[0]
    nx, ny, nz = mesh.shape
    pg = (nx, ny, nz) if n == nx * ny * nz else (1, 1, n)
    wp, ws, keep = _wargs(weight, n, mesh.device)
    call("jpm_cic_paint_f32", stream(), ptr(mesh, torch.float32), ptr(pos), wp, ws, n, nx, ny, nz, *pg)
    return mesh


def cic_paint_dx_(mesh, disp, weight=1.0, halo=(0, 0)):
    """In place: padded mesh += paint_dx(disp).  jaxpm/painting.py:161-189."""
    d = as_f32(disp, mesh.device)
    nx, ny, nz = d.shape[:3]
    hx, hy = halo
    assert tuple(mesh.shape) == (nx + 2 * hx, ny + 2 * hy, nz), (mesh.shape, d.shape, halo)
    wp, ws, keep = _wargs(weight, nx * ny * nz, mesh.device)
    call("jpm_cic_paint_dx_f32", stream(), ptr(mesh, torch.float32), ptr(d), wp, ws, nx, ny, nz, hx, hy)
    return mesh


def cic_read(mesh, positions):
    m = as_f32(mesh)
    pos = as_f32(positions, m.device)
    flat = pos.reshape(-1, 3)
    out = torch.empty(flat.shape[0], dtype=torch.float32, device=m.device)
    call("jpm_cic_read_f32", stream(), ptr(out), ptr(m), ptr(flat), flat.shape[0], *m.shape)
    return out.reshape(pos.shape[:-1])


def cic_read_dx(mesh, disp, halo=(0, 0)):
    m = as_f32(mesh)
    d = as_f32(disp, m.device)
    nx, ny, nz = d.shape[:3]
    hx, hy = halo
    assert tuple(m.shape) == (nx + 2 * hx, ny + 2 * hy, nz)
    out = torch.empty((nx, ny, nz), dtype=torch.float32, device=m.device)
    call("jpm_cic_read_dx_f32", stream(), ptr(out), ptr(m), ptr(d), nx, ny, nz, hx, hy)
    return out


def cic_read3(force3, pos_or_disp, scale=1.0, relative=False, halo=(0, 0)):
    """force3: [3,mx,my,mz] -> [...,3] (the 3 reads + stack of jaxpm/pm.py:54-56)."""
    f = as_f32(force3)
    p = as_f32(pos_or_disp, f.device)
    flat = p.reshape(-1, 3)
    out = torch.empty_like(flat)
    mx, my, mz = f.shape[1:]
    call("jpm_cic_read3_f32", stream(), ptr(out), ptr(f[0]), ptr(f[1]), ptr(f[2]), ptr(flat),
         float(scale), flat.shape[0], mx, my, mz, halo[0], halo[1], int(relative))
    return out.reshape(p.shape)


def cic_readgrad(mesh, pos_or_disp, relative=False, halo=(0, 0), want_value=True, want_grad=True,
                 grad_scale=1.0):
    """value = read(mesh, p); grad = grad_scale * d value / d p (grad_scale: scalar or per-particle)."""
    m = as_f32(mesh)
    p = as_f32(pos_or_disp, m.device)
    flat = p.reshape(-1, 3)
    sp, ss, keep = _wargs(grad_scale, flat.shape[0], m.device)
    val = torch.empty(flat.shape[0], dtype=torch.float32, device=m.device) if want_value else None
    grad = torch.empty_like(flat) if want_grad else None
    call("jpm_cic_readgrad_f32", stream(), ptr(val), ptr(grad), ptr(m), ptr(flat), sp, ss,
         flat.shape[0], *m.shape, halo[0], halo[1], int(relative))
    return (val.reshape(p.shape[:-1]) if want_value else None,
            grad.reshape(p.shape) if want_grad else None)


def cic_readgrad3(mesh3, pos_or_disp, cotangent, relative=False, halo=(0, 0), scale=1.0, out=None):
    """sum_d cotangent[..., d] * d read(mesh3[d]) / d x in one pass (the read half of the pm_forces adjoint)."""
    m = as_f32(mesh3)
    p = as_f32(pos_or_disp, m.device).reshape(-1, 3)
    u = as_f32(cotangent, m.device).reshape(-1, 3)
    acc = out is not None
    g = out if acc else torch.empty_like(p)
    call("jpm_cic_readgrad3_f32", stream(), ptr(g, torch.float32), ptr(m[0]), ptr(m[1]), ptr(m[2]), ptr(p), ptr(u),
         float(scale), p.shape[0], *m.shape[1:], halo[0], halo[1], int(relative), int(acc))
    return g


def cic_readgrad1_(grad, mesh, pos_or_disp, relative=False, halo=(0, 0), scale=1.0):
    """grad += scale * d read(mesh) / d x (the paint adjoint), accumulated in place."""
    m = as_f32(mesh)
    p = as_f32(pos_or_disp, m.device).reshape(-1, 3)
    call("jpm_cic_readgrad3_f32", stream(), ptr(grad, torch.float32), ptr(m), None, None, ptr(p), None, float(scale),
         p.shape[0], *m.shape, halo[0], halo[1], int(relative), 1)
    return grad


def cic_paint3_(mesh3, pos_or_disp, weights3, relative=False, halo=(0, 0), scale=1.0):
    """mesh3[d] += paint(x; weight = scale * weights3[..., d]) for the three components in one pass."""
    p = as_f32(pos_or_disp, mesh3.device).reshape(-1, 3)
    w = as_f32(weights3, mesh3.device).reshape(-1, 3)
    call("jpm_cic_paint3_f32", stream(), ptr(mesh3, torch.float32), ptr(p), ptr(w), float(scale), p.shape[0],
         *mesh3.shape[1:], halo[0], halo[1], int(relative))
    return mesh3


def fd_divergence3(g3):
    """sum_d D_d g3[d] with the 4th-order central difference (the real-space form of the gradient kernel)."""
    out = torch.empty(g3.shape[1:], dtype=torch.float32, device=g3.device)
    call("jpm_fd_divergence3_f32", stream(), ptr(out), ptr(g3, torch.float32), *g3.shape[1:])
    return out


def read3_kick_drift_(force3, pos, vel, kick, drift, relative=False, halo=(0, 0), pos_prev=None,
                      vel_prev=None, use_new_vel=True, forces_out=None):
    """Fused read3 + kick + drift.  Kick-drift form is in place on (pos, vel); with
    (pos_prev, vel_prev) the result is written into those buffers (leapfrog-midpoint form)."""
    f = force3
    n = pos.numel() // 3
    po = pos if pos_prev is None else pos_prev
    vo = vel if vel_prev is None else vel_prev
    mx, my, mz = f.shape[1:]
    call("jpm_cic_read3_kick_drift_f32", stream(), ptr(po, torch.float32), ptr(vo, torch.float32),
         ptr(forces_out), ptr(f[0]), ptr(f[1]), ptr(f[2]), ptr(pos, torch.float32),
         ptr(vel, torch.float32), ptr(po), ptr(vo), float(kick), float(drift), int(use_new_vel), n,
         mx, my, mz, halo[0], halo[1], int(relative))
    return po, vo


def cell_index(pos_or_disp, mesh_shape, relative=False, halo=(0, 0)):
    p = as_f32(pos_or_disp).reshape(-1, 3)
    out = torch.empty(p.shape[0], dtype=torch.int32, device=p.device)
    call("jpm_cic_cell_index_i32", stream(), ptr(out), ptr(p), p.shape[0], *mesh_shape, halo[0],
         halo[1], int(relative))
    return out


# ---- FFT / k-space ---------------------------------------------------------------
def rfft3(x, plan=None):
    """Unnormalised forward R2C; returns complex64 [nx,ny,nz/2+1]."""
    x = as_f32(x)
    plan = plan or get_plan(x.shape, x.device)
    out = torch.empty(plan.spec_shape, dtype=torch.complex64, device=x.device)
    call("jpm_fft3d_r2c", plan.handle, stream(), ptr(x), ptr(out))
    return out


def irfft3_(spec, plan, batch=1):
    """UNNORMALISED inverse C2R of `batch` stacked half-spectra; destroys `spec`."""
    shape = plan.shape if batch == 1 else (batch, *plan.shape)
    out = torch.empty(shape, dtype=torch.float32, device=spec.device)
    call("jpm_ifft3d_c2r", plan.handle, stream(), ptr(spec, torch.complex64), ptr(out), batch)
    return out


def _ftab(filter_tab, device):
    if filter_tab is None:
        return None, 0, 0.0, None
    tab, kmax = filter_tab
    t = as_f32(tab, device).reshape(-1)
    return ptr(t), t.numel(), float(kmax), t


def greens_grad(delta_k, plan, norm=None, r_split=0.0, filter_tab=None):
    """delta_k -> 3 force spectra (jaxpm/pm.py:49-56 in one pass)."""
    norm = 1.0 / plan.ncell if norm is None else norm
    out = torch.empty((3, *plan.spec_shape), dtype=torch.complex64, device=delta_k.device)
    fp, nt, km, keep = _ftab(filter_tab, delta_k.device)
    call("jpm_greens_grad_c64", plan.handle, stream(), ptr(delta_k, torch.complex64), ptr(out),
         float(norm), float(r_split), fp, nt, km)
    return out


def greens_div(spec3, plan, norm=None, r_split=0.0, filter_tab=None):
    """Transpose of greens_grad: 3 spectra -> 1."""
    norm = 1.0 / plan.ncell if norm is None else norm
    out = torch.empty(plan.spec_shape, dtype=torch.complex64, device=spec3.device)
    fp, nt, km, keep = _ftab(filter_tab, spec3.device)
    call("jpm_greens_div_c64", plan.handle, stream(), ptr(spec3, torch.complex64), ptr(out),
         float(norm), float(r_split), fp, nt, km)
    return out


def force_meshes_from_spectrum(delta_k, plan, r_split=0.0, filter_tab=None):
    """[3,nx,ny,nz] force meshes from a half-spectrum."""
    return irfft3_(greens_grad(delta_k, plan, None, r_split, filter_tab), plan, 3)


def force_meshes_from_density(density, plan, r_split=0.0, filter_tab=None):
    d = as_f32(density)
    out = torch.empty((3, *plan.shape), dtype=torch.float32, device=d.device)
    fp, nt, km, keep = _ftab(filter_tab, d.device)
    call("jpm_density_to_force_meshes", plan.handle, stream(), ptr(d), ptr(out), float(r_split), fp,
         nt, km)
    return out


def force_meshes_from_density_fused(density, plan, r_split=0.0, filter_tab=None):
    """Same as force_meshes_from_density through the fused FFT chain on the ghost-zone meshes."""
    d = as_f32(density)
    out = torch.empty((3, *plan.shape), dtype=torch.float32, device=d.device)
    fp, nt, km, keep = _ftab(filter_tab, d.device)
    call("jpm_density_to_force_meshes_fused", plan.handle, stream(), ptr(d), ptr(out), float(r_split), fp,
         nt, km)
    return out


def potential_from_density_fused(density, plan, r_split=0.0, filter_tab=None):
    """psi = IFFT(G delta_k / k^2) = -phi through the one-inverse-transform chain (csrc/pmfft.cu)."""
    d = as_f32(density)
    out = torch.empty(plan.shape, dtype=torch.float32, device=d.device)
    fp, nt, km, keep = _ftab(filter_tab, d.device)
    call("jpm_density_to_potential_fused", plan.handle, stream(), ptr(d), ptr(out), float(r_split), fp, nt, km)
    return out


def lpt2_source(delta_k, plan, return_shear=False):
    """delta2 of jaxpm/pm.py:88-109 from the first-order spectrum (optionally also the 6 shear meshes)."""
    sh = torch.empty((6, *plan.spec_shape), dtype=torch.complex64, device=delta_k.device)
    call("jpm_lpt2_shear_c64", plan.handle, stream(), ptr(delta_k, torch.complex64), ptr(sh),
         1.0 / plan.ncell)
    s6 = torch.empty((6, *plan.shape), dtype=torch.float32, device=delta_k.device)
    for b in range(2):
        call("jpm_ifft3d_c2r", plan.handle, stream(), ptr(sh[3 * b:3 * b + 3]), ptr(s6[3 * b:3 * b + 3]), 3)
    out = torch.empty(plan.shape, dtype=torch.float32, device=delta_k.device)
    call("jpm_lpt2_source_f32", stream(), ptr(out), ptr(s6), plan.ncell)
    return (out, s6) if return_shear else out


def lpt2_source_vjp(s6, cot, plan):
    """Reverse mode of lpt2_source with respect to the linear field: sum_q L_q(cot * d delta2 / d s_q) with the
    self-adjoint shear operators L_q = C2R diag(a_i a_j / k^2 / Nc) R2C."""
    cot = as_f32(cot)
    t6 = torch.empty_like(s6)
    call("jpm_lpt2_source_adj_f32", stream(), ptr(t6), ptr(s6), ptr(cot), plan.ncell)
    tk = torch.empty((6, *plan.spec_shape), dtype=torch.complex64, device=cot.device)
    for q in range(6):
        call("jpm_fft3d_r2c", plan.handle, stream(), ptr(t6[q]), ptr(tk[q]))
    acc = torch.empty(plan.spec_shape, dtype=torch.complex64, device=cot.device)
    call("jpm_lpt2_shear_adj_c64", plan.handle, stream(), ptr(tk), ptr(acc), 1.0 / plan.ncell)
    return irfft3_(acc, plan, 1)


def kfilter_logtab(spec, plan, tab, log10_kmin, log10_kmax, kscale, norm=1.0):
    t = as_f32(tab, spec.device).reshape(-1)
    out = torch.empty_like(spec)
    call("jpm_kfilter_logtab_c64", plan.handle, stream(), ptr(spec, torch.complex64), ptr(out), ptr(t),
         t.numel(), float(log10_kmin), float(log10_kmax), float(kscale[0]), float(kscale[1]),
         float(kscale[2]), float(norm))
    return out


def linear_field_fused(white, plan, tab, log10_kmin, log10_kmax, kscale, dc_amp):
    """pm.py:134-143 on the fused FFT chain: IFFT(FFT(white) * tab(|k_phys|)) / Nc, k = 0 times dc_amp."""
    w = as_f32(white)
    t = as_f32(tab, w.device).reshape(-1)
    out = torch.empty(plan.shape, dtype=torch.float32, device=w.device)
    call("jpm_linear_field_f32", plan.handle, stream(), ptr(w), ptr(out), ptr(t), t.numel(), float(log10_kmin),
         float(log10_kmax), float(kscale[0]), float(kscale[1]), float(kscale[2]), float(dc_amp))
    return out


def axpby(a, x, b=0.0, y=None, out=None):
    x = as_f32(x)
    y = None if y is None else as_f32(y, x.device)
    out = torch.empty_like(x) if out is None else out
    call("jpm_axpby_f32", stream(), ptr(out), float(a), ptr(x), float(b), ptr(y), x.numel())
    return out


def grid_plus_disp(disp, offset=(0, 0)):
    d = as_f32(disp)
    out = torch.empty_like(d)
    nx, ny, nz = d.shape[:3]
    call("jpm_grid_plus_disp_f32", stream(), ptr(out), ptr(d), nx, ny, nz, offset[0], offset[1])
    return out


def pm_step_(plan, pos, vel, kick, drift, relative=False):
    call("jpm_pm_step_f32", plan.handle, stream(), ptr(pos, torch.float32), ptr(vel, torch.float32),
         float(kick), float(drift), int(relative))


def pm_step_host_(plan, pos_host, vel_host, pos_dev, vel_dev, kick, drift, relative=False):
    """Host-buffer entry (pinned CPU tensors in/out); synchronises the stream."""
    assert not pos_host.is_cuda and not vel_host.is_cuda
    call("jpm_pm_step_host_f32", plan.handle, stream(), pos_host.data_ptr(), vel_host.data_ptr(),
         ptr(pos_dev, torch.float32), ptr(vel_dev, torch.float32), float(kick), float(drift),
         int(relative))


def pack_box(mesh, x0, x1, y0, y1):
    nx, ny, nz = mesh.shape
    out = torch.empty((x1 - x0, y1 - y0, nz), dtype=torch.float32, device=mesh.device)
    call("jpm_pack_box_f32", stream(), ptr(out), ptr(mesh, torch.float32), ny, nz, x0, x1, y0, y1)
    return out


def unpack_box_(mesh, packed, x0, x1, y0, y1, accumulate=False):
    nx, ny, nz = mesh.shape
    assert tuple(packed.shape) == (x1 - x0, y1 - y0, nz)
    call("jpm_unpack_box_f32", stream(), ptr(mesh, torch.float32), ptr(packed, torch.float32), ny, nz,
         x0, x1, y0, y1, int(accumulate))
    return mesh


class Sim:
    """Tile-sorted resident particle state (jpm_sim): load once, step many times, store back."""

    def __init__(self, mesh_shape, particle_shape, relative, device, halo=(0, 0), tile=None, margin=2,
                 with_plan=True, plan=None, positions_only=False):
        self.mesh_shape = tuple(int(s) for s in mesh_shape)
        self.pshape = tuple(int(s) for s in particle_shape)
        self.relative, self.device, self.halo = bool(relative), device, halo
        if tile is None:
            tile = 16 if min(self.mesh_shape) >= 64 else 8
        self.tile, self.margin = tile, margin
        # `plan`: any object with a jpm_plan `.handle` (ops.Plan, slab.SlabPlan); kept alive by the sim
        self.plan = plan if plan is not None else (get_plan(self.mesh_shape, device) if with_plan else None)
        h = C.c_void_p()
        with torch.cuda.device(device):
            call("jpm_sim_create_ex", C.byref(h), self.plan.handle if self.plan else None, *self.mesh_shape,
                 *self.pshape, halo[0], halo[1], int(self.relative), tile, margin, 1 if positions_only else 0)
        self.handle = h
        self.positions_only = bool(positions_only)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().jpm_sim_destroy(self.handle)
        except Exception:
            pass

    def load(self, pos, vel=None):
        call("jpm_sim_load", self.handle, stream(), ptr(pos, torch.float32), ptr(vel, torch.float32))

    def forces(self, out=None, scale=1.0, r_split=0.0, filter_tab=None):
        """pm_forces of the loaded state on the tile kernels -> [np, 3] in the caller's particle order."""
        npart = self.pshape[0] * self.pshape[1] * self.pshape[2]
        if out is None:
            out = torch.empty((npart, 3), dtype=torch.float32, device=self.device)
        fp, nt, km, keep = _ftab(filter_tab, self.device)
        call("jpm_sim_forces", self.handle, stream(), ptr(out, torch.float32), float(scale), float(r_split), fp, nt, km)
        return out

    def store(self, pos, vel):
        call("jpm_sim_store", self.handle, stream(), ptr(pos, torch.float32), ptr(vel, torch.float32))

    def paint_(self, mesh):
        assert tuple(mesh.shape) == self.mesh_shape
        call("jpm_sim_paint", self.handle, stream(), ptr(mesh, torch.float32))
        return mesh

    def read_kick_drift(self, force3, kick, drift):
        call("jpm_sim_read_kick_drift", self.handle, stream(), ptr(force3[0]), ptr(force3[1]),
             ptr(force3[2]), float(kick), float(drift))

    def step(self, kick, drift):
        call("jpm_sim_step", self.handle, stream(), float(kick), float(drift))

    def step_host(self, pos_host, vel_host, pos_dev, vel_dev, kick, drift):
        """Host-buffer entry (pinned CPU tensors in / out) on the tile kernels; synchronises the stream."""
        assert not pos_host.is_cuda and not vel_host.is_cuda
        call("jpm_sim_step_host_f32", self.handle, stream(), pos_host.data_ptr(), vel_host.data_ptr(),
             ptr(pos_dev, torch.float32), ptr(vel_dev, torch.float32), float(kick), float(drift))

    def steps_host(self, pos_hosts, vel_hosts, kicks, drifts):
        """A batch of independent host-resident states, one step each, with upload / compute / download of consecutive
        elements overlapped (jpm_sim_steps_host_f32).  pos_hosts / vel_hosts: lists of pinned CPU tensors, updated in place."""
        n = len(pos_hosts)
        assert len(vel_hosts) == n and len(kicks) == n and len(drifts) == n
        assert all(not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in (*pos_hosts, *vel_hosts))
        pp = (C.c_void_p * n)(*[t.data_ptr() for t in pos_hosts])
        vv = (C.c_void_p * n)(*[t.data_ptr() for t in vel_hosts])
        kk = (C.c_float * n)(*[float(x) for x in kicks])
        dd = (C.c_float * n)(*[float(x) for x in drifts])
        call("jpm_sim_steps_host_f32", self.handle, stream(), n, pp, vv, kk, dd)

    def step_profile(self, kick, drift):
        """One step with per-stage CUDA-event timing: [(stage name, milliseconds), ...]."""
        names = (C.c_char_p * 24)()
        ms = (C.c_float * 24)()
        n = C.c_int32(0)
        call("jpm_sim_step_profile", self.handle, stream(), float(kick), float(drift), names, ms, 24, C.byref(n))
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    FORCE_MODES = {"spectral": 0, "potential": 1, "auto": 2}

    def set_force_mode(self, mode):
        """'spectral' (three inverse transforms), 'potential' (one + difference stencil) or 'auto'."""
        call("jpm_sim_set_force_mode", self.handle, int(self.FORCE_MODES.get(mode, mode)))

    def force_info(self):
        out = (C.c_double * 6)()
        call("jpm_sim_force_info", self.handle, stream(), out)
        names = {v: k for k, v in self.FORCE_MODES.items()}
        return {"mode": names[int(out[0])], "next": names[int(out[1])], "error_bound": float(out[2]),
                "steps_spectral": int(out[3]), "steps_potential": int(out[4]), "potential_available": bool(out[5])}

    def fallback_counts(self):
        out = (C.c_int64 * 4)()
        call("jpm_sim_stats_host", self.handle, stream(), out)
        return tuple(int(v) for v in out)
