"""Multi-GPU force loop: halo reduce (paint), halo fill (read) and the sharded pm_forces / LPT /
drift-kick driver.  One process per GPU; every array is the LOCAL block of the calling rank.

Reference protocol (/root/reference/jaxpm/distributed.py:45-113, painting.py:192-215, :239-260,
SURVEY.md §3.3):
  paint: pad the local mesh by h, scatter, halo_exchange(extents h//2) [ext jaxdecomp], then
         slice_unpad adds the received outer bands onto the first/last h//2 interior cells;
  read : pad by h//2, halo_exchange(extents h//2) (ordinary halo fill), gather with the
         Lagrangian sites offset by h//2.
Here the exchange and the add/copy are fused: bands are packed by `jpm_pack_box_f32`, sent to the
+-1 neighbour of each sharded axis with NCCL send/recv (x phase, then y phase over the already
corrected rows so corners propagate exactly like the sequential per-axis exchange), and unpacked
with accumulate (`jpm_unpack_box_f32`).  The 3x-padded mesh of the reference is only materialised
for the paint target; force meshes are padded by h//2.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops
from ._lib import as_f32
from .distributed import get_halo_size


def _sendrecv(sh, axis, to_lo, to_hi):
    """Send `to_lo` to the low neighbour and `to_hi` to the high neighbour along `axis`;
    return (from_lo, from_hi)."""
    lo = sh.neighbor(-1, 0) if axis == 0 else sh.neighbor(0, -1)
    hi = sh.neighbor(+1, 0) if axis == 0 else sh.neighbor(0, +1)
    from_lo, from_hi = torch.empty_like(to_hi), torch.empty_like(to_lo)
    if lo == sh.rank and hi == sh.rank:
        from_lo.copy_(to_hi)
        from_hi.copy_(to_lo)
        return from_lo, from_hi
    reqs = dist.batch_isend_irecv([
        dist.P2POp(dist.isend, to_hi, sh.global_rank(hi), group=sh.group),
        dist.P2POp(dist.irecv, from_lo, sh.global_rank(lo), group=sh.group),
        dist.P2POp(dist.isend, to_lo, sh.global_rank(lo), group=sh.group),
        dist.P2POp(dist.irecv, from_hi, sh.global_rank(hi), group=sh.group),
    ])
    for r in reqs:
        r.wait()
    return from_lo, from_hi


def pad(x, pad_width):
    """slice_pad per shard (distributed.py:88-99): zero-pad axes 0/1."""
    hx, hy = pad_width[0][0], pad_width[1][0]
    out = torch.zeros((x.shape[0] + 2 * hx, x.shape[1] + 2 * hy, x.shape[2]), dtype=torch.float32, device=x.device)
    ops.unpack_box_(out, as_f32(x), hx, hx + x.shape[0], hy, hy + x.shape[1])
    return out


def halo_reduce_(padded, hx, hy, sh):
    """halo_exchange(extents h//2) + slice_unpad (distributed.py:61-85) fused; returns the cropped
    local mesh.  `padded` [lx+2hx, ly+2hy, nz] is modified."""
    S0, S1, nz = padded.shape
    ex, ey = hx // 2, hy // 2
    if ex > 0:
        to_hi = ops.pack_box(padded, S0 - 2 * ex, S0 - ex, 0, S1)
        to_lo = ops.pack_box(padded, ex, 2 * ex, 0, S1)
        from_lo, from_hi = _sendrecv(sh, 0, to_lo, to_hi)
        ops.unpack_box_(padded, from_lo, hx, hx + ex, 0, S1, accumulate=True)
        ops.unpack_box_(padded, from_hi, S0 - hx - ex, S0 - hx, 0, S1, accumulate=True)
    if ey > 0:
        to_hi = ops.pack_box(padded, hx, S0 - hx, S1 - 2 * ey, S1 - ey)
        to_lo = ops.pack_box(padded, hx, S0 - hx, ey, 2 * ey)
        from_lo, from_hi = _sendrecv(sh, 1, to_lo, to_hi)
        ops.unpack_box_(padded, from_lo, hx, S0 - hx, hy, hy + ey, accumulate=True)
        ops.unpack_box_(padded, from_hi, hx, S0 - hx, S1 - hy - ey, S1 - hy, accumulate=True)
    return ops.pack_box(padded, hx, S0 - hx, hy, S1 - hy)


def unpad_reduce(x, pad_width):
    """slice_unpad_impl alone (distributed.py:68-85) on an already exchanged padded block."""
    hx, hy = pad_width[0][0], pad_width[1][0]
    S0, S1, nz = x.shape
    x = x.clone()
    if hx > 0:
        ops.unpack_box_(x, ops.pack_box(x, 0, hx // 2, 0, S1), hx, hx + hx // 2, 0, S1, accumulate=True)
        ops.unpack_box_(x, ops.pack_box(x, S0 - hx // 2, S0, 0, S1), S0 - hx - hx // 2, S0 - hx, 0, S1, accumulate=True)
    if hy > 0:
        ops.unpack_box_(x, ops.pack_box(x, 0, S0, 0, hy // 2), 0, S0, hy, hy + hy // 2, accumulate=True)
        ops.unpack_box_(x, ops.pack_box(x, 0, S0, S1 - hy // 2, S1), 0, S0, S1 - hy - hy // 2, S1 - hy, accumulate=True)
    return ops.pack_box(x, hx, S0 - hx, hy, S1 - hy)


def halo_exchange(x, halo_extents, sh):
    """[ext] jaxdecomp.halo_exchange on a padded block: pads [0,e) / [S-e,S) receive the
    neighbours' [S-2e,S-e) / [e,2e); axis 0 then axis 1."""
    x = as_f32(x).clone()
    S0, S1, nz = x.shape
    ex, ey = halo_extents
    if ex > 0:
        from_lo, from_hi = _sendrecv(sh, 0, ops.pack_box(x, ex, 2 * ex, 0, S1), ops.pack_box(x, S0 - 2 * ex, S0 - ex, 0, S1))
        ops.unpack_box_(x, from_lo, 0, ex, 0, S1)
        ops.unpack_box_(x, from_hi, S0 - ex, S0, 0, S1)
    if ey > 0:
        from_lo, from_hi = _sendrecv(sh, 1, ops.pack_box(x, 0, S0, ey, 2 * ey), ops.pack_box(x, 0, S0, S1 - 2 * ey, S1 - ey))
        ops.unpack_box_(x, from_lo, 0, S0, 0, ey)
        ops.unpack_box_(x, from_hi, 0, S0, S1 - ey, S1)
    return x


def halo_fill(meshes, hx, hy, sh, ex=None, ey=None):
    """Pad each local mesh [lx,ly,nz] of `meshes` [nb,...] by (hx, hy) and fill the innermost
    (ex, ey) <= (hx, hy) cells of the pads from the neighbours (slice_pad + halo_exchange of
    painting.py:248-252; there pad == extent == h//2), all meshes in one message per direction."""
    ex = hx if ex is None else ex
    ey = hy if ey is None else ey
    nb, lx, ly, nz = meshes.shape
    S0, S1 = lx + 2 * hx, ly + 2 * hy
    out = torch.zeros((nb, S0, S1, nz), dtype=torch.float32, device=meshes.device)
    for b in range(nb):
        ops.unpack_box_(out[b], meshes[b], hx, hx + lx, hy, hy + ly)
    if ex > 0:
        to_hi = torch.stack([ops.pack_box(out[b], hx + lx - ex, hx + lx, 0, S1) for b in range(nb)])
        to_lo = torch.stack([ops.pack_box(out[b], hx, hx + ex, 0, S1) for b in range(nb)])
        from_lo, from_hi = _sendrecv(sh, 0, to_lo, to_hi)
        for b in range(nb):
            ops.unpack_box_(out[b], from_lo[b], hx - ex, hx, 0, S1)
            ops.unpack_box_(out[b], from_hi[b], hx + lx, hx + lx + ex, 0, S1)
    if ey > 0:
        to_hi = torch.stack([ops.pack_box(out[b], 0, S0, hy + ly - ey, hy + ly) for b in range(nb)])
        to_lo = torch.stack([ops.pack_box(out[b], 0, S0, hy, hy + ey) for b in range(nb)])
        from_lo, from_hi = _sendrecv(sh, 1, to_lo, to_hi)
        for b in range(nb):
            ops.unpack_box_(out[b], from_lo[b], 0, S0, hy - ey, hy)
            ops.unpack_box_(out[b], from_hi[b], 0, S0, hy + ly, hy + ly + ey)
    return out


def _halos(halo_size, sh):
    padw, _ = get_halo_size(halo_size, sh)
    return padw[0][0], padw[1][0]


def cic_paint_dx(displacements, weight, halo_size, sh):
    """Sharded cic_paint_dx (painting.py:192-215): local displacements -> local mesh block."""
    d = as_f32(displacements)
    hx, hy = _halos(halo_size, sh)
    lx, ly, nz = d.shape[:3]
    padded = torch.zeros((lx + 2 * hx, ly + 2 * hy, nz), dtype=torch.float32, device=d.device)
    ops.cic_paint_dx_(padded, d, weight, (hx, hy))
    return halo_reduce_(padded, hx, hy, sh)


def cic_read_dx(grid_mesh, disp, halo_size, sh):
    """Sharded cic_read_dx (painting.py:239-260)."""
    hx, hy = _halos(halo_size, sh)
    m = halo_fill(as_f32(grid_mesh).unsqueeze(0), hx // 2, hy // 2, sh)[0]
    return ops.cic_read_dx(m, as_f32(disp), (hx // 2, hy // 2))


def _fft(local_shape, sh, device):
    from .pfft import get_pfft
    return get_pfft(sh.global_shape(local_shape), sh, device)


def force_meshes(rho, sh, r_split=0.0, filter_tab=None):
    """Local density block -> the three local force-mesh blocks [3, lx, ly, nz]."""
    fft = _fft(rho.shape, sh, rho.device)
    dk = fft.forward(as_f32(rho).unsqueeze(0))
    out3 = torch.empty((3, *fft.spec_shape), dtype=torch.complex64, device=rho.device)
    tabs = fft.kspace_tables(lambda a: torch.as_tensor(a).to(rho.device))
    fft.backend.kspace(0, dk[0], out3, tabs, fft.spec_shape, fft.axis_map, 1.0 / fft.ncell, r_split, filter_tab)
    return fft.inverse(out3)


def pm_forces(positions, mesh_shape, delta, r_split, relative, halo_size, sh, filter_tab=None):
    """Sharded pm_forces (pm.py:12-58), relative mode only."""
    if not relative:
        raise NotImplementedError("multi-device pm_forces supports paint_absolute_pos=False only "
                                  "(the reference's absolute multi-device mode is self-inconsistent)")
    d = as_f32(positions)
    hx, hy = _halos(halo_size, sh)
    if delta is None:
        rho = cic_paint_dx(d, 1.0, halo_size, sh)
    elif isinstance(delta, torch.Tensor) and delta.is_complex():
        from .pfft import pifft3d
        rho = pifft3d(delta, sh)
    else:
        rho = as_f32(delta)
    f3 = force_meshes(rho, sh, r_split, filter_tab)
    f3p = halo_fill(f3, hx // 2, hy // 2, sh)
    return ops.cic_read3(f3p, d, 1.0, relative=True, halo=(hx // 2, hy // 2))


def lpt2_source(ic, sh):
    """delta2 of pm.py:88-109 for a local block of the linear field."""
    fft = _fft(ic.shape, sh, ic.device)
    dk = fft.forward(as_f32(ic).unsqueeze(0))
    sh6 = torch.empty((6, *fft.spec_shape), dtype=torch.complex64, device=ic.device)
    tabs = fft.kspace_tables(lambda a: torch.as_tensor(a).to(ic.device))
    fft.backend.kspace(1, dk[0], sh6, tabs, fft.spec_shape, fft.axis_map, 1.0 / fft.ncell)
    s6 = fft.inverse(sh6).contiguous()
    out = torch.empty(ic.shape, dtype=torch.float32, device=ic.device)
    from ._lib import call, ptr, stream
    call("jpm_lpt2_source_f32", stream(), ptr(out), ptr(s6), out.numel())
    return out


class ShardedStepper:
    """Per-rank state of the sharded drift-kick loop.

    resident=False: the order-preserving kernels on (disp, vel) every step — paint into the
      h-padded mesh, halo reduce, distributed FFT chain, h//2 halo fill, fused read3+kick+drift.
    resident=True : the tile-sorted state of csrc/sim.cu on the h-padded local mesh; the force
      meshes are padded by h as well (same geometry for paint and read, so one tile sort serves
      both) with the inner h//2 of the pad filled — identical to the reference for every particle
      within its halo reach (max|disp| < h//2, the reference's own validity limit)."""

    def __init__(self, disp, vel, halo_size, sh, resident=True, tile=None, margin=2, halos=None):
        self.sh, self.halo_size = sh, halo_size
        # `halos` overrides the reference's rule "no halo on an axis that is not split" (tests: a (1, 1)
        # process grid where every rank is its own neighbour)
        self.hx, self.hy = halos if halos is not None else _halos(halo_size, sh)
        self.disp, self.vel = disp, vel
        self.lshape = tuple(disp.shape[:3])
        self.resident = resident
        self.sim = None
        if resident:
            lx, ly, nz = self.lshape
            self.pshape = (lx + 2 * self.hx, ly + 2 * self.hy, nz)
            if tile is None:
                tile = 16 if min(self.pshape) >= 64 else 8
            self.sim = ops.Sim(self.pshape, self.lshape, True, disp.device, halo=(self.hx, self.hy), tile=tile,
                               margin=margin, with_plan=False)
            self.sim.load(disp, vel)

    def load(self, disp, vel):
        self.disp, self.vel = disp, vel
        if self.sim is not None:
            self.sim.load(disp, vel)

    def store(self, disp, vel):
        if self.sim is not None:
            self.sim.store(disp, vel)

    def step(self, kick, drift):
        sh, hx, hy = self.sh, self.hx, self.hy
        if self.sim is None:
            lx, ly, nz = self.lshape
            padded = torch.zeros((lx + 2 * hx, ly + 2 * hy, nz), dtype=torch.float32, device=self.disp.device)
            ops.cic_paint_dx_(padded, self.disp, 1.0, (hx, hy))
            rho = halo_reduce_(padded, hx, hy, sh)
            f3p = halo_fill(force_meshes(rho, sh), hx // 2, hy // 2, sh)
            ops.read3_kick_drift_(f3p, self.disp, self.vel, kick, drift, True, halo=(hx // 2, hy // 2))
            return
        padded = torch.zeros(self.pshape, dtype=torch.float32, device=self.disp.device)
        self.sim.paint_(padded)
        rho = halo_reduce_(padded, hx, hy, sh)
        f3p = halo_fill(force_meshes(rho, sh), hx, hy, sh, hx // 2, hy // 2)
        self.sim.read_kick_drift(f3p, kick, drift)

    def timing_summary(self):
        return None

    def close(self):
        self.sim = None


def make_stepper(disp, vel, halo_size, sh, resident=True, tile=None, margin=None, fused=True,
                 force_mode="spectral"):
    """The fastest stepper that serves this decomposition: the peer-memory stepper (slab.py: halo reduce / FFT
    transposes / halo fill inside the FFT kernels, no NCCL on the data path) for slab (P, 1) and pencil (px, py)
    process grids on power-of-two meshes, else the NCCL stepper above."""
    from . import slab
    hx, hy = _halos(halo_size, sh)
    gshape = sh.global_shape(tuple(disp.shape[:3]))
    # the slab stepper's ghost planes play the role of the reference's halo; its validity limit is the
    # reference's (|displacement| < halo // 2 is exchanged, painting.py:192-215).  A grid with px == 1 has no halo in x
    # in the reference (get_halo_size zeroes it): its x ghost planes are the rank's own periodic images, any width does
    if sh.pdims[1] > 1 and sh.pdims[0] == 1 and hx < 1:
        hx = max(hy, 4)
    if fused and resident and slab.slab_supported(gshape, sh.pdims, hx, hy):
        return slab.SlabStepper(disp, vel, hx, sh.size, sh.rank, group=sh.group, tile=tile,
                                margin=1 if margin is None else margin, force_mode=force_mode, pdims=sh.pdims, gy=hy)
    if force_mode != "spectral":
        raise NotImplementedError("force_mode != 'spectral' needs the fused peer-memory path (power-of-two mesh, "
                                  "nx, ny divisible by the rank count, ny / py a multiple of 16)")
    return ShardedStepper(disp, vel, halo_size, sh, resident=resident, tile=tile, margin=2 if margin is None else margin)


def nbody_kick_drift(disp, vel, d, k, mesh_shape, halo_size, sh, callback=None, resident=True, fused=True,
                     force_mode="spectral"):
    """Sharded drift-kick loop (the first drift has already been applied by the caller)."""
    st = make_stepper(disp, vel, halo_size, sh, resident=resident, fused=fused, force_mode=force_mode)
    nsteps = len(k)
    for n in range(nsteps):
        st.step(k[n], d[n + 1] if n + 1 < nsteps else 0.0)
        if callback is not None:
            st.store(disp, vel)
            callback(n, disp, vel)
    st.store(disp, vel)
    st.close()
    return disp, vel


def linear_field(field, mesh_shape, box_size, pk, sh):
    raise NotImplementedError("sharded linear_field: generate the ICs once and feed every rank its "
                              "block (the reference's own sharded RNG differs from the unsharded one, "
                              "distributed.py:204-215)")
