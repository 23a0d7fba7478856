"""Multi-GPU force loop over NVLink peer memory: x slabs (pdims = (P, 1)) and pencil grids (px, py), one process per
GPU.  On a pencil grid the particle domain is the pencil, the FFT chain keeps x slabs and the z passes do the
row-group transpose while they load / store (include/jaxpm_b200.h: jpm_slab_create_ex).

What the reference does with three library collectives per force evaluation — halo_exchange after the
paint, [ext] jaxdecomp.pfft3d / pifft3d all-to-alls, halo_exchange before the read
(/root/reference/jaxpm/distributed.py:37-85, painting.py:192-215, :239-260, pm.py:41-56) — is done
here INSIDE the FFT kernels of csrc/pmfft.cu: the z pass folds the neighbours' density ghost planes
while it loads, the y and x passes store their results straight into the buffer of the rank that owns
the row (the transposes), the last pass writes the neighbours' force ghost planes.  Ranks meet in
four in-stream flag barriers per step; no NCCL call, no pack/unpack kernel, no staging copy.

`torch.distributed` is only used once, to gather the 64-byte CUDA IPC handles of the ranks' blocks.
"""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import as_f32, call, ptr, stream

_HANDLE_BYTES = 64


def slab_supported(global_shape, pdims, halo_x, halo_y=None):
    """True when the fused peer-memory path can serve this decomposition: x slabs (P, 1) or a pencil grid (px, py)."""
    nx, ny, nz = (int(s) for s in global_shape)
    px, py = pdims
    P = px * py
    pow2 = lambda n: 16 <= n <= 1024 and (n & (n - 1)) == 0
    if px < 1 or py < 1 or P > 8 or not (pow2(nx) and pow2(ny) and pow2(nz)):
        return False
    if nx % P or ny % P:
        return False
    Lx, Ly = nx // px, ny // py
    if not (1 <= halo_x <= Lx and Lx + 2 * halo_x >= 16):
        return False
    if py > 1:
        hy = halo_x if halo_y is None else halo_y
        if not (1 <= hy <= Ly and Ly % 16 == 0):
            return False
    return True


def pencil_targets(pdims, rank, global_shape, gx, gy, xl, y0, ge, rows=16, G=4):
    """Host restatement of the routing rule of the z passes on a pencil grid (`pencil_targets` in csrc/pmfft.cu), for
    documentation and the CPU tests: which arrays hold plane `xl` of FFT slab `rank`, rows `y0 .. y0 + rows - 1`.

    Returns [(owner_rank, plane_in_its_array, first_array_row, r0, r1)]: tile rows r0 <= r < r1 live in the array of
    `owner_rank` at plane `plane` (0 = its first ghost plane), rows `first_array_row + r`.  The first entry is the
    pencil that owns the rows (all of them); the others are the ghost images held by the x / y / corner neighbours
    (`ge` ghost planes / rows per side in use).  The forward z pass SUMS the entries (halo reduce + row-group
    transpose), the inverse z pass WRITES to all of them (halo fill)."""
    px, py = pdims
    nx, ny = global_shape[:2]
    lx, Lx, Ly = nx // (px * py), nx // px, ny // py
    a, b = divmod(rank, py)
    xq = b * lx + xl
    by, yl0 = divmod(y0, Ly)
    gex, gey = min(gx, ge), min(gy, ge)
    xs = [(a, gx + xq)]
    if xq < gex:
        xs.append(((a - 1) % px, gx + Lx + xq))
    if xq >= Lx - gex:
        xs.append(((a + 1) % px, xq - (Lx - gx)))
    ys = [(by, G + gy + yl0, 0, rows)]
    if yl0 < gey:
        ys.append(((by - 1) % py, G + gy + Ly + yl0, 0, min(rows, gey - yl0)))
    if yl0 + rows > Ly - gey:
        ys.append(((by + 1) % py, G + gy + yl0 - Ly, max(0, Ly - gey - yl0), rows))
    return [(xr * py + yc, xp, yr, r0, r1) for xr, xp in xs for yc, yr, r0, r1 in ys]


class SlabPlan:
    """One rank's jpm_plan of the slab / pencil decomposition (buffers + peer mappings).  `nranks` is the rank count
    of an x-slab grid, or pass `pdims=(px, py)` (rank = a py + b) with the y ghost width `gy` for a pencil grid."""

    def __init__(self, global_shape, nranks, rank, gx, device, pdims=None, gy=0):
        self.global_shape = tuple(int(s) for s in global_shape)
        self.pdims = (int(nranks), 1) if pdims is None else (int(pdims[0]), int(pdims[1]))
        self.nranks, self.rank, self.gx, self.device = self.pdims[0] * self.pdims[1], int(rank), int(gx), torch.device(device)
        self.gy = int(gy) if self.pdims[1] > 1 else 0
        nx, ny, nz = self.global_shape
        self.lx = nx // self.nranks                            # FFT slab of this rank
        self.Lx, self.Ly = nx // self.pdims[0], ny // self.pdims[1]
        self.local_shape = (self.Lx, self.Ly, nz)              # the rank's particle / mesh block
        self.mesh_shape = (self.Lx + 2 * self.gx, self.Ly + 2 * self.gy, nz)     # what the particle kernels see
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            call("jpm_slab_create_ex", C.byref(h), nx, ny, nz, self.pdims[0], self.pdims[1], self.rank, self.gx, self.gy)
        self.handle = h
        self.attached = False

    def ipc_handle(self):
        buf = (C.c_ubyte * _HANDLE_BYTES)()
        call("jpm_slab_ipc_handle", self.handle, buf, _HANDLE_BYTES)
        return bytes(buf)

    def base(self):
        b, n = C.c_void_p(), C.c_int64()
        call("jpm_slab_base", self.handle, C.byref(b), C.byref(n))
        return b.value

    def attach_ipc(self, handles):
        """handles: list of nranks 64-byte strings in rank order (from every rank's ipc_handle())."""
        blob = b"".join(handles)
        assert len(blob) == _HANDLE_BYTES * self.nranks
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        with torch.cuda.device(self.device):
            call("jpm_slab_attach_ipc", self.handle, buf, self.nranks)
        self.attached = True

    def attach_local(self, plans):
        """Peers living in this process (tests: several ranks on one device, or one process driving
        several peer-enabled devices)."""
        arr = (C.c_void_p * self.nranks)(*[p.base() for p in plans])
        with torch.cuda.device(self.device):
            call("jpm_slab_attach_ptrs", self.handle, arr, self.nranks)
        self.attached = True

    def set_density(self, rho_local):
        call("jpm_slab_set_density_f32", self.handle, stream(), ptr(as_f32(rho_local), torch.float32))

    def forces(self, r_split=0.0):
        """COLLECTIVE: every rank calls it on its own stream."""
        call("jpm_slab_forces", self.handle, stream(), float(r_split))

    def interior(self, which):
        out = torch.empty(self.local_shape, dtype=torch.float32, device=self.device)
        call("jpm_slab_get_interior_f32", self.handle, stream(), int(which), ptr(out))
        return out

    def check(self):
        call("jpm_slab_check", self.handle, stream())

    def ghost_width(self):
        v = C.c_int32(0)
        call("jpm_slab_ghost_width", self.handle, stream(), C.byref(v))
        return int(v.value)

    def halo_exceeded(self):
        v = C.c_int32(0)
        call("jpm_slab_halo_exceeded", self.handle, stream(), C.byref(v))
        return bool(v.value)

    def destroy(self):
        if self.handle:
            _lib.load().jpm_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def connect(plan, group=None):
    """Exchange the IPC handles of every rank's block over `torch.distributed` and map the peers."""
    import torch.distributed as dist
    handles = [None] * plan.nranks
    dist.all_gather_object(handles, plan.ipc_handle(), group=group)
    plan.attach_ipc(handles)
    dist.barrier(group=group)
    return plan


class SlabStepper:
    """Per-rank resident state of the slab drift-kick loop: tile-sorted particles (csrc/sim.cu) on the
    rank's ghost-zone mesh + the fused peer-memory FFT chain.  Same contract as halo.ShardedStepper:
    identical to the reference for every particle within its halo reach (|disp_x| < gx)."""

    def __init__(self, disp, vel, gx, nranks, rank, group=None, tile=None, margin=1, plan=None,
                 force_mode="spectral", pdims=None, gy=0):
        d = as_f32(disp)
        Lx, Ly, nz = d.shape[:3]
        self.device = d.device
        self.group = group
        pd = (int(nranks), 1) if pdims is None else (int(pdims[0]), int(pdims[1]))
        self.plan = plan if plan is not None else SlabPlan((Lx * pd[0], Ly * pd[1], nz), nranks, rank, gx, d.device,
                                                           pdims=pd, gy=gy)
        if not self.plan.attached and plan is None:
            connect(self.plan, group)
        ms = self.plan.mesh_shape
        if tile is None:
            tile = 16 if min(ms) >= 64 else 8
        self.sim = ops.Sim(ms, (Lx, Ly, nz), True, d.device, halo=(self.plan.gx, self.plan.gy), tile=tile,
                           margin=margin, plan=self.plan)
        if force_mode != "spectral":
            # "potential": one inverse transform + the gradient pass (psi ghosts over NVLink instead of three force
            # meshes); "auto": per step from the GLOBAL error bound, the same decision on every rank
            self.sim.set_force_mode(force_mode)
        self.sim.load(d, as_f32(vel))

    def load(self, disp, vel):
        self.sim.load(as_f32(disp), as_f32(vel))

    def store(self, disp, vel):
        self.sim.store(disp, vel)
        self.plan.check()
        if self.plan.halo_exceeded():
            import warnings
            warnings.warn(f"rank {self.plan.rank}: particles reached the outermost of the {self.plan.gx} ghost planes - "
                          "halo_size is too small for this displacement field (the reference silently wraps them too)")

    def step(self, kick, drift):
        self.sim.step(kick, drift)

    def step_profile(self, kick, drift):
        return self.sim.step_profile(kick, drift)

    def force_info(self):
        return self.sim.force_info()

    def close(self, barrier=True):
        """Ranks must have finished using each other's memory before any block is freed."""
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if barrier and self.plan.nranks > 1 and dist.is_available() and dist.is_initialized():
            dist.barrier(group=self.group)
        self.sim = None
        self.plan.destroy()
