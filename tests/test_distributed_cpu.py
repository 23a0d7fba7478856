"""World-size-2/4 gloo tests (CPU) of the multi-GPU host logic: the all-to-all schedule of the
pencil FFT, the halo reduce / fill protocol and the particle->rank rule.  The per-rank compute
kernels are CUDA-only (no CPU fallback in the product), so the TESTS inject NumPy stand-ins for
them (oracle functions) — what is exercised here is the communication schedule and index plans."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import distributed as OD
from oracle import painting as OP


class NumpyBackend:
    """CPU stand-in for jaxpm_b200.pfft.CudaBackend (same interface)."""

    def empty_c(self, n):
        return torch.zeros(n, dtype=torch.complex64)

    def empty_r(self, shape):
        return torch.zeros(shape, dtype=torch.float32)

    def rfft_z(self, x, out):
        out.copy_(torch.from_numpy(np.fft.rfft(x.numpy().astype(np.float64), axis=-1).astype(np.complex64)).reshape(-1))

    def irfft_z(self, spec, out):
        n = out.shape[-1]
        s = spec.numpy().reshape(*out.shape[:-1], n // 2 + 1).astype(np.complex128)
        out.copy_(torch.from_numpy((np.fft.irfft(s, n=n, axis=-1) * n).astype(np.float32)))

    def cfft(self, buf, n, inverse):
        a = buf.numpy().reshape(-1, n).astype(np.complex128)
        r = np.fft.ifft(a, axis=-1) * n if inverse else np.fft.fft(a, axis=-1)
        buf.copy_(torch.from_numpy(r.astype(np.complex64)).reshape(-1))

    def copy2d(self, dst, doff, src, soff, nrows, ncols, srs, drs):
        d, s = dst.numpy(), src.numpy()
        for r in range(nrows):
            d[doff + r * drs:doff + r * drs + ncols] = s[soff + r * srs:soff + r * srs + ncols]

    def transpose(self, dst, doff, src, soff, ni, nj, nb, ssi, ssb, dsj, dsb):
        d, s = dst.numpy(), src.numpy()
        i, j = np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")
        for b in range(nb):
            d[doff + b * dsb + j * dsj + i] = s[soff + b * ssb + i * ssi + j]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _block(x, sh, lx, ly):
    return x[sh.rx * lx:(sh.rx + 1) * lx, sh.ry * ly:(sh.ry + 1) * ly]


def _pfft_worker(rank, world, pdims, shape):
    from jaxpm_b200.distributed import Sharding
    from jaxpm_b200.pfft import PencilFFT
    sh = Sharding(pdims)
    nx, ny, nz = shape
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, *shape)).astype(np.float32)            # two fields (batched)
    fft = PencilFFT(shape, sh, backend=NumpyBackend())
    loc = np.ascontiguousarray(np.stack([_block(f, sh, fft.lx, fft.ly) for f in x]))
    spec = fft.forward(torch.from_numpy(loc))
    ref = np.fft.rfftn(x.astype(np.float64), axes=(1, 2, 3))                # [2, nx, ny, nzh]
    mine = ref[:, :, fft.yoff:fft.yoff + fft.ny2, fft.zoff:fft.zoff + fft.nzl].transpose(0, 2, 3, 1)
    assert spec.shape == (2, fft.ny2, fft.nzl, nx)
    err = np.abs(spec.numpy() - mine).max() / np.abs(ref).max()
    assert err < 1e-5, err
    back = fft.inverse(spec.clone()).numpy() / fft.ncell
    assert np.abs(back - loc).max() < 1e-4
    # every z-mode and every y index is owned exactly once across the groups
    owned = torch.tensor([fft.ny2 * fft.nzl * nx], dtype=torch.int64)
    dist.all_reduce(owned)
    assert int(owned) == nx * ny * (nz // 2 + 1)
    # local k tables line up with the global ones
    tabs = fft.kspace_tables(lambda a: a)
    from jaxpm_b200.pfft import kspace_tables_1d
    np.testing.assert_array_equal(tabs[0], kspace_tables_1d(ny)[0][fft.yoff:fft.yoff + fft.ny2])
    np.testing.assert_array_equal(tabs[1], kspace_tables_1d(nz, nz // 2 + 1)[0][fft.zoff:fft.zoff + fft.nzl])


@pytest.mark.parametrize("pdims,shape", [((2, 1), (8, 8, 6)), ((1, 2), (8, 12, 10)), ((2, 2), (8, 8, 12))])
def test_pencil_fft_schedule(pdims, shape):
    _run(_pfft_worker, pdims[0] * pdims[1], pdims, shape)


def _patch_cpu_kernels():
    """NumPy stand-ins for the CUDA-only per-rank kernels (test infrastructure only)."""
    from jaxpm_b200 import halo, ops

    def pack_box(mesh, x0, x1, y0, y1):
        return mesh[x0:x1, y0:y1].contiguous().clone()

    def unpack_box_(mesh, packed, x0, x1, y0, y1, accumulate=False):
        if accumulate:
            mesh[x0:x1, y0:y1] += packed
        else:
            mesh[x0:x1, y0:y1] = packed
        return mesh

    def cic_paint_dx_(mesh, disp, weight=1.0, halo=(0, 0)):
        mesh += torch.from_numpy(OP.cic_paint_dx_padded(disp.numpy(), weight, halo))
        return mesh

    def cic_read_dx(mesh, disp, halo=(0, 0)):
        return torch.from_numpy(OP.cic_read_dx_padded(mesh.numpy(), disp.numpy(), halo))

    ops.pack_box, ops.unpack_box_ = pack_box, unpack_box_
    ops.cic_paint_dx_, ops.cic_read_dx = cic_paint_dx_, cic_read_dx
    halo.as_f32 = lambda x, device=None: x.to(torch.float32).contiguous()


def _halo_worker(rank, world, pdims, shape, h):
    from jaxpm_b200 import halo
    from jaxpm_b200.distributed import Sharding, get_halo_size, get_local_shape
    _patch_cpu_kernels()
    sh = Sharding(pdims)
    assert (sh.rx, sh.ry) == OD.owner_rank(sh.rx * (shape[0] // pdims[0]), sh.ry * (shape[1] // pdims[1]), shape, pdims)
    assert get_halo_size(h, sh) == OD.get_halo_size(h, pdims)
    lx, ly, _ = get_local_shape(shape, sh)
    rng = np.random.default_rng(0)
    disp = np.clip(rng.standard_normal((*shape, 3)), -h / 2 + 0.1, h / 2 - 0.1).astype(np.float32)
    mesh = rng.standard_normal(shape).astype(np.float32)
    dloc = torch.from_numpy(np.ascontiguousarray(_block(disp, sh, lx, ly)))
    got = halo.cic_paint_dx(dloc, 1.0, h, sh).numpy()
    ref, _ = OD.cic_paint_dx(disp, h, pdims)
    assert np.abs(got - _block(ref, sh, lx, ly)).max() < 1e-5
    mloc = torch.from_numpy(np.ascontiguousarray(_block(mesh, sh, lx, ly)))
    got = halo.cic_read_dx(mloc, dloc, h, sh).numpy()
    ref = OD.cic_read_dx(mesh, disp, h, pdims)
    assert np.abs(got - _block(ref, sh, lx, ly)).max() < 1e-5
    # the stand-alone exchange / unpad pair reproduces the fused reduce
    padw, ext = get_halo_size(h, sh)
    hx, hy = padw[0][0], padw[1][0]
    padded = torch.from_numpy(OP.cic_paint_dx_padded(dloc.numpy(), 1.0, (hx, hy)))
    two_step = halo.unpad_reduce(halo.halo_exchange(padded, ext, sh), padw).numpy()
    fused = halo.halo_reduce_(padded.clone(), hx, hy, sh).numpy()
    assert np.abs(two_step - fused).max() < 1e-5


@pytest.mark.parametrize("pdims", [(2, 1), (1, 2), (2, 2)])
def test_halo_protocol(pdims):
    _run(_halo_worker, pdims[0] * pdims[1], pdims, (16, 16, 6), 4)


# ---- slab path: host-side logic (the kernels are CUDA-only; the handle exchange and the routing rule are not) ----
class _FakeSlabPlan:
    """Stands in for slab.SlabPlan: a rank-tagged 64-byte 'IPC handle' and a record of what was attached."""

    def __init__(self, nranks, rank):
        self.nranks, self.rank, self.attached, self.got = nranks, rank, False, None

    def ipc_handle(self):
        return bytes([self.rank]) * 64

    def attach_ipc(self, handles):
        self.got, self.attached = list(handles), True


def _slab_connect_worker(rank, world):
    from jaxpm_b200 import slab
    plan = _FakeSlabPlan(world, rank)
    slab.connect(plan)
    assert plan.attached and len(plan.got) == world
    for r, h in enumerate(plan.got):          # rank order, every rank sees every handle
        assert h == bytes([r]) * 64


@pytest.mark.parametrize("world", [2, 4])
def test_slab_handle_exchange(world):
    """slab.connect gathers the ranks' 64-byte handles in rank order (gloo here, NCCL on the box)."""
    _run(_slab_connect_worker, world)


def test_slab_routing_rule():
    """Which decompositions take the fused peer-memory stepper (slab.slab_supported): slab (P, 1) and pencil (px, py)
    grids on power-of-two meshes with the halo inside one block and ny / py a multiple of 16; everything else stays on
    the NCCL path."""
    from jaxpm_b200.slab import slab_supported
    assert slab_supported((512, 512, 512), (8, 1), 64)
    assert slab_supported((1024, 1024, 1024), (8, 1), 64)
    assert slab_supported((32, 32, 32), (2, 1), 8)
    assert slab_supported((512, 512, 512), (1, 8), 64, 64)        # y slabs = a (1, P) pencil grid
    assert slab_supported((512, 512, 512), (4, 2), 64)            # pencils (tests/test_distributed_pm.py:28)
    assert slab_supported((512, 512, 512), (2, 4), 64, 32)
    assert not slab_supported((64, 64, 64), (1, 8), 8, 8)         # ny / py = 8 rows: below the 16-row tiles of the z passes
    assert not slab_supported((512, 512, 512), (2, 4), 64, 129)   # halo wider than a pencil
    assert not slab_supported((512, 512, 512), (4, 4), 64)        # more than the 8 GPUs of a box
    assert not slab_supported((32, 32, 24), (2, 1), 8)            # not a power of two
    assert not slab_supported((512, 512, 512), (8, 1), 65)        # halo wider than a slab
    assert not slab_supported((64, 64, 64), (8, 1), 0)


@pytest.mark.parametrize("pdims,shape,gx,gy,ge", [((2, 2), (32, 32), 8, 8, 5), ((2, 4), (64, 64), 8, 8, 8), ((4, 2), (64, 64), 16, 6, 9),
                                                  ((1, 4), (32, 64), 4, 8, 3), ((1, 2), (64, 32), 8, 16, 16)])
def test_pencil_exchange_rule(pdims, shape, gx, gy, ge):
    """The routing rule of the fused path on pencil grids (slab.pencil_targets, the host restatement of the device
    function the z passes use), on a 2-D model (z is local): every rank paints into its pencil + `ge` ghost planes /
    rows; the forward pass must hand every FFT slab the periodic global sum (halo reduce of
    jaxpm/distributed.py:61-85 + the row-group transpose), the inverse pass must fill every rank's pencil AND ghost
    frame, corners included, with the global field (halo fill, jaxpm/painting.py:248-252)."""
    from jaxpm_b200.slab import pencil_targets
    px, py = pdims
    nx, ny = shape
    P, G, rows = px * py, 4, 16
    lx, Lx, Ly = nx // P, nx // px, ny // py
    rng = np.random.default_rng(3)
    gex, gey = min(gx, ge), min(gy, ge)
    local, truth = [], np.zeros(shape)
    for r in range(P):
        a, b = divmod(r, py)
        arr = np.zeros((Lx + 2 * gx, Ly + 2 * gy + 2 * G))
        # what a paint can touch: the pencil and ge ghost planes / rows around it
        x0, x1, y0, y1 = gx - gex, gx + Lx + gex, G + gy - gey, G + gy + Ly + gey
        arr[x0:x1, y0:y1] = rng.standard_normal((x1 - x0, y1 - y0))
        local.append(arr)
        xs = (np.arange(x0, x1) - gx + a * Lx) % nx
        ys = (np.arange(y0, y1) - G - gy + b * Ly) % ny
        np.add.at(truth, (xs[:, None], ys[None, :]), arr[x0:x1, y0:y1])
    # forward: slab r, plane xl, row tile y0 = sum over the targets
    got = np.zeros(shape)
    for r in range(P):
        for xl in range(lx):
            for y0 in range(0, ny, rows):
                acc = np.zeros(rows)
                for owner, plane, row0, r0, r1 in pencil_targets(pdims, r, shape, gx, gy, xl, y0, ge, rows, G):
                    acc[r0:r1] += local[owner][plane, row0 + r0:row0 + r1]
                got[r * lx + xl, y0:y0 + rows] = acc
    np.testing.assert_allclose(got, truth, rtol=0, atol=1e-12)
    # inverse: every rank's pencil + ghost frame receives the global field
    field = rng.standard_normal(shape)
    out = [np.full_like(a_, np.nan) for a_ in local]
    for r in range(P):
        for xl in range(lx):
            for y0 in range(0, ny, rows):
                for owner, plane, row0, r0, r1 in pencil_targets(pdims, r, shape, gx, gy, xl, y0, ge, rows, G):
                    out[owner][plane, row0 + r0:row0 + r1] = field[r * lx + xl, y0 + r0:y0 + r1]
    for r in range(P):
        a, b = divmod(r, py)
        x0, x1, y0, y1 = gx - gex, gx + Lx + gex, G + gy - gey, G + gy + Ly + gey
        xs = (np.arange(x0, x1) - gx + a * Lx) % nx
        ys = (np.arange(y0, y1) - G - gy + b * Ly) % ny
        np.testing.assert_array_equal(out[r][x0:x1, y0:y1], field[xs[:, None], ys[None, :]])
        rest = out[r].copy()
        rest[x0:x1, y0:y1] = np.nan
        assert np.isnan(rest).all()          # nothing is written outside the frame in use
