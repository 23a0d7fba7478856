"""Oracle (test infrastructure): the 2-D domain decomposition, simulated in one process.

Restates /root/reference/jaxpm/distributed.py:45-58 (get_halo_size), :61-65
(halo_exchange -> jaxdecomp.halo_exchange), :68-85 (slice_unpad_impl), :88-113
(slice_pad / slice_unpad), :116-129 (get_local_shape), :168-190
(uniform_particles: particle -> rank rule) and the distributed wrappers
/root/reference/jaxpm/painting.py:192-215 (cic_paint_dx), :239-260 (cic_read_dx).

[ext] jaxdecomp.halo_exchange (absent): restated as the standard periodic halo
update with extents e on an array whose "interior" is [e, S-e): pad [0,e)
receives the low neighbour's [S-2e, S-e), pad [S-e,S) receives the high
neighbour's [e, 2e); axis 0 first, then axis 1, each over the full extent of
the other axes so that corners propagate.
"""
import numpy as np

from . import painting as P


def get_halo_size(halo_size, pdims):
    """distributed.py:45-58.  Accepts an int or a 2-tuple."""
    if pdims is None or tuple(pdims) == (1, 1):
        return ((0, 0), (0, 0), (0, 0)), (0, 0)
    if np.isscalar(halo_size):
        halo_size = (halo_size, halo_size)
    hx = (0, 0) if pdims[0] == 1 else (halo_size[0],) * 2
    hy = (0, 0) if pdims[1] == 1 else (halo_size[1],) * 2
    ex = 0 if pdims[0] == 1 else halo_size[0] // 2
    ey = 0 if pdims[1] == 1 else halo_size[1] // 2
    return (hx, hy, (0, 0)), (ex, ey)


def get_local_shape(mesh_shape, pdims):
    return [mesh_shape[0] // pdims[0], mesh_shape[1] // pdims[1], mesh_shape[2]]


def owner_rank(i, j, mesh_shape, pdims):
    """Particle (Lagrangian index i,j,*) -> (rx, ry); distributed.py:176-184."""
    return i // (mesh_shape[0] // pdims[0]), j // (mesh_shape[1] // pdims[1])


def split(x, pdims):
    """Global array -> {(rx,ry): local block} over axes 0/1."""
    px, py = pdims
    nx, ny = x.shape[0] // px, x.shape[1] // py
    return {(rx, ry): x[rx * nx:(rx + 1) * nx, ry * ny:(ry + 1) * ny].copy()
            for rx in range(px) for ry in range(py)}


def assemble(blocks, pdims):
    px, py = pdims
    return np.concatenate([np.concatenate([blocks[(rx, ry)] for ry in range(py)], axis=1)
                           for rx in range(px)], axis=0)


def halo_exchange(blocks, extents, pdims):
    """[ext] jaxdecomp.halo_exchange, periodic, sequential per axis."""
    px, py = pdims
    ex, ey = extents
    blocks = {k: v.copy() for k, v in blocks.items()}
    if ex > 0:
        old = {k: v.copy() for k, v in blocks.items()}
        for (rx, ry), b in blocks.items():
            lo, hi = old[((rx - 1) % px, ry)], old[((rx + 1) % px, ry)]
            S = b.shape[0]
            b[:ex] = lo[S - 2 * ex:S - ex]
            b[S - ex:] = hi[ex:2 * ex]
    if ey > 0:
        old = {k: v.copy() for k, v in blocks.items()}
        for (rx, ry), b in blocks.items():
            lo, hi = old[(rx, (ry - 1) % py)], old[(rx, (ry + 1) % py)]
            S = b.shape[1]
            b[:, :ey] = lo[:, S - 2 * ey:S - ey]
            b[:, S - ey:] = hi[:, ey:2 * ey]
    return blocks


def slice_unpad_impl(x, pad_width):
    """distributed.py:68-85 (even halos assumed)."""
    hx, hy = pad_width[0][0], pad_width[1][0]
    x = x.copy()
    if hx > 0:
        x[hx:hx + hx // 2] += x[:hx // 2]
        x[-(hx + hx // 2):-hx] += x[-(hx // 2):]
    if hy > 0:
        x[:, hy:hy + hy // 2] += x[:, :hy // 2]
        x[:, -(hy + hy // 2):-hy] += x[:, -(hy // 2):]
    sl = [slice(None)] * 3
    if hx > 0:
        sl[0] = slice(hx, -hx)
    if hy > 0:
        sl[1] = slice(hy, -hy)
    return x[tuple(sl)]


def cic_paint_dx(displacements, halo_size, pdims, weight=1.0):
    """painting.py:192-215 over a simulated (px,py) device mesh.  Returns the
    global mesh and the per-rank padded meshes before the exchange."""
    pad, ext = get_halo_size(halo_size, pdims)
    dblocks = split(np.asarray(displacements), pdims)
    wblocks = None if np.ndim(weight) == 0 else split(np.asarray(weight), pdims)
    padded = {k: P.cic_paint_dx_padded(d, weight if wblocks is None else wblocks[k],
                                       (pad[0][0], pad[1][0])) for k, d in dblocks.items()}
    ex = halo_exchange(padded, ext, pdims)
    out = {k: slice_unpad_impl(v, pad) for k, v in ex.items()}
    return assemble(out, pdims), padded


def cic_read_dx(grid_mesh, disp, halo_size, pdims):
    """painting.py:239-260 over a simulated (px,py) device mesh."""
    pad, ext = get_halo_size(halo_size, pdims)
    pad = tuple((a // 2, b // 2) for a, b in pad)
    mblocks = {k: np.pad(v, pad) for k, v in split(np.asarray(grid_mesh), pdims).items()}
    mblocks = halo_exchange(mblocks, ext, pdims)
    dblocks = split(np.asarray(disp), pdims)
    out = {k: P.cic_read_dx_padded(mblocks[k], dblocks[k], (pad[0][0], pad[1][0]))
           for k in dblocks}
    return assemble(out, pdims)
