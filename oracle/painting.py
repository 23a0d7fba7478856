"""Oracle (test infrastructure): CIC paint / read, absolute and relative mode.

NumPy restatement of
  * /root/reference/jaxpm/painting.py:15-45   (_cic_paint_impl)
  * /root/reference/jaxpm/painting.py:78-106  (_cic_read_impl)
  * /root/reference/jaxpm/painting_utils.py:28-96 (enmesh, the relative-mode rule)
  * /root/reference/jaxpm/painting.py:161-189,218-236 (_cic_paint_dx_impl, _cic_read_dx_impl)

All arithmetic that decides a cell index or a CIC weight is done in the dtype
of the particle array (float32 by default) with the same operation order as
the reference; only the final scatter accumulation is done in float64 (the
reference's atomics have no defined order, so the oracle returns the exactly
rounded sum the fp32 result must sit within 1e-5 of).
"""
import numpy as np

# corner enumeration of the absolute path, painting.py:24-25
_CONN_ABS = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1],
                      [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]])
# binary counting of the relative path, painting_utils.py:43-45 (bit d of c -> axis d)
_CONN_REL = (np.arange(8)[:, None] >> np.arange(3)) & 1


def cic_indices_weights(positions, mesh_shape):
    """Cell indices [Np,8,3] int32 and weights [Np,8] following painting.py:19-37.

    floor -> + corner -> kernel = 1-|x - corner| -> kx*ky*kz ; index = int32(corner) mod N
    (Python-sign modulo).
    """
    pos = np.asarray(positions)
    dt = pos.dtype
    pos = pos.reshape(-1, 1, 3)
    fl = np.floor(pos)
    nc = fl + _CONN_ABS.astype(dt)[None]
    k = (1.0 - np.abs(pos - nc)).astype(dt)
    w = (k[..., 0] * k[..., 1]).astype(dt) * k[..., 2]
    idx = np.mod(nc.astype(np.int32), np.asarray(mesh_shape, dtype=np.int32))
    return idx.astype(np.int32), w.astype(dt)


def _flat(idx, shape):
    return (idx[..., 0].astype(np.int64) * shape[1] + idx[..., 1]) * shape[2] + idx[..., 2]


def cic_paint(grid_mesh, positions, weight=1.0):
    """painting.py:15-45.  Accumulates into (a copy of) ``grid_mesh``."""
    mesh = np.asarray(grid_mesh)
    dt = mesh.dtype
    idx, w = cic_indices_weights(np.asarray(positions, dtype=dt), mesh.shape)
    if np.isscalar(weight) or np.ndim(weight) == 0:
        w = (np.asarray(weight, dtype=dt) * w).astype(dt)
    else:
        w = (np.asarray(weight, dtype=dt).reshape(-1, 1) * w).astype(dt)
    acc = np.bincount(_flat(idx, mesh.shape).ravel(),
                      weights=w.ravel().astype(np.float64),
                      minlength=mesh.size)
    return (mesh.astype(np.float64) + acc.reshape(mesh.shape)).astype(dt)


def cic_read(grid_mesh, positions):
    """painting.py:78-106.  Returns positions.shape[:-1]."""
    mesh = np.asarray(grid_mesh)
    pos = np.asarray(positions, dtype=mesh.dtype)
    idx, w = cic_indices_weights(pos, mesh.shape)
    vals = mesh.reshape(-1)[_flat(idx, mesh.shape)]
    out = (vals * w).astype(mesh.dtype)
    # jnp .sum(axis=-1) over 8 terms in fp32
    return out.sum(axis=-1, dtype=mesh.dtype).reshape(pos.shape[:-1])


def enmesh_rel(pmid, disp, mesh_shape):
    """painting_utils.py:28-96 with cell_size=new_cell_size=1, offset=0,
    base_shape=new_shape=mesh_shape (the only way the reference calls it,
    painting_utils.py:106-107,180-181).

    Returns int32 indices [Np,8,3] (may contain the out-of-range value N, see
    SURVEY.md §2.2: such entries are dropped by the scatter / read as 0) and
    weights [Np,8].
    """
    disp = np.asarray(disp)
    dt = disp.dtype
    one = dt.type(1.0)
    L = np.asarray(mesh_shape, dtype=np.int32).astype(dt)          # grid_length
    pp = (np.asarray(pmid, dtype=np.int32) * one + disp).astype(dt)  # :48
    pp = pp[:, None, :]
    ni = (pp + _CONN_REL.astype(dt)[None] * one).astype(dt)         # :51
    # jnp float mod: C fmod, then + L when the sign differs (python-sign modulo), :54
    r = np.fmod(ni, L)
    r = np.where((r != 0) & (r < 0), (r + L).astype(dt), r)
    ni = np.floor(r / one) * 1                                       # :56 floor-div by 1
    nd = (pp - ni * one).astype(dt)                                  # :57
    nd = (nd - (np.rint(nd / L) * L).astype(dt)).astype(dt)          # :60-62
    idx = ni.astype(np.int32)
    w = (one - np.abs(nd)).astype(dt)
    w = w.prod(axis=-1, dtype=dt)                                    # :93
    return idx, w


def _pmid(shape, halo_x, halo_y):
    a, b, c = np.meshgrid(np.arange(shape[0], dtype=np.int32),
                          np.arange(shape[1], dtype=np.int32),
                          np.arange(shape[2], dtype=np.int32), indexing='ij')
    return np.stack([a + halo_x, b + halo_y, c], axis=-1).reshape(-1, 3)


def cic_paint_dx_padded(displacements, weight=1.0, halo=(0, 0)):
    """painting.py:161-189 for ONE shard: returns the zero-padded local mesh
    (nx+2hx, ny+2hy, nz) before halo exchange."""
    disp = np.asarray(displacements)
    dt = disp.dtype
    shp = disp.shape[:-1]
    hx, hy = halo
    pshape = (shp[0] + 2 * hx, shp[1] + 2 * hy, shp[2])
    idx, w = enmesh_rel(_pmid(shp, hx, hy), disp.reshape(-1, 3), pshape)
    if np.isscalar(weight) or np.ndim(weight) == 0:
        w = (np.asarray(weight, dtype=dt) * w).astype(dt)
    else:
        w = (np.asarray(weight, dtype=dt).reshape(-1, 1) * w).astype(dt)
    ok = np.all((idx >= 0) & (idx < np.asarray(pshape)), axis=-1)   # OOB dropped by scatter
    flat = _flat(np.where(ok[..., None], idx, 0), pshape)
    acc = np.bincount(flat.ravel(), weights=np.where(ok, w, 0).ravel().astype(np.float64),
                      minlength=int(np.prod(pshape)))
    return acc.reshape(pshape).astype(dt)


def cic_paint_dx(displacements, weight=1.0):
    """painting.py:192-215, single device (halo 0)."""
    return cic_paint_dx_padded(displacements, weight, (0, 0))


def cic_read_dx_padded(padded_mesh, disp, halo=(0, 0)):
    """painting.py:218-236 for ONE shard: ``padded_mesh`` is the local mesh
    already padded by ``halo`` and halo-filled."""
    mesh = np.asarray(padded_mesh)
    disp = np.asarray(disp, dtype=mesh.dtype)
    hx, hy = halo
    shp = (mesh.shape[0] - 2 * hx, mesh.shape[1] - 2 * hy, mesh.shape[2])
    idx, w = enmesh_rel(_pmid(shp, hx, hy), disp.reshape(-1, 3), mesh.shape)
    ok = np.all((idx >= 0) & (idx < np.asarray(mesh.shape)), axis=-1)
    vals = mesh.reshape(-1)[_flat(np.where(ok[..., None], idx, 0), mesh.shape)]
    vals = np.where(ok, vals, 0)                                    # mode='drop', fill 0
    return (vals * w).astype(mesh.dtype).sum(axis=1, dtype=mesh.dtype).reshape(shp)


def cic_read_dx(grid_mesh, disp):
    """painting.py:239-260, single device."""
    return cic_read_dx_padded(grid_mesh, disp, (0, 0))


# ---------------------------------------------------------------------------
# Adjoints (what JAX autodiff of the functions above produces); used to check
# the custom adjoint kernels.  d/dx (1-|x-c|) = -sign(x-c), sign(0)=0.
# ---------------------------------------------------------------------------
def cic_weight_grads(positions):
    """Per-corner weights [Np,8] and d(weight)/d(pos) [Np,8,3] (absolute rule)."""
    pos = np.asarray(positions)
    dt = pos.dtype
    p = pos.reshape(-1, 1, 3)
    nc = np.floor(p) + _CONN_ABS.astype(dt)[None]
    d = p - nc
    k = 1.0 - np.abs(d)
    s = -np.sign(d)
    w = k[..., 0] * k[..., 1] * k[..., 2]
    g = np.stack([s[..., 0] * k[..., 1] * k[..., 2],
                  k[..., 0] * s[..., 1] * k[..., 2],
                  k[..., 0] * k[..., 1] * s[..., 2]], axis=-1)
    return w.astype(dt), g.astype(dt)


def cic_read_vjp(grid_mesh, positions, cot):
    """VJP of cic_read wrt (mesh, positions) for cotangent ``cot`` [Np]."""
    mesh = np.asarray(grid_mesh)
    pos = np.asarray(positions, dtype=mesh.dtype)
    idx, _ = cic_indices_weights(pos, mesh.shape)
    w, g = cic_weight_grads(pos)
    flat = _flat(idx, mesh.shape)
    c = np.asarray(cot, dtype=np.float64).reshape(-1, 1)
    gmesh = np.bincount(flat.ravel(), weights=(w * c).ravel(), minlength=mesh.size)
    vals = mesh.reshape(-1)[flat].astype(np.float64)
    gpos = (vals[..., None] * g).sum(axis=1) * c
    return gmesh.reshape(mesh.shape).astype(mesh.dtype), gpos.reshape(pos.shape).astype(mesh.dtype)


def cic_paint_vjp(mesh_shape, positions, weight, cot):
    """VJP of cic_paint wrt (positions, weight-array) for cotangent mesh ``cot``."""
    cot = np.asarray(cot)
    pos = np.asarray(positions, dtype=cot.dtype)
    idx, _ = cic_indices_weights(pos, mesh_shape)
    w, g = cic_weight_grads(pos)
    vals = cot.reshape(-1)[_flat(idx, mesh_shape)].astype(np.float64)
    wt = np.broadcast_to(np.asarray(weight, dtype=np.float64).reshape(-1, 1) if np.ndim(weight) else
                         np.float64(weight), (pos.reshape(-1, 3).shape[0], 1))
    gpos = (vals[..., None] * g).sum(axis=1) * wt
    gw = (vals * w).sum(axis=1)
    return gpos.reshape(pos.shape).astype(cot.dtype), gw.astype(cot.dtype)


def cic_paint_2d(mesh, positions, weight):
    """painting.py:131-158: 2-D CIC, kernel = (1-|dx|)(1-|dy|) [* weight], int32 cast, python mod."""
    mesh = np.asarray(mesh)
    dt = mesh.dtype
    pos = np.asarray(positions, dtype=dt).reshape(-1, 1, 2)
    fl = np.floor(pos)
    conn = np.array([[0, 0], [1., 0], [0., 1], [1., 1]], dtype=dt)
    nc = fl + conn
    k = (1. - np.abs(pos - nc)).astype(dt)
    k = (k[..., 0] * k[..., 1]).astype(dt)
    if weight is not None:
        k = (k * np.asarray(weight, dtype=dt).reshape(-1, 1)).astype(dt)
    idx = np.mod(nc.astype(np.int32), np.array(mesh.shape))
    flat = idx[..., 0].astype(np.int64) * mesh.shape[1] + idx[..., 1]
    acc = np.bincount(flat.ravel(), weights=k.ravel().astype(np.float64), minlength=mesh.size)
    return (mesh.astype(np.float64) + acc.reshape(mesh.shape)).astype(dt)


def density_plane(positions, box_shape, center, width, plane_resolution):
    """lensing.py:11-44 without the optional smoothing: mod, rescale, slab mask, cic_paint_2d, normalisation."""
    nx, ny, nz = box_shape
    pos = np.asarray(positions, dtype=np.float32)
    xy = np.mod(pos[..., :2], np.float32(nx)).astype(np.float32)
    xy = (xy / np.float32(nx) * np.float32(plane_resolution)).astype(np.float32)
    d = pos[..., 2]
    w = np.where((d > (center - width / 2)) & (d <= (center + width / 2)), 1., 0.).astype(np.float32)
    plane = cic_paint_2d(np.zeros([plane_resolution, plane_resolution], np.float32), xy, w)
    return (plane / np.float32((nx / plane_resolution) * (ny / plane_resolution) * width)).astype(np.float32)
