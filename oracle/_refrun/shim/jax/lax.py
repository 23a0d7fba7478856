"""`jax.lax` stand-in: scatter_add (point scatter), scan, axis_index, psum."""
from dataclasses import dataclass

import numpy as _np

from . import numpy as jnp
from . import sharding as _sh


class FftType:
    FFT, IFFT, RFFT, IRFFT = range(4)


@dataclass
class ScatterDimensionNumbers:
    update_window_dims: tuple
    inserted_window_dims: tuple
    scatter_dims_to_operand_dims: tuple


SCATTER_LOG = []   # (indices, updates) of every scatter_add call, for the golden generator


def scatter_add(operand, indices, updates, dnums, **kw):
    """Point scatter: indices [..., ndim] address single elements (the only form the reference
    uses, painting.py:39-44).  Out-of-bounds updates are dropped (XLA semantics)."""
    op = _np.array(jnp._raw(operand), copy=True)
    idx = _np.asarray(jnp._raw(indices))
    upd = _np.asarray(jnp._raw(updates))
    assert tuple(dnums.inserted_window_dims) == tuple(range(op.ndim)) and not dnums.update_window_dims
    assert idx.shape[-1] == op.ndim and idx.shape[:-1] == upd.shape
    SCATTER_LOG.append((idx.copy(), upd.copy()))
    ok = _np.all((idx >= 0) & (idx < _np.asarray(op.shape)), axis=-1)
    sel = tuple(idx[..., d][ok] for d in range(op.ndim))
    _np.add.at(op, sel, upd[ok].astype(op.dtype))
    return jnp._wrap(op)


def scan(f, init, xs, length=None):
    leaves = xs if isinstance(xs, (list, tuple)) else [xs]
    n = length if length is not None else len(leaves[0])
    def leaf(v):   # scan traces its carry: Python scalars become (weak) 0-d arrays
        if isinstance(v, (tuple, list)):
            return type(v)(leaf(u) for u in v)
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            return v
        return jnp._wrap(_np.asarray(v, dtype=(_np.int64 if jnp.X64 else _np.int32) if isinstance(v, int)
                                     else (_np.float64 if jnp.X64 else _np.float32)))
    carry, ys = leaf(init), []
    for i in range(n):
        x = type(xs)(l[i] for l in leaves) if isinstance(xs, (list, tuple)) else xs[i]
        carry, y = f(carry, x)
        ys.append(y)
    if ys and ys[0] is not None:
        ys = jnp._wrap(_np.stack([_np.asarray(jnp._raw(y)) for y in ys]))
    else:
        ys = None
    return carry, ys


def axis_index(name):
    return _sh.current_axis_index(name)


def psum(x, axis_name):
    m = _sh.active_mesh()
    return x * m.devices.shape[m.axis_names.index(axis_name)]
