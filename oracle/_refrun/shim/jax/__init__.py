"""NumPy stand-in for the part of `jax` the reference's hot path uses (test infrastructure;
see oracle/_refrun/__init__.py).  Eager, single process; `shard_map` loops over the blocks of
a simulated device mesh."""
import functools as _ft

import numpy as _np

from . import lax, numpy, random, sharding, tree  # noqa: F401
from .sharding import _shard_map as shard_map  # noqa: F401

Array = numpy.Arr


def jit(f=None, static_argnums=None, static_argnames=None, **kw):
    if f is None:
        return lambda g: g
    return f


def device_count():
    m = sharding.active_mesh()
    return 1 if m is None else int(_np.prod(m.devices.shape))


def vmap(f, in_axes=0, out_axes=0):
    raise NotImplementedError("vmap is not emulated")


class _Config:
    def update(self, key, val):
        if key == "jax_enable_x64":
            numpy.X64 = bool(val)


config = _Config()
