"""PRNG stand-in: NOT threefry — the reference's random streams are not reproducible here, so
the golden generator always passes explicit arrays."""
import numpy as _np

from . import numpy as jnp


def PRNGKey(seed):
    return int(seed)


key = PRNGKey


def split(k, n=2):
    return [int(k) * 7919 + i + 1 for i in range(n)]


def normal(key, shape, dtype=float):
    dt = _np.float32 if (dtype is float and not jnp.X64) else dtype
    return jnp._wrap(_np.random.default_rng(int(key)).standard_normal(shape).astype(dt))
