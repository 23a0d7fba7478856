"""Shared test inputs (seeded) and error metrics."""
import numpy as np


def lagrangian_grid(shape, dtype=np.float32):
    return np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).astype(dtype)


def displaced(shape, sigma, seed=1, dtype=np.float32):
    rng = np.random.default_rng(seed)
    disp = (sigma * rng.standard_normal((*shape, 3))).astype(dtype)
    return lagrangian_grid(shape, dtype), disp


def rel_err(a, b):
    """max |a-b| / max |b| (the 1e-5 field criterion of BASELINE.md §5)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def gaussian_ic(shape, box, cosmo_pk, seed=0):
    """White noise coloured by sqrt(P(k) Nc / V) (pm.py:129-144), float32."""
    rng = np.random.default_rng(seed)
    wn = rng.standard_normal(shape).astype(np.float32)
    from oracle.pm import linear_field
    return linear_field(wn, box, cosmo_pk)
