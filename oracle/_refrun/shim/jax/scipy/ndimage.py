"""Stub of jax.scipy.ndimage for the NumPy stand-in (import-time only, jaxpm/lensing.py:5)."""


def map_coordinates(*args, **kwargs):
    raise NotImplementedError("map_coordinates is not part of the NumPy stand-in (lensing convergence is out of scope)")
