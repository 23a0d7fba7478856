"""Stub of jax_cosmo.constants for the NumPy stand-in (import-time only: jaxpm/lensing.py:4 imports it; the
golden generator never calls convergence_Born, the only user of these numbers)."""
rh = 2997.92458          # c / H0 in Mpc/h
H0 = 100.0               # km/s/(Mpc/h)
c = 299792.458           # km/s
