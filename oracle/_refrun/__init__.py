"""Reference runner (test infrastructure, never imported by the product).

JAX, jaxdecomp and jax_cosmo cannot be installed in this image, so the reference
(`/root/reference/jaxpm`) cannot be imported as is.  `shim/` is a NumPy stand-in for
the small part of the `jax` / `jaxdecomp` / `jax_cosmo` API that the reference's hot
path touches; with it on `sys.path`, `make_golden.py` imports the reference's OWN,
UNMODIFIED source files from `/root/reference/jaxpm` and records what they compute
as fixtures under `tests/golden/`.  Nothing from the reference is copied.

What this pins: every line of Python logic in the reference (corner enumeration,
weight products, modulo / floor-div / rint wrap, halo pad / add rule, kernel
formulas, LPT and ODE factor algebra) is executed verbatim.  What it does not pin:
the primitives themselves (`lax.scatter_add`, `jnp.fft`, XLA's fp32 reduction
order) are NumPy's here, JAX type promotion is emulated (x64 off: everything stays
32-bit, ints never widen floats), and the [ext] packages jaxdecomp (FFT, fftfreq3d,
halo_exchange) and jax_cosmo (background / growth tables) are restated from their
published behaviour.  DESIGN.md states the same.
"""
