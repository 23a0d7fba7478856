"""jaxpm_b200 — B200-native particle-mesh force loop behind the `jaxpm` API.

Like the reference's `jaxpm/__init__.py` (empty), users import from the
submodules: `jaxpm_b200.painting`, `.pm`, `.ode`, `.kernels`, `.distributed`,
`.growth`, `.utils`.  Importing this package does not load the CUDA library;
the first op does, and fails loudly if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"
