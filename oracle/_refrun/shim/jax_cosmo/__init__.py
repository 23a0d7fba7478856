"""[ext] jax_cosmo stand-in: delegates to oracle/cosmology.py, the restatement of jax_cosmo's
published background / growth-table algorithm (the package is absent and unpinned).
The reference's own growth.py formulas run verbatim on top of these tables."""
from . import background  # noqa: F401
from oracle.cosmology import Cosmology, Planck15  # noqa: F401
