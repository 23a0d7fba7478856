"""CPU oracle for the JaxPM particle-mesh force loop.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy/SciPy restatement of the
reference's algorithm (``/root/reference/jaxpm``) for the hot path named in
``BASELINE.json``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``jaxpm_b200``) never does: it fails loudly when the CUDA library is
missing.

Parity pin: the reference cannot be imported here (JAX, jaxdecomp, jax_cosmo
are not installed and cannot be installed) and it ships no golden vectors
(SURVEY.md §0.10).  The oracle is pinned two ways instead:

* ``oracle/_refrun`` executes the reference's *own source files* from
  ``/root/reference/jaxpm`` on top of a NumPy stand-in for the ``jax`` API and
  records their outputs as fixtures under ``tests/golden/`` (the generating
  script is ``oracle/_refrun/make_golden.py``).  Every oracle function is
  checked against those fixtures in ``tests/test_oracle_golden.py``.
* analytic known-answer tests (mass conservation, plane-wave LPT, adjointness,
  sharded == unsharded) in ``tests/test_oracle_kat.py``.

Third-party arithmetic the reference delegates to packages that are absent
from ``/root/reference`` (jaxdecomp>=0.2.9 FFT/halo, jax_cosmo growth tables)
is restated from the published algorithm and marked ``[ext]`` where it appears.
"""

from . import cosmology, distributed, kernels, ode, painting, pm, utils  # noqa: F401
