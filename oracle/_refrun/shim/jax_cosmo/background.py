import numpy as _np

import jax.numpy as jnp
from oracle import cosmology as _C


def _w(f):
    def g(cosmo, a, *args, **kw):
        return jnp._wrap(_np.asarray(f(cosmo, _np.asarray(jnp._raw(a), dtype=_np.float64), *args, **kw)))
    return g


Esqr, Omega_m_a, Omega_de_a, w, f_de = (_w(_C.Esqr), _w(_C.Omega_m_a), _w(_C.Omega_de_a), _w(_C.w),
                                        _w(_C.f_de))
growth_factor, growth_rate = _w(_C.growth_factor), _w(_C.growth_rate)
growth_factor_second, growth_rate_second = _w(_C.growth_factor_second), _w(_C.growth_rate_second)


def _compute_growth_tables(cosmo):
    return tuple(jnp._wrap(_np.asarray(t)) for t in _C.growth_tables(cosmo))
