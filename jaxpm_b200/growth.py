"""FastPM growth functions, same names as /root/reference/jaxpm/growth.py.

E :31-52, df_de :55-85, dEa :88-116, Gf :124-154, Gf2 :157-187, dGfa :190-224,
dGf2a :227-261, gp :264-293, Dplusdada :296-314, Dplus_to_a :317-334, plus the
jax_cosmo re-exports (:6-10).  Host-side NumPy float64 scalars.
"""
import numpy as np

from .cosmology import (Esqr, Omega_de_a, Omega_m_a, _compute_growth_tables, f_de,  # noqa: F401
                        growth_factor, growth_factor_second, growth_rate, growth_rate_second, w)

__all__ = ["growth_factor", "growth_rate", "growth_factor_second", "growth_rate_second", "E", "df_de",
           "dEa", "Gf", "Gf2", "dGfa", "dGf2a", "gp", "Dplusdada", "Dplus_to_a"]


def E(cosmo, a):
    return np.sqrt(Esqr(cosmo, a))


def df_de(cosmo, a, epsilon=1e-5):
    return (3 * cosmo.wa * (np.log(a - epsilon) - (a - 1) / (a - epsilon)) /
            np.power(np.log(a - epsilon), 2))


def dEa(cosmo, a):
    a = np.asarray(a, dtype=np.float64)
    return (0.5 * (-3 * cosmo.Omega_m * np.power(a, -4) - 2 * cosmo.Omega_k * np.power(a, -3) +
                   df_de(cosmo, a) * cosmo.Omega_de * np.exp(f_de(cosmo, a))) / E(cosmo, a))


def gp(cosmo, a):
    return growth_rate(cosmo, a) * growth_factor(cosmo, a) / a


def Gf(cosmo, a):
    return gp(cosmo, a) * np.power(a, 3) * E(cosmo, a)


def Gf2(cosmo, a):
    D2f = growth_rate_second(cosmo, a) * growth_factor_second(cosmo, a) / a
    return D2f * np.power(a, 3) * E(cosmo, a)


def _second_derivative(cosmo, a, hcol, gcol):
    t = _compute_growth_tables(cosmo)
    return np.interp(np.log(a), np.log(t[0]), t[hcol] / t[0] * t[gcol])


def dGfa(cosmo, a):
    D1f = gp(cosmo, a)
    Ea = E(cosmo, a)
    return (_second_derivative(cosmo, a, 3, 1) * a**3 * Ea + D1f * a**3 * dEa(cosmo, a) +
            3 * a**2 * Ea * D1f)


def dGf2a(cosmo, a):
    D2f = growth_rate_second(cosmo, a) * growth_factor_second(cosmo, a) / a
    Ea = E(cosmo, a)
    return (_second_derivative(cosmo, a, 6, 4) * a**3 * Ea + D2f * a**3 * dEa(cosmo, a) +
            3 * a**2 * Ea * D2f)


def Dplusdada(cosmo, a):
    return _second_derivative(cosmo, np.atleast_1d(a), 3, 1)


def Dplus_to_a(cosmo, D):
    t = _compute_growth_tables(cosmo)
    return np.interp(np.atleast_1d(D), t[1], t[0])
