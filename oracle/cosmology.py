"""Oracle (test infrastructure): scalar background cosmology and FastPM growth functions.

Restates /root/reference/jaxpm/growth.py:31-52 (E), :55-85 (df_de), :88-116 (dEa),
:124-154 (Gf), :157-187 (Gf2), :190-224 (dGfa), :227-261 (dGf2a), :264-293 (gp).

[ext] jax_cosmo (unpinned; CI uses ASKabalan/jax_cosmo@better-cache,
.github/workflows/tests.yml:58-60) is absent.  Its published algorithm is
restated: Esqr = Om a^-3 + Ok a^-2 + Ode exp(f_de), f_de Linder w0-wa, and the
linear growth ODE system for (D1, D2) integrated from a=1e-3 with
matter-dominated initial conditions, normalised to 1 at a=1
(jax_cosmo.background._compute_growth_tables: cache = (a, g, f, h, g2, f2, h2)
with f = dlnD/dlna and h = a*D''/D).  Here the ODE is solved with SciPy's
DOP853 at rtol 1e-11 on a dense log-a grid, so table interpolation error is
negligible; parity of these O(1) scalars with a real jax_cosmo run is unpinned.
"""
import numpy as np
from scipy.integrate import solve_ivp


class Cosmology:
    def __init__(self, Omega_c, Omega_b, h, n_s, sigma8, Omega_k=0.0, w0=-1.0, wa=0.0):
        self.Omega_c, self.Omega_b, self.h, self.n_s = Omega_c, Omega_b, h, n_s
        self.sigma8, self.Omega_k, self.w0, self.wa = sigma8, Omega_k, w0, wa
        self._tab = None

    @property
    def Omega_m(self):
        return self.Omega_b + self.Omega_c

    @property
    def Omega_de(self):
        return 1.0 - self.Omega_k - self.Omega_m


def Planck15(**kw):
    # [ext, from memory] jax_cosmo.parameters.Planck15
    p = dict(Omega_c=0.2589, Omega_b=0.04860, Omega_k=0.0, h=0.6774, n_s=0.9667,
             sigma8=0.8159, w0=-1.0, wa=0.0)
    p.update(kw)
    return Cosmology(**p)


def Planck18_tests():
    # /root/reference/tests/conftest.py:52-70
    return Cosmology(Omega_c=0.2607, Omega_b=0.0490, Omega_k=0.0, h=0.6766, n_s=0.9665,
                     sigma8=0.8102, w0=-1.0, wa=0.0)


def w(cosmo, a):
    return cosmo.w0 + (1.0 - a) * cosmo.wa


def f_de(cosmo, a):
    return -3.0 * (1.0 + cosmo.w0 + cosmo.wa) * np.log(a) + 3.0 * cosmo.wa * (a - 1.0)


def Esqr(cosmo, a):
    a = np.asarray(a, dtype=np.float64)
    return cosmo.Omega_m * a**-3 + cosmo.Omega_k * a**-2 + cosmo.Omega_de * np.exp(f_de(cosmo, a))


def Omega_m_a(cosmo, a):
    return cosmo.Omega_m * np.power(a, -3.0) / Esqr(cosmo, a)


def Omega_de_a(cosmo, a):
    return cosmo.Omega_de * np.exp(f_de(cosmo, a)) / Esqr(cosmo, a)


def _rhs(cosmo, a, y):
    g1, g2, f1, f2 = y
    q = (2.0 - 0.5 * (Omega_m_a(cosmo, a) + (1.0 + 3.0 * w(cosmo, a)) * Omega_de_a(cosmo, a))) / a
    r = 1.5 * Omega_m_a(cosmo, a) / a / a
    return np.array([f1, f2, -q * f1 + r * g1, -q * f2 + r * g2 - r * g1**2])


def growth_tables(cosmo, log10_amin=-3.0, steps=2048):
    if cosmo._tab is None:
        atab = np.logspace(log10_amin, 0.0, steps)
        a0 = atab[0]
        y0 = np.array([a0, -3.0 / 7 * a0**2, 1.0, -6.0 / 7 * a0])
        sol = solve_ivp(lambda a, y: _rhs(cosmo, a, y), (a0, 1.0), y0, t_eval=atab,
                        method='DOP853', rtol=1e-11, atol=1e-14)
        y = sol.y
        d2 = np.array([_rhs(cosmo, a, y[:, i]) for i, a in enumerate(atab)]).T
        y1, y2 = y[0], y[1]
        g = y1 / y1[-1]
        g2 = y2 / y2[-1]
        f = y[2] / y1[-1] * atab / g
        f2 = y[3] / y2[-1] * atab / g2
        h = d2[2] / y1[-1] * atab / g
        h2 = d2[3] / y2[-1] * atab / g2
        cosmo._tab = (atab, g, f, h, g2, f2, h2)
    return cosmo._tab


def _interp(cosmo, a, col):
    t = growth_tables(cosmo)
    return np.interp(np.log(np.asarray(a, dtype=np.float64)), np.log(t[0]), t[col])


def growth_factor(cosmo, a):
    return _interp(cosmo, a, 1)


def growth_rate(cosmo, a):
    return _interp(cosmo, a, 2)


def growth_factor_second(cosmo, a):
    return _interp(cosmo, a, 4)


def growth_rate_second(cosmo, a):
    return _interp(cosmo, a, 5)


# ---- reference growth.py ---------------------------------------------------
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


def dGfa(cosmo, a):
    D1f = gp(cosmo, a)
    t = growth_tables(cosmo)
    f1p = np.interp(np.log(a), np.log(t[0]), t[3] / t[0] * t[1])
    Ea = E(cosmo, a)
    return f1p * a**3 * Ea + D1f * a**3 * dEa(cosmo, a) + 3 * a**2 * Ea * D1f


def dGf2a(cosmo, a):
    D2f = growth_rate_second(cosmo, a) * growth_factor_second(cosmo, a) / a
    t = growth_tables(cosmo)
    f2p = np.interp(np.log(a), np.log(t[0]), t[6] / t[0] * t[4])
    Ea = E(cosmo, a)
    return f2p * a**3 * Ea + D2f * a**3 * dEa(cosmo, a) + 3 * a**2 * Ea * D2f
