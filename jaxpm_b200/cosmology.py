"""Host-side scalar cosmology (NumPy, float64): background, growth tables, linear P(k).

These are the O(1)-per-step scalars the reference obtains from jax_cosmo
(`jc.background.Esqr`, `growth_factor`, ... — /root/reference/jaxpm/growth.py:1-10,
/root/reference/jaxpm/pm.py:77-87) plus its own FastPM factors
(/root/reference/jaxpm/growth.py:31-293).  They stay on the host: the CUDA kernels
only ever receive the resulting kick/drift coefficients.

[ext] jax_cosmo is not installed; the published algorithms are restated
(growth ODE of `_compute_growth_tables`, Eisenstein & Hu 1998 transfer function
with baryon wiggles, sigma8 normalisation).  The growth ODE is integrated with
a fixed-step RK4 on a log-a grid (independent of the oracle's SciPy solver).
"""
import numpy as np


class Cosmology:
    def __init__(self, Omega_c, Omega_b, h, n_s, sigma8, Omega_k=0.0, w0=-1.0, wa=0.0):
        self.Omega_c, self.Omega_b, self.h, self.n_s = float(Omega_c), float(Omega_b), float(h), float(n_s)
        self.sigma8, self.Omega_k, self.w0, self.wa = float(sigma8), float(Omega_k), float(w0), float(wa)
        self._cache = {}

    @property
    def Omega_m(self):
        return self.Omega_b + self.Omega_c

    @property
    def Omega_de(self):
        return 1.0 - self.Omega_k - self.Omega_m

    def __repr__(self):
        return (f"Cosmology(Omega_c={self.Omega_c}, Omega_b={self.Omega_b}, h={self.h}, n_s={self.n_s}, "
                f"sigma8={self.sigma8}, Omega_k={self.Omega_k}, w0={self.w0}, wa={self.wa})")


def Planck15(**kw):
    p = dict(Omega_c=0.2589, Omega_b=0.04860, Omega_k=0.0, h=0.6774, n_s=0.9667, sigma8=0.8159,
             w0=-1.0, wa=0.0)
    p.update(kw)
    return Cosmology(**p)


# ---- background --------------------------------------------------------------------
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


# ---- growth tables -------------------------------------------------------------------
def _derivs(cosmo, a, y):
    g1, g2, f1, f2 = y
    om, ode = Omega_m_a(cosmo, a), Omega_de_a(cosmo, a)
    q = (2.0 - 0.5 * (om + (1.0 + 3.0 * w(cosmo, a)) * ode)) / a
    r = 1.5 * om / a / a
    return np.array([f1, f2, -q * f1 + r * g1, -q * f2 + r * g2 - r * g1 * g1])


def _compute_growth_tables(cosmo, log10_amin=-3.0, steps=2048, substeps=4):
    """(atab, gtab, ftab, htab, g2tab, f2tab, h2tab), same layout as jax_cosmo's cache."""
    if "growth" not in cosmo._cache:
        atab = np.logspace(log10_amin, 0.0, steps)
        y = np.empty((steps, 4))
        a0 = atab[0]
        y[0] = [a0, -3.0 / 7 * a0**2, 1.0, -6.0 / 7 * a0]
        for i in range(steps - 1):
            yy, a = y[i].copy(), atab[i]
            h = (atab[i + 1] - atab[i]) / substeps
            for _ in range(substeps):
                k1 = _derivs(cosmo, a, yy)
                k2 = _derivs(cosmo, a + 0.5 * h, yy + 0.5 * h * k1)
                k3 = _derivs(cosmo, a + 0.5 * h, yy + 0.5 * h * k2)
                k4 = _derivs(cosmo, a + h, yy + h * k3)
                yy = yy + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
                a += h
            y[i + 1] = yy
        d2 = np.array([_derivs(cosmo, a, y[i]) for i, a in enumerate(atab)])
        y1, y2 = y[:, 0], y[:, 1]
        g, g2 = y1 / y1[-1], y2 / y2[-1]
        f = y[:, 2] / y1[-1] * atab / g
        f2 = y[:, 3] / y2[-1] * atab / g2
        hh = d2[:, 2] / y1[-1] * atab / g
        h2 = d2[:, 3] / y2[-1] * atab / g2
        cosmo._cache["growth"] = (atab, g, f, hh, g2, f2, h2)
    return cosmo._cache["growth"]


def _interp(cosmo, a, col):
    t = _compute_growth_tables(cosmo)
    return np.interp(np.log(np.asarray(a, dtype=np.float64)), np.log(t[0]), t[col])


def growth_factor(cosmo, a):
    return _interp(cosmo, a, 1)


def growth_rate(cosmo, a):
    return _interp(cosmo, a, 2)


def growth_factor_second(cosmo, a):
    return _interp(cosmo, a, 4)


def growth_rate_second(cosmo, a):
    return _interp(cosmo, a, 5)


# ---- linear matter power (Eisenstein & Hu 1998, with wiggles) ------------------------
def _eh_transfer(cosmo, k):
    """k in h/Mpc."""
    h = cosmo.h
    w_m = cosmo.Omega_m * h**2
    w_b = cosmo.Omega_b * h**2
    fb = cosmo.Omega_b / cosmo.Omega_m
    fc = (cosmo.Omega_m - cosmo.Omega_b) / cosmo.Omega_m
    T_2_7_sqr = (2.7255 / 2.7)**2
    k = np.asarray(k, dtype=np.float64) * h  # 1/Mpc
    k_eq = 7.46e-2 * w_m / T_2_7_sqr
    z_eq = 2.50e4 * w_m / T_2_7_sqr**2
    b1 = 0.313 * w_m**-0.419 * (1.0 + 0.607 * w_m**0.674)
    b2 = 0.238 * w_m**0.223
    z_d = 1291.0 * w_m**0.251 / (1.0 + 0.659 * w_m**0.828) * (1.0 + b1 * w_b**b2)
    R_d = 31.5 * w_b / T_2_7_sqr**2 * (1.0e3 / z_d)
    R_eq = 31.5 * w_b / T_2_7_sqr**2 * (1.0e3 / z_eq)
    sh_d = 2.0 / (3.0 * k_eq) * np.sqrt(6.0 / R_eq) * np.log(
        (np.sqrt(1.0 + R_d) + np.sqrt(R_eq + R_d)) / (1.0 + np.sqrt(R_eq)))
    k_silk = 1.6 * w_b**0.52 * w_m**0.73 * (1.0 + (10.4 * w_m)**-0.95)
    a1 = (46.9 * w_m)**0.670 * (1.0 + (32.1 * w_m)**-0.532)
    a2 = (12.0 * w_m)**0.424 * (1.0 + (45.0 * w_m)**-0.582)
    alpha_c = a1**-fb * a2**(-fb**3)
    bb1 = 0.944 / (1.0 + (458.0 * w_m)**-0.708)
    bb2 = (0.395 * w_m)**-0.0266
    beta_c = 1.0 / (1.0 + bb1 * (fc**bb2 - 1.0))
    y = (1.0 + z_eq) / (1.0 + z_d)
    G = y * (-6.0 * np.sqrt(1.0 + y) + (2.0 + 3.0 * y) * np.log((np.sqrt(1.0 + y) + 1.0) /
                                                               (np.sqrt(1.0 + y) - 1.0)))
    alpha_b = 2.07 * k_eq * sh_d * (1.0 + R_d)**-0.75 * G
    beta_node = 8.41 * w_m**0.435
    beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * np.sqrt((17.2 * w_m)**2 + 1.0)
    q = k / (13.41 * k_eq)

    def T0(kk, ac, bc):
        qq = kk / (13.41 * k_eq)
        Cc = 14.2 / ac + 386.0 / (1.0 + 69.9 * qq**1.08)
        L = np.log(np.e + 1.8 * bc * qq)
        return L / (L + Cc * qq * qq)

    f = 1.0 / (1.0 + (k * sh_d / 5.4)**4)
    Tc = f * T0(k, 1.0, beta_c) + (1.0 - f) * T0(k, alpha_c, beta_c)
    s_tilde = sh_d / (1.0 + (beta_node / (k * sh_d))**3)**(1.0 / 3.0)
    x = k * s_tilde
    j0 = np.where(x == 0, 1.0, np.sin(x) / np.where(x == 0, 1.0, x))
    Tb = (T0(k, 1.0, 1.0) / (1.0 + (k * sh_d / 5.2)**2) +
          alpha_b / (1.0 + (beta_b / (k * sh_d))**3) * np.exp(-(k / k_silk)**1.4)) * j0
    return fb * Tb + fc * Tc


def linear_matter_power(cosmo, k, a=1.0):
    """P_lin(k) in (Mpc/h)^3, k in h/Mpc, normalised to sigma8 at a=1."""
    k = np.asarray(k, dtype=np.float64)

    def unnorm(kk):
        with np.errstate(divide="ignore", invalid="ignore"):     # k = 0: the k^n_s factor makes P(0) = 0
            return kk**cosmo.n_s * _eh_transfer(cosmo, kk)**2

    if "pknorm" not in cosmo._cache:
        lk = np.linspace(np.log(1e-5), np.log(1e3), 8192)
        kk = np.exp(lk)
        x = kk * 8.0
        wth = 3.0 * (np.sin(x) - x * np.cos(x)) / x**3
        integrand = kk**3 * unnorm(kk) * wth**2 / (2.0 * np.pi**2)
        sig2 = np.trapezoid(integrand, lk)
        cosmo._cache["pknorm"] = cosmo.sigma8**2 / sig2
    return cosmo._cache["pknorm"] * unnorm(k) * growth_factor(cosmo, a)**2
