"""Known-answer / property tests pinning the CPU oracle."""
import numpy as np
import pytest

from helpers import displaced, lagrangian_grid
from oracle import cosmology as C
from oracle import distributed as D
from oracle import kernels as K
from oracle import ode, painting as P, pm, utils as U


@pytest.mark.parametrize("shape", [(8, 8, 8), (8, 12, 16)])
def test_mass_conservation_and_on_grid(shape):
    grid, disp = displaced(shape, 1.3)
    m = P.cic_paint(np.zeros(shape, np.float32), grid + disp)
    assert abs(m.sum(dtype=np.float64) - np.prod(shape)) < 1e-3
    ones = P.cic_paint(np.zeros(shape, np.float32), grid)
    np.testing.assert_array_equal(ones, np.ones(shape, np.float32))
    np.testing.assert_array_equal(P.cic_paint_dx(np.zeros((*shape, 3), np.float32)), ones)


def test_read_constant_and_linear_field():
    shape = (8, 8, 8)
    grid, disp = displaced(shape, 2.0)
    np.testing.assert_allclose(P.cic_read(np.full(shape, 3.5, np.float32), grid + disp), 3.5, rtol=1e-6)
    # trilinear interpolation reproduces a (periodic-safe) linear ramp away from the wrap
    ramp = np.broadcast_to(np.arange(8, dtype=np.float64)[:, None, None], shape).copy()
    pos = np.random.default_rng(0).uniform(1, 6, (50, 3))
    np.testing.assert_allclose(P.cic_read(ramp, pos), pos[:, 0], rtol=1e-12)


def test_adjointness_paint_read():
    shape = (8, 10, 12)
    grid, disp = displaced(shape, 1.7, dtype=np.float64)
    rng = np.random.default_rng(3)
    w, m = rng.standard_normal(shape), rng.standard_normal(shape)
    lhs = (P.cic_paint(np.zeros(shape), grid + disp, w) * m).sum()
    rhs = (w * P.cic_read(m, grid + disp)).sum()
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)


def test_absolute_and_relative_rules_agree():
    shape = (8, 8, 12)
    grid, disp = displaced(shape, 1.5, dtype=np.float64)
    np.testing.assert_allclose(P.cic_paint_dx(disp), P.cic_paint(np.zeros(shape), grid + disp), atol=1e-12)
    m = np.random.default_rng(1).standard_normal(shape)
    np.testing.assert_allclose(P.cic_read_dx(m, disp), P.cic_read(m, grid + disp), atol=1e-12)


def test_relative_rule_float_mod_edge_case():
    # a coordinate in (-3.8e-6, 0) wraps to index N in fp32 and is dropped
    shape = (8, 8, 8)
    disp = np.zeros((*shape, 3), np.float32)
    disp[0, 0, 0, 0] = -1e-7
    idx, w = P.enmesh_rel(P._pmid(shape, 0, 0), disp.reshape(-1, 3), shape)
    assert idx[0, 0, 0] == 8  # out of range -> dropped
    m = P.cic_paint_dx(disp)
    assert m.sum(dtype=np.float64) < np.prod(shape)  # the dropped corners lose mass


def test_vjp_against_finite_differences():
    shape = (6, 6, 6)
    rng = np.random.default_rng(5)
    pos = rng.uniform(0, 6, (20, 3))
    mesh = rng.standard_normal(shape)
    cot = rng.standard_normal(20)
    gmesh, gpos = P.cic_read_vjp(mesh, pos, cot)
    eps = 1e-6
    for p in range(3):
        for d in range(3):
            pp, pm_ = pos.copy(), pos.copy()
            pp[p, d] += eps
            pm_[p, d] -= eps
            fd = ((P.cic_read(mesh, pp) - P.cic_read(mesh, pm_)) * cot).sum() / (2 * eps)
            assert abs(fd - gpos[p, d]) < 1e-6
    # paint VJP wrt positions
    cotm = rng.standard_normal(shape)
    gp, gw = P.cic_paint_vjp(shape, pos, 1.0, cotm)
    pp, pm_ = pos.copy(), pos.copy()
    pp[4, 1] += eps
    pm_[4, 1] -= eps
    fd = ((P.cic_paint(np.zeros(shape), pp) - P.cic_paint(np.zeros(shape), pm_)) * cotm).sum() / (2 * eps)
    assert abs(fd - gp[4, 1]) < 1e-6


def test_plane_wave_force_is_fd_kernel():
    # delta = cos(k0 x): F_x = Re IFFT(i a(k0) delta_k / k0^2) = -a(k0)/k0^2 sin(k0 x)
    N = 16
    k0 = 2 * np.pi * 2 / N
    x = np.arange(N)
    delta = np.broadcast_to(np.cos(k0 * x)[:, None, None], (N, N, N)).copy()
    grid = lagrangian_grid((N, N, N), np.float64)
    F = pm.pm_forces(grid, delta=delta)
    a = (8 * np.sin(k0) - np.sin(2 * k0)) / 6
    np.testing.assert_allclose(F[..., 0], np.broadcast_to((-a / k0**2 * np.sin(k0 * x))[:, None, None],
                                                          (N, N, N)), atol=1e-12)
    np.testing.assert_allclose(F[..., 1:], 0, atol=1e-12)


def test_lpt1_plane_wave():
    N = 16
    cosmo = C.Planck15()
    k0 = 2 * np.pi / N
    x = np.arange(N)
    ic = 0.01 * np.broadcast_to(np.cos(k0 * x)[:, None, None], (N, N, N)).copy()
    dx, p, f = pm.lpt(cosmo, ic, a=0.1, order=1)
    a_ = (8 * np.sin(k0) - np.sin(2 * k0)) / 6
    expect = C.growth_factor(cosmo, 0.1) * (-a_ / k0**2) * 0.01 * np.sin(k0 * x)
    np.testing.assert_allclose(dx[:, 0, 0, 0], expect, atol=1e-12)
    np.testing.assert_allclose(p, 0.1**2 * C.growth_rate(cosmo, 0.1) * C.E(cosmo, 0.1) * dx, rtol=1e-10,
                               atol=1e-14)


@pytest.mark.parametrize("pdims", [(1, 2), (2, 1), (2, 2), (4, 2), (2, 4)])
def test_sharded_equals_unsharded(pdims):
    # mirrors tests/test_distributed_pm.py:26 (MSE < 1e-12 in f64); halo = mesh // 2 there
    shape = (16, 16, 8)
    rng = np.random.default_rng(0)
    disp = np.clip(rng.standard_normal((*shape, 3)), -1.9, 1.9)
    h = 4
    g, padded = D.cic_paint_dx(disp, h, pdims)
    assert U.MSE(g, P.cic_paint_dx(disp)) < 1e-24
    m = rng.standard_normal(shape)
    assert U.MSE(D.cic_read_dx(m, disp, h, pdims), P.cic_read_dx(m, disp)) < 1e-24
    for (rx, ry), b in padded.items():
        hx = h if pdims[0] > 1 else 0
        hy = h if pdims[1] > 1 else 0
        assert b.shape == (shape[0] // pdims[0] + 2 * hx, shape[1] // pdims[1] + 2 * hy, shape[2])


def test_owner_rank_rule():
    assert D.owner_rank(5, 9, (16, 16, 8), (2, 4)) == (0, 2)
    assert D.get_halo_size(8, (1, 4)) == (((0, 0), (8, 8), (0, 0)), (0, 4))
    assert D.get_halo_size((8, 6), (2, 2)) == (((8, 8), (6, 6), (0, 0)), (4, 3))


def test_growth_ode_matter_dominated_limits():
    eds = C.Cosmology(Omega_c=0.95, Omega_b=0.05, h=0.7, n_s=1.0, sigma8=0.8)
    for a in (0.1, 0.5, 1.0):  # Einstein-de Sitter: D1 = a, f1 = 1, f2 = 2, E = a^-1.5
        assert abs(C.growth_factor(eds, a) - a) < 1e-6
        assert abs(C.growth_rate(eds, a) - 1) < 1e-6
        assert abs(C.growth_rate_second(eds, a) - 2) < 1e-5
        assert abs(C.E(eds, a) - a**-1.5) < 1e-12
        assert abs(C.Gf(eds, a) - a**1.5) < 1e-5          # D' a^3 E = a^1.5
        assert abs(C.dGfa(eds, a) - 1.5 * a**0.5) < 1e-4


def test_power_spectrum_white_noise_and_parseval():
    N = 32
    rng = np.random.default_rng(0)
    f = rng.standard_normal((N, N, N))
    k, pk = U.power_spectrum(f, box_shape=(100., 100., 100.))
    assert abs(np.mean(pk) / (100.**3 / N**3) - 1) < 0.05     # white noise: P = V/Nc
    assert np.all(np.diff(k) > 0)


def test_kdk_integrators_run_and_agree_at_small_dt():
    N = 8
    cosmo = C.Planck15()
    grid, disp = displaced((N, N, N), 0.3)
    vel = np.zeros_like(disp)
    drift, kick = ode.symplectic_ode((N, N, N), cosmo)
    p1, v1 = ode.semi_implicit_euler(drift, kick, grid + disp, vel, 0.1, 0.12, 4)
    y = ode.leapfrog_midpoint(ode.make_diffrax_ode((N, N, N)), np.stack([grid + disp, vel]), 0.1, 0.12, 4,
                              cosmo)
    assert np.abs(p1 - y[0]).max() < 5e-3


def _fd4(psi, axis):
    """4th-order central difference of a periodic array."""
    r = lambda s: np.roll(psi, -s, axis=axis)
    return (2.0 / 3.0) * (r(1) - r(-1)) - (1.0 / 12.0) * (r(2) - r(-2))


def test_gradient_kernel_is_the_fd4_symbol():
    """The identity the potential force path of csrc/sim.cu rests on: the reference's gradient kernel
    i (8 sin w - sin 2w) / 6 (kernels.py:62-66) is the Fourier symbol of the 4th-order central difference, so
    IFFT(-gradient_kernel(d) * pot_k) (pm.py:54-56) == D_d psi with psi = -IFFT(pot_k).  Float64, oracle kernels."""
    from oracle import kernels as OK
    shape = (16, 12, 20)
    x = np.random.default_rng(0).standard_normal(shape)
    dk = OK.fft3d(x)
    kvec = OK.fftk(dk)
    pot = dk * OK.invlaplace_kernel(kvec)
    psi = -OK.ifft3d(pot)
    for d in range(3):
        ref = OK.ifft3d(-OK.gradient_kernel(kvec, d) * pot)
        assert np.abs(_fd4(psi, d) - ref).max() < 1e-12 * np.abs(ref).max()


def test_philox_known_answers_and_normal_field_moments():
    """oracle/rng.py (the restatement that pins the device Gaussian generator) against the published Random123
    known-answer vectors of philox4x32-10, and the moments of the Box-Muller field built on it."""
    from oracle import rng
    kat = {(0, 0): (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8),
           (0xffffffff, 0xffffffff): (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)}
    for (c, k), want in kat.items():
        got = rng.philox4x32_10(c, c, c, c, k, k)
        assert tuple(int(v) for v in got) == want
    got = rng.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert tuple(int(v) for v in got) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    z = rng.normal_field(5, (32, 32, 64))
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.01 and abs((z**4).mean() - 3) < 0.1   # n = 65536: 5 sigma
    # independent of how the mesh is traversed: a sub-block equals the slice of the full field by construction,
    # different seeds / streams decorrelate
    z2 = rng.normal_field(6, (32, 32, 64))
    assert abs(np.mean(z * z2)) < 0.02
    assert abs(np.mean(z * rng.normal_field(5, (32, 32, 64), stream_id=1))) < 0.02
