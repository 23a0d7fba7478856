"""GPU parity of the POTENTIAL force path and of the fast functional API.

* potential chain (csrc/pmfft.cu pmfft_potential): psi = IFFT(G delta_k / k^2) vs the float64 oracle, and the
  identity it rests on: the reference's gradient kernel i (8 sin w - sin 2w) / 6 (kernels.py:62-66) is the symbol
  of the 4th-order central difference, so D_d psi == IFFT(-gradient_kernel(d) * pot_k) (pm.py:54-56);
* resident step with force_mode = potential / auto vs spectral vs the oracle (gradient pass of csrc/pmfft.cu + the read kernel);
* pm_forces on the tile kernels (jpm_sim_forces) vs the oracle and vs the order-preserving kernels.
Tolerances as everywhere: fields 1e-5 relative (max-norm)."""
import numpy as np
import pytest
import torch

from helpers import displaced, lagrangian_grid, rel_err
from oracle import cosmology as OC
from oracle import kernels as OK
from oracle import ode as OO
from oracle import pm as OPM

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-5


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


def fd4(psi, axis):
    """4th-order central difference of a periodic array (float64)."""
    r = lambda s: np.roll(psi, -s, axis=axis)
    return (2.0 / 3.0) * (r(1) - r(-1)) - (1.0 / 12.0) * (r(2) - r(-2))


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 16, 64), (64, 128, 32), (256, 64, 128)])
def test_potential_chain(cuda, shape):
    from jaxpm_b200 import ops
    from jaxpm_b200.kernels import pgd_filter_table
    rng = np.random.default_rng(3)
    x = rng.standard_normal(shape).astype(np.float32)
    plan = ops.get_plan(shape, cuda)
    dk = OK.fft3d(x.astype(np.float64))
    kvec = OK.fftk(dk)
    for r_split, filt, tab, tol in ((0.0, None, None, FIELD_TOL),
                                    (1.3, OK.PGD_kernel(kvec, 0.4, 2.5), pgd_filter_table(0.4, 2.5, 1 << 16), 1e-4)):
        pot = dk * OK.invlaplace_kernel(kvec) * OK.longrange_kernel(kvec, r_split)
        if filt is not None:
            pot = pot * filt
        ref = -OK.ifft3d(pot)
        psi = ops.potential_from_density_fused(T(x, cuda), plan, r_split, tab).cpu().numpy()
        assert rel_err(psi, ref) < tol, r_split
        # differences of the fp32 mesh == the three-transform force meshes
        for d in range(3):
            fref = OK.ifft3d(-OK.gradient_kernel(kvec, d) * pot)
            assert rel_err(fd4(psi.astype(np.float64), d), fref) < 10 * tol, (d, r_split)


@pytest.mark.parametrize("relative", [False, True])
@pytest.mark.parametrize("shape,tile", [((32, 32, 32), 8), ((32, 64, 32), 16), ((64, 64, 64), 16)])
def test_sim_step_potential(cuda, relative, shape, tile):
    """K resident steps with the potential force path (one inverse transform, one real-space gradient pass, the
    ordinary read kernel) == spectral path == oracle, including particles at the periodic edges (generic stencil
    inside the box) and beyond the margin (global-memory fallback)."""
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import nbody_kick_drift
    grid, disp = displaced(shape, 1.0)
    x = disp if relative else grid + disp
    vel = (0.3 * np.random.default_rng(9).standard_normal(x.shape)).astype(np.float32)
    vel[3, 4, 5] = (40.0, -35.0, 50.0)        # leaves its box within one step: global-memory fallback
    cosmo, ocos = Planck15(), OC.Planck15()
    drift, kick = OO.symplectic_ode(shape, ocos, paint_absolute_pos=not relative)
    rp, rv = OO.semi_implicit_euler(drift, kick, x, vel, 0.5, 0.8, 3)
    res = {}
    for mode in ("spectral", "potential", "auto"):
        info = {}
        p, v = nbody_kick_drift(cosmo, T(x, cuda), T(vel, cuda), 0.5, 0.8, 3, mesh_shape=shape,
                                paint_absolute_pos=not relative, tile=tile, margin=1, force_mode=mode, info=info)
        res[mode] = (p.cpu().numpy(), v.cpu().numpy(), info)
        assert np.abs(res[mode][0] - rp).max() < 2e-4, mode
        assert rel_err(res[mode][1], rv) < 1e-4, mode
    assert res["potential"][2]["steps_potential"] == 3 and res["potential"][2]["steps_spectral"] == 0
    assert res["spectral"][2]["steps_spectral"] == 3
    assert res["potential"][2]["fallbacks"][1] >= 1      # the fast particle took the global path of the read
    # white-noise displacements give a rough density: psi is small against F, the two paths agree to rounding
    assert rel_err(res["potential"][1], res["spectral"][1]) < 2e-5


def test_force_mode_auto_follows_the_error_bound(cuda):
    """A smooth, large-scale density (|psi| >> |F|: differencing psi in fp32 would cost accuracy) keeps AUTO on the
    three-transform chain; a rough one lets it switch to the potential chain after the first measured step."""
    from jaxpm_b200 import ops
    shape = (64, 64, 64)
    grid = lagrangian_grid(shape)
    smooth = np.zeros((*shape, 3), np.float32)
    smooth[..., 0] = 0.3 * np.sin(2 * np.pi * grid[..., 0] / shape[0])      # one fundamental mode
    rough = (1.0 * np.random.default_rng(4).standard_normal((*shape, 3))).astype(np.float32)
    out = {}
    for name, disp in (("smooth", smooth), ("rough", rough)):
        sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=1)
        sim.set_force_mode("auto")
        sim.load(T(disp, cuda), torch.zeros((*shape, 3), device=cuda))
        for _ in range(4):
            sim.step(1e-6, 1e-6)
            torch.cuda.synchronize()
        out[name] = sim.force_info()
    assert out["smooth"]["steps_potential"] == 0 and out["smooth"]["error_bound"] > 6e-6, out["smooth"]
    assert out["rough"]["steps_potential"] >= 2 and 0 < out["rough"]["error_bound"] < 4e-6, out["rough"]


def test_potential_mode_availability(cuda):
    """Margin 2 runs the potential chain through the gradient pass; a mesh the fused FFT chain does not serve
    (not a power of two) refuses the mode loudly instead of silently falling back."""
    from jaxpm_b200 import ops
    from jaxpm_b200._lib import JpmError
    shape = (32, 32, 32)
    _, disp = displaced(shape, 1.0)
    vel = np.zeros_like(disp)
    out = {}
    for mode in ("spectral", "potential"):
        sim = ops.Sim(shape, shape, True, cuda, tile=8, margin=2)
        sim.set_force_mode(mode)
        sim.load(T(disp, cuda), T(vel, cuda))
        sim.step(1.0, 0.0)
        p, v = torch.empty(disp.shape, device=cuda), torch.empty(disp.shape, device=cuda)
        sim.store(p, v)
        out[mode] = v.cpu().numpy()
    assert rel_err(out["potential"], out["spectral"]) < FIELD_TOL
    sim24 = ops.Sim((24, 40, 20), (24, 40, 20), True, cuda, tile=8, margin=1)     # not a power-of-two mesh
    with pytest.raises(JpmError):
        sim24.set_force_mode("auto")


# ---- pm_forces on the tile kernels -----------------------------------------------------------------------
@pytest.mark.parametrize("absolute", [True, False])
@pytest.mark.parametrize("shape,sigma", [((16, 16, 16), 1.0), ((32, 32, 64), 3.0), ((64, 64, 64), 6.0)])
def test_pm_forces_fast_api(cuda, shape, sigma, absolute):
    """pm_forces (no gradient requested, power-of-two mesh) runs tile-sort -> shared-memory paint -> fused FFT chain
    -> shared-memory gather; same forces, in the caller's particle order, as the oracle and the slow kernels."""
    from jaxpm_b200 import _lib, pm
    from jaxpm_b200.kernels import pgd_filter_table
    grid, disp = displaced(shape, sigma)
    disp[0, 0, 0] = (-1e-7, 0.3, -0.2)
    x = grid + disp if absolute else disp
    ref = OPM.pm_forces(x, mesh_shape=shape, paint_absolute_pos=absolute)
    n0 = _lib.launch_count()
    got = pm.pm_forces(T(x, cuda), mesh_shape=shape, paint_absolute_pos=absolute).cpu().numpy()
    assert got.shape == (*shape, 3)
    assert rel_err(got, ref) < FIELD_TOL
    assert _lib.launch_count() - n0 <= 12          # count, scan, fill, paint, 5 FFT passes, forces
    try:
        pm._FAST_API = False
        slow = pm.pm_forces(T(x, cuda), mesh_shape=shape, paint_absolute_pos=absolute).cpu().numpy()
    finally:
        pm._FAST_API = True
    assert rel_err(got, slow) < FIELD_TOL
    ref = OPM.pm_forces(x, mesh_shape=shape, paint_absolute_pos=absolute, r_split=1.3,
                        kfilter=OK.PGD_kernel(OK.fftk(shape), 0.4, 2.5))
    got = pm.pm_forces(T(x, cuda), mesh_shape=shape, paint_absolute_pos=absolute, r_split=1.3,
                       filter_tab=pgd_filter_table(0.4, 2.5, 1 << 16)).cpu().numpy()
    assert rel_err(got, ref) < 1e-4


def test_pm_forces_fast_api_unstructured_and_repeat(cuda):
    """Absolute positions as a flat list with np != ncell, called twice on different particle sets (the cached
    positions-only state is re-loaded), and the differentiable call still takes the adjoint-capable path."""
    from jaxpm_b200 import pm
    shape = (32, 32, 32)
    rng = np.random.default_rng(21)
    for n in (5000, 5000, 777):
        pos = rng.uniform(-3, 35, (n, 3)).astype(np.float32)
        ref = OPM.pm_forces(pos, mesh_shape=shape)
        got = pm.pm_forces(T(pos, cuda), mesh_shape=shape).cpu().numpy()
        assert got.shape == (n, 3)
        assert rel_err(got, ref) < FIELD_TOL
    x = T(pos, cuda).requires_grad_(True)
    f = pm.pm_forces(x, mesh_shape=shape)
    f.square().sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    assert rel_err(f.detach().cpu().numpy(), ref) < FIELD_TOL
