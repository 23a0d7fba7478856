"""Pin the CPU oracle against tests/golden/*.npz — outputs of the reference's own source files
(/root/reference/jaxpm, run unmodified on the NumPy stand-in for jax by oracle/_refrun/make_golden.py).

Integer results (cell indices, halo extents, particle->rank) must match bit for bit; CIC weights
bit for bit (same fp32 operation order); fields to 2e-6 relative (the fixture's scatter / 8-term sums
are sequential fp32, the oracle accumulates in float64)."""
import os

import numpy as np
import pytest

from helpers import rel_err
from oracle import cosmology as C
from oracle import distributed as D
from oracle import kernels as K
from oracle import ode as O
from oracle import painting as P
from oracle import pm as PM
from oracle import utils as U

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-6


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_absolute_paint_read_indices_weights_fields():
    g = gold("paint_read_abs")
    shape = g["base"].shape
    idx, w = P.cic_indices_weights(g["pos"], shape)
    np.testing.assert_array_equal(idx, g["idx"])                 # integer contract: bit-exact
    np.testing.assert_array_equal(w, g["kernel"])                # (kx*ky)*kz in fp32: bit-exact
    assert rel_err(P.cic_paint(np.zeros(shape, np.float32), g["pos"]), g["mesh_w1"]) < TOL
    assert rel_err(P.cic_paint(g["base"], g["pos"], g["weight"]), g["mesh_warr_on_base"]) < TOL
    assert rel_err(P.cic_paint(g["base"], g["pos"], 2.5), g["mesh_w2p5_on_base"]) < TOL
    assert rel_err(P.cic_read(g["base"], g["pos"]), g["read_base"]) < TOL


@pytest.mark.parametrize("tag,halo", [("h00", (0, 0)), ("h23", (2, 3))])
def test_relative_paint_read_indices_weights_fields(tag, halo):
    g = gold("paint_read_rel")
    disp = g["disp"]
    shp = disp.shape[:3]
    pshape = (shp[0] + 2 * halo[0], shp[1] + 2 * halo[1], shp[2])
    idx, w = P.enmesh_rel(P._pmid(shp, *halo), disp.reshape(-1, 3), pshape)
    np.testing.assert_array_equal(idx, g[f"idx_{tag}"])          # incl. the out-of-range index N
    np.testing.assert_array_equal(w, g[f"w_{tag}"])
    assert (g[f"idx_{tag}"] == np.asarray(pshape)).any() or tag != "h00"   # the edge case is in the fixture
    assert rel_err(P.cic_paint_dx_padded(disp, 1.0, halo), g[f"mesh_{tag}"]) < TOL
    assert rel_err(P.cic_paint_dx_padded(disp, g["weight"], halo), g[f"mesh_warr_{tag}"]) < TOL
    assert rel_err(P.cic_read_dx_padded(g[f"field_{tag}"], disp, halo), g[f"read_{tag}"]) < TOL
    if tag == "h00":
        assert rel_err(P.cic_paint_dx(disp), g["paint_dx_api"]) < TOL
        assert rel_err(P.cic_read_dx(g["field_h00"], disp), g["read_dx_api"]) < TOL


def test_kspace_kernels():
    g = gold("kernels")
    shape = g["invlap"].shape
    kvec = K.fftk(shape, dtype=np.float32)
    for d in range(3):
        np.testing.assert_array_equal(kvec[d], g[f"k{d}"])
        np.testing.assert_allclose(K.gradient_kernel(kvec, d), g[f"grad{d}_o1"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(K.gradient_kernel(kvec, d, order=0), g[f"grad{d}_o0"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(K.invlaplace_kernel(kvec), g["invlap"], rtol=1e-6)
    np.testing.assert_allclose(K.invlaplace_kernel(kvec, fd=True), g["invlap_fd"], rtol=2e-6)
    assert K.longrange_kernel(kvec, 0) == float(g["longrange_r0"]) == 1.0
    np.testing.assert_allclose(K.longrange_kernel(kvec, 1.5), g["longrange_r1p5"], rtol=1e-5, atol=1e-30)
    np.testing.assert_allclose(K.cic_compensation(kvec), g["cic_comp"], rtol=1e-5)


def test_pm_forces():
    g = gold("pm_forces")
    shape = g["delta"].shape
    tol = 5e-6   # two fp32 FFTs in both; scipy (oracle) vs scipy (fixture) with different fp32 paint sums
    assert rel_err(PM.pm_forces(g["pos"], mesh_shape=shape), g["f_abs"]) < tol
    assert rel_err(PM.pm_forces(g["disp"], mesh_shape=shape, paint_absolute_pos=False), g["f_rel"]) < tol
    assert rel_err(PM.pm_forces(g["pos"], mesh_shape=shape, r_split=2.0), g["f_abs_rsplit2"]) < tol
    assert rel_err(PM.pm_forces(g["pos"], delta=g["delta"]), g["f_abs_delta_real"]) < tol
    assert rel_err(PM.pm_forces(g["disp"], delta=K.fft3d(g["delta"]), paint_absolute_pos=False),
                   g["f_rel_delta_cplx"]) < tol


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("mode", ["rel", "abs"])
def test_lpt(order, mode):
    g = gold("lpt")
    ic = g["ic"]
    part = None
    if mode == "abs":
        part = np.stack(np.meshgrid(*[np.arange(s) for s in ic.shape], indexing="ij"), -1).astype(np.float32)
    dx, p, f = PM.lpt(C.Planck15(), ic, particles=part, a=float(g["a"]), order=order)
    for got, name in ((dx, "dx"), (p, "p"), (f, "f")):
        assert rel_err(got, g[f"{mode}_o{order}_{name}"]) < 5e-6, name


def test_growth_scalars_and_ode_terms():
    g = gold("growth_ode")
    cosmo = C.Planck15()
    a = g["a"]
    for name in ("E", "dEa", "gp", "Gf", "Gf2", "dGfa", "dGf2a", "growth_factor", "growth_rate",
                 "growth_factor_second", "growth_rate_second"):
        # the fixture ran the reference's formulas in fp32 (x64 off); the oracle's are float64
        np.testing.assert_allclose(getattr(C, name)(cosmo, a), g["g_" + name], rtol=3e-6, err_msg=name)
    shape = g["pos"].shape[:3]
    pos, vel, a0, dt0 = g["pos"], g["vel"], float(g["ode_a"]), float(g["fpm_dt0"])
    dpos, dvel = O.make_ode_fn(shape)((pos, vel), a0, cosmo)
    assert rel_err(dpos, g["ode_dpos"]) < 5e-6 and rel_err(dvel, g["ode_dvel"]) < 5e-6
    assert rel_err(O.make_diffrax_ode(shape)(a0, np.stack([pos, vel]), cosmo), g["diffrax_rhs"]) < 5e-6
    drift, kick = O.symplectic_ode(shape, cosmo)
    assert rel_err(drift(a0, vel, None), g["sym_drift"]) < 5e-6
    assert rel_err(kick(a0, pos, None), g["sym_kick"]) < 5e-6
    drift, kick, first = O.symplectic_fpm_ode(shape, dt0, cosmo)
    assert rel_err(drift(a0, vel, None), g["fpm_drift"]) < 2e-5
    assert rel_err(kick(a0, pos, None), g["fpm_kick"]) < 2e-5
    assert rel_err(first(a0, pos, cosmo), g["fpm_first_kick"]) < 2e-5


@pytest.mark.parametrize("pd", [(2, 2), (1, 4), (4, 1), (2, 4)])
def test_sharded_protocol(pd):
    g = gold("distributed")
    tag, halo = f"p{pd[0]}{pd[1]}", int(g["halo"])
    shape = g["field"].shape
    hs, ext = D.get_halo_size((halo, halo), pd)
    np.testing.assert_array_equal(np.asarray(hs), g[f"{tag}_halo_size"])     # integer contract
    np.testing.assert_array_equal(np.asarray(ext), g[f"{tag}_halo_ext"])
    np.testing.assert_array_equal(D.get_local_shape(shape, pd), g[f"{tag}_local_shape"])
    mesh, _ = D.cic_paint_dx(g["disp"], (halo, halo), pd)
    assert rel_err(mesh, g[f"{tag}_paint"]) < TOL
    assert rel_err(D.cic_read_dx(g["field"], g["disp"], (halo, halo), pd), g[f"{tag}_read"]) < TOL
    # the reference itself: sharded == single device (tests/test_distributed_pm.py:176-179)
    assert rel_err(g[f"{tag}_paint"], g["single_paint"]) < TOL
    assert rel_err(g[f"{tag}_read"], g["single_read"]) < TOL
    assert rel_err(g[f"{tag}_forces"], g["single_forces"]) < 5e-6
    # particle -> rank: the sharded Lagrangian grid is the global grid cut into blocks
    np.testing.assert_array_equal(g[f"{tag}_particles"], g["single_particles"])
    i, j = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    rx, ry = D.owner_rank(i, j, shape, pd)
    lx, ly = shape[0] // pd[0], shape[1] // pd[1]
    np.testing.assert_array_equal(rx, g[f"{tag}_particles"][:, :, 0, 0].astype(int) // lx)
    np.testing.assert_array_equal(ry, g[f"{tag}_particles"][:, :, 0, 1].astype(int) // ly)


@pytest.mark.parametrize("pd", [(2, 2), (2, 4)])
def test_sharded_protocol_pencil_fixture(pd):
    """The oracle's sharded paint against the pencil fixture the GPU pencil path is judged on
    (tests/golden/distributed_pencil.npz, 32 x 64 x 16, halo 8; tests/test_gpu_golden.py::test_golden_pencil_path)."""
    g = gold("distributed_pencil")
    tag, halo = f"p{pd[0]}{pd[1]}", int(g["halo"])
    mesh, _ = D.cic_paint_dx(g["disp"], (halo, halo), pd)
    assert rel_err(mesh, g[f"{tag}_paint"]) < TOL
    assert rel_err(g[f"{tag}_paint"], g["single_paint"]) < 1e-5
    shape = g["disp"].shape[:3]
    i, j = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    rx, ry = D.owner_rank(i, j, shape, pd)
    np.testing.assert_array_equal(rx, g[f"{tag}_particles"][:, :, 0, 0].astype(int) // (shape[0] // pd[0]))
    np.testing.assert_array_equal(ry, g[f"{tag}_particles"][:, :, 0, 1].astype(int) // (shape[1] // pd[1]))


def test_slice_unpad_rule():
    g = gold("distributed")
    np.testing.assert_allclose(D.slice_unpad_impl(g["unpad_in"], ((4, 4), (4, 4), (0, 0))), g["unpad_out_h44"],
                               rtol=0, atol=0)
    np.testing.assert_allclose(D.slice_unpad_impl(g["unpad_in_h40"], ((4, 4), (0, 0), (0, 0))),
                               g["unpad_out_h40"], rtol=0, atol=0)


def test_power_spectrum():
    g = gold("power_spectrum")
    box = tuple(g["box"])
    k, pk = U.power_spectrum(g["f1"], box_shape=box)
    np.testing.assert_allclose(k, g["k"], rtol=1e-6)
    np.testing.assert_allclose(pk, g["pk"], rtol=2e-5)
    _, pkx = U.power_spectrum(g["f1"], g["f2"], box_shape=box)
    np.testing.assert_allclose(pkx, g["pk_cross"], rtol=2e-5)
    kp, pkl = U.power_spectrum(g["f1"], box_shape=box, multipoles=[0, 2], kedges=5)
    np.testing.assert_allclose(kp, g["k_poles"], rtol=1e-6)
    np.testing.assert_allclose(pkl, g["pk_poles"], rtol=2e-4, atol=1e-3 * np.abs(g["pk_poles"]).max())
    kc, pkc = U.power_spectrum(g["f1"])
    np.testing.assert_allclose(pkc, g["pk_cell"], rtol=2e-5)


def test_widened_rows():
    """SURVEY.md §8f rows 3-4 as computed by the reference's own source: compensate_cic, cic_paint_2d."""
    from oracle import kernels as K2
    from oracle import painting as P2
    g = gold("widened")
    dk = K2.fft3d(g["field"].astype(np.float64))
    comp = K2.ifft3d(K2.cic_compensation(K2.fftk(dk)) * dk)
    assert np.abs(comp - g["compensated"]).max() / np.abs(g["compensated"]).max() < 2e-6
    for base, w, key in ((g["base2"], g["w2"], "mesh2_weighted"), (np.zeros_like(g["base2"]), None, "mesh2_unit")):
        got = P2.cic_paint_2d(base, g["pos2"], w)
        assert np.abs(got - g[key]).max() / np.abs(g[key]).max() < 2e-6
    dp = P2.density_plane(g["pos3"], (16, 16, 16), 8.0, 4.0, 12)      # lensing.py:11-44
    assert np.abs(dp - g["density_plane"]).max() / np.abs(g["density_plane"]).max() < 2e-6
