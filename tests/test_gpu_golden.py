"""GPU parity against the committed golden fixtures (tests/golden/*.npz): the CUDA path, through the
C ABI, compared DIRECTLY with what the reference's own source computed (no oracle in between).
Bars (BASELINE.json north_star): cell indices bit-exact, fields 1e-5 relative (fp32)."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELD_TOL = 1e-5


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def test_golden_cell_indices_bit_exact(cuda):
    from jaxpm_b200 import ops
    g = gold("paint_read_abs")
    shape = g["base"].shape
    c0 = g["idx"][:, 0]                                           # corner (0,0,0) of every particle
    flat = (c0[:, 0] * shape[1] + c0[:, 1]) * shape[2] + c0[:, 2]
    np.testing.assert_array_equal(N(ops.cell_index(T(g["pos"], cuda), shape)), flat)
    r = gold("paint_read_rel")
    for tag, halo in (("h00", (0, 0)), ("h23", (2, 3))):
        shp = r["disp"].shape[:3]
        pshape = (shp[0] + 2 * halo[0], shp[1] + 2 * halo[1], shp[2])
        c0 = r[f"idx_{tag}"][:, 0]
        ok = np.all((c0 >= 0) & (c0 < np.asarray(pshape)), axis=-1)
        flat = np.where(ok, (c0[:, 0] * pshape[1] + c0[:, 1]) * pshape[2] + c0[:, 2], -1)
        got = N(ops.cell_index(T(r["disp"], cuda), pshape, relative=True, halo=halo))
        np.testing.assert_array_equal(got, flat)


def test_golden_paint_read_absolute(cuda):
    from jaxpm_b200.painting import cic_paint, cic_read
    g = gold("paint_read_abs")
    shape = g["base"].shape
    pos = T(g["pos"], cuda)
    assert rel_err(N(cic_paint(torch.zeros(shape, device=cuda), pos)), g["mesh_w1"]) < FIELD_TOL
    assert rel_err(N(cic_paint(T(g["base"], cuda), pos, T(g["weight"], cuda))), g["mesh_warr_on_base"]) < FIELD_TOL
    assert rel_err(N(cic_paint(T(g["base"], cuda), pos, 2.5)), g["mesh_w2p5_on_base"]) < FIELD_TOL
    assert rel_err(N(cic_read(T(g["base"], cuda), pos)), g["read_base"]) < FIELD_TOL


@pytest.mark.parametrize("tag,halo", [("h00", (0, 0)), ("h23", (2, 3))])
def test_golden_paint_read_relative(cuda, tag, halo):
    from jaxpm_b200 import ops
    from jaxpm_b200.painting import cic_paint_dx, cic_read_dx
    g = gold("paint_read_rel")
    disp = T(g["disp"], cuda)
    pshape = g[f"mesh_{tag}"].shape
    m = ops.cic_paint_dx_(torch.zeros(pshape, device=cuda), disp, 1.0, halo)
    assert rel_err(N(m), g[f"mesh_{tag}"]) < FIELD_TOL
    m = ops.cic_paint_dx_(torch.zeros(pshape, device=cuda), disp, T(g["weight"], cuda), halo)
    assert rel_err(N(m), g[f"mesh_warr_{tag}"]) < FIELD_TOL
    assert rel_err(N(ops.cic_read_dx(T(g[f"field_{tag}"], cuda), disp, halo)), g[f"read_{tag}"]) < FIELD_TOL
    if tag == "h00":
        assert rel_err(N(cic_paint_dx(disp)), g["paint_dx_api"]) < FIELD_TOL
        assert rel_err(N(cic_read_dx(T(g["field_h00"], cuda), disp)), g["read_dx_api"]) < FIELD_TOL


def test_golden_resident_state_paint_read(cuda):
    """The tile-sorted resident path (csrc/sim.cu) against the same fixtures."""
    from jaxpm_b200 import ops
    g = gold("paint_read_rel")
    disp = T(g["disp"], cuda)
    shp = tuple(disp.shape[:3])
    sim = ops.Sim(shp, shp, True, cuda, tile=8, margin=2, with_plan=False)
    sim.load(disp, torch.zeros_like(disp))
    m = sim.paint_(torch.zeros(shp, device=cuda))
    assert rel_err(N(m), g["mesh_h00"]) < FIELD_TOL
    a = gold("paint_read_abs")
    pos = T(a["pos"], cuda)
    sim = ops.Sim(shp, shp, False, cuda, tile=8, margin=2, with_plan=False)
    sim.load(pos, torch.zeros_like(pos))
    m = sim.paint_(torch.zeros(shp, device=cuda))
    assert rel_err(N(m), a["mesh_w1"]) < FIELD_TOL


def test_golden_pm_forces(cuda):
    from jaxpm_b200.distributed import fft3d
    from jaxpm_b200.pm import pm_forces
    g = gold("pm_forces")
    shape = g["delta"].shape
    pos, disp, delta = T(g["pos"], cuda), T(g["disp"], cuda), T(g["delta"], cuda)
    assert rel_err(N(pm_forces(pos, mesh_shape=shape)), g["f_abs"]) < FIELD_TOL
    assert rel_err(N(pm_forces(disp, mesh_shape=shape, paint_absolute_pos=False)), g["f_rel"]) < FIELD_TOL
    assert rel_err(N(pm_forces(pos, mesh_shape=shape, r_split=2.0)), g["f_abs_rsplit2"]) < FIELD_TOL
    assert rel_err(N(pm_forces(pos, delta=delta)), g["f_abs_delta_real"]) < FIELD_TOL
    assert rel_err(N(pm_forces(disp, delta=fft3d(delta), paint_absolute_pos=False)), g["f_rel_delta_cplx"]) < FIELD_TOL


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("mode", ["rel", "abs"])
def test_golden_lpt(cuda, order, mode):
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.distributed import uniform_particles
    from jaxpm_b200.pm import lpt
    g = gold("lpt")
    ic = T(g["ic"], cuda)
    part = uniform_particles(ic.shape, device=cuda) if mode == "abs" else None
    dx, p, f = lpt(Planck15(), ic, particles=part, a=float(g["a"]), order=order)
    for got, name in ((dx, "dx"), (p, "p"), (f, "f")):
        assert rel_err(N(got), g[f"{mode}_o{order}_{name}"]) < FIELD_TOL, name


def test_golden_growth_and_ode_terms(cuda):
    from jaxpm_b200 import growth
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import make_diffrax_ode, make_ode_fn, symplectic_fpm_ode, symplectic_ode
    g = gold("growth_ode")
    cosmo = Planck15()
    for name in ("E", "dEa", "gp", "Gf", "Gf2", "dGfa", "dGf2a", "growth_factor", "growth_rate",
                 "growth_factor_second", "growth_rate_second"):
        np.testing.assert_allclose(getattr(growth, name)(cosmo, g["a"]), g["g_" + name], rtol=2e-5, err_msg=name)
    shape = g["pos"].shape[:3]
    pos, vel, a0, dt0 = T(g["pos"], cuda), T(g["vel"], cuda), float(g["ode_a"]), float(g["fpm_dt0"])
    dpos, dvel = make_ode_fn(shape)((pos, vel), a0, cosmo)
    assert rel_err(N(dpos), g["ode_dpos"]) < FIELD_TOL and rel_err(N(dvel), g["ode_dvel"]) < FIELD_TOL
    assert rel_err(N(make_diffrax_ode(shape)(a0, torch.stack([pos, vel]), cosmo)), g["diffrax_rhs"]) < FIELD_TOL
    drift, kick = symplectic_ode(shape, cosmo)
    assert rel_err(N(drift(a0, vel, None)), g["sym_drift"]) < FIELD_TOL
    assert rel_err(N(kick(a0, pos, None)), g["sym_kick"]) < FIELD_TOL
    drift, kick, first = symplectic_fpm_ode(shape, dt0, cosmo)
    assert rel_err(N(drift(a0, vel, None)), g["fpm_drift"]) < 3e-5     # growth-table (host scalar) tolerance
    assert rel_err(N(kick(a0, pos, None)), g["fpm_kick"]) < 3e-5
    assert rel_err(N(first(a0, pos, cosmo)), g["fpm_first_kick"]) < 3e-5


def test_golden_power_spectrum_and_widened_rows(cuda):
    """Device power_spectrum, compensate_cic and cic_paint_2d directly against what the reference's own source
    produced (tests/golden/power_spectrum.npz, widened.npz)."""
    from jaxpm_b200.painting import cic_paint_2d, compensate_cic
    from jaxpm_b200.utils import power_spectrum
    g = gold("power_spectrum")
    box = tuple(float(b) for b in g["box"])
    k, pk = power_spectrum(T(g["f1"], cuda), box_shape=box)
    np.testing.assert_allclose(k, g["k"], rtol=1e-5)
    np.testing.assert_allclose(N(pk), g["pk"], rtol=1e-4)
    _, pkx = power_spectrum(T(g["f1"], cuda), T(g["f2"], cuda), box_shape=box)
    np.testing.assert_allclose(N(pkx), g["pk_cross"], rtol=1e-4)
    kp, pkl = power_spectrum(T(g["f1"], cuda), box_shape=box, multipoles=[0, 2], kedges=5)
    np.testing.assert_allclose(kp, g["k_poles"], rtol=1e-5)
    assert np.abs(N(pkl) - g["pk_poles"]).max() < 1e-4 * np.abs(g["pk_poles"][0]).max()
    kc, pkc = power_spectrum(T(g["f1"], cuda))
    np.testing.assert_allclose(N(pkc), g["pk_cell"], rtol=1e-4)
    w = gold("widened")
    assert rel_err(N(compensate_cic(T(w["field"], cuda))), w["compensated"]) < FIELD_TOL
    assert rel_err(N(cic_paint_2d(T(w["base2"], cuda), T(w["pos2"], cuda), T(w["w2"], cuda))), w["mesh2_weighted"]) < FIELD_TOL
    assert rel_err(N(cic_paint_2d(torch.zeros(tuple(w["base2"].shape), device=cuda), T(w["pos2"], cuda), None)),
                   w["mesh2_unit"]) < FIELD_TOL
    from jaxpm_b200.lensing import density_plane
    assert rel_err(N(density_plane(T(w["pos3"], cuda), (16, 16, 16), 8.0, 4.0, 12)), w["density_plane"]) < FIELD_TOL


# ---- multi-GPU decompositions against the reference-source fixtures ------------------------
def _padded_density(plan, cuda):
    """The rank's ghost-zone density array [nxp][nyp][nzp] (painted, ghosts not folded)."""
    import ctypes as C
    from jaxpm_b200._lib import call, ptr, stream
    dims = (C.c_int32 * 3)()
    call("jpm_plan_padded_get_f32", plan.handle, stream(), 0, None, dims)
    out = torch.empty(tuple(int(d) for d in dims), dtype=torch.float32, device=cuda)
    call("jpm_plan_padded_get_f32", plan.handle, stream(), 0, ptr(out), dims)
    return out


@pytest.mark.parametrize("P", [2, 4])
def test_golden_slab_path(cuda, P):
    """The fused peer-memory slab path (jaxpm_b200/slab.py, csrc/pmfft.cu, csrc/sim.cu) with P ranks in this process,
    fed the inputs of tests/golden/distributed_slab.npz and compared DIRECTLY with what the reference's own sharded
    cic_paint_dx / pm_forces computed for pdims (P, 1), halo 8 (jaxpm/painting.py:192-215, pm.py:12-58,
    distributed.py:45-113): the painted density (per-rank ghost-zone arrays folded on the host) and the force on
    every particle (one kick of unit coefficient from zero velocity)."""
    from jaxpm_b200.slab import SlabPlan, SlabStepper
    g = gold("distributed_slab")
    disp, gx = g["disp"], int(g["halo"])
    shape = disp.shape[:3]
    nx, ny, nz = shape
    lx = nx // P
    plans = [SlabPlan(shape, P, r, gx, cuda) for r in range(P)]
    for p in plans:
        p.attach_local(plans)
    streams = [torch.cuda.Stream(cuda) for _ in range(P)]
    dl = [T(disp[r * lx:(r + 1) * lx], cuda) for r in range(P)]
    vl = [torch.zeros_like(d) for d in dl]
    torch.cuda.synchronize()
    st = []
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st.append(SlabStepper(dl[r], vl[r], gx, P, r, tile=8, margin=1, plan=plans[r]))
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st[r].step(1.0, 0.0)                     # vel = 0 + 1 * F(disp); no drift
    G = 4
    rho = np.zeros(shape, np.float64)
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st[r].store(dl[r], vl[r])
            a = N(_padded_density(plans[r], cuda)).astype(np.float64)
        # local plane xl of the rank's mesh (its slab + gx ghost planes per side) is global plane r lx - gx + xl;
        # y / z carry G periodic ghost cells per side; the kGhost spare planes around the x range stay empty
        xs = (r * lx - gx + np.arange(lx + 2 * gx)) % nx
        ys = (np.arange(ny + 2 * G) - G) % ny
        zs = (np.arange(nz + 2 * G) - G) % nz
        assert a[:G].sum() == 0 and a[G + lx + 2 * gx:].sum() == 0
        np.add.at(rho, (xs[:, None, None], ys[None, :, None], zs[None, None, :]), a[G:G + lx + 2 * gx])
    torch.cuda.synchronize()
    assert rel_err(rho, g[f"p{P}1_paint"]) < FIELD_TOL
    forces = np.concatenate([N(v) for v in vl])
    assert rel_err(forces, g[f"p{P}1_forces"]) < FIELD_TOL
    np.testing.assert_array_equal(np.concatenate([N(d) for d in dl]), disp)       # zero drift: positions untouched
    for s_ in st:
        s_.close(barrier=False)


@pytest.mark.parametrize("pdims", [(2, 2), (2, 4)])
def test_golden_pencil_path(cuda, pdims):
    """The fused peer-memory path on PENCIL process grids (pencil particle domains, x-slab FFT chain, row-group transpose
    inside the z passes; jaxpm_b200/slab.py, csrc/pmfft.cu) with px py ranks in this process, fed the inputs of
    tests/golden/distributed_pencil.npz and compared DIRECTLY with what the reference's own sharded cic_paint_dx /
    pm_forces computed for these pdims, halo 8 (jaxpm/painting.py:192-215, pm.py:12-58, distributed.py:45-129;
    tests/test_distributed_pm.py:28): the painted density (per-rank ghost-zone arrays folded on the host) and the
    force on every particle (one kick of unit coefficient from zero velocity)."""
    from jaxpm_b200.slab import SlabPlan, SlabStepper
    g = gold("distributed_pencil")
    disp, h = g["disp"], int(g["halo"])
    shape = disp.shape[:3]
    nx, ny, nz = shape
    px, py = pdims
    P = px * py
    Lx, Ly = nx // px, ny // py
    blk = lambda a_: [a_[rx * Lx:(rx + 1) * Lx, ry * Ly:(ry + 1) * Ly] for rx in range(px) for ry in range(py)]
    plans = [SlabPlan(shape, P, r, h, cuda, pdims=pdims, gy=h) for r in range(P)]
    for p in plans:
        p.attach_local(plans)
    streams = [torch.cuda.Stream(cuda) for _ in range(P)]
    dl = [T(b, cuda) for b in blk(disp)]
    vl = [torch.zeros_like(d) for d in dl]
    torch.cuda.synchronize()
    st = []
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st.append(SlabStepper(dl[r], vl[r], h, P, r, tile=8, margin=1, plan=plans[r], pdims=pdims, gy=h))
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            st[r].step(1.0, 0.0)                     # vel = 0 + 1 * F(disp); no drift
    G = 4
    rho = np.zeros(shape, np.float64)
    for r in range(P):
        rx, ry = divmod(r, py)
        with torch.cuda.stream(streams[r]):
            st[r].store(dl[r], vl[r])
            a = N(_padded_density(plans[r], cuda)).astype(np.float64)
        # local (xl, yl) of the rank's mesh (its pencil + h ghost planes / rows per side) is global
        # (rx Lx - h + xl, ry Ly - h + yl); z carries G periodic ghost cells per side; the G spare planes / rows stay empty
        xs = (rx * Lx - h + np.arange(Lx + 2 * h)) % nx
        ys = (ry * Ly - h + np.arange(Ly + 2 * h)) % ny
        zs = (np.arange(nz + 2 * G) - G) % nz
        assert a[:G].sum() == 0 and a[G + Lx + 2 * h:].sum() == 0 and a[:, :G].sum() == 0 and a[:, G + Ly + 2 * h:].sum() == 0
        np.add.at(rho, (xs[:, None, None], ys[None, :, None], zs[None, None, :]), a[G:G + Lx + 2 * h, G:G + Ly + 2 * h])
    torch.cuda.synchronize()
    tag = f"p{px}{py}"
    assert rel_err(rho, g[f"{tag}_paint"]) < FIELD_TOL
    join = lambda l: np.concatenate([np.concatenate([N(t) for t in l[rx * py:(rx + 1) * py]], axis=1) for rx in range(px)])
    assert rel_err(join(vl), g[f"{tag}_forces"]) < FIELD_TOL
    np.testing.assert_array_equal(join(dl), disp)       # zero drift: positions untouched
    for s_ in st:
        s_.close(barrier=False)


@pytest.mark.parametrize("fixture,pdims", [("distributed", (2, 2)), ("distributed", (1, 4)), ("distributed", (4, 1)),
                                           ("distributed", (2, 4)), ("distributed_slab", (2, 1)),
                                           ("distributed_slab", (4, 1)), ("distributed_pencil", (2, 2)),
                                           ("distributed_pencil", (2, 4))])
def test_golden_particle_to_rank_assignment(cuda, fixture, pdims):
    """uniform_particles(sharding=Sharding(pdims, rank=r)) of the PRODUCT == the block of the reference's sharded
    uniform_particles (jaxpm/distributed.py:168-190) that rank r owns: bit-exact (integers), every rank."""
    from jaxpm_b200.distributed import Sharding, get_local_shape, uniform_particles
    g = gold(fixture)
    ref = g[f"p{pdims[0]}{pdims[1]}_particles"]
    shape = ref.shape[:3]
    for r in range(pdims[0] * pdims[1]):
        sh = Sharding(pdims, rank=r)
        loc = get_local_shape(shape, sh)
        if fixture == "distributed":
            np.testing.assert_array_equal(loc, g[f"p{pdims[0]}{pdims[1]}_local_shape"])
        got = N(uniform_particles(shape, sharding=sh, device=cuda))
        assert got.shape == (*loc, 3)
        blk = ref[sh.rx * loc[0]:(sh.rx + 1) * loc[0], sh.ry * loc[1]:(sh.ry + 1) * loc[1]]
        np.testing.assert_array_equal(got.astype(np.int32), blk)
        np.testing.assert_array_equal(got, blk.astype(np.float32))
