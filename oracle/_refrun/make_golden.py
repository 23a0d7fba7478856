#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the reference's OWN source (/root/reference/jaxpm,
unmodified, imported in place) on the NumPy stand-in for jax in oracle/_refrun/shim.

    python oracle/_refrun/make_golden.py          # only works where /root/reference exists

Every fixture stores the inputs next to the reference's outputs, so the tests need neither the
reference nor this script at run time.  See oracle/_refrun/__init__.py for what this does and does
not pin.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("JAXPM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "shim"), ROOT, REF]

import numpy as np  # noqa: E402

import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402
from jax import lax  # noqa: E402
from jax.sharding import NamedSharding, PartitionSpec as P, clear_mesh, make_mesh  # noqa: E402

import jaxpm  # noqa: E402
assert os.path.realpath(jaxpm.__path__[0] if hasattr(jaxpm, "__path__") else jaxpm.__file__).startswith(
    os.path.realpath(REF)), "jaxpm must come from the reference tree"
from jaxpm import distributed, growth, kernels, ode, painting, painting_utils, pm, utils  # noqa: E402

from oracle import cosmology as OC  # noqa: E402  ([ext] jax_cosmo restatement behind the shim)
from oracle.pm import linear_field as _colour  # noqa: E402  (input generation only)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
A = np.asarray
SHAPE = (8, 12, 16)


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: A(v) for k, v in arrs.items()})
    print(f"{name}: " + ", ".join(f"{k}{A(v).shape}" for k, v in arrs.items()))


def grid(shape):
    return np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).astype(np.float32)


def gen_paint_read_abs():
    rng = np.random.default_rng(11)
    pos = (grid(SHAPE) + 2.5 * rng.standard_normal((*SHAPE, 3))).astype(np.float32)
    pos[0, 0, 0] = (-0.25, -1e-7, SHAPE[2] + 2.5)        # negative / tiny-negative / beyond the box
    pos[1, 2, 3] = (7.9999995, 11.5, -17.25)
    pos[2, 2, 2] = (3.0, 4.0, 5.0)                       # exactly on a grid point
    w = rng.uniform(0.5, 1.5, SHAPE).astype(np.float32)
    base = rng.standard_normal(SHAPE).astype(np.float32)
    lax.SCATTER_LOG.clear()
    m1 = painting.cic_paint(jnp.zeros(SHAPE), jnp.asarray(pos))
    idx, ker = lax.SCATTER_LOG[-1]
    mw = painting.cic_paint(jnp.asarray(base), jnp.asarray(pos), jnp.asarray(w))
    m25 = painting.cic_paint(jnp.asarray(base), jnp.asarray(pos), 2.5)
    rd = painting.cic_read(jnp.asarray(base), jnp.asarray(pos))
    save("paint_read_abs", pos=pos, weight=w, base=base, idx=idx.astype(np.int32), kernel=ker,
         mesh_w1=m1, mesh_warr_on_base=mw, mesh_w2p5_on_base=m25, read_base=rd)


def gen_paint_read_rel():
    rng = np.random.default_rng(12)
    disp = (1.5 * rng.standard_normal((*SHAPE, 3))).astype(np.float32)
    disp[0, 0, 0] = (-1e-7, 0.3, -0.2)                   # float-mod edge: index N, dropped
    disp[1, 1, 1] = (-1.0 - 1e-7, 0.0, 0.0)
    disp[2, 2, 2] = (0.0, 0.0, 0.0)
    disp[7, 11, 15] = (0.99999994, 0.5, 0.99999994)
    w = rng.uniform(0.5, 1.5, SHAPE).astype(np.float32)
    out = dict(disp=disp, weight=w)
    for tag, (hx, hy) in (("h00", (0, 0)), ("h23", (2, 3))):
        hs = ((hx, hx), (hy, hy), (0, 0))
        pshape = (SHAPE[0] + 2 * hx, SHAPE[1] + 2 * hy, SHAPE[2])
        a, b, c = np.meshgrid(*[np.arange(s) for s in SHAPE], indexing="ij")
        pmid = jnp.asarray(np.stack([a + hx, b + hy, c], -1).reshape(-1, 3).astype(np.int32))
        one = jnp.asarray(np.float32(1.0))
        zero = jnp.asarray(np.int32(0))
        ind, frac = painting_utils.enmesh(pmid, jnp.asarray(disp.reshape(-1, 3)), one, pshape, zero, one, pshape)
        m = painting._cic_paint_dx_impl(jnp.asarray(disp), halo_size=hs)
        mw = painting._cic_paint_dx_impl(jnp.asarray(disp), jnp.asarray(w), halo_size=hs)
        field = rng.standard_normal(pshape).astype(np.float32)
        rd = painting._cic_read_dx_impl(jnp.asarray(field), jnp.asarray(disp), hs)
        out.update({f"idx_{tag}": A(ind).astype(np.int32), f"w_{tag}": frac, f"mesh_{tag}": m,
                    f"mesh_warr_{tag}": mw, f"field_{tag}": field, f"read_{tag}": rd})
    out["paint_dx_api"] = painting.cic_paint_dx(jnp.asarray(disp))
    out["read_dx_api"] = painting.cic_read_dx(jnp.asarray(out["field_h00"]), jnp.asarray(disp))
    save("paint_read_rel", **out)


def gen_kernels():
    dk = jnp.zeros(SHAPE, dtype=np.complex64)
    kvec = kernels.fftk(dk)
    out = {f"k{d}": kvec[d] for d in range(3)}
    for d in range(3):
        out[f"grad{d}_o1"] = kernels.gradient_kernel(kvec, d)
        out[f"grad{d}_o0"] = kernels.gradient_kernel(kvec, d, order=0)
    out["invlap"] = kernels.invlaplace_kernel(kvec)
    out["invlap_fd"] = kernels.invlaplace_kernel(kvec, fd=True)
    out["longrange_r0"] = np.float32(kernels.longrange_kernel(kvec, 0))
    out["longrange_r1p5"] = kernels.longrange_kernel(kvec, 1.5)
    out["cic_comp"] = kernels.cic_compensation(kvec)
    save("kernels", **out)


def gen_pm_forces():
    rng = np.random.default_rng(13)
    pos = (grid(SHAPE) + 1.2 * rng.standard_normal((*SHAPE, 3))).astype(np.float32)
    disp = (1.2 * rng.standard_normal((*SHAPE, 3))).astype(np.float32)
    delta = (1.0 + 0.5 * rng.standard_normal(SHAPE)).astype(np.float32)
    out = dict(pos=pos, disp=disp, delta=delta)
    out["f_abs"] = pm.pm_forces(jnp.asarray(pos), mesh_shape=SHAPE)
    out["f_rel"] = pm.pm_forces(jnp.asarray(disp), mesh_shape=SHAPE, paint_absolute_pos=False)
    out["f_abs_rsplit2"] = pm.pm_forces(jnp.asarray(pos), mesh_shape=SHAPE, r_split=2.0)
    out["f_abs_delta_real"] = pm.pm_forces(jnp.asarray(pos), delta=jnp.asarray(delta))
    out["f_rel_delta_cplx"] = pm.pm_forces(jnp.asarray(disp), delta=distributed.fft3d(jnp.asarray(delta)),
                                           paint_absolute_pos=False)
    save("pm_forces", **out)


def _ic(shape, box, seed):
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    c = Planck15()
    wn = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    return _colour(wn, box, lambda k: linear_matter_power(c, k))


def gen_lpt():
    shape, box = (16, 16, 24), (64.0, 64.0, 96.0)
    ic = _ic(shape, box, 14)
    cosmo = OC.Planck15()
    out = dict(ic=ic, a=np.float64(0.1), box=np.asarray(box))
    part = grid(shape)
    for order in (1, 2):
        dx, p, f = pm.lpt(cosmo, jnp.asarray(ic), a=0.1, order=order)
        out.update({f"rel_o{order}_dx": dx, f"rel_o{order}_p": p, f"rel_o{order}_f": f})
        dx, p, f = pm.lpt(cosmo, jnp.asarray(ic), particles=jnp.asarray(part), a=0.1, order=order)
        out.update({f"abs_o{order}_dx": dx, f"abs_o{order}_p": p, f"abs_o{order}_f": f})
    save("lpt", **out)


def gen_growth_ode():
    cosmo = OC.Planck15()
    a = np.array([0.1, 0.3, 0.7, 1.0])
    out = dict(a=a)
    for name in ("E", "dEa", "gp", "Gf", "Gf2", "dGfa", "dGf2a", "growth_factor", "growth_rate",
                 "growth_factor_second", "growth_rate_second"):
        out["g_" + name] = np.asarray(getattr(growth, name)(cosmo, jnp.asarray(a)), dtype=np.float64)
    rng = np.random.default_rng(15)
    pos = (grid(SHAPE) + 0.8 * rng.standard_normal((*SHAPE, 3))).astype(np.float32)
    vel = (0.05 * rng.standard_normal((*SHAPE, 3))).astype(np.float32)
    out.update(pos=pos, vel=vel)
    jp, jv = jnp.asarray(pos), jnp.asarray(vel)
    dpos, dvel = ode.make_ode_fn(SHAPE)((jp, jv), 0.3, cosmo)
    out.update(ode_dpos=dpos, ode_dvel=dvel)
    out["diffrax_rhs"] = ode.make_diffrax_ode(SHAPE)(0.3, jnp.stack([jp, jv]), cosmo)
    drift, kick = ode.symplectic_ode(SHAPE, cosmo)
    out.update(sym_drift=drift(0.3, jv, None), sym_kick=kick(0.3, jp, None))
    drift, kick, first = ode.symplectic_fpm_ode(SHAPE, 0.05, cosmo)
    out.update(fpm_drift=drift(0.3, jv, None), fpm_kick=kick(0.3, jp, None), fpm_first_kick=first(0.3, jp, cosmo),
               fpm_dt0=np.float64(0.05), ode_a=np.float64(0.3))
    save("growth_ode", **out)


def gen_distributed():
    shape, halo = (16, 16, 8), 4
    rng = np.random.default_rng(16)
    # |disp| < halo//2 = 2 cells: the reach within which the reference's halo protocol is exact
    disp = np.clip(0.7 * rng.standard_normal((*shape, 3)), -1.9, 1.9).astype(np.float32)
    field = rng.standard_normal(shape).astype(np.float32)
    out = dict(disp=disp, field=field, halo=np.int32(halo))
    clear_mesh()
    out["single_paint"] = painting.cic_paint_dx(jnp.asarray(disp))
    out["single_read"] = painting.cic_read_dx(jnp.asarray(field), jnp.asarray(disp))
    out["single_forces"] = pm.pm_forces(jnp.asarray(disp), mesh_shape=shape, paint_absolute_pos=False)
    out["single_particles"] = distributed.uniform_particles(shape)
    for pd in ((2, 2), (1, 4), (4, 1), (2, 4)):
        tag = f"p{pd[0]}{pd[1]}"
        sh = NamedSharding(make_mesh(pd), P('x', 'y'))
        hs, ext = distributed.get_halo_size((halo, halo), sh)
        out[f"{tag}_halo_size"] = np.asarray(hs, dtype=np.int32)
        out[f"{tag}_halo_ext"] = np.asarray(ext, dtype=np.int32)
        out[f"{tag}_paint"] = painting.cic_paint_dx(jnp.asarray(disp), halo_size=(halo, halo), sharding=sh)
        out[f"{tag}_read"] = painting.cic_read_dx(jnp.asarray(field), jnp.asarray(disp), halo_size=(halo, halo),
                                                  sharding=sh)
        out[f"{tag}_forces"] = pm.pm_forces(jnp.asarray(disp), mesh_shape=shape, paint_absolute_pos=False,
                                            halo_size=(halo, halo), sharding=sh)
        out[f"{tag}_particles"] = distributed.uniform_particles(shape, sharding=sh)
        out[f"{tag}_local_shape"] = np.asarray(distributed.get_local_shape(shape, sh), dtype=np.int32)
        clear_mesh()
    # slice_unpad_impl on one padded block (the halo add rule, distributed.py:68-85)
    blk = rng.standard_normal((8 + 8, 4 + 8, 8)).astype(np.float32)
    out["unpad_in"] = blk
    out["unpad_out_h44"] = distributed.slice_unpad_impl(jnp.asarray(blk), ((4, 4), (4, 4), (0, 0)))
    blk2 = rng.standard_normal((8 + 8, 4, 8)).astype(np.float32)
    out["unpad_in_h40"] = blk2
    out["unpad_out_h40"] = distributed.slice_unpad_impl(jnp.asarray(blk2), ((4, 4), (0, 0), (0, 0)))
    save("distributed", **out)


def gen_distributed_slab():
    """x-slab decompositions (pdims (P, 1)) of a power-of-two mesh, the shapes the fused peer-memory slab path
    (jaxpm_b200/slab.py) serves: the reference's own paint / read / pm_forces / uniform_particles, sharded."""
    shape, halo = (32, 16, 16), 8
    rng = np.random.default_rng(23)
    disp = np.clip(1.4 * rng.standard_normal((*shape, 3)), -3.9, 3.9).astype(np.float32)
    field = rng.standard_normal(shape).astype(np.float32)
    out = dict(disp=disp, field=field, halo=np.int32(halo))
    clear_mesh()
    out["single_paint"] = painting.cic_paint_dx(jnp.asarray(disp))
    out["single_forces"] = pm.pm_forces(jnp.asarray(disp), mesh_shape=shape, paint_absolute_pos=False)
    for pd in ((2, 1), (4, 1)):
        tag = f"p{pd[0]}{pd[1]}"
        sh = NamedSharding(make_mesh(pd), P('x', 'y'))
        out[f"{tag}_paint"] = painting.cic_paint_dx(jnp.asarray(disp), halo_size=(halo, halo), sharding=sh)
        out[f"{tag}_read"] = painting.cic_read_dx(jnp.asarray(field), jnp.asarray(disp), halo_size=(halo, halo),
                                                  sharding=sh)
        out[f"{tag}_forces"] = pm.pm_forces(jnp.asarray(disp), mesh_shape=shape, paint_absolute_pos=False,
                                            halo_size=(halo, halo), sharding=sh)
        out[f"{tag}_particles"] = distributed.uniform_particles(shape, sharding=sh)
        clear_mesh()
    save("distributed_slab", **out)


def gen_distributed_pencil():
    """Pencil decompositions (pdims (px, py), tests/test_distributed_pm.py:28) of a power-of-two mesh in the shapes the
    fused peer-memory path serves (ny / py a multiple of 16): the reference's own sharded paint / pm_forces /
    uniform_particles (two grids, to keep the fixture small)."""
    shape, halo = (32, 64, 16), 8
    rng = np.random.default_rng(29)
    lim = 3.5
    disp = (lim * np.tanh(1.4 * rng.standard_normal((*shape, 3)) / lim)).astype(np.float32)
    out = dict(disp=disp, halo=np.int32(halo))
    clear_mesh()
    out["single_paint"] = painting.cic_paint_dx(jnp.asarray(disp))
    for pd in ((2, 2), (2, 4)):
        tag = f"p{pd[0]}{pd[1]}"
        sh = NamedSharding(make_mesh(pd), P('x', 'y'))
        out[f"{tag}_paint"] = painting.cic_paint_dx(jnp.asarray(disp), halo_size=(halo, halo), sharding=sh)
        out[f"{tag}_forces"] = pm.pm_forces(jnp.asarray(disp), mesh_shape=shape, paint_absolute_pos=False,
                                            halo_size=(halo, halo), sharding=sh)
        out[f"{tag}_particles"] = np.asarray(distributed.uniform_particles(shape, sharding=sh)).astype(np.int16)
        clear_mesh()
    save("distributed_pencil", **out)


def gen_power_spectrum():
    rng = np.random.default_rng(17)
    shape, box = (16, 16, 24), (100.0, 100.0, 150.0)
    f1 = rng.standard_normal(shape).astype(np.float32)
    f2 = (0.7 * f1 + 0.3 * rng.standard_normal(shape)).astype(np.float32)
    k, pk = utils.power_spectrum(jnp.asarray(f1), box_shape=box)
    k2, pkx = utils.power_spectrum(jnp.asarray(f1), jnp.asarray(f2), box_shape=box)
    k3, pkl = utils.power_spectrum(jnp.asarray(f1), box_shape=box, multipoles=[0, 2], kedges=5)
    kc, pkc = utils.power_spectrum(jnp.asarray(f1))
    save("power_spectrum", f1=f1, f2=f2, box=np.asarray(box), k=k, pk=pk, pk_cross=pkx, k_poles=k3,
         pk_poles=pkl, k_cell=kc, pk_cell=pkc)


def gen_widened():
    """compensate_cic (painting.py:263-275) and cic_paint_2d (:131-158)."""
    rng = np.random.default_rng(19)
    shape = (8, 12, 16)
    f = rng.standard_normal(shape).astype(np.float32)
    comp = painting.compensate_cic(jnp.asarray(f))
    pshape, n = (12, 10), 500
    pos = rng.uniform(-3, 15, (n, 2)).astype(np.float32)
    pos[0] = (-1e-7, 9.9999995)
    w = rng.uniform(0.5, 1.5, n).astype(np.float32)
    base = rng.standard_normal(pshape).astype(np.float32)
    m_w = painting.cic_paint_2d(jnp.asarray(base), jnp.asarray(pos), jnp.asarray(w))
    m_1 = painting.cic_paint_2d(jnp.zeros(pshape), jnp.asarray(pos), None)
    from jaxpm import lensing
    pos3 = rng.uniform(-4, 20, (3000, 3)).astype(np.float32)
    pos3[0] = (-1e-7, 16.0, 10.0)
    dplane = lensing.density_plane(jnp.asarray(pos3), (16, 16, 16), 8.0, 4.0, 12)
    save("widened", field=f, compensated=comp, pos2=pos, w2=w, base2=base, mesh2_weighted=m_w, mesh2_unit=m_1,
         pos3=pos3, density_plane=dplane)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (gen_paint_read_abs, gen_paint_read_rel, gen_kernels, gen_pm_forces, gen_lpt, gen_growth_ode,
               gen_distributed, gen_distributed_slab, gen_distributed_pencil, gen_power_spectrum, gen_widened):
        if only and fn.__name__ not in only:
            continue
        fn()
    print("golden fixtures written to", OUT)
