"""Parity at the sizes the metric is quoted on.

* configs[1] of BASELINE.json (256^3 / 256^3, 2LPT at a = 0.1, 40 PM steps to a = 1): the CUDA path against
  tests/golden/config2_256.npz, an END-TO-END run of the CPU oracle (tests/golden/make_config2.py, ~2 h of CPU):
  LPT fields, one force evaluation, the state after steps 1, 2 and 40, cell-index checksums over all 1.7e7
  particles, coarse-grained density, P(k).
* 512^3: on the clustered a ~ 1 state the benchmark times, one more step by three routes - resident tile kernels
  with the three-transform chain, resident with the potential chain, order-preserving kernels + cuFFT (the functional
  API's slow path) - must agree to the field tolerance; the measured difference between the two force paths is
  checked against the error bound the AUTO mode switches on, at a = 0.1 (smooth) and a ~ 1 (clustered).
"""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FIELD_TOL = 1e-5


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


def _sample(t, ids):
    return t.reshape(-1, t.shape[-1])[torch.as_tensor(ids, device=t.device)].cpu().numpy()


def _cell_sums(disp, shape):
    """Same wrapped int64 checksums as tests/golden/make_config2.py:cell_index_sums, from the CUDA cell indices."""
    from jaxpm_b200 import ops
    idx = ops.cell_index(disp, shape, relative=True).to(torch.int64)
    ids = torch.arange(1, idx.numel() + 1, device=idx.device, dtype=torch.int64)
    return np.array([int(idx.sum()), int((idx * ids).sum())], dtype=np.int64)


def _block_sums(field, coarse=32):
    n = field.shape[0]
    b = n // coarse
    return field.double().reshape(coarse, b, coarse, b, coarse, b).sum(dim=(1, 3, 5)).cpu().numpy()


@pytest.mark.parametrize("force_mode", ["spectral", "auto"])
def test_config2_256_against_oracle_run(cuda, force_mode):
    path = os.path.join(HERE, "golden", "config2_256.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/config2_256.npz not generated (python tests/golden/make_config2.py)")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_config2 as MC
    from jaxpm_b200 import ops
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import kick_drift_coefficients
    from jaxpm_b200.painting import cic_paint_dx
    from jaxpm_b200.pm import lpt, pm_forces
    from jaxpm_b200.utils import power_spectrum
    g = np.load(path)
    n, nsteps = int(g["n"]), int(g["nsteps"])
    have_final = f"step{nsteps}_pk" in g.files
    ids = g["sample_ids"]
    shape, box = (n, n, n), (float(n),) * 3
    cosmo = Planck15()
    ic = MC.initial_conditions(n, box)                      # same seeded input as the oracle run (CPU, ~10 s)
    np.testing.assert_array_equal(ic.reshape(-1)[ids], g["ic_sample"])
    ict = T(ic, cuda)
    # ---- 2LPT (and 1LPT) ---------------------------------------------------------------------------------
    dx, p, f = lpt(cosmo, ict, a=0.1, order=2)
    for got, key in ((dx, "lpt_dx"), (p, "lpt_p"), (f, "lpt_f")):
        ref = g[key]
        assert np.abs(_sample(got, ids) - ref).max() / np.abs(ref).max() < FIELD_TOL, key
    assert abs(float(dx.abs().max()) / float(g["lpt_dx_max"]) - 1) < 1e-5
    dx1 = lpt(cosmo, ict, a=0.1, order=1)[0]
    assert np.abs(_sample(dx1, ids) - g["lpt1_dx"]).max() / np.abs(g["lpt1_dx"]).max() < FIELD_TOL
    del dx1, f, ict
    # ---- cell indices: bit-exact over ALL particles (checksums) and on the sample ----------------------------
    # (computed from the CUDA displacement; they equal the oracle's as long as no particle of the CUDA field sits
    # on the other side of a cell boundary, i.e. fields agree far better than their distance to a boundary)
    sums = _cell_sums(dx, shape)
    if not np.array_equal(sums, g["lpt_cell_sums"]):
        # a handful of particles within rounding distance of a cell face may differ: bound their number
        from oracle import painting as OP
        d_ref, d_got = g["lpt_pos"], _sample(dx, ids)
        pm = np.stack(np.unravel_index(ids, shape), -1).astype(np.int32)
        i_ref, _ = OP.enmesh_rel(pm, d_ref, shape)
        i_got, _ = OP.enmesh_rel(pm, d_got, shape)
        assert (np.any(i_ref[:, 0] != i_got[:, 0], axis=-1)).mean() < 1e-4
    # ---- one force evaluation at the LPT state (fast functional API) ---------------------------------------
    F = pm_forces(dx, mesh_shape=shape, paint_absolute_pos=False)
    # (a) fast functional API == order-preserving kernels + cuFFT on the SAME displacement: every particle, strict
    from jaxpm_b200 import pm as jpm_pm
    jpm_pm._FAST_API = False
    try:
        F_slow = pm_forces(dx, mesh_shape=shape, paint_absolute_pos=False)
    finally:
        jpm_pm._FAST_API = True
    assert float((F - F_slow).abs().max()) / float(g["force0_max"]) < FIELD_TOL
    del F_slow
    # (b) against the oracle's run.  Its displacement differs from the CUDA one by fp32 rounding, and the reference's
    # relative-mode rule (painting_utils.py:48-65) is DISCONTINUOUS in two kinds of places, where a particle can take
    # the other branch in ONE of the two runs and move up to one particle mass:
    #   * the dropped-corner window at the periodic edge (+ mode='drop'): a corner coordinate in (-ulp(N)/2, 0) wraps
    #     to N in fp32 and is dropped (see test_nbody_config1);
    #   * just below every power of two 2^k: pp and pp + 1 lie in different binades, pp + 1 rounds UP to 2^k + 1, the
    #     second corner lands one cell too far with weight ~ -ulp and the particle paints (almost) nothing.
    # Such particles are found here explicitly; every sampled particle must then agree to 1e-5 PLUS the field of a
    # unit point mass at each of them (1 / 4 pi r^2 - at 1e-5 of max|F| = 3.5 one flipped particle is felt out to
    # r ~ 48 cells), and strictly to 1e-5 when there is none.
    w = float(np.spacing(np.float32(n))) / 2
    grid_i = [torch.arange(m, device=cuda, dtype=torch.float32) for m in shape]
    border = torch.zeros(shape, dtype=torch.bool, device=cuda)
    for ax in range(3):
        view = [1, 1, 1]
        view[ax] = -1
        x = grid_i[ax].view(view) + dx[..., ax]
        for edge in (0.0, -w, -1.0, -1.0 - w):
            border |= (x - edge).abs() < 5e-6
        for kpow in range(0, int(np.log2(n)) + 1):
            e2 = float(2 ** kpow)
            border |= ((x - e2) > -(float(np.spacing(np.float32(e2))) + 3e-6)) & ((x - e2) < 3e-6)
    bpos = border.nonzero().to(torch.float32)                     # Lagrangian sites; |dx| << the radii that matter
    err = np.abs(_sample(F, ids) - g["force0"]).max(-1) / float(g["force0_max"])
    allow = np.full(err.shape, FIELD_TOL)
    if bpos.shape[0]:
        spos = torch.as_tensor(np.stack(np.unravel_index(ids, shape), -1), device=cuda, dtype=torch.float32)
        dd = (spos[:, None, :] - bpos[None, :, :]).abs()
        dd = torch.minimum(dd, float(n) - dd)
        r2 = (dd * dd).sum(-1).clamp_min(1.0)
        allow = allow + (2.0 / (4 * np.pi * r2) / float(g["force0_max"])).sum(-1).cpu().numpy()
    nbad = int((err > FIELD_TOL).sum())
    print(f"[config2 {force_mode}] force0: median {np.median(err):.2e}, max {err.max():.2e}, {nbad} of {err.size} sampled "
          f"particles above {FIELD_TOL}; {bpos.shape[0]} particle(s) within 5e-6 of a discontinuity of the reference rule: "
          f"{bpos.cpu().numpy().astype(int).tolist()[:4]}; worst err / allowance {float((err / allow).max()):.2f}")
    assert (err <= allow).all()
    assert bpos.shape[0] <= 64
    assert abs(float(F.abs().max()) / float(g["force0_max"]) - 1) < 1e-4
    del F
    rho = cic_paint_dx(dx)
    # density block sums (8^3 cells per block): to 1e-5 of the largest block everywhere, except in the blocks a particle
    # found above paints into (its Lagrangian block and the neighbours |dx| < 8 away), where up to ONE particle mass
    # per such particle may differ
    dblk = np.abs(_block_sums(rho) - g["lpt_rho_blocks"])
    bw = n // 32
    may = np.zeros(dblk.shape, dtype=bool)
    for site in bpos.cpu().numpy().astype(int):
        c = site // bw
        for o in np.ndindex(3, 3, 3):
            may[tuple((c + np.array(o) - 1) % 32)] = True
    tol_blk = FIELD_TOL * np.abs(g["lpt_rho_blocks"]).max()
    print(f"[config2 {force_mode}] density blocks: max diff {dblk.max():.3f} particle masses, {(dblk > tol_blk).sum()} blocks above "
          f"{tol_blk:.4f}, all next to a flagged particle: {bool(may[dblk > tol_blk].all())}")
    assert (dblk[~may] < tol_blk).all() and dblk.max() < 1.01 and (dblk > tol_blk).sum() <= 2 * max(1, bpos.shape[0])
    _, pk = power_spectrum(rho, box_shape=box)
    assert np.abs(pk.cpu().numpy() / g["lpt_pk"] - 1).max() < 1e-4
    del rho
    # ---- resident steps -----------------------------------------------------------------------------------
    d, k = kick_drift_coefficients(cosmo, 0.1, 1.0, nsteps, "symplectic")
    pos, vel = dx.clone(), p.clone()
    ops.axpby(1.0, pos, d[0], vel, out=pos)
    sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=1)
    sim.set_force_mode(force_mode)
    sim.load(pos, vel)
    last = nsteps if have_final else 2
    for s in range(last):
        sim.step(k[s], d[s + 1] if s + 1 < nsteps else 0.0)
        tag = f"step{s + 1}"
        if s + 1 in (1, 2, nsteps) and f"{tag}_pk" in g.files:
            sim.store(pos, vel)
            # the stored positions already carry the NEXT drift (fused kernel): take it back out
            if s + 1 < nsteps:
                pos_now = ops.axpby(1.0, pos, -d[s + 1], vel)
            else:
                pos_now = pos
            err_p = np.abs(_sample(pos_now, ids) - g[f"{tag}_pos"]).max(-1)
            err_v = np.abs(_sample(vel, ids) - g[f"{tag}_vel"]).max(-1) / np.abs(g[f"{tag}_vel"]).max()
            rho = cic_paint_dx(pos_now)
            _, pk = power_spectrum(rho, box_shape=box)
            pk_err = np.abs(pk.cpu().numpy() / g[f"{tag}_pk"] - 1).max()
            blocks = _block_sums(rho)
            blk_err = np.abs(blocks - g[f"{tag}_rho_blocks"]).max() / np.abs(g[f"{tag}_rho_blocks"]).max()
            print(f"[config2 {force_mode}] {tag}: median|dpos| {np.median(err_p):.2e} max {err_p.max():.2e} cells, "
                  f"max dvel {err_v.max():.2e}, P(k) {pk_err:.2e}, density blocks {blk_err:.2e}, "
                  f"rho max {float(rho.max()):.1f} vs {float(g[f'{tag}_rho_max']):.1f}")
            if s + 1 <= 2:
                # every sampled particle to rounding - except the neighbourhood of a particle that crosses the periodic
                # edge within fp32 noise (the reference's dropped-corner discontinuity, see test_nbody_config1): with
                # 1.7e7 particles about one such event per paint is expected, each moving one particle mass in ONE cell
                assert np.quantile(err_p, 0.999) < 1e-4 and np.quantile(err_v, 0.999) < 2e-5, tag
                dblk = np.abs(blocks - g[f"{tag}_rho_blocks"])
                assert dblk.max() < 1.01 and (dblk > FIELD_TOL * np.abs(blocks).max()).sum() <= 8, tag
            else:
                # 40 steps of a chaotic system in fp32: individual particles in collapsed regions decorrelate,
                # the statistics the north star names (the final power spectrum) must not
                assert np.median(err_p) < 1e-3, tag
            assert pk_err < 1e-4, tag
            del rho
    info = sim.force_info()
    print(f"[config2 {force_mode}] force path: {info}")
    if force_mode == "auto" and have_final:
        assert info["steps_potential"] > 0


def _state_512(cuda, n, nsteps_total, upto):
    """(disp, vel, d, k) of the benchmark workload after `upto` of `nsteps_total` resident steps (spectral)."""
    from jaxpm_b200 import ops
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    from jaxpm_b200.ode import kick_drift_coefficients
    from jaxpm_b200.pm import linear_field, lpt
    shape = (n, n, n)
    cosmo = Planck15()
    ic = linear_field(shape, (float(n),) * 3, lambda kk: linear_matter_power(cosmo, kk), seed=0, device=cuda)
    dx, p, _ = lpt(cosmo, ic, a=0.1, order=1)
    del ic
    disp, vel = dx.contiguous(), p.contiguous()
    d, k = kick_drift_coefficients(cosmo, 0.1, 1.0, nsteps_total, "symplectic")
    ops.axpby(1.0, disp, d[0], vel, out=disp)
    if upto > 0:
        sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=1)
        sim.load(disp, vel)
        for s in range(upto):
            sim.step(k[s], d[s + 1])
        sim.store(disp, vel)
        del sim
    torch.cuda.empty_cache()
    return disp, vel, d, k


@pytest.mark.parametrize("upto", [0, 39])
def test_512_resident_potential_and_order_preserving_agree(cuda, upto):
    from jaxpm_b200 import ops
    n = int(os.environ.get("JPM_SCALE_N", "512"))
    shape = (n, n, n)
    disp, vel, d, k = _state_512(cuda, n, 40, upto)
    kk, dd = float(k[upto]), float(d[min(upto + 1, 39)])
    out = {}
    # (3) order-preserving kernels + cuFFT (the slow path of the functional API)
    p3, v3 = disp.clone(), vel.clone()
    ops.pm_step_(ops.get_plan(shape, cuda), p3, v3, kk, dd, True)
    out["direct"] = (p3, v3)
    for mode in ("spectral", "potential"):
        sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=1)
        sim.set_force_mode(mode)
        sim.load(disp, vel)
        sim.step(kk, dd)
        p, v = torch.empty_like(disp), torch.empty_like(vel)
        sim.store(p, v)
        out[mode] = (p, v)
        if mode == "potential":
            info = sim.force_info()
        fb = sim.fallback_counts()
        del sim
        torch.cuda.empty_cache()
    dv_ref = out["direct"][1] - vel
    dp_ref = out["direct"][0] - disp
    res = {}
    for mode in ("spectral", "potential"):
        dv = out[mode][1] - vel
        dp = out[mode][0] - disp
        res[mode] = (float((dv - dv_ref).abs().max() / dv_ref.abs().max()),
                     float((dp - dp_ref).abs().max() / dp_ref.abs().max()))
    e_sp = float((out["potential"][1] - out["spectral"][1]).abs().max() / dv_ref.abs().max())
    # the bound AUTO evaluates: 2.7e-6 * rms(psi) / max|F| (computed by the potential step itself)
    bound = None
    sim = ops.Sim(shape, shape, True, cuda, tile=16, margin=1)
    sim.set_force_mode("auto")
    sim.load(disp, vel)
    sim.step(0.0, 0.0)
    torch.cuda.synchronize()
    sim.step(0.0, 0.0)
    bound = sim.force_info()["error_bound"]
    del sim
    print(f"[{n}^3 step {upto}] kick update vs order-preserving path: spectral {res['spectral'][0]:.2e}, "
          f"potential {res['potential'][0]:.2e}; potential vs spectral {e_sp:.2e}; AUTO bound {bound:.2e}; "
          f"fallbacks (paint, read, generic paint, generic read) {fb}")
    assert res["spectral"][0] < FIELD_TOL and res["spectral"][1] < 2e-5
    assert e_sp < max(2.0 * bound, 2e-6), "the error bound of the AUTO force mode must cover the measured difference"
    if bound < 4e-6:       # where AUTO would run the potential chain it must meet the tolerance
        assert res["potential"][0] < FIELD_TOL and res["potential"][1] < 2e-5
