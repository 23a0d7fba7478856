"""The multi-GPU paths exercised on ONE device, so that the round-end single-GPU box covers them:

* the slab decomposition over peer memory (jaxpm_b200/slab.py, csrc/pmfft.cu) with P ranks living in
  this process, each on its own stream, their blocks attached by pointer instead of by CUDA IPC handle;
* the NCCL halo protocol of jaxpm_b200/halo.py with a (1, 1) process grid, where every rank is its own
  neighbour (the send/recv degenerates to a copy, the pack/unpack/accumulate logic is the same).

Reference behaviour: sharded == unsharded (/root/reference/tests/test_distributed_pm.py:37-179).
The real multi-process runs are tests/test_multi_gpu.py (needs >= 2 GPUs)."""
import numpy as np
import pytest
import torch

from helpers import displaced, rel_err

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-5


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


def _make_plans(shape, P, gx, dev):
    from jaxpm_b200.slab import SlabPlan
    plans = [SlabPlan(shape, P, r, gx, dev) for r in range(P)]
    for p in plans:
        p.attach_local(plans)
    return plans


@pytest.mark.parametrize("shape,P,gx", [((32, 32, 32), 1, 4), ((32, 32, 32), 2, 8), ((32, 32, 32), 2, 16),
                                        ((64, 32, 16), 4, 5), ((64, 64, 64), 8, 8), ((128, 32, 64), 2, 7)])
def test_slab_forces_equal_single_gpu(cuda, shape, P, gx):
    """density block per rank -> fused peer-memory FFT chain -> force blocks == the single-GPU chain."""
    from jaxpm_b200 import ops
    rng = np.random.default_rng(5)
    rho = rng.standard_normal(shape).astype(np.float32)
    ref = ops.force_meshes_from_density(T(rho, cuda), ops.get_plan(shape, cuda)).cpu().numpy()
    plans = _make_plans(shape, P, gx, cuda)
    streams = [torch.cuda.Stream(cuda) for _ in range(P)]
    lx = shape[0] // P
    torch.cuda.synchronize()
    for rep in range(2):      # twice: the barrier epochs and buffer reuse of a second evaluation
        for r, (p, s) in enumerate(zip(plans, streams)):
            with torch.cuda.stream(s):
                p.set_density(T(rho[r * lx:(r + 1) * lx], cuda))
        for p, s in zip(plans, streams):
            with torch.cuda.stream(s):
                p.forces()
        for r, (p, s) in enumerate(zip(plans, streams)):
            with torch.cuda.stream(s):
                p.check()
                for d in range(3):
                    got = p.interior(1 + d).cpu().numpy()
                    err = np.abs(got - ref[d][r * lx:(r + 1) * lx]).max() / np.abs(ref[d]).max()
                    assert err < FIELD_TOL, (rep, r, d, err)
    torch.cuda.synchronize()
    for p in plans:
        p.destroy()


@pytest.mark.parametrize("force_mode", ["spectral", "potential", "auto"])
@pytest.mark.parametrize("shape,P,gx,tile", [((32, 32, 32), 2, 8, 8), ((64, 64, 64), 2, 16, 16), ((64, 32, 32), 4, 8, 8),
                                             ((32, 32, 32), 1, 8, 8)])
def test_slab_stepper_equals_single_gpu(cuda, shape, P, gx, tile, force_mode):
    """K drift-kick steps of the slab stepper (ghost-fold / transposes / ghost-fill inside the FFT kernels)
    == the single-GPU resident stepper on the same particles, for the three force paths (potential: psi ghost
    planes + gradient pass per rank; auto: the ranks share their force statistics and switch together)."""
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import kick_drift_coefficients, nbody_kick_drift
    from jaxpm_b200.slab import SlabStepper
    from jaxpm_b200 import ops
    _, disp = displaced(shape, 1.0)
    disp = np.clip(disp, -gx / 2 + 0.5, gx / 2 - 0.5).astype(np.float32)
    vel = (0.2 * np.random.default_rng(9).standard_normal(disp.shape)).astype(np.float32)
    cosmo = Planck15()
    K = 7 if force_mode == "auto" else 3      # long enough for the lagged AUTO decision to switch the ranks over
    rp, rv = nbody_kick_drift(cosmo, T(disp, cuda), T(vel, cuda), 0.5, 0.8, K, paint_absolute_pos=False,
                              resident=True, tile=tile, margin=1)
    d, k = kick_drift_coefficients(cosmo, 0.5, 0.8, K, "symplectic")
    lx = shape[0] // P
    plans = _make_plans(shape, P, gx, cuda)
    streams = [torch.cuda.Stream(cuda) for _ in range(P)]
    dl = [T(disp[r * lx:(r + 1) * lx], cuda) for r in range(P)]
    vl = [T(vel[r * lx:(r + 1) * lx], cuda) for r in range(P)]
    for r in range(P):
        ops.axpby(1.0, dl[r], d[0], vl[r], out=dl[r])
    torch.cuda.synchronize()
    steppers = []
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            steppers.append(SlabStepper(dl[r], vl[r], gx, P, r, tile=tile, margin=1, plan=plans[r],
                                        force_mode=force_mode))
    for n in range(K):
        for r in range(P):
            with torch.cuda.stream(streams[r]):
                # several ranks driven by ONE host thread: the lagged AUTO decision waits for an earlier step of
                # this rank, which needs the peers' kernels of that step to have been enqueued - they have, the
                # ranks are stepped in lock step and the lag is >= 2 steps
                steppers[r].step(k[n], d[n + 1] if n + 1 < K else 0.0)
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            steppers[r].store(dl[r], vl[r])
    torch.cuda.synchronize()
    p = torch.cat(dl).cpu().numpy()
    v = torch.cat(vl).cpu().numpy()
    assert np.abs(p - rp.cpu().numpy()).max() < 2e-4
    assert rel_err(v, rv.cpu().numpy()) < 1e-4
    infos = [s.force_info() for s in steppers]
    if force_mode == "potential":
        assert all(i["steps_potential"] == K for i in infos)
    if force_mode == "auto":
        # every rank took the same decisions (white-noise displacements: the bound is small, they switch over)
        assert len({(i["steps_spectral"], i["steps_potential"]) for i in infos}) == 1, infos
        assert infos[0]["steps_potential"] >= 1, infos
    for s in steppers:
        s.close(barrier=False)


def _make_pencil_plans(shape, pdims, gx, gy, dev):
    from jaxpm_b200.slab import SlabPlan
    P = pdims[0] * pdims[1]
    plans = [SlabPlan(shape, P, r, gx, dev, pdims=pdims, gy=gy) for r in range(P)]
    for p in plans:
        p.attach_local(plans)
    return plans


def _blocks(a, pdims):
    """[(rank, block)] of the leading two axes split over the (px, py) grid, rank = rx * py + ry."""
    px, py = pdims
    Lx, Ly = a.shape[0] // px, a.shape[1] // py
    return [a[rx * Lx:(rx + 1) * Lx, ry * Ly:(ry + 1) * Ly] for rx in range(px) for ry in range(py)]


PENCIL_CASES = [((32, 32, 32), (2, 2), 8, 8), ((64, 64, 32), (2, 4), 8, 5), ((64, 64, 64), (4, 2), 16, 16),
                ((32, 64, 32), (1, 4), 4, 8), ((64, 32, 16), (1, 2), 8, 16), ((32, 128, 32), (2, 2), 3, 7)]


@pytest.mark.parametrize("shape,pdims,gx,gy", PENCIL_CASES)
def test_pencil_forces_equal_single_gpu(cuda, shape, pdims, gx, gy):
    """Pencil process grids (jaxpm/distributed.py:116-129; tests/test_distributed_pm.py:28 pdims (4,2), (2,4), (1,8)):
    density block per rank -> the z passes transpose within the row group while they load / store, slab FFT chain in
    between -> force blocks == the single-GPU chain."""
    from jaxpm_b200 import ops
    rng = np.random.default_rng(5)
    rho = rng.standard_normal(shape).astype(np.float32)
    ref = ops.force_meshes_from_density(T(rho, cuda), ops.get_plan(shape, cuda)).cpu().numpy()
    P = pdims[0] * pdims[1]
    plans = _make_pencil_plans(shape, pdims, gx, gy, cuda)
    streams = [torch.cuda.Stream(cuda) for _ in range(P)]
    rb = _blocks(rho, pdims)
    torch.cuda.synchronize()
    for rep in range(2):
        for r, (p, s) in enumerate(zip(plans, streams)):
            with torch.cuda.stream(s):
                p.set_density(T(rb[r], cuda))
        for p, s in zip(plans, streams):
            with torch.cuda.stream(s):
                p.forces()
        for r, (p, s) in enumerate(zip(plans, streams)):
            with torch.cuda.stream(s):
                p.check()
                for d in range(3):
                    got = p.interior(1 + d).cpu().numpy()
                    err = np.abs(got - _blocks(ref[d], pdims)[r]).max() / np.abs(ref[d]).max()
                    assert err < FIELD_TOL, (rep, r, d, err)
    torch.cuda.synchronize()
    for p in plans:
        p.destroy()


@pytest.mark.parametrize("force_mode,K", [("spectral", 1), ("potential", 1), ("spectral", 3), ("auto", 7)])
@pytest.mark.parametrize("shape,pdims,gx,gy,tile", [((32, 32, 32), (2, 2), 8, 8, 8), ((64, 64, 64), (2, 2), 16, 16, 16),
                                                    ((64, 64, 32), (2, 4), 8, 8, 8), ((64, 64, 32), (4, 2), 8, 8, 8),
                                                    ((32, 64, 32), (1, 4), 8, 8, 8)])
def test_pencil_stepper_equals_single_gpu(cuda, shape, pdims, gx, gy, tile, force_mode, K):
    """K drift-kick steps of the fused stepper on a pencil grid (ghost planes / rows / corners folded and filled by the
    z passes across the 3 x 3 neighbourhood, ghost width from the particles' actual reach) == the single-GPU resident
    stepper on the same particles.

    One step is held to rounding for EVERY particle.  Over several steps the comparison has to live with the
    reference's own rule (painting_utils.py:48-65), which is discontinuous in fp32 just below every power-of-two
    coordinate (pp + 1 rounds up into the next binade: the particle paints ~nothing for one step) and at the periodic
    edge (dropped corner): those windows sit at different particles in global and in per-shard coordinates, so about
    one particle per few 1e5 particle-steps paints in one run and not in the other (the reference's own sharded and
    unsharded runs differ the same way).  Several steps therefore: the bulk to rounding, and at most a few
    neighbourhoods (a 5-cell ball per event) off."""
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import kick_drift_coefficients, nbody_kick_drift
    from jaxpm_b200.slab import SlabStepper
    from jaxpm_b200 import ops
    _, disp = displaced(shape, 1.0)
    g = min(gx, gy)
    lim = g / 2 - 0.5
    disp = (lim * np.tanh(disp / lim)).astype(np.float32)        # bounded by the halo reach, no pile-up at the bound
    vel = (0.2 * np.random.default_rng(9).standard_normal(disp.shape)).astype(np.float32)
    cosmo = Planck15()
    rp, rv = nbody_kick_drift(cosmo, T(disp, cuda), T(vel, cuda), 0.5, 0.8, K, paint_absolute_pos=False,
                              resident=True, tile=tile, margin=1)
    d, k = kick_drift_coefficients(cosmo, 0.5, 0.8, K, "symplectic")
    P = pdims[0] * pdims[1]
    plans = _make_pencil_plans(shape, pdims, gx, gy, cuda)
    streams = [torch.cuda.Stream(cuda) for _ in range(P)]
    dl = [T(b, cuda) for b in _blocks(disp, pdims)]
    vl = [T(b, cuda) for b in _blocks(vel, pdims)]
    for r in range(P):
        ops.axpby(1.0, dl[r], d[0], vl[r], out=dl[r])
    torch.cuda.synchronize()
    steppers = []
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            steppers.append(SlabStepper(dl[r], vl[r], gx, P, r, tile=tile, margin=1, plan=plans[r],
                                        force_mode=force_mode, pdims=pdims, gy=gy))
    for n in range(K):
        for r in range(P):
            with torch.cuda.stream(streams[r]):
                steppers[r].step(k[n], d[n + 1] if n + 1 < K else 0.0)
    for r in range(P):
        with torch.cuda.stream(streams[r]):
            steppers[r].store(dl[r], vl[r])
    torch.cuda.synchronize()
    px, py = pdims
    p = torch.cat([torch.cat(dl[rx * py:(rx + 1) * py], dim=1) for rx in range(px)], dim=0).cpu().numpy()
    v = torch.cat([torch.cat(vl[rx * py:(rx + 1) * py], dim=1) for rx in range(px)], dim=0).cpu().numpy()
    infos = [s.force_info() for s in steppers]
    ep = np.abs(p - rp.cpu().numpy()).max(-1)
    ev = np.abs(v - rv.cpu().numpy()).max(-1) / np.abs(rv.cpu().numpy()).max()
    bad = float(((ep > 2e-4) | (ev > 1e-4)).mean())
    print(f"[pencil {pdims} {force_mode} K={K}] max|dpos| {ep.max():.2e}, rel dvel {ev.max():.2e}, median {np.median(ep):.1e}, "
          f"bad fraction {bad:.4f}, ghost width in use {[pl.ghost_width() for pl in plans]}, halo exceeded "
          f"{[pl.halo_exceeded() for pl in plans]}, {infos[0]}")
    assert not any(pl.halo_exceeded() for pl in plans)
    assert np.median(ep) < 5e-6 and np.median(ev) < 5e-6
    if K == 1:
        assert ep.max() < 2e-4 and ev.max() < 1e-4
    else:
        assert bad < 0.02 * K / 3, "more than a few event neighbourhoods differ"
    if force_mode == "potential":
        assert all(i["steps_potential"] == K for i in infos)
    if force_mode == "auto":
        assert len({(i["steps_spectral"], i["steps_potential"]) for i in infos}) == 1, infos
    for s in steppers:
        s.close(barrier=False)


@pytest.mark.parametrize("resident", [False, True])
def test_halo_protocol_self_neighbour(cuda, resident):
    """halo.ShardedStepper on a (1, 1) process grid with halo 8: pad / paint / halo reduce / halo fill /
    read run exactly as on a real process grid, the neighbour being the rank itself."""
    from jaxpm_b200 import halo, ops
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.distributed import Sharding
    from jaxpm_b200.ode import kick_drift_coefficients, nbody_kick_drift
    shape, h, K = (32, 32, 24), 8, 3
    _, disp = displaced(shape, 1.0)
    disp = np.clip(disp, -h / 2 + 0.5, h / 2 - 0.5).astype(np.float32)
    vel = (0.2 * np.random.default_rng(9).standard_normal(disp.shape)).astype(np.float32)
    cosmo = Planck15()
    rp, rv = nbody_kick_drift(cosmo, T(disp, cuda), T(vel, cuda), 0.5, 0.8, K, paint_absolute_pos=False, resident=False)
    d, k = kick_drift_coefficients(cosmo, 0.5, 0.8, K, "symplectic")
    sh = Sharding((1, 1), rank=0)
    p, v = T(disp, cuda), T(vel, cuda)
    ops.axpby(1.0, p, d[0], v, out=p)
    st = halo.ShardedStepper(p, v, h, sh, resident=resident, halos=(h, h))
    for n in range(K):
        st.step(k[n], d[n + 1] if n + 1 < K else 0.0)
    st.store(p, v)
    assert np.abs(p.cpu().numpy() - rp.cpu().numpy()).max() < 2e-4
    assert rel_err(v.cpu().numpy(), rv.cpu().numpy()) < 1e-4


def test_slab_reports_halo_too_small(cuda):
    """A displacement field that reaches the outermost ghost plane is flagged (the reference's halo_size has the same
    validity limit but fails silently); a field inside the reach is not."""
    import warnings
    from jaxpm_b200.slab import SlabStepper
    shape, P, gx = (32, 32, 32), 2, 4
    lx = shape[0] // P
    for amp, expect in ((1.0, False), (6.0, True)):
        plans = _make_plans(shape, P, gx, cuda)
        streams = [torch.cuda.Stream(cuda) for _ in range(P)]
        disp = np.zeros((*shape, 3), np.float32)
        disp[..., 0] = amp * np.sin(2 * np.pi * np.arange(shape[1]) / shape[1])[None, :, None]
        vel = np.zeros_like(disp)
        dl = [T(disp[r * lx:(r + 1) * lx], cuda) for r in range(P)]
        vl = [T(vel[r * lx:(r + 1) * lx], cuda) for r in range(P)]
        torch.cuda.synchronize()
        st = []
        for r in range(P):
            with torch.cuda.stream(streams[r]):
                st.append(SlabStepper(dl[r], vl[r], gx, P, r, tile=8, margin=1, plan=plans[r]))
        for r in range(P):
            with torch.cuda.stream(streams[r]):
                st[r].step(0.0, 0.0)
        flagged = False
        for r in range(P):
            with torch.cuda.stream(streams[r]), warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                st[r].store(dl[r], vl[r])
                flagged = flagged or any("halo_size is too small" in str(x.message) for x in w)
        torch.cuda.synchronize()
        assert flagged == expect, (amp, flagged)
        for s_ in st:
            s_.close(barrier=False)
