#!/usr/bin/env python
"""Generate tests/golden/config2_256.npz: BASELINE.json configs[1] (256^3 particles on a 256^3 mesh,
2LPT at a = 0.1, 40 PM drift-kick steps to a = 1) run END TO END with the CPU oracle (oracle/, the
restatement of the reference pinned by tests/test_oracle_golden.py).

    python tests/golden/make_config2.py [N] [NSTEPS]        # ~2 h on 8 cores, ~12 GB of RAM at N = 256

The full fields (3 x 200 MB per checkpoint) cannot live in the repository, so the fixture keeps, per
checkpoint, (i) every field at a fixed random SAMPLE of 32768 particles, (ii) order-independent checksums
over ALL particles / cells (float64 block sums of the painted density on a 32^3 coarse grid, wrapped
int64 sums of the cell indices), (iii) the power spectrum of the painted density.  Inputs are
regenerated from the seed at test time (the white noise is numpy's PCG64 stream, the colouring is the
oracle's linear_field), so nothing but this file and the oracle is needed to reproduce it.

Checkpoints: LPT output (dx, p, f), the force evaluation at the LPT state, the state after drift-kick
steps 1 and 2 (of 40) and after step 40 (a = 1).
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

from oracle import cosmology as OC  # noqa: E402
from oracle import ode as OO  # noqa: E402
from oracle import painting as OP  # noqa: E402
from oracle import pm as OPM  # noqa: E402
from oracle import utils as OU  # noqa: E402

NSAMPLE = 32768
COARSE = 32


def sample_ids(n3):
    return np.sort(np.random.default_rng(123).choice(n3, NSAMPLE, replace=False)).astype(np.int64)


def white_noise(n):
    return np.random.default_rng(0).standard_normal((n, n, n)).astype(np.float32)


def initial_conditions(n, box):
    """The IC array both sides start from (input generation, not part of the parity claim)."""
    from jaxpm_b200.cosmology import Planck15, linear_matter_power  # pure NumPy, no GPU
    c = Planck15()
    return OPM.linear_field(white_noise(n), box, lambda k: linear_matter_power(c, k))


def block_sums(field, coarse=COARSE):
    n = field.shape[0]
    b = n // coarse
    return field.astype(np.float64).reshape(coarse, b, coarse, b, coarse, b).sum(axis=(1, 3, 5))


def cell_index_sums(disp):
    """Wrapped int64 sums over ALL particles of the corner-0 flat cell index (relative rule,
    painting_utils.py:53-65; -1 = dropped) and of index * (particle id + 1)."""
    shape = disp.shape[:3]
    n3 = int(np.prod(shape))
    s0 = np.int64(0)
    s1 = np.int64(0)
    step = 1 << 22
    flat_disp = disp.reshape(-1, 3)
    pm = None
    with np.errstate(over="ignore"):
        for b in range(0, n3, step):
            e = min(b + step, n3)
            ids = np.arange(b, e, dtype=np.int64)
            k = ids % shape[2]
            j = (ids // shape[2]) % shape[1]
            i = ids // (shape[1] * shape[2])
            pm = np.stack([i, j, k], -1).astype(np.int32)
            idx, _ = OP.enmesh_rel(pm, flat_disp[b:e], shape)
            c0 = idx[:, 0]
            ok = np.all((c0 >= 0) & (c0 < np.asarray(shape)), axis=-1)
            f = np.where(ok, (c0[:, 0].astype(np.int64) * shape[1] + c0[:, 1]) * shape[2] + c0[:, 2], -1)
            s0 = s0 + f.sum(dtype=np.int64)
            s1 = s1 + (f * (ids + 1)).sum(dtype=np.int64)
    return np.array([s0, s1], dtype=np.int64)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out_path = os.path.join(HERE, f"config2_{n}.npz")
    shape, box = (n, n, n), (float(n),) * 3
    ocos = OC.Planck15()
    ids = sample_ids(n**3)
    out = {"n": np.int64(n), "nsteps": np.int64(nsteps), "sample_ids": ids}
    t0 = time.time()

    def log(msg):
        print(f"[{time.time() - t0:7.0f} s] {msg}", flush=True)

    def sample(a):
        return np.ascontiguousarray(a.reshape(-1, a.shape[-1])[ids])

    def checkpoint(tag, pos, vel):
        field = OP.cic_paint_dx(pos)
        _, pk = OU.power_spectrum(field, box_shape=box)
        out[f"{tag}_pos"] = sample(pos)
        out[f"{tag}_vel"] = sample(vel)
        out[f"{tag}_pk"] = np.asarray(pk, dtype=np.float64)
        out[f"{tag}_rho_blocks"] = block_sums(field)
        out[f"{tag}_rho_max"] = np.float64(field.max())
        out[f"{tag}_cell_sums"] = cell_index_sums(pos)
        out[f"{tag}_pos_abs_mean"] = np.float64(np.abs(pos.astype(np.float64)).mean())
        np.savez_compressed(out_path, **out)
        log(f"checkpoint {tag}: rho max {field.max():.1f}, |disp| mean {out[f'{tag}_pos_abs_mean']:.3f}")

    ic = initial_conditions(n, box)
    out["ic_sample"] = ic.reshape(-1)[ids]
    out["ic_std"] = np.float64(ic.astype(np.float64).std())
    log("ICs done")
    dx, p, f = OPM.lpt(ocos, ic, a=0.1, order=2)
    out["lpt_dx"], out["lpt_p"], out["lpt_f"] = sample(dx), sample(p), sample(f)
    out["lpt_dx_max"] = np.float64(np.abs(dx).max())
    log("2LPT done")
    dx1 = OPM.lpt(ocos, ic, a=0.1, order=1)[0]
    out["lpt1_dx"] = sample(dx1)
    del dx1, f, ic
    forces = OPM.pm_forces(dx, mesh_shape=shape, paint_absolute_pos=False)
    out["force0"] = sample(forces)
    out["force0_max"] = np.float64(np.abs(forces).max())
    del forces
    checkpoint("lpt", dx, p)
    drift, kick = OO.symplectic_ode(shape, ocos, paint_absolute_pos=False)
    ts = np.linspace(0.1, 1.0, nsteps + 1)
    pos, vel = dx, p
    for s in range(nsteps):
        dt = ts[s + 1] - ts[s]
        pos = (pos + dt * drift(ts[s], vel, None)).astype(pos.dtype)
        vel = (vel + dt * kick(ts[s], pos, None)).astype(vel.dtype)
        log(f"step {s + 1}/{nsteps}")
        if s + 1 in (1, 2, nsteps):
            checkpoint(f"step{s + 1}", pos, vel)
    log(f"wrote {out_path}")


if __name__ == "__main__":
    main()
