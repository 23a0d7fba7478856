#!/usr/bin/env python
"""bench.py — PM particle-steps/s of the B200-native force loop (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl reference] [--grad] [--pdims PXxPY]

A "step" is one pass of the hot path over the resident particle set: zero mesh -> CIC paint -> fused FFT chain with the
k-space kernels inside (three inverse transforms, or one + the 4th-order difference pass where the measured fp32 error
bound allows) -> fused read3 + kick + drift (jaxpm/ode.py:100-117 around jaxpm/pm.py:12-58).  Workload: SIZE^3 particles
on a SIZE^3 mesh (default 512^3, the size the metric is quoted on), Planck15 Gaussian ICs from the device generator
(L = SIZE Mpc/h), 2LPT at a = 0.1, then the 40 drift-kick steps to a = 1 of BASELINE.json's configs (relative /
displacement mode, the mode of the reference's published runs); the LAST K steps are timed (the most clustered
state), at least W run untimed before.

Prints ONE JSON line (rank 0): `value` (device-timed), `roofline` (per-kernel table from CUDA events at the stage
boundaries, chain against 48 B/cell, ncu DRAM traffic), `e2e` (a stream of pinned host states through the pipelined
host entry), `e2e_run` (host ICs -> LPT -> 40 steps -> host state), `api` (pm_forces through the reference's
signature), `lpt`, `parity` (final P(k) against the order-preserving path / the one-GPU run), `cpu_baseline`.
N > 1 (bench_multi.py): strong scaling of the same problem on the fused peer-memory path - x slabs up to 4 ranks,
4x2 pencils at 8 - with min / max stage times over the ranks.  `--grad`: BASELINE.json config 5 (reverse-mode
gradient of the final P(k) w.r.t. the ICs).  `--impl reference` times the CPU restatement of the reference (oracle/,
NumPy/SciPy; JAX is not installable here) on a bounded sub-box of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pm_particle_steps_per_sec"
UNIT = "particle-steps/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy/SciPy restatement of the reference) on a bounded sample
# ---------------------------------------------------------------------------------------
def cpu_step_rate(n_sample, steps, warmup=0, a_start=0.775):
    """Time `steps` drift-kick steps of an n_sample^3 sub-box (same 1 Mpc/h cells, same Planck15 spectrum) with the
    oracle, starting from a CLUSTERED state at the epoch the GPU arm times (the last steps of a 40-step run to a = 1):
    seeded Gaussian ICs, 1LPT displacement / momentum extrapolated to a_start (shell-crossed pancakes and knots -
    the workload family of the GPU arm; evolving the sample with 30 oracle steps would take ~10 minutes).
    Returns (particle-steps/s, seconds per step, threads)."""
    import numpy as np
    from jaxpm_b200.cosmology import Planck15 as P15, linear_matter_power
    from oracle import cosmology as OC
    from oracle import ode as OO
    from oracle import pm as OPM
    shape = (n_sample,) * 3
    rng = np.random.default_rng(0)
    cosmo = OC.Planck15()
    OC.growth_tables(cosmo)
    c = P15()
    wn = rng.standard_normal(shape).astype(np.float32)
    ic = OPM.linear_field(wn, (float(n_sample),) * 3, lambda k: linear_matter_power(c, k))
    disp, vel, _ = OPM.lpt(cosmo, ic, a=a_start, order=1)
    drift, kick = OO.symplectic_ode(shape, cosmo, paint_absolute_pos=False)
    da = 0.9 / 40
    if warmup:
        disp, vel = OO.semi_implicit_euler(drift, kick, disp, vel, a_start - warmup * da, a_start, warmup)
    t0 = time.perf_counter()
    OO.semi_implicit_euler(drift, kick, disp, vel, a_start, a_start + steps * da, steps)
    dt = time.perf_counter() - t0
    return n_sample**3 * steps / dt, dt / steps, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = 128 if (args.steps + args.warmup) <= 20 else 64     # ~10 s per 128^3 step on 16 host cores
    rate, sps, cores = cpu_step_rate(n_s, args.steps, args.warmup)
    sample = (f"{args.steps} drift-kick steps of a {n_s}^3-particle / {n_s}^3-mesh sub-box of the {args.size}^3 "
              f"workload (same 1 Mpc/h cells, Planck15 ICs, clustered 1LPT state at a = 0.775), oracle NumPy/SciPy "
              f"port, scipy.fft workers=all")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.size}^3 particles on {args.size}^3 mesh, PM drift-kick steps "
                               f"(timed on a {n_s}^3 sub-box)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def schedule(cosmo, args, kick_drift_coefficients):
    """(n_pre, d, k): untimed step count and the drift/kick coefficients of n_pre + K equal steps of
    da = 0.9 / schedule_steps starting at a = 0.1 (the timed K steps end at a = 1 when K <= schedule)."""
    S, K, W = args.schedule_steps, args.steps, args.warmup
    n_pre = max(W, S - K)
    total = n_pre + K
    # more steps than the schedule asked for (a large --steps): still end at a = 1, with proportionally smaller steps
    a_end = 0.1 + min(total, S) * 0.9 / S
    d, k = kick_drift_coefficients(cosmo, 0.1, a_end, total, "symplectic")
    return n_pre, d, k


def time_kernel(fn, iters=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from jaxpm_b200 import _lib, ops
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    from jaxpm_b200.ode import kick_drift_coefficients
    from jaxpm_b200.pm import linear_field, lpt

    if world > 1:
        from bench_multi import run_multi  # sharded path (jaxpm_b200/halo.py, pfft.py)
        return run_multi(args, world, rank, dev)

    N = args.size
    shape = (N, N, N)
    npart = N**3
    cosmo = Planck15()
    box = (float(N),) * 3
    K, W = args.steps, args.warmup

    # ---- workload set-up (untimed): ICs -> 1LPT at a=0.1 -> displacement + momentum ------------
    ic = linear_field(shape, box, lambda k: linear_matter_power(cosmo, k), seed=0, device=dev)
    # LPT is set-up, reported separately: 1LPT = 1 fwd + 3 inv FFT + 3 reads, 2LPT = 2 fwd + 12 inv
    lpt_ms = {}
    for order in (2, 1):
        lpt(cosmo, ic, a=0.1, order=order)   # warm (plans, tables)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        dx, p, _ = lpt(cosmo, ic, a=0.1, order=order)
        ev1.record()
        torch.cuda.synchronize()
        lpt_ms[f"order{order}_ms"] = round(ev0.elapsed_time(ev1), 3)
    lpt_order = args.lpt_order
    if lpt_order == 2:
        dx, p, _ = lpt(cosmo, ic, a=0.1, order=2)
    ic_host = ic.cpu().pin_memory() if args.e2e_run else None
    del ic
    disp, vel = dx.contiguous(), p.contiguous()
    del dx, p
    plan = ops.get_plan(shape, dev)
    # physical schedule: `--schedule-steps` (40, the step count of BASELINE.json's configs) equal steps in
    # a from 0.1 to 1; the timed region is the LAST K steps of that run (the most clustered state), the
    # steps before it are untimed (at least W of them)
    n_pre, d, k = schedule(cosmo, args, kick_drift_coefficients)
    ops.axpby(1.0, disp, d[0], vel, out=disp)
    torch.cuda.empty_cache()
    # resident tile-sorted state (jaxpm_b200/csrc/sim.cu): loaded once, like the LPT set-up
    sim = ops.Sim(shape, shape, True, dev, tile=args.tile, margin=args.margin)
    if args.force_mode != "spectral":
        sim.set_force_mode(args.force_mode)
    sim.load(disp, vel)

    def step(n):
        sim.step(k[n], d[n + 1] if n + 1 < n_pre + K else 0.0)

    for n in range(n_pre):
        step(n)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for n in range(n_pre, n_pre + K):
        step(n)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    fallbacks = sim.fallback_counts()
    finfo = sim.force_info()
    sim.store(disp, vel)
    t_dev = e0.elapsed_time(e1) * 1e-3
    value = npart * K / t_dev

    # ---- per-stage durations of the SAME step (CUDA events on the launching stream, recorded inside
    #      jpm_sim_step at every stage boundary; 5 more steps on the evolved particle set) -------------
    peak, peak_kind = _peaks()
    nc = npart
    alg = {  # algorithmic bytes per launch
        "mesh_memset": 4 * nc, "sim_paint": 12 * npart + 4 * nc,
        "tile_scan+sim_read3_kick_drift": 48 * npart + 12 * nc,
        "fft_z_r2c+ghost_fold": 8 * nc, "fft_y_fwd+transpose": 8 * nc,
        "fft_x_fwd+greens_grad+ifft_x_x2+transpose": 12 * nc, "ifft_y_x3": 20 * nc,
        "ifft_z_c2r_x3+ghost_fill": 24 * nc,
        # potential chain: one spectrum through the inverse half, one mesh box per tile in the read
        "fft_x_fwd+greens+ifft_x+transpose": 8 * nc, "ifft_y": 8 * nc, "ifft_z_c2r+ghost_fill": 8 * nc,
        "tile_scan+sim_readpot_kick_drift": 48 * npart + 4 * nc, "fd_gradient": 16 * nc,
        "fft_z_r2c+ghost_fold|fft_y_fwd (chunked pairs)": 16 * nc, "ifft_y_x3|ifft_z_c2r_x3 (chunked pairs)": 44 * nc,
        "ghost_fold": 0, "ghost_fill": 0, "fft_r2c(cuFFT)": 8 * nc, "greens_grad": 16 * nc,
        "ifft_c2r_x3(cuFFT)": 24 * nc,
    }
    acc, reps = {}, 5
    for _ in range(reps):
        for name, ms in sim.step_profile(0.0, 0.0):
            acc[name] = acc.get(name, 0.0) + ms / reps
    kernels = {n: {"s": ms * 1e-3, "alg_bytes": alg.get(n, 0)} for n, ms in acc.items()}
    if args.direct:
        # the order-preserving kernels of the functional API on the same particle set, for comparison
        mesh = torch.zeros(shape, dtype=torch.float32, device=dev)
        f3 = torch.zeros((3, *shape), dtype=torch.float32, device=dev)
        scratch_p, scratch_v = disp.clone(), vel.clone()
        t_zero = time_kernel(lambda: mesh.zero_())

        def k_paint_direct():
            mesh.zero_()
            ops.cic_paint_dx_(mesh, disp)

        kernels["direct_paint_dx(order-preserving)"] = {"s": time_kernel(k_paint_direct) - t_zero,
                                                        "alg_bytes": 12 * npart + 4 * nc}
        kernels["direct_read3_kick_drift(order-preserving)"] = {
            "s": time_kernel(lambda: ops.read3_kick_drift_(f3, scratch_p, scratch_v, 0.0, 0.0, True)),
            "alg_bytes": 48 * npart + 12 * nc}
        del mesh, f3, scratch_p, scratch_v
    for v in kernels.values():
        v["GBps"] = v["alg_bytes"] / v["s"] / 1e9 if v["s"] > 0 else 0.0
        v["frac"] = v["GBps"] / peak
    # DRAM traffic of each kernel from the committed `ncu --set full` capture of this workload (profiles/), per launch
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", f"traffic_{N}.json")) as f:
            tj = json.load(f)["kernels"]
        names = {"sim_paint": "sim_paint_kernel", "tile_scan+sim_read3_kick_drift": "sim_read_kernel",
                 "tile_scan+sim_readpot_kick_drift": "sim_readpot_kernel",
                 "fft_z_r2c+ghost_fold": "zfwd_kernel", "fft_y_fwd+transpose": "yfwd_kernel",
                 "fft_x_fwd+greens_grad+ifft_x_x2+transpose": "xfused_kernel", "ifft_y_x3": "yinv_kernel",
                 "ifft_z_c2r_x3+ghost_fill": "zinv_kernel", "fft_x_fwd+greens+ifft_x+transpose": "xpot_kernel",
                 "ifft_y": "ypot_kernel", "ifft_z_c2r+ghost_fill": "zinv_kernel(potential)", "fd_gradient": "fdgrad_kernel"}
        traffic = {stage: tj[kern]["traffic"] for stage, kern in names.items() if kern in tj}
    except Exception:
        pass
    own = {n: v for n, v in kernels.items() if "cuFFT" not in n and n != "mesh_memset" and "direct" not in n}
    dom_name = max(own, key=lambda n: own[n]["s"])
    dom = own[dom_name]
    step_alg_bytes = 60 * npart + 64 * nc
    # the FFT chain as a whole against SURVEY.md's 48 B/cell (8 fwd + 16 k-space + 24 inv), next to the per-pass numbers
    chain_s = sum(v["s"] for n, v in kernels.items() if n.startswith("fft_") or n.startswith("ifft_"))
    chain = {"ms": round(chain_s * 1e3, 4), "alg_bytes_48_per_cell": 48 * nc,
             "GBps_on_48B": round(48 * nc / chain_s / 1e9, 1) if chain_s > 0 else None,
             "frac_on_48B": round(48 * nc / chain_s / 1e9 / peak, 4) if chain_s > 0 else None}
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": dom["GBps"], "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic.get(dom_name),
                "alg_bytes": dom["alg_bytes"],
                "step_achieved": step_alg_bytes * K / t_dev / 1e9,
                "step_frac": step_alg_bytes * K / t_dev / 1e9 / peak,
                "fft_chain": chain,
                "kernels": {n: {"ms": round(v["s"] * 1e3, 4), "GBps": round(v["GBps"], 1),
                                "frac": round(v["frac"], 4)} for n, v in kernels.items()}}
    torch.cuda.empty_cache()

    # ---- the reference's functional API on the same (clustered) state: pm_forces(disp) -> forces, particle order
    #      kept (jaxpm/pm.py:12-58): tile sort on entry, shared-memory paint, fused FFT chain, shared-memory gather
    from jaxpm_b200 import pm as jpm_pm
    api = {}
    for label, fast in (("api_pm_forces_ms", True), ("api_pm_forces_slow_path_ms", False)):
        if not fast and not args.direct:
            continue
        jpm_pm._FAST_API = fast
        api[label] = round(time_kernel(lambda: jpm_pm.pm_forces(disp, mesh_shape=shape, paint_absolute_pos=False),
                                       iters=3, warm=1) * 1e3, 3)
    jpm_pm._FAST_API = True
    ops._force_sims.clear()
    torch.cuda.empty_cache()

    # ---- parity carried by the bench line: the final matter power spectrum of the timed run (device estimator,
    #      jaxpm/utils.py:76-128) against the SAME workload on the order-preserving kernels + cuFFT, three-transform
    #      forces (the path the round-1 parity suite pins to the oracle); max relative difference over the k bins
    from jaxpm_b200.painting import cic_paint_dx
    from jaxpm_b200.utils import power_spectrum
    parity = None
    if not args.no_parity:
        total = n_pre + K
        _, pk_fast = power_spectrum(cic_paint_dx(disp), box_shape=box)
        ic2 = linear_field(shape, box, lambda kk: linear_matter_power(cosmo, kk), seed=0, device=dev)
        dx2, p2, _ = lpt(cosmo, ic2, a=0.1, order=lpt_order)
        del ic2
        d2, v2 = dx2.contiguous(), p2.contiguous()
        del dx2, p2
        ops.axpby(1.0, d2, d[0], v2, out=d2)
        for n in range(total):
            ops.pm_step_(plan, d2, v2, k[n], d[n + 1] if n + 1 < total else 0.0, True)
        _, pk_ref = power_spectrum(cic_paint_dx(d2), box_shape=box)
        rel = (pk_fast / pk_ref - 1).abs()
        med = float((d2 - disp).abs().max(-1).values.median())
        parity = {"final_pk_max_rel_diff": float(rel.max()), "bins": int(rel.numel()), "tolerance": 1e-4,
                  "median_abs_dpos_cells": med,
                  "against": "same ICs and schedule on the order-preserving kernels + cuFFT (functional-API slow path)"}
        del d2, v2
        torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI entry on the SAME tile kernels (pinned host state, H2D + D2H of
    #      the full particle state every step): PCIe-bound by construction (24 B in + 24 B out per particle-step)
    e2e_steps = max(1, args.e2e_steps)
    bytes_state = 2 * npart * 12
    # (a) one state at a time (latency form): H2D -> sort -> step -> un-sort -> D2H, serial
    ph = torch.empty(disp.shape, dtype=torch.float32).pin_memory()
    vh = torch.empty(vel.shape, dtype=torch.float32).pin_memory()
    ph.copy_(disp)
    vh.copy_(vel)
    sim.step_host(ph, vh, disp, vel, 0.0, 0.0)  # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(min(e2e_steps, 3)):
        sim.step_host(ph, vh, disp, vel, 1e-6, 1e-6)
    torch.cuda.synchronize()
    t_serial = (time.perf_counter() - t0) / min(e2e_steps, 3)
    # (b) a stream of host-resident states (throughput form, the headline): the same per-step work and bytes, the
    #     upload of state b + 1, the step of state b and the download of state b - 1 overlapped on three streams
    #     (jpm_sim_steps_host_f32); a ring of three pinned host state buffers
    ring_p = [ph] + [torch.empty_like(ph).pin_memory() for _ in range(2)]
    ring_v = [vh] + [torch.empty_like(vh).pin_memory() for _ in range(2)]
    for t in ring_p[1:]:
        t.copy_(ph)
    for t in ring_v[1:]:
        t.copy_(vh)
    del disp, vel
    torch.cuda.empty_cache()
    sim.steps_host(ring_p[:2], ring_v[:2], [0.0] * 2, [0.0] * 2)  # warm (allocates the staging buffers)
    torch.cuda.synchronize()
    lp0 = _lib.launch_count()
    t0 = time.perf_counter()
    sim.steps_host([ring_p[i % 3] for i in range(e2e_steps)], [ring_v[i % 3] for i in range(e2e_steps)],
                   [1e-6] * e2e_steps, [1e-6] * e2e_steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    e2e = {"value": npart * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_state,
           "d2h_bytes_per_step": bytes_state, "steps": e2e_steps,
           "entry": "jpm_sim_steps_host_f32: a stream of pinned host (pos, vel) states, each H2D -> tile sort -> resident "
                    "step -> un-sort -> D2H; the three legs of consecutive states overlap (two copy streams + compute)",
           "pcie_GBps_each_way": round(bytes_state * e2e_steps / t_e2e / 1e9, 1),
           "gpu_launches": _lib.launch_count() - lp0,
           "serial_one_state": {"value": npart / t_serial, "ms": round(t_serial * 1e3, 2),
                                "entry": "jpm_sim_step_host_f32 (no overlap: latency of one state)"}}
    disp = torch.empty(ph.shape, dtype=torch.float32, device=dev)
    vel = torch.empty(vh.shape, dtype=torch.float32, device=dev)
    del ph, vh, ring_p, ring_v
    # run-level end to end, the call a user of the reference makes (notebooks/05-MultiHost_PM.py:85-135): host ICs in,
    # LPT + the whole step schedule on the device, host particle state out
    e2e_run = None
    if args.e2e_run:
        total = n_pre + K
        out_p = torch.empty(disp.shape, dtype=torch.float32).pin_memory()
        out_v = torch.empty(vel.shape, dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        icd = ic_host.to(dev, non_blocking=True)
        dxr, pr, _ = lpt(cosmo, icd, a=0.1, order=lpt_order)
        dr, vr = dxr.contiguous(), pr.contiguous()
        ops.axpby(1.0, dr, d[0], vr, out=dr)
        sim.load(dr, vr)
        for n in range(total):
            sim.step(k[n], d[n + 1] if n + 1 < total else 0.0)
        sim.store(dr, vr)
        out_p.copy_(dr, non_blocking=True)
        out_v.copy_(vr, non_blocking=True)
        torch.cuda.synchronize()
        t_run = time.perf_counter() - t0
        e2e_run = {"value": npart * total / t_run, "unit": UNIT, "seconds": round(t_run, 4), "steps": total,
                   "lpt_order": lpt_order, "h2d_bytes": int(ic_host.numel() * 4), "d2h_bytes": bytes_state,
                   "entry": "host ICs -> lpt -> resident steps -> host (pos, vel)"}
        del icd, dxr, pr, dr, vr, out_p, out_v

    # ---- CPU baseline (oracle port) on a bounded sample ----------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu:
        rate, sps, cores = cpu_step_rate(128, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "1 drift-kick step of a 128^3 sub-box (same 1 Mpc/h cells, Planck15 ICs, clustered 1LPT "
                         f"state at a = 0.775), oracle NumPy/SciPy port of the reference, {sps:.1f} s"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{N}^3 particles on {N}^3 mesh, {lpt_order}LPT at a=0.1 then {args.schedule_steps} PM "
                                   f"drift-kick steps to a=1 (relative mode), Planck15, L={N} Mpc/h; timed = the "
                                   f"last {K} steps, {n_pre} untimed before",
                       "l2": "inputs larger than L2 (particle state 3.2 GB, mesh 0.5 GB at 512^3)",
                       "parallelism": "single GPU", "force_mode": args.force_mode},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_run": e2e_run, "gpu_launches": launches,
            "api": api, "lpt": lpt_ms, "parity": parity, "force_path": finfo,
            "clocks": clocks, "sim": {"tile": sim.tile, "margin": sim.margin,
                                      "global_fallback_particles_paint_read": fallbacks[:2],
                                      "generic_stencil_particles_paint_read": fallbacks[2:]},
        }))


def run_grad(args):
    """BASELINE.json config 5: reverse-mode gradient of the final matter power spectrum with respect to the initial
    conditions (SIZE^3 particles / mesh, default 256^3): ic -> lpt -> K drift-kick steps (per-step recompute) ->
    cic_paint_dx -> power_spectrum -> scalar -> backward.  One JSON line: forward+backward particle-steps/s."""
    import numpy as np
    import torch
    from jaxpm_b200 import _lib
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    from jaxpm_b200.ode import nbody_kick_drift_grad
    from jaxpm_b200.painting import cic_paint_dx
    from jaxpm_b200.pm import linear_field, lpt
    from jaxpm_b200.utils import power_spectrum
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    N, K = args.size, args.steps
    shape, box = (N, N, N), (float(N),) * 3
    cosmo = Planck15()
    ic0 = linear_field(shape, box, lambda k: linear_matter_power(cosmo, k), seed=0, device=dev)

    def run():
        ic = ic0.clone().requires_grad_(True)
        dx, p, _ = lpt(cosmo, ic, a=0.1, order=args.lpt_order)
        pos, vel = nbody_kick_drift_grad(cosmo, dx, p, 0.1, 1.0, K, paint_absolute_pos=False)
        _, pk = power_spectrum(cic_paint_dx(pos), box_shape=box)
        loss = pk.log().sum()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        loss.backward()
        torch.cuda.synchronize()
        return ic.grad, float(loss), t1

    run()   # warm
    torch.cuda.reset_peak_memory_stats()
    l0 = _lib.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    grad, loss, t1 = run()
    t2 = time.perf_counter()
    npart = N**3
    print(json.dumps({
        "metric": "reverse_mode_pm_particle_steps_per_sec", "value": npart * K / (t2 - t0), "unit": UNIT, "n_gpus": 1,
        "steps": K, "forward_ms": (t1 - t0) * 1e3, "backward_ms": (t2 - t1) * 1e3, "higher_is_better": True,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"d sum_b log P(k_b) / d initial_conditions: {N}^3 particles on {N}^3 mesh, "
                               f"{args.lpt_order}LPT at a=0.1, {K} drift-kick steps to a=1 (per-step recompute), "
                               "cic_paint_dx, power_spectrum; adjoint passes per step: force meshes recomputed on the fused "
                               "chain, readgrad3 (one gather pass), paint3 (one scatter pass), real-space divergence + "
                               "one transform pair on the potential chain, paint adjoint; 2LPT source VJP, P(k) adjoint"},
        "gpu_launches": _lib.launch_count() - l0, "loss": loss, "grad_rms": float(grad.square().mean().sqrt()),
        "grad_finite": bool(torch.isfinite(grad).all()), "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--force-mode", default="auto", choices=["spectral", "potential", "auto"],
                    help="force path of the resident step (include/jaxpm_b200.h): three inverse transforms, one + "
                         "difference stencil, or per step by the measured fp32 error bound")
    ap.add_argument("--lpt-order", type=int, default=2, choices=[1, 2], help="LPT order of the initial state")
    ap.add_argument("--no-parity", action="store_true", help="skip the P(k) parity leg")
    ap.add_argument("--grad", action="store_true",
                    help="BASELINE.json config 5 instead: reverse-mode gradient of the final P(k) w.r.t. the ICs (use --size 256)")
    ap.add_argument("--no-e2e-run", dest="e2e_run", action="store_false", help="skip the run-level end-to-end leg")
    ap.add_argument("--tile", type=int, default=16)
    ap.add_argument("--margin", type=int, default=1)
    ap.add_argument("--schedule-steps", type=int, default=40,
                    help="number of equal steps in a from 0.1 to 1 (the timed steps are the last K of them)")
    ap.add_argument("--halo", type=int, default=64, help="halo width of the sharded path (N > 1)")
    ap.add_argument("--no-resident", action="store_true", help="N > 1: order-preserving kernels")
    ap.add_argument("--pdims", default=None, help="N > 1: process grid PXxPY (default Nx1 slabs; PY > 1 = pencil particle domains; both on the fused peer-memory path unless --nccl)")
    ap.add_argument("--nccl", action="store_true", help="N > 1: NCCL halo / all-to-all path instead of the fused slab path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--direct", action="store_true",
                    help="also time the order-preserving paint/read kernels of the functional API")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.grad:
        return run_grad(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
