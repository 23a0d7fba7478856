#!/usr/bin/env python
"""bench.py — PM particle-steps/s of the B200-native force loop (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl reference]

A "step" is one pass of the hot path over the resident particle set: zero mesh -> CIC paint ->
R2C FFT -> fused Green's x gradient pass -> 3x C2R FFT -> fused read3 + kick + drift
(jaxpm/ode.py:100-117 around jaxpm/pm.py:12-58).  Workload: SIZE^3 particles on a SIZE^3 mesh
(default 512^3, the size the metric is quoted on), Planck15 Gaussian ICs (L = SIZE Mpc/h), 1LPT at
a = 0.1, then W untimed + K timed drift-kick steps towards a = 1 (relative/displacement mode, the
mode of the reference's published runs).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, NumPy/SciPy; JAX is not installable here) on a bounded sub-box of the same workload.
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
def cpu_step_rate(n_sample, steps, warmup=0):
    """Time `steps` drift-kick steps of an n_sample^3 sub-box with the oracle.  Returns
    (particle-steps/s, seconds per step, threads)."""
    import numpy as np
    from oracle import cosmology as OC
    from oracle import ode as OO
    shape = (n_sample,) * 3
    rng = np.random.default_rng(0)
    grid = np.stack(np.meshgrid(*[np.arange(n_sample)] * 3, indexing="ij"), -1).astype(np.float32)
    disp = (0.5 * rng.standard_normal(grid.shape)).astype(np.float32)
    vel = (0.01 * rng.standard_normal(grid.shape)).astype(np.float32)
    cosmo = OC.Planck15()
    OC.growth_tables(cosmo)
    drift, kick = OO.symplectic_ode(shape, cosmo, paint_absolute_pos=False)
    if warmup:
        disp, vel = OO.semi_implicit_euler(drift, kick, disp, vel, 0.1, 0.1 + 0.01 * warmup, warmup)
    t0 = time.perf_counter()
    OO.semi_implicit_euler(drift, kick, disp, vel, 0.2, 0.2 + 0.01 * steps, steps)
    dt = time.perf_counter() - t0
    return n_sample**3 * steps / dt, dt / steps, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = 128 if (args.steps + args.warmup) <= 12 else 64
    rate, sps, cores = cpu_step_rate(n_s, args.steps, args.warmup)
    sample = (f"{args.steps} drift-kick steps of a {n_s}^3-particle / {n_s}^3-mesh sub-box of the "
              f"{args.size}^3 workload (same cell size), oracle NumPy/SciPy port, scipy.fft workers=all")
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
    d, k = kick_drift_coefficients(cosmo, 0.1, 0.1 + total * 0.9 / S, total, "symplectic")
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
    dx, p, _ = lpt(cosmo, ic, a=0.1, order=1)
    del ic
    disp, vel = dx.contiguous(), p.contiguous()
    plan = ops.get_plan(shape, dev)
    # physical schedule: `--schedule-steps` (40, the step count of BASELINE.json's configs) equal steps in
    # a from 0.1 to 1; the timed region is the LAST K steps of that run (the most clustered state), the
    # steps before it are untimed (at least W of them)
    n_pre, d, k = schedule(cosmo, args, kick_drift_coefficients)
    ops.axpby(1.0, disp, d[0], vel, out=disp)
    torch.cuda.empty_cache()
    # resident tile-sorted state (jaxpm_b200/csrc/sim.cu): loaded once, like the LPT set-up
    sim = ops.Sim(shape, shape, True, dev, tile=args.tile, margin=args.margin)
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
        traffic = {"sim_paint": tj["sim_paint_kernel"]["traffic"], "tile_scan+sim_read3_kick_drift": tj["sim_read_kernel"]["traffic"],
                   "fft_z_r2c+ghost_fold": tj["zfwd_kernel"]["traffic"], "fft_y_fwd+transpose": tj["yfwd_kernel"]["traffic"],
                   "fft_x_fwd+greens_grad+ifft_x_x2+transpose": tj["xfused_kernel"]["traffic"],
                   "ifft_y_x3": tj["yinv_kernel"]["traffic"], "ifft_z_c2r_x3+ghost_fill": tj["zinv_kernel"]["traffic"]}
    except Exception:
        pass
    own = {n: v for n, v in kernels.items() if "cuFFT" not in n and n != "mesh_memset" and "direct" not in n}
    dom_name = max(own, key=lambda n: own[n]["s"])
    dom = own[dom_name]
    step_alg_bytes = 60 * npart + 64 * nc
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": dom["GBps"], "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic.get(dom_name),
                "alg_bytes": dom["alg_bytes"],
                "step_achieved": step_alg_bytes * K / t_dev / 1e9,
                "step_frac": step_alg_bytes * K / t_dev / 1e9 / peak,
                "kernels": {n: {"ms": round(v["s"] * 1e3, 4), "GBps": round(v["GBps"], 1),
                                "frac": round(v["frac"], 4)} for n, v in kernels.items()}}
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI entry (pinned host state, H2D + D2H per step) -
    e2e_steps = max(1, min(K, args.e2e_steps))
    ph = torch.empty(disp.shape, dtype=torch.float32).pin_memory()
    vh = torch.empty(vel.shape, dtype=torch.float32).pin_memory()
    ph.copy_(disp)
    vh.copy_(vel)
    ops.pm_step_host_(plan, ph, vh, disp, vel, 0.0, 0.0, True)  # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(e2e_steps):
        ops.pm_step_host_(plan, ph, vh, disp, vel, 1e-6, 1e-6, True)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    bytes_state = 2 * npart * 12
    e2e = {"value": npart * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_state,
           "d2h_bytes_per_step": bytes_state, "steps": e2e_steps,
           "entry": "jpm_pm_step_host_f32 (pinned host pos/vel in, pos/vel out)"}

    # ---- CPU baseline (oracle port) on a bounded sample ----------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu:
        rate, sps, cores = cpu_step_rate(128, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "1 drift-kick step of a 128^3 sub-box (same cell size), oracle NumPy/SciPy port "
                         f"of the reference, {sps:.1f} s"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{N}^3 particles on {N}^3 mesh, 1LPT at a=0.1 then {args.schedule_steps} PM "
                                   f"drift-kick steps to a=1 (relative mode), Planck15, L={N} Mpc/h; timed = the "
                                   f"last {K} steps, {n_pre} untimed before",
                       "l2": "inputs larger than L2 (particle state 3.2 GB, mesh 0.5 GB at 512^3)",
                       "parallelism": "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "sim": {"tile": sim.tile, "margin": sim.margin,
                                      "global_fallback_particles_paint_read": fallbacks[:2],
                                      "generic_stencil_particles_paint_read": fallbacks[2:]},
        }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--tile", type=int, default=16)
    ap.add_argument("--margin", type=int, default=1)
    ap.add_argument("--schedule-steps", type=int, default=40,
                    help="number of equal steps in a from 0.1 to 1 (the timed steps are the last K of them)")
    ap.add_argument("--halo", type=int, default=64, help="halo width of the sharded path (N > 1)")
    ap.add_argument("--no-resident", action="store_true", help="N > 1: order-preserving kernels")
    ap.add_argument("--nccl", action="store_true", help="N > 1: NCCL halo / all-to-all path instead of the fused slab path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--direct", action="store_true",
                    help="also time the order-preserving paint/read kernels of the functional API")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
