"""N > 1 leg of bench.py: the sharded force loop (slab decomposition over N GPUs of one box).

Strong scaling: the SIZE^3 workload is fixed and split over the ranks; `value` = all particles
advanced per second, timed on the device (CUDA events), max over ranks."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def run_multi(args, world, rank, dev):
    from bench import METRIC, UNIT, ClockSampler, _peaks, schedule
    from jaxpm_b200 import _lib, halo, ops
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    from jaxpm_b200.distributed import Sharding
    from jaxpm_b200.ode import kick_drift_coefficients
    from jaxpm_b200.pm import linear_field, lpt

    N = args.size
    pdims = (world, 1)
    sh = Sharding(pdims)
    h = args.halo
    shape = (N, N, N)
    cosmo = Planck15()
    K, W = args.steps, args.warmup
    lx, ly = N // pdims[0], N // pdims[1]

    # identical ICs on every rank (generated redundantly, untimed), then keep the local block
    ic = linear_field(shape, (float(N),) * 3, lambda k: linear_matter_power(cosmo, k), seed=0, device=dev)
    dx, p, _ = lpt(cosmo, ic, a=0.1, order=1)
    blk = lambda a: a[sh.rx * lx:(sh.rx + 1) * lx, sh.ry * ly:(sh.ry + 1) * ly].contiguous()
    disp, vel = blk(dx), blk(p)
    del ic, dx, p
    ops.clear_plans()
    torch.cuda.empty_cache()
    n_pre, d, k = schedule(cosmo, args, kick_drift_coefficients)
    ops.axpby(1.0, disp, d[0], vel, out=disp)
    stepper = halo.make_stepper(disp, vel, h, sh, resident=not args.no_resident, tile=args.tile,
                                margin=args.margin, fused=not args.nccl)
    fused = type(stepper).__name__ == "SlabStepper"

    def step(n):
        stepper.step(k[n], d[n + 1] if n + 1 < n_pre + K else 0.0)

    for n in range(n_pre):
        step(n)
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for n in range(n_pre, n_pre + K):
        step(n)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev = float(t)
    npart = N**3
    value = npart * K / t_dev

    # end to end: host-resident local state in, one step, host-resident state out (every rank)
    e2e_steps = max(1, min(K, args.e2e_steps))
    stepper.store(disp, vel)
    ph, vh = disp.cpu().pin_memory(), vel.cpu().pin_memory()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(e2e_steps):
        disp.copy_(ph, non_blocking=True)
        vel.copy_(vh, non_blocking=True)
        stepper.load(disp, vel)
        stepper.step(1e-6, 1e-6)
        stepper.store(disp, vel)
        ph.copy_(disp, non_blocking=True)
        vh.copy_(vel, non_blocking=True)
        torch.cuda.synchronize()
    dist.barrier()
    te = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    bytes_state = 2 * npart * 12
    e2e = {"value": npart * e2e_steps / float(te), "unit": UNIT, "h2d_bytes_per_step": bytes_state,
           "d2h_bytes_per_step": bytes_state, "steps": e2e_steps,
           "entry": "per-rank pinned host state -> sharded PM step -> host state (all ranks concurrently)"}
    # per-stage durations of one more step (CUDA events at the stage boundaries of this rank's stream; a stage
    # ends with the flag barrier that follows it, so waiting for the slowest rank is inside), max over ranks
    timing = None
    if fused:
        acc, reps = {}, 5
        for _ in range(reps):
            for name, ms in stepper.step_profile(0.0, 0.0):
                acc[name] = acc.get(name, 0.0) + ms / reps
        names = list(acc)
        tt = torch.tensor([acc[n] for n in names], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        timing = {"stage_ms_max_over_ranks": {n: round(float(v), 4) for n, v in zip(names, tt)},
                  "ghost_planes_used": stepper.plan.ghost_width(), "ghost_planes_allocated": h}
    # NVLink bytes this rank sends + receives per step on the slab path (remote stores of the two FFT transposes,
    # ghost planes read for the halo reduce, ghost planes written for the halo fill), against the measured
    # 770 GB/s per direction of /opt/skills/guides/B200_PROFILING.md
    nvlink = None
    if fused:
        nzc = (N // 2 + 1 + 7) // 8 * 8
        lxl = N // world
        remote = (world - 1) / world
        transposes = (lxl * N * nzc * 8) * remote * 3          # AT (1 spectrum) + T01 (2 spectra) leaving the rank
        ge = timing["ghost_planes_used"]
        plane = (N + 8) * (N + 8) * 4
        ghosts_out = 3 * 2 * ge * plane                          # force ghost planes written to the two neighbours
        ghosts_in = 2 * ge * plane                               # density ghost planes read from them
        sent = transposes + ghosts_out
        nvlink = {"sent_bytes_per_rank_per_step": int(sent), "read_bytes_per_rank_per_step": int(ghosts_in),
                  "sent_GBps_over_whole_step": sent * K / t_dev / 1e9, "peak_GBps_per_direction": 770.0,
                  "frac_of_link_if_not_overlapped": sent / 770e9 / (t_dev / K)}
    peak, peak_kind = _peaks()
    step_alg_bytes = (60 + 64) * npart
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{N}^3 particles on {N}^3 mesh, 1LPT at a=0.1 then {args.schedule_steps} PM "
                                   f"drift-kick steps to a=1 (relative mode), Planck15, L={N} Mpc/h; timed = the "
                                   f"last {K} steps, {n_pre} untimed before",
                       "l2": "inputs larger than L2", "parallelism": f"slab pdims={pdims}, halo={h}",
                       "resident": not args.no_resident,
                       "exchange": ("halo reduce / FFT transposes / halo fill inside the FFT kernels over NVLink peer "
                                    "memory (slab.py), 4 flag barriers per step, no NCCL on the data path") if fused
                       else "NCCL send/recv halos + all-to-all FFT transposes (halo.py, pfft.py)"},
            "roofline": {"bound": "hbm", "kernel": "whole step (per-GPU share of 124 B/particle-step)",
                         "achieved": step_alg_bytes * K / t_dev / 1e9 / world, "peak": peak, "peak_kind": peak_kind,
                         "unit": "GB/s", "frac": step_alg_bytes * K / t_dev / 1e9 / world / peak, "traffic": None},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "timing": timing, "nvlink": nvlink,
        }))
    stepper.close()
    dist.destroy_process_group()
