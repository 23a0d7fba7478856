"""N > 1 leg of bench.py: the sharded force loop (slab decomposition over N GPUs of one box).

Strong scaling: the SIZE^3 workload is fixed and split over the ranks; `value` = all particles
advanced per second, timed on the device (CUDA events), max over ranks."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def nvlink_counters(gpu_index):
    """(tx_bytes, rx_bytes) summed over the NVLink links of one GPU from the driver's cumulative data counters
    (`nvidia-smi nvlink -gt d`, KiB per link), or None when the tool does not report them."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu_index)], capture_output=True, text=True,
                             timeout=20).stdout
    except Exception:
        return None
    tx = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    if not tx or not rx:
        return None
    return sum(tx) * 1024, sum(rx) * 1024


def run_multi(args, world, rank, dev):
    from bench import METRIC, UNIT, ClockSampler, _peaks, schedule
    from jaxpm_b200 import _lib, halo, ops
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    from jaxpm_b200.distributed import Sharding
    from jaxpm_b200.ode import kick_drift_coefficients
    from jaxpm_b200.pm import linear_field, lpt

    N = args.size
    # default process grid: x slabs up to 4 ranks; at 8 ranks 4x2 pencils - a 512^3 slab is then only 64 planes thick
    # against 2 x 29 ghost planes, and the clustered state loads the slabs unevenly (paint 0.20 ... 0.38 ms over the
    # ranks, 1.62 ms/step) where the pencils stay balanced (0.21 ... 0.25 ms, 1.41 ms/step; profiles/r02m8c_*)
    pdims = (4, 2) if world == 8 else (world, 1)
    if getattr(args, "pdims", None):
        pdims = tuple(int(v) for v in args.pdims.lower().split("x"))
        assert pdims[0] * pdims[1] == world, f"--pdims {args.pdims} needs {pdims[0] * pdims[1]} ranks"
    sh = Sharding(pdims)
    h = args.halo
    shape = (N, N, N)
    cosmo = Planck15()
    K, W = args.steps, args.warmup
    lx, ly = N // pdims[0], N // pdims[1]

    # identical ICs on every rank (generated redundantly, untimed), then keep the local block
    ic = linear_field(shape, (float(N),) * 3, lambda k: linear_matter_power(cosmo, k), seed=0, device=dev)
    dx, p, _ = lpt(cosmo, ic, a=0.1, order=args.lpt_order)
    blk = lambda a: a[sh.rx * lx:(sh.rx + 1) * lx, sh.ry * ly:(sh.ry + 1) * ly].contiguous()
    disp, vel = blk(dx), blk(p)
    del ic, dx, p
    ops.clear_plans()
    torch.cuda.empty_cache()
    n_pre, d, k = schedule(cosmo, args, kick_drift_coefficients)
    ops.axpby(1.0, disp, d[0], vel, out=disp)
    force_mode = args.force_mode if not (args.nccl or args.no_resident) else "spectral"
    stepper = halo.make_stepper(disp, vel, h, sh, resident=not args.no_resident, tile=args.tile,
                                margin=args.margin, fused=not args.nccl, force_mode=force_mode)
    fused = type(stepper).__name__ == "SlabStepper"

    def step(n):
        stepper.step(k[n], d[n + 1] if n + 1 < n_pre + K else 0.0)

    for n in range(n_pre):
        step(n)
    torch.cuda.synchronize()
    # driver NVLink counters around the timed region (rank 0's GPU) - read BEFORE the barrier that starts it: the ranks
    # spin on each other inside the step, a late rank would be timed by all the others
    nv0 = nvlink_counters(dev.index) if rank == 0 else None
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
    nv1 = nvlink_counters(dev.index) if rank == 0 else None
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev = float(t)
    npart = N**3
    value = npart * K / t_dev

    finfo = stepper.force_info() if hasattr(stepper, "force_info") else None
    stepper.store(disp, vel)
    # ---- parity carried by the bench line: the final matter power spectrum of this N-GPU run against the SAME
    #      workload run on ONE GPU (rank 0, resident tile kernels, three-transform forces); device estimator of
    #      jaxpm/utils.py:76-128 on the gathered particle set; max relative difference over the k bins
    parity = None
    if not args.no_parity:
        from jaxpm_b200.painting import cic_paint_dx
        from jaxpm_b200.utils import power_spectrum
        parts = [torch.empty_like(disp) for _ in range(world)] if rank == 0 else None
        dist.gather(disp, parts, dst=0)
        if rank == 0:
            full = torch.cat([torch.cat(parts[rx * pdims[1]:(rx + 1) * pdims[1]], dim=1) for rx in range(pdims[0])], dim=0)
            del parts
            box = (float(N),) * 3
            _, pk_multi = power_spectrum(cic_paint_dx(full), box_shape=box)
            ic = linear_field(shape, box, lambda kk: linear_matter_power(cosmo, kk), seed=0, device=dev)
            dx1, p1, _ = lpt(cosmo, ic, a=0.1, order=args.lpt_order)
            del ic
            d1, v1 = dx1.contiguous(), p1.contiguous()
            del dx1, p1
            ops.axpby(1.0, d1, d[0], v1, out=d1)
            sim1 = ops.Sim(shape, shape, True, dev, tile=args.tile, margin=args.margin)
            sim1.load(d1, v1)
            total = n_pre + K
            for n in range(total):
                sim1.step(k[n], d[n + 1] if n + 1 < total else 0.0)
            sim1.store(d1, v1)
            _, pk_one = power_spectrum(cic_paint_dx(d1), box_shape=box)
            rel = (pk_multi / pk_one - 1).abs()
            parity = {"final_pk_max_rel_diff_vs_1gpu": float(rel.max()), "bins": int(rel.numel()), "tolerance": 1e-4,
                      "median_abs_dpos_cells": float((full - d1).abs().max(-1).values.median()),
                      "against": "same ICs and schedule on one GPU (rank 0), resident tile kernels, spectral forces"}
            del full, d1, v1, sim1
            ops.clear_plans()
            torch.cuda.empty_cache()
        dist.barrier()
    # end to end: host-resident local state in, one step, host-resident state out (every rank)
    e2e_steps = max(1, min(K, args.e2e_steps))
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
        tmin = tt.clone()
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        timing = {"stage_ms_max_over_ranks": {n: round(float(v), 4) for n, v in zip(names, tt)},
                  "stage_ms_min_over_ranks": {n: round(float(v), 4) for n, v in zip(names, tmin)},
                  "ghost_planes_used": stepper.plan.ghost_width(), "ghost_planes_allocated": h}
    # NVLink bytes this rank sends + receives per step on the slab path (remote stores of the two FFT transposes,
    # ghost planes read for the halo reduce, ghost planes written for the halo fill), against the measured
    # 770 GB/s per direction of /opt/skills/guides/B200_PROFILING.md
    nvlink = None
    if fused:
        nzc = (N // 2 + 1 + 7) // 8 * 8
        lxl = N // world
        remote = (world - 1) / world
        pot_steps = finfo["steps_potential"] if finfo else 0
        pot = pot_steps > 0 and finfo["next"] == "potential"     # the timed steps ran the potential chain
        nspec = 2 if pot else 3                                  # AT (1 spectrum) + psi (1) | + T01 (2 spectra)
        transposes = (lxl * N * nzc * 8) * remote * nspec
        ge = timing["ghost_planes_used"]
        ncomp = 1 if pot else 3
        geo = min(ge + 2, h) if pot else ge
        if pdims[1] == 1:
            plane = (N + 8) * (N + 8) * 4
            # ghost planes written to the two neighbours: three force meshes, or psi with 2 more planes for the stencil
            ghosts_out = ncomp * 2 * geo * plane
            ghosts_in = 2 * ge * plane                           # density ghost planes read from them
            regroup_in = regroup_out = 0
        else:
            # pencil grid: the z passes also move the slab's rows between the pencils of the row group (the first
            # transpose of a pencil FFT, done in real space), and the ghost frame has four sides + corners
            Lx, Ly = N // pdims[0], N // pdims[1]
            frame = lambda g: ((Lx + 2 * g) * (Ly + 2 * g) - Lx * Ly) * (N + 8) * 4
            ghosts_out = ncomp * frame(geo)
            ghosts_in = frame(ge)
            regroup_in = lxl * N * N * 4 * (pdims[1] - 1) / pdims[1]
            regroup_out = ncomp * regroup_in
        sent = transposes + ghosts_out + regroup_out
        ghosts_in = ghosts_in + regroup_in
        measured = None
        if nv0 is not None and nv1 is not None:
            measured = {"tx_bytes_per_step": (nv1[0] - nv0[0]) / K, "rx_bytes_per_step": (nv1[1] - nv0[1]) / K,
                        "source": "nvidia-smi nvlink -gt d on rank 0's GPU around the timed steps (all links, KiB counters)"}
        nvlink = {"sent_bytes_per_rank_per_step": int(sent), "read_bytes_per_rank_per_step": int(ghosts_in),
                  "measured_rank0": measured,
                  "sent_GBps_over_whole_step": sent * K / t_dev / 1e9, "peak_GBps_per_direction": 770.0,
                  "frac_of_link_if_not_overlapped": sent / 770e9 / (t_dev / K)}
    peak, peak_kind = _peaks()
    step_alg_bytes = (60 + 64) * npart
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{N}^3 particles on {N}^3 mesh, {args.lpt_order}LPT at a=0.1 then {args.schedule_steps} PM "
                                   f"drift-kick steps to a=1 (relative mode), Planck15, L={N} Mpc/h; timed = the "
                                   f"last {K} steps, {n_pre} untimed before",
                       "l2": "inputs larger than L2",
                       "parallelism": f"{'slab' if pdims[1] == 1 else 'pencil'} pdims={pdims}, halo={h}",
                       "force_mode": force_mode,
                       "resident": not args.no_resident,
                       "exchange": ("halo reduce / FFT transposes / halo fill inside the FFT kernels over NVLink peer "
                                    "memory (slab.py), 4 flag barriers per step, no NCCL on the data path"
                                    + ("; pencil particle domain, x-slab FFT chain, row-group transpose inside the z passes"
                                       if pdims[1] > 1 else "")) if fused
                       else "NCCL send/recv halos + all-to-all FFT transposes (halo.py, pfft.py)"},
            "roofline": {"bound": "hbm", "kernel": "whole step (per-GPU share of 124 B/particle-step)",
                         "achieved": step_alg_bytes * K / t_dev / 1e9 / world, "peak": peak, "peak_kind": peak_kind,
                         "unit": "GB/s", "frac": step_alg_bytes * K / t_dev / 1e9 / world / peak, "traffic": None},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "timing": timing, "nvlink": nvlink, "parity": parity, "force_path": finfo,
        }))
    stepper.close()
    dist.destroy_process_group()
