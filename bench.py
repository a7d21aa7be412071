#!/usr/bin/env python
"""Benchmark of the power-spectrum-likelihood hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the CPU path, on the box's host cores

One step = ONE batched emcee step of BASELINE config 3: 64 walkers x 256 x 256 x 1024 float32 lightcone, 4 chunks,
20 log-k bins, inputs already resident in HBM.  With --gpus N the 64 walkers are sharded 64/N per GPU (strong
scaling) and the per-walker lnL are all-gathered over NCCL; `weak` in the line is the same step with 64
walkers on every GPU.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (256, 256, 1024)
NCHUNKS = 4
NBINS = 20
CELL = 1.5
BOX_LEN = SHAPE[0] * CELL
MIN_K, MAX_K, MODEL_UNC = 0.1, 1.0, 0.15
MU, SIGMA = 20.0, 10.0
METRIC = "ps_likelihood_evals_per_sec_256x256x1024_lightcone"
UNIT = "evals/s"
MODE = "spherical"
NREDSHIFT = 1   # coeval configurations: cubes per walker (one spectrum and one data set per redshift)
WORKLOAD = "BASELINE config 3: one 64-walker emcee step"

# The headline metric is quoted on config 3 (default).  --config 1 / 5 run the other lightcone configurations
# of BASELINE.json for information (they are parity-test cases, not the bench line).
CONFIGS = {
    3: dict(shape=(256, 256, 1024), nchunks=4, mode="spherical", metric=METRIC, workload=WORKLOAD),
    1: dict(shape=(128, 128, 1024), nchunks=4, mode="spherical",
            metric="ps_likelihood_evals_per_sec_128x128x1024_lightcone", workload="BASELINE config 1"),
    5: dict(shape=(256, 256, 2048), nchunks=8, mode="cylindrical",
            metric="cylindrical_ps_evals_per_sec_256x256x2048_lightcone", workload="BASELINE config 5"),
    2: dict(shape=(200, 200, 200), nchunks=1, mode="spherical", nredshift=3,
            metric="ps_likelihood_evals_per_sec_3x200cube_coeval", workload="BASELINE config 2"),
    4: dict(shape=(512, 512, 512), nchunks=1, mode="spherical", nredshift=16,
            metric="ps_likelihood_evals_per_sec_16x512cube_coeval", workload="BASELINE config 4"),
}


def select_config(n):
    global SHAPE, NCHUNKS, BOX_LEN, METRIC, MODE, WORKLOAD, NREDSHIFT
    c = CONFIGS[n]
    SHAPE, NCHUNKS, MODE, METRIC, WORKLOAD = c["shape"], c["nchunks"], c["mode"], c["metric"], c["workload"]
    NREDSHIFT = c.get("nredshift", 1)
    BOX_LEN = SHAPE[0] * CELL


def workload_string(total_walkers):
    """The same string in both arms (`same_config`): what one step evaluates."""
    cyl = MODE == "cylindrical"
    return (f"{WORKLOAD}: {total_walkers} walkers x "
            + (f"{NREDSHIFT} coeval cubes " if NREDSHIFT > 1 else "")
            + f"{SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} float32, {NCHUNKS} chunk(s), "
            f"{NBINS} {'k_perp x k_par' if cyl else 'log-k'} bins")


def default_total_walkers(config):
    return {4: 2}.get(config, 64)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes_per_eval():
    n0, n1, n2 = SHAPE
    nz = n2 // NCHUNKS
    N, Nh = n0 * n1 * nz, n0 * n1 * (nz // 2 + 1)
    ns = NCHUNKS * NREDSHIFT  # spectra per evaluation
    return ns * (4 * N + 16 * Nh), ns * (4 * N + 8 * Nh), ns * 8 * Nh


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    The sampler is started ahead of the warm-up (nvidia-smi needs ~0.1 s to produce its first line) and
    every line carries a timestamp; only samples inside the timed window are summarised.
    """

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime

        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [nm for nm, v in zip(names, f[4:8])
                                                            if v.lower().startswith("active")]))
            except ValueError:
                continue
        if not rows:
            return None
        sel = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not sel:  # region shorter than the sampling period: nearest samples
            sel = sorted(rows, key=lambda r: abs(r[0] - (t0 or r[0])))[:3]
            window = "nearest samples (timed region shorter than the sampling period)"
        reasons = sorted({x for r in sel for x in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in sel), "sm_max_mhz": max(r[2] for r in sel),
                "reasons": reasons, "samples": len(sel), "window": window}


# ---------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port of the reference's powerbox path (float32 FFT, as
# NumPy >= 2 / pyFFTW run it), one process per core -- the reference's own parallel model
# (ProcessPoolExecutor over walkers, src/py21cmmc/mcmc.py:580-583).
# ---------------------------------------------------------------------------------------------
_CPU_BOX = {}


def _cpu_make_box(slot):
    """Untimed: the synthetic chunk a worker process transforms -- one box per process, made once (the cost of
    the transform does not depend on the values)."""
    import numpy as np

    if "box" not in _CPU_BOX:
        rng = np.random.default_rng(1234 + os.getpid())
        nz = SHAPE[2] // NCHUNKS
        _CPU_BOX["box"] = (rng.standard_normal((SHAPE[0], SHAPE[1], nz), dtype=np.float32) * SIGMA + MU)
    return slot


def _cpu_chunk_task(slot):
    import numpy as np

    from oracle import ps_oracle as O

    _cpu_make_box(slot)   # no-op once the process has its box (cpu_prepare and the warm-up steps see to that)
    box = _CPU_BOX["box"]
    nz = box.shape[2]
    t0 = time.perf_counter()
    L = (BOX_LEN, BOX_LEN, nz * CELL)
    if MODE == "cylindrical":
        w2 = np.ones(box.shape[:2])
        w2[box.shape[0] // 2, box.shape[0] // 2] = 0
        p = O.get_power(box, L, bins=NBINS, res_ndim=2, k_weights=w2, fft_dtype=np.float32)[0]
        return time.perf_counter() - t0, float(np.nansum(p))
    p, k = O.compute_power_lightcone(box, L, NBINS, log_bins=True, fft_dtype=np.float32)
    delta = p * k**3 / (2 * np.pi**2)
    return time.perf_counter() - t0, float(np.nansum(delta))


def cpu_prepare(pool, workers):
    """Generate the boxes in the worker processes before anything is timed."""
    list(pool.map(_cpu_make_box, range(8 * workers), chunksize=1))


def cpu_step(pool, workers, step):
    """One bounded sample: `workers` chunks (one spectrum each) in parallel, boxes already in the workers' memory.
    Returns evals/s."""
    t0 = time.perf_counter()
    res = list(pool.map(_cpu_chunk_task, range(workers), chunksize=1))
    dt = time.perf_counter() - t0
    return workers / (NCHUNKS * NREDSHIFT) / dt, dt


def cpu_workers():
    """All host cores, bounded by memory: one 256^3 chunk needs ~2 GB inside the NumPy port."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        import psutil

        n = min(n, max(1, int(psutil.virtual_memory().available // (3 << 30))))
    except Exception:
        n = min(n, 16)
    return max(1, min(n, 128))


def cpu_sample_string(workers):
    nzc = SHAPE[2] // NCHUNKS
    return (f"{workers} chunks of {SHAPE[0]}x{SHAPE[1]}x{nzc} (= {workers / (NCHUNKS * NREDSHIFT):g} evals) per step, "
            "one process per core, boxes generated before the timed region, oracle port of powerbox.get_power in "
            "float32-FFT mode (NumPy pocketfft)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from concurrent.futures import ProcessPoolExecutor

    workers = cpu_workers()
    with ProcessPoolExecutor(max_workers=workers) as pool:
        cpu_prepare(pool, workers)
        for w in range(args.warmup):
            cpu_step(pool, workers, w)
        t0 = time.perf_counter()
        for s in range(args.steps):
            cpu_step(pool, workers, s)
        total = time.perf_counter() - t0
    evals = args.steps * workers / (NCHUNKS * NREDSHIFT)
    value = evals / total
    sample = cpu_sample_string(workers)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 FFT / f64 bin accumulation",
        "data": "synthetic",
        "config": {"workload": workload_string(args.walkers), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
class _PinnedSim:
    """Simulator stand-in for the plugin-level e2e: walker w's lightcone lives in page-locked host memory (what a
    21cmFAST wrapper that writes into pre-pinned walker buffers hands to the chain, core.py:545-561)."""

    def __init__(self, boxes):
        self.boxes = boxes

    def __call__(self, params):
        return self.boxes[int(round(float(params.walker))) if len(params) else 0]


def _timed_steps(torch, dist, world, dev, step, nsteps):
    """nsteps calls of step() between barrier + synchronize on both sides; CUDA events; max over ranks (ms)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for _ in range(nsteps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall1 = time.time()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), wall0, wall1


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import py21cmmc_b200  # noqa: F401
    from py21cmmc_b200 import chain as chain_mod, core as core_mod, likelihood as lk_mod
    from py21cmmc_b200.plan import PowerPlan, pinned_empty

    total_walkers = args.walkers            # one emcee step of the whole job
    if total_walkers % world:
        raise SystemExit(f"--walkers {total_walkers} must be divisible by the number of GPUs ({world})")
    walkers = total_walkers // world        # this rank's shard (strong scaling: fixed total work)
    bounds = [(i * SHAPE[2] // NCHUNKS, (i + 1) * SHAPE[2] // NCHUNKS) for i in range(NCHUNKS)]
    lengths = [(BOX_LEN, BOX_LEN, (b - a) * CELL) for a, b in bounds]
    cyl = MODE == "cylindrical"
    plan = PowerPlan(SHAPE, bounds, lengths, NBINS, not cyl, True, False, False, "legacy_1.001", MODE, local,
                     args.path, workspace_bytes=args.workspace_mb << 20)

    # synthetic Gaussian brightness-temperature boxes, generated on the device.  The weak-scaling
    # record needs `total_walkers` boxes on every rank; the strong step uses the first `walkers` of them.  Walker w
    # of the job has the same seed whatever the number of GPUs.
    nweak = total_walkers if (world > 1 and not args.no_weak) else walkers
    nboxes_all = nweak * NREDSHIFT
    boxes_all = torch.empty((nboxes_all,) + SHAPE, dtype=torch.float32, device=dev)
    for i in range(nboxes_all):
        w = i // NREDSHIFT
        gw = (rank * walkers + w) if w < walkers else (total_walkers + rank * total_walkers + w)
        g = torch.Generator(device=dev).manual_seed(1234 + gw * NREDSHIFT + (i % NREDSHIFT))
        boxes_all[i].normal_(MU, SIGMA, generator=g)
    nboxes = walkers * NREDSHIFT
    boxes = boxes_all[:nboxes]
    # lnL of this rank's walkers is written straight into its slice of the gather buffer (in-place all-gather)
    gathered = torch.zeros((world * nboxes,), dtype=torch.float64, device=dev)
    lnl = gathered[rank * nboxes:(rank + 1) * nboxes]
    if not cyl:
        # "data" = spectrum of walker 0 of rank 0 (the simulate=True convention), broadcast to every rank
        power0 = plan.power_device(boxes_all[:NREDSHIFT].contiguous())
        kc = torch.from_numpy(np.concatenate(list(plan.kcentres))).to(dev)
        delta0 = power0 * kc**3 / (2 * np.pi**2)
        if world > 1:
            dist.broadcast(delta0, 0)
        data = delta0.cpu().numpy().reshape(NREDSHIFT, NCHUNKS, NBINS)
        kmask = np.repeat(((plan.kcentres >= MIN_K) & (plan.kcentres <= MAX_K)).astype(np.uint8)[None], NREDSHIFT, 0)
        plan.set_data(np.nan_to_num(data), None, kmask & np.isfinite(data), MODEL_UNC, True)

        def compute(b=boxes, out=lnl):
            plan.lnl_device(b, lnl=out)
    else:  # config 5: the 2-D spectra are the product (no chi^2 is defined on them in the reference)
        p2d = torch.empty((nboxes_all, plan.out_per_box), dtype=torch.float64, device=dev)

        def compute(b=boxes, out=lnl):
            plan.power_device(b, out=p2d[:b.shape[0]])
            out.copy_(p2d[:b.shape[0], 0])

    def step():
        compute()
        if world > 1:
            dist.all_gather_into_tensor(gathered, lnl)

    use_graph = not args.no_graph
    plan.set_graphs(use_graph)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = plan.info()["launches_last_call"]
    ms_total, wall0, wall1 = _timed_steps(torch, dist, world, dev, step, args.steps)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = total_walkers / (ms_per_step * 1e-3)
    graph_stats = plan.graph_stats()
    lnl_host = lnl.cpu().numpy().copy()
    self_check = float(lnl_host[0]) if (rank == 0 and not cyl) else None

    # ---- per-kernel times (CUDA events recorded by the library around its launches; direct launches, no graph)
    plan.set_graphs(False)
    plan.set_profiling(True)
    ms_direct, _, _ = _timed_steps(torch, dist, world, dev, compute, args.steps)
    prof = plan.profile()
    plan.set_profiling(False)
    plan.set_graphs(use_graph)
    kernels_ms = sum(v[0] for v in prof.values()) / args.steps

    # ---- where a step's time goes beyond its kernels (launch gaps, the collective), for the scaling residual
    residual = None
    if world > 1:
        ms_nogather, _, _ = _timed_steps(torch, dist, world, dev, compute, args.steps)
        residual = {"step_ms": ms_per_step, "step_ms_without_all_gather": ms_nogather / args.steps,
                    "kernels_ms_rank0": kernels_ms,
                    "all_gather_ms": ms_per_step - ms_nogather / args.steps,
                    "launch_gaps_and_tails_ms": ms_nogather / args.steps - kernels_ms,
                    "note": "kernels = sum of the CUDA-event times of the step's kernels on rank 0 (direct launches); "
                            "the step is timed as the max over ranks"}

    # ---- sustained: the same step looped for >= args.sustain_s seconds, its own clocks sample
    sustained = None
    if args.sustain_s > 0:
        nsus = max(args.steps, int(args.sustain_s * 1e3 / ms_per_step) + 1)
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
            time.sleep(0.15)
        ms_sus, w0, w1 = _timed_steps(torch, dist, world, dev, step, nsus)
        c2 = s2.stop(w0, w1) if rank == 0 else None
        sustained = {"value": total_walkers / (ms_sus / nsus * 1e-3), "unit": UNIT, "steps": nsus,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / nsus, "clocks": c2}

    # ---- weak scaling record (N > 1): every GPU evaluates all `total_walkers` walkers per step
    weak = None
    if world > 1 and not args.no_weak:
        gathered_w = torch.zeros((world * nboxes_all,), dtype=torch.float64, device=dev)
        lnl_w = gathered_w[rank * nboxes_all:(rank + 1) * nboxes_all]

        def step_weak():
            compute(boxes_all, lnl_w)
            dist.all_gather_into_tensor(gathered_w, lnl_w)

        for _ in range(3):
            step_weak()
        ms_w, _, _ = _timed_steps(torch, dist, world, dev, step_weak, max(3, args.steps // 2))
        ms_w /= max(3, args.steps // 2)
        weak = {"value": world * total_walkers / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w,
                "walkers_per_gpu": total_walkers, "scaling": "weak"}

    # ------------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not args.no_e2e and not cyl and NREDSHIFT == 1:
        ew = min(walkers, args.e2e_walkers)
        esteps = max(1, min(args.steps, args.e2e_steps))
        # walker-owned page-locked lightcones (the layout a simulator wrapper with pinned walker buffers produces)
        hboxes = []
        for w in range(ew):
            hb = pinned_empty(SHAPE)
            torch.from_numpy(hb).copy_(boxes[w])
            hboxes.append(hb)
        torch.cuda.synchronize()
        # (1) the plugin: chain.call_batch -> Likelihood1DPowerLightcone.likelihood_batch on contexts holding NumPy
        #     lightcones; cores, H2D of every lightcone, kernels, D2H of lnL and the Delta^2 blobs, model dicts, store
        from py21cmmc_b200.chain import Params

        lk_mod.set_defaults(device=local, edge_rule="legacy_1.001", kernel_path=args.path)
        up = core_mod.UserParams(BOX_LEN=BOX_LEN, HII_DIM=SHAPE[0])
        sim = _PinnedSim(hboxes)
        core = core_mod.CoreLightConeModule(sim, up)
        lk = lk_mod.Likelihood1DPowerLightcone(nchunks=NCHUNKS, n_psbins=NBINS, min_k=MIN_K, max_k=MAX_K,
                                               model_uncertainty=MODEL_UNC, use_data=False, device=local,
                                               edge_rule="legacy_1.001", kernel_path=args.path)
        ch = chain_mod.build_computation_chain(core, lk, params=Params(("walker", [0.0, 0.0, 1e6, 1.0])))
        # the data every rank compares with: Delta^2 of walker 0 of rank 0 (as in the device-resident step)
        lk.data = [{"k": plan.kcentres[c], "delta": np.nan_to_num(data[0, c])} for c in range(NCHUNKS)]
        lk.noise = None
        plist = [[float(w)] for w in range(ew)]
        res = ch.call_batch(plist)   # warm-up (staging buffers, plan of the plugin's geometry)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(esteps):
            res = ch.call_batch(plist)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt_plugin = float(dt.item())
        l_plugin = np.array([r[0] for r in res])
        # (2) the C ABI alone: ps21_lnl_host_ptrs on the same buffers
        plan.lnl_host(hboxes)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            l_h, _, _ = plan.lnl_host(hboxes)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt_abi = float(dt.item())
        # (2b) the same plugin call on ordinary (pageable) NumPy lightcones, as an unmodified simulator returns them
        sim.boxes = [np.array(h) for h in hboxes[:min(ew, 4)]]
        plist_p = [[float(w)] for w in range(len(sim.boxes))]
        ch.call_batch(plist_p)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            ch.call_batch(plist_p)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt_pageable = float(dt.item())
        n_pageable = len(sim.boxes)
        sim.boxes = hboxes
        # (3) the H2D line rate of this box: a plain pinned copy of the same bytes, nothing else
        dst = torch.empty((min(ew, 4),) + SHAPE, dtype=torch.float32, device=dev)
        srcs = [torch.from_numpy(h) for h in hboxes[:dst.shape[0]]]
        for i, sbuf in enumerate(srcs):
            dst[i].copy_(sbuf, non_blocking=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            for i, sbuf in enumerate(srcs):
                dst[i].copy_(sbuf, non_blocking=True)
        torch.cuda.synchronize()
        dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        box_bytes = int(np.prod(SHAPE)) * 4
        line_rate = 3 * len(srcs) * box_bytes / float(dtc.item()) / 1e9
        del dst
        h2d = int(ew * box_bytes)
        e2e = {"value": world * ew * esteps / dt_plugin, "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": int(world * ew * (8 + 8 * NCHUNKS * NBINS)),
               "walkers_per_step": ew * world, "steps": esteps,
               "through": "chain.call_batch -> Likelihood1DPowerLightcone.likelihood_batch -> ps21_lnl_host_ptrs "
                          "(contexts hold page-locked NumPy lightcones; cores, H2D, kernels, lnL + Delta^2 blobs D2H, "
                          "store)",
               "bound": "pcie", "h2d_gbs_per_gpu": ew * esteps * box_bytes / dt_plugin / 1e9,
               "h2d_line_rate_gbs_per_gpu": line_rate,
               "pageable": {"value": world * n_pageable * esteps / dt_pageable, "unit": UNIT,
                            "h2d_gbs_per_gpu": n_pageable * esteps * box_bytes / dt_pageable / 1e9,
                            "through": "the same plugin call on pageable NumPy lightcones (staged by the library)"},
               "c_abi": {"value": world * ew * esteps / dt_abi, "unit": UNIT,
                         "h2d_gbs_per_gpu": ew * esteps * box_bytes / dt_abi / 1e9,
                         "through": "ps21_lnl_host_ptrs on the same buffers"},
               "device_resident_over_e2e": value / (world * ew * esteps / dt_plugin)}
        tol = 1e-12 if plan.path == "fused" else 1e-7
        assert np.allclose(l_h, lnl_host[:ew], rtol=tol, atol=1e-12), "host-buffer path differs from device path"
        assert np.allclose(l_plugin, lnl_host[:ew], rtol=tol, atol=1e-12), "plugin path differs from device path"

    # ---- library comparator (context, never on the product path): cuFFT R2C of the same chunks, transform only
    comparator = None
    if rank == 0 and world == 1 and not args.no_comparator:
        nzc = SHAPE[2] // NCHUNKS
        nsp = min(16, nboxes * NCHUNKS)
        x = boxes[: max(1, nsp // NCHUNKS)].reshape(-1, SHAPE[0], SHAPE[1], NCHUNKS, nzc).permute(0, 3, 1, 2, 4)
        x = x.reshape(-1, SHAPE[0], SHAPE[1], nzc).contiguous()
        for _ in range(2):
            y = torch.fft.rfftn(x, dim=(1, 2, 3))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            y = torch.fft.rfftn(x, dim=(1, 2, 3))
        b.record()
        torch.cuda.synchronize()
        comparator = {"what": "torch.fft.rfftn (cuFFT) on the same chunks: R2C transform only, no |delta_k|^2, no "
                              "binning, no chi^2",
                      "us_per_spectrum": 1e3 * a.elapsed_time(b) / 5 / x.shape[0],
                      "this_library_us_per_spectrum": 1e3 * ms_per_step / (walkers * NCHUNKS * NREDSHIFT),
                      "note": "this library's figure includes power, binning and chi^2"}
        del x, y

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peak, peak_src = peaks()
    b_all, b_p1, b_last = alg_bytes_per_eval()
    kinds = {k: v for k, v in prof.items() if v[1] > 0}
    dom = max(kinds, key=lambda k: kinds[k][0])
    dom_ms, dom_n = kinds[dom]
    evals_per_launch = walkers * args.steps / dom_n
    alg = {"fft_pass1": b_p1, "fft_last_pass_bin": b_last, "fft_pass_mid": 2 * b_last}.get(dom, 0) * evals_per_launch
    achieved = alg / (dom_ms / dom_n * 1e-3) / 1e9 if dom_n else 0.0
    # DRAM traffic of that kernel: from the committed ncu --set full capture, scaled to this run's spectra per launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            trj = json.load(f)
        kname = {"fft_pass1": "k_fused_p1", "fft_last_pass_bin": "k_fused_p2"}.get(dom)
        if plan.path == "fused" and kname in trj["kernels"] and SHAPE[:2] == (256, 256):
            traffic = trj["kernels"][kname]["dram_bytes_per_spectrum"] * NCHUNKS * evals_per_launch
            traffic_src = "static: profiles/ncu_traffic.json (" + trj["source"] + "), not measured in this run"
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": alg, "avg_launch_ms": dom_ms / dom_n,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kinds.items()},
                "kernel_timing": "CUDA events on the launching stream around each kernel, direct launches, "
                                 f"{args.steps} steps right after the timed region ({ms_direct / args.steps:.3f} ms/step)",
                "pipeline": {"alg_bytes_per_eval": b_all,
                             "achieved": value / world * b_all / 1e9, "frac": value / world * b_all / 1e9 / peak,
                             "note": f"whole step, per GPU: evals/s x (4N + 16 N_h) x {NCHUNKS} chunks"}}
    if sustained:
        sustained["pipeline_frac"] = sustained["value"] / world * b_all / 1e9 / peak

    # ------------------------------------------------------------------ CPU baseline (bounded sample)
    cpu = None
    if not args.no_cpu and world == 1:
        from concurrent.futures import ProcessPoolExecutor

        workers = cpu_workers()
        with ProcessPoolExecutor(max_workers=workers) as pool:
            cpu_prepare(pool, workers)
            cpu_step(pool, workers, 0)   # warm-up
            v, dt = cpu_step(pool, workers, 1)
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": cpu_sample_string(workers) + f"; one step of {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 FFT / f64 bin accumulation", "data": "synthetic",
        "config": {"workload": workload_string(total_walkers),
                   "l2": f"inputs ({nboxes * int(np.prod(SHAPE)) * 4 / 2**30:.1f} GiB/GPU) exceed the 126 MB L2; "
                         "no flush needed",
                   "kernel_path": plan.path, "walkers_total": total_walkers, "walkers_per_gpu": walkers,
                   "parallelism": f"walkers sharded {total_walkers}/{world} per GPU",
                   "collective": "in-place all_gather of per-walker lnL (fp64) per step" if world > 1 else "none",
                   "cuda_graph": bool(use_graph)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "sustained": sustained,
        "weak": weak, "scaling_residual": residual, "comparator": comparator,
        "graph": graph_stats, "ms_per_step_direct_launches": ms_direct / args.steps,
        "gpu_launches": int(launches_per_step * args.steps), "self_check_lnl_walker0": self_check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers", type=int, default=None,
                    help="walkers of ONE step of the whole job, sharded over the GPUs (default 64; config 4: 2, its "
                         "inputs are 8 GiB per walker)")
    ap.add_argument("--path", default="auto", choices=["auto", "general", "fused"])
    ap.add_argument("--workspace-mb", type=int, default=0)
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration (3 = the headline metric)")
    ap.add_argument("--nbins", type=int, default=20,
                    help="n_psbins (20 = BASELINE; 116 = the reference default for 256^3 chunks)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="direct kernel launches instead of the CUDA graph of a step")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling record at N > 1")
    ap.add_argument("--no-comparator", action="store_true")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of the `sustained` record (0 = skip)")
    ap.add_argument("--e2e-walkers", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    select_config(args.config)
    if args.walkers is None:
        args.walkers = default_total_walkers(args.config)
    globals()["NBINS"] = args.nbins
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
