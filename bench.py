#!/usr/bin/env python
"""Benchmark of the power-spectrum-likelihood hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the CPU path, on the box's host cores

One step = one batched likelihood evaluation of WALKERS lightcones per GPU (256 x 256 x 1024 float32,
4 chunks, 20 log-k bins: BASELINE config 3) already resident in HBM.  Prints ONE JSON line on rank 0.
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
WORKLOAD = "BASELINE config 3: batched emcee step"

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
def _cpu_chunk_task(seed):
    import numpy as np

    from oracle import ps_oracle as O

    rng = np.random.default_rng(seed)
    nz = SHAPE[2] // NCHUNKS
    box = (rng.standard_normal((SHAPE[0], SHAPE[1], nz), dtype=np.float32) * SIGMA + MU)
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


def cpu_step(pool, workers, step):
    """One bounded sample: `workers` chunks (256^3 each = 1/4 eval) in parallel.  Returns evals/s."""
    t0 = time.perf_counter()
    list(pool.map(_cpu_chunk_task, [1000 * step + w for w in range(workers)]))
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from concurrent.futures import ProcessPoolExecutor

    workers = cpu_workers()
    with ProcessPoolExecutor(max_workers=workers) as pool:
        for w in range(args.warmup):
            cpu_step(pool, workers, 100000 + w)   # seeds must be non-negative
        t0 = time.perf_counter()
        for s in range(args.steps):
            cpu_step(pool, workers, s)
        total = time.perf_counter() - t0
    evals = args.steps * workers / (NCHUNKS * NREDSHIFT)
    value = evals / total
    nzc = SHAPE[2] // NCHUNKS
    sample = (f"{workers} chunks of {SHAPE[0]}x{SHAPE[1]}x{nzc} (= {workers / NCHUNKS:g} evals) per step, one process "
              "per core, oracle port of powerbox.get_power in float32-FFT mode (NumPy pocketfft)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 FFT / f64 bins",
        "data": "synthetic",
        "config": {"workload": f"{SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} lightcone, {NCHUNKS} chunks, 20 bins ({WORKLOAD})",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
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
    from py21cmmc_b200.plan import PowerPlan

    walkers = args.walkers
    bounds = [(i * SHAPE[2] // NCHUNKS, (i + 1) * SHAPE[2] // NCHUNKS) for i in range(NCHUNKS)]
    lengths = [(BOX_LEN, BOX_LEN, (b - a) * CELL) for a, b in bounds]
    cyl = MODE == "cylindrical"
    plan = PowerPlan(SHAPE, bounds, lengths, NBINS, not cyl, True, False, False, "legacy_1.001", MODE, local,
                     args.path, workspace_bytes=args.workspace_mb << 20)

    # synthetic Gaussian brightness-temperature boxes, generated on the device
    nboxes = walkers * NREDSHIFT  # coeval: [walker][redshift] cubes, each a one-chunk box
    boxes = torch.empty((nboxes,) + SHAPE, dtype=torch.float32, device=dev)
    for w in range(nboxes):
        g = torch.Generator(device=dev).manual_seed(1234 + 1000 * rank + w)
        boxes[w].normal_(MU, SIGMA, generator=g)
    # "data" = spectrum of walker 0 of rank 0 (the simulate=True convention), broadcast to every rank
    lnl = torch.empty((nboxes,), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * nboxes,), dtype=torch.float64, device=dev) if world > 1 else None
    if not cyl:
        power0 = plan.power_device(boxes[:NREDSHIFT].contiguous())
        kc = torch.from_numpy(np.concatenate(list(plan.kcentres))).to(dev)
        delta0 = power0 * kc**3 / (2 * np.pi**2)
        if world > 1:
            dist.broadcast(delta0, 0)
        data = delta0.cpu().numpy().reshape(NREDSHIFT, NCHUNKS, NBINS)
        kmask = np.repeat(((plan.kcentres >= MIN_K) & (plan.kcentres <= MAX_K)).astype(np.uint8)[None], NREDSHIFT, 0)
        plan.set_data(np.nan_to_num(data), None, kmask & np.isfinite(data), MODEL_UNC, True)

        def step():
            plan.lnl_device(boxes, lnl=lnl)
            if world > 1:
                dist.all_gather_into_tensor(gathered, lnl)
    else:  # config 5: the 2-D spectra are the product (no chi^2 is defined on them in the reference)
        p2d = torch.empty((nboxes, plan.out_per_box), dtype=torch.float64, device=dev)

        def step():
            plan.power_device(boxes, out=p2d)
            if world > 1:
                lnl.copy_(p2d[:, 0])
                dist.all_gather_into_tensor(gathered, lnl)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = plan.info()["launches_last_call"]
    plan.set_profiling(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall1 = time.time()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    prof = plan.profile()
    plan.set_profiling(False)
    ms_per_step = ms_total / args.steps
    value = world * walkers / (ms_per_step * 1e-3)

    # parity spot check inside the bench: the timed path's lnL of walker 0 is 0 against its own data
    lnl_host = lnl.cpu().numpy()
    self_check = float(lnl_host[0]) if (rank == 0 and not cyl) else None

    # ------------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not args.no_e2e and not cyl:
        ew = min(walkers, args.e2e_walkers)
        hbuf = torch.empty((ew * NREDSHIFT,) + SHAPE, dtype=torch.float32).pin_memory()
        hbuf.copy_(boxes[:ew * NREDSHIFT])
        hnp = hbuf.numpy()
        esteps = max(1, min(args.steps, args.e2e_steps))
        plan.lnl_host(hnp)  # warm-up (staging buffers)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(esteps):
            l_h, _, _ = plan.lnl_host(hnp)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": world * ew * esteps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(ew * NREDSHIFT * np.prod(SHAPE) * 4),
               "d2h_bytes_per_step": int(ew * NREDSHIFT * 8),
               "walkers_per_step": ew, "steps": esteps,
               "note": "ps21_lnl_host: pinned host lightcones -> H2D (double buffered) -> kernels -> lnL D2H"}
        # fused path: fixed-order reductions (bitwise equal); general path: fp64 shared-memory atomics inside a
        # CTA reorder the last bits, which the chi^2 amplifies by Delta/(Delta - data)
        tol = 1e-12 if plan.path == "fused" else 1e-7
        assert np.allclose(l_h, lnl_host[:ew * NREDSHIFT], rtol=tol, atol=1e-12), \
            "host-buffer path differs from device path"

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
    # DRAM traffic of that kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), scaled
    # from the profiled launch to this run's spectra per launch
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f)["kernels"]
        kname = {"fft_pass1": "k_fused_p1", "fft_last_pass_bin": "k_fused_p2"}.get(dom)
        if plan.path == "fused" and kname in tr and SHAPE[:2] == (256, 256):
            traffic = tr[kname]["dram_bytes_per_spectrum"] * NCHUNKS * evals_per_launch
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg, "avg_launch_ms": dom_ms / dom_n,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kinds.items()},
                "pipeline": {"alg_bytes_per_eval": b_all,
                             "achieved": value / world * b_all / 1e9, "frac": value / world * b_all / 1e9 / peak,
                             "note": f"whole step, per GPU: evals/s x (4N + 16 N_h) x {NCHUNKS} chunks"}}

    # ------------------------------------------------------------------ CPU baseline (bounded sample)
    cpu = None
    if not args.no_cpu and world >= 1:
        from concurrent.futures import ProcessPoolExecutor

        workers = cpu_workers()
        with ProcessPoolExecutor(max_workers=workers) as pool:
            v, dt = cpu_step(pool, workers, 0)
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": f"{workers} chunks of {SHAPE[0]}x{SHAPE[1]}x{SHAPE[2] // NCHUNKS} (= {workers / NCHUNKS:g} "
                         f"evals) in {dt:.1f} s, one process per core, oracle port of powerbox.get_power "
                         "(float32 FFT)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 FFT / f64 bin accumulation", "data": "synthetic",
        "config": {"workload": f"{walkers} walkers/GPU x "
                               + (f"{NREDSHIFT} coeval cubes " if NREDSHIFT > 1 else "")
                               + f"{SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} float32, "
                               f"{NCHUNKS} chunk(s), {NBINS} {'k_perp x k_par' if cyl else 'log-k'} bins ({WORKLOAD})",
                   "l2": f"inputs ({nboxes * int(np.prod(SHAPE)) * 4 / 2**30:.1f} GiB/GPU) exceed the 126 MB L2; "
                         "no flush needed",
                   "kernel_path": plan.path, "walkers_per_gpu": walkers, "parallelism": f"walker-sharded x{world}",
                   "collective": "all_gather of per-walker lnL (fp64) per step" if world > 1 else "none"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches_per_step * args.steps), "self_check_lnl_walker0": self_check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers", type=int, default=None,
                    help="walkers per GPU per step (default 64; config 4: 2, its inputs are 8 GiB per walker)")
    ap.add_argument("--path", default="auto", choices=["auto", "general", "fused"])
    ap.add_argument("--workspace-mb", type=int, default=0)
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json lightcone configuration (3 = the headline metric)")
    ap.add_argument("--nbins", type=int, default=20,
                    help="n_psbins (20 = BASELINE; 116 = the reference default for 256^3 chunks)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-walkers", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    select_config(args.config)
    if args.walkers is None:
        args.walkers = 2 if args.config == 4 else 64
    globals()["NBINS"] = args.nbins
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
