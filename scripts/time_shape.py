"""Time the device path on an arbitrary lightcone shape (context numbers for shapes outside BASELINE.json).
usage: python scripts/time_shape.py n0 n_slices nchunks [walkers] [nbins]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import py21cmmc_b200  # noqa: E402,F401
from py21cmmc_b200 import binning  # noqa: E402
from py21cmmc_b200.plan import get_plan  # noqa: E402

n0, ns, nchunks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
walkers = int(sys.argv[4]) if len(sys.argv) > 4 else 16
nbins = int(sys.argv[5]) if len(sys.argv) > 5 else 20
cell = 1.5
idx = binning.chunk_indices(ns, nchunks)
bounds = [(idx[i], idx[i + 1]) for i in range(nchunks)]
lengths = [(n0 * cell, n0 * cell, (b - a) * cell) for a, b in bounds]
plan = get_plan((n0, n0, ns), bounds, lengths, nbins, True)
torch.manual_seed(1)
x = torch.randn((walkers, n0, n0, ns), device="cuda:0") * 10 + 20
plan.set_profiling(True)
for _ in range(3):
    out = plan.power_device(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 5
for _ in range(reps):
    out = plan.power_device(x)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"{n0}x{n0}x{ns}, {nchunks} chunks {[e - s for s, e in bounds]}, path {plan.path}: {ms:.3f} ms per {walkers} walkers = "
      f"{1e3 * walkers / ms:.1f} evals/s; kernels {plan.profile()}; checksum {float(torch.nan_to_num(out.double()).sum()):.10e}")
