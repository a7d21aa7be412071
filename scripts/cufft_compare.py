"""Context number, not a product path: cuFFT (torch.fft.rfftn) on the benchmark's chunk shape, R2C only
(no |delta_k|^2, no binning), timed with CUDA events.  usage: python scripts/cufft_compare.py [nspectra]"""
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
for shape in [(256, 256, 256), (128, 128, 256), (200, 200, 200), (512, 512, 512)]:
    nb = n if shape[0] < 512 else max(1, n // 8)
    x = torch.randn((nb,) + shape, device=dev, dtype=torch.float32)
    for _ in range(3):
        y = torch.fft.rfftn(x, dim=(1, 2, 3))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        y = torch.fft.rfftn(x, dim=(1, 2, 3))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    N = shape[0] * shape[1] * shape[2]
    Nh = shape[0] * shape[1] * (shape[2] // 2 + 1)
    print(f"cuFFT rfftn {shape} x{nb}: {1e3 * ms / nb:8.1f} us/spectrum, "
          f"{nb * (4 * N + 8 * Nh) / ms / 1e6:7.1f} GB/s of 4N+8N_h (compulsory) bytes")
    del x, y
    torch.cuda.empty_cache()
