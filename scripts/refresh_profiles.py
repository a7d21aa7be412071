"""Regenerate the tracked ncu evidence under profiles/ from one gpurun call's scratch output.

usage: python scripts/refresh_profiles.py <rep.ncu-rep> <launches.csv> <launches_default.csv> <plain_default.log> [round]
Writes profiles/<round>_ncu_fused_final_summary.txt, profiles/ncu_traffic.json,
profiles/<round>_ncu_launches_{fused_final.csv,default_bench.csv,default_bench_summary.txt}  (round defaults to r02;
the .ncu-rep binaries stay in gpurun_out/, they are not tracked).
"""
import csv
import json
import shutil
import subprocess
import sys
from collections import OrderedDict

rep, launches, launches_default, plain_default = sys.argv[1:5]
RND = sys.argv[5] if len(sys.argv) > 5 else "r02"
P = "profiles/" + RND + "_"

shutil.copy(launches, P + "ncu_launches_fused_final.csv")
shutil.copy(launches_default, P + "ncu_launches_default_bench.csv")

summary = subprocess.run([sys.executable, "scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout
with open(P + "ncu_fused_final_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:k_fused_p -s 6 -c 2  "
            "python bench.py --steps 1 --warmup 3 --walkers 8 --no-e2e --no-cpu\n")
    f.write(summary)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]


def to_bytes(val, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(val.replace(",", "")) * scale


kernels = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    key = "k_fused_p1" if "k_fused_p1" in name else "k_fused_p2" if "k_fused_p2" in name else None
    if key is None:
        continue
    rd = to_bytes(r[h.index("dram__bytes_read.sum")], rows[1][h.index("dram__bytes_read.sum")])
    wr = to_bytes(r[h.index("dram__bytes_write.sum")], rows[1][h.index("dram__bytes_write.sum")])
    nspec = 8 * 4   # --walkers 8, four chunks per lightcone
    kernels[key] = {"spectra_per_profiled_launch": nspec, "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "dram_bytes_per_spectrum": (rd + wr) / nspec}
with open("profiles/ncu_traffic.json", "w") as f:
    json.dump({"source": RND + " capture: ncu --set full of bench.py --steps 1 --warmup 3 --walkers 8 --no-e2e --no-cpu "
                         "(summary: " + P + "ncu_fused_final_summary.txt)",
               "kernels": kernels}, f, indent=1)

# launch list of the default bench command
rows = [r for r in csv.reader(open(launches_default, errors="replace")) if len(r) > 5]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[h.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += t
step = ("k_fused_p1", "k_fused_p2", "k_finalize", "k_slab_mean")
tot = sum(v[1] for k, v in agg.items() if any(s in k for s in step))
line = [l for l in open(plain_default) if l.startswith("{")][-1]
rec = json.loads(line)
br = rec.get("roofline", {}).get("kernel_ms_per_step", {})
with open(P + "ncu_launches_default_bench_summary.txt", "w") as f:
    f.write("===== launch list of the default bench command: ncu --metrics gpu__time_duration.sum --clock-control none "
            "python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --sustain-s 0 --no-comparator\n")
    f.write("      (64 walkers; per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes)\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f" share_of_step_kernels={t / tot:.3f}" if any(s in k for s in step) else ""
        f.write(f"  {k[:70]:70s} n={n:3d} total={t:9.3f} ms avg={1e3 * t / n:9.1f} us{share}\n")
    if br:
        s = sum(br.values())
        f.write("      live CUDA-event shares of the same command (no profiler): "
                + ", ".join(f"{k}={v / s:.3f}" for k, v in br.items()) + f"; ms_per_step={rec['ms_per_step']:.3f}\n")
print(open(P + "ncu_fused_final_summary.txt").read())
print(open(P + "ncu_launches_default_bench_summary.txt").read())
print(open("profiles/ncu_traffic.json").read())
