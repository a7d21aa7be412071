# A/B of the general path on one box: config-3 workload (16 walkers), config 2 (200^3), config 4 (512^3), 116 bins
for args in "--path general --walkers 16" "--config 1 --path general" "--config 2" "--config 4 --e2e-walkers 1" "--path general --walkers 16 --nbins 116" "--config 5 --path general --walkers 8"; do
for v in A B; do
  PS21_LIBRARY=$PWD/ab/libps21_$v.so python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e $args > gpurun_out/abg.log 2>&1
  python - "$v" "$args" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/abg.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print(sys.argv[1], sys.argv[2], "|", round(d["value"],1), round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
else:
    print(sys.argv[1], sys.argv[2], "failed"); print(open("gpurun_out/abg.log").read()[-800:])
PY
done
done
