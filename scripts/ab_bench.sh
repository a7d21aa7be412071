# A/B of two library builds on the same box: ab/libps21_A.so vs ab/libps21_B.so, interleaved
# usage: bash scripts/ab_bench.sh [extra bench.py args]
for rep in 1 2 3; do
for v in A B; do
  PS21_LIBRARY=$PWD/ab/libps21_$v.so python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ab_${v}_$rep.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/ab_${v}_$rep.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$v rep $rep", round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])
else:
    print("$v failed"); print(open("gpurun_out/ab_${v}_$rep.log").read()[-500:])
PY
done
done
