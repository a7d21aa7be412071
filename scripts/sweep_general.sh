# tile-size sweep of the general path (tuning experiment)
for args in "--path general --walkers 16" "--config 2" "--config 4 --e2e-walkers 1"; do
for t2 in 4 8 16; do
for t3 in 4 8 16; do
  PS21_G2_TILE=$t2 PS21_G3_TILE=$t3 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $args > gpurun_out/abg.log 2>&1
  python - "$t2 $t3" "$args" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/abg.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("SW", sys.argv[1], sys.argv[2], "|", round(d["value"],1), round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["kernel_ms_per_step"].items()})
else:
    print("SW", sys.argv[1], sys.argv[2], "failed"); print(open("gpurun_out/abg.log").read()[-300:])
PY
done
done
done
