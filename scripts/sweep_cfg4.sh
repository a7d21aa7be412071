# tile sweep for 512-point axes (config 4)
for t1 in 4 8 16; do
  PS21_G1_TILE=$t1 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/abg.log 2>&1
  python - "g1=$t1" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/abg.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print("SW", sys.argv[1], round(d["value"],1) if d else "failed", {k: round(x,2) for k,x in d["roofline"]["kernel_ms_per_step"].items()} if d else open("gpurun_out/abg.log").read()[-300:])
PY
done
for t2 in 4 8 16; do
for t3 in 4 8; do
  PS21_G2_TILE=$t2 PS21_G3_TILE=$t3 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/abg.log 2>&1
  python - "g2=$t2 g3=$t3" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/abg.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print("SW", sys.argv[1], round(d["value"],1) if d else "failed", {k: round(x,2) for k,x in d["roofline"]["kernel_ms_per_step"].items()} if d else open("gpurun_out/abg.log").read()[-300:])
PY
done
done
