for v in A B; do
for env in "X=1" "PS21_G3_TILE=8" "PS21_G2_TILE=16 PS21_G3_TILE=8" "PS21_G1_TILE=16"; do
  env $env PS21_LIBRARY=$PWD/ab/libps21_$v.so python bench.py --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/abg.log 2>&1
  python - "$v $env" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/abg.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print("SW", sys.argv[1], round(d["value"],1) if d else "failed", {k: round(x,2) for k,x in d["roofline"]["kernel_ms_per_step"].items()} if d else open("gpurun_out/abg.log").read()[-300:])
PY
done
done
