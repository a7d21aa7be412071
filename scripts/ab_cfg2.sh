for rep in 1 2; do
for v in A B; do
  PS21_LIBRARY=$PWD/ab/libps21_$v.so python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/abg.log 2>&1
  python - "$v" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/abg.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print("SW", sys.argv[1], round(d["value"],1) if d else "failed", {k: round(x,2) for k,x in d["roofline"]["kernel_ms_per_step"].items()} if d else open("gpurun_out/abg.log").read()[-300:])
PY
done
done
