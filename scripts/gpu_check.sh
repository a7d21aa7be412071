python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_v21.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ["bench_v21"]:
    l=[x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")]
    if not l:
        print(f, "no json"); print(open(f"gpurun_out/{f}.log").read()[-2000:]); continue
    d=json.loads(l[-1]); print(f, round(d["value"],1), "evals/s", round(d["ms_per_step"],3), "ms pipeline_frac", round(d["roofline"]["pipeline"]["frac"],3), d["roofline"]["kernel_ms_per_step"], d["e2e"] and round(d["e2e"]["value"],1), d["clocks"])
PY
