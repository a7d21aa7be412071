python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --path general --walkers 16 > gpurun_out/bench_general_v4.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ["bench_general_v4"]:
    l=[x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")]
    if not l:
        print(f, "no json"); print(open(f"gpurun_out/{f}.log").read()[-2000:]); continue
    d=json.loads(l[-1]); print(f, round(d["value"],1), "evals/s", round(d["ms_per_step"],3), "ms pipeline_frac", round(d["roofline"]["pipeline"]["frac"],3), d["roofline"]["kernel_ms_per_step"])
PY
