# informational bench lines for the other configurations of BASELINE.json (parity-test cases, not the bench line)
python bench.py --config 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg1.log 2>&1
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.log 2>&1
python bench.py --config 4 --walkers 2 --steps 5 --warmup 3 --no-cpu --e2e-walkers 1 > gpurun_out/bench_cfg4.log 2>&1
python bench.py --config 5 --walkers 16 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5.log 2>&1
python - <<'PY'
import json
for f in ["bench_cfg1","bench_cfg2","bench_cfg4","bench_cfg5"]:
    l=[x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")]
    if not l:
        print(f, "no json"); print(open(f"gpurun_out/{f}.log").read()[-1500:]); continue
    d=json.loads(l[-1]); print(f, d["metric"], round(d["value"],1), round(d["ms_per_step"],3), "ms pipeline_frac", round(d["roofline"]["pipeline"]["frac"],3), d["config"]["kernel_path"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["e2e"] and round(d["e2e"]["value"],1), d.get("self_check_lnl_walker0"))
PY
