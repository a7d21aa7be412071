python bench.py --config 4 --walkers 2 --steps 5 --warmup 3 --no-cpu --e2e-walkers 1 > gpurun_out/bench_cfg4.log 2>&1
python - <<'PY'
import json
for f in ["bench_cfg4"]:
    l=[x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")]
    if not l:
        print(f, "no json"); print(open(f"gpurun_out/{f}.log").read()[-1500:]); continue
    d=json.loads(l[-1]); print(f, d["metric"], round(d["value"],2), round(d["ms_per_step"],3), "ms pipeline_frac", round(d["roofline"]["pipeline"]["frac"],3), d["config"]["kernel_path"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["e2e"] and round(d["e2e"]["value"],2), d.get("self_check_lnl_walker0"))
PY
