"""bench.py contract checks that need no GPU: the reference arm's JSON line, the clock sampler's parsing, the
roofline byte accounting."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench

    bench.select_config(3)
    b_all, b_p1, b_last = bench.alg_bytes_per_eval()
    assert b_all == 809_500_672  # 809.5 MB per config-3 evaluation
    assert b_p1 + b_last == b_all
    bench.select_config(1)
    assert abs(bench.alg_bytes_per_eval()[0] / 1e6 - 202.4) < 0.1
    bench.select_config(3)


def test_clock_sampler_parses_and_windows(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    fake = tmp_path / "nvidia-smi"
    fake.write_text(
        "#!/usr/bin/env python3\n"
        "import time, datetime, sys\n"
        "for i in range(200):\n"
        "    t = datetime.datetime.now().strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]\n"
        "    mhz = 1965 if i % 2 else 1900\n"
        "    print(f'{t}, {mhz}, 1965, 900.5, Not Active, Not Active, Not Active, ' + ('Active' if i > 5 else 'Not Active'), flush=True)\n"
        "    time.sleep(0.02)\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.4)
    t0 = time.time()
    time.sleep(0.5)
    t1 = time.time()
    out = s.stop(t0, t1)
    assert out is not None and out["window"] == "timed region" and out["samples"] >= 5
    assert out["sm_max_mhz"] == 1965.0 and 1900.0 <= out["sm_mhz"] <= 1965.0
    assert out["reasons"] == ["sw_power_cap"]
