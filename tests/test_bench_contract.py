"""bench.py's reference arm runs on CPU (the oracle on all host cores): check the one-JSON-line contract the driver parses, and that the
product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--sites", "64", "--cpu-sites", "64", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                                  # nothing but the JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "site-days/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "site-days/sec (hourly soil T+moisture)" and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "site-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_only_rank_zero_works_under_torchrun_env():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--sites", "64", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, "bench.py", "--sites", "64", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU path" in (p.stderr + p.stdout)


def test_weather_probe_self_test_and_bench_attachment():
    """tools/gpu_weather_probe.py (what bench.py attaches as `widening.weather_kernel` at N = 1) run against the host emulation of mc_weather.cu: the script's
    struct filling, launch loop, byte accounting and oracle check are exercised here; on a box without a GPU bench.weather_probe() degrades to an error object."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_weather_probe.py"), "--emulate", "--sites", "1500", "--days", "2", "--reps", "1", "--check", "300"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-800:]
    p = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert p["parity"] == {"sites_checked": 300, "bit_identical_to_oracle": True}
    assert p["roofline"]["bound"] == "hbm" and abs(p["roofline"]["alg_bytes_per_site_day"] - 2232.0) < 1.0 and p["site_days_per_s"] > 0
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_sand_probe.py"), "--emulate", "--sites", "12", "--steps", "1"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-800:]
    q = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert set(q) >= {"with_sand", "without_sand", "with_snow", "ratio", "snow_ratio", "workload"} and q["with_sand"]["work_per_site_day"]["n_infil_steps"] == 240.0
    assert "soil_chunk_kernel<19, 8>" in q["with_snow"]["kernel"] and q["with_snow"]["sites_with_snow_at_the_end"] >= 0
    assert 1.0 <= q["with_tolerance_convergence"]["day_passes_per_site_day"] <= 6.0 and q["tolerance_ratio"] > 0
    sys.path.insert(0, root)
    import bench
    import torch
    if not torch.cuda.is_available():
        assert "error" in bench.weather_probe(120)
