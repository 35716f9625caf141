"""Lean profiling target: PROFILE_SITES synthetic sites, PROFILE_DAYS one-day launches of the soil kernel (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microclimate_jl_b200 import marshal as M, synthetic as syn  # noqa: E402
from microclimate_jl_b200.simulation import SoilEngine  # noqa: E402

S = int(os.environ.get("PROFILE_SITES", "65536"))
nd = int(os.environ.get("PROFILE_DAYS", "2"))
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
sites = syn.make_sites(0, S, S)
eng = SoilEngine(desc, S)
eng.set_sites(sites)
for d in range(nd):
    f = syn.make_forcing(0, S, S, d, 1)
    out = M.OutputBatch(1, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
    rc = eng.step_days(d, f, out)
    c = eng.counters()
    print(f"day {d}: rc {rc} kernel {c['gpu_ms']:.1f} ms -> {S / (c['gpu_ms'] * 1e-3):.0f} site-days/s", flush=True)
