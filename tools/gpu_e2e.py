"""Development probe: wall time of mc_step_days (pinned host buffers) per day for several MC_SLICES settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from microclimate_jl_b200 import marshal as M, synthetic as syn, _abi as A
from microclimate_jl_b200.simulation import SoilEngine
import bench

S = int(os.environ.get("PROFILE_SITES", str(1 << 20))); nd = 3
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
sites = syn.make_sites(0, S, S)
days = [bench._pin_forcing(syn.make_forcing(0, S, S, d, 1)) for d in range(nd)]
out = M.OutputBatch(1, desc.n_soil, S, ())
for f in bench.FIELDS:
    setattr(out, f, bench._pinned((24, desc.n_soil, S)))
for slices in os.environ.get("SLICE_LIST", "1 4").split():
    os.environ["MC_SLICES"] = slices
    eng = SoilEngine(desc, S); eng.set_sites(sites)
    for d in range(nd):
        t = time.perf_counter(); eng.step_days(d, days[d], out); w = time.perf_counter() - t
        print(f"slices={slices} day {d}: wall {w:.3f} s  pipeline(ev0->ev1) {eng.counters()['gpu_ms']/1e3:.3f} s -> {S/w:.0f} site-days/s", flush=True)
    eng.close()
