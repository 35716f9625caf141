"""Lean profiling / A-B target: PROFILE_SITES synthetic sites, PROFILE_DAYS one-day launches of the soil kernel with the forcing
resident in HBM (kernel-only time from the engine's CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from microclimate_jl_b200 import _abi as A, marshal as M, synthetic as syn  # noqa: E402
from microclimate_jl_b200.simulation import SoilEngine  # noqa: E402

S = int(os.environ.get("PROFILE_SITES", "65536"))
nd = int(os.environ.get("PROFILE_DAYS", "2"))
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
sites = syn.make_sites(0, S, S)
eng = SoilEngine(desc, S)
eng.set_sites(sites)
out = {f: torch.zeros((24, desc.n_soil, S), dtype=torch.float64, device="cuda") for f in ("soil_temperature", "soil_moisture")}
for d in range(nd):
    f = syn.make_forcing(0, S, S, d, 1)
    dev = {k: torch.from_numpy(getattr(f, k)).cuda() for k in A.FORCING_HOURLY + A.FORCING_DAILY if getattr(f, k) is not None}
    torch.cuda.synchronize()
    eng.step_days_device(d, 1, {k: v.data_ptr() for k, v in dev.items()}, {k: v.data_ptr() for k, v in out.items()})
    eng.synchronize()
    c = eng.counters()
    print(f"day {d}: flagged {int((eng.status() != 0).sum())} kernel {c['gpu_ms']:.1f} ms -> {S / (c['gpu_ms'] * 1e-3):.0f} site-days/s", flush=True)
