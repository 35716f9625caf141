"""Headroom probe: soil_conductivity over independent items vs its rate inside the per-site kernel."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from microclimate_jl_b200 import marshal as M, synthetic as syn
from microclimate_jl_b200.simulation import SoilEngine
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
eng = SoilEngine(desc, 64)
for n, reps in ((1 << 20, 64), (1 << 22, 32)):
    v = C.c_double()
    rc = eng._L.mc_bench_conductivity(eng.h, n, reps, C.byref(v))
    print(f"items {n} reps {reps}: rc {rc} {v.value / 1e9:.2f} G evaluations/s", flush=True)
print("fp64 peak", eng.measure_fp64_peak())
