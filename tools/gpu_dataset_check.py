"""GPU check of the file-driven path: tools/dataset.py make + run (CUDA engine, memory-mapped inputs) must equal the in-memory run bit for bit."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import microclimate_jl_b200  # noqa
from microclimate_jl_b200 import marshal as M, storage as st, synthetic as syn
from microclimate_jl_b200.simulation import SoilEngine
S, ND = 4096, 2
d = tempfile.mkdtemp()
subprocess.check_call([sys.executable, "tools/dataset.py", "make", d, "--sites", str(S), "--days", str(ND), "--chunk", "1"], cwd=ROOT)
subprocess.check_call([sys.executable, "tools/dataset.py", "run", d], cwd=ROOT)
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
eng = SoilEngine(desc, S); eng.set_sites(syn.make_sites(0, S, S))
out = M.OutputBatch(ND, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
eng.step_days(0, syn.make_forcing(0, S, S, 0, ND), out)
got = np.concatenate([st.load_outputs(os.path.join(d, f"out_d{k:04d}_s0-{S}.mcb")).soil_temperature for k in range(ND)])
print("file-driven == in-memory:", np.array_equal(got, out.soil_temperature), "max|d| =", float(np.abs(got - out.soil_temperature).max()))
