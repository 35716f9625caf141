"""Torch-free quick GPU check for a short box slot.
  python tools/gpu_quick_ab.py parity            512 synthetic sites x 2 days, default library vs the oracle (bit-for-bit expected)
  python tools/gpu_quick_ab.py perf [SITES]      kernel time of 2 one-day launches (device counters) through mc_step_days, library from MC_LIB
Each mode appends one line to gpurun_out/quick_ab.log."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from microclimate_jl_b200 import marshal as M, synthetic as syn  # noqa: E402
from microclimate_jl_b200.simulation import SoilEngine  # noqa: E402

LOG = os.path.join(ROOT, "gpurun_out", "quick_ab.log")
os.makedirs(os.path.dirname(LOG), exist_ok=True)


def say(msg):
    with open(LOG, "a") as f:
        f.write(msg + "\n")
    print(msg, flush=True)


mode = sys.argv[1]
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
tag = os.path.basename(os.environ.get("MC_LIB", "default")) + (" pad=" + os.environ["MC_COOP_SMEM_PAD"] if os.environ.get("MC_COOP_SMEM_PAD") else "")
if mode == "parity":
    from oracle.oracle import OracleEngine
    S, D = 512, 2
    sites, forcing = syn.make_sites(0, S, S), syn.make_forcing(0, S, S, 0, D, threads=8)
    outs = []
    for E in (OracleEngine, SoilEngine):
        out = M.OutputBatch(D, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
        e = E(desc, S)
        e.set_sites(sites)
        e.step_days(0, forcing, out)
        outs.append(out)
    dT = float(np.abs(outs[0].soil_temperature - outs[1].soil_temperature).max())
    dth = float(np.abs(outs[0].soil_moisture - outs[1].soil_moisture).max())
    say(f"parity {tag}: {S} sites x {D} days max|dT|={dT:.3g} max|dtheta|={dth:.3g} {'BIT-IDENTICAL' if dT == 0 and dth == 0 else 'MISMATCH'}")
else:
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 151552
    t0 = time.time()
    sites = syn.make_sites(0, S, S)
    eng = SoilEngine(desc, S)
    eng.set_sites(sites)
    ms = []
    for d in range(2):
        f = syn.make_forcing(0, S, S, d, 1, threads=8)
        out = M.OutputBatch(1, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
        eng.step_days(d, f, out)
        ms.append(eng.counters()["gpu_ms"])
    say(f"perf {tag}: {S} sites kernel ms/day {ms[0]:.1f} {ms[1]:.1f} -> {S / (ms[1] * 1e-3):.0f} site-days/s (checksum {float(out.soil_temperature[-1, 0, :64].sum()):.9f}, wall {time.time() - t0:.1f} s)")
