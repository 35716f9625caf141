"""Development probe run on the GPU box: parity statistics vs the oracle, kernel timing, FP64 peak.  Not part of the product."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import configs  # noqa: E402
from microclimate_jl_b200 import marshal as M, synthetic as syn  # noqa: E402
from microclimate_jl_b200.simulation import SoilEngine  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402


def stats(a, b, tol):
    d = np.abs(a - b).ravel()
    return dict(max=float(d.max()), p999=float(np.quantile(d, 0.999)), median=float(np.median(d)), frac_gt_tol=float((d > tol).mean()))


def main():
    S = int(os.environ.get("PROBE_SITES", "32768"))
    res = {}
    for name, mk in (("config1", configs.config1_monthly), ("config2", configs.config2_daily)):
        p, _ = mk()
        o1, e1 = configs.run_engine(OracleEngine, p)
        o2, e2 = configs.run_engine(SoilEngine, p)
        res[name] = dict(T=stats(o1.soil_temperature, o2.soil_temperature, 1e-3), theta=stats(o1.soil_moisture, o2.soil_moisture, 1e-5),
                         counters_oracle=e1.counters(), counters_gpu=e2.counters())
        print(name, json.dumps(res[name]), flush=True)
    # synthetic parity, 256 sites x 2 days
    model = syn.synthetic_model()
    desc = M.model_desc(model, syn.synthetic_time_mode())
    Sp = 256
    sites, f = syn.make_sites(0, Sp, Sp), syn.make_forcing(0, Sp, Sp, 0, 2)
    outs = []
    for E in (OracleEngine, SoilEngine):
        out = M.OutputBatch(2, desc.n_soil, Sp, ("soil_temperature", "soil_moisture"))
        e = E(desc, Sp)
        e.set_sites(sites)
        e.step_days(0, f, out)
        outs.append(out)
    res["synthetic256"] = dict(T=stats(outs[0].soil_temperature, outs[1].soil_temperature, 1e-3),
                               theta=stats(outs[0].soil_moisture, outs[1].soil_moisture, 1e-5))
    print("synthetic256", json.dumps(res["synthetic256"]), flush=True)
    # timing
    t = time.time()
    sites = syn.make_sites(0, S, S)
    nd = 4
    f = syn.make_forcing(0, S, S, 0, nd)
    print("generated", S, "sites x", nd, "days in", round(time.time() - t, 2), "s", flush=True)
    eng = SoilEngine(desc, S)
    print("fp64 peak TFLOP/s", eng.measure_fp64_peak(), flush=True)
    eng.set_sites(sites)
    for d in range(nd):
        out = M.OutputBatch(1, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
        t = time.time()
        rc = eng.step_days(d, f.day_slice(d, d + 1), out)
        wall = time.time() - t
        c = eng.counters()
        print(f"day {d}: rc {rc} wall {wall:.3f}s kernel {c['gpu_ms']:.1f} ms -> {S / (c['gpu_ms'] * 1e-3):.0f} site-days/s (kernel), {S / wall:.0f} e2e; "
              f"rhs/site-day {c['n_rhs'] / c['n_site_days']:.0f} newton/site-day {c['n_newton_iter'] / c['n_site_days']:.0f}", flush=True)
    res["timing"] = dict(S=S, gpu_ms=c["gpu_ms"], site_days_per_s=S / (c["gpu_ms"] * 1e-3))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
