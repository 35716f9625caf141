"""Kernel throughput of two harder variants of the benchmark grid next to the headline grid, at one full wave (151 552 sites), on one B200: the pure-sand
texture class (SURVEY 8d's six classes) and the SnowModel (8 snow nodes; the per-site kernel).  Prints ONE JSON object.

    python tools/gpu_sand_probe.py [--sites 151552] [--steps 3]

The headline grid of bench.py replaces "sand" by loamy sand because pure sand drives the infiltration Newton loop to its 500-iteration cap; both
arms of the benchmark run the same grid, so the ratio is fair, but the absolute number is on the easier grid.  This probe integrates the same sites, days and forcing
three times — headline textures, textures with sand, headline textures with the SnowModel — through mc_step_days (host buffers), one warm-up day then `steps` timed days, kernel time from the CUDA events of the
library (mc_counters.gpu_ms).  bench.py runs it in a subprocess at N = 1 and attaches the object as `grid_variants`.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STATUS_BITS = {1: "NAN", 2: "DT_UNDERFLOW", 4: "NEWTON_MAXITER", 8: "STEP_LIMIT"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=151552)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--emulate", action="store_true", help="CPU-tier self-test of this script on the host emulation of the kernel (tests/hostemu); no timing claim")
    args = ap.parse_args()
    import contextlib
    import time
    from microclimate_jl_b200 import marshal as M, synthetic as syn
    if args.emulate:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from hostemu_engine import HostEmuEngine
        SoilEngine = lambda d, n: contextlib.closing(HostEmuEngine(d, n))      # noqa: E731
    else:
        from microclimate_jl_b200.simulation import SoilEngine
    S, K = args.sites, args.steps
    import dataclasses
    import numpy as np
    from microclimate_jl_b200 import types as T
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    # third variant: the same grid with the SnowModel (8 snow nodes on the soil column, density function of the reference's snow fixtures) — it runs the per-site
    # kernel soil_chunk_kernel<19, 8>; the year starts in January, so the high-latitude sites of the northern sweep carry snow from the first cold wet hour on
    snow_model = dataclasses.replace(syn.synthetic_model(), snow_model=T.SnowModel(density_function=(0.5979, 0.2178, 0.001, 0.0038)))
    desc_snow = M.model_desc(snow_model, syn.synthetic_time_mode())
    # fourth variant: tolerance-based day-pass convergence (SoilTemperatureConvergenceTolerance, days repeated until the start-of-day profile moves < 0.5 K, at most 6
    # passes, NonConsecutiveDayMode) — per-site pass counts diverge, so it runs the per-site kernel soil_chunk_kernel<19, 0> as well
    base = syn.synthetic_model()
    tol_model = dataclasses.replace(base, config=dataclasses.replace(base.config, convergence=T.SoilTemperatureConvergenceTolerance(tolerance=0.5, max_iterations_per_day=6)))
    desc_tol = M.model_desc(tol_model, T.NonConsecutiveDayMode(iterations_per_day=6))
    days = [syn.make_forcing(0, S, S, d, 1, threads=os.cpu_count() or 1) for d in range(K + 1)]
    out = M.OutputBatch(1, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
    res = {}
    for key, sand, dsc in (("without_sand", False, desc), ("with_sand", True, desc), ("with_snow", False, desc_snow), ("with_tolerance_convergence", False, desc_tol)):
        if key == "with_tolerance_convergence":
            out = M.OutputBatch(1, dsc.n_soil, S, ("soil_temperature", "soil_moisture"))
        if key == "with_snow":
            out = M.OutputBatch(1, dsc.n_soil, S, ("soil_temperature", "soil_moisture", "snow_depth"), dsc.n_snow)
        with SoilEngine(dsc, S) as eng:
            eng.set_sites(syn.make_sites(0, S, S, with_sand=sand))
            eng.step_days(0, days[0], out)                      # warm-up day (also the spin-up from the uniform initial profile)
            c0 = eng.counters()                                  # counters of the warm-up day are subtracted (mc_reset_counters is a product-only entry point)
            ms = []
            for d in range(1, K + 1):
                t0 = time.perf_counter()
                eng.step_days(d, days[d], out)
                ms.append(eng.counters().get("gpu_ms") or 1e3 * (time.perf_counter() - t0))      # the emulation has no device clock
            c, st = eng.counters(), (out.status if args.emulate else eng.status())
        res[key] = {
            "site_days_per_s": S * K / (sum(ms) * 1e-3), "kernel_ms_per_day": ms,
            "flagged_by_bit": {name: int(((st & bit) != 0).sum()) for bit, name in STATUS_BITS.items()},
            "work_per_site_day": {k: (c[k] - c0[k]) / (S * K) for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps")}}
        if key == "with_tolerance_convergence":
            res[key]["kernel"] = "soil_chunk_kernel<19, 0> (per-site kernel: per-site pass counts are not on the cooperative kernel)"
            res[key]["day_passes_per_site_day"] = (c["n_day_passes"] - c0["n_day_passes"]) / (S * K)
        if key == "with_snow":
            res[key]["kernel"] = "soil_chunk_kernel<19, 8> (per-site kernel: SnowModel is not on the cooperative kernel)"
            res[key]["sites_with_snow_at_the_end"] = int((out.snow_depth[-1] > 0.0).sum())
    res["snow_ratio"] = res["with_snow"]["site_days_per_s"] / res["without_sand"]["site_days_per_s"]
    res["tolerance_ratio"] = res["with_tolerance_convergence"]["site_days_per_s"] / res["without_sand"]["site_days_per_s"]
    res["ratio"] = res["with_sand"]["site_days_per_s"] / res["without_sand"]["site_days_per_s"]
    res["workload"] = ("%d sites (one wave) x %d timed days after 1 warm-up day, same sites / forcing in all runs: the headline grid's five texture classes, the six incl. "
                       "pure sand, the headline textures with the SnowModel, and with tolerance-based day-pass convergence (non-consecutive days)" % (S, K))
    if args.emulate:
        res["emulated"] = "host emulation - self-test of the probe, not a measurement"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
