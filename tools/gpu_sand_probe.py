"""Kernel throughput of the benchmark grid WITH the pure-sand texture class (SURVEY 8d's six classes) next to the headline grid's five, at one full wave
(151 552 sites), on one B200: prints ONE JSON object.

    python tools/gpu_sand_probe.py [--sites 151552] [--steps 3]

The headline grid of bench.py replaces "sand" by loamy sand because pure sand drives the infiltration Newton loop to its 500-iteration cap; both
arms of the benchmark run the same grid, so the ratio is fair, but the absolute number is on the easier grid.  This probe integrates the same sites, days and forcing
twice — textures with and without sand — through mc_step_days (host buffers), one warm-up day then `steps` timed days, kernel time from the CUDA events of the
library (mc_counters.gpu_ms).  bench.py runs it in a subprocess at N = 1 and attaches the object as `with_sand`.
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
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    days = [syn.make_forcing(0, S, S, d, 1, threads=os.cpu_count() or 1) for d in range(K + 1)]
    out = M.OutputBatch(1, desc.n_soil, S, ("soil_temperature", "soil_moisture"))
    res = {}
    for sand in (False, True):
        with SoilEngine(desc, S) as eng:
            eng.set_sites(syn.make_sites(0, S, S, with_sand=sand))
            eng.step_days(0, days[0], out)                      # warm-up day (also the spin-up from the uniform initial profile)
            c0 = eng.counters()                                  # counters of the warm-up day are subtracted (mc_reset_counters is a product-only entry point)
            ms = []
            for d in range(1, K + 1):
                t0 = time.perf_counter()
                eng.step_days(d, days[d], out)
                ms.append(eng.counters().get("gpu_ms") or 1e3 * (time.perf_counter() - t0))      # the emulation has no device clock
            c, st = eng.counters(), (out.status if args.emulate else eng.status())
        res["with_sand" if sand else "without_sand"] = {
            "site_days_per_s": S * K / (sum(ms) * 1e-3), "kernel_ms_per_day": ms,
            "flagged_by_bit": {name: int(((st & bit) != 0).sum()) for bit, name in STATUS_BITS.items()},
            "work_per_site_day": {k: (c[k] - c0[k]) / (S * K) for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps")}}
    res["ratio"] = res["with_sand"]["site_days_per_s"] / res["without_sand"]["site_days_per_s"]
    res["workload"] = "%d sites (one wave) x %d timed days after 1 warm-up day, same sites / forcing in both runs, textures incl. pure sand vs the headline grid's" % (S, K)
    if args.emulate:
        res["emulated"] = "host emulation - self-test of the probe, not a measurement"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
