"""Operation-count weights of the algorithmic roofline numerator (SURVEY 8d:  W_alg(site-day) = sum_k count_k * w_k).

Runs the CPU oracle compiled with an operation-counting scalar in place of `double` (oracle/opcount/opcount.cpp: the oracle source unchanged,
every + - * / sqrt counted as 1, fma as 2, math-kernel calls by class) on a sample of the benchmark grid, checks that the counted build
reproduces the plain oracle bit for bit, and divides the per-region tallies by the event counts of the same run - the very events the device
counters count (mc_counters) - to get flop-equivalents per event.  Writes profiles/walg_weights.json, which bench.py multiplies with the
counters of its timed region.

    python tools/walg_weights.py [--sites 192] [--days 3]

Folding weights for math-kernel calls: exp / exp2 / exp10 / log / log10 = 20, pow = 40, atan / cos = 25.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microclimate_jl_b200 import _abi as A, marshal as M, synthetic as syn  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402

REGIONS = ["other", "rhs", "obukhov_iter", "newton_iter", "infil_step", "step_control", "stomatal_iter"]
OPS = ["add_sub", "mul", "div", "sqrt", "fma", "exp", "log", "pow", "trig"]
FOLD = {"add_sub": 1, "mul": 1, "div": 1, "sqrt": 1, "fma": 2, "exp": 20, "log": 20, "pow": 40, "trig": 25}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=192)
    ap.add_argument("--days", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "walg_weights.json"))
    args = ap.parse_args()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "opcount/libmc_oracle_opcount.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, "oracle", "opcount", "libmc_oracle_opcount.so"))
    L.mco_create.restype = C.c_void_p
    L.mco_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64]
    L.mco_set_sites.argtypes = [C.c_void_p, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
    L.mco_step_days.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA), C.c_int]
    L.mco_get_counters.argtypes = [C.c_void_p, C.POINTER(A.Counters)]
    L.opc_collect.argtypes = [C.POINTER(C.c_ulonglong)]
    S, D = args.sites, args.days
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    # sites spread over the whole latitude sweep of the 1 Mi-site grid (the benchmark's mix of climates and textures)
    S_total = 1 << 20
    idx = np.linspace(0, S_total - 1, S).astype(np.int64)
    sites = M.SitesBatch(desc.n_soil, S, True)
    forcing0 = M.ForcingBatch(1, S, hourly_rain=True)
    forcing = M.ForcingBatch(D, S, hourly_rain=True)
    for k, s in enumerate(idx):
        one = syn.make_sites(int(s), int(s) + 1, S_total)
        for f in one.SITE_FIELDS:
            getattr(sites, f)[k] = getattr(one, f)[0]
        for f in one.NODE_FIELDS + ("initial_soil_moisture", "initial_soil_temperature"):
            getattr(sites, f)[:, k] = getattr(one, f)[:, 0]
        f0 = syn.make_forcing(int(s), int(s) + 1, S_total, 0, 1, threads=1)
        f1 = syn.make_forcing(int(s), int(s) + 1, S_total, 1, D, threads=1)
        for f in A.FORCING_HOURLY + A.FORCING_DAILY:
            if getattr(f0, f) is not None:
                getattr(forcing0, f)[:, k] = getattr(f0, f)[:, 0]
                getattr(forcing, f)[:, k] = getattr(f1, f)[:, 0]
    forcing.day_of_year[:] = [(d % 365) + 1 for d in range(1, 1 + D)]
    fields = ("soil_temperature", "soil_moisture")

    def run(step, counters):
        outs = []
        for d0, fb in ((0, forcing0), (1, forcing)):
            out = M.OutputBatch(fb.ndays, desc.n_soil, S, fields)
            step(d0, fb, out)
            outs.append(out)
            if d0 == 0:
                base = counters()
        return outs[1], base, counters()

    # plain oracle
    eng = OracleEngine(desc, S); eng.set_sites(sites)
    ref, _, _ = run(lambda d0, fb, out: eng.step_days(d0, fb, out, n_threads=0), eng.counters)
    # counted oracle (single thread: the tallies are thread-local)
    h = L.mco_create(C.byref(desc), S)
    cs, ci = sites.c_sites(), sites.c_initial()
    assert L.mco_set_sites(h, C.byref(cs), C.byref(ci)) == 0

    def cstep(d0, fb, out):
        cf, co = fb.c_struct(), out.c_struct()
        assert L.mco_step_days(h, d0, fb.ndays, C.byref(cf), C.byref(co), 1) >= 0
        if d0 == 0:
            L.opc_reset()                      # the first day (initialisation transients) is warm-up: weights come from days 1..D

    def ccounters():
        c = A.Counters(); L.mco_get_counters(h, C.byref(c)); return c.as_dict()
    got, base, cnt = run(cstep, ccounters)
    assert np.array_equal(got.soil_temperature, ref.soil_temperature) and np.array_equal(got.soil_moisture, ref.soil_moisture), "counted build differs from the oracle"
    tally = (C.c_ulonglong * (len(REGIONS) * len(OPS)))()
    L.opc_collect(tally)
    T = np.array(list(tally), dtype=np.float64).reshape(len(REGIONS), len(OPS))
    ev = {k: cnt[k] - base[k] for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes")}
    site_days = S * D
    events = {"other": 24 * ev["n_day_passes"], "rhs": ev["n_rhs"], "obukhov_iter": ev["n_obukhov_iter"], "newton_iter": ev["n_newton_iter"],
              "infil_step": ev["n_infil_steps"], "step_control": ev["n_accept"] + ev["n_reject"], "stomatal_iter": ev["n_infil_steps"]}
    fold = np.array([FOLD[o] for o in OPS], dtype=np.float64)
    table = {}
    for r, name in enumerate(REGIONS):
        ops = {o: T[r, k] / events[name] for k, o in enumerate(OPS)}
        table[name] = {"per_event_ops": ops, "flop_eq_per_event": float((T[r] * fold).sum() / events[name]), "events_per_site_day": events[name] / site_days,
                       "flop_eq_per_site_day": float((T[r] * fold).sum() / site_days)}
    credited = [n for n in REGIONS if n != "stomatal_iter"]
    out = {"what": "FP64 flop-equivalents per event of the reference algorithm (oracle with an operation-counting scalar), synthetic benchmark grid, N=19, NoSnow, "
                   "DynamicSoilMoisture; event counts are the device counters' (mc_counters)",
           "sample": f"{S} sites spread over the 1 Mi-site grid x days 1..{D} (day 0 excluded as warm-up)", "fold": FOLD,
           "event_of_region": {"other": "24 x n_day_passes (site-hours)", "rhs": "n_rhs", "obukhov_iter": "n_obukhov_iter", "newton_iter": "n_newton_iter",
                               "infil_step": "n_infil_steps", "step_control": "n_accept + n_reject", "stomatal_iter": "n_infil_steps (literal loop; NOT credited in W_alg: "
                               "the kernel resolves non-terminating stomatal iterations from their cycle)"},
           "regions": table,
           "w_alg_per_site_day_sample": float(sum(table[n]["flop_eq_per_site_day"] for n in credited)),
           "raw_ops_per_site_day_sample": {o: float(T[[REGIONS.index(n) for n in credited], k].sum() / site_days) for k, o in enumerate(OPS)}}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (v["flop_eq_per_event"], v["flop_eq_per_site_day"]) for k, v in table.items()}, indent=1))
    print("W_alg per site-day (sample):", out["w_alg_per_site_day_sample"])


if __name__ == "__main__":
    main()
