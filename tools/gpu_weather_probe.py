"""Throughput probe of the weather kernel (widening rows f2 / f3) on one B200: prints ONE JSON object.

    python tools/gpu_weather_probe.py [--sites 262144] [--days 8] [--reps 20]

Device-resident inputs and outputs (mc_populate_weather_device), W = 3 warm-up launches, then `reps` back-to-back launches between two
mc_synchronize calls, host clock (each launch is >= 1 ms, the launch overhead ~10 us).  The hourly outputs of one launch (9 x 24 x days x sites x 8 B = 3.6 GB
at the defaults) are far larger than the 126 MB L2, so every launch streams from / to HBM.  Roofline: HBM — algorithmic bytes = everything the kernel must read
(value series, site arrays, clear-sky arrays) plus everything it must write, divided by the launch time, against MEASURED_PEAKS.json's copy bandwidth.
The first `--check` sites are compared bit for bit with the CPU oracle.  bench.py runs this in a subprocess (N = 1 only) and attaches the object as
`widening.weather_kernel`; a failure here can never take the headline measurement down.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=262144)
    ap.add_argument("--days", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--check", type=int, default=2048)
    ap.add_argument("--emulate", action="store_true", help="CPU-tier self-test of this script: host tensors and the host emulation of mc_weather.cu (tests/hostemu); no timing claim")
    args = ap.parse_args()
    import numpy as np
    import torch
    from microclimate_jl_b200 import _abi as A, marshal as M, synthetic as syn, weather as W
    from microclimate_jl_b200.simulation import SoilEngine
    S, D = args.sites, args.days
    H = 24 * D
    rng = np.random.default_rng(20260719)
    days = (np.arange(D) * 37 % 365 + 1).astype(np.int32)
    lat = rng.uniform(-60, 70, S)
    slope = np.where(np.arange(S) % 3 == 0, 15.0, 0.0)
    aspect = rng.uniform(0, 360, S)
    tmin = rng.uniform(255.0, 295.0, (D, S))
    f = W.minmax_forcings(reference_temperature_min=tmin, reference_temperature_max=tmin + rng.uniform(3.0, 18.0, (D, S)),
                          reference_wind_speed_min=rng.uniform(0.1, 1.0, (D, S)), reference_wind_speed_max=rng.uniform(1.5, 8.0, (D, S)),
                          reference_humidity_min=rng.uniform(0.1, 0.6, (D, S)), reference_humidity_max=rng.uniform(0.6, 1.05, (D, S)),
                          cloud_cover_min=rng.uniform(0.0, 0.5, (D, S)), cloud_cover_max=rng.uniform(0.5, 1.05, (D, S)))
    direct = rng.uniform(0, 900, (H, S)) * (rng.random((H, S)) < 0.5)
    diffuse = rng.uniform(0, 150, (H, S))
    psite = rng.uniform(7.5e4, 1.02e5, S)
    win = W.WeatherInBatch(days, S, {q: v.values for q, v in f.items()}, direct_horizontal=direct, diffuse_horizontal=diffuse, pressure_site=psite)
    sites = W.SolarSitesBatch(lat, slope, aspect)
    w = W.weather_desc()

    if args.emulate:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import hostemu_engine as HE
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).clone()      # noqa: E731
        device, sync = "cpu", (lambda: None)
    else:
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()       # noqa: E731
        device, sync = "cuda", torch.cuda.synchronize
    pd = C.POINTER(C.c_double)
    ptr = lambda t: C.cast(C.c_void_p(t.data_ptr()), pd)                      # noqa: E731
    keep = []

    def up(a):
        t = dev(a)
        keep.append(t)
        return ptr(t)
    cs = A.SolarSitesSoA(up(lat), up(slope), up(aspect), None)
    ci = A.WeatherInSoA()
    read_bytes = 3 * S * 8 + S * 8
    for q in A.WEATHER_QUANTITIES:
        arr = getattr(ci, q + "_values")
        for i, v in enumerate(win.forcings[q]):
            arr[i] = up(v)
            read_bytes += v.nbytes
    ci.direct_horizontal, ci.diffuse_horizontal, ci.pressure_site = up(direct), up(diffuse), up(psite)
    read_bytes += direct.nbytes + diffuse.nbytes
    ci.day_of_year = days.ctypes.data_as(C.POINTER(C.c_int32))
    co = A.WeatherOutSoA()
    outs = {}
    write_bytes = 0
    for k in A.WEATHER_OUT_HOURLY + A.WEATHER_OUT_DAILY:
        outs[k] = torch.full((H if k in A.WEATHER_OUT_HOURLY else D, S), float("nan"), dtype=torch.float64, device=device)
        setattr(co, k, ptr(outs[k]))
        write_bytes += outs[k].numel() * 8
    sync()

    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    import contextlib
    eng_cm = contextlib.closing(HE.WeatherApiEmu(desc, S)) if args.emulate else SoilEngine(desc, S)
    with eng_cm as eng:
        L, h = eng._L, eng.h
        eng_sync = (lambda: None) if args.emulate else eng.synchronize
        last_error = L.wa_last_error if args.emulate else L.mc_last_error
        if args.emulate:
            L.wa_allow_foreign_pointers(1)          # the emulation tells "device" from host memory; this entry point is handed caller-owned device memory

        def launch():
            rc = L.mc_populate_weather_device(h, D, C.byref(w), C.byref(cs), C.byref(ci), C.byref(co))
            if rc < 0:
                raise RuntimeError("mc_populate_weather_device rc=%d: %s" % (rc, last_error(h).decode()))
        for _ in range(1 if args.emulate else 3):
            launch()
        eng_sync()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            launch()
        eng_sync()
        ms = 1e3 * (time.perf_counter() - t0) / args.reps
        sync()

    # parity of the first `check` sites against the CPU oracle (test infrastructure, used here as the checker only)
    from oracle import oracle as O
    n = min(args.check, S)
    sub = W.WeatherInBatch(days, n, {q: tuple(np.ascontiguousarray(v[:, :n]) for v in win.forcings[q]) for q in A.WEATHER_QUANTITIES},
                           direct_horizontal=direct[:, :n], diffuse_horizontal=diffuse[:, :n], pressure_site=psite[:n])
    ref = W.WeatherOutBatch(D, n)
    O.populate_weather(w, sub, ref, W.SolarSitesBatch(lat[:n], slope[:n], aspect[:n]))
    identical = all(np.array_equal(getattr(ref, k), outs[k][:, :n].cpu().numpy(), equal_nan=True) for k in A.WEATHER_OUT_HOURLY + A.WEATHER_OUT_DAILY)
    # the CPU path beside it: the oracle (OpenMP over sites, all host cores) on a bounded sample of the same workload
    nc = min(S, 1024 if args.emulate else 32768)
    subc = W.WeatherInBatch(days, nc, {q: tuple(np.ascontiguousarray(v[:, :nc]) for v in win.forcings[q]) for q in A.WEATHER_QUANTITIES},
                            direct_horizontal=direct[:, :nc], diffuse_horizontal=diffuse[:, :nc], pressure_site=psite[:nc])
    refc, sitesc = W.WeatherOutBatch(D, nc), W.SolarSitesBatch(lat[:nc], slope[:nc], aspect[:nc])
    O.populate_weather(w, subc, refc, sitesc)                     # warm-up (page the output arrays in)
    t0 = time.perf_counter()
    O.populate_weather(w, subc, refc, sitesc)
    cpu_s = time.perf_counter() - t0

    peak = 6536.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fp:
            peak = float(json.load(fp).get("hbm_gbs", peak))
    except Exception:
        pass
    gbs = (read_bytes + write_bytes) / (ms * 1e-3) / 1e9
    print(json.dumps({
        **({"emulated": "host emulation - self-test of the probe, not a measurement"} if args.emulate else {}),
        "kernel": "weather_kernel (populate_weather! + solar geometry, one thread per site-day)", "sites": S, "days_per_launch": D, "launches_timed": args.reps, "warmup": 3,
        "ms_per_launch": ms, "site_days_per_s": S * D / (ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                     "alg_bytes_per_site_day": (read_bytes + write_bytes) / (S * D), "how": "bytes the kernel must read + write per launch / host-timed launch duration "
                     "(back-to-back launches between two stream synchronisations); peak = MEASURED_PEAKS.json hbm_gbs"},
        "l2": "outputs_larger_than_L2 (%.1f GB written per launch)" % (write_bytes / 1e9),
        "parity": {"sites_checked": n, "bit_identical_to_oracle": bool(identical)},
        "cpu_baseline": {"value": nc * D / cpu_s, "unit": "site-days/s", "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": "%d sites x %d days of the same workload, oracle/mc_weather_oracle.cpp with OpenMP over sites, %.2f s" % (nc, D, cpu_s)}}))


if __name__ == "__main__":
    main()
