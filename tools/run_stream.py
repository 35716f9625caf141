"""Time-chunked streaming run on one GPU: S synthetic sites x D consecutive days, one day-chunk per launch, the carried state resident
in HBM, the next chunk's forcing generated on host threads while the GPU integrates the current one (BASELINE config 5 shape; the full
1 Mi x 365 d target is this loop with D=365).  Prints per-day and whole-run site-days/s and checks NSPOT sites against the oracle.

    python tools/run_stream.py --sites 1048576 --days 20 --spot 32
"""
import argparse
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from microclimate_jl_b200 import marshal as M, synthetic as syn  # noqa: E402
from microclimate_jl_b200.simulation import SoilEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1 << 20)
    ap.add_argument("--days", type=int, default=20)
    ap.add_argument("--spot", type=int, default=32)
    args = ap.parse_args()
    S, D = args.sites, args.days
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    eng = SoilEngine(desc, S)
    eng.set_sites(syn.make_sites(0, S, S))
    out = M.OutputBatch(1, desc.n_soil, S, ())
    for f in bench.FIELDS:
        setattr(out, f, bench._pinned((24, desc.n_soil, S)))
    bufs = [bench._pin_forcing(syn.make_forcing(0, S, S, 0, 1)), bench._pin_forcing(syn.make_forcing(0, S, S, 0, 1))]
    spot = np.linspace(0, S - 1, args.spot).astype(np.int64) if args.spot else np.zeros(0, np.int64)
    spot_T = np.zeros((D * 24, desc.n_soil, len(spot)))
    spot_th = np.zeros_like(spot_T)
    t_run = time.perf_counter()
    gen = None
    for d in range(D):
        if gen is not None:
            gen.join()                  # forcing of day d was generated while day d-1 ran
        if d + 1 < D:                   # generate day d+1 into the buffer day d-1 used, on host threads, while the GPU integrates day d
            gen = threading.Thread(target=syn.make_forcing, args=(0, S, S, d + 1, 1), kwargs={"out": bufs[(d + 1) % 2]})
            gen.start()
        else:
            gen = None
        t = time.perf_counter()
        rc = eng.step_days(d, bufs[d % 2], out)
        w = time.perf_counter() - t
        spot_T[d * 24:(d + 1) * 24] = out.soil_temperature[:, :, spot]
        spot_th[d * 24:(d + 1) * 24] = out.soil_moisture[:, :, spot]
        print(f"day {d:3d}: {w:6.2f} s  {S / w:9.0f} site-days/s  flagged sites {rc}", flush=True)
    total = time.perf_counter() - t_run
    print(f"RUN: {S} sites x {D} days in {total:.1f} s = {S * D / total:.0f} site-days/s end to end (forcing generation overlapped)", flush=True)
    if args.spot:
        from oracle.oracle import OracleEngine
        worst_T = worst_th = 0.0
        for k, s in enumerate(spot):
            o = OracleEngine(desc, 1)
            o.set_sites(syn.make_sites(int(s), int(s) + 1, S))
            ref = M.OutputBatch(D, desc.n_soil, 1, bench.FIELDS)
            o.step_days(0, syn.make_forcing(int(s), int(s) + 1, S, 0, D, threads=1), ref, n_threads=1)
            worst_T = max(worst_T, float(np.abs(ref.soil_temperature[:, :, 0] - spot_T[:, :, k]).max()))
            worst_th = max(worst_th, float(np.abs(ref.soil_moisture[:, :, 0] - spot_th[:, :, k]).max()))
        print(f"SPOT CHECK vs oracle on {len(spot)} sites x {D} days (all hours, all nodes): max|dT| = {worst_T:.3g} K, max|dtheta| = {worst_th:.3g}", flush=True)


if __name__ == "__main__":
    main()
