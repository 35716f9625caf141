"""Iteration-count census with the oracle's trace hooks (CPU only): the numbers quoted in profiles/README.md.

    python tools/oracle_census.py warp   [--sites 4096]            # attempts / Newton iterations per site-hour, lane efficiencies, round-count model
    python tools/oracle_census.py tails  [--chunks 96 --days 8]    # scan of chunks spread over the 1 Mi-site grid for day-long stragglers

Uses mco_set_trace / mco_set_trace_sub of oracle/mc_oracle.cpp (analysis hooks; they do not change results).
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import microclimate_jl_b200  # noqa: E402,F401
from microclimate_jl_b200 import marshal as M, synthetic as syn  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402
import oracle.oracle as oo  # noqa: E402


def traced_run(s0, S, S_total, nd, threads=0):
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    eng = OracleEngine(desc, S)
    eng.set_sites(syn.make_sites(s0, s0 + S, S_total))
    rows = nd * 24
    att, nw, sub = np.zeros((S, rows), np.int32), np.zeros((S, rows), np.int32), np.zeros((S, rows, 10), np.int32)
    lib = oo.lib()
    lib.mco_set_trace(att.ctypes.data_as(ctypes.c_void_p), nw.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(rows))
    lib.mco_set_trace_sub(sub.ctypes.data_as(ctypes.c_void_p))
    out = M.OutputBatch(nd, desc.n_soil, S, fields=("soil_temperature",))
    eng.step_days(0, syn.make_forcing(s0, s0 + S, S_total, 0, nd), out, threads or (os.cpu_count() or 1))
    lib.mco_set_trace(None, None, ctypes.c_int64(0))
    lib.mco_set_trace_sub(None)
    eng.close()
    return att, nw, sub


def warp(args):
    S = args.sites
    att, nw, sub = traced_run(0, S, S, 2)
    a, n, x = att[:, 24:].astype(np.int64), nw[:, 24:].astype(np.int64), sub[:, 24:, :].astype(np.int64)      # day 1
    print(f"attempts per site-hour: mean {a.mean():.1f} p99 {np.percentile(a, 99):.0f} max {a.max()};  Newton iterations per site-hour: mean {n.mean():.1f} p99 {np.percentile(n, 99):.0f} max {n.max()}")
    w = x.reshape(S // 32, 32, 24, 10)
    print(f"Newton lane efficiency of a 32-site warp: per sub-step lock-step {x.mean() * 10 / w.max(axis=1).sum(axis=2).mean():.3f}, per hour {x.mean() * 10 / w.sum(axis=3).max(axis=1).mean():.3f}")
    print("Newton iterations per sub-step histogram (0..12+):", np.bincount(np.minimum(x.ravel(), 12)))
    rhs = 6 * a + 1
    for TB in (1024, 256, 128):
        if S % TB:
            continue
        b, mm = rhs.reshape(S // TB, TB, 24), (n + 10).reshape(S // TB, TB, 24)
        both = b + mm
        print(f"block of {TB}: ODE rounds/day lock-step per hour {b.max(axis=1).sum(axis=1).mean():.0f} vs demand {rhs.sum(axis=1).mean():.0f}; "
              f"ODE + moisture asynchronous over the day {both.sum(axis=2).max(axis=1).mean():.0f} vs demand {both.sum(axis=2).mean():.0f}")


def tails(args):
    ST, chunk = 1 << 20, 256
    A, N, ids = [], [], []
    for c in range(args.chunks):
        s0 = (c * (ST // args.chunks)) // 32 * 32
        att, nw, _ = traced_run(s0, chunk, ST, args.days)
        A.append(att.reshape(chunk, args.days, 24)); N.append(nw.reshape(chunk, args.days, 24)); ids.append(np.arange(s0, s0 + chunk))
    A, N, ids = np.concatenate(A), np.concatenate(N), np.concatenate(ids)
    for d in range(args.days):
        a, n = A[:, d].sum(axis=1), N[:, d].sum(axis=1)
        print(f"day {d}: attempts/site-day mean {a.mean():.0f} max {a.max()} (max per hour {A[:, d].max()}) | Newton/site-day mean {n.mean():.0f} p99 {np.percentile(n, 99):.0f} "
              f"max {n.max()} at site {ids[n.argmax()]} (max per hour {N[:, d].max()})")


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("warp"); a.add_argument("--sites", type=int, default=4096)
    b = sub.add_parser("tails"); b.add_argument("--chunks", type=int, default=96); b.add_argument("--days", type=int, default=8)
    args = ap.parse_args()
    (warp if args.cmd == "warp" else tails)(args)


if __name__ == "__main__":
    main()
