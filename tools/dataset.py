"""File-driven runs (microclimate.jl_b200/storage.py containers).

    python tools/dataset.py make  DIR --sites 4096 --days 8 --chunk 2        # synthetic grid -> DIR/sites.mcb + DIR/forcing_dDDDD.mcb (one file per day chunk)
    python tools/dataset.py run   DIR [--lo 0 --hi N] [--resume] [--engine cuda|hostemu]
                                                                             # integrate every chunk in order: DIR/out_dDDDD_sLO-HI.mcb + a state checkpoint after each chunk

`run` memory-maps the chunk files (a whole-grid day chunk is passed to mc_step_days without a copy), writes the selected output
fields of each chunk as one container and the carried state as DIR/state_sLO-HI.mcb; `--resume` restarts after the last
checkpointed chunk.  Ranks of a multi-GPU job call it with disjoint [lo, hi) ranges of the same files (no collective).
`--engine hostemu` runs the kernel body on the CPU through the test harness (development only: no GPU needed, not a product path).
"""
import argparse
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import microclimate_jl_b200  # noqa: E402,F401
from microclimate_jl_b200 import marshal as M, storage as st, synthetic as syn  # noqa: E402

FIELDS = ("soil_temperature", "soil_moisture", "surface_water")


def make(args):
    os.makedirs(args.dir, exist_ok=True)
    S = args.sites
    st.save_sites(os.path.join(args.dir, "sites.mcb"), syn.make_sites(0, S, S), generator="synthetic", seed=syn.SEED)
    for d0 in range(0, args.days, args.chunk):
        nd = min(args.chunk, args.days - d0)
        st.save_forcing(os.path.join(args.dir, f"forcing_d{d0:04d}.mcb"), syn.make_forcing(0, S, S, d0, nd), day0=d0)
    print(f"wrote {S} sites x {args.days} days in chunks of {args.chunk} to {args.dir}")


def run(args):
    chunks = sorted(glob.glob(os.path.join(args.dir, "forcing_d*.mcb")))
    if not chunks:
        raise SystemExit(f"no forcing_d*.mcb under {args.dir}")
    S_all = st.Container(os.path.join(args.dir, "sites.mcb")).header["S"]
    lo, hi = args.lo, (S_all if args.hi is None else args.hi)
    tag = f"s{lo}-{hi}"
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    if args.engine == "hostemu":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from hostemu_engine import HostEmuEngine as Engine
    else:
        from microclimate_jl_b200.simulation import SoilEngine as Engine
    eng = Engine(desc, hi - lo)
    eng.set_sites(st.load_sites(os.path.join(args.dir, "sites.mcb"), lo, hi))
    ck = os.path.join(args.dir, f"state_{tag}.mcb")
    done = 0
    if args.resume and os.path.exists(ck):
        state = st.load_state(ck)
        eng.set_state(state)
        done = int(state.days_done[0])
        print(f"resumed at day {done}")
    t0, n = time.perf_counter(), 0
    for path in chunks:
        hdr = st.Container(path).header
        d0, nd = hdr["day0"], hdr["ndays"]
        if d0 + nd <= done:
            continue
        if d0 != done:
            raise SystemExit(f"{path} starts at day {d0} but the state is at day {done}")
        out = M.OutputBatch(nd, desc.n_soil, hi - lo, FIELDS)
        eng.step_days(d0, st.load_forcing(path, 0, nd, lo, hi), out)
        st.save_outputs(os.path.join(args.dir, f"out_d{d0:04d}_{tag}.mcb"), out, day0=d0)
        st.save_state(ck, eng.get_state())
        done += nd
        n += nd * (hi - lo)
        print(f"days {d0}..{d0 + nd - 1}: {int((out.status != 0).sum())} flagged sites", flush=True)
    dt = time.perf_counter() - t0
    print(f"{n} site-days in {dt:.1f} s = {n / max(dt, 1e-9):.0f} site-days/s incl. file reads and writes")


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("make"); a.add_argument("dir"); a.add_argument("--sites", type=int, default=4096)
    a.add_argument("--days", type=int, default=8); a.add_argument("--chunk", type=int, default=2)
    b = sub.add_parser("run"); b.add_argument("dir"); b.add_argument("--lo", type=int, default=0); b.add_argument("--hi", type=int, default=None)
    b.add_argument("--resume", action="store_true"); b.add_argument("--engine", default="cuda", choices=["cuda", "hostemu"])
    args = ap.parse_args()
    (make if args.cmd == "make" else run)(args)


if __name__ == "__main__":
    main()
