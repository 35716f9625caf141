"""The north-star run: S synthetic sites x D consecutive days (default 1 Mi x 365), the ONE grid sharded over the ranks of the job
(strong scaling; `torchrun --nproc-per-node N tools/run_year.py ...`, or plain `python` for one GPU).

Per chunk of days every rank generates its shard's forcing ON THE DEVICE (microclimate_jl_b200.synthetic_device: same definition as the
host generator, no host round trip), integrates it with the state resident in HBM (SoilEngine.step_days_device), and records the kernel time
(CUDA events on the engine's stream), the per-status-bit count of flagged sites and the outputs of its spot-check sites.  At the end the
spot-check sites are re-integrated by the CPU oracle from exactly the forcing the device used (copied back column-wise) and compared at every
hour and node.  Rank 0 writes the per-chunk trace and a JSON summary (value = whole-grid site-days / max-over-ranks device time).

    python tools/run_year.py --sites 1048576 --days 365 --chunk 1 --spot 64 --out gpurun_out/year_1gpu
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STATUS_BITS = {1: "NAN", 2: "DT_UNDERFLOW", 4: "NEWTON_MAXITER", 8: "STEP_LIMIT"}      # include/microclimate_b200.h MC_STATUS_*


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1 << 20, help="sites of the whole grid (sharded over the ranks)")
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--chunk", type=int, default=1, help="days per kernel launch")
    ap.add_argument("--spot", type=int, default=64, help="spot-check sites over the whole grid (0 = none)")
    ap.add_argument("--spot-budget-s", type=float, default=240.0, help="wall-clock bound of the CPU spot check per rank (sites not reached are reported as skipped)")
    ap.add_argument("--with-sand", action="store_true")
    ap.add_argument("--out", default="gpurun_out/year")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    from microclimate_jl_b200 import _abi as A, marshal as M, synthetic as syn, synthetic_device as sd
    from microclimate_jl_b200.simulation import SoilEngine, site_shards
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev)
        # the result gather is host-side pickles after a CPU spot check that may take minutes on some ranks: its own gloo group with a long timeout,
        # so that a slow rank cannot run into NCCL's 10-minute watchdog (the first 8-GPU year run lost its summary exactly there)
        host_pg = dist.new_group(backend="gloo", timeout=datetime.timedelta(hours=2))
    S_total, D, C = args.sites, args.days, args.chunk
    lo, hi = site_shards(S_total, world)[rank]
    S = hi - lo
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    N = desc.n_soil
    eng = SoilEngine(desc, S, device=local_rank)
    eng.set_sites(syn.make_sites(lo, hi, S_total, with_sand=args.with_sand))
    gen = sd.DeviceForcing(lo, hi, S_total, C, dev)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    open(f"{args.out}_rank{rank}_trace.jsonl", "w").close()
    out_dev = {f: torch.zeros((24 * C, N, S), dtype=torch.float64, device=dev) for f in ("soil_temperature", "soil_moisture")}
    optr = {f: t.data_ptr() for f, t in out_dev.items()}
    # spot-check sites: spread over the whole grid, each rank keeps the ones in its shard
    spot_global = np.unique(np.linspace(0, S_total - 1, args.spot).astype(np.int64)) if args.spot else np.zeros(0, np.int64)
    spot = spot_global[(spot_global >= lo) & (spot_global < hi)] - lo
    spot_idx = torch.as_tensor(spot, device=dev)
    spot_T, spot_th, spot_forcing = [], [], []
    trace = []
    bits_names = sorted(STATUS_BITS)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
    torch.cuda.synchronize()
    t_run = time.perf_counter()
    d = 0
    while d < D:
        c = min(C, D - d)
        if c != C:                                   # ragged last chunk
            gen = sd.DeviceForcing(lo, hi, S_total, c, dev)
        ptrs, doy = gen.generate(d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.step_days_device(d, c, ptrs, optr)
        eng.synchronize()
        wall = time.perf_counter() - t0
        ms = eng.counters()["gpu_ms"]
        st = eng.status()
        by_bit = {STATUS_BITS[b]: int(((st & b) != 0).sum()) for b in bits_names}
        trace.append({"day0": d, "days": c, "gpu_ms": ms, "wall_s": wall, "flagged": int((st != 0).sum()), "by_bit": by_bit})
        with open(f"{args.out}_rank{rank}_trace.jsonl", "a") as tf:      # incremental: a cut-off run keeps what it measured
            tf.write(json.dumps(trace[-1]) + "\n")
        if len(spot):
            spot_T.append(out_dev["soil_temperature"][:24 * c].index_select(2, spot_idx).cpu().numpy())
            spot_th.append(out_dev["soil_moisture"][:24 * c].index_select(2, spot_idx).cpu().numpy())
            spot_forcing.append(gen.host_columns(spot))
        if rank == 0 and (d % 10 == 0 or d + c >= D):
            print(f"[rank0] day {d:3d}+{c}: {ms / 1e3:7.3f} s  {S * c / (ms * 1e-3):9.0f} site-days/s/GPU  flagged {trace[-1]['flagged']} {by_bit}", file=sys.stderr, flush=True)
        d += c
    torch.cuda.synchronize()
    run_wall = time.perf_counter() - t_run
    dev_s = sum(t["gpu_ms"] for t in trace) * 1e-3
    counters = eng.counters()
    # spot check against the oracle, from the very forcing the device used
    worst_T = worst_th = 0.0
    spot_done = 0
    t_spot = time.perf_counter()
    if len(spot):
        from oracle.oracle import OracleEngine
        T_dev, th_dev = np.concatenate(spot_T, axis=0), np.concatenate(spot_th, axis=0)
        sites_all = syn.make_sites(lo, hi, S_total, with_sand=args.with_sand)
        for k, s in enumerate(spot):
            if time.perf_counter() - t_spot > args.spot_budget_s:      # a site whose Newton solves run to the cap costs the CPU ~200x a normal one
                break
            o = OracleEngine(desc, 1)
            o.set_sites(sites_all.slice(int(s), int(s) + 1))
            d0 = 0
            rows_T, rows_th = [], []
            for fb in spot_forcing:
                f1 = fb.site_slice(k, k + 1)
                f1.day_of_year = fb.day_of_year
                ref = M.OutputBatch(fb.ndays, N, 1, ("soil_temperature", "soil_moisture"))
                o.step_days(d0, f1, ref, n_threads=1)
                rows_T.append(ref.soil_temperature[:, :, 0]); rows_th.append(ref.soil_moisture[:, :, 0])
                d0 += fb.ndays
            worst_T = max(worst_T, float(np.abs(np.concatenate(rows_T) - T_dev[:, :, k]).max()))
            worst_th = max(worst_th, float(np.abs(np.concatenate(rows_th) - th_dev[:, :, k]).max()))
            spot_done += 1
    mine = {"rank": rank, "sites": S, "device_s": dev_s, "wall_s": run_wall, "trace": trace, "spot_sites": int(spot_done), "spot_sites_planned": int(len(spot)), "spot_max_dT": worst_T,
            "spot_max_dtheta": worst_th, "counters": {k: int(v) for k, v in counters.items() if k.startswith("n_")}}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=host_pg)
    else:
        gathered = [mine]
    if rank == 0:
        nchunks = len(trace)
        per_chunk_max = [max(g["trace"][i]["gpu_ms"] for g in gathered) for i in range(nchunks)]
        dev_max = max(g["device_s"] for g in gathered)
        flagged_last = {b: sum(g["trace"][-1]["by_bit"][b] for g in gathered) for b in STATUS_BITS.values()}
        med = float(np.median(per_chunk_max[:-1] if (D % C and nchunks > 1) else per_chunk_max))
        summary = {"workload": f"synthetic grid (BASELINE config 5): {S_total} sites x {D} days, N=19, hourly T+theta out, {C}-day launches, "
                               f"sharded over {world} GPU(s) (strong scaling), forcing generated on device", "n_gpus": world, "sites": S_total, "days": D,
                   "with_sand": args.with_sand,
                   "value": S_total * D / dev_max, "unit": "site-days/s", "device_s_max_over_ranks": dev_max, "wall_s": max(g["wall_s"] for g in gathered),
                   "device_s_per_rank": [g["device_s"] for g in gathered],
                   "chunk_ms_median": med, "chunk_ms_max": max(per_chunk_max), "slowest_chunk_over_median": max(per_chunk_max) / med,
                   "chunks_slower_than_1p10_median": int(sum(1 for x in per_chunk_max if x > 1.10 * med)),
                   "flagged_by_bit_final": flagged_last, "flagged_final": sum(g["trace"][-1]["flagged"] for g in gathered),
                   "spot_check": {"sites": sum(g["spot_sites"] for g in gathered), "sites_planned": sum(g["spot_sites_planned"] for g in gathered), "days": D, "max_abs_dT_K": max(g["spot_max_dT"] for g in gathered),
                                  "max_abs_dtheta": max(g["spot_max_dtheta"] for g in gathered)},
                   "work_per_site_day": {k: sum(g["counters"][k] for g in gathered) / (S_total * D) for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps")}}
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out + ".json", "w") as f:
            json.dump(summary, f, indent=1)
        with open(args.out + "_trace.txt", "w") as f:
            f.write("# day0 days  max-over-ranks gpu_ms  site-days/s  flagged(all ranks)  by status bit; per-rank gpu_ms\n")
            for i in range(nchunks):
                t0 = gathered[0]["trace"][i]
                bb = {b: sum(g["trace"][i]["by_bit"][b] for g in gathered) for b in STATUS_BITS.values()}
                f.write(f"{t0['day0']:4d} {t0['days']:2d} {per_chunk_max[i]:10.2f} {S_total * t0['days'] / (per_chunk_max[i] * 1e-3):10.0f} "
                        f"{sum(g['trace'][i]['flagged'] for g in gathered):7d} {json.dumps(bb)} {' '.join('%.1f' % g['trace'][i]['gpu_ms'] for g in gathered)}\n")
        print(json.dumps(summary), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
