#!/usr/bin/env python
"""bench.py — site-days/s of hourly soil temperature + moisture on N B200s (BASELINE.json metric).

A "step" integrates ONE DAY (24 hourly outputs) for every site of the rank's shard of the synthetic grid
(BASELINE.json config 5: N=19 nodes, consecutive days, dynamic soil moisture, hourly rain, no snow); successive steps
are successive days, state resident in HBM.  Weak scaling: every rank owns --sites sites.

  value   kernel-only throughput, forcing already in HBM (CUDA events on the engine's stream, max over ranks)
  e2e     the same metric through the C-ABI call with pinned HOST buffers: H2D forcing + kernel + D2H of the hourly
          soil T and moisture profiles inside the timed region
  --impl reference   the CPU path (the oracle port: Julia is not installable here) on all host cores, bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIELDS = ("soil_temperature", "soil_moisture")
ALG_BYTES_PER_SITE_DAY = (10 * 24 + 7) * 8 + 2 * 19 * 24 * 8     # hourly+daily forcing read, hourly T+theta written
STATUS_BITS = {1: "NAN", 2: "DT_UNDERFLOW", 4: "NEWTON_MAXITER", 8: "STEP_LIMIT"}      # include/microclimate_b200.h MC_STATUS_*


def library_hash():
    """Identity of the built CUDA library: numbers calibrated with ncu (profiles/traffic.json) are only valid for the build they were taken on."""
    import hashlib
    from microclimate_jl_b200 import _lib
    h = hashlib.sha256()
    for d in sorted(_lib.DEPS):
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def walg_per_site_day(counters, site_days):
    """W_alg = sum_k count_k * w_k, counts from the device counters of the timed region, weights = source-level FP64 operation counts of
    the oracle's functions (tools/walg_weights.py, profiles/walg_weights.json).  Stomatal iterations are not credited (see the weights file)."""
    path = os.path.join(ROOT, "profiles", "walg_weights.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        W = json.load(f)["regions"]
    ev = {"other": 24 * counters["n_day_passes"], "rhs": counters["n_rhs"], "obukhov_iter": counters["n_obukhov_iter"], "newton_iter": counters["n_newton_iter"],
          "infil_step": counters["n_infil_steps"], "step_control": counters["n_accept"] + counters["n_reject"]}
    parts = {k: ev[k] * W[k]["flop_eq_per_event"] / site_days for k in ev}
    return sum(parts.values()), parts


def kernel_sass_hash(name="soil_chunk_coop_kernelILi19E"):
    """Identity of the compiled soil kernel itself: hash of its SASS (instruction text + encodings) from `cuobjdump -sass` of the built library.
    Unlike library_hash() it survives edits that do not change the kernel's machine code (new entry points, other translation units)."""
    import hashlib
    import re
    import shutil
    from microclimate_jl_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        txt = subprocess.run([exe, "-sass", _lib.SO], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    h, on, n = hashlib.sha256(), False, 0
    for l in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            on = name in m.group(1)
            continue
        if on and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            h.update(l.strip().encode())
            n += 1
    return h.hexdigest()[:16] if n else None


def calibrated_traffic():
    """DRAM bytes and executed FP64 flops per site-day from the ncu capture of THIS kernel build (tools/ncu_traffic.py writes profiles/traffic.json with the
    library's source hash and the kernel's SASS hash); a capture of another build is refused.  Valid when either identity matches: the source hash
    (nothing changed) or the SASS hash of soil_chunk_coop_kernel<19> (sources changed elsewhere, the kernel's machine code did not)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    t["matched_by"] = "library_hash" if t.get("library_hash") == library_hash() else None
    if t["matched_by"] is None and t.get("kernel_sass_hash") and t["kernel_sass_hash"] == kernel_sass_hash():
        t["matched_by"] = "kernel_sass_hash"
    t["stale"] = t["matched_by"] is None
    return t


def _probe(script, timeout_s):
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script)], capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": "probe exited %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}
    except Exception as e:          # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def sand_probe(timeout_s=300):
    """The six-class texture grid of SURVEY 8d (incl. pure sand) and the SnowModel grid (per-site kernel) next to the headline grid, one wave of sites
    (tools/gpu_sand_probe.py), in a separate process after the headline measurement."""
    return _probe("gpu_sand_probe.py", timeout_s)


def weather_probe(timeout_s=240):
    """Throughput + HBM roofline + oracle parity of the weather kernel, measured by tools/gpu_weather_probe.py in a SEPARATE process
    after the headline measurement is complete: whatever happens there (a CUDA error included) cannot touch this process or its numbers."""
    return _probe("gpu_weather_probe.py", timeout_s)


def _pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()


def _pin_forcing(fb):
    from microclimate_jl_b200 import _abi as A
    for f in A.FORCING_HOURLY + A.FORCING_DAILY:
        a = getattr(fb, f)
        if a is not None:
            p = _pinned(a.shape)
            p[...] = a
            setattr(fb, f, p)
    return fb


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def __enter__(self):
        def loop():
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([x.strip() for x in o.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_run(desc, S_total, s0, n_sites, day0, ndays, threads=0):
    """Time the oracle (CPU restatement of the reference path) on `n_sites` sites x `ndays` days with all host cores."""
    from microclimate_jl_b200 import marshal as M, synthetic as syn
    from oracle.oracle import OracleEngine
    sites = syn.make_sites(s0, s0 + n_sites, S_total)
    eng = OracleEngine(desc, n_sites)
    eng.set_sites(sites)
    times = []
    for d in range(day0, day0 + ndays):
        f = syn.make_forcing(s0, s0 + n_sites, S_total, d, 1)
        out = M.OutputBatch(1, desc.n_soil, n_sites, FIELDS)
        t = time.perf_counter()
        eng.step_days(d, f, out, n_threads=threads)
        times.append(time.perf_counter() - t)
    return times


def main():
    # keep stdout clean for the ONE JSON line: libraries (e.g. NCCL's version banner) print to fd 1
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", type=int, default=1 << 20, help="sites per GPU (weak scaling)")
    ap.add_argument("--cpu-sites", type=int, default=0, help="sites of the bounded CPU sample (0 = sized for ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probes", action="store_true", help="skip the two post-measurement probes (tools/gpu_sand_probe.py, tools/gpu_weather_probe.py; N = 1 only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from microclimate_jl_b200 import marshal as M, synthetic as syn
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    S = args.sites
    S_total = S * max(world, args.gpus)
    cores = os.cpu_count() or 1
    config = {"workload": "synthetic site grid (BASELINE config 5): %d sites/GPU x 1 day per step, N=19 soil nodes, ConsecutiveDayMode, "
                          "DynamicSoilMoisture(1e-6,500,360 s), HourlyRainfall, no snow, hourly T+theta profiles out; every GPU's shard is one full latitude sweep" % S,
              "sites_per_gpu": S, "total_sites": S_total, "days_per_step": 1, "n_soil": 19, "seed": syn.SEED,
              "l2": "inputs_larger_than_L2 (%.2f GB forcing + %.2f GB outputs per step)" % (S * 247 * 8 / 1e9, S * 38 * 24 * 8 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_sites or max(64, min(S, 768 * cores))
        times = cpu_run(desc, n, 0, n, 0, args.warmup + args.steps, threads=cores)      # n sites spanning the whole latitude sweep: same mix as a GPU shard
        tt = times[args.warmup:]
        v = n * len(tt) / sum(tt)
        config = dict(config, sample_sites=n, sample_note="CPU arm: a %d-site sample of the same grid (one full latitude sweep, same per-site workload mix) x 1 day per step; "
                                                            "the metric is a per-site-day rate, the label sites_per_gpu is the GPU arm's" % n)
        line = {"impl": "reference", "metric": "site-days/sec (hourly soil T+moisture)", "value": v, "unit": "site-days/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(tt) / len(tt), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "site-days/s", "cores": cores, "kind": "port",
                                 "sample": "%d sites (one full latitude sweep of the synthetic grid) x 1 day per step, OpenMP over sites; Julia reference not installable here" % n},
                "e2e": {"value": v, "unit": "site-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=real_stdout, flush=True)
        return

    import torch
    import torch.distributed as dist
    from microclimate_jl_b200.simulation import SoilEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ndays = args.warmup + args.steps
    s0 = rank * S
    t_gen = time.perf_counter()
    sites = syn.make_sites(s0, s0 + S, S)            # every rank's shard is one full latitude sweep of S sites (site ids rank*S ..): identical workload mix per GPU
    eng = SoilEngine(desc, S, device=local_rank)
    peak_tf = eng.measure_fp64_peak()
    eng.set_sites(sites)
    # forcing for all warm-up + timed days: pinned host copy (e2e leg) and device-resident copy (value leg)
    from microclimate_jl_b200 import _abi as A
    host_days = [_pin_forcing(syn.make_forcing(s0, s0 + S, S, d, 1)) for d in range(ndays)]
    t_gen = time.perf_counter() - t_gen
    out_host = M.OutputBatch(1, desc.n_soil, S, ())
    for f in FIELDS:
        setattr(out_host, f, _pinned((24, desc.n_soil, S)))
    h2d = sum(getattr(host_days[0], f).nbytes for f in A.FORCING_HOURLY + A.FORCING_DAILY if getattr(host_days[0], f) is not None)
    d2h = sum(getattr(out_host, f).nbytes for f in FIELDS) + 4 * S

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value leg: device-resident forcing + outputs -------------------------------------------------------------
    dev_days = []
    for fb in host_days[:ndays]:
        d = {f: torch.from_numpy(getattr(fb, f)).cuda(non_blocking=True) for f in A.FORCING_HOURLY + A.FORCING_DAILY if getattr(fb, f) is not None}
        dev_days.append(d)
    out_dev = {f: torch.zeros((24, desc.n_soil, S), dtype=torch.float64, device="cuda") for f in FIELDS}
    torch.cuda.synchronize()
    optr = {f: t.data_ptr() for f, t in out_dev.items()}
    day = 0
    for w in range(args.warmup):
        if w == 0:      # first warm-up through the host-buffer entry so its staging buffers exist before the timed e2e leg
            eng.step_days(day, host_days[day], out_host)
        else:
            eng.step_days_device(day, 1, {k: v.data_ptr() for k, v in dev_days[day].items()}, optr)
            eng.synchronize()
        day += 1
    eng.reset_counters()
    snapshot = eng.get_state()          # both legs integrate the SAME days from the same state
    day_timed0 = day
    barrier()
    kernel_ms = []
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.step_days_device(day, 1, {k: v.data_ptr() for k, v in dev_days[day].items()}, optr)
            eng.synchronize()
            kernel_ms.append(eng.counters()["gpu_ms"])          # CUDA events around the launch, on the engine's stream
            day += 1
        barrier()
        wall_value = time.perf_counter() - t0
    dev_time = reduce_max(sum(kernel_ms) * 1e-3)
    counters = eng.counters()
    st = eng.status()
    status_bad = int((st != 0).sum())
    flagged_by_bit = {name: int(((st & bit) != 0).sum()) for bit, name in STATUS_BITS.items()}
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "kernel_ms": kernel_ms, "flagged": status_bad, "flagged_by_bit": flagged_by_bit})
    else:
        per_rank = [{"rank": 0, "kernel_ms": kernel_ms, "flagged": status_bad, "flagged_by_bit": flagged_by_bit}]
    value = world * S * args.steps / dev_time
    del dev_days
    torch.cuda.empty_cache()

    # ---- e2e leg: pinned host buffers through mc_step_days (H2D + kernel + D2H inside the timed region) --------------
    eng.set_state(snapshot)
    day = day_timed0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):           # mc_step_days_async: the H2D of step k+1 and the D2H of step k-1 overlap the kernel of step k
        eng.step_days_async(day, host_days[day], out_host)
        day += 1
    eng.wait()
    barrier()
    e2e_time = reduce_max(time.perf_counter() - t0)
    e2e = world * S * args.steps / e2e_time
    checksum = float(out_host.soil_temperature[-1, 0, :1024].sum())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    per_gpu = value / world
    timed_site_days = S * args.steps
    w_alg, w_parts = walg_per_site_day(counters, timed_site_days)
    cal = calibrated_traffic()
    exe = cal["fp64_flop_per_site_day"] if cal and not cal["stale"] else None
    roof = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_tf, "peak_how": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
            # algorithmic work of THIS run: device counters of the timed region x source-level operation-count weights of the oracle (profiles/walg_weights.json)
            "w_alg": w_alg, "w_alg_parts": w_parts, "achieved": (w_alg * per_gpu / 1e12) if w_alg else None, "frac": (w_alg * per_gpu / 1e12 / peak_tf) if w_alg else None,
            "frac_alg": (w_alg * per_gpu / 1e12 / peak_tf) if w_alg else None,
            "achieved_how": "W_alg (flop-equivalents per site-day: + - * / sqrt = 1, fma = 2, exp/log = 20, pow = 40, atan/cos = 25) x site-days/s per GPU",
            # second view: FP64 flops the kernel EXECUTES (DADD + DMUL + 2 DFMA thread instructions, ncu capture of this build)
            "executed_flop_per_site_day": exe, "achieved_executed": (exe * per_gpu / 1e12) if exe else None, "frac_executed": (exe * per_gpu / 1e12 / peak_tf) if exe else None,
            "traffic": (cal["dram_bytes_per_site_day"] * S) if cal and not cal["stale"] else None,
            "traffic_build_hash": cal.get("library_hash") if cal else None, "library_hash": library_hash(),
            "traffic_kernel_sass_hash": cal.get("kernel_sass_hash") if cal else None, "traffic_matched_by": cal.get("matched_by") if cal else None,
            "traffic_how": ("DRAM bytes per launch (one launch = one step of %d site-days per GPU) from the ncu capture of this build (profiles/traffic.json)" % S) if cal and not cal["stale"]
                           else "no ncu capture of this build (profiles/traffic.json absent or taken on another build): null rather than a stale number",
            "hbm": {"achieved": ALG_BYTES_PER_SITE_DAY * per_gpu / 1e9, "peak": 6536.0, "unit": "GB/s", "frac": ALG_BYTES_PER_SITE_DAY * per_gpu / 1e9 / 6536.0,
                    "alg_bytes_per_site_day": ALG_BYTES_PER_SITE_DAY,
                    "dram_traffic_GBps": (cal["dram_bytes_per_site_day"] * per_gpu / 1e9) if cal and not cal["stale"] else None},
            "work_per_site_day": {k: counters[k] / max(1, timed_site_days) for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps")}}
    line = {"metric": "site-days/sec (hourly soil T+moisture)", "value": value, "unit": "site-days/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_time / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clk.summary(),
            "e2e": {"value": e2e, "unit": "site-days/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_time / args.steps},
            "gpu_launches": args.steps, "roofline": roof, "flagged_sites": status_bad, "flagged_by_bit": flagged_by_bit,
            "per_rank": per_rank, "e2e_pipelined": True, "wall_value_s": wall_value, "setup_s": t_gen, "checksum": checksum}
    if world == 1 and not args.no_cpu_baseline:
        n = args.cpu_sites or max(64, min(S, 768 * cores))
        times = cpu_run(desc, n, 0, n, 0, 2, threads=cores)
        line["cpu_baseline"] = {"value": n / times[1], "unit": "site-days/s", "cores": cores, "kind": "port",
                                "sample": "%d sites (one full latitude sweep) x 1 day (day 1 after a 1-day warm-up), OpenMP over sites, %.1f s; C++ restatement of the reference (Julia not installable here)" % (n, times[1])}
    if world == 1 and not args.no_probes:
        eng.close()                       # the probes are separate processes on the same GPU: give the memory back first
        line["grid_variants"] = sand_probe()
        line["widening"] = {"weather_kernel": weather_probe()}
    print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
