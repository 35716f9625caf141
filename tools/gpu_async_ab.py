"""A/B runs of the asynchronous kernel's launch knobs in ONE process (development): for every configuration integrate PROFILE_DAYS days of
PROFILE_SITES synthetic sites (forcing resident in HBM) and print the kernel time of each day plus the kernel's work census
(MC_ASYNC_STATS: claims / items / cycles per task type, idle polls).  Configurations: "threads slots sermin" triples, or "coop" / "legacy".

    python tools/gpu_async_ab.py "1024 64 8" "1024 64 1" coop
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from microclimate_jl_b200 import _abi as A, marshal as M, synthetic as syn  # noqa: E402
from microclimate_jl_b200.simulation import SoilEngine  # noqa: E402

S = int(os.environ.get("PROFILE_SITES", "37888"))
nd = int(os.environ.get("PROFILE_DAYS", "2"))
TYPES = ["COND", "P2", "HOIST", "MU", "MF", "SURF", "CTRL", "THOMAS", "SUB", "SITE"]
desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
sites = syn.make_sites(0, S, S)
days = []
for d in range(nd):
    f = syn.make_forcing(0, S, S, d, 1)
    days.append({k: torch.from_numpy(getattr(f, k)).cuda() for k in A.FORCING_HOURLY + A.FORCING_DAILY if getattr(f, k) is not None})
out = {f: torch.zeros((24, desc.n_soil, S), dtype=torch.float64, device="cuda") for f in ("soil_temperature", "soil_moisture")}
torch.cuda.synchronize()
os.environ["MC_ASYNC_STATS"] = "1"
ref = None
for cfg in (sys.argv[1:] or ["1024 64 8"]):
    for k in ("MC_KERNEL", "MC_ASYNC_THREADS", "MC_ASYNC_SLOTS", "MC_ASYNC_SERMIN", "MC_ASYNC_AFFINITY"):
        os.environ.pop(k, None)
    if cfg in ("coop", "legacy"):
        os.environ["MC_KERNEL"] = cfg
    else:
        os.environ["MC_KERNEL"] = "async"
        th, sl, sm, *rest = cfg.split()
        os.environ["MC_ASYNC_THREADS"], os.environ["MC_ASYNC_SLOTS"], os.environ["MC_ASYNC_SERMIN"] = th, sl, sm
        os.environ["MC_ASYNC_AFFINITY"] = rest[0] if rest else "0"
    eng = SoilEngine(desc, S)
    eng.set_sites(sites)
    ms = []
    for d in range(nd):
        eng.step_days_device(d, 1, {k: v.data_ptr() for k, v in days[d].items()}, {k: v.data_ptr() for k, v in out.items()})
        eng.synchronize()
        ms.append(eng.counters()["gpu_ms"])
    chk = float(out["soil_temperature"].sum().item())
    if ref is None:
        ref = chk
    print(f"== {os.path.basename(os.environ.get('MC_LIB', 'default'))} {cfg:14s} ms/day {' '.join('%.1f' % x for x in ms)}  -> {S / (ms[-1] * 1e-3):9.0f} site-days/s  checksum {'same' if chk == ref else 'DIFFERENT'}", flush=True)
    if cfg not in ("coop", "legacy"):
        buf = (C.c_ulonglong * 64)()
        n = eng._L.mc_get_async_stats(eng.h, buf, 64)
        v = list(buf)[:n]
        tot_cyc = sum(v[3 * i + 2] for i in range(len(TYPES))) + v[3 * len(TYPES) + 1]
        for i, name in enumerate(TYPES):
            claims, items, cyc = v[3 * i:3 * i + 3]
            if claims:
                print(f"   {name:7s} claims {claims:11d} items/claim {items / claims:6.2f} cycles/claim {cyc / claims:9.0f} share of warp time {100 * cyc / max(1, tot_cyc):5.1f} %")
        print(f"   idle    polls  {v[3 * len(TYPES)]:11d} cycles/poll {v[3 * len(TYPES) + 1] / max(1, v[3 * len(TYPES)]):9.0f} share of warp time {100 * v[3 * len(TYPES) + 1] / max(1, tot_cyc):5.1f} %", flush=True)
    eng.close()
