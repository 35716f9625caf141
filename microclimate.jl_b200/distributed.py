"""Multi-GPU orchestration: one process per GPU, sites sharded contiguously, NO data-path collective.

Sites are independent (reference: one MicroCache per site, `reinit!` per inputs, src/simulation.jl:266-277), so rank r
integrates sites [lo_r, hi_r) on its own GPU and the only communication is a host-side gather of the selected output
rows to rank 0 (gloo, CPU tensors) after the integration.  Results are independent of the number of ranks.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from . import _abi as A
from . import marshal as M
from .simulation import site_shards


def run_shard(engine_factory: Callable, desc, sites: M.SitesBatch, forcing: M.ForcingBatch, rank: int, world: int,
              fields: Sequence[str] = M.DEFAULT_OUTPUT_FIELDS, chunk_days: Optional[int] = None):
    """Integrate this rank's contiguous shard of a batch that every rank can see (or regenerate)."""
    lo, hi = site_shards(sites.S, world)[rank]
    out = M.OutputBatch(forcing.ndays, desc.n_soil, hi - lo, fields, desc.n_snow, desc.n_heights)
    if hi == lo:
        return out, (lo, hi)
    eng = engine_factory(desc, hi - lo)
    try:
        eng.set_sites(sites.slice(lo, hi))
        f = forcing.site_slice(lo, hi)
        if not chunk_days:
            eng.step_days(0, f, out)
        else:
            d0 = 0
            while d0 < forcing.ndays:
                d1 = min(forcing.ndays, d0 + chunk_days)
                part = M.OutputBatch(d1 - d0, desc.n_soil, hi - lo, fields, desc.n_snow, desc.n_heights)
                eng.step_days(d0, f.day_slice(d0, d1), part)
                out.assign_rows(24 * d0, part)
                d0 = d1
    finally:
        close = getattr(eng, "close", None)
        if close:
            close()
    return out, (lo, hi)


def run_shard_minmax(engine_factory: Callable, desc, wdesc, sites: M.SitesBatch, win, rest: M.ForcingBatch, solar_sites, rank: int, world: int,
                     fields: Sequence[str] = M.DEFAULT_OUTPUT_FIELDS, chunk_days: Optional[int] = None):
    """run_shard for a min/max-driven problem: this rank expands the weather of, and integrates, its contiguous shard (mc_step_days_minmax); `win` / `rest` /
    `solar_sites` are the whole batch's weather inputs, remaining forcing and solar site table, sliced here."""
    lo, hi = site_shards(sites.S, world)[rank]
    out = M.OutputBatch(win.ndays, desc.n_soil, hi - lo, fields, desc.n_snow, desc.n_heights)
    if hi == lo:
        return out, (lo, hi)
    eng = engine_factory(desc, hi - lo)
    try:
        eng.set_sites(sites.slice(lo, hi))
        w, f, ss = win.site_slice(lo, hi), rest.site_slice(lo, hi), (None if solar_sites is None else solar_sites.slice(lo, hi))
        cd = win.ndays if not chunk_days else min(int(chunk_days), win.ndays)
        d0 = 0
        while d0 < win.ndays:
            d1 = min(win.ndays, d0 + cd)
            part = out if (d0 == 0 and d1 == win.ndays) else M.OutputBatch(d1 - d0, desc.n_soil, hi - lo, fields, desc.n_snow, desc.n_heights)
            eng.step_days_minmax(d0, wdesc, w.day_slice(d0, d1), f.day_slice(d0, d1), part, ss)
            if part is not out:
                out.assign_rows(24 * d0, part)
            d0 = d1
    finally:
        close = getattr(eng, "close", None)
        if close:
            close()
    return out, (lo, hi)


def gather_outputs(local: M.OutputBatch, n_sites_total: int, group=None, dst: int = 0) -> Optional[M.OutputBatch]:
    """Host gather of per-rank output shards into one OutputBatch on `dst` (None elsewhere).  `group` must be a
    CPU-capable (gloo) process group; with an nccl default group create one with `dist.new_group(backend="gloo")`."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shards = site_shards(n_sites_total, world)
    width = max(hi - lo for lo, hi in shards)
    # every selected array travels: node fields, scalars, the snow temperature profile and the solve_air! profile fields
    fields = [f for f in A.OUTPUT_NODE_FIELDS + A.OUTPUT_SCALAR_FIELDS + ("snow_temperature",) + A.OUTPUT_PROFILE_FIELDS + A.OUTPUT_PROFILE_SCALARS
              if getattr(local, f, None) is not None]
    n_snow = 0 if local.snow_temperature is None else local.snow_temperature.shape[1]
    n_heights = max([getattr(local, f).shape[1] for f in A.OUTPUT_PROFILE_FIELDS if getattr(local, f, None) is not None], default=0)
    if n_heights == 0 and any(getattr(local, f, None) is not None for f in A.OUTPUT_PROFILE_SCALARS):
        n_heights = 1          # profile scalars only: OutputBatch allocates them whenever heights exist
    full = M.OutputBatch(local.ndays, local.N, n_sites_total, fields, n_snow, n_heights) if rank == dst else None

    def padded(a):
        t = torch.zeros(a.shape[:-1] + (width,), dtype=torch.float64)
        t[..., :a.shape[-1]] = torch.from_numpy(np.ascontiguousarray(a))
        return t
    for f in fields:
        t = padded(getattr(local, f))
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst, group=group)
        if rank == dst:
            for (lo, hi), b in zip(shards, bufs):
                getattr(full, f)[..., lo:hi] = b[..., :hi - lo].numpy()
    st = torch.zeros(width, dtype=torch.int32)
    st[:local.S] = torch.from_numpy(local.status)
    bufs = [torch.empty_like(st) for _ in range(world)] if rank == dst else None
    dist.gather(st, bufs, dst=dst, group=group)
    if rank == dst:
        for (lo, hi), b in zip(shards, bufs):
            full.status[lo:hi] = b[:hi - lo].numpy()
    return full
