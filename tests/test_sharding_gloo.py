"""N>1 host path on CPU: world_size-2 gloo, contiguous site shards, no data-path collective, host gather to rank 0.
The per-rank engine here is the host-emulated kernel body (tests/hostemu) because this tier has no GPU; the orchestration
code under test (microclimate.jl_b200/distributed.py) is exactly what runs one-process-per-GPU with SoilEngine."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, nd, path):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from hostemu_engine import HostEmuEngine
    from microclimate_jl_b200 import distributed as D, marshal as M, synthetic as syn
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    sites, forcing = syn.make_sites(0, S, S), syn.make_forcing(0, S, S, 0, nd, threads=1)
    f = ("soil_temperature", "soil_moisture", "surface_water")
    local, (lo, hi) = D.run_shard(HostEmuEngine, desc, sites, forcing, rank, world, fields=f, chunk_days=1)
    full = D.gather_outputs(local, S)
    if rank == 0:
        np.savez(path, T=full.soil_temperature, th=full.soil_moisture, pool=full.surface_water, status=full.status)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_process():
    from hostemu_engine import HostEmuEngine
    import parity
    from microclimate_jl_b200 import marshal as M, synthetic as syn
    from microclimate_jl_b200.simulation import site_shards
    S, nd = 13, 2          # ragged: 7 + 6 sites
    assert site_shards(S, 2) == [(0, 7), (7, 13)] and site_shards(5, 8)[-1] == (5, 5)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.npz")
        mp.spawn(_worker, args=(2, _free_port(), S, nd, path), nprocs=2, join=True)
        got = np.load(path)
        desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
        ref, _ = parity.run(HostEmuEngine, desc, syn.make_sites(0, S, S), syn.make_forcing(0, S, S, 0, nd, threads=1),
                            fields=("soil_temperature", "soil_moisture", "surface_water"))
        assert np.array_equal(got["T"], ref.soil_temperature)
        assert np.array_equal(got["th"], ref.soil_moisture)
        assert np.array_equal(got["pool"], ref.surface_water)
        assert np.array_equal(got["status"], ref.status)


def _snow_air_problem(S=3):
    """Config-1-with-snow model (8 snow nodes) with four air heights, replicated over S sites whose air temperature is shifted site by site."""
    import copy
    import configs
    from microclimate_jl_b200 import marshal as M
    p, _ = configs.config1_monthly(snow=True)
    p = copy.deepcopy(p)
    p.model.heights = np.array([0.002, 0.01, 0.5, 2.0])
    inputs = []
    for s in range(S):
        i = copy.deepcopy(p.inputs)
        i.environment_hourly.reference_temperature = i.environment_hourly.reference_temperature - 1.5 * s
        inputs.append(i)
    desc = M.model_desc(p.model, p.time_mode, inputs0=inputs[0])
    return desc, M.SitesBatch.from_inputs(p.model, inputs), M.ForcingBatch.from_inputs(p.model, inputs, p.days[:2])


def _worker_snow_air(rank, world, port, path):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import parity
    from hostemu_engine import HostEmuEngine
    from microclimate_jl_b200 import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc, sites, forcing = _snow_air_problem()
    f = parity.ALL_FIELDS + parity.SNOW_FIELDS + parity.AIR_FIELDS
    local, _ = D.run_shard(HostEmuEngine, desc, sites, forcing, rank, world, fields=f, chunk_days=1)
    full = D.gather_outputs(local, sites.S)
    if rank == 0:
        np.savez(path, **{k: getattr(full, k) for k in f}, status=full.status)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_carries_snow_and_air_profile_outputs():
    """the host gather used to drop snow_temperature and the solve_air! fields; every selected array must arrive on rank 0."""
    from hostemu_engine import HostEmuEngine
    import parity
    desc, sites, forcing = _snow_air_problem()
    f = parity.ALL_FIELDS + parity.SNOW_FIELDS + parity.AIR_FIELDS
    assert desc.n_snow == 8 and desc.n_heights == 4 and "snow_temperature" in f and "air_temperature" in f
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.npz")
        mp.spawn(_worker_snow_air, args=(2, _free_port(), path), nprocs=2, join=True)
        got = np.load(path)
        ref, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
        for k in f:
            assert got[k].shape == getattr(ref, k).shape, k
            assert np.array_equal(got[k], getattr(ref, k)), k
        assert np.array_equal(got["status"], ref.status)


def _worker_minmax(rank, world, port, path):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from hostemu_engine import WeatherApiEmu
    from microclimate_jl_b200 import distributed as D, weather as W
    from test_weather import _soil_case
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc, sites, rest, win, solar = _soil_case(S=9, D=2, seed=31)
    f = ("soil_temperature", "soil_moisture", "surface_water")
    local, _ = D.run_shard_minmax(WeatherApiEmu, desc, W.weather_desc(), sites, win, rest, solar, rank, world, fields=f, chunk_days=1)
    full = D.gather_outputs(local, sites.S)
    if rank == 0:
        np.savez(path, T=full.soil_temperature, th=full.soil_moisture, pool=full.surface_water, status=full.status)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_minmax_run_equals_single_process():
    """One process per GPU for a min/max-driven problem: each rank expands the weather of its shard and integrates it (mc_step_days_minmax, here the host emulation of
    the weather unit over the soil host emulation), no data-path collective, host gather — equal to the single-handle run."""
    from hostemu_engine import WeatherApiEmu
    from microclimate_jl_b200 import marshal as M, weather as W
    from test_weather import _soil_case
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.npz")
        mp.spawn(_worker_minmax, args=(2, _free_port(), path), nprocs=2, join=True)
        got = np.load(path)
        desc, sites, rest, win, solar = _soil_case(S=9, D=2, seed=31)
        eng = WeatherApiEmu(desc, 9)
        try:
            eng.set_sites(sites)
            ref = M.OutputBatch(2, desc.n_soil, 9, ("soil_temperature", "soil_moisture", "surface_water"))
            eng.step_days_minmax(0, W.weather_desc(), win, rest, ref, solar)
        finally:
            eng.close()
        assert np.array_equal(got["T"], ref.soil_temperature) and np.array_equal(got["th"], ref.soil_moisture)
        assert np.array_equal(got["pool"], ref.surface_water) and np.array_equal(got["status"], ref.status)
