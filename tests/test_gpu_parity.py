"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA engine, called through the C ABI
(include/microclimate_b200.h), against the CPU oracle on the same marshalled buffers.

Bar: the north-star tolerance (<= 1e-3 K, <= 1e-5 m3/m3) on EVERY entry.  Because both sides evaluate the same
IEEE operation sequence with the shared bit-reproducible math kernel (include/mc_math.h), the observed difference is
exactly zero and the tests additionally assert bit-identity — which is what makes parity checkable at all for a path
this ill-conditioned (tests/test_conditioning.py).
"""
import numpy as np
import pytest

import configs
import parity
from microclimate_jl_b200 import marshal as M
from microclimate_jl_b200 import synthetic as syn
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def SoilEngine():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from microclimate_jl_b200.simulation import SoilEngine as E
    return E


def _problem_buffers(problem):
    desc = M.model_desc(problem.model, problem.time_mode, inputs0=problem.inputs)
    sites = M.SitesBatch.from_inputs(problem.model, [problem.inputs])
    forcing = M.ForcingBatch.from_inputs(problem.model, [problem.inputs], problem.days)
    return desc, sites, forcing


@pytest.mark.parametrize("mk", [configs.config1_monthly, configs.config2_daily], ids=["config1_monthly", "config2_daily"])
def test_reference_configs_match_oracle(SoilEngine, mk):
    """BASELINE.json configs 1 and 2 (the reference's own CI cases): every output field, every hour."""
    p, _ = mk()
    desc, sites, forcing = _problem_buffers(p)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, eg = parity.run(SoilEngine, desc, sites, forcing)
    parity.assert_within_north_star(out, ref)
    parity.assert_bit_identical(out, ref)
    co, cg = eo.counters(), eg.counters()
    for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes"):
        assert co[k] == cg[k], k
    assert cg["n_kernel_launches"] == 1


def test_reference_goldens_through_cuda(SoilEngine):
    """The reference's own assertions (test/micro_testrun_monthly.jl:172-182, micro_testrun_daily.jl:206-209) on the CUDA output."""
    p, g1 = configs.config1_monthly()
    out, _ = configs.run_engine(SoilEngine, p)
    T = out.soil_temperature[:, :, 0]
    assert np.abs(T / g1["soil_nmr"] - 1).max() < 1e-1 and np.abs(T - g1["soil_nmr"]).max() < 1.5
    assert np.abs(out.sky_temperature[:, 0] / g1["TSKYC"] - 1).max() < 2e-3
    p, g2 = configs.config2_daily()
    out, _ = configs.run_engine(SoilEngine, p)
    ci = np.arange(0, 19, 2)
    T, th = out.soil_temperature[:, ci, 0], out.soil_moisture[:, ci, 0]
    assert np.abs(T[:, :3] / g2["soil_nmr"][:, :3] - 1).max() < 1e-1
    assert np.sqrt(((th - g2["soilmoist_nmr"]) ** 2).mean()) < 0.03
    assert np.all(out.soil_moisture[1:, 18, 0] == 1.0 - 1.3 / 2.56)


@pytest.fixture(scope="module")
def synth():
    S, nd = 512, 3
    model = syn.synthetic_model()
    desc = M.model_desc(model, syn.synthetic_time_mode())
    return desc, syn.make_sites(0, S, S, with_sand=True), syn.make_forcing(0, S, S, 0, nd)


def test_synthetic_grid_matches_oracle(SoilEngine, synth):
    """512 synthetic sites (all texture classes incl. the non-converging dry-sand regime) x 3 days."""
    desc, sites, forcing = synth
    ref, _ = parity.run(OracleEngine, desc, sites, forcing)
    out, _ = parity.run(SoilEngine, desc, sites, forcing)
    parity.assert_within_north_star(out, ref)
    parity.assert_bit_identical(out, ref)


def test_time_chunking_and_site_sharding_are_invariant(SoilEngine, synth):
    desc, sites, forcing = synth
    whole, _ = parity.run(SoilEngine, desc, sites, forcing, fields=("soil_temperature", "soil_moisture", "surface_water"))
    chunked, _ = parity.run(SoilEngine, desc, sites, forcing, fields=("soil_temperature", "soil_moisture", "surface_water"), chunks=[1, 2])
    parity.assert_bit_identical(chunked, whole, ("soil_temperature", "soil_moisture", "surface_water"))
    lo, hi = 100, 357          # a ragged shard: not a multiple of the block or warp size
    part, _ = parity.run(SoilEngine, desc, sites.slice(lo, hi), forcing.site_slice(lo, hi), fields=("soil_temperature", "soil_moisture"))
    assert np.array_equal(part.soil_temperature, whole.soil_temperature[:, :, lo:hi])
    assert np.array_equal(part.soil_moisture, whole.soil_moisture[:, :, lo:hi])


def test_state_checkpoint_roundtrip(SoilEngine, synth):
    desc, sites, forcing = synth
    f = ("soil_temperature", "soil_moisture")
    a = SoilEngine(desc, sites.S); a.set_sites(sites)
    o1 = M.OutputBatch(1, desc.n_soil, sites.S, f); a.step_days(0, forcing.day_slice(0, 1), o1)
    st = a.get_state()
    assert st.days_done[0] == 1 and np.array_equal(st.status, o1.status)
    st.status[::5] |= 4                       # the status word travels with the checkpoint: bits set before the resume survive it
    b = SoilEngine(desc, sites.S); b.set_sites(sites); b.set_state(st)
    oa = M.OutputBatch(1, desc.n_soil, sites.S, f); ob = M.OutputBatch(1, desc.n_soil, sites.S, f)
    a.step_days(1, forcing.day_slice(1, 2), oa); b.step_days(1, forcing.day_slice(1, 2), ob)
    for k in f:                                # (parity.assert_bit_identical would also compare the status words, which differ here by construction)
        assert np.array_equal(getattr(oa, k), getattr(ob, k)), k
    assert np.all(ob.status[::5] & 4) and np.array_equal(ob.status & ~4, oa.status & ~4)


def test_error_behaviour(SoilEngine, synth):
    desc, sites, forcing = synth
    e = SoilEngine(desc, sites.S)
    out = M.OutputBatch(1, desc.n_soil, sites.S, ("soil_temperature",))
    with pytest.raises(RuntimeError, match="mc_set_sites"):
        e.step_days(0, forcing.day_slice(0, 1), out)
    e.set_sites(sites)
    with pytest.raises(RuntimeError, match="day0"):
        e.step_days(3, forcing.day_slice(0, 1), out)
    import copy
    bad = copy.copy(desc); bad.n_soil = 2
    with pytest.raises(RuntimeError, match="n_soil"):
        SoilEngine(bad, 4)


def test_fp64_peak_is_measured(SoilEngine, synth):
    desc, sites, _ = synth
    e = SoilEngine(desc, sites.S)
    tf = e.measure_fp64_peak()
    assert 10.0 < tf < 60.0     # B200: 148 SM x 64 FP64 lanes x 2 x ~1.9 GHz ~ 36-37 TFLOP/s


def test_snow_model_matches_oracle(SoilEngine):
    """BASELINE config 3: SnowModel (8 snow nodes on top of the soil column) + each apparent_heat_capacity closure, and a
    config-4-shaped problem (consecutive days, dynamic moisture, snow).  Every output incl. the snow fields, bit for bit."""
    import copy
    import dataclasses
    import microclimate_jl_b200 as mc
    f = parity.ALL_FIELDS + ("snow_fall", "snow_depth", "snow_density", "snow_temperature")
    p, g = configs.config1_monthly(snow=True)
    for ahc in (mc.BonacinaStep(), mc.TanhSmoothed(), mc.Gaussian(), mc.WestermannSigmoid()):
        q = copy.deepcopy(p)
        q.model.snow_model = dataclasses.replace(q.model.snow_model, apparent_heat_capacity=ahc)
        desc, sites, forcing = _problem_buffers(q)
        ref, eo = parity.run(OracleEngine, desc, sites, forcing, fields=f)
        out, eg = parity.run(SoilEngine, desc, sites, forcing, fields=f)
        parity.assert_within_north_star(out, ref, f)
        parity.assert_bit_identical(out, ref, f)
        assert eo.counters()["n_rhs"] == eg.counters()["n_rhs"]
        if isinstance(ahc, mc.BonacinaStep):        # the reference's own assertions on the CUDA output (micro_testrun_monthly_snow.jl:194-196)
            assert np.allclose(out.snow_fall[:, 0], g["SNOWFALL"], rtol=1e-3, atol=1e-9)
            assert np.allclose(out.snow_depth[:72, 0], g["SNOWDEP"][:72], rtol=1e-2)
            assert np.allclose(out.snow_density[:72, 0], g["SNOWDENS"][:72], rtol=1e-2)
    p2, _ = configs.config2_daily(6)
    p2 = copy.deepcopy(p2)
    p2.model.snow_model = mc.SnowModel(density_function=(0.5979, 0.2178, 0.001, 0.0038))
    h = p2.inputs.environment_hourly
    h.reference_temperature = np.asarray(h.reference_temperature) - 12.0
    p2.inputs.environment_daily.rainfall = np.asarray(p2.inputs.environment_daily.rainfall) + 4.0
    desc, sites, forcing = _problem_buffers(p2)
    ref, _ = parity.run(OracleEngine, desc, sites, forcing, fields=f)
    out, _ = parity.run(SoilEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(out, ref, f)
    st = SoilEngine(desc, 1)
    st.set_sites(sites)
    o1 = M.OutputBatch(2, desc.n_soil, 1, f, desc.n_snow)
    st.step_days(0, forcing.day_slice(0, 2), o1)
    state = st.get_state()
    assert state.snow_state is not None and state.snow_state[0, 0] == o1.snow_depth[-1, 0]      # current depth is carried


def test_async_pipeline_and_multi_handle_match_single_call(SoilEngine, synth):
    """mc_step_days_async / mc_wait (H2D, kernel and D2H of neighbouring day chunks overlap on three streams over two staging sets) and the
    multi-device handle (sites of ONE set of host arrays sharded over devices, here three ragged shards; on a one-GPU box all on device 0, which
    also exercises handles with different site counts sharing a device) must reproduce the single synchronous call bit for bit."""
    import torch
    from microclimate_jl_b200.simulation import MultiSoilEngine
    desc, sites, forcing = synth
    f = ("soil_temperature", "soil_moisture", "surface_water", "sky_temperature")
    whole, _ = parity.run(SoilEngine, desc, sites, forcing, fields=f)
    # three one-day chunks enqueued back to back, one wait at the end
    eng = SoilEngine(desc, sites.S); eng.set_sites(sites)
    parts = [M.OutputBatch(1, desc.n_soil, sites.S, f) for _ in range(3)]
    fparts = [forcing.day_slice(d, d + 1) for d in range(3)]
    for d in range(3):
        eng.step_days_async(d, fparts[d], parts[d])
    nbad = eng.wait()
    piped = M.OutputBatch(3, desc.n_soil, sites.S, f)
    for d in range(3):
        piped.assign_rows(24 * d, parts[d])
    parity.assert_bit_identical(piped, whole, f)
    assert nbad == int((whole.status != 0).sum())
    assert eng.counters()["n_site_days"] == 3 * sites.S
    # multi-device handle: as many GPUs as the box has, at least three shards
    ndev = torch.cuda.device_count()
    devices = [g % ndev for g in range(max(3, ndev))]
    with MultiSoilEngine(desc, 509, devices) as me:
        assert me.ranges()[0][0] == 0 and me.ranges()[-1][1] == 509
        me.set_sites(sites.slice(0, 509))
        out = M.OutputBatch(3, desc.n_soil, 509, f)
        me.step_days(0, forcing.site_slice(0, 509), out)
        for name in f:
            assert np.array_equal(getattr(out, name), getattr(whole, name)[..., :509]), name
        assert np.array_equal(out.status, whole.status[:509])
        assert me.counters()["n_site_days"] == 3 * 509


def test_handles_with_different_models_share_a_device(SoilEngine):
    """Two handles with DIFFERENT models (N = 10 prescribed-moisture vs N = 19 dynamic) alternate on one device through the asynchronous device-pointer
    entry: the per-device constant bank must never be overwritten under a running kernel."""
    import torch
    p1, _ = configs.config1_monthly()
    p2, _ = configs.config2_daily(3)
    d1, s1, f1 = _problem_buffers(p1)
    d2, s2, f2 = _problem_buffers(p2)
    ref1, _ = parity.run(OracleEngine, d1, s1, f1, fields=("soil_temperature",))
    ref2, _ = parity.run(OracleEngine, d2, s2, f2, fields=("soil_temperature", "soil_moisture"))
    e1 = SoilEngine(d1, 1); e1.set_sites(s1)
    e2 = SoilEngine(d2, 1); e2.set_sites(s2)
    o1 = [M.OutputBatch(1, d1.n_soil, 1, ("soil_temperature",)) for _ in range(f1.ndays)]
    o2 = [M.OutputBatch(1, d2.n_soil, 1, ("soil_temperature", "soil_moisture")) for _ in range(f2.ndays)]
    k1 = [f1.day_slice(d, d + 1) for d in range(f1.ndays)]
    k2 = [f2.day_slice(d, d + 1) for d in range(f2.ndays)]
    for d in range(max(f1.ndays, f2.ndays)):         # interleave asynchronous chunks of the two models
        if d < f1.ndays:
            e1.step_days_async(d, k1[d], o1[d])
        if d < f2.ndays:
            e2.step_days_async(d, k2[d], o2[d])
    e1.wait(); e2.wait()
    got1 = np.concatenate([o.soil_temperature for o in o1]); got2 = np.concatenate([o.soil_temperature for o in o2])
    assert np.array_equal(got1, ref1.soil_temperature) and np.array_equal(got2, ref2.soil_temperature)
    assert np.array_equal(np.concatenate([o.soil_moisture for o in o2]), ref2.soil_moisture)


@pytest.mark.parametrize("n_nodes", [7, 24], ids=["N7", "N24"])
def test_other_node_counts_match_oracle(SoilEngine, n_nodes):
    """Node counts below / between the kernels' compile-time array bounds (NA = 10 / 19 / 24)."""
    p, _ = configs.config2_daily(3)
    depths = np.concatenate([[0.0], np.geomspace(0.01, 2.0, n_nodes - 1)])
    desc, sites, forcing = _problem_buffers(configs.regrid(p, depths))
    ref, _ = parity.run(OracleEngine, desc, sites, forcing)
    out, _ = parity.run(SoilEngine, desc, sites, forcing)
    parity.assert_within_north_star(out, ref)
    parity.assert_bit_identical(out, ref)


def test_capped_newton_solves_match_oracle(SoilEngine):
    """Sandy sites whose moisture Newton solve runs to the 500-iteration cap (1260, 1331, 1353 of the 4096-site grid): the kernel skips whole
    periods of exactly periodic orbits; outputs, status flags and iteration counts must be those of the oracle's full loop."""
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    sites = syn.make_sites(1256, 1360, 4096, with_sand=True)
    forcing = syn.make_forcing(1256, 1360, 4096, 0, 3)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, eg = parity.run(SoilEngine, desc, sites, forcing)
    assert (ref.status != 0).any()
    parity.assert_bit_identical(out, ref)
    assert np.array_equal(out.status, ref.status)
    assert eo.counters()["n_newton_iter"] == eg.counters()["n_newton_iter"]


def test_cooperative_kernel_block_shapes(SoilEngine, synth, monkeypatch):
    """The block-cooperative kernel must not depend on how sites fall into blocks: full-machine blocks of 1024 threads with a ragged last
    block, a grid-stride loop over several batches per block, and the legacy per-site kernel all give the oracle's bits."""
    desc, sites, forcing = synth
    f = ("soil_temperature", "soil_moisture", "surface_water", "soil_water_potential")
    ref, eo = parity.run(OracleEngine, desc, sites, forcing, fields=f)
    monkeypatch.setenv("MC_COOP_BLOCK", "1024")                 # one ragged block: 512 sites + 512 mirror threads
    a, ea = parity.run(SoilEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(a, ref, f)
    assert ea.counters()["n_rhs"] == eo.counters()["n_rhs"] and ea.counters()["n_newton_iter"] == eo.counters()["n_newton_iter"]
    monkeypatch.setenv("MC_COOP_BLOCK", "96")                   # 6 batches ...
    monkeypatch.setenv("MC_COOP_GRID", "2")                     # ... over 2 blocks: grid-stride reuse of the block scratch, ragged tail
    b, eb = parity.run(SoilEngine, desc, sites.slice(0, 509), forcing.site_slice(0, 509), fields=f)
    for name in f:
        assert np.array_equal(getattr(b, name), getattr(ref, name)[..., :509]), name
    monkeypatch.delenv("MC_COOP_BLOCK"); monkeypatch.delenv("MC_COOP_GRID")
    monkeypatch.setenv("MC_KERNEL", "legacy")                   # the per-site kernel (what snow / flux-potential models run)
    c, ec = parity.run(SoilEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(c, ref, f)
    monkeypatch.delenv("MC_KERNEL")
    # the opt-in asynchronous kernel under odd shapes: 3 blocks x 5 slots x 96 threads (passes claimed in pieces, slots reloaded many times), one full block, defaults
    monkeypatch.setenv("MC_KERNEL", "async")
    for grid, slots, threads in (("3", "5", "96"), ("1", "64", "1024"), ("148", "64", "1024")):
        monkeypatch.setenv("MC_ASYNC_GRID", grid); monkeypatch.setenv("MC_ASYNC_SLOTS", slots); monkeypatch.setenv("MC_ASYNC_THREADS", threads)
        d, ed = parity.run(SoilEngine, desc, sites.slice(0, 203), forcing.site_slice(0, 203), fields=f)
        for name in f:
            assert np.array_equal(getattr(d, name), getattr(ref, name)[..., :203]), name
    monkeypatch.delenv("MC_ASYNC_GRID"); monkeypatch.delenv("MC_ASYNC_SLOTS"); monkeypatch.delenv("MC_ASYNC_THREADS"); monkeypatch.delenv("MC_KERNEL")


def test_solve_air_matches_oracle(SoilEngine):
    """solve_air! (the first "next" row of the scope table): MOST profiles from the final surface temperature, snow surface included."""
    import copy
    p, g = configs.config1_monthly(snow=True)
    p = copy.deepcopy(p)
    p.model.heights = np.array([0.002, 0.01, 0.5, 2.0])
    f = parity.ALL_FIELDS + parity.SNOW_FIELDS + parity.AIR_FIELDS
    desc, sites, forcing = _problem_buffers(p)
    ref, _ = parity.run(OracleEngine, desc, sites, forcing, fields=f)
    out, _ = parity.run(SoilEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(out, ref, f)
    chunked, _ = parity.run(SoilEngine, desc, sites, forcing, fields=f, chunks=[7, 5])
    parity.assert_bit_identical(chunked, ref, parity.AIR_FIELDS)
    # without soil_temperature selected the call must refuse rather than read garbage
    bad = M.OutputBatch(12, desc.n_soil, 1, parity.AIR_FIELDS, desc.n_snow, desc.n_heights)
    e = SoilEngine(desc, 1); e.set_sites(sites)
    with pytest.raises(RuntimeError, match="solve_air"):
        e.step_days(0, forcing, bad)


# ---- strategy variants on the B200 (scope rows a8 / a12 / a14): the same cases the CPU tier runs through the host emulation ----
def _variant(problem, **cfg):
    import copy
    p = copy.deepcopy(problem)
    for k, v in cfg.items():
        if k == "time_mode":
            p.time_mode = v
        elif k == "hydraulics":
            p.model.soil_hydraulic_model = v
        else:
            setattr(p.model.config, k, v)
    return p


def _assert_gpu_equals_oracle(SoilEngine, p, fields=parity.ALL_FIELDS):
    desc, sites, forcing = _problem_buffers(p)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing, fields=fields)
    out, eg = parity.run(SoilEngine, desc, sites, forcing, fields=fields)
    parity.assert_within_north_star(out, ref, fields)
    parity.assert_bit_identical(out, ref, fields)
    co, cg = eo.counters(), eg.counters()
    for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes"):
        assert co[k] == cg[k], k
    return ref, out


def test_tolerance_convergence_matches_oracle(SoilEngine):
    """SoilTemperatureConvergenceTolerance (src/soil_temperature_convergence/tolerance.jl:8-22): the day-pass count differs per day
    (and, in a batch, per site) - the per-lane `is_last` divergence of survey H-c."""
    import microclimate_jl_b200 as mc
    p1, _ = configs.config1_monthly()
    p = _variant(p1, convergence=mc.SoilTemperatureConvergenceTolerance(tolerance=0.5, max_iterations_per_day=6),
                 time_mode=mc.NonConsecutiveDayMode(iterations_per_day=6))
    ref, _ = _assert_gpu_equals_oracle(SoilEngine, p)
    desc, sites, forcing = _problem_buffers(p)
    _, eo = parity.run(OracleEngine, desc, sites, forcing)
    assert 12 < eo.counters()["n_day_passes"] < 72          # neither the minimum nor the cap on every day: the tolerance decides


def test_spinup_first_day_matches_oracle(SoilEngine):
    """ConsecutiveDayMode(spinup_first_day=true) (src/time_mode/consecutive.jl:9-23): day 1 repeats `iterations` passes."""
    import microclimate_jl_b200 as mc
    p2, _ = configs.config2_daily(4)
    p = _variant(p2, time_mode=mc.ConsecutiveDayMode(spinup_first_day=True), convergence=mc.FixedSoilTemperatureIterations(3))
    _assert_gpu_equals_oracle(SoilEngine, p)


def test_matric_flux_potential_infiltration_matches_oracle(SoilEngine):
    """MatricFluxPotentialAlgorithm (src/soil_hydraulics/infiltration_algorithm/matric_flux_potential.jl:19-101)."""
    import microclimate_jl_b200 as mc
    p2, _ = configs.config2_daily(4)
    p = _variant(p2, hydraulics=mc.CampbellSoilHydraulics(infiltration_algorithm=mc.MatricFluxPotentialAlgorithm()))
    _assert_gpu_equals_oracle(SoilEngine, p)


def test_swinbank_and_supplied_longwave_match_oracle(SoilEngine):
    """SwinbankAtmosphericRadiation (src/radiation/atmospheric/swinbank.jl:9-11) and user-supplied longwave
    (src/radiation/longwave/view_factor.jl:23-67, `longwave_radiation` hourly forcing)."""
    import copy
    import microclimate_jl_b200 as mc
    p1, _ = configs.config1_monthly()
    p_sw = copy.deepcopy(p1)
    p_sw.model.radiation = mc.RadiationModel(longwave_model=mc.ViewFactorLongwave(atmospheric_radiation_model=mc.SwinbankAtmosphericRadiation()))
    ref_sw, _ = _assert_gpu_equals_oracle(SoilEngine, p_sw)
    ref_cn, _ = _assert_gpu_equals_oracle(SoilEngine, p1)
    assert not np.array_equal(ref_sw.soil_temperature, ref_cn.soil_temperature)       # the strategy really changes the answer
    p_lw = copy.deepcopy(p1)
    h = p_lw.inputs.environment_hourly
    h.longwave_radiation = 300.0 + 40.0 * np.cos(np.arange(len(h.zenith_angle)) * (2 * np.pi / 24.0))
    ref_lw, _ = _assert_gpu_equals_oracle(SoilEngine, p_lw)
    assert not np.array_equal(ref_lw.sky_temperature, ref_cn.sky_temperature)


def test_sloped_site_matches_oracle(SoilEngine):
    """Sloped site with a reduced sky view (src/soil_balance/energy/heat_transport_1d.jl:223-232): slope-zenith and diffuse fraction forcing."""
    import copy
    p, _ = configs.config1_monthly()
    p = copy.deepcopy(p)
    p.inputs.site.slope = 20.0
    p.inputs.site.sky_view_fraction = 0.8
    h = p.inputs.environment_hourly
    h.diffuse_fraction = np.full(len(h.zenith_angle), 0.3)
    h.zenith_slope_angle = np.minimum(np.asarray(h.zenith_angle) + 10.0, 90.0)
    _assert_gpu_equals_oracle(SoilEngine, p)
    # and a dynamic-moisture, consecutive-day sloped site (the block-cooperative kernel's slope branch)
    p2, _ = configs.config2_daily(3)
    p2 = copy.deepcopy(p2)
    p2.inputs.site.slope = 15.0
    p2.inputs.site.sky_view_fraction = 0.9
    h = p2.inputs.environment_hourly
    h.diffuse_fraction = np.full(len(h.zenith_angle), 0.25)
    h.zenith_slope_angle = np.minimum(np.asarray(h.zenith_angle) + 8.0, 90.0)
    _assert_gpu_equals_oracle(SoilEngine, p2)


def test_file_driven_run_matches_in_memory(SoilEngine, tmp_path):
    """On-disk formats (scope row f4): site / forcing containers memory-mapped into mc_step_days, chunked outputs and a state checkpoint written per
    chunk (tools/dataset.py) - the file-driven run, interrupted after the first chunk and resumed from its checkpoint, must equal the in-memory run."""
    import os
    import subprocess
    import sys
    from microclimate_jl_b200 import storage as st
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    S, ND = 1536, 3
    d = str(tmp_path)
    subprocess.check_call([sys.executable, "tools/dataset.py", "make", d, "--sites", str(S), "--days", str(ND), "--chunk", "1"], cwd=root)
    chunks = sorted(p for p in os.listdir(d) if p.startswith("forcing_d"))
    assert len(chunks) == ND
    hidden = [os.path.join(d, c) for c in chunks[1:]]
    for p in hidden:                                   # first invocation only sees day 0 ...
        os.rename(p, p + ".later")
    subprocess.check_call([sys.executable, "tools/dataset.py", "run", d], cwd=root)
    for p in hidden:                                   # ... the second resumes from the checkpoint for days 1-2
        os.rename(p + ".later", p)
    subprocess.check_call([sys.executable, "tools/dataset.py", "run", d, "--resume"], cwd=root)
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    f = ("soil_temperature", "soil_moisture", "surface_water")
    ref, _ = parity.run(SoilEngine, desc, syn.make_sites(0, S, S), syn.make_forcing(0, S, S, 0, ND), fields=f)
    for name in f:
        got = np.concatenate([getattr(st.load_outputs(os.path.join(d, f"out_d{k:04d}_s0-{S}.mcb")), name) for k in range(ND)])
        assert np.array_equal(got, getattr(ref, name)), name
    assert int(st.load_state(os.path.join(d, f"state_s0-{S}.mcb")).days_done[0]) == ND


def test_config4_scan329_full_year_matches_oracle(SoilEngine):
    """BASELINE config 4 (SCAN / SNOTEL 329, test/micro_testrun_snow_scan329.jl:1-238): N=19 + 8 snow nodes, 365 consecutive days, dynamic moisture,
    DailyRainfall, snow - every output field incl. the snow fields over the whole year, bit for bit, in 73-day chunks (the snowpack is carried state)."""
    p, g = configs.config4_scan329()
    desc, sites, forcing = _problem_buffers(p)
    f = parity.ALL_FIELDS + parity.SNOW_FIELDS
    ref, eo = parity.run(OracleEngine, desc, sites, forcing, fields=f, chunks=[73] * 5)
    out, eg = parity.run(SoilEngine, desc, sites, forcing, fields=f, chunks=[73] * 5)
    parity.assert_within_north_star(out, ref, f)
    parity.assert_bit_identical(out, ref, f)
    co, cg = eo.counters(), eg.counters()
    for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes"):
        assert co[k] == cg[k], k
    # the reference's usual NicheMapR bar on the CUDA output itself
    T, G = out.soil_temperature[:, ::2, 0], g["soil_nmr"].astype(float)
    assert np.abs(T / G - 1).max() < 1e-1 and np.corrcoef(out.snow_depth[:, 0], g["SNOWDEP"].astype(float))[0, 1] > 0.85


@pytest.mark.parametrize("case", ["config1_monthly", "config2_daily", "tolerance", "synthetic"])
def test_async_kernel_matches_oracle(SoilEngine, case, monkeypatch, synth):
    """The opt-in asynchronous shared-memory-resident kernel (MC_KERNEL=async, mc_async.cuh): per-site state machines, node items and single-lane tasks
    claimed by worker warps.  Same bits and counters as the oracle, incl. the per-site pass counts of the tolerance-based convergence."""
    import microclimate_jl_b200 as mc
    monkeypatch.setenv("MC_KERNEL", "async")
    if case == "synthetic":
        desc, sites, forcing = synth
    else:
        p = {"config1_monthly": lambda: configs.config1_monthly()[0], "config2_daily": lambda: configs.config2_daily(6)[0],
             "tolerance": lambda: _variant(configs.config1_monthly()[0], convergence=mc.SoilTemperatureConvergenceTolerance(tolerance=0.5, max_iterations_per_day=6),
                                           time_mode=mc.NonConsecutiveDayMode(iterations_per_day=6))}[case]()
        desc, sites, forcing = _problem_buffers(p)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, eg = parity.run(SoilEngine, desc, sites, forcing)
    parity.assert_bit_identical(out, ref)
    co, cg = eo.counters(), eg.counters()
    for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes"):
        assert co[k] == cg[k], k
