"""CPU tier: the product's per-site kernel body (microclimate.jl_b200/csrc/mc_device.cuh), compiled for the host by the
tests/hostemu harness, against the oracle.  This is what lets the no-GPU tier check the code the CUDA kernel executes;
the GPU tier (test_gpu_parity.py) repeats the same comparisons on the device through the C ABI."""
import copy
import dataclasses

import numpy as np
import pytest

import configs
import parity
import microclimate_jl_b200 as mc
from hostemu_engine import HostEmuEngine
from microclimate_jl_b200 import marshal as M
from microclimate_jl_b200 import synthetic as syn
from oracle.oracle import OracleEngine


def _buffers(problem):
    desc = M.model_desc(problem.model, problem.time_mode, inputs0=problem.inputs)
    return desc, M.SitesBatch.from_inputs(problem.model, [problem.inputs]), M.ForcingBatch.from_inputs(problem.model, [problem.inputs], problem.days)


@pytest.mark.parametrize("mk", [configs.config1_monthly, configs.config2_daily], ids=["config1_monthly", "config2_daily"])
def test_reference_configs_bit_identical(mk):
    p, _ = mk()
    desc, sites, forcing = _buffers(p)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, ee = parity.run(HostEmuEngine, desc, sites, forcing)
    parity.assert_within_north_star(out, ref)
    parity.assert_bit_identical(out, ref)
    co, ce = eo.counters(), ee.counters()
    assert all(co[k] == ce[k] for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes"))


@pytest.fixture(scope="module")
def synth():
    S, nd = 48, 3
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    return desc, syn.make_sites(0, S, S, with_sand=True), syn.make_forcing(0, S, S, 0, nd)


def test_synthetic_with_sand_bit_identical(synth):
    desc, sites, forcing = synth
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, ee = parity.run(HostEmuEngine, desc, sites, forcing)
    parity.assert_bit_identical(out, ref)
    assert np.all((out.soil_moisture >= 0) & (out.soil_moisture <= 1)) and np.all(np.isfinite(out.soil_temperature))
    assert np.array_equal(out.status, ref.status)
    assert eo.counters()["n_newton_iter"] == ee.counters()["n_newton_iter"]


def test_capped_newton_solves_bit_identical():
    """Sites whose moisture Newton solve runs to the 500-iteration cap (sandy sites 1260, 1331, 1353 of the 4096-site grid from day 1 on):
    the kernel skips whole periods of an exactly periodic orbit; outputs, status flags and iteration counts must be the full loop's."""
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    sites = syn.make_sites(1256, 1360, 4096, with_sand=True)
    forcing = syn.make_forcing(1256, 1360, 4096, 0, 3)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, ee = parity.run(HostEmuEngine, desc, sites, forcing)
    assert (ref.status & mc.STATUS_NEWTON_MAXITER).any() if hasattr(mc, "STATUS_NEWTON_MAXITER") else (ref.status != 0).any()
    parity.assert_bit_identical(out, ref)
    assert np.array_equal(out.status, ref.status)
    co, ce = eo.counters(), ee.counters()
    assert all(co[k] == ce[k] for k in ("n_rhs", "n_accept", "n_reject", "n_newton_iter", "n_infil_steps"))


def test_chunking_and_sharding_invariance(synth):
    desc, sites, forcing = synth
    f = ("soil_temperature", "soil_moisture", "surface_water", "soil_water_potential")
    whole, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
    chunked, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f, chunks=[2, 1])
    parity.assert_bit_identical(chunked, whole, f)
    ref_chunked, _ = parity.run(OracleEngine, desc, sites, forcing, fields=f, chunks=[2, 1])
    parity.assert_bit_identical(chunked, ref_chunked, f)
    part, _ = parity.run(HostEmuEngine, desc, sites.slice(7, 29), forcing.site_slice(7, 29), fields=f)
    assert np.array_equal(part.soil_temperature, whole.soil_temperature[:, :, 7:29])
    assert np.array_equal(part.soil_moisture, whole.soil_moisture[:, :, 7:29])


def test_cooperative_moisture_option_bit_identical(synth):
    """MC_COOP_MOIST=1 (block-cooperative soil water balance; compiled out by default because it is slower on B200) must stay
    bit-identical: it relies on a sub-step's initial state being the previous sub-step's final state."""
    import hostemu_engine

    class CoopMoistEngine(HostEmuEngine):
        LIB = hostemu_engine.variant("coopmoist", ["-DMC_COOP_MOIST=1"])

    p, _ = configs.config2_daily(3)
    desc, sites, forcing = _buffers(p)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, ee = parity.run(CoopMoistEngine, desc, sites, forcing)
    parity.assert_bit_identical(out, ref)
    co, ce = eo.counters(), ee.counters()
    assert all(co[k] == ce[k] for k in ("n_rhs", "n_accept", "n_reject", "n_newton_iter", "n_infil_steps"))
    desc, sites, forcing = synth
    ref, _ = parity.run(OracleEngine, desc, sites, forcing)
    out, _ = parity.run(CoopMoistEngine, desc, sites, forcing)
    parity.assert_bit_identical(out, ref)


@pytest.mark.parametrize("n_nodes", [7, 22, 24], ids=["N7", "N22", "N24"])
def test_other_node_counts_bit_identical(n_nodes):
    """Node counts that are not an array bound of the kernel templates (NA = 10 / 19 / 24): ragged rows of the block scratch."""
    p, _ = configs.config2_daily(3)
    depths = np.concatenate([[0.0], np.geomspace(0.01, 2.0, n_nodes - 1)])
    desc, sites, forcing = _buffers(configs.regrid(p, depths))
    assert desc.n_soil == n_nodes
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    out, ee = parity.run(HostEmuEngine, desc, sites, forcing)
    parity.assert_bit_identical(out, ref)
    assert eo.counters()["n_rhs"] == ee.counters()["n_rhs"]


def test_stomatal_cycle_shortcut_equals_plain_loop():
    """The kernel resolves a stomatal Newton iteration that never meets |residual| <= 10 from its cycle instead of running all
    moisture_max_iterations; the value after the last iteration must be the plain loop's, bit for bit, whatever the orbit."""
    import ctypes as C
    import hostemu_engine
    L = hostemu_engine.lib()
    L.he_stomatal_newton.restype = C.c_double
    L.he_stomatal_newton.argtypes = [C.c_void_p] + [C.c_double] * 3
    L.he_stomatal_newton_plain.restype = C.c_double
    L.he_stomatal_newton_plain.argtypes = [C.c_void_p] + [C.c_double] * 3 + [C.POINTER(C.c_int)]
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    rng = np.random.default_rng(7)
    n, capped = 20000, 0
    psi_bar = -np.exp(rng.uniform(np.log(0.5), np.log(3000.0), n))          # J/kg
    R_bar = np.exp(rng.uniform(np.log(1e5), np.log(1e12), n))               # m4/kg/s
    TP = np.exp(rng.uniform(np.log(1e-12), np.log(2e-4), n))                # kg/m2/s
    for a, b, c in zip(psi_bar, R_bar, TP):
        it = C.c_int(0)
        plain = L.he_stomatal_newton_plain(C.byref(desc), a, b, c, C.byref(it))
        fast = L.he_stomatal_newton(C.byref(desc), a, b, c)
        capped += it.value >= desc.moisture_max_iterations
        assert (plain == fast) or (plain != plain and fast != fast), (a, b, c, plain, fast, it.value)
    assert capped > 100          # the sample must exercise non-terminating orbits


@pytest.mark.parametrize("defs", [(), ("-DMC_COOP_MOIST=1",)], ids=["default", "coop_moist"])
def test_block_protocol_with_real_threads(defs, monkeypatch):
    """The cooperative kernel's block protocol — request lists with counter parity, item passes over other threads' sites, role swaps,
    barriers, mirror threads of a ragged last block — run with 8 OS threads per block and a real barrier (hostemu -DMC_HOST_THREADS)."""
    import hostemu_engine

    class ThreadedEngine(HostEmuEngine):
        LIB = hostemu_engine.variant("threads" + ("_cm" if defs else ""), ["-DMC_HOST_THREADS", "-std=c++20", "-pthread", *defs])

    monkeypatch.setenv("HE_TB", "8")
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    sites, forcing = syn.make_sites(0, 21, 21, with_sand=True), syn.make_forcing(0, 21, 21, 0, 2)      # 3 batches, the last one 5 sites + 3 mirrors
    ref, eo = parity.run(OracleEngine, desc, sites, forcing)
    for grid in ("0", "1"):                      # one batch per block / one block striding over the three batches (counter rotation carried)
        monkeypatch.setenv("HE_GRID", grid)
        out, ee = parity.run(ThreadedEngine, desc, sites, forcing)
        parity.assert_bit_identical(out, ref)
        co, ce = eo.counters(), ee.counters()
        assert all(co[k] == ce[k] for k in ("n_rhs", "n_accept", "n_reject", "n_newton_iter", "n_infil_steps"))


@pytest.mark.parametrize("slots,lanes", [(6, 4), (3, 2), (64, 32)], ids=["6slots_4lanes", "3slots_2lanes", "64slots_32lanes"])
def test_async_protocol_with_real_threads(slots, lanes, monkeypatch):
    """The asynchronous kernel's claim / retire protocol (request masks, per-slot cursor words claimed by compare-and-swap, retire counters,
    sites handed out from a shared cursor) with real threads: every worker "warp" is an OS thread, several blocks run concurrently, slots and
    lanes are few so that passes are claimed in pieces by different warps.  Results and counters must be the oracle's whatever the interleaving."""
    import ctypes as C
    import hostemu_engine

    class AsyncEngine(HostEmuEngine):
        LIB = hostemu_engine.variant(f"async_{slots}_{lanes}", ["-DMC_HOST_THREADS", "-std=c++20", "-pthread", f"-DHE_SLOTS={slots}", f"-DMCA_LANES={lanes}"])

    AsyncEngine.LIB.he_async_runs.restype = C.c_longlong
    monkeypatch.setenv("MC_KERNEL", "async")      # opt-in: the default NoSnow kernel is the block-synchronous cooperative one
    desc = M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode())
    sites, forcing = syn.make_sites(0, 29, 29, with_sand=True), syn.make_forcing(0, 29, 29, 0, 2)
    ref, eo = parity.run(OracleEngine, desc, sites, forcing, chunks=[1, 1])
    before = AsyncEngine.LIB.he_async_runs()
    for warps, grid in (("5", "2"), ("2", "3"), ("7", "1")):
        monkeypatch.setenv("HE_WARPS", warps)
        monkeypatch.setenv("HE_GRID", grid)
        out, ee = parity.run(AsyncEngine, desc, sites, forcing, chunks=[1, 1])
        parity.assert_bit_identical(out, ref)
        co, ce = eo.counters(), ee.counters()
        assert all(co[k] == ce[k] for k in ("n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps", "n_day_passes"))
    assert AsyncEngine.LIB.he_async_runs() - before == 6          # the asynchronous body really ran (2 chunks x 3 configurations)
    # config 1: non-consecutive days, 3 passes per day, prescribed moisture; and the tolerance-based pass count (per-site, no block-uniform constraint)
    p1, _ = configs.config1_monthly()
    for p in (p1, _variant(p1, convergence=mc.SoilTemperatureConvergenceTolerance(tolerance=0.5, max_iterations_per_day=6), time_mode=mc.NonConsecutiveDayMode(iterations_per_day=6))):
        d1, s1, f1 = _buffers(p)
        ref1, eo1 = parity.run(OracleEngine, d1, s1, f1)
        out1, ee1 = parity.run(AsyncEngine, d1, s1, f1)
        parity.assert_bit_identical(out1, ref1)
        assert eo1.counters()["n_day_passes"] == ee1.counters()["n_day_passes"] and eo1.counters()["n_rhs"] == ee1.counters()["n_rhs"]


def _variant(problem, **cfg):
    p = copy.deepcopy(problem)
    for k, v in cfg.items():
        if k == "time_mode":
            p.time_mode = v
        elif k == "hydraulics":
            p.model.soil_hydraulic_model = v
        else:
            setattr(p.model.config, k, v)
    return p


def test_strategy_variants_bit_identical():
    """Every strategy object on the path: tolerance convergence, spin-up first day, flux-potential infiltration, Swinbank."""
    p1, _ = configs.config1_monthly()
    p2, _ = configs.config2_daily(4)
    variants = [
        _variant(p1, convergence=mc.SoilTemperatureConvergenceTolerance(tolerance=0.5, max_iterations_per_day=6),
                 time_mode=mc.NonConsecutiveDayMode(iterations_per_day=6)),
        _variant(p2, time_mode=mc.ConsecutiveDayMode(spinup_first_day=True), convergence=mc.FixedSoilTemperatureIterations(3)),
        _variant(p2, hydraulics=mc.CampbellSoilHydraulics(infiltration_algorithm=mc.MatricFluxPotentialAlgorithm())),
    ]
    p_sw = copy.deepcopy(p1)
    p_sw.model.radiation = mc.RadiationModel(longwave_model=mc.ViewFactorLongwave(atmospheric_radiation_model=mc.SwinbankAtmosphericRadiation()))
    variants.append(p_sw)
    for p in variants:
        desc, sites, forcing = _buffers(p)
        ref, eo = parity.run(OracleEngine, desc, sites, forcing)
        out, ee = parity.run(HostEmuEngine, desc, sites, forcing)
        parity.assert_bit_identical(out, ref)
        assert eo.counters()["n_day_passes"] == ee.counters()["n_day_passes"]


def test_sloped_site_bit_identical():
    p, _ = configs.config1_monthly()
    p = copy.deepcopy(p)
    p.inputs.site.slope = 20.0
    p.inputs.site.sky_view_fraction = 0.8
    h = p.inputs.environment_hourly
    h.diffuse_fraction = np.full(len(h.zenith_angle), 0.3)
    h.zenith_slope_angle = np.minimum(np.asarray(h.zenith_angle) + 10.0, 90.0)
    desc, sites, forcing = _buffers(p)
    ref, _ = parity.run(OracleEngine, desc, sites, forcing)
    out, _ = parity.run(HostEmuEngine, desc, sites, forcing)
    parity.assert_bit_identical(out, ref)


def test_snow_model_and_apparent_heat_capacity_variants_bit_identical():
    """SnowModel (config 3) with each apparent_heat_capacity closure (src/apparent_heat_capacity/*.jl), plus a chunked run
    (the snowpack state is part of the carried state)."""
    p, _ = configs.config1_monthly(snow=True)
    f = parity.ALL_FIELDS + ("snow_fall", "snow_depth", "snow_density", "snow_temperature")
    for ahc in (mc.BonacinaStep(), mc.TanhSmoothed(), mc.Gaussian(), mc.WestermannSigmoid()):
        q = copy.deepcopy(p)
        q.model.snow_model = dataclasses.replace(q.model.snow_model, apparent_heat_capacity=ahc)
        desc, sites, forcing = _buffers(q)
        assert desc.n_snow == 8
        ref, eo = parity.run(OracleEngine, desc, sites, forcing, fields=f)
        out, ee = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
        parity.assert_bit_identical(out, ref, f)
        assert eo.counters()["n_rhs"] == ee.counters()["n_rhs"]
        assert ref.snow_depth.max() > 10.0
    desc, sites, forcing = _buffers(p)
    whole, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
    chunked, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f, chunks=[5, 4, 3])
    parity.assert_bit_identical(chunked, whole, f)
    ref_chunked, _ = parity.run(OracleEngine, desc, sites, forcing, fields=f, chunks=[5, 4, 3])
    parity.assert_bit_identical(chunked, ref_chunked, f)


def test_snow_with_dynamic_moisture_consecutive_days_bit_identical():
    """Config-4-shaped problem (consecutive days, dynamic moisture, snow): config 2's site with a SnowModel and cold forcing."""
    p, _ = configs.config2_daily(6)
    p = copy.deepcopy(p)
    p.model.snow_model = mc.SnowModel(density_function=(0.5979, 0.2178, 0.001, 0.0038))
    h = p.inputs.environment_hourly
    h.reference_temperature = np.asarray(h.reference_temperature) - 12.0        # push the desert winter below freezing
    p.inputs.environment_daily.rainfall = np.asarray(p.inputs.environment_daily.rainfall) + 4.0
    desc, sites, forcing = _buffers(p)
    f = parity.ALL_FIELDS + ("snow_fall", "snow_depth", "snow_density", "snow_temperature")
    ref, _ = parity.run(OracleEngine, desc, sites, forcing, fields=f)
    out, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(out, ref, f)
    # chunked == whole for everything except soil_humidity's first row per chunk (upstream writes initialise_soil_humidity! to
    # row 1 of the output buffer only, src/soil_moisture/dynamic.jl:32-35; a chunk's buffer starts at its own row 1)
    g = tuple(x for x in f if x != "soil_humidity")
    chunked, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f, chunks=[2, 4])
    parity.assert_bit_identical(chunked, ref, g)
    assert np.array_equal(np.delete(chunked.soil_humidity, 48, axis=0), np.delete(ref.soil_humidity, 48, axis=0))
    assert ref.snow_depth.max() > 2.0 and np.all(np.isfinite(ref.soil_temperature))


def test_solve_air_bit_identical_incl_snow_surface_and_chunking():
    """solve_air! rows (air temperature / wind / humidity profiles at several heights, convective flux, friction velocity)."""
    p, _ = configs.config1_monthly(snow=True)
    p = copy.deepcopy(p)
    p.model.heights = np.array([0.002, 0.01, 0.5, 2.0])          # one height below the roughness length, one interior, the reference height
    f = parity.ALL_FIELDS + parity.SNOW_FIELDS + parity.AIR_FIELDS
    desc, sites, forcing = _buffers(p)
    assert desc.n_heights == 4
    ref, _ = parity.run(OracleEngine, desc, sites, forcing, fields=f)
    out, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(out, ref, f)
    assert np.all(out.wind_speed[:, 0, 0] == 0.0) and np.array_equal(out.air_temperature[:, 3, 0], forcing.reference_temperature[:, 0])
    chunked, _ = parity.run(HostEmuEngine, desc, sites, forcing, fields=f, chunks=[7, 5])
    parity.assert_bit_identical(chunked, ref, parity.AIR_FIELDS)


def test_config4_scan329_bit_identical():
    """BASELINE config 4 (SCAN 329: snow + dynamic moisture + DailyRainfall, consecutive days): the first 40 days (mid-winter, snowpack building) through the
    kernel body; the full year runs on the GPU (tests/test_gpu_parity.py)."""
    p, _ = configs.config4_scan329(40)
    desc, sites, forcing = _buffers(p)
    f = parity.ALL_FIELDS + parity.SNOW_FIELDS
    ref, eo = parity.run(OracleEngine, desc, sites, forcing, fields=f)
    out, ee = parity.run(HostEmuEngine, desc, sites, forcing, fields=f)
    parity.assert_bit_identical(out, ref, f)
    assert ref.snow_depth.max() > 5.0
    co, ce = eo.counters(), ee.counters()
    assert all(co[k] == ce[k] for k in ("n_rhs", "n_accept", "n_reject", "n_newton_iter", "n_infil_steps"))
