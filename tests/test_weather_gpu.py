"""GPU tests of the widening rows f2 / f3 (run on the B200 with `-m gpu`): weather_kernel through the C ABI against the oracle, bit for bit, on the cases
of tests/test_weather.py; the golden ZEN vectors through CUDA; and the fused min/max step (expand on the device, then solve_soil!) against
populate-then-solve with the CPU oracle on both halves."""
import copy
import os

import numpy as np
import pytest

import parity
from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import marshal as M
from microclimate_jl_b200 import synthetic as syn
from microclimate_jl_b200 import weather as W
from oracle import oracle as O
from oracle.oracle import OracleEngine
from test_weather import ALL_OUT, CUSTOM_CURVES, GOLD, _derived_case, _random_problem, _soil_case, check_example_problem, check_minmax_snow_and_air_outputs, check_solve_minmax_batched

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def SoilEngine():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from microclimate_jl_b200.simulation import SoilEngine as E
    return E


def _engine(SoilEngine, S):
    return SoilEngine(M.model_desc(syn.synthetic_model(), syn.synthetic_time_mode()), S)


@pytest.mark.parametrize("case", ["standard", "custom", "given_solar"])
def test_weather_kernel_matches_oracle(SoilEngine, case):
    """4 099 sites (a ragged last block) x 6 days incl. polar, sloped and clamped cases: every hourly and daily output of weather_kernel == oracle, bit for bit."""
    w, win, sites, _ = _random_problem({"standard": 31, "custom": 32, "given_solar": 33}[case], S=4099, D=6, curves=CUSTOM_CURVES if case == "custom" else None,
                                       with_given_solar=case == "given_solar")
    ref, got = W.WeatherOutBatch(win.ndays, win.S), W.WeatherOutBatch(win.ndays, win.S)
    O.populate_weather(w, win, ref, sites)
    with _engine(SoilEngine, win.S) as eng:
        eng.populate_weather(w, win, got, sites)
        for k in ALL_OUT:
            a, b = getattr(ref, k), getattr(got, k)
            assert np.array_equal(a, b, equal_nan=True), (k, float(np.nanmax(np.abs(a - b))))
        part = W.WeatherOutBatch(win.ndays, win.S, ("diffuse_fraction", "zenith_slope_angle", "hour_angle_sunrise"))
        eng.populate_weather(w, win, part, sites)
        for k in ("diffuse_fraction", "zenith_slope_angle", "hour_angle_sunrise"):
            assert np.array_equal(getattr(part, k), getattr(ref, k), equal_nan=True), k
        with pytest.raises(RuntimeError, match="clear-sky"):
            eng.populate_weather(w, W.WeatherInBatch(win.days, win.S, win.forcings, pressure_site=win.pressure_site), W.WeatherOutBatch(win.ndays, win.S), sites)


@pytest.mark.parametrize("name", ["monthly", "daily", "scan329"])
def test_golden_zenith_through_cuda(SoilEngine, name):
    """The reference's golden ZEN columns (9 768 hourly values, three sites) reproduced by the device geometry to <= 1e-5 degrees."""
    doy, zen, lat = GOLD[f"zen_{name}_doy"], GOLD[f"zen_{name}_zen"], float(GOLD[f"zen_{name}_lat"])
    days = doy[::24]
    out = W.WeatherOutBatch(len(days), 1, ("zenith_angle", "zenith_slope_angle"))
    with _engine(SoilEngine, 1) as eng:
        eng.populate_weather(W.weather_desc(), W.WeatherInBatch(days, 1, {}), out, W.SolarSitesBatch([lat]))
    assert np.abs(out.zenith_angle[:, 0] - zen).max() <= 1e-5
    assert np.array_equal(out.zenith_slope_angle, out.zenith_angle)


def test_step_days_minmax_matches_populate_then_solve(SoilEngine):
    """mc_step_days_minmax on the B200 (weather expanded on the device into the soil kernel's forcing, then the soil kernel) == oracle weather followed by
    the oracle soil integration: every soil output and the expanded weather bit for bit; then the same run day by day."""
    desc, sites, rest, win, solar = _soil_case(S=96, D=3, seed=41)
    w = W.weather_desc()
    S, D = win.S, win.ndays
    wx = W.WeatherOutBatch(D, S)
    O.populate_weather(w, win, wx, solar)
    hourly = copy.deepcopy(rest)
    for k in A.WEATHER_OUT_HOURLY:
        getattr(hourly, k)[...] = getattr(wx, k)
    fields = parity.ALL_FIELDS
    ref, _ = parity.run(OracleEngine, desc, sites, hourly, fields=fields)
    with SoilEngine(desc, S) as eng:
        eng.set_sites(sites)
        out = M.OutputBatch(D, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
        got_wx = W.WeatherOutBatch(D, S)
        eng.step_days_minmax(0, w, win, rest, out, solar, got_wx)
        parity.assert_within_north_star(out, ref)
        parity.assert_bit_identical(out, ref, fields)
        for k in ALL_OUT:
            assert np.array_equal(getattr(got_wx, k), getattr(wx, k)), k
    with SoilEngine(desc, S) as eng:
        eng.set_sites(sites)
        out = M.OutputBatch(D, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
        for d in range(D):
            part = M.OutputBatch(1, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
            wd = W.WeatherInBatch(win.days[d:d + 1], S, {q: tuple(v[d:d + 1] for v in win.forcings[q]) for q in A.WEATHER_QUANTITIES},
                                  direct_horizontal=win.direct_horizontal[24 * d:24 * d + 24], diffuse_horizontal=win.diffuse_horizontal[24 * d:24 * d + 24],
                                  pressure_site=win.pressure_site)
            eng.step_days_minmax(d, w, wd, hourly.day_slice(d, d + 1), part, solar, None, use_rest_zenith=(d == 1))
            out.assign_rows(24 * d, part)
        # soil_humidity: see tests/test_weather.py (row 1 quirk of initialise_soil_humidity! under chunking)
        parity.assert_bit_identical(out, ref, tuple(k for k in fields if k != "soil_humidity"))


def test_solve_minmax_batched_on_the_b200(SoilEngine):
    """solve_minmax_batched (config 1 from its min/max inputs, three sites incl. sloped ones, whole and in day chunks) through the CUDA library == oracle weather + oracle soil."""
    check_solve_minmax_batched(None)


def test_derived_humidity_matches_oracle(SoilEngine):
    """Derived(RelativeHumidityFromVapourPressureAndTemperature) on the device: humidity = clamp(e / e_sat(T), 0, 1) from the vapour-pressure curve and the hour's temperature."""
    w, win, sites, _, _ = _derived_case(S=3001, D=3, seed=51)
    fields = ("reference_temperature", "reference_humidity", "cloud_cover")
    ref = W.WeatherOutBatch(win.ndays, win.S, fields)
    O.populate_weather(w, win, ref, sites)
    with _engine(SoilEngine, win.S) as eng:
        for flds in (fields, ("reference_humidity",)):
            got = W.WeatherOutBatch(win.ndays, win.S, flds)
            eng.populate_weather(w, win, got, sites)
            for k in flds:
                assert np.array_equal(getattr(got, k), getattr(ref, k)), k


def test_multi_device_minmax_step_on_the_b200(SoilEngine):
    """mc_multi_step_days_minmax: one host process, the sites of ONE set of host arrays sharded over three pipelines (the same GPU listed three times on a 1-GPU box, as
    many distinct GPUs as there are otherwise) — soil outputs, weather read-back and status equal to the single-handle call."""
    import torch
    from microclimate_jl_b200.simulation import MultiSoilEngine
    desc, sites, rest, win, solar = _soil_case(S=70, D=2, seed=43)
    w = W.weather_desc()
    S, D = win.S, win.ndays
    fields = ("soil_temperature", "soil_moisture", "surface_water", "sky_temperature")
    with SoilEngine(desc, S) as one:
        one.set_sites(sites)
        ref, ref_wx = M.OutputBatch(D, desc.n_soil, S, fields), W.WeatherOutBatch(D, S)
        one.step_days_minmax(0, w, win, rest, ref, solar, ref_wx)
    n = torch.cuda.device_count()
    devices = [g % n for g in range(3)]
    with MultiSoilEngine(desc, S, devices) as multi:
        multi.set_sites(sites)
        out, wx = M.OutputBatch(D, desc.n_soil, S, fields), W.WeatherOutBatch(D, S)
        multi.step_days_minmax(0, w, win, rest, out, solar, wx)
        parity.assert_bit_identical(out, ref, fields)
        assert np.array_equal(out.status, ref.status)
        for k in ALL_OUT:
            assert np.array_equal(getattr(wx, k), getattr(ref_wx, k)), k


def test_example_microclimate_problem_solves_on_the_b200(SoilEngine):
    """The reference's own example problem (test/example_constructors.jl: min/max-driven Madison year, N = 19) end to end through the CUDA library."""
    check_example_problem(None)


def test_minmax_step_with_snow_and_air_outputs_on_the_b200(SoilEngine):
    """Every output family through mc_step_days_minmax on the B200: soil fields, SnowModel scalars + 8-node profile, solve_air! profiles (config 3 from its min/max inputs)."""
    check_minmax_snow_and_air_outputs(None)
