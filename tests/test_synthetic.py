"""The synthetic grid generator (BASELINE config 5) is a pure function of (seed, site index, hour): shards and day-chunks
generated independently are slices of the monolithic arrays; values are physical."""
import numpy as np

from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import synthetic as syn


def test_random_access_in_sites_and_time():
    S = 96
    f = syn.make_forcing(0, S, S, 0, 3, threads=1)
    g = syn.make_forcing(16, 48, S, 1, 2, threads=2, block=8)
    for name in A.FORCING_HOURLY:
        a, b = getattr(f, name), getattr(g, name)
        if a is None:
            assert b is None
            continue
        assert np.array_equal(b, a[24:, 16:48]), name
    for name in A.FORCING_DAILY:
        a, b = getattr(f, name), getattr(g, name)
        if a is not None:
            assert np.array_equal(b, a[1:, 16:48]), name
    s_all, s_part = syn.make_sites(0, S, S), syn.make_sites(16, 48, S)
    for name in s_all.NODE_FIELDS + ("initial_soil_moisture", "initial_soil_temperature"):
        assert np.array_equal(getattr(s_part, name), getattr(s_all, name)[:, 16:48]), name
    for name in s_all.SITE_FIELDS:
        assert np.array_equal(getattr(s_part, name), getattr(s_all, name)[16:48]), name


def test_values_are_physical_and_sites_are_latitude_ordered():
    S = 512
    f = syn.make_forcing(0, S, S, 180, 2)
    assert 180.0 < f.reference_temperature.min() and f.reference_temperature.max() < 340.0
    assert f.reference_humidity.min() >= 0.05 and f.reference_humidity.max() <= 1.0
    assert f.reference_wind_speed.min() >= 0.1
    assert f.global_radiation.min() >= 0.0 and f.global_radiation.max() <= 1000.0
    assert f.zenith_angle.min() >= 0.0 and f.zenith_angle.max() <= 90.0
    assert 0.0 <= f.cloud_cover.min() and f.cloud_cover.max() <= 1.0
    assert 0.01 < (f.rainfall_hourly > 0).mean() < 0.08
    assert 60000.0 < f.pressure.min() and f.pressure.max() <= 101325.0
    assert np.allclose(f.rainfall_daily, f.rainfall_hourly.reshape(2, 24, S).sum(axis=1))
    s = syn.make_sites(0, S, S)
    sat = 1.0 - s.bulk_density / s.mineral_density
    assert np.all((s.initial_soil_moisture > 0) & (s.initial_soil_moisture < sat))
    # day length grows with latitude in northern summer (doy ~181): sites are ordered south -> north
    daylight = (f.zenith_angle[:24] < 90.0).sum(axis=0)
    assert daylight[:32].mean() < daylight[-32:].mean()
