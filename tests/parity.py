"""Shared helpers for the parity tests: run the same marshalled buffers through two engines and compare."""
import numpy as np

from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import marshal as M

ALL_FIELDS = A.OUTPUT_NODE_FIELDS + ("surface_water", "sky_temperature")
AIR_FIELDS = A.OUTPUT_PROFILE_FIELDS + A.OUTPUT_PROFILE_SCALARS
SNOW_FIELDS = ("snow_fall", "snow_depth", "snow_density", "snow_temperature")

# north_star tolerances (BASELINE.json): <= 1e-3 K on hourly soil temperature, <= 1e-5 m3/m3 on moisture.
TOL_T_K = 1e-3
TOL_THETA = 1e-5


def run(engine_cls, desc, sites, forcing, fields=ALL_FIELDS, chunks=None, **kw):
    """Integrate `forcing` on `engine_cls`; `chunks` = list of day counts to split the run into."""
    S, nd = sites.S, forcing.ndays
    out = M.OutputBatch(nd, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
    eng = engine_cls(desc, S, **kw)
    eng.set_sites(sites)
    if not chunks:
        eng.step_days(0, forcing, out)
    else:
        assert sum(chunks) == nd
        d0 = 0
        for c in chunks:
            part = M.OutputBatch(c, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
            eng.step_days(d0, forcing.day_slice(d0, d0 + c), part)
            out.assign_rows(24 * d0, part)
            d0 += c
    return out, eng


def max_abs_diff(a: M.OutputBatch, b: M.OutputBatch, fields=ALL_FIELDS):
    res = {}
    for f in fields:
        x, y = getattr(a, f), getattr(b, f)
        if x is None or y is None:
            continue
        res[f] = float(np.max(np.abs(x - y))) if x.size else 0.0
    return res


def assert_within_north_star(a, b, fields=ALL_FIELDS):
    """The tolerance BASELINE.json states, applied to EVERY entry (no percentile)."""
    d = max_abs_diff(a, b, fields)
    assert d.get("soil_temperature", 0.0) <= TOL_T_K, d
    assert d.get("soil_moisture", 0.0) <= TOL_THETA, d
    return d


def assert_bit_identical(a, b, fields=ALL_FIELDS):
    for f in fields:
        x, y = getattr(a, f, None), getattr(b, f, None)
        if x is None or y is None:
            continue
        assert np.array_equal(x, y, equal_nan=True), (f, float(np.nanmax(np.abs(x - y))))
    assert np.array_equal(a.status, b.status)
