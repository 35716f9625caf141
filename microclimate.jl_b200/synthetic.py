"""Deterministic synthetic site grid + hourly forcing for BASELINE.json config 5
("synthetic 1M-site grid x 365 days (varied soil profiles/forcing)"), SURVEY §8d.

Random access in BOTH axes: every value is a pure function of (seed, stream, global site index, time
counter) through a splitmix64 hash, so any site range [s0, s1) and any day range can be generated
independently — shards on different GPUs and successive day-chunks see exactly the arrays a single
monolithic generation would produce.  Sites are ordered by latitude (monotone in the site index) so the
32 sites of a warp share day/night and climate regime.

Model: N=19 DEFAULT_DEPTHS, ConsecutiveDayMode, DynamicSoilMoisture defaults, HourlyRainfall, no snow,
heights (0.01, 2) m — the configuration the headline metric (site-days/s, hourly soil T + moisture) is quoted on.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
from scipy.special import ndtri

from . import marshal as M
from . import types as T

SEED = 20260719
_U64 = np.uint64
_K_SITE = _U64(0x9E3779B97F4A7C15)
_K_TIME = _U64(0xC2B2AE3D27D4EB4F)
_K_STREAM = _U64(0x165667B19E3779F9)

# stream ids
(ST_LAT, ST_TMEAN, ST_ADIU, ST_ELEV, ST_LAI, ST_ALBEDO, ST_TEXTURE, ST_PERT, ST_LAMM, ST_ROOT,
 ST_TNOISE, ST_RH, ST_WIND, ST_CLOUD, ST_RAIN) = range(15)

# Campbell (1985)-style hydraulic classes: (psi_e J/kg, b, K_s kg s/m3, rho_b kg/m3).
# The headline grid uses TEXTURES (loamy sand ... clay).  Pure sand is kept separately (TEXTURES_WITH_SAND): under the
# reference's Campbell retention curve (no residual water) air-dry sand drives the infiltration Newton loop into its
# 500-iteration cap (sites come back flagged MC_STATUS_NEWTON_MAXITER); tests cover that regime on purpose, the
# throughput benchmark does not measure a non-converging loop.
_SAND = (0.7, 1.7, 5.8e-3, 1600.0)
TEXTURES = np.array([
    (0.9, 2.1, 1.7e-3, 1550.0),    # loamy sand
    (1.5, 3.1, 7.2e-4, 1500.0),    # sandy loam
    (1.1, 4.5, 3.7e-4, 1400.0),    # loam
    (2.1, 4.7, 1.9e-4, 1350.0),    # silt loam
    (2.6, 5.2, 6.4e-5, 1300.0),    # clay loam
    (3.7, 7.6, 1.7e-5, 1250.0),    # clay
])
TEXTURES_WITH_SAND = np.vstack([np.array([_SAND]), TEXTURES])
# root density profile of example_campbell_hydraulic_profile (src/soil_hydraulics/campbell.jl:61), m/m3
ROOTS = np.array([0, 0, 8.2, 8.0, 7.8, 7.4, 7.1, 6.4, 5.8, 4.8, 4.0, 1.8, 0.9, 0.6, 0.8, 0.4, 0.4, 0, 0]) * 1e4


def _mix(x):
    x = (x ^ (x >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> _U64(27))) * _U64(0x94D049BB133111EB)
    return x ^ (x >> _U64(31))


def u01(stream: int, site, counter=0):
    """Uniform (0,1) double from the counter-based hash; `site`/`counter` broadcast like numpy arrays."""
    with np.errstate(over="ignore"):
        key = (np.asarray(site, dtype=_U64) * _K_SITE + np.asarray(counter, dtype=_U64) * _K_TIME
               + _U64(stream + 1) * _K_STREAM + _U64(SEED))
        h = _mix(_mix(key))
    return ((h >> _U64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def synthetic_model() -> T.MicroModel:
    cfg = T.MicroConfig(convergence=T.FixedSoilTemperatureIterations(1), rainfall_schedule=T.HourlyRainfall(),
                        soil_moisture_strategy=T.DynamicSoilMoisture())
    hyd = T.CampbellSoilHydraulics(root_resistance=2.5e10, stomatal_closure_potential=-1500.0, leaf_resistance=2.0e6,
                                   stomatal_stability_parameter=10.0, root_radius=0.001)
    return T.MicroModel(depths=T.DEFAULT_DEPTHS, heights=np.array([0.01, 2.0]), soil_hydraulic_model=hyd, config=cfg)


def synthetic_time_mode():
    return T.ConsecutiveDayMode(spinup_first_day=False)


def _site_scalars(s0: int, s1: int, S_total: int):
    s = np.arange(s0, s1, dtype=np.int64)
    # monotone in s within one sweep of S_total sites (warps are latitude-coherent); site ids beyond S_total start another sweep with fresh
    # random streams, so N shards of S_total sites each are statistically identical workloads (bench.py's weak scaling)
    lat = -60.0 + 130.0 * ((s % S_total) + u01(ST_LAT, s)) / S_total
    t_mean = 27.0 - 0.55 * np.abs(lat) + 3.0 * ndtri(u01(ST_TMEAN, s))  # degC
    a_ann = 2.0 + 0.25 * np.abs(lat)
    a_diu = 4.0 + 10.0 * u01(ST_ADIU, s)
    elev = 2500.0 * u01(ST_ELEV, s)
    pressure = 101325.0 * (1.0 - 2.25577e-5 * elev) ** 5.2553
    return s, lat, t_mean, a_ann, a_diu, pressure


def make_sites(s0: int, s1: int, S_total: int, *, with_sand: bool = False) -> M.SitesBatch:
    """Sites [s0, s1) of an S_total-site grid: texture class per site, parameters perturbed +-10 %."""
    n = len(T.DEFAULT_DEPTHS)
    s, lat, t_mean, _, _, _ = _site_scalars(s0, s1, S_total)
    S = s1 - s0
    b = M.SitesBatch(n, S, True)
    b.roughness_height[:] = 0.004
    b.slope[:] = 0.0
    b.sky_view_fraction[:] = 1.0
    b.albedo[:] = 0.1 + 0.2 * u01(ST_ALBEDO, s)
    TEX = TEXTURES_WITH_SAND if with_sand else TEXTURES
    tex = np.minimum((u01(ST_TEXTURE, s) * len(TEX)).astype(np.int64), len(TEX) - 1)
    pert = lambda k: 0.9 + 0.2 * u01(ST_PERT, s, k)
    psi_e = TEX[tex, 0] * pert(0)
    bb = TEX[tex, 1] * pert(1)
    ks = TEX[tex, 2] * pert(2)
    rho_b = TEX[tex, 3] * pert(3)
    b.bulk_density[:] = rho_b[None, :]
    b.mineral_density[:] = 2650.0
    b.mineral_conductivity[:] = (1.2 + 1.7 * u01(ST_LAMM, s))[None, :]
    b.mineral_heat_capacity[:] = 870.0
    b.air_entry_water_potential[:] = psi_e[None, :]
    b.saturated_hydraulic_conductivity[:] = ks[None, :]
    b.campbell_b_parameter[:] = bb[None, :]
    b.root_density[:] = ROOTS[:, None] * u01(ST_ROOT, s)[None, :]
    b.initial_soil_temperature[:] = (t_mean + 273.15)[None, :]
    b.initial_soil_moisture[:] = (0.5 * (1.0 - rho_b / 2650.0))[None, :]
    return b


def _smooth_noise(stream, s, hours, block):
    """Temporally correlated N(0,1)-ish value noise: smoothstep blend of per-(site, block) normals."""
    blk = hours // block
    f = (hours % block) / float(block)
    w = f * f * (3.0 - 2.0 * f)
    ublk = np.unique(np.concatenate([blk, blk + 1]))
    g = ndtri(u01(stream, s[None, :], ublk[:, None]))              # [nblk, S]
    i0 = np.searchsorted(ublk, blk)
    i1 = np.searchsorted(ublk, blk + 1)
    return (1.0 - w)[:, None] * g[i0] + w[:, None] * g[i1]


def _forcing_block(s0, s1, S_total, day0, ndays, out: M.ForcingBatch, o0: int):
    s, lat, t_mean, a_ann, a_diu, pressure = _site_scalars(s0, s1, S_total)
    S = s1 - s0
    H = 24 * ndays
    hours = np.arange(24 * day0, 24 * (day0 + ndays), dtype=np.int64)           # global hour index
    hr = (hours % 24).astype(np.float64)
    doy = ((hours // 24) % 365 + 1).astype(np.float64)
    sl = slice(o0, o0 + S)
    peak = np.where(lat >= 0.0, 200.0, 17.0)
    seasonal = a_ann[None, :] * np.cos(2.0 * np.pi * (doy[:, None] - peak[None, :]) / 365.0)
    t_air = (t_mean[None, :] + seasonal + a_diu[None, :] * np.cos(2.0 * np.pi * (hr[:, None] - 15.0) / 24.0)
             + 1.5 * _smooth_noise(ST_TNOISE, s, hours, 6))
    t_daymin = t_mean[None, :] + seasonal - a_diu[None, :]
    out.reference_temperature[:, sl] = t_air + 273.15
    rh_noise = 0.05 * _smooth_noise(ST_RH, s, hours, 3)
    out.reference_humidity[:, sl] = np.clip(0.85 - 0.03 * (t_air - t_daymin) + rh_noise, 0.05, 1.0)
    blk6 = hours // 6
    ub = np.unique(blk6)
    wind_b = np.maximum(0.1, np.exp(np.log(2.5) + 0.5 * ndtri(u01(ST_WIND, s[None, :], ub[:, None]))))
    out.reference_wind_speed[:, sl] = wind_b[np.searchsorted(ub, blk6)]
    days = np.arange(day0, day0 + ndays, dtype=np.int64)
    uc = u01(ST_CLOUD, s[None, :], days[:, None])
    cloud_d = (1.0 - (1.0 - uc) ** (1.0 / 0.9)) ** (1.0 / 0.6)                   # Kumaraswamy(0.6, 0.9), a closed-form Beta look-alike
    cloud = np.repeat(cloud_d, 24, axis=0)
    out.cloud_cover[:, sl] = cloud
    # solar geometry: declination + hour angle (Cooper 1969)
    decl = np.deg2rad(23.45) * np.sin(2.0 * np.pi * (284.0 + doy) / 365.0)
    ha = np.deg2rad(15.0 * (hr - 12.0))
    latr = np.deg2rad(lat)
    cosz = (np.sin(latr)[None, :] * np.sin(decl)[:, None] + np.cos(latr)[None, :] * np.cos(decl)[:, None] * np.cos(ha)[:, None])
    cosz = np.clip(cosz, 0.0, 1.0)
    zen = np.minimum(90.0, np.rad2deg(np.arccos(cosz)))
    out.zenith_angle[:, sl] = zen
    out.zenith_slope_angle[:, sl] = zen
    solar = np.maximum(0.0, 1000.0 * cosz * (0.36 + 0.64 * (1.0 - cloud)))
    out.global_radiation[:, sl] = solar
    kt = np.where(cosz > 1e-6, solar / np.maximum(1367.0 * cosz, 1e-9), 0.0)    # clearness index -> Erbs et al. (1982)
    fd = np.where(kt <= 0.22, 1.0 - 0.09 * kt,
                  np.where(kt <= 0.80, 0.9511 - 0.1604 * kt + 4.388 * kt ** 2 - 16.638 * kt ** 3 + 12.336 * kt ** 4, 0.165))
    out.diffuse_fraction[:, sl] = np.clip(fd, 0.0, 1.0)
    out.pressure[:, sl] = pressure[None, :]
    ur = u01(ST_RAIN, s[None, :], hours[:, None])
    out.rainfall_hourly[:, sl] = np.where(ur < 0.04, -1.5 * np.log(np.maximum(ur, 1e-300) / 0.04), 0.0)
    out.leaf_area_index[:, sl] = (0.1 + 2.9 * u01(ST_LAI, s))[None, :]
    out.shade[:, sl] = 0.0
    out.surface_emissivity[:, sl] = 0.96
    out.cloud_emissivity[:, sl] = 0.96
    out.soil_wetness[:, sl] = 0.0
    out.deep_soil_temperature[:, sl] = (t_mean + 273.15)[None, :]
    out.rainfall_daily[:, sl] = out.rainfall_hourly[:, sl].reshape(ndays, 24, S).sum(axis=1)


def make_forcing(s0: int, s1: int, S_total: int, day0: int, ndays: int, *, threads: int = 0, block: int = 16384,
                 out: M.ForcingBatch | None = None) -> M.ForcingBatch:
    """Hourly + daily forcing of sites [s0, s1) for days [day0, day0+ndays). `out` may be a pre-allocated
    (e.g. pinned) ForcingBatch."""
    S = s1 - s0
    if out is None:
        out = M.ForcingBatch(ndays, S, hourly_rain=True, day_of_year=[(d % 365) + 1 for d in range(day0, day0 + ndays)])
    else:
        out.day_of_year[:] = [(d % 365) + 1 for d in range(day0, day0 + ndays)]
    jobs = [(a, min(s1, a + block)) for a in range(s0, s1, block)]
    threads = threads or min(len(jobs), os.cpu_count() or 1)
    if threads <= 1:
        for a, b in jobs:
            _forcing_block(a, b, S_total, day0, ndays, out, a - s0)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda ab: _forcing_block(ab[0], ab[1], S_total, day0, ndays, out, ab[0] - s0), jobs))
    return out
