"""Host-side mirror of the reference's min/max weather interface (the widening rows f2 / f3 of SURVEY 8) over the C ABI.

Same names and argument meaning as the reference:
  TimeOfDay anchors `Sunrise, Sunset, Midday, Midnight, ClockTime`, shapes `Sine, Decay, Linear`, `DielCurve`, `ForcingSpec`,
  `minmax_forcing_model`, `minmax_forcings`, `MonthlyMinMaxEnvironment`, `DailyMinMaxEnvironment`, `example_monthly_weather`
      (src/inputs/diel.jl, src/inputs/monthly_minmax.jl, src/inputs/daily_minmax.jl),
  `Angstrom`, `ErbsDiffuseFraction`, `AngstromMaxwellShortwave`
      (src/radiation/sunshine_fraction/angstrom.jl, src/radiation/diffuse/erbs.jl, src/radiation/shortwave/angstrom_maxwell.jl),
and `populate_weather` / `SoilEngine.step_days_minmax` in place of `populate_weather!` + `solve_soil!` (src/simulation.jl:20-36,246-260).
Unit-stripped SI: temperatures in K, fractions 0..1, angles in degrees, hours as decimal hours.

Only marshalling lives here; the arithmetic runs in the CUDA library (csrc/mc_weather.cuh).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _abi as A


# ---- TimeOfDay (diel.jl:19-85) ----
@dataclass(frozen=True)
class Sunrise:
    offset: float = 0.0
    KIND = A.MC_TOD_SUNRISE
    NOMINAL = 6.0


@dataclass(frozen=True)
class Sunset:
    offset: float = 0.0
    KIND = A.MC_TOD_SUNSET
    NOMINAL = 18.0


@dataclass(frozen=True)
class Midday:
    offset: float = 0.0
    KIND = A.MC_TOD_MIDDAY
    NOMINAL = 12.0


@dataclass(frozen=True)
class Midnight:
    offset: float = 0.0
    KIND = A.MC_TOD_MIDNIGHT
    NOMINAL = 0.0


@dataclass(frozen=True)
class ClockTime:
    hour: float
    KIND = A.MC_TOD_CLOCK

    @property
    def offset(self):
        return self.hour
    NOMINAL = 0.0


def nominal_hour(t) -> float:                               # diel.jl:87-91
    return (t.NOMINAL + t.offset) % 24.0


def _tod(t) -> A.TimeOfDay:
    return A.TimeOfDay(t.KIND, 0, float(t.offset))


# ---- Shape (diel.jl:98-150) ----
@dataclass(frozen=True)
class Sine:
    trough: object
    peak: object
    KIND = A.MC_SHAPE_SINE

    def endpoints(self):
        return (self.trough, self.peak)


@dataclass(frozen=True)
class Decay:
    from_: object
    to: object
    KIND = A.MC_SHAPE_DECAY

    def endpoints(self):
        return (self.from_, self.to)


@dataclass(frozen=True)
class Linear:
    from_: object
    to: object
    KIND = A.MC_SHAPE_LINEAR

    def endpoints(self):
        return (self.from_, self.to)


def is_space_filling(s) -> bool:
    return isinstance(s, Sine)


class DielCurve:
    """DielCurve(shapes; inputs) with the reference constructor's checks (diel.jl:172-222)."""

    def __init__(self, shapes: Sequence, inputs: Sequence):
        self.shapes, self.inputs = tuple(shapes), tuple(inputs)
        if not 1 <= len(self.shapes) <= A.MC_DIEL_MAX_SHAPES or len(self.inputs) > A.MC_DIEL_MAX_INPUTS:
            raise ValueError(f"DielCurve: at most {A.MC_DIEL_MAX_SHAPES} shapes and {A.MC_DIEL_MAX_INPUTS} inputs fit the C ABI")
        self._validate_coverage()
        self._validate_inputs()

    def _validate_coverage(self):
        nfill = sum(is_space_filling(s) for s in self.shapes)
        if nfill > 1:
            raise ValueError(f"DielCurve: {nfill} space-filling shapes; coverage is ambiguous (allow at most one)")
        arcs = []
        for s in self.shapes:
            if is_space_filling(s):
                continue
            f, t = s.endpoints()
            a, b = nominal_hour(f), nominal_hour(t)
            ln = (b - a) % 24.0
            if ln == 0.0:
                raise ValueError(f"DielCurve: shape {s} spans zero length")
            arcs.append((a, ln))
        if not arcs:
            return
        arcs.sort()
        for i, (a, ln) in enumerate(arcs):
            cover_end = a + ln
            next_start = arcs[i + 1][0] if i + 1 < len(arcs) else arcs[0][0] + 24.0
            if cover_end > next_start + 1e-9:
                raise ValueError(f"DielCurve: shapes overlap near hour {next_start % 24.0}")
            if nfill == 0 and cover_end < next_start - 1e-9:
                raise ValueError(f"DielCurve: day not covered - gap from hour {cover_end % 24.0} to {next_start % 24.0}; add a shape or a space-filling one")

    def _validate_inputs(self):
        eps = [e for s in self.shapes for e in s.endpoints()]
        for inp in self.inputs:
            if inp not in eps:
                raise ValueError(f"DielCurve: input {inp} is not an endpoint of any shape")

    def seams(self):
        return tuple(e for s in self.shapes for e in s.endpoints() if e not in self.inputs)

    def c_struct(self) -> A.DielCurve:
        c = A.DielCurve()
        c.n_shapes, c.n_inputs = len(self.shapes), len(self.inputs)
        for i, s in enumerate(self.shapes):
            a, b = s.endpoints()
            c.shapes[i] = A.DielShape(s.KIND, 0, _tod(a), _tod(b))
        for i, t in enumerate(self.inputs):
            c.inputs[i] = _tod(t)
        return c


@dataclass(frozen=True)
class ForcingSpec:                                          # diel.jl:244-262
    curve: DielCurve
    inputs: Tuple[str, ...]


def minmax_forcing_model(temperature_sunrise_offset=0, temperature_midday_offset=1, wind_sunrise_offset=0, wind_midday_offset=1,
                         humidity_sunrise_offset=1, humidity_midday_offset=0, cloud_sunrise_offset=1, cloud_midday_offset=0) -> Dict[str, ForcingSpec]:
    """The classic daily min/max within-day structure (diel.jl:456-495)."""
    def three(so, mo, names):
        return ForcingSpec(DielCurve((Linear(Midnight(), Sunrise(so)), Linear(Sunrise(so), Midday(mo)), Linear(Midday(mo), Midnight())),
                                     inputs=(Midnight(), Sunrise(so), Midday(mo))), names)
    ts, tm = float(temperature_sunrise_offset), float(temperature_midday_offset)
    return {
        "reference_temperature": ForcingSpec(DielCurve((Sine(Sunrise(ts), Midday(tm)), Decay(Sunset(), Sunrise(ts))), inputs=(Sunrise(ts), Midday(tm))),
                                             ("reference_temperature_min", "reference_temperature_max")),
        "reference_wind_speed": three(float(wind_sunrise_offset), float(wind_midday_offset),
                                      ("reference_wind_speed_mean", "reference_wind_speed_min", "reference_wind_speed_max")),
        "reference_humidity": three(float(humidity_sunrise_offset), float(humidity_midday_offset),
                                    ("reference_humidity_mean", "reference_humidity_max", "reference_humidity_min")),
        "cloud_cover": three(float(cloud_sunrise_offset), float(cloud_midday_offset), ("cloud_cover_mean", "cloud_cover_max", "cloud_cover_min")),
    }


MINMAX_FORCING_MODEL = minmax_forcing_model()


@dataclass
class DielForcing:                                          # diel.jl:224-236: a curve plus its per-day value series
    curve: DielCurve
    values: Tuple[np.ndarray, ...]


def bind_forcings(model: Dict[str, object], valuesource) -> Dict[str, object]:
    """bind_forcings(model, valuesource) (diel.jl:414-431): ForcingSpecs are bound to the value series `valuesource(name)` of their inputs, Derived nodes pass through."""
    return {k: (DielForcing(n.curve, tuple(np.asarray(valuesource(i), dtype=np.float64) for i in n.inputs)) if isinstance(n, ForcingSpec) else n) for k, n in model.items()}


def minmax_forcings(*, reference_temperature_min, reference_temperature_max, reference_wind_speed_min, reference_wind_speed_max,
                    reference_humidity_min, reference_humidity_max, cloud_cover_min, cloud_cover_max, **offsets) -> Dict[str, DielForcing]:
    """Bind the standard min/max model to per-day series (diel.jl:516-542); the means are (min + max) / 2 as in the reference."""
    g = {k: np.asarray(v, dtype=np.float64) for k, v in dict(
        reference_temperature_min=reference_temperature_min, reference_temperature_max=reference_temperature_max,
        reference_wind_speed_min=reference_wind_speed_min, reference_wind_speed_max=reference_wind_speed_max,
        reference_humidity_min=reference_humidity_min, reference_humidity_max=reference_humidity_max,
        cloud_cover_min=cloud_cover_min, cloud_cover_max=cloud_cover_max).items()}
    for q in ("reference_wind_speed", "reference_humidity", "cloud_cover"):
        g[q + "_mean"] = (g[q + "_min"] + g[q + "_max"]) / 2
    model = minmax_forcing_model(**offsets) if offsets else MINMAX_FORCING_MODEL
    return bind_forcings(model, lambda name: g[name])


# ---- Derived quantities (diel.jl:264-312): computed hour by hour from other quantities instead of interpolated from a curve ----
@dataclass(frozen=True)
class GoffGratch:                                           # FluidProperties.jl's default vapour-pressure equation (the only one the path uses)
    pass


@dataclass(frozen=True)
class RelativeHumidityFromVapourPressureAndTemperature:     # diel.jl:280-287: (e, T) -> clamp(e / e_sat(T), 0, 1)
    saturation_method: object = field(default_factory=GoffGratch)


@dataclass(frozen=True)
class VapourPressureFromRelativeHumidityAndTemperature:     # diel.jl:289-301: (rh, T) -> rh * e_sat(T); an intermediate, never an output of populate_weather!
    saturation_method: object = field(default_factory=GoffGratch)


@dataclass(frozen=True)
class Derived:
    functor: object
    inputs: Tuple[str, ...]


def humidity_forcing(forcings: Dict[str, object]):
    """The DielForcing that fills the ABI's reference_humidity slot, and mc_weather_desc.humidity_source: the humidity's own curve, or — for
    `reference_humidity = Derived(RelativeHumidityFromVapourPressureAndTemperature(GoffGratch()), ("actual_vapour_pressure", "reference_temperature"))` — the
    curve and series of the vapour pressure it is derived from (Pa)."""
    f = forcings["reference_humidity"]
    if isinstance(f, DielForcing):
        return f, A.MC_HUMIDITY_CURVE
    if (isinstance(f, Derived) and isinstance(f.functor, RelativeHumidityFromVapourPressureAndTemperature) and isinstance(f.functor.saturation_method, GoffGratch)
            and len(f.inputs) == 2 and f.inputs[1] == "reference_temperature" and isinstance(forcings.get(f.inputs[0]), DielForcing)):
        return forcings[f.inputs[0]], A.MC_HUMIDITY_FROM_VAPOUR_PRESSURE
    raise ValueError("reference_humidity must be a DielForcing or Derived(RelativeHumidityFromVapourPressureAndTemperature(GoffGratch()), "
                     "(<vapour pressure forcing>, 'reference_temperature')); other Derived quantities are not on the device")


def forcing_of(forcings: Dict[str, object], q: str) -> DielForcing:
    """The DielForcing behind the ABI slot of output quantity q."""
    if q == "reference_humidity":
        return humidity_forcing(forcings)[0]
    f = forcings[q]
    if not isinstance(f, DielForcing):
        raise ValueError(f"{q} must be a DielForcing (Derived is supported for reference_humidity only)")
    return f


@dataclass
class MonthlyMinMaxEnvironment:                             # src/inputs/monthly_minmax.jl:13-16
    forcings: Dict[str, DielForcing]


@dataclass
class DailyMinMaxEnvironment:                               # src/inputs/daily_minmax.jl:13-16
    forcings: Dict[str, DielForcing]


def example_monthly_weather() -> MonthlyMinMaxEnvironment:   # src/inputs/monthly_minmax.jl:24-39 (Madison, CRU CL 2.0), in K and fractions
    tmin = np.array([-14.3, -12.1, -5.1, 1.2, 6.9, 12.3, 15.2, 13.6, 8.9, 3, -3.2, -10.6]) + 273.15
    tmax = np.array([-3.2, 0.1, 6.8, 14.6, 21.3, 26.4, 29, 27.7, 23.3, 16.6, 7.8, -0.4]) + 273.15
    wind = np.array([4.9, 4.8, 5.2, 5.3, 4.6, 4.3, 3.8, 3.7, 4, 4.6, 4.9, 4.8])
    rhmin = np.array([50.2, 48.4, 48.7, 40.8, 40, 42.1, 45.5, 47.3, 47.6, 45, 51.3, 52.8]) / 100.0
    cloud = np.array([50.3, 47, 48.2, 47.5, 40.9, 35.7, 34.1, 36.6, 42.6, 48.4, 61.1, 60.1]) / 100.0
    return MonthlyMinMaxEnvironment(minmax_forcings(
        reference_temperature_min=tmin, reference_temperature_max=tmax, reference_wind_speed_min=wind * 0.1, reference_wind_speed_max=wind,
        reference_humidity_min=rhmin, reference_humidity_max=np.full(12, 1.0), cloud_cover_min=cloud, cloud_cover_max=cloud))


# ---- shortwave ----
@dataclass(frozen=True)
class Angstrom:                                             # sunshine_fraction/angstrom.jl:8-12
    a: float = 0.36
    b: float = 0.64
    gamma: float = 1.0


@dataclass(frozen=True)
class ErbsDiffuseFraction:                                  # diffuse/erbs.jl:12
    pass


@dataclass(frozen=True)
class AngstromMaxwellShortwave:                             # shortwave/angstrom_maxwell.jl:16-19
    diffuse_fraction_model: object = field(default_factory=ErbsDiffuseFraction)
    sunshine_fraction_model: Angstrom = field(default_factory=Angstrom)


@dataclass(frozen=True)
class SolarGeometry:
    """Constants of the built-in solar geometry (McCullough & Porter 1971 as in NicheMapR's SOLRAD; include/microclimate_b200.h)."""
    orbital_eccentricity: float = 0.0167238
    sin_obliquity: float = 0.39784993
    days_per_year: float = 365.0
    equinox_day: float = 80.0
    refraction: bool = True


def weather_desc(curves: Optional[Dict[str, DielCurve]] = None, shortwave: AngstromMaxwellShortwave = AngstromMaxwellShortwave(),
                 geometry: SolarGeometry = SolarGeometry(), humidity_source: int = A.MC_HUMIDITY_CURVE, forcings: Optional[Dict[str, object]] = None) -> A.WeatherDesc:
    """Flatten forcing curves + shortwave model + geometry constants into mc_weather_desc.  Pass either `curves` (one DielCurve per output quantity) or an
    environment's `forcings` (DielForcing / Derived nodes, as bind_forcings leaves them): the latter also resolves a derived humidity."""
    if forcings is not None:
        curves = {q: forcing_of(forcings, q).curve for q in A.WEATHER_QUANTITIES}
        humidity_source = humidity_forcing(forcings)[1]
    if not isinstance(shortwave.diffuse_fraction_model, ErbsDiffuseFraction):
        raise ValueError("only ErbsDiffuseFraction exists (src/radiation/diffuse)")
    curves = curves or {k: s.curve for k, s in MINMAX_FORCING_MODEL.items()}
    w = A.WeatherDesc()
    for q in A.WEATHER_QUANTITIES:
        setattr(w, q, curves[q].c_struct())
    sf = shortwave.sunshine_fraction_model
    w.angstrom_a, w.angstrom_b, w.angstrom_gamma = sf.a, sf.b, sf.gamma
    w.solar_constant = 1367.0
    w.diffuse_model = A.MC_DIFFUSE_ERBS
    w.refraction = 1 if geometry.refraction else 0
    w.humidity_source = humidity_source
    w.orbital_eccentricity, w.sin_obliquity = geometry.orbital_eccentricity, geometry.sin_obliquity
    w.days_per_year, w.equinox_day = geometry.days_per_year, geometry.equinox_day
    return w


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a, shape, name):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != shape:
        raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
    return a


class SolarSitesBatch:
    """mc_solar_sites_soa: what SolarTerrain(site) takes from Site (src/site/site.jl:48-60) for S sites."""

    def __init__(self, latitude, slope=None, aspect=None, solar_noon_offset_h=None):
        self.latitude = np.ascontiguousarray(latitude, dtype=np.float64)
        S = self.S = self.latitude.shape[0]
        self.slope, self.aspect = _arr(slope, (S,), "slope"), _arr(aspect, (S,), "aspect")
        self.solar_noon_offset_h = _arr(solar_noon_offset_h, (S,), "solar_noon_offset_h")

    def slice(self, lo: int, hi: int) -> "SolarSitesBatch":
        """Contiguous site range [lo, hi) — the unit of multi-GPU sharding."""
        cut = lambda a: None if a is None else np.ascontiguousarray(a[lo:hi])      # noqa: E731
        return SolarSitesBatch(cut(self.latitude), cut(self.slope), cut(self.aspect), cut(self.solar_noon_offset_h))

    def c_struct(self) -> A.SolarSitesSoA:
        c = A.SolarSitesSoA()
        for f in A.SOLAR_SITE_FIELDS:
            setattr(c, f, _ptr(getattr(self, f)))
        return c


class WeatherInBatch:
    """mc_weather_in_soa for `ndays` days of S sites.  `forcings[q]` = tuple of [ndays][S] arrays in curve-input order."""

    def __init__(self, days: Sequence[int], n_sites: int, forcings: Dict[str, Sequence[np.ndarray]], *, direct_horizontal=None, diffuse_horizontal=None,
                 hour_solar_noon=None, hour_angle_sunrise=None, global_radiation_hourly=None, pressure_hourly=None, pressure_site=None):
        self.days = np.ascontiguousarray(days, dtype=np.int32)
        D, S = self.ndays, self.S = len(self.days), n_sites
        self.forcings = {q: tuple(_arr(v, (D, S), f"{q} values") for v in forcings.get(q, ())) for q in A.WEATHER_QUANTITIES}
        self.hour_solar_noon, self.hour_angle_sunrise = _arr(hour_solar_noon, (D, S), "hour_solar_noon"), _arr(hour_angle_sunrise, (D, S), "hour_angle_sunrise")
        H = 24 * D
        self.direct_horizontal, self.diffuse_horizontal = _arr(direct_horizontal, (H, S), "direct_horizontal"), _arr(diffuse_horizontal, (H, S), "diffuse_horizontal")
        self.global_radiation_hourly, self.pressure_hourly = _arr(global_radiation_hourly, (H, S), "global_radiation_hourly"), _arr(pressure_hourly, (H, S), "pressure_hourly")
        self.pressure_site = _arr(pressure_site, (S,), "pressure_site")

    @classmethod
    def from_environments(cls, days, envs: Sequence, **kw):
        """One MonthlyMinMaxEnvironment / DailyMinMaxEnvironment per site (the reference's per-site inputs), transposed to site-fastest arrays."""
        S = len(envs)
        f = {q: tuple(np.stack([np.asarray(forcing_of(e.forcings, q).values[i], dtype=np.float64) for e in envs], axis=1)
                      for i in range(len(forcing_of(envs[0].forcings, q).values))) for q in A.WEATHER_QUANTITIES}
        return cls(days, S, f, **kw)

    def site_slice(self, lo: int, hi: int) -> "WeatherInBatch":
        """Sites [lo, hi) of this batch (multi-GPU shards)."""
        cut = lambda a: None if a is None else np.ascontiguousarray(a[..., lo:hi])      # noqa: E731
        return WeatherInBatch(self.days, hi - lo, {q: tuple(cut(v) for v in self.forcings[q]) for q in A.WEATHER_QUANTITIES},
                              direct_horizontal=cut(self.direct_horizontal), diffuse_horizontal=cut(self.diffuse_horizontal),
                              hour_solar_noon=cut(self.hour_solar_noon), hour_angle_sunrise=cut(self.hour_angle_sunrise),
                              global_radiation_hourly=cut(self.global_radiation_hourly), pressure_hourly=cut(self.pressure_hourly),
                              pressure_site=cut(self.pressure_site))

    def day_slice(self, d0: int, d1: int) -> "WeatherInBatch":
        """Days [d0, d1) of this batch (time-chunked runs)."""
        cut = lambda a, k: None if a is None else np.ascontiguousarray(a[k * d0:k * d1])      # noqa: E731
        return WeatherInBatch(self.days[d0:d1], self.S, {q: tuple(cut(v, 1) for v in self.forcings[q]) for q in A.WEATHER_QUANTITIES},
                              direct_horizontal=cut(self.direct_horizontal, 24), diffuse_horizontal=cut(self.diffuse_horizontal, 24),
                              hour_solar_noon=cut(self.hour_solar_noon, 1), hour_angle_sunrise=cut(self.hour_angle_sunrise, 1),
                              global_radiation_hourly=cut(self.global_radiation_hourly, 24), pressure_hourly=cut(self.pressure_hourly, 24),
                              pressure_site=self.pressure_site)

    def c_struct(self) -> A.WeatherInSoA:
        c = A.WeatherInSoA()
        for q in A.WEATHER_QUANTITIES:
            arr = getattr(c, q + "_values")
            for i, v in enumerate(self.forcings[q]):
                arr[i] = _ptr(v)
        for f in A.WEATHER_IN_DAILY + A.WEATHER_IN_HOURLY + ("pressure_site",):
            setattr(c, f, _ptr(getattr(self, f)))
        c.day_of_year = self.days.ctypes.data_as(C.POINTER(C.c_int32))
        return c


class WeatherOutBatch:
    """mc_weather_out_soa host buffers: hourly [24*ndays][S], daily [ndays][S]; unselected = None."""

    def __init__(self, ndays: int, n_sites: int, fields: Sequence[str] = A.WEATHER_OUT_HOURLY + A.WEATHER_OUT_DAILY):
        self.ndays, self.S = ndays, n_sites
        for f in A.WEATHER_OUT_HOURLY:
            setattr(self, f, np.full((24 * ndays, n_sites), np.nan) if f in fields else None)
        for f in A.WEATHER_OUT_DAILY:
            setattr(self, f, np.full((ndays, n_sites), np.nan) if f in fields else None)

    def c_struct(self) -> A.WeatherOutSoA:
        c = A.WeatherOutSoA()
        for f in A.WEATHER_OUT_HOURLY + A.WEATHER_OUT_DAILY:
            setattr(c, f, _ptr(getattr(self, f)))
        return c


# ---- (de)serialisation of the forcing model for the on-disk containers (storage.py): structure only, no data ----
_TOD = {"Sunrise": Sunrise, "Sunset": Sunset, "Midday": Midday, "Midnight": Midnight, "ClockTime": ClockTime}
_SHAPES = {"Sine": Sine, "Decay": Decay, "Linear": Linear}


def curve_to_json(c: DielCurve) -> dict:
    tod = lambda t: [type(t).__name__, float(t.offset)]          # noqa: E731
    return {"shapes": [[type(sh).__name__, tod(sh.endpoints()[0]), tod(sh.endpoints()[1])] for sh in c.shapes], "inputs": [tod(t) for t in c.inputs]}


def curve_from_json(d: dict) -> DielCurve:
    tod = lambda x: _TOD[x[0]](float(x[1]))                      # noqa: E731
    return DielCurve(tuple(_SHAPES[k](tod(a), tod(b)) for k, a, b in d["shapes"]), tuple(tod(t) for t in d["inputs"]))


def validate_weather(w: A.WeatherDesc) -> None:
    """mc_validate_weather: the DielCurve constructor checks as the library applies them to a flattened descriptor."""
    from . import _lib
    buf = C.create_string_buffer(512)
    if _lib.lib().mc_validate_weather(C.byref(w), buf, 512):
        raise ValueError(buf.value.decode())
