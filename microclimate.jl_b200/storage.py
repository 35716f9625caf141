"""On-disk containers for the batched-sites engine: site-fastest columnar files for gridded forcing,
chunked outputs, site tables and state checkpoints.

One file = one batch object (`SitesBatch`, `ForcingBatch`, `OutputBatch`, `StateBatch`) stored exactly as the C ABI
wants it in memory — little-endian float64 (int32 for status / day-of-year, int64 for days_done), C order, **site index
fastest** — so reading is `np.memmap` and a day chunk of a forcing file, or a field of an output file, goes to
`mc_step_days` / comes from it without a transpose or a copy.  Layout (documented in INTEGRATION.md §6 so that the Julia
side can `Mmap.mmap` the same files):

    bytes 0..7     magic  b"MCB200\\x00\\x01"
    bytes 8..15    uint64 little-endian: length L of the JSON header
    bytes 16..16+L JSON (utf-8): {"kind", "S", "n_soil", "ndays", "day0", "meta", "arrays": [{"name", "dtype", "shape", "offset"}]}
    arrays         each at its absolute byte `offset`, 4096-byte aligned, raw C-order data

Multi-GPU / multi-process use: every rank memory-maps the same file and slices its site range `[lo, hi)` (a strided
view; `ascontiguous=True` makes the pinned-copy source contiguous); time chunking slices days, which are contiguous.
"""
import json
import os
import struct
from typing import Dict, Optional

import numpy as np

from . import _abi as A
from . import marshal as M

MAGIC = b"MCB200\x00\x01"
ALIGN = 4096


def _align(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def write_container(path: str, kind: str, arrays: Dict[str, Optional[np.ndarray]], **header) -> None:
    """Write named arrays (None entries are skipped) into one container file, atomically (temp file + rename)."""
    path = os.fspath(path)
    items = [(k, np.ascontiguousarray(v)) for k, v in arrays.items() if v is not None]
    for k, v in items:
        if v.dtype.byteorder == ">" or v.dtype.kind not in "fiu":
            raise ValueError(f"{k}: unsupported dtype {v.dtype}")
    # two passes: header length depends on the offsets, offsets depend on the header length -> reserve a fixed header block
    entries = [{"name": k, "dtype": v.dtype.newbyteorder("<").str, "shape": list(v.shape), "offset": 0} for k, v in items]
    hdr = dict(header, kind=kind, arrays=entries)
    reserve = _align(16 + len(json.dumps(hdr).encode()) + 32 * len(entries) + 256)
    off = reserve
    for e, (_, v) in zip(entries, items):
        e["offset"] = off
        off = _align(off + v.nbytes)
    blob = json.dumps(hdr).encode()
    assert 16 + len(blob) <= reserve
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for e, (_, v) in zip(entries, items):
            f.seek(e["offset"])
            f.write(v.astype(v.dtype.newbyteorder("<"), copy=False).tobytes(order="C"))
        f.truncate(off)
    os.replace(tmp, path)


class Container:
    """A memory-mapped container: `.header` (dict) and `.arrays[name]` (read-only np.memmap views, zero-copy)."""

    def __init__(self, path: str, mode: str = "r"):
        path = os.fspath(path)
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{path}: not an MCB200 container")
            (n,) = struct.unpack("<Q", f.read(8))
            self.header = json.loads(f.read(n).decode())
        self.path, self.kind = path, self.header["kind"]
        self.arrays = {e["name"]: np.memmap(path, dtype=np.dtype(e["dtype"]), mode=mode, offset=e["offset"], shape=tuple(e["shape"]))
                       for e in self.header["arrays"]}

    def get(self, name: str):
        return self.arrays.get(name)


def _site_view(a, lo, hi, contiguous):
    if a is None:
        return None
    v = a[..., lo:hi]
    return np.ascontiguousarray(v) if contiguous else v


# ------------------------------------------------------------------------------------------------ sites
def save_sites(path: str, b: M.SitesBatch, **meta) -> None:
    arrays = {f: getattr(b, f) for f in b.SITE_FIELDS + b.NODE_FIELDS}
    arrays["initial_soil_moisture"] = b.initial_soil_moisture
    arrays["initial_soil_temperature"] = b.initial_soil_temperature
    write_container(path, "sites", arrays, S=b.S, n_soil=b.N, meta=meta)


def load_sites(path: str, lo: int = 0, hi: Optional[int] = None) -> M.SitesBatch:
    c = Container(path)
    assert c.kind == "sites"
    hi = c.header["S"] if hi is None else hi
    b = M.SitesBatch.__new__(M.SitesBatch)
    b.N, b.S = c.header["n_soil"], hi - lo
    for f in b.SITE_FIELDS + b.NODE_FIELDS + ("initial_soil_moisture", "initial_soil_temperature"):
        setattr(b, f, _site_view(c.get(f), lo, hi, True))
    return b


# ------------------------------------------------------------------------------------------------ forcing
def save_forcing(path: str, b: M.ForcingBatch, day0: int = 0, **meta) -> None:
    arrays = {f: getattr(b, f) for f in A.FORCING_HOURLY + A.FORCING_DAILY}
    arrays["prescribed_soil_moisture"] = b.prescribed_soil_moisture
    arrays["day_of_year"] = np.asarray(b.day_of_year, dtype=np.int32)
    write_container(path, "forcing", arrays, S=b.S, ndays=b.ndays, day0=day0, meta=meta)


def load_forcing(path: str, d0: int = 0, d1: Optional[int] = None, lo: int = 0, hi: Optional[int] = None, *, contiguous: bool = True) -> M.ForcingBatch:
    """Days [d0, d1) (relative to the file's first day) and sites [lo, hi) of a forcing file.  With the full site range the
    hourly / daily arrays of a day chunk are contiguous file ranges and are returned as memmap views (no copy)."""
    c = Container(path)
    assert c.kind == "forcing"
    d1 = c.header["ndays"] if d1 is None else d1
    hi = c.header["S"] if hi is None else hi
    whole = lo == 0 and hi == c.header["S"]
    b = M.ForcingBatch.__new__(M.ForcingBatch)
    b.ndays, b.S = d1 - d0, hi - lo
    for f in A.FORCING_HOURLY:
        a = c.get(f)
        setattr(b, f, None if a is None else (a[24 * d0:24 * d1] if whole else _site_view(a[24 * d0:24 * d1], lo, hi, contiguous)))
    for f in A.FORCING_DAILY:
        a = c.get(f)
        setattr(b, f, None if a is None else (a[d0:d1] if whole else _site_view(a[d0:d1], lo, hi, contiguous)))
    a = c.get("prescribed_soil_moisture")
    b.prescribed_soil_moisture = None if a is None else (a[d0:d1] if whole else _site_view(a[d0:d1], lo, hi, contiguous))
    b.day_of_year = np.ascontiguousarray(c.get("day_of_year")[d0:d1])
    return b


# ------------------------------------------------------------------------------------------------ min/max weather (populate_weather! inputs)
def save_weather(path: str, win, solar_sites=None, curves=None, humidity_source: int = A.MC_HUMIDITY_CURVE, day0: int = 0, **meta) -> None:
    """One container for what mc_populate_weather / mc_step_days_minmax read: the per-day value series of every curve input (`<quantity>_values_<i>`, [ndays][S]),
    the per-day / hourly solar inputs, pressure, day_of_year, the solar site table ([S]) — and, in the header, the forcing model's structure (the DielCurves are
    model-level: no data) so that a reader rebuilds mc_weather_desc without the writer's code."""
    from . import weather as W
    arrays = {}
    for q in A.WEATHER_QUANTITIES:
        for i, v in enumerate(win.forcings[q]):
            arrays[f"{q}_values_{i}"] = v
    for f in A.WEATHER_IN_DAILY + A.WEATHER_IN_HOURLY + ("pressure_site",):
        arrays[f] = getattr(win, f)
    arrays["day_of_year"] = np.asarray(win.days, dtype=np.int32)
    if solar_sites is not None:
        for f in A.SOLAR_SITE_FIELDS:
            arrays[f] = getattr(solar_sites, f)
    model = None if curves is None else {"curves": {q: W.curve_to_json(curves[q]) for q in A.WEATHER_QUANTITIES}, "humidity_source": int(humidity_source)}
    write_container(path, "weather", arrays, S=win.S, ndays=win.ndays, day0=day0, forcing_model=model, meta=meta)


def load_weather(path: str, d0: int = 0, d1: Optional[int] = None, lo: int = 0, hi: Optional[int] = None):
    """Days [d0, d1) and sites [lo, hi) of a weather container: (WeatherInBatch, SolarSitesBatch or None, curves or None, humidity_source)."""
    from . import weather as W
    c = Container(path)
    assert c.kind == "weather"
    d1 = c.header["ndays"] if d1 is None else d1
    hi = c.header["S"] if hi is None else hi

    def cut(name, k):
        a = c.get(name)
        return None if a is None else _site_view(a[k * d0:k * d1], lo, hi, True)
    forcings = {}
    for q in A.WEATHER_QUANTITIES:
        vals, i = [], 0
        while c.get(f"{q}_values_{i}") is not None:
            vals.append(cut(f"{q}_values_{i}", 1))
            i += 1
        forcings[q] = tuple(vals)
    ps = c.get("pressure_site")
    win = W.WeatherInBatch(np.ascontiguousarray(c.get("day_of_year")[d0:d1]), hi - lo, forcings,
                           direct_horizontal=cut("direct_horizontal", 24), diffuse_horizontal=cut("diffuse_horizontal", 24),
                           hour_solar_noon=cut("hour_solar_noon", 1), hour_angle_sunrise=cut("hour_angle_sunrise", 1),
                           global_radiation_hourly=cut("global_radiation_hourly", 24), pressure_hourly=cut("pressure_hourly", 24),
                           pressure_site=None if ps is None else np.ascontiguousarray(ps[lo:hi]))
    lat = c.get("latitude")
    solar = None if lat is None else W.SolarSitesBatch(*[None if c.get(f) is None else np.ascontiguousarray(c.get(f)[lo:hi]) for f in A.SOLAR_SITE_FIELDS])
    fm = c.header.get("forcing_model")
    curves = None if not fm else {q: W.curve_from_json(fm["curves"][q]) for q in A.WEATHER_QUANTITIES}
    return win, solar, curves, (fm or {}).get("humidity_source", A.MC_HUMIDITY_CURVE)


# ------------------------------------------------------------------------------------------------ outputs
_OUT_FIELDS = A.OUTPUT_NODE_FIELDS + A.OUTPUT_SCALAR_FIELDS + ("snow_temperature",) + A.OUTPUT_PROFILE_FIELDS + A.OUTPUT_PROFILE_SCALARS


def save_outputs(path: str, b: M.OutputBatch, day0: int = 0, **meta) -> None:
    arrays = {f: getattr(b, f, None) for f in _OUT_FIELDS}
    arrays["status"] = b.status
    write_container(path, "output", arrays, S=b.S, n_soil=b.N, ndays=b.ndays, day0=day0, meta=meta)


def load_outputs(path: str, lo: int = 0, hi: Optional[int] = None, *, contiguous: bool = False) -> M.OutputBatch:
    c = Container(path)
    assert c.kind == "output"
    hi = c.header["S"] if hi is None else hi
    b = M.OutputBatch.__new__(M.OutputBatch)
    b.ndays, b.N, b.S = c.header["ndays"], c.header["n_soil"], hi - lo
    for f in _OUT_FIELDS:
        setattr(b, f, _site_view(c.get(f), lo, hi, contiguous))
    b.status = np.ascontiguousarray(c.get("status")[lo:hi])
    return b


# ------------------------------------------------------------------------------------------------ state checkpoints
def save_state(path: str, st: M.StateBatch, **meta) -> None:
    arrays = {f: getattr(st, f) for f, _ in A.StateSoA._fields_}
    n_soil, S = st.soil_temperature.shape
    write_container(path, "state", arrays, S=S, n_soil=n_soil, n_snow=0 if st.snow_temperature is None else st.snow_temperature.shape[0],
                    days_done=int(st.days_done[0]), meta=meta)


def load_state(path: str, lo: int = 0, hi: Optional[int] = None) -> M.StateBatch:
    c = Container(path)
    assert c.kind == "state"
    hi = c.header["S"] if hi is None else hi
    st = M.StateBatch(c.header["n_soil"], c.header["n_snow"], hi - lo)
    for f, _ in A.StateSoA._fields_:
        a = c.get(f)
        if f == "days_done":
            st.days_done[0] = int(a[0])
        elif f == "status":
            if a is not None:                       # checkpoints written before the status word travelled with the state have none: flags start at 0
                st.status[...] = a[lo:hi]
        elif a is not None:
            getattr(st, f)[...] = a[..., lo:hi]
    return st
