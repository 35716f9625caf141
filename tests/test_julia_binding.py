"""The Julia binding (julia/MicroclimateB200.jl) cannot run in this container (no Julia).  What CAN rot silently is its POD mirrors of the C
header and the entry points it ccalls, so both are checked here by parsing the Julia source: every struct must have the header's fields in the
header's order with the same sizes (via the ctypes mirror _abi.py, which tests/test_abi.py pins against a C compile of the header), and every
ccall'd symbol must be exported by the library and declared in include/microclimate_b200.h."""
import ctypes as C
import os
import re

from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "julia", "MicroclimateB200.jl")).read()
CONSTS = {"MC_MAX_SOIL_NODES": A.MC_MAX_SOIL_NODES, "MC_SNOW_NODES": A.MC_SNOW_NODES, "MC_MAX_HEIGHTS": A.MC_MAX_HEIGHTS,
          "MC_DIEL_MAX_SHAPES": A.MC_DIEL_MAX_SHAPES, "MC_DIEL_MAX_INPUTS": A.MC_DIEL_MAX_INPUTS}
SIZES = {"Cdouble": 8, "Cint": 4, "Ptr{Cdouble}": 8, "Ptr{Cint}": 8}


def julia_struct(name):
    """[(field, type string)] of `struct name ... end` (fields separated by ';' or newlines, comments stripped)."""
    m = re.search(r"^struct\s+%s\b;?(.*?)\bend\b" % name, SRC, re.S | re.M)
    assert m, f"struct {name} not found in MicroclimateB200.jl"
    body = re.sub(r"#.*", "", m.group(1))
    fields = []
    for part in re.split(r"[;\n]", body):
        part = part.strip()
        if part:
            fname, ftype = part.split("::")
            fields.append((fname.strip(), ftype.strip()))
    return fields


def julia_sizeof(t, structs):
    m = re.fullmatch(r"NTuple\{(\w+),\s*(.+)\}", t)
    if m:
        n = CONSTS.get(m.group(1)) or int(m.group(1))
        return n * julia_sizeof(m.group(2), structs)[0], n
    if t in structs:
        return C.sizeof(structs[t]), 1
    return SIZES[t], 1


def flat(ct):
    """(name, size, count) of every field of a ctypes Structure, in order."""
    out = []
    for n, t in ct._fields_:
        out.append((n, C.sizeof(t), getattr(t, "_length_", 1)))
    return out


def test_julia_pod_structs_match_the_header_field_by_field():
    structs = {"McFluid": A.FluidConstants, "McSnow": A.SnowDesc}
    for jname, ct in (("McFluid", A.FluidConstants), ("McSnow", A.SnowDesc), ("McModel", A.ModelDesc)):
        jf = julia_struct(jname)
        cf = flat(ct)
        assert [n for n, _ in jf] == [n for n, _, _ in cf], f"{jname}: field names / order differ from the header"
        for (n, t), (_, size, count) in zip(jf, cf):
            js, jc = julia_sizeof(t, structs)
            assert js == size and (jc == count or count == 1), f"{jname}.{n}: Julia {t} is {js} bytes, header has {size}"
        # isbits Julia structs use C layout rules: same field sizes in the same order => same offsets and total size
        assert sum(julia_sizeof(t, structs)[0] for _, t in jf) <= C.sizeof(ct)


def test_julia_pointer_tables_have_the_header_arity():
    def arity(jname):
        return sum(julia_sizeof(t, {})[0] // 8 for _, t in julia_struct(jname))
    assert arity("McSites") == len(A.SitesSoA._fields_) == 12
    assert arity("McInitial") == len(A.InitialSoA._fields_) == 2
    assert arity("McForcing") == len(A.ForcingSoA._fields_) == len(A.FORCING_HOURLY) + len(A.FORCING_DAILY) + 2
    assert arity("McOutput") == len(A.OutputSoA._fields_)
    f = dict(julia_struct("McForcing"))
    assert f["hourly"] == "NTuple{11,Ptr{Cdouble}}" and f["daily"] == "NTuple{8,Ptr{Cdouble}}" and f["doy"] == "Ptr{Cint}"
    o = dict(julia_struct("McOutput"))
    assert o["node"] == "NTuple{7,Ptr{Cdouble}}" and o["scalar"] == "NTuple{5,Ptr{Cdouble}}" and o["status"] == "Ptr{Cint}"


def test_julia_ccalls_only_exported_declared_symbols():
    called = set(re.findall(r"\(\s*:(mc_\w+)\s*,\s*LIB\s*\)", SRC))
    called |= {s for s in re.findall(r":(mc_\w+)", SRC)}          # symbols chosen at run time (f_set / f_async / f_wait)
    header = open(os.path.join(ROOT, "include", "microclimate_b200.h")).read()
    assert {"mc_create", "mc_multi_create", "mc_step_days_async", "mc_multi_step_days_async", "mc_wait", "mc_multi_wait", "mc_set_sites"} <= called
    for sym in called:
        assert sym in _lib.EXPORTS, f"{sym} is ccall'd by the Julia binding but not exported"
        assert re.search(r"\b%s\s*\(" % sym, header), f"{sym} is not declared in the header"


def test_julia_binding_marshals_what_round_1_dropped():
    """prescribed moisture, daily albedo, supplied longwave, the snow descriptor, Swinbank, chunk_days and several devices."""
    for needle in ("precomputed_soil_moisture", "environment_daily.albedo", "longwave_radiation", "snow_desc(model.snow_model, inputs0)",
                   "SwinbankAtmosphericRadiation", "chunk_days", "mc_multi_create", "snow_temperature"):
        assert needle in SRC, needle
    assert "m.fluid, m.snow)" not in SRC          # the zeroed default snow descriptor must not be passed for a SnowModel


def test_julia_weather_structs_match_the_header():
    """The widening structs (populate_weather! / solar geometry): nested PODs field by field, pointer tables by arity and total size."""
    structs = {"McTimeOfDay": A.TimeOfDay, "McDielShape": A.DielShape, "McDielCurve": A.DielCurve}
    for jname, ct in (("McTimeOfDay", A.TimeOfDay), ("McDielShape", A.DielShape), ("McDielCurve", A.DielCurve), ("McWeather", A.WeatherDesc)):
        jf, cf = julia_struct(jname), flat(ct)
        assert [n for n, _ in jf] == [n for n, _, _ in cf], f"{jname}: field names / order differ from the header"
        for (n, t), (_, size, _count) in zip(jf, cf):
            assert julia_sizeof(t, structs)[0] == size, f"{jname}.{n}"
        assert sum(julia_sizeof(t, structs)[0] for _, t in jf) == C.sizeof(ct), jname          # no padding holes in these structs
    for jname, ct in (("McSolarSites", A.SolarSitesSoA), ("McWeatherIn", A.WeatherInSoA), ("McWeatherOut", A.WeatherOutSoA)):
        assert sum(julia_sizeof(t, {})[0] for _, t in julia_struct(jname)) == C.sizeof(ct), jname
    wi = dict(julia_struct("McWeatherIn"))
    assert wi["values"] == "NTuple{24,Ptr{Cdouble}}" and 24 == 4 * A.MC_DIEL_MAX_INPUTS and wi["doy"] == "Ptr{Cint}"
    assert dict(julia_struct("McWeatherOut"))["hourly"] == "NTuple{9,Ptr{Cdouble}}" and len(A.WEATHER_OUT_HOURLY) == 9
    # the enum values the Julia marshaller hard-codes (tod / shape)
    for pat, val in ((r"tod\(t::ClockTime\) = McTimeOfDay\((\d)", A.MC_TOD_CLOCK), (r"tod\(t::Sunrise\) = McTimeOfDay\((\d)", A.MC_TOD_SUNRISE),
                     (r"tod\(t::Sunset\) = McTimeOfDay\((\d)", A.MC_TOD_SUNSET), (r"tod\(t::Midday\) = McTimeOfDay\((\d)", A.MC_TOD_MIDDAY),
                     (r"tod\(t::Midnight\) = McTimeOfDay\((\d)", A.MC_TOD_MIDNIGHT), (r"shape\(s::Sine\) = McDielShape\((\d)", A.MC_SHAPE_SINE),
                     (r"shape\(s::Decay\) = McDielShape\((\d)", A.MC_SHAPE_DECAY), (r"shape\(s::Linear\) = McDielShape\((\d)", A.MC_SHAPE_LINEAR)):
        assert int(re.search(pat, SRC).group(1)) == val, pat
    assert "mc_populate_weather" in set(re.findall(r"\(\s*:(mc_\w+)\s*,\s*LIB\s*\)", SRC)) and "mc_populate_weather" in _lib.EXPORTS
