"""The C-ABI library loads on a machine without a GPU, exports every symbol include/microclimate_b200.h declares,
mirrors the reference's defaults, and FAILS LOUDLY (never falls back) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import microclimate_jl_b200 as mc
from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import _lib
from microclimate_jl_b200 import marshal as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    return _lib.lib()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "microclimate_b200.h")).read()
    declared = set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mc_handle"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.mc_abi_version() == A.MC_ABI_VERSION == 2 and L.mc_backend_name() == b"cuda-sm_100a"


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors (microclimate.jl_b200/_abi.py) vs sizes computed by the C compiler from the header."""
    import subprocess, tempfile
    names = ["mc_fluid_constants", "mc_snow_desc", "mc_model_desc", "mc_sites_soa", "mc_initial_soa", "mc_state_soa", "mc_forcing_soa", "mc_output_soa", "mc_counters",
             "mc_time_of_day", "mc_diel_shape", "mc_diel_curve", "mc_weather_desc", "mc_solar_sites_soa", "mc_weather_in_soa", "mc_weather_out_soa"]
    mirrors = [A.FluidConstants, A.SnowDesc, A.ModelDesc, A.SitesSoA, A.InitialSoA, A.StateSoA, A.ForcingSoA, A.OutputSoA, A.Counters,
               A.TimeOfDay, A.DielShape, A.DielCurve, A.WeatherDesc, A.SolarSitesSoA, A.WeatherInSoA, A.WeatherOutSoA]
    # sizes AND the offset of every field, computed by the C compiler from the header
    lines = []
    for n, m in zip(names, mirrors):
        lines.append('printf("%%zu", sizeof(%s));' % n)
        for fname, _ in m._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (n, fname))
        lines.append('printf("\\n");')
    src = '#include "microclimate_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){%s return MC_ABI_VERSION == %d ? 0 : 1;}\n' % ("".join(lines), A.MC_ABI_VERSION)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        rows = [[int(x) for x in l.split()] for l in subprocess.check_output([os.path.join(d, "s")]).decode().splitlines()]
    for n, m, row in zip(names, mirrors, rows):
        assert row[0] == C.sizeof(m), n
        assert row[1:] == [getattr(m, fname).offset for fname, _ in m._fields_], n


def test_default_model_mirrors_reference_defaults(L):
    """mc_default_model == MicroModel()/MicroConfig() (reference src/types.jl:71-120, src/constants.jl:8)."""
    d = A.ModelDesc()
    L.mc_default_model(C.byref(d))
    py = M.model_desc(mc.MicroModel(), mc.NonConsecutiveDayMode())
    assert d.n_soil == py.n_soil == 19 and d.n_snow == 0
    assert np.allclose(list(d.depths_m)[:19], mc.DEFAULT_DEPTHS, rtol=0, atol=0)
    for f in ("reference_height_m", "local_height_m", "reltol", "abstol", "maximum_surface_temperature_K", "de_vries_shape_factor",
              "recirculation_power", "return_flow_threshold", "root_resistance", "stomatal_closure_potential", "leaf_resistance",
              "stomatal_stability_parameter", "root_radius", "karman_constant", "dyer_constant", "max_surface_pool",
              "moisture_tolerance", "moisture_max_iterations", "moisture_timestep_s", "rainfall_schedule", "moisture_strategy",
              "convergence_kind", "convergence_iterations", "atmospheric_radiation", "infiltration_algorithm"):
        assert getattr(d, f) == getattr(py, f), f
    for f, _ in A.FluidConstants._fields_:
        assert getattr(d.fluid, f) == getattr(py.fluid, f), f


def test_fails_loudly_without_a_gpu_and_on_bad_models(L):
    import torch
    d = A.ModelDesc()
    L.mc_default_model(C.byref(d))
    h = C.c_void_p()
    if not torch.cuda.is_available():
        rc = L.mc_create(C.byref(d), 8, 0, C.byref(h))
        assert rc == -2 and not h.value
        assert b"no CUDA device" in L.mc_last_error(None) and b"no CPU path" in L.mc_last_error(None)
        with pytest.raises(RuntimeError, match="no CPU path"):
            mc.SoilEngine(d, 8)
    d.n_soil = 2
    assert L.mc_create(C.byref(d), 8, 0, C.byref(h)) == -1 and b"n_soil" in L.mc_last_error(None)
    d.n_soil = 19
    d.n_snow = 3
    assert L.mc_create(C.byref(d), 8, 0, C.byref(h)) == -3 and b"n_snow" in L.mc_last_error(None)
    assert L.mc_step_days(None, 0, 1, None, None) == -1


def test_product_does_not_reach_into_test_infrastructure():
    """Nothing under microclimate.jl_b200/ may import, link or execute oracle/ or tests/hostemu (they are checkers)."""
    pkg = os.path.join(ROOT, "microclimate.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "libmc_oracle" not in txt and "mc_oracle" not in txt and "libmc_hostemu" not in txt and "from oracle" not in txt and "import oracle" not in txt, fn


def test_every_export_is_documented_against_the_reference_interface():
    """INTEGRATION.md section 1 maps each C-ABI entry point to the reference interface it replaces: no exported symbol may be missing from it."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert [e for e in _lib.EXPORTS if e not in doc] == []
