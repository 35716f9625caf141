"""Generate the committed golden fixtures from the reference's own test data.

Run ONCE in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Reads the NicheMapR fixture CSVs the reference's tests read (test/micro_testrun_monthly.jl:11-153,
test/micro_testrun_daily.jl:11-180) and writes compact .npz files next to this script:
  config1_monthly.npz  inputs + NicheMapR goldens of config 1 (12 non-consecutive days, N=10)
  config2_daily.npz    inputs + NicheMapR goldens of config 2 (30 consecutive days, N=19, dynamic moisture)
Hourly forcing for config 1 is NicheMapR's own expansion found in the golden output
(metout_monthly.csv TAREF/RH/VREF/SOLR/ZEN) because the diel engine / solar model are pre-processing
outside the hot path.
"""
import os
import sys

import numpy as np
import pandas as pd

REF = "/root/reference/test/data"
OUT = os.path.dirname(os.path.abspath(__file__))

MICROINPUT_NAMES = [
    "doynum", "RUF", "ERR", "Usrhyt", "Refhyt", "Numtyps", "Z01", "Z02", "ZH1", "ZH2",
    "idayst", "ida", "HEMIS", "ALAT", "AMINUT", "ALONG", "ALMINT", "ALREF", "slope",
    "azmuth", "ALTT", "CMH2O", "microdaily", "tannul", "EC", "VIEWF", "snowtemp",
    "snowdens", "snowmelt", "undercatch", "rainmult", "runshade", "runmoist",
    "maxpool", "evenrain", "snowmodel", "rainmelt", "writecsv", "densfun1", "densfun2",
    "densfun3", "densfun4", "hourly", "rainhourly", "lamb", "IUV", "RW", "PC", "RL", "SP", "R1",
    "IM", "MAXCOUNT", "IR", "message", "fail", "snowcond", "intercept", "grasshade", "solonly",
    "ZH", "D0", "TIMAXS1", "TIMAXS2", "TIMAXS3", "TIMAXS4", "TIMINS1", "TIMINS2", "TIMINS3", "TIMINS4",
    "spinup", "dewrain", "moiststep", "maxsurf", "ndmax"]


def col(path, c=1):
    return pd.read_csv(path).iloc[:, c].to_numpy(float)


def microinput(d):
    v = col(f"{REF}/{d}/microinput.csv")
    return dict(zip(MICROINPUT_NAMES, v))


def pressure_from_elevation(h):
    return 101325.0 * (1.0 - 0.0065 * h / 288.0) ** (1.0 / 0.190284)


def expand_fine(coarse):
    n = len(coarse)
    out = np.empty(2 * n - 1)
    out[0::2] = coarse
    out[1::2] = 0.5 * (coarse[:-1] + coarse[1:])
    return out


def config1(tag="monthly", init="init_monthly", metout="metout_monthly.csv", soil="soil_monthly.csv"):
    mi = microinput(init)
    met = pd.read_csv(f"{REF}/{metout}", na_values="NA")
    soil_nmr = pd.read_csv(f"{REF}/{soil}").iloc[:, 3:13].to_numpy(float)          # degC, 288 x 10
    sp = pd.read_csv(f"{REF}/{init}/soilprop.csv").iloc[0, 1:].to_numpy(float)      # BD, ?, cond, cp, dens
    ndays = 12
    cc = col(f"{REF}/{init}/CCMINN.csv")[:ndays] / 100.0
    d = dict(
        days=np.array([15, 46, 74, 105, 135, 166, 196, 227, 258, 288, 319, 349]),
        depths=col(f"{REF}/{init}/DEP.csv") / 100.0,
        heights=np.array([mi["Usrhyt"], mi["Refhyt"]]),
        roughness_height=mi["RUF"], slope=mi["slope"], elevation=mi["ALTT"],
        albedo=col(f"{REF}/{init}/REFLS.csv")[0],
        bulk_density=sp[0] * 1000.0, mineral_conductivity=sp[2], mineral_heat_capacity=sp[3], mineral_density=sp[4] * 1000.0,
        moists=pd.read_csv(f"{REF}/{init}/moists.csv").iloc[:, 1:13].to_numpy(float),      # 10 x 12
        shade=col(f"{REF}/{init}/Minshades.csv")[:ndays] / 100.0,
        soil_wetness=col(f"{REF}/{init}/PCTWET.csv")[:ndays] / 100.0,
        surface_emissivity=col(f"{REF}/{init}/SLES.csv")[:ndays],
        rainfall=col(f"{REF}/{init}/rain.csv")[:ndays],
        deep_soil_temperature=col(f"{REF}/{init}/tannulrun.csv")[:ndays] + 273.15,
        iterations_per_day=int(mi["ndmax"]), maxsurf=mi["maxsurf"],
        pressure=pressure_from_elevation(mi["ALTT"]),
        # NicheMapR's hourly expansion (golden output doubles as forcing)
        TAREF=met["TAREF"].to_numpy(float) + 273.15, RH=met["RH"].to_numpy(float) / 100.0,
        VREF=met["VREF"].to_numpy(float), SOLR=met["SOLR"].to_numpy(float), ZEN=met["ZEN"].to_numpy(float),
        cloud=np.repeat(cc, 24),
        # goldens
        soil_nmr=soil_nmr + 273.15, TSKYC=met["TSKYC"].to_numpy(float) + 273.15,
        TALOC=met["TALOC"].to_numpy(float) + 273.15, RHLOC=met["RHLOC"].to_numpy(float) / 100.0,
        VLOC=met["VLOC"].to_numpy(float),
    )
    if tag != "monthly":
        for k in ("snowtemp", "snowdens", "snowmelt", "undercatch", "rainmult", "rainmelt", "densfun1", "densfun2",
                  "densfun3", "densfun4", "snowcond", "intercept"):
            d[k] = mi[k]
        d["SNOWFALL"] = met["SNOWFALL"].to_numpy(float)
        d["SNOWDEP"] = met["SNOWDEP"].to_numpy(float)
        d["SNOWDENS"] = met["SNOWDENS"].to_numpy(float)
    np.savez_compressed(f"{OUT}/config1_{tag}.npz", **d)
    print("wrote", f"config1_{tag}.npz", {k: np.shape(v) for k, v in d.items() if np.ndim(v)})


def config2():
    init = "init_daily"
    mi = microinput(init)
    ndays, H = 30, 720
    met = pd.read_csv(f"{REF}/metout_FordDryLake.csv")

    def gold(name):
        return pd.read_csv(f"{REF}/{name}_FordDryLake.csv").iloc[:H, 4:14].to_numpy(float)
    sp = pd.read_csv(f"{REF}/{init}/soilprop.csv").iloc[0, 1:].to_numpy(float)
    coarse_depths = col(f"{REF}/{init}/DEP.csv") / 100.0

    def hr(name):
        return pd.read_csv(f"{REF}/{init}/{name}.csv")["x"].to_numpy(float)[:H]
    d = dict(
        days=np.arange(1, ndays + 1), coarse_depths=coarse_depths, depths=expand_fine(coarse_depths),
        heights=np.array([mi["Usrhyt"], mi["Refhyt"]]),
        roughness_height=mi["RUF"], slope=mi["slope"], elevation=mi["ALTT"], albedo=col(f"{REF}/{init}/REFLS.csv")[0],
        bulk_density=col(f"{REF}/{init}/BD.csv") * 1000.0, mineral_density=col(f"{REF}/{init}/DD.csv") * 1000.0,
        mineral_conductivity=sp[2], mineral_heat_capacity=sp[3],
        air_entry_water_potential=col(f"{REF}/{init}/PE.csv"), saturated_hydraulic_conductivity=col(f"{REF}/{init}/KS.csv"),
        campbell_b_parameter=col(f"{REF}/{init}/BB.csv"), root_density=col(f"{REF}/{init}/L.csv"),
        root_resistance=mi["RW"], stomatal_closure_potential=-mi["PC"], leaf_resistance=mi["RL"],
        stomatal_stability_parameter=mi["SP"], root_radius=mi["R1"],
        moisture_tolerance=mi["IM"], moisture_max_iterations=int(mi["MAXCOUNT"]), moisture_timestep=mi["moiststep"],
        max_surface_pool=mi["maxpool"] * 1000.0, ndmax=int(mi["ndmax"]), spinup=int(mi["spinup"]),
        rainhourly=int(mi["rainhourly"]),
        shade=col(f"{REF}/{init}/Minshades.csv")[:ndays] / 100.0,
        soil_wetness=col(f"{REF}/{init}/PCTWET.csv")[:ndays] / 100.0,
        surface_emissivity=col(f"{REF}/{init}/SLES.csv")[:ndays],
        rainfall_daily=col(f"{REF}/{init}/rain.csv")[:ndays],
        deep_soil_temperature=col(f"{REF}/{init}/tannulrun.csv")[:ndays] + 273.15,
        leaf_area_index=col(f"{REF}/{init}/LAI.csv")[:ndays],
        pressure=pressure_from_elevation(mi["ALTT"]),
        TAIRhr=hr("TAIRhr") + 273.15, RHhr=np.clip(hr("RHhr"), 0, 100) / 100.0, WNhr=np.maximum(hr("WNhr"), 0.1),
        SOLRhr=hr("SOLRhr"), CLDhr=np.clip(hr("CLDhr"), 0, 100) / 100.0, RAINhr=np.maximum(hr("RAINhr"), 0.0),
        ZEN=met["ZEN"].to_numpy(float)[:H],
        initial_soil_temperature=expand_fine(col(f"{REF}/{init}/soilinit.csv") + 273.15),
        initial_soil_moisture=expand_fine(pd.read_csv(f"{REF}/{init}/moists.csv").iloc[:, 1].to_numpy(float)),
        # goldens
        soil_nmr=gold("soil") + 273.15, soilmoist_nmr=gold("soilmoist"), soilpot_nmr=gold("soilpot"),
        tcond_nmr=gold("tcond"), specheat_nmr=gold("specheat"), densit_nmr=gold("densit"),
        TAREF=met["TAREF"].to_numpy(float)[:H] + 273.15, TALOC=met["TALOC"].to_numpy(float)[:H] + 273.15,
        RH=met["RH"].to_numpy(float)[:H] / 100.0, RHLOC=met["RHLOC"].to_numpy(float)[:H] / 100.0,
        VREF=met["VREF"].to_numpy(float)[:H], VLOC=met["VLOC"].to_numpy(float)[:H],
        TSKYC=met["TSKYC"].to_numpy(float)[:H] + 273.15,
    )
    np.savez_compressed(f"{OUT}/config2_daily.npz", **d)
    print("wrote config2_daily.npz", {k: np.shape(v) for k, v in d.items() if np.ndim(v)})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference test data not found (this script only runs in the build container)")
    config1()
    config1(tag="monthly_snow", init="init_monthly_snow", metout="metout_monthly_snow.csv", soil="soil_monthly_snow.csv")
    config2()
