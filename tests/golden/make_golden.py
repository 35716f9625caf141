"""Generate the committed golden fixtures from the reference's own test data.

Run ONCE in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Reads the NicheMapR fixture CSVs the reference's tests read (test/micro_testrun_monthly.jl:11-153,
test/micro_testrun_daily.jl:11-180) and writes compact .npz files next to this script:
  config1_monthly.npz  inputs + NicheMapR goldens of config 1 (12 non-consecutive days, N=10)
  config2_daily.npz    inputs + NicheMapR goldens of config 2 (30 consecutive days, N=19, dynamic moisture)
  config4_scan329.npz  inputs + NicheMapR goldens of config 4 (SCAN 329: 365 consecutive days, N=19 + 8 snow nodes, dynamic moisture, snow)
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


def config4():
    """BASELINE config 4, SCAN/SNOTEL 329 (test/micro_testrun_snow_scan329.jl:1-238): N=19 + 8 snow nodes, 365 consecutive days of 2013, dynamic
    moisture, DailyRainfall, snow.  The reference expands daily min/max to hourly with its diel engine (pre-processing outside the hot path); as for
    config 1 the hourly forcing here is NicheMapR's own expansion found in the golden output (metout.csv TAREF/RH/VREF/SOLR/ZEN); cloud cover is
    held at the day's mean of (CCmin, CCmax) because metout carries no hourly cloud column.  Goldens: soil.csv, soilmoist.csv, metout.csv snow columns."""
    d = f"{REF}/scan329"
    met = pd.read_csv(f"{d}/metout.csv")
    soil = pd.read_csv(f"{d}/soil.csv")
    moist = pd.read_csv(f"{d}/soilmoist.csv")
    sun = pd.read_csv(f"{d}/sunsnow.csv")
    frc = pd.read_csv(f"{d}/nmr_forcing.csv")
    ini = pd.read_csv(f"{d}/nmr_initial.csv")
    par = pd.read_csv(f"{d}/nmr_params.csv").iloc[0]
    nd = 365
    H = 24 * nd
    sm10 = np.array([ini[f"SM{i}"][0] for i in range(1, 11)], float)
    st = np.full(19, 274.15)
    st[18] = 273.15 + 6.4; st[17] = 273.15 + (6.4 + 3.0) / 2.0; st[16] = st[15] = st[14] = 273.15 + 3.0          # script lines 84-89
    cc = 0.5 * (frc["CCmin_pct"].to_numpy(float)[:nd] + frc["CCmax_pct"].to_numpy(float)[:nd]) / 100.0
    out = dict(
        days=frc["doy"].to_numpy(int)[:nd],
        depths=np.array([0.0, 1.25, 2.5, 3.75, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0, 25.0, 30.0, 40.0, 50.0, 75.0, 100.0, 150.0, 200.0]) / 100.0,
        heights=np.array([0.01, 2.0]), roughness_height=float(par["RUF"]), elevation=float(par["elevation_m"]), albedo=float(par["REFL"]),
        emissivity=float(par["SLE"]), bulk_density=1300.0, mineral_density=2560.0,
        mineral_conductivity=np.array([0.2, 0.2, 0.2, 1.35] + [2.5] * 15), mineral_heat_capacity=np.array([1920.0, 1920.0, 1920.0, 1395.0] + [870.0] * 15),
        de_vries_shape_factor=0.0167238, air_entry_water_potential=1.1, saturated_hydraulic_conductivity=0.0037, campbell_b_parameter=4.5,
        root_density=np.array([0.0, 0.0, 82000.0, 80000.0, 78000.0, 74000.0, 71000.0, 64000.0, 58000.0, 48000.0, 40000.0, 18000.0, 9000.0, 6000.0, 8000.0,
                               4000.0, 4000.0, 0.0, 0.0]),
        initial_soil_moisture=expand_fine(sm10), initial_soil_temperature=st,
        rainfall=frc["Rain_mm"].to_numpy(float)[:nd], deep_soil_temperature=soil["D200cm"].to_numpy(float)[0:H:24] + 273.15,
        pressure=pressure_from_elevation(float(par["elevation_m"])),
        snowtemp=float(par["snowtemp"]), snowdens=float(par["snowdens"]), snowmelt=float(par["snowmelt"]), undercatch=float(par["undercatch"]),
        rainmelt=float(par["rainmelt"]), densfun=np.array([par["densfun1"], par["densfun2"], par["densfun3"], par["densfun4"]], float),
        snowcond=float(par["snowcond"]), intercept=float(par["intercept"]),
        TAREF=met["TAREF"].to_numpy(float)[:H] + 273.15, RH=np.clip(met["RH"].to_numpy(float)[:H], 0, 100) / 100.0,
        VREF=met["VREF"].to_numpy(float)[:H], SOLR=met["SOLR"].to_numpy(float)[:H], ZEN=met["ZEN"].to_numpy(float)[:H], cloud=np.repeat(cc, 24),
        # goldens (float32 keeps the fixture small; the comparisons are at 1e-1 relative / kelvin level)
        soil_nmr=(soil.iloc[:H, 2:12].to_numpy(float) + 273.15).astype(np.float32), soilmoist_nmr=moist.iloc[:H, 2:12].to_numpy(float).astype(np.float32),
        SNOWDEP=met["SNOWDEP"].to_numpy(float)[:H].astype(np.float32), SNOWFALL=met["SNOWFALL"].to_numpy(float)[:H].astype(np.float32),
        SNOWDENS=met["SNOWDENS"].to_numpy(float)[:H].astype(np.float32), TSKYC=(met["TSKYC"].to_numpy(float)[:H] + 273.15).astype(np.float32),
        sunsnow=(sun.iloc[:H, 2:11].to_numpy(float) + 273.15).astype(np.float32),
    )
    for k in ("TAREF", "RH", "VREF", "SOLR", "ZEN"):
        out[k] = out[k].astype(np.float64)
    np.savez_compressed(f"{OUT}/config4_scan329.npz", **out)
    print("wrote config4_scan329.npz", os.path.getsize(f"{OUT}/config4_scan329.npz") // 1024, "KiB")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference test data not found (this script only runs in the build container)")
    config1()
    config1(tag="monthly_snow", init="init_monthly_snow", metout="metout_monthly_snow.csv", soil="soil_monthly_snow.csv")
    config2()
    config4()
