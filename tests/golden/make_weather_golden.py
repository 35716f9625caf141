"""Generate tests/golden/weather_golden.npz from the reference's own test data (widening rows f2 / f3).

Run ONCE in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_weather_golden.py
Contents:
  zen_<set>_{doy,hour,zen}, zen_<set>_lat   the NicheMapR ZEN column (hourly zenith angle, degrees, 90 = below the horizon) of
        test/data/metout_monthly.csv (288 rows, Madison 43.07305 N), metout_FordDryLake.csv (720 rows, 33.6547 N) and
        scan329/metout.csv (8760 rows, 39.1368 N): the anchors of the built-in solar geometry (SolarRadiation.jl is un-vendored);
  monthly_*   config 1's min/max inputs (init_monthly/{TMINN,TMAXX,WNMINN,WNMAXX,RHMINN,RHMAXX,CCMINN,CCMAXX}.csv, as the reference's
        test reads them, test/micro_testrun_monthly.jl:93-102) and NicheMapR's own hourly expansion of them (metout_monthly.csv
        TAREF / RH / VREF / SOLR), which the reference's test compares the solved 2 m air temperature / humidity / wind against
        (test/micro_testrun_monthly.jl:172-180);
  scan329_*   config 4's daily min/max forcing (scan329/nmr_forcing.csv: Tmin / Tmax, RH min / max, one wind value, cloud max / min, 365 days).
"""
import os

import numpy as np
import pandas as pd

REF = "/root/reference/test/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def col(path, c=1):
    return pd.read_csv(path).iloc[:, c].to_numpy(float)


def main():
    g = {}
    for name, path, lat in (("monthly", "metout_monthly.csv", col(f"{REF}/init_monthly/longlat.csv")[1]),
                            ("daily", "metout_FordDryLake.csv", col(f"{REF}/init_daily/longlat.csv")[1]),
                            ("scan329", "scan329/metout.csv", float(pd.read_csv(f"{REF}/scan329/nmr_params.csv").latitude[0]))):
        m = pd.read_csv(f"{REF}/{path}")
        g[f"zen_{name}_doy"] = m.DOY.to_numpy(np.int32)
        g[f"zen_{name}_hour"] = (m.TIME.to_numpy(float) / 60.0)
        g[f"zen_{name}_zen"] = m.ZEN.to_numpy(float)
        g[f"zen_{name}_lat"] = np.float64(lat)
    d = f"{REF}/init_monthly"
    g["monthly_days"] = np.array([15, 46, 74, 105, 135, 166, 196, 227, 258, 288, 319, 349], dtype=np.int32)
    for k, f, scale in (("tmin_C", "TMINN", 1.0), ("tmax_C", "TMAXX", 1.0), ("wind_min", "WNMINN", 1.0), ("wind_max", "WNMAXX", 1.0),
                        ("rh_min", "RHMINN", 0.01), ("rh_max", "RHMAXX", 0.01), ("cloud_min", "CCMINN", 0.01), ("cloud_max", "CCMAXX", 0.01)):
        g["monthly_" + k] = col(f"{d}/{f}.csv")[:12] * scale
    m = pd.read_csv(f"{REF}/metout_monthly.csv")
    g["monthly_TAREF_C"], g["monthly_RH_pct"], g["monthly_VREF"], g["monthly_SOLR"] = (m[c].to_numpy(float) for c in ("TAREF", "RH", "VREF", "SOLR"))
    # config 4 (SCAN 329): the daily min/max forcing the reference's script feeds its DailyMinMaxEnvironment (test/micro_testrun_snow_scan329.jl:93-110;
    # a single wind value is used for both min and max there)
    f = pd.read_csv(f"{REF}/scan329/nmr_forcing.csv")
    g["scan329_doy"] = f.doy.to_numpy(np.int32)
    for k, c, scale in (("tmin_C", "Tmin_C", 1.0), ("tmax_C", "Tmax_C", 1.0), ("rh_min", "RHmin_pct", 0.01), ("rh_max", "RHmax_pct", 0.01), ("wind", "Wind_ms", 1.0),
                        ("cloud_max", "CCmax_pct", 0.01), ("cloud_min", "CCmin_pct", 0.01)):
        g["scan329_" + k] = f[c].to_numpy(float) * scale
    np.savez_compressed(os.path.join(OUT, "weather_golden.npz"), **g)
    print({k: np.shape(v) for k, v in g.items()})


if __name__ == "__main__":
    main()
