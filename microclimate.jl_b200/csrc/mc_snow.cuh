/*
 * mc_snow.cuh - SnowModel{8}: snow state, apparent heat capacity, mass/energy bookkeeping and the snow-aware RHS.
 * Part of mc_device.cuh (included from there, inside namespace mcd, in this order); not a stand-alone header.
 */

/* ------------------------------------------------------------------------------------------
 * Snow: SnowModel{8} (src/snow/snow_model.jl) + apparent-heat-capacity closures (src/apparent_heat_capacity/*.jl).
 * Depths in cm, densities in g/cm3, as upstream.  The reference's snow_depth_hourly[nsteps] scratch array is only read
 * at [step] and [step-1]; it is carried as two scalars (sd_cur, sd_prev).
 * ------------------------------------------------------------------------------------------ */
constexpr int NSN = MC_SNOW_NODES;
struct SnowState { double depth, age, days_since, density, prev_density, cum_melt, extra, sum_phase; int active; };
struct SnowScratch { double node_depths[NSN]; double sd_prev, sd_cur; double meanT[NSN], meanTpast[NSN]; };
struct SnowCtx { SnowState st; SnowScratch sc; double depths[MC_MAX_ODE_NODES + 1]; double Q_freeze; const SnowState* rhs_state; };

MC_DEV double mc_tanh(double x) { const double e = mcm::exp(2.0 * x); return x > 20.0 ? 1.0 : (x < -20.0 ? -1.0 : (e - 1.0) / (e + 1.0)); }
MC_DEV_NOINLINE double apparent_heat_capacity(double T, double cp_base)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    const double Tc = T - T0C, Lf = LATENT_FUSION;
    switch (sm.ahc_kind) {
    case MC_AHC_TANH: {
        const double lo = (Tc - sm.ahc_T_low_C) / sm.ahc_smoothing_K, hi = (Tc - sm.ahc_T_high_C) / sm.ahc_smoothing_K;
        return cp_base + (Lf / sm.ahc_window_K) * (0.5 * (mc_tanh(lo) - mc_tanh(hi)));
    }
    case MC_AHC_GAUSSIAN: {
        const double x = (Tc - sm.ahc_T_melt_C) / sm.ahc_width_K;
        return cp_base + (Lf / (sm.ahc_width_K * sqrt(2.0 * PI_D))) * mcm::exp(-0.5 * x * x);
    }
    case MC_AHC_WESTERMANN: {
        const double x = (Tc - sm.ahc_T_melt_C) / sm.ahc_width_K, th = mc_tanh(x);
        return cp_base + (Lf / (2.0 * sm.ahc_width_K)) * (1.0 - th * th);
    }
    default:
        if (Tc > sm.ahc_T_low_C && Tc <= sm.ahc_T_high_C) return cp_base + Lf / sm.ahc_window_K;
        return cp_base;
    }
}
MC_DEV SnowState initial_snow_state(double depth, double density_or_neg)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    const double dflt = sm.density_function[0] > 0.0 ? sm.density_function[1] : sm.snow_density;
    const double d = density_or_neg > 0.0 ? density_or_neg : dflt;
    SnowState st; st.depth = depth; st.age = 0.0; st.days_since = 0.3; st.density = d; st.prev_density = d;
    st.cum_melt = 0.0; st.extra = 0.0; st.active = 0; st.sum_phase = 0.0;
    return st;
}
MC_DEV double evolve_snow_density(const SnowState& st)
{
    const double* df = MC_CONST_MODEL.snow.density_function;
    if (!(df[0] > 0.0)) return st.density;
    if (df[2] > 0.0) return (df[0] - df[1]) * (1.0 - mcm::exp(-df[2] * st.depth - df[3] * st.age)) + df[1];
    return jmin(0.9167, df[0] * st.age + df[1]);
}
MC_DEV double snow_albedo(double days_since) { return (-9.874 * mcm::log(jmax(days_since, 0.3)) + 78.3434) / 100.0; }
MC_DEV double snow_specific_heat(double dens, double P)
{
    const DevModel& m = MC_CONST_MODEL;
    const double e = goff_gratch(273.15) * 1.0;
    const double mr = (m.eps * m.dep * e) / (P - m.dep * e);
    return 2100.0 * dens + (1005.0 + 1820.0 * (mr / 1.0 + mr)) * (1.0 - dens);
}
MC_DEV_NOINLINE void activate_snow_nodes(SnowState& st, SnowScratch& sc, double& sd)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    const double cursnow = sd;
    int active = st.active;
    if (cursnow > 300.0) active = 0;
    if (cursnow < sm.min_snow_depth_cm) {
        MC_ROLLED for (int i = 0; i < NSN; i++) sc.node_depths[i] = 0.0;
        sd = 0.0; st.active = 0; st.depth = 0.0;
        return;
    }
    int maxsnode = 0;
    MC_ROLLED for (int i = 1; i <= NSN; i++) if (cursnow > sm.node_thresholds_cm[i - 1]) maxsnode = i;
    maxsnode = maxsnode < NSN - 1 ? maxsnode : NSN - 1;
    MC_ROLLED for (int i = 1; i <= maxsnode; i++) sc.node_depths[i + (NSN - maxsnode) - 1] = sm.node_thresholds_cm[i - 1];
    if (maxsnode < active) { MC_ROLLED for (int i = 1; i <= NSN - maxsnode; i++) sc.node_depths[i - 1] = 0.0; }
    st.depth = cursnow; st.active = maxsnode;
}
MC_DEV void compute_effective_depths(const SnowState& st, const SnowScratch& sc, double* out)
{
    const DevModel& m = MC_CONST_MODEL;
    const double off = st.depth >= m.snow.min_snow_depth_cm ? st.depth : 0.0;
    MC_ROLLED for (int i = 0; i < NSN; i++) out[i] = sc.node_depths[i] / 100.0;
    MC_ROLLED for (int i = 0; i < m.N; i++) out[NSN + i] = m.depths[i] + off / 100.0;
}
MC_DEV void snow_layer_geometry(const SnowState& st, const SnowScratch& sc, int layer_index, int& node_index, double& thickness)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    const int maxsnode = st.active;
    int ni = layer_index + NSN - 1 - maxsnode; ni = ni > NSN ? NSN : ni; ni = ni < 1 ? 1 : ni;
    if (layer_index != maxsnode + 1) {
        int upper = layer_index + NSN - maxsnode; upper = upper < 1 ? 1 : upper;
        thickness = sc.node_depths[upper - 1] - sc.node_depths[ni - 1];
    } else {
        thickness = jmax(0.0, st.depth - sm.node_thresholds_cm[(maxsnode > 1 ? maxsnode : 1) - 1]);
    }
    node_index = ni - 1;
}
MC_DEV_NOINLINE double snow_thermal_melt(const SnowState& st, SnowScratch& sc, const double* Ts, const double* Ts_before, double P, double T_soil, double T_soil_before)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    MC_ROLLED for (int i = 0; i < NSN; i++) {
        const double below = i < NSN - 1 ? Ts[i + 1] : T_soil, below_past = i < NSN - 1 ? Ts_before[i + 1] : T_soil_before;
        sc.meanT[i] = (Ts[i] + below) / 2.0 - T0C;
        sc.meanTpast[i] = (Ts_before[i] + below_past) / 2.0 - T0C;
    }
    const double cp = snow_specific_heat(st.density, P);
    double total = 0.0;
    MC_ROLLED for (int j = 1; j <= st.active + 1; j++) {
        int ni; double thick; snow_layer_geometry(st, sc, j, ni, thick);
        const double mass = thick * st.density * 10.0;
        if (sc.meanT[ni] > sm.melt_threshold_C) {
            const double dT = sc.meanT[ni] - sm.melt_threshold_C;
            const double sensible = dT * cp * mass, latent = LATENT_FUSION * mass;
            total += (sensible + latent) / LATENT_FUSION;
        }
    }
    return total / 10.0;
}
MC_DEV_NOINLINE void snow_phase_transition(SnowState& st, SnowScratch& sc, double* Ts, double P, double& T_soil_surface)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    if (st.depth >= 200.0) return;
    if (st.depth < sm.min_snow_depth_cm) return;
    const double cp = snow_specific_heat(st.density, P);
    bool clamp_soil = false; bool clamp[NSN]; double phase_heat[NSN], layer_mass[NSN];
    MC_ROLLED for (int i = 0; i < NSN; i++) { clamp[i] = false; phase_heat[i] = 0.0; layer_mass[i] = 0.0; }
    MC_ROLLED for (int j = 1; j <= st.active + 1; j++) {
        if (j > NSN) continue;
        int ni; double thick; snow_layer_geometry(st, sc, j, ni, thick);
        if (!(sc.node_depths[ni] > 0.0)) continue;
        const double mass = thick * st.density * 10.0;
        if (sc.meanTpast[ni] >= 0.0 && sc.meanT[ni] <= 0.0) {
            layer_mass[ni] = jmax(0.0, mass);
            phase_heat[ni] = (sc.meanTpast[ni] - sc.meanT[ni]) * cp * layer_mass[ni];
            clamp[ni] = true;
            if (ni == NSN - 1) clamp_soil = true; else clamp[ni + 1] = true;
        }
    }
    double sp = 0.0, total_mass = 0.0;
    MC_ROLLED for (int i = 0; i < NSN; i++) sp += phase_heat[i];
    double new_sum = st.sum_phase + sp;
    MC_ROLLED for (int i = 0; i < NSN; i++) total_mass += layer_mass[i];
    if (total_mass > 0.0 && new_sum > LATENT_FUSION * total_mass) {
        bool over[NSN]; MC_ROLLED for (int i = 0; i < NSN; i++) over[i] = false;
        MC_ROLLED for (int i = 0; i < NSN; i++) if (clamp[i] && phase_heat[i] > 0.0) { over[i] = true; phase_heat[i] = 0.0; if (i < NSN - 1) over[i + 1] = true; }
        MC_ROLLED for (int i = 0; i < NSN; i++) if (over[i]) Ts[i] = T0C - 0.5;
        if (clamp_soil) T_soil_surface = T0C - 0.5;
        new_sum = 0.0;
    } else {
        MC_ROLLED for (int i = 0; i < NSN; i++) if (clamp[i]) Ts[i] = T0C;
        if (clamp_soil) T_soil_surface = T0C;
    }
    st.sum_phase = new_sum;
}
struct SnowUpdateOut { double snowfall, thermal_melt, rain_melt; };
MC_DEV_NOINLINE SnowUpdateOut update_snow(SnowState& st, SnowScratch& sc, const double* Ts, const double* Ts_before, double T_air, double rainfall, double P,
                                          int64_t step, double T_soil, double T_soil_before, double Q_evap, bool hourly, double shade, int day_of_year)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    const bool first = step == 1, midnight = ((step - 1) % 24) == 0;
    const double prev_dens = st.density;
    const double new_dens = evolve_snow_density(st);
    const double ratio = new_dens / prev_dens;
    if (step > 1 && ratio != 0.0 && isfinite(ratio)) sc.sd_prev /= ratio;
    double snowfall = 0.0, days_since = st.days_since;
    if (T_air - T0C <= sm.snow_temperature_threshold_C && rainfall > 0.0) {
        if (midnight || hourly) {
            snowfall = rainfall * sm.rain_multiplier * sm.undercatch / new_dens / 10.0;
            if (shade > 0.0) snowfall *= (1.0 - sm.canopy_interception);
            days_since = 0.3;
        } else days_since += 1.0 / 25.0;
    } else days_since += 1.0 / 25.0;
    days_since = jmax(days_since, 0.3);
    double sublimation = 0.0;
    if (T_soil - T0C > 0.0 && Q_evap > 0.0) {
        const double gw = Q_evap / latent_heat(T_soil);
        sublimation = gw * 3600.0 / new_dens / 10.0;
    }
    const double previous_depth = step > 1 ? sc.sd_prev : st.depth;
    double rain_melt = 0.0;
    if (T_air - T0C >= sm.snow_temperature_threshold_C && rainfall > 0.0 && previous_depth >= sm.min_snow_depth_cm) {
        if (midnight || hourly) {
            const double water_depth = rainfall / WATER_RHO * 100.0;
            rain_melt = sm.rain_melt_factor * water_depth * (T_air - T0C) * 24.0 / new_dens;
            if (hourly) rain_melt /= 24.0;
            rain_melt = jmax(0.0, rain_melt);
        }
    }
    SnowState for_melt = st; for_melt.density = new_dens;
    double thermal_melt;
    if (first || previous_depth < sm.min_snow_depth_cm) {
        snow_thermal_melt(for_melt, sc, Ts, Ts_before, P, T_soil, T_soil_before);
        thermal_melt = 0.0;
    } else {
        const double raw = snow_thermal_melt(for_melt, sc, Ts, Ts_before, P, T_soil, T_soil_before);
        thermal_melt = day_of_year > 1 ? raw * sm.snow_melt_factor : raw;
    }
    double cum_melt = st.cum_melt + thermal_melt;
    if (step % 24 == 0) cum_melt = 0.0;
    const double carried = st.extra / new_dens / 10.0;
    const double net_acc = jmax(0.0, snowfall - sublimation) + carried;
    double extra = 0.0;
    double new_depth = jmax(0.0, previous_depth + net_acc - rain_melt - thermal_melt);
    if (new_depth < sm.min_snow_depth_cm && (step <= 1 || sc.sd_prev < 1e-8)) { extra = new_depth * new_dens * 10.0; new_depth = 0.0; }
    sc.sd_cur = new_depth;
    double age;
    if (first) age = 0.0;
    else if (sc.sd_cur > 0.0 && sc.sd_prev > 0.0) age = st.age + 1.0 / 25.0;
    else age = 0.0;
    st.depth = new_depth; st.age = age; st.days_since = days_since; st.density = new_dens; st.prev_density = new_dens;
    st.cum_melt = cum_melt; st.extra = extra;
    SnowUpdateOut o; o.snowfall = snowfall; o.thermal_melt = thermal_melt; o.rain_melt = rain_melt;
    return o;
}
MC_DEV_NOINLINE void adjust_snow_near_nodes(SnowState& st, SnowScratch& sc, int64_t step)
{
    const mc_snow_desc& sm = MC_CONST_MODEL.snow;
    double snowout = sc.sd_cur;
    double threshold = sm.node_thresholds_cm[(st.active > 1 ? st.active : 1) - 1];
    double extra = st.extra;
    if (step > 1 && snowout > 0.0) {
        if ((snowout - threshold < 0.5) && (snowout <= sc.sd_prev)) {
            extra += (snowout - threshold) * st.density * 10.0;
            sc.sd_cur = jmax(0.0, threshold - 0.1);
        }
    }
    st.extra = extra;
    activate_snow_nodes(st, sc, sc.sd_cur);
    snowout = sc.sd_cur;
    if (snowout > 0.0) {
        threshold = sm.node_thresholds_cm[(st.active > 1 ? st.active : 1) - 1];
        if (snowout < sm.min_snow_depth_cm + 0.5 || snowout < threshold + 0.5) {
            extra = 0.5 * st.density * 10.0;
            sc.sd_cur = threshold + 0.5;
        }
    }
    st.extra = extra; st.depth = sc.sd_cur;
}
/* surface_convection_evaporation with its default zenith of 90 degrees (neutral branch), as the snow block calls it (simulation.jl:660-671) */
MC_DEV_NOINLINE double neutral_surface_evaporation(double T_surf, double T_air, double wind, double rh, double P, double z0, double lz, double wetness)
{
    const DevModel& m = MC_CONST_MODEL;
    const double dT = T_air - T_surf;
    const double rho_cp = calc_rho_cp((T_surf + T_air) / 2.0);
    const double ustar = jmax(m.kappa * wind / lz, 1.0e-6);
    const double Qc = convective_flux(rho_cp, dT, ustar, 0.64 / lz, sublayer_stanton(z0, ustar));
    const double hc = dT == 0.0 ? 0.5 : jmax(fabs(Qc / dT), 0.5);
    const WetAir wa = wet_air(T_air, rh, P);
    const double hd = (hc / (wa.specific_heat * wa.density)) * m.schmidt_factor;
    const double Tsc = T_surf < T_LO ? T_LO : T_surf;
    return (wetness * hd * (vapour_density_sat(Tsc) - wa.vapour_density)) * latent_heat(Tsc);
}

/* soil_energy_balance in its general form (snow nodes on top, first active node j, co-located nodes), heat_transport_1d.jl:137-302.
 * hs.cap holds rho*c of the soil nodes (volumetric); depths change hourly with the snowpack. */
template <int NA>
MC_DEV_NOINLINE void soil_rhs_snow(const HourSoil<NA>& hs, const HourPar& hp, const double* __restrict__ u, double t, double* __restrict__ du, Counters& cn)
{
    const DevModel& m = MC_CONST_MODEL; const SnowCtx& sn = *hp.sn; const mc_snow_desc& sm = m.snow;
    constexpr int ns = NSN; const int N = m.N; const int M = N + ns;
    cn.n_rhs++;
    const double xk = t / 60.0;
    int ik = (int)floor(xk);
    ik = ik < 0 ? 0 : (ik > 23 ? 23 : ik);
    const double dk = xk - (double)ik;
#if MC_LERP_LOOP
    double fv[7];
    MC_ROLLED for (int k = 0; k < 7; k++) fv[k] = knot_lerp(hp.ftab[k], hp.S, ik, dk);
    const double P = fv[0], T_air = fv[1];
    const double wind = jmax(fv[2], 0.1);
    const double zen = jmin(90.0, nearbyint(fv[3] * 1000.0) / 1000.0);
    const double solar = jmax(0.0, fv[4]);
    const double cloud = jclamp(fv[5], 0.0, 1.0);
    const double rh = jclamp(fv[6], 0.0, 1.0);
#else
    const double P = knot_lerp(hp.fP, hp.S, ik, dk);
    const double T_air = knot_lerp(hp.fT, hp.S, ik, dk);
    const double wind = jmax(knot_lerp(hp.fwind, hp.S, ik, dk), 0.1);
    const double zen = jmin(90.0, nearbyint(knot_lerp(hp.fzen, hp.S, ik, dk) * 1000.0) / 1000.0);
    const double solar = jmax(0.0, knot_lerp(hp.fsolar, hp.S, ik, dk));
    const double cloud = jclamp(knot_lerp(hp.fcloud, hp.S, ik, dk), 0.0, 1.0);
    const double rh = jclamp(knot_lerp(hp.frh, hp.S, ik, dk), 0.0, 1.0);
#endif
    double T[NA + NSN], lam[NA + NSN], vhc[NA + NSN], C[NA + NSN], G[NA + NSN];
    MC_ROLLED for (int i = 0; i < M; i++) T[i] = jclamp(u[i], T_LO, m.T_hi);
    const double p0p = m.P0 / P, pp0 = P / m.P0;
    MC_ROLLED for (int i = 0; i < N; i++) {
        lam[ns + i] = soil_conductivity(T[ns + i], P, p0p, pp0, hs.theta[i], hs.phim[i], hs.phig[i], hs.fw[i], hs.lam_m[i]);
        vhc[ns + i] = hs.cap[i];
    }
    {   /* write_snow_properties!, snow_model.jl:171-208 */
        const SnowState& st = *sn.rhs_state;
        const double snowdens = evolve_snow_density(st);
        if (st.depth >= sm.min_snow_depth_cm) {
            const double cps = snow_specific_heat(snowdens, P);
            double k;
            if (sm.snow_conductivity > 0.0) k = sm.snow_conductivity;
            else { const double r = snowdens * 1000.0; k = 0.00395 + 0.00084 * r - 1.7756e-6 * (r * r) + 3.80635e-9 * (r * r * r); }
            MC_ROLLED for (int i = 0; i < ns; i++) {
                lam[i] = k;
                const double cpi = apparent_heat_capacity(T[i], cps);
                const int nx = i + 1 < ns - 1 ? i + 1 : ns - 1;
                const bool active = sn.sc.node_depths[i] > 0.0 || (i < ns - 1 && sn.sc.node_depths[nx] > 0.0);
                vhc[i] = (active ? snowdens * 1000.0 : 0.0) * cpi;
            }
        } else {
            /* snow nodes take the first soil node's conductivity and specific heat with zero density: rho*c = 0 */
            MC_ROLLED for (int i = 0; i < ns; i++) { lam[i] = lam[ns]; vhc[i] = 0.0; }
        }
    }
    const double* ld = sn.depths;                     /* ld[M] = 0 sentinel */
    MC_ROLLED for (int i = 0; i < M; i++) {
        if (i == 0) {
            const double d2 = ld[1];
            C[0] = vhc[0] * d2 / 2.0;
            G[0] = d2 == 0.0 ? 0.0 : lam[0] / d2;
        } else {
            const double dd = ld[i + 1] - ld[i];
            if (i < ns && ld[i] == 0.0 && ld[i + 1] != 0.0) C[i] = vhc[i] * ld[i + 1];
            else C[i] = vhc[i] * (ld[i + 1] - ld[i - 1]) / 2.0;
            const double pdd = ld[i] - ld[i - 1];
            G[i] = dd == 0.0 ? (pdd == 0.0 ? 0.0 : lam[i] / pdd) : lam[i] / dd;
        }
    }
    int j = 0; while (j < M - 1 && C[j] == 0.0) j++;
    double Q_solar = (1.0 - hp.albedo) * solar * (1.0 - hp.shade);
    if (hp.slope > 0.0 && zen < 90.0) {
        const double szen = jmin(90.0, knot_lerp(hp.fszen, hp.S, ik, dk));
        const double fd = jclamp(knot_lerp(hp.fdiff, hp.S, ik, dk), 0.0, 1.0);
        const double cz = mcm::cos(zen * PI_D / 180.0), csz = mcm::cos(szen * PI_D / 180.0);
        Q_solar = jmax(0.0, (((1.0 - fd) * Q_solar) / cz) * csz) + (fd * Q_solar) * hp.svf;
    }
    const WetAir wa = wet_air(T_air, rh, P);
    const double atm = atmospheric_radiation(wa.vapour_pressure, T_air);
    const double cloud_rad = m.sigma * hp.eps_cloud * mcm::pow4(T_air - 2.0);
    const double hill_rad = m.sigma * hp.eps_cloud * mcm::pow4(T_air);
    const double lw_sky = (atm * (1.0 - cloud) + cloud_rad * cloud) * (1.0 - hp.shade);
    const double incoming = (lw_sky + hp.shade * hill_rad) * hp.svf + hill_rad * (1.0 - hp.svf);
    const double Q_ir = incoming - ((1.0 - hp.shade) * m.sigma * hp.eps_surf) * mcm::pow4(T[j]) - hp.shade * hill_rad;
    const double cond_ref = (t == 0.0 && j > 0) ? T[0] : T[j];
    const double Q_cond = j < M - 1 ? G[j] * (T[j + 1] - cond_ref) : 0.0;
    double Ts = T[j];
    if (fabs(Ts - T0C) > 100.0) Ts = T_air + 1.0;
    if (Ts == T_air) Ts = Ts + 0.1;
    const double dT = T_air - Ts;
    const double rho_cp = calc_rho_cp((Ts + T_air) / 2.0);
    double Qc;
    if (T_air >= Ts || zen >= 90.0) {
        const double ustar = jmax(m.kappa * wind / hp.lz, 1.0e-6);
        Qc = convective_flux(rho_cp, dT, ustar, 0.64 / hp.lz, sublayer_stanton(hp.z0, ustar));
    } else {
        const Obukhov o = calc_obukhov(wind, hp.z0, m.zref, hp.lz, dT, rho_cp, hp.L_prev >= 0.0 ? -0.3 : hp.L_prev);
        cn.n_obukhov += o.iters;
        Qc = o.Q;
    }
    const double hc = dT == 0.0 ? 0.5 : jmax(fabs(Qc / dT), 0.5);
    const double hd = (hc / (wa.specific_heat * wa.density)) * m.schmidt_factor;
    const double Tsc = Ts < T_LO ? T_LO : Ts;
    const double Qe = (hp.soil_wetness * hd * (vapour_density_sat(Tsc) - wa.vapour_density)) * latent_heat(Tsc);
    const double surf = 60.0 * (Q_solar + Q_ir + Q_cond + Qc + sn.Q_freeze - Qe) / C[j];
    MC_ROLLED for (int i = 0; i < M; i++) {
        if (i <= j) du[i] = surf;
        else if (i == M - 1) du[i] = 0.0;
        else du[i] = C[i] == 0.0 ? 0.0 : 60.0 * (G[i - 1] * (T[i - 1] - T[i]) + G[i] * (T[i + 1] - T[i])) / C[i];
    }
}

template <int NA, int NS>
MC_DEV void rhs_dispatch(const HourSoil<NA>& hs, const HourPar& hp, const double* u, double t, double* du, Counters& cn)
{
    if (NS == 0) soil_rhs<NA>(hs, hp, u, t, du, cn);
    else soil_rhs_snow<NA>(hs, hp, u, t, du, cn);
}
