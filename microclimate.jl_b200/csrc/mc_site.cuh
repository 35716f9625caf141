/*
 * mc_site.cuh - solve_soil! / solve_air! for one site (per-site kernel body and cooperative kernel body) and the host-side model flattening.
 * Part of mc_device.cuh (included from there, inside namespace mcd, in this order); not a stand-alone header.
 */

/* theta-only soil property terms for the coming ODE hour (theta is constant while the ODE advances) */
template <int NA, int NS>
MC_DEV_NOINLINE void hoist_soil(const SiteRef& r, const double* theta, HourSoil<NA>& hs)
{
    const DevModel& m = MC_CONST_MODEL; const KArgs& a = *r.a;
    MC_ROLLED for (int i = 0; i < m.N; i++) {
        const double bd = r.node(a.bd, i), md = r.node(a.md, i);
        const double th = theta[i];
        const double cp = bd / md * r.node(a.c_m, i) + th * WATER_CP;
        const double rho = th * WATER_RHO + bd;
        hs.cap[i] = NS == 0 ? (rho * cp) * m.dzc[i] / 2.0 : rho * cp;      /* with snow the node geometry changes hourly: keep rho*c */
        hs.theta[i] = th; hs.phim[i] = bd / md; hs.phig[i] = jmax(0.0, 1.0 - th - bd / md);
        hs.fw[i] = 1.0 / (1.0 + mcm::pow(th / m.theta0, -m.recirc_pow));
    }
}
MC_DEV_NOINLINE void put_potential_row(const SiteRef& r, int row, const double* theta)
{
    const KArgs& a = *r.a;
    if (!a.o_psi) return;
    MC_ROLLED for (int i = 0; i < MC_CONST_MODEL.N; i++)
        r.put_node(a.o_psi, row, i, r.node(a.psi_e, i) * mcm::pow((1.0 - r.node(a.bd, i) / r.node(a.md, i)) / theta[i], r.node(a.bb, i)));
}
/* initialise_soil_humidity!, dynamic.jl:32-35: always row 1 of the buffer */
MC_DEV_NOINLINE void init_humidity(const SiteRef& r, const double* theta, const double* Tn)
{
    const DevModel& m = MC_CONST_MODEL; const KArgs& a = *r.a;
    if (!m.dynamic || !a.o_h) return;
    MC_ROLLED for (int i = 0; i < m.N; i++) {
        const double psi = r.node(a.psi_e, i) * mcm::pow((1.0 - r.node(a.bd, i) / r.node(a.md, i)) / theta[i], r.node(a.bb, i));
        r.put_node(a.o_h, 0, i, jclamp(mcm::exp(MW_INFIL * psi / (m.R * Tn[i])), 0.0, 1.0));
    }
}
MC_DEV_NOINLINE void write_props(const SiteRef& r, int row, const double* Tn, const double* theta, double P)
{
    const KArgs& a = *r.a;
    if (!a.o_lam && !a.o_cp && !a.o_rho) return;
    MC_ROLLED for (int i = 0; i < MC_CONST_MODEL.N; i++) {
        double lam, cp, rho;
        soil_properties_full(Tn[i], P, theta[i], r.node(a.bd, i), r.node(a.md, i), r.node(a.lam_m, i), r.node(a.c_m, i), lam, cp, rho);
        r.put_node(a.o_lam, row, i, lam); r.put_node(a.o_cp, row, i, cp); r.put_node(a.o_rho, row, i, rho);
    }
}

MC_DEV bool is_converged(int iter, int niter, const double* T0, const double* T0_prev, int N)
{
    const DevModel& m = MC_CONST_MODEL;
    if (m.conv_kind == MC_CONV_FIXED) return iter >= niter;
    if (iter >= niter) return true;
    if (iter <= 1) return niter <= 1;
    double mx = 0.0; MC_ROLLED for (int i = 0; i < N; i++) mx = fmax(mx, fabs(T0[i] - T0_prev[i]));
    return mx < m.conv_tol;
}

/* ------------------------------------------------------------------------------------------
 * solve_soil! for days [day0, day0+ndays) of site s (simulation.jl:355-819).  NS = 0 (NoSnow) or 8 (SnowModel):
 * with snow the ODE state is [T_snow(8); T_soil(N)] over hourly-changing effective depths.
 * ------------------------------------------------------------------------------------------ */
template <int NA, int NS>
MC_DEV void solve_site(const KArgs& a, int64_t s, Counters& cn)
{
    const DevModel& m = MC_CONST_MODEL;
    const int N = m.N;
    constexpr int ns = NS;
    constexpr bool snow = NS > 0;
    const int nrows = a.ndays * 24;
    const bool dynamic = m.dynamic != 0, nonconsec = m.nonconsec != 0;
    SiteRef r; r.a = &a; r.s = s;
    const int64_t S = a.S;

    HourSoil<NA> hs; HourPar hp; Integ<NA + NS> I; SnowCtx sn;
    double T0[NA], theta[NA], sum_phase[NA], sum_phase_day_start[NA], T0_prev_start[NA], T0_day_start[NA], T0_before[NA], T0_output[NA];
    double psi_out[NA], h_out[NA];
    double Ts[NSN], Ts_before[NSN], Ts_day_start[NSN];
    MC_ROLLED for (int i = 0; i < N; i++) { T0[i] = a.T[(int64_t)i * S + s]; theta[i] = a.theta[(int64_t)i * S + s]; sum_phase[i] = a.sum_phase[(int64_t)i * S + s]; hs.lam_m[i] = r.node(a.lam_m, i); }
    double pool = a.pool[s], wet_dyn = a.wetness[s], L_ode = a.L_ode[s], L_prof = a.L_prof[s], tdeep_stale = a.tdeep_prev[s];
    int flags = a.flags[s];
    const double site_albedo = MC_LDG(a.albedo + s);
    hp.z0 = MC_LDG(a.z0 + s); hp.slope = MC_LDG(a.slope + s); hp.svf = MC_LDG(a.svf + s); hp.lz = mcm::log(m.zref / hp.z0); hp.S = S;
    hp.sn = snow ? &sn : nullptr;
    I.flags = 0; I.iter = 0; I.accept = false; I.u_modified = false; I.t = 0; I.dt = 0; I.dtpropose = 0; I.qold = 1e-4; I.q11 = 1; I.tstop = 1440.0;
    MC_ROLLED for (int i = 0; i < N + ns; i++) { I.k1[i] = a.fsal[(int64_t)i * S + s]; I.k7[i] = I.k1[i]; }
    double sd_last = 0.0;
    SnowState proto_state;
    if (snow) {
        MC_ROLLED for (int i = 0; i < NSN; i++) { Ts[i] = a.snow_T[(int64_t)i * S + s]; sn.sc.node_depths[i] = a.snow_nd[(int64_t)i * S + s]; sn.sc.meanT[i] = 0.0; sn.sc.meanTpast[i] = 0.0; }
        const double* q = a.snow_state + s;
        sn.st.depth = q[0]; sn.st.age = q[S]; sn.st.days_since = q[2 * S]; sn.st.density = q[3 * S]; sn.st.prev_density = q[4 * S];
        sn.st.cum_melt = q[5 * S]; sn.st.extra = q[6 * S]; sn.st.active = (int)q[7 * S]; sn.st.sum_phase = q[8 * S]; sd_last = q[9 * S];
        sn.sc.sd_prev = 0.0; sn.sc.sd_cur = 0.0;
        sn.Q_freeze = a.q_freeze[s]; sn.rhs_state = &sn.st;
        MC_ROLLED for (int i = 0; i <= N + ns; i++) sn.depths[i] = 0.0;
    }

    auto set_day_forcing = [&](int jl) {
        const int64_t off = (int64_t)jl * 24 * S + s;
        hp.fP = a.f_P + off; hp.fT = a.f_T + off; hp.fwind = a.f_wind + off; hp.frh = a.f_rh + off; hp.fcloud = a.f_cloud + off;
        hp.fsolar = a.f_solar + off; hp.fdiff = a.f_diffuse + off; hp.fzen = a.f_zen + off; hp.fszen = a.f_szen + off;
        hp.ftab[0] = hp.fP; hp.ftab[1] = hp.fT; hp.ftab[2] = hp.fwind; hp.ftab[3] = hp.fzen; hp.ftab[4] = hp.fsolar; hp.ftab[5] = hp.fcloud; hp.ftab[6] = hp.frh;
    };
    auto tins_mean = [&](int jl) { double ssum = 0.0; MC_ROLLED for (int h = 0; h < 24; h++) ssum += r.hourly(a.f_T, jl * 24 + h); ssum += r.hourly(a.f_T, jl * 24 + 23); return ssum / 25.0; };
    double sky_T = 0.0;
    bool is_last = false;
    /* apply_snow_overrides (simulation.jl:46-81) + per-hour parameter update */
    auto set_hour_params = [&](const EnvInstant& e, int64_t gstep, double depth_prev_hour) {
        hp.shade = e.shade; hp.eps_surf = e.eps_surf; hp.eps_cloud = e.eps_cloud;
        const bool snow_active = snow && is_last && ((gstep > 1 ? depth_prev_hour : sn.st.depth) > 0.0);
        if (snow_active) { hp.albedo = snow_albedo(sn.st.days_since); hp.soil_wetness = 1.0; }
        else { hp.albedo = e.albedo; hp.soil_wetness = dynamic ? wet_dyn : e.wet_daily; }
        sky_T = sky_temperature(hp.svf, e.P, e.rh, e.T_ref, e.eps_cloud, e.cloud, e.shade, a.has_lw != 0, e.lw);
    };
    auto combine = [&](double* u) { MC_ROLLED for (int i = 0; i < ns; i++) u[i] = Ts[i]; MC_ROLLED for (int i = 0; i < N; i++) u[ns + i] = T0[i]; };
    auto sync_inactive = [&]() { const double sy = Ts[0]; MC_ROLLED for (int k = 0; k < ns; k++) if (!(sn.sc.node_depths[k] > 0.0)) Ts[k] = sy; };

    EnvInstant env;
    if (a.day0 == 0) {
        /* init(): SciMLBase.init evaluates the RHS once at prototype parameters to seed fsalfirst (simulation.jl:196-238,853-860) */
        set_day_forcing(0);
        env = get_instant(r, 0, 0, site_albedo);
        MC_ROLLED for (int i = 0; i < ns; i++) I.tmp[i] = m.snow.initial_snow_temperature_K;
        MC_ROLLED for (int i = 0; i < N; i++) I.tmp[ns + i] = a.has_init_T ? T0[i] : 280.0;
        hoist_soil<NA, NS>(r, theta, hs);
        hp.shade = env.shade; hp.eps_surf = env.eps_surf; hp.eps_cloud = env.eps_cloud; hp.soil_wetness = 0.0; hp.albedo = site_albedo; hp.L_prev = L_ode;
        if (snow) {           /* depths_proto = combined_depths(initial snow state) but p.snow_state = initial_snow_state(model) */
            proto_state = initial_snow_state(0.0, -1.0);
            compute_effective_depths(sn.st, sn.sc, sn.depths);
            sn.rhs_state = &proto_state;
            const double qf = sn.Q_freeze; sn.Q_freeze = 0.0;
            Counters scratch = {};
            rhs_dispatch<NA, NS>(hs, hp, I.tmp, 0.0, I.k1, scratch);
            sn.rhs_state = &sn.st; sn.Q_freeze = qf;
        } else {
            Counters scratch = {};
            rhs_dispatch<NA, NS>(hs, hp, I.tmp, 0.0, I.k1, scratch);
        }
        /* solve_soil! pre-loop (simulation.jl:395-452) */
        if (!a.has_init_T) { const double t_init = tins_mean(0); MC_ROLLED for (int i = 0; i < N; i++) T0[i] = t_init; }
        MC_ROLLED for (int i = 0; i < N; i++) sum_phase[i] = 0.0;
        write_props(r, 0, T0, theta, 101325.0);
        put_potential_row(r, 0, theta);
        MC_ROLLED for (int i = 0; i < N; i++) { r.put_node(a.o_T, 0, i, T0[i]); r.put_node(a.o_theta, 0, i, theta[i]); }
        init_humidity(r, theta, T0);
        tdeep_stale = env.T_deep;
        pool = 0.0;
        if (snow) sn.Q_freeze = 0.0;
    }

    for (int jl = 0; jl < a.ndays; jl++) {
        const int64_t j = a.day0 + jl;
        const int doy = a.doy ? (int)MC_LDG(a.doy + jl) : (int)(j + 1);
        set_day_forcing(jl);
        const int niter = nonconsec ? m.iters_per_day : ((m.spinup && j == 0) ? m.conv_iters : 1);
        int noon_i = 0; double best = r.hourly(a.f_zen, jl * 24);
        MC_ROLLED for (int i = 1; i < 24; i++) { const double zz = r.hourly(a.f_zen, jl * 24 + i); if (zz < best) { best = zz; noon_i = i; } }
        const int midnight_i = (noon_i + 12) % 24;
        if (nonconsec) {                                                      /* per-day resets (simulation.jl:480-516); the snowpack is NOT reset */
            MC_ROLLED for (int i = 0; i < N; i++) sum_phase[i] = 0.0;
            if (snow) sn.st.sum_phase = 0.0;
            if (!dynamic && a.has_presc) { MC_ROLLED for (int i = 0; i < N; i++) theta[i] = MC_LDG(a.presc + ((int64_t)jl * N + i) * S + s); }
            else { MC_ROLLED for (int i = 0; i < N; i++) theta[i] = r.node(a.init_theta, i); }
            const double t_reset = tins_mean(jl);
            if (!a.has_init_T) { MC_ROLLED for (int i = 0; i < N; i++) T0[i] = t_reset; }
            else { MC_ROLLED for (int i = 0; i < N; i++) T0[i] = r.node(a.init_T, i); }
            MC_ROLLED for (int i = 0; i < ns; i++) Ts[i] = t_reset;
        }
        T0[N - 1] = tdeep_stale;
        int iter = 0; is_last = false;
        MC_ROLLED for (int i = 0; i < N; i++) { T0_prev_start[i] = T0[i]; T0_day_start[i] = T0[i]; sum_phase_day_start[i] = sum_phase[i]; }
        MC_ROLLED for (int i = 0; i < ns; i++) Ts_day_start[i] = Ts[i];
        while (!is_last) {                                                    /* outer fixed point over day passes (simulation.jl:530) */
            iter++; cn.n_day_passes++;
            is_last = is_converged(iter, niter, T0, T0_prev_start, N);
            MC_ROLLED for (int i = 0; i < N; i++) T0_prev_start[i] = T0[i];
            if (!nonconsec) { MC_ROLLED for (int i = 0; i < N; i++) T0[i] = T0_day_start[i]; MC_ROLLED for (int i = 0; i < ns; i++) Ts[i] = Ts_day_start[i]; }
            MC_ROLLED for (int i = 0; i < N; i++) { sum_phase[i] = sum_phase_day_start[i]; T0_output[i] = T0[i]; }
            if (snow && nonconsec) sn.st.sum_phase = 0.0;
            const int day_row = jl * 24;
            if (is_last) {                                                    /* day row 1 = state at the start of the output pass (:557-592) */
                MC_ROLLED for (int i = 0; i < N; i++) { r.put_node(a.o_T, day_row, i, T0[i]); r.put_node(a.o_theta, day_row, i, theta[i]); }
                r.put_scalar(a.o_pool, day_row, pool);
                put_potential_row(r, day_row, theta);
                init_humidity(r, theta, T0);
                const EnvInstant e0 = get_instant(r, jl, day_row, site_albedo);
                r.put_scalar(a.o_sky, day_row, sky_temperature(hp.svf, e0.P, e0.rh, e0.T_ref, e0.eps_cloud, e0.cloud, e0.shade, a.has_lw != 0, e0.lw));
                write_props(r, day_row, T0, theta, e0.P);
            }
            env = get_instant(r, jl, day_row, site_albedo);
            T0[N - 1] = env.T_deep;
            set_hour_params(env, a.day0 * 24 + day_row + 1, sd_last);
            if (snow) {
                if (!is_last) { double tmp = sn.st.depth; activate_snow_nodes(sn.st, sn.sc, tmp); }
                compute_effective_depths(sn.st, sn.sc, sn.depths);
                int fa = -1; MC_ROLLED for (int k = ns - 1; k >= 0; k--) if (sn.sc.node_depths[k] > 0.0) fa = k;
                const double sync_temp = fa < 0 ? T0[0] : Ts[fa];
                MC_ROLLED for (int k = 0; k < ns; k++) if (fa < 0 || k < fa) Ts[k] = sync_temp;       /* _sync_snow_to_active */
            }
            hp.L_prev = L_ode;
            hoist_soil<NA, NS>(r, theta, hs);
            /* reinit!(integrator, u0; t0=0, reset_dt=true, reinit_cache=false): k1 (fsalfirst) deliberately stays stale */
            combine(I.u);
            MC_ROLLED for (int i = 0; i < N + ns; i++) I.uprev[i] = I.u[i];
            I.t = 0.0; I.iter = 0; I.u_modified = false; I.accept = false; I.qold = 1e-4; I.q11 = 1.0; I.tstop = 1440.0;
            I.dt = determine_initdt<NA, NS>(I, hs, hp, cn);
            for (int i = 0; i < 24; i++) {
                const int step = jl * 24 + i;
                const int64_t gstep = a.day0 * 24 + step + 1;
                MC_ROLLED for (int k = 0; k < N; k++) T0_before[k] = T0[k];
                MC_ROLLED for (int k = 0; k < ns; k++) Ts_before[k] = Ts[k];
                if (i > 0) hoist_soil<NA, NS>(r, theta, hs);
                hp.L_prev = L_ode;
                /* step_to_hour! (simulation.jl:888-903) */
                const double target = (i + 1) * 60.0;
                I.tstop = target;
                while (I.t < target - 1e-10) {
                    tsit5_step<NA, NS>(I, hs, hp, cn);
                    bool changed = false;
                    MC_ROLLED for (int k = 0; k < N + ns; k++) { const double c = jclamp(I.u[k], T_LO, m.T_hi); if (c != I.u[k]) { I.u[k] = c; changed = true; } }
                    if (changed) { MC_ROLLED for (int k = 0; k < N + ns; k++) I.uprev[k] = I.u[k]; I.u_modified = true; }
                }
                MC_ROLLED for (int k = 0; k < ns; k++) Ts[k] = I.u[k];
                MC_ROLLED for (int k = 0; k < N; k++) T0[k] = I.u[ns + k];
                if (is_last) {
                    phase_transition(N, m.depths, T0, T0_before, sum_phase, theta);
                    MC_ROLLED for (int k = 0; k < N; k++) T0_output[k] = T0[k];
                }
                double snowfall_hour = 0.0, thermal_melt = 0.0, rain_melt_snow = 0.0;
                if (snow && is_last) {                                        /* snow post-ODE block (simulation.jl:656-708) */
                    sn.sc.sd_prev = sd_last;
                    const double Qe = neutral_surface_evaporation(T0[0], env.T_ref, env.wind, env.rh, env.P, hp.z0, hp.lz, hp.soil_wetness);
                    const SnowUpdateOut so = update_snow(sn.st, sn.sc, Ts, Ts_before, env.T_ref, env.rain_daily, env.P, gstep, T0[0], T0_before[0], Qe,
                                                         m.rain_hourly != 0, env.shade, doy);
                    snowfall_hour = so.snowfall; thermal_melt = so.thermal_melt; rain_melt_snow = so.rain_melt;
                    if (doy > 1 && m.snow.snow_melt_factor <= 1.0) sn.Q_freeze = thermal_melt * 10.0 * LATENT_FUSION / 3600.0 * (1.0 - m.snow.snow_melt_factor);
                    else sn.Q_freeze = 0.0;
                    int prev_active = sn.st.active;
                    activate_snow_nodes(sn.st, sn.sc, sn.sc.sd_cur);
                    if (sn.st.active < prev_active) sync_inactive();
                    double soil_surf = T0[0];
                    snow_phase_transition(sn.st, sn.sc, Ts, env.P, soil_surf);      /* soil-surface clamp deliberately not applied upstream */
                    MC_ROLLED for (int k = 0; k < ns; k++) if (sn.sc.node_depths[k] > 0.0 && Ts[k] > T0C) Ts[k] = T0C;     /* clamp_snow_temperatures */
                    /* freeze_new_snow selects an empty index range upstream: no-op */
                    adjust_snow_near_nodes(sn.st, sn.sc, gstep);
                    prev_active = sn.st.active;
                    activate_snow_nodes(sn.st, sn.sc, sn.sc.sd_cur);
                    if (sn.st.active < prev_active) sync_inactive();
                    sd_last = sn.sc.sd_cur;
                }
                if (i < 23) {
                    env = get_instant(r, jl, step + 1, site_albedo);
                    T0[N - 1] = env.T_deep;
                    set_hour_params(env, gstep + 1, sd_last);
                    if (snow) {
                        if (!is_last) { double tmp = sn.st.depth; activate_snow_nodes(sn.st, sn.sc, tmp); }
                        compute_effective_depths(sn.st, sn.sc, sn.depths);
                    }
                }
                combine(I.u);
                I.u_modified = true;
                {   /* init_surface_fluxes! (monin_obukhov.jl:247-264): hourly warm start of the Obukhov length */
                    const double T_air = MC_LDG(hp.fT + (int64_t)i * S);
                    const double wind = jmax(MC_LDG(hp.fwind + (int64_t)i * S), 0.1);
                    const double zen = jmin(90.0, nearbyint(MC_LDG(hp.fzen + (int64_t)i * S) * 1000.0) / 1000.0);
                    const double Tsf = T0[0];
                    if (T_air < Tsf && zen < 90.0) {
                        const Obukhov ob = calc_obukhov(wind, hp.z0, m.zref, hp.lz, T_air - Tsf, calc_rho_cp((Tsf + T_air) / 2.0), L_ode >= 0.0 ? -0.3 : L_ode);
                        cn.n_obukhov += ob.iters;
                        L_ode = ob.L;
                    }
                }
                bool have_infil = false;
                if (is_last) {
                    double rain;
                    if (m.rain_hourly) rain = r.hourly(a.f_rain, step);
                    else rain = (i == midnight_i) ? env.rain_daily : 0.0;
                    bool snow_present = false;
                    if (snow) {
                        rain = rain * m.snow.rain_multiplier;
                        snow_present = sn.sc.sd_cur > 0.0;
                        const double melted_water = thermal_melt * 10.0, rain_melt_water = rain_melt_snow * sn.st.density * 10.0;
                        if (snow_present && env.T_ref - T0C < m.snow.snow_temperature_threshold_C) pool = jclamp(pool + melted_water, 0.0, m.max_pool);
                        else pool = jclamp(pool + rain + rain_melt_water + melted_water, 0.0, m.max_pool);
                    } else {
                        pool = jclamp(pool + rain, 0.0, m.max_pool);
                    }
                    if (dynamic) {
                        wet_dyn = soil_water_hour<NA>(r, env, T0, theta, pool, wet_dyn, hp.z0, L_prof, psi_out, h_out, flags, cn, snow_present);
                        have_infil = true;
                    }
                }
                const int out_row = step + 1;
                if (is_last && out_row < nrows && !(i == 23 && nonconsec)) {
                    r.put_scalar(a.o_pool, out_row, pool);
                    MC_ROLLED for (int k = 0; k < N; k++) { r.put_node(a.o_T, out_row, k, T0_output[k]); r.put_node(a.o_theta, out_row, k, theta[k]); }
                    r.put_scalar(a.o_sky, out_row, sky_T);
                    if (have_infil) { MC_ROLLED for (int k = 0; k < N; k++) { r.put_node(a.o_psi, out_row, k, psi_out[k]); r.put_node(a.o_h, out_row, k, h_out[k]); } }
                    write_props(r, out_row, T0, theta, r.hourly(a.f_P, out_row));
                }
                if (is_last && snow) {                                        /* snow outputs are recorded at `step` (simulation.jl:803-813) */
                    r.put_scalar(a.o_snow_fall, step, snowfall_hour);
                    r.put_scalar(a.o_snow_depth, step, sn.sc.sd_cur);
                    r.put_scalar(a.o_snow_dens, step, evolve_snow_density(sn.st));
                    if (a.o_snow_T) { MC_ROLLED for (int k = 0; k < ns; k++) a.o_snow_T[((int64_t)step * ns + k) * S + s] = Ts[k]; }
                }
            }
        }
        tdeep_stale = env.T_deep;
    }
    MC_ROLLED for (int i = 0; i < N; i++) {
        a.T[(int64_t)i * S + s] = T0[i]; a.theta[(int64_t)i * S + s] = theta[i]; a.sum_phase[(int64_t)i * S + s] = sum_phase[i];
        if (!(isfinite(T0[i]) && isfinite(theta[i]))) flags |= MC_STATUS_NAN;
    }
    MC_ROLLED for (int i = 0; i < N + ns; i++) a.fsal[(int64_t)i * S + s] = I.k1[i];
    a.pool[s] = pool; a.wetness[s] = wet_dyn; a.L_ode[s] = L_ode; a.L_prof[s] = L_prof; a.tdeep_prev[s] = tdeep_stale;
    a.flags[s] = flags | I.flags;
    if (snow) {
        MC_ROLLED for (int i = 0; i < NSN; i++) { a.snow_T[(int64_t)i * S + s] = Ts[i]; a.snow_nd[(int64_t)i * S + s] = sn.sc.node_depths[i]; }
        double* q = a.snow_state + s;
        q[0] = sn.st.depth; q[S] = sn.st.age; q[2 * S] = sn.st.days_since; q[3 * S] = sn.st.density; q[4 * S] = sn.st.prev_density;
        q[5 * S] = sn.st.cum_melt; q[6 * S] = sn.st.extra; q[7 * S] = (double)sn.st.active; q[8 * S] = sn.st.sum_phase; q[9 * S] = sd_last;
        a.q_freeze[s] = sn.Q_freeze;
    }
}


/* ------------------------------------------------------------------------------------------
 * solve_soil! for one site with the block-cooperative ODE phase (NoSnow, fixed day-pass counts so that every site of the block makes
 * the same sequence of coop2_rounds calls).  Everything outside the ODE hour is solve_site<NA,0> verbatim.
 * ------------------------------------------------------------------------------------------ */
template <int NA>
MC_DEV void solve_site_coop2(const KArgs& a, int64_t s, Counters& cn, Coop2Shared& sh, MoistShared& ms, double* scr, int& rot_ode, int& rot_moist)
{
    const DevModel& m = MC_CONST_MODEL;
    const int N = m.N;
    const int nrows = a.ndays * 24;
    const bool dynamic = m.dynamic != 0, nonconsec = m.nonconsec != 0;
    SiteRef r; r.a = &a; r.s = s;
    const int64_t S = a.S;
    const int tid = MC_TID;

    HourPar hp; Own2 I;
    const bool coop_moist = MC_COOP_MOIST && !m.flux_alg && m.niter_moist >= 1 && m.moist_max_it >= 1;      /* water-potential form of infiltration_step! */
    int& mpar = rot_moist; (void)mpar; (void)coop_moist; (void)ms;
    double T0[NA], theta[NA], sum_phase[NA], sum_phase_day_start[NA], T0_day_start[NA], T0_before[NA], T0_output[NA];
    double psi_out[NA], h_out[NA];
    MC_ROLLED for (int i = 0; i < N; i++) { T0[i] = a.T[(int64_t)i * S + s]; theta[i] = a.theta[(int64_t)i * S + s]; sum_phase[i] = a.sum_phase[(int64_t)i * S + s]; MC_SCR(P_LM, i, tid) = r.node(a.lam_m, i); }
    double pool = a.pool[s], wet_dyn = a.wetness[s], L_ode = a.L_ode[s], L_prof = a.L_prof[s], tdeep_stale = a.tdeep_prev[s];
    int flags = a.flags[s];
    const double site_albedo = MC_LDG(a.albedo + s);
    hp.z0 = MC_LDG(a.z0 + s); hp.slope = MC_LDG(a.slope + s); hp.svf = MC_LDG(a.svf + s); hp.lz = mcm::log(m.zref / hp.z0); hp.S = S;
    hp.sn = nullptr;
    I.flags = 0; I.iter = 0; I.accept = false; I.u_modified = false; I.t = 0; I.dt = 0; I.dtpropose = 0; I.qold = 1e-4; I.q11 = 1; I.tstop = 1440.0;
    I.st = 0; I.guard = 0; I.need_initdt = false; I.u_stale = false; I.kswap = false; I.uswap = false; I.stage = ST_NONE; I.q0 = 0.0; I.par = rot_ode;      /* rotation of the block's list counters continues across site batches */
    MC_ROLLED for (int i = 0; i < N; i++) { const double f = a.fsal[(int64_t)i * S + s]; MC_SCR(PK_A, i, tid) = f; MC_SCR(PK_B, i, tid) = f; }

    auto set_day_forcing = [&](int jl) {
        const int64_t off = (int64_t)jl * 24 * S + s;
        hp.fP = a.f_P + off; hp.fT = a.f_T + off; hp.fwind = a.f_wind + off; hp.frh = a.f_rh + off; hp.fcloud = a.f_cloud + off;
        hp.fsolar = a.f_solar + off; hp.fdiff = a.f_diffuse + off; hp.fzen = a.f_zen + off; hp.fszen = a.f_szen + off;
    };
    auto tins_mean = [&](int jl) { double ssum = 0.0; MC_ROLLED for (int h = 0; h < 24; h++) ssum += r.hourly(a.f_T, jl * 24 + h); ssum += r.hourly(a.f_T, jl * 24 + 23); return ssum / 25.0; };
    double sky_T = 0.0;
    auto set_hour_params = [&](const EnvInstant& e) {
        hp.shade = e.shade; hp.eps_surf = e.eps_surf; hp.eps_cloud = e.eps_cloud;
        hp.albedo = e.albedo; hp.soil_wetness = dynamic ? wet_dyn : e.wet_daily;
        sky_T = sky_temperature(hp.svf, e.P, e.rh, e.T_ref, e.eps_cloud, e.cloud, e.shade, a.has_lw != 0, e.lw);
    };
    /* theta-only soil property terms of the coming ODE hour, straight into the block's scratch (hoist_soil) */
    auto hoist_publish = [&]() {
        MC_ROLLED for (int i = 0; i < N; i++) {
            const double bd = r.node(a.bd, i), md = r.node(a.md, i);
            const double th = theta[i];
            const double cp = bd / md * r.node(a.c_m, i) + th * WATER_CP;
            const double rho = th * WATER_RHO + bd;
            MC_SCR(P_CAP, i, tid) = (rho * cp) * m.dzc[i] / 2.0;
            MC_SCR(P_TH, i, tid) = th; MC_SCR(P_PM, i, tid) = bd / md; MC_SCR(P_PG, i, tid) = jmax(0.0, 1.0 - th - bd / md);
            MC_SCR(P_FW, i, tid) = 1.0 / (1.0 + mcm::pow(th / m.theta0, -m.recirc_pow));
        }
    };

    EnvInstant env;
    if (a.day0 == 0) {
        /* init(): SciMLBase.init evaluates the RHS once at prototype parameters to seed fsalfirst (simulation.jl:196-238,853-860) */
        set_day_forcing(0);
        env = get_instant(r, 0, 0, site_albedo);
        MC_ROLLED for (int i = 0; i < N; i++) MC_SCR(I.uparr(), i, tid) = a.has_init_T ? T0[i] : 280.0;
        hoist_publish();
        hp.shade = env.shade; hp.eps_surf = env.eps_surf; hp.eps_cloud = env.eps_cloud; hp.soil_wetness = 0.0; hp.albedo = site_albedo; hp.L_prev = L_ode;
        {
            Counters scratch = {};
            I.st = 10; I.stage = ST_REFRESH; I.req_t = 0.0;
            coop2_rounds<NA>(I, true, hp, sh, scr, scratch);
        }
        /* solve_soil! pre-loop (simulation.jl:395-452) */
        if (!a.has_init_T) { const double t_init = tins_mean(0); MC_ROLLED for (int i = 0; i < N; i++) T0[i] = t_init; }
        MC_ROLLED for (int i = 0; i < N; i++) sum_phase[i] = 0.0;
        write_props(r, 0, T0, theta, 101325.0);
        put_potential_row(r, 0, theta);
        MC_ROLLED for (int i = 0; i < N; i++) { r.put_node(a.o_T, 0, i, T0[i]); r.put_node(a.o_theta, 0, i, theta[i]); }
        init_humidity(r, theta, T0);
        tdeep_stale = env.T_deep;
        pool = 0.0;
    }

    for (int jl = 0; jl < a.ndays; jl++) {
        const int64_t j = a.day0 + jl;
        set_day_forcing(jl);
        const int niter = nonconsec ? m.iters_per_day : ((m.spinup && j == 0) ? m.conv_iters : 1);
        int noon_i = 0; double best = r.hourly(a.f_zen, jl * 24);
        MC_ROLLED for (int i = 1; i < 24; i++) { const double zz = r.hourly(a.f_zen, jl * 24 + i); if (zz < best) { best = zz; noon_i = i; } }
        const int midnight_i = (noon_i + 12) % 24;
        if (nonconsec) {                                                      /* per-day resets (simulation.jl:480-516) */
            MC_ROLLED for (int i = 0; i < N; i++) sum_phase[i] = 0.0;
            if (!dynamic && a.has_presc) { MC_ROLLED for (int i = 0; i < N; i++) theta[i] = MC_LDG(a.presc + ((int64_t)jl * N + i) * S + s); }
            else { MC_ROLLED for (int i = 0; i < N; i++) theta[i] = r.node(a.init_theta, i); }
            const double t_reset = tins_mean(jl);
            if (!a.has_init_T) { MC_ROLLED for (int i = 0; i < N; i++) T0[i] = t_reset; }
            else { MC_ROLLED for (int i = 0; i < N; i++) T0[i] = r.node(a.init_T, i); }
        }
        T0[N - 1] = tdeep_stale;
        MC_ROLLED for (int i = 0; i < N; i++) { T0_day_start[i] = T0[i]; sum_phase_day_start[i] = sum_phase[i]; }
        for (int iter = 1; iter <= niter; iter++) {                           /* fixed day-pass count (block-uniform) */
            cn.n_day_passes++;
            const bool is_last = iter >= niter;
            if (!nonconsec) { MC_ROLLED for (int i = 0; i < N; i++) T0[i] = T0_day_start[i]; }
            MC_ROLLED for (int i = 0; i < N; i++) { sum_phase[i] = sum_phase_day_start[i]; T0_output[i] = T0[i]; }
            const int day_row = jl * 24;
            if (is_last) {                                                    /* day row 1 = state at the start of the output pass (:557-592) */
                MC_ROLLED for (int i = 0; i < N; i++) { r.put_node(a.o_T, day_row, i, T0[i]); r.put_node(a.o_theta, day_row, i, theta[i]); }
                r.put_scalar(a.o_pool, day_row, pool);
                put_potential_row(r, day_row, theta);
                init_humidity(r, theta, T0);
                const EnvInstant e0 = get_instant(r, jl, day_row, site_albedo);
                r.put_scalar(a.o_sky, day_row, sky_temperature(hp.svf, e0.P, e0.rh, e0.T_ref, e0.eps_cloud, e0.cloud, e0.shade, a.has_lw != 0, e0.lw));
                write_props(r, day_row, T0, theta, e0.P);
            }
            env = get_instant(r, jl, day_row, site_albedo);
            T0[N - 1] = env.T_deep;
            set_hour_params(env);
            /* reinit!(integrator, u0; t0=0, reset_dt=true, reinit_cache=false): k1 (fsalfirst) deliberately stays stale */
            MC_ROLLED for (int i = 0; i < N; i++) { MC_SCR(PU_A, i, tid) = T0[i]; MC_SCR(PU_B, i, tid) = T0[i]; }
            I.u_stale = false;
            I.t = 0.0; I.iter = 0; I.u_modified = false; I.accept = false; I.qold = 1e-4; I.q11 = 1.0; I.tstop = 1440.0;
            I.need_initdt = true;                                             /* the two init-dt RHS calls join hour 0's rounds */
            for (int i = 0; i < 24; i++) {
                const int step = jl * 24 + i;
                MC_ROLLED for (int k = 0; k < N; k++) T0_before[k] = T0[k];
                hoist_publish();
                hp.L_prev = L_ode;
                /* step_to_hour! (simulation.jl:888-903), cooperatively */
                I.target = (i + 1) * 60.0; I.tstop = I.target; I.st = 0;
                {
                    const bool pend = coop2_advance<NA>(I, sh, scr, cn);
                    coop2_rounds<NA>(I, pend, hp, sh, scr, cn);
                }
                { const int U = I.uarr(); MC_ROLLED for (int k = 0; k < N; k++) T0[k] = MC_SCR(U, k, tid); }
                if (is_last) {
                    phase_transition(N, m.depths, T0, T0_before, sum_phase, theta);
                    MC_ROLLED for (int k = 0; k < N; k++) T0_output[k] = T0[k];
                }
                if (i < 23) {
                    env = get_instant(r, jl, step + 1, site_albedo);
                    T0[N - 1] = env.T_deep;
                    set_hour_params(env);
                }
                { const int U = I.uarr(); MC_ROLLED for (int k = 0; k < N; k++) MC_SCR(U, k, tid) = T0[k]; }
                I.u_modified = true;
                {   /* init_surface_fluxes! (monin_obukhov.jl:247-264): hourly warm start of the Obukhov length */
                    const double T_air = MC_LDG(hp.fT + (int64_t)i * S);
                    const double wind = jmax(MC_LDG(hp.fwind + (int64_t)i * S), 0.1);
                    const double zen = jmin(90.0, nearbyint(MC_LDG(hp.fzen + (int64_t)i * S) * 1000.0) / 1000.0);
                    const double Tsf = T0[0];
                    if (T_air < Tsf && zen < 90.0) {
                        const Obukhov ob = calc_obukhov(wind, hp.z0, m.zref, hp.lz, T_air - Tsf, calc_rho_cp((Tsf + T_air) / 2.0), L_ode >= 0.0 ? -0.3 : L_ode);
                        cn.n_obukhov += ob.iters;
                        L_ode = ob.L;
                    }
                }
                bool have_infil = false;
                if (is_last) {
                    double rain;
                    if (m.rain_hourly) rain = r.hourly(a.f_rain, step);
                    else rain = (i == midnight_i) ? env.rain_daily : 0.0;
                    pool = jclamp(pool + rain, 0.0, m.max_pool);
                    if (dynamic) {
#if MC_COOP_MOIST
                        if (coop_moist) wet_dyn = soil_water_hour_coop<NA>(r, env, T0, theta, pool, hp.z0, L_prof, psi_out, h_out, flags, cn, sh, ms, scr, mpar);
                        else
#endif
                        wet_dyn = soil_water_hour<NA>(r, env, T0, theta, pool, wet_dyn, hp.z0, L_prof, psi_out, h_out, flags, cn, false);
                        have_infil = true;
                    }
                }
                const int out_row = step + 1;
                if (is_last && out_row < nrows && !(i == 23 && nonconsec)) {
                    r.put_scalar(a.o_pool, out_row, pool);
                    MC_ROLLED for (int k = 0; k < N; k++) { r.put_node(a.o_T, out_row, k, T0_output[k]); r.put_node(a.o_theta, out_row, k, theta[k]); }
                    r.put_scalar(a.o_sky, out_row, sky_T);
                    if (have_infil) { MC_ROLLED for (int k = 0; k < N; k++) { r.put_node(a.o_psi, out_row, k, psi_out[k]); r.put_node(a.o_h, out_row, k, h_out[k]); } }
                    write_props(r, out_row, T0, theta, r.hourly(a.f_P, out_row));
                }
            }
        }
        tdeep_stale = env.T_deep;
    }
    MC_ROLLED for (int i = 0; i < N; i++) {
        a.T[(int64_t)i * S + s] = T0[i]; a.theta[(int64_t)i * S + s] = theta[i]; a.sum_phase[(int64_t)i * S + s] = sum_phase[i];
        if (!(isfinite(T0[i]) && isfinite(theta[i]))) flags |= MC_STATUS_NAN;
    }
    { const int K1 = I.k1arr(); MC_ROLLED for (int i = 0; i < N; i++) a.fsal[(int64_t)i * S + s] = MC_SCR(K1, i, tid); }
    a.pool[s] = pool; a.wetness[s] = wet_dyn; a.L_ode[s] = L_ode; a.L_prof[s] = L_prof; a.tdeep_prev[s] = tdeep_stale;
    a.flags[s] = flags | I.flags;
    rot_ode = I.par;
}


/* ------------------------------------------------------------------------------------------
 * solve_air! (src/simulation.jl:822-851) for the rows of one chunk of site s: atmospheric_surface_profile! (monin_obukhov.jl:85-206)
 * at every height from the row's final surface temperature.  Runs as its own small kernel after the soil kernel of the chunk.
 * ------------------------------------------------------------------------------------------ */
MC_DEV void solve_air_site(const KArgs& a, int64_t s, Counters& cn)
{
    const DevModel& m = MC_CONST_MODEL;
    const int N = m.N, ns = a.n_snow, nh = m.n_heights, H = a.ndays * 24;
    const int64_t S = a.S;
    const double z = m.zref, z0 = MC_LDG(a.z0 + s), kappa = m.kappa, gamma = m.gamma;
    const double lzr = mcm::log(z / z0);
    int ref_idx = -1; for (int k = nh - 1; k >= 0; k--) if (m.heights[k] == z) ref_idx = k;
    double L_air = a.L_air[s];
    for (int row = 0; row < H; row++) {
        const bool has_snow = ns > 0 && a.o_snow_depth && a.o_snow_depth[(int64_t)row * S + s] > 0.0;
        const double T_surf = has_snow ? a.o_snow_T[((int64_t)row * ns + 0) * S + s] : a.o_T[((int64_t)row * N + 0) * S + s];
        const double T_ref = MC_LDG(a.f_T + (int64_t)row * S + s), wind = MC_LDG(a.f_wind + (int64_t)row * S + s);
        const double rh = MC_LDG(a.f_rh + (int64_t)row * S + s), zen = MC_LDG(a.f_zen + (int64_t)row * S + s);
        const double dT = T_ref - T_surf;
        const double rho_cp = calc_rho_cp((T_surf + T_ref) / 2.0);
        double Tair[MC_MAX_HEIGHTS], Qc, ustar;
        if (T_ref >= T_surf || zen >= 90.0) {
            ustar = jmax(kappa * wind / lzr, 1.0e-6);
            const double Stb = 0.64 / lzr, Sts = sublayer_stanton(z0, ustar);
            Qc = convective_flux(rho_cp, dT, ustar, Stb, Sts);
            const double Tz0 = (T_ref * Stb + T_surf * Sts) / (Stb + Sts);
            for (int k = 0; k < nh; k++) {
                const double h = m.heights[k]; double v;
                if (h < z0) { v = 0.0; Tair[k] = T_surf + (Tz0 - T_surf) * (h / z0); }
                else if (k == ref_idx) { v = wind; Tair[k] = T_ref; }
                else { v = (ustar / kappa) * mcm::log(h / z0 + 0.0); Tair[k] = Tz0 + (T_ref - Tz0) * mcm::log(h / z0) / lzr; }
                if (a.o_wind) a.o_wind[((int64_t)row * nh + k) * S + s] = v;
            }
        } else {
            const Obukhov ob = calc_obukhov(wind, z0, z, lzr, dT, rho_cp, L_air);
            cn.n_obukhov += ob.iters;
            L_air = ob.L; Qc = ob.Q; ustar = ob.ustar;
            const double Tz0 = (T_ref * ob.Stb + T_surf * ob.Sts) / (ob.Stb + ob.Sts);
            for (int k = 0; k < nh; k++) {
                const double h = m.heights[k]; double v;
                if (h < z0) { v = 0.0; Tair[k] = T_surf + (Tz0 - T_surf) * (h / z0); }
                else if (k == ref_idx) { v = wind; Tair[k] = T_ref; }
                else {
                    const double x = calc_phi_m(h, gamma, ob.L);
                    const double psi_m1 = calc_psi_m(x), psi_h2 = calc_psi_h(x);
                    const double hr = h / z0;
                    v = (ob.ustar / kappa) * jmax(mcm::log(hr) - psi_m1, 1e-6);
                    Tair[k] = Tz0 + (T_ref - Tz0) * jmax(mcm::log(hr) - psi_h2, 1e-6) / (mcm::log(z / z0) - ob.psi_h);
                }
                if (a.o_wind) a.o_wind[((int64_t)row * nh + k) * S + s] = v;
            }
        }
        const double e_ref = goff_gratch(T_ref) * rh;
        for (int k = 0; k < nh; k++) {
            if (a.o_air_T) a.o_air_T[((int64_t)row * nh + k) * S + s] = Tair[k];
            if (a.o_rh) a.o_rh[((int64_t)row * nh + k) * S + s] = jclamp(e_ref / goff_gratch(Tair[k]) * 1.0, 0.0, 1.0);
        }
        if (a.o_qc) a.o_qc[(int64_t)row * S + s] = Qc;
        if (a.o_ustar) a.o_ustar[(int64_t)row * S + s] = ustar;
    }
    a.L_air[s] = L_air;
}

/* host side: flatten mc_model_desc into DevModel (run once per handle) */
static inline void build_dev_model(const mc_model_desc& md, DevModel& d)
{
    memset(&d, 0, sizeof(d));
    d.N = md.n_soil;
    d.dynamic = md.moisture_strategy == MC_MOIST_DYNAMIC;
    d.nonconsec = md.time_mode == MC_TIME_NONCONSECUTIVE;
    d.conv_kind = md.convergence_kind; d.conv_iters = md.convergence_iterations; d.iters_per_day = md.iterations_per_day;
    d.spinup = md.spinup_first_day; d.rain_hourly = md.rainfall_schedule == MC_RAIN_HOURLY;
    d.flux_alg = md.infiltration_algorithm == MC_INFIL_MATRIC_FLUX_POTENTIAL;
    d.swinbank = md.atmospheric_radiation == MC_ATMRAD_SWINBANK;
    d.moist_max_it = md.moisture_max_iterations;
    d.niter_moist = (int)(3600.0 / md.moisture_timestep_s);
    const int N = md.n_soil;
    double ld[MC_MAX_SOIL_NODES + 1];
    MC_ROLLED for (int i = 0; i < N; i++) { d.depths[i] = md.depths_m[i]; ld[i] = md.depths_m[i]; }
    ld[N] = 0.0;                                          /* zero sentinel (heat_transport_1d.jl:127,173) */
    MC_ROLLED for (int i = 0; i < N; i++) {
        d.dzc[i] = i == 0 ? ld[1] : ld[i + 1] - ld[i - 1];
        d.dzg[i] = ld[i + 1] - ld[i];
    }
    {   /* same expressions as soil_water_hour: depth[1] = -1e10, depth[2] = 0, depth[i+1] = depths[i-1], depth[L+1] = 1e20 */
        const int L = N;
        auto depth_final = [&](int i) -> double { return i == 1 ? -1e10 : (i == 2 ? 0.0 : (i == L + 1 ? 1e20 : d.depths[i - 2])); };
        auto depth_mass = [&](int i) -> double { return i <= 2 ? 0.0 : d.depths[i - 2]; };
        for (int i = 0; i < MAXN + 2; i++) { d.mdz[i] = 0.0; d.mmass[i] = 0.0; }
        for (int i = 1; i <= L; i++) d.mdz[i] = depth_final(i + 1) - depth_final(i);
        for (int i = 2; i <= L; i++) d.mmass[i] = WATER_RHO * (depth_mass(i + 1) - depth_mass(i - 1)) / 2.0;
    }
    d.zref = md.reference_height_m; d.zloc = md.local_height_m; d.reltol = md.reltol; d.abstol = md.abstol;
    d.T_hi = md.maximum_surface_temperature_K;
    d.ga = md.de_vries_shape_factor; d.gc = 1.0 - 2.0 * d.ga; d.recirc_pow = md.recirculation_power; d.theta0 = md.return_flow_threshold;
    d.root_res = md.root_resistance; d.PC = md.stomatal_closure_potential; d.RL = md.leaf_resistance;
    d.SP = md.stomatal_stability_parameter; d.root_radius = md.root_radius;
    d.kappa = md.karman_constant; d.gamma = md.dyer_constant; d.conv_tol = md.convergence_tolerance_K;
    d.moist_tol = md.moisture_tolerance; d.moist_dt = md.moisture_timestep_s; d.max_pool = md.max_surface_pool;
    const mc_fluid_constants& f = md.fluid;
    d.Mw = f.molar_mass_water; d.eps = f.molar_mass_water / f.molar_mass_air; d.dep = f.ideal_gas_departure;
    d.cpa = f.cp_dry_air; d.cpv = f.cp_water_vapour; d.R = f.gas_constant; d.sigma = f.stefan_boltzmann; d.g = f.standard_gravity;
    d.P0 = f.standard_pressure; d.Lv_a = f.latent_heat_a; d.Lv_b = f.latent_heat_b; d.Hv_a = f.molar_latent_heat_a; d.Hv_b = f.molar_latent_heat_b;
    d.Ma_over_R = f.molar_mass_air / f.gas_constant;
    d.schmidt_factor = mcm::pow(0.71 / 0.60, 0.666);
    d.snow = md.snow;
    d.n_heights = md.n_heights;
    for (int i = 0; i < MC_MAX_HEIGHTS; i++) d.heights[i] = i < md.n_heights ? md.heights_m[i] : 0.0;
}
