/*
 * mc_physics.cuh - Psychrometrics, longwave, Monin-Obukhov, Campbell-de Vries soil properties and the no-snow soil energy balance (RHS of the ODE).
 * Part of mc_device.cuh (included from there, inside namespace mcd, in this order); not a stand-alone header.
 */

/* ------------------------------------------------------------------------------------------
 * FluidProperties.jl psychrometrics (un-vendored; constants are data, see mc_fluid_constants)
 * ------------------------------------------------------------------------------------------ */
/* vapour_pressure(GoffGratch(), T) [Pa] */
MC_DEV_NOINLINE double goff_gratch(double T)
{
    const double Tp = T + 0.01;
    double loge;
    if (Tp <= 273.16) {
        const double r = MCM_DIV(273.16, Tp);
        loge = -9.09718 * (r - 1.0) - 3.56654 * mcm::log10(r) + 0.876793 * (1.0 - MCM_DIV(Tp, 273.16)) + 0.7858350313586662 /* mcm::log10(6.1071) */;
    } else {
        const double r = MCM_DIV(373.16, Tp);
        loge = -7.90298 * (r - 1.0) + 5.02808 * mcm::log10(r)
             - 1.3816e-7 * (mcm::exp10(11.344 * (1.0 - MCM_DIV(Tp, 373.16))) - 1.0)
             + 8.1328e-3 * (mcm::exp10(-3.49149 * (r - 1.0)) - 1.0) + 3.0057148979490314 /* mcm::log10(1013.246) */;
    }
    return 100.0 * mcm::exp10(loge);
}

/* goff_gratch at T-1, T, T+1 (the central difference of campbell_devries.jl:117-122) evaluated side by side: the three
 * evaluations are independent, so interleaving them triples the instruction-level parallelism of the hottest code in the
 * path.  Every element goes through exactly the operations of goff_gratch() above (bit-identical results). */
struct GG3 { double em, e0, ep; };
MC_DEV_NOINLINE GG3 goff_gratch3(double T)
{
    GG3 o;
    const double t0 = (T - 1.0) + 0.01, t1 = T + 0.01, t2 = (T + 1.0) + 0.01;
    if (t0 > 273.16 && t2 < 400.0) {                       /* all three on the over-water branch (t0 < t1 < t2) */
        const double r0 = 373.16 / t0, r1 = 373.16 / t1, r2 = 373.16 / t2;
        const double a0 = 11.344 * (1.0 - t0 / 373.16), a1 = 11.344 * (1.0 - t1 / 373.16), a2 = 11.344 * (1.0 - t2 / 373.16);
        const double b0 = -3.49149 * (r0 - 1.0), b1 = -3.49149 * (r1 - 1.0), b2 = -3.49149 * (r2 - 1.0);
        const double pa0 = mcm::exp10_arg(a0), pa1 = mcm::exp10_arg(a1), pa2 = mcm::exp10_arg(a2);
        const double pb0 = mcm::exp10_arg(b0), pb1 = mcm::exp10_arg(b1), pb2 = mcm::exp10_arg(b2);
        /* arguments are O(1) for 273 K < T < 400 K: always inside the fast range of exp / log */
        const double ea0 = mcm::exp10_finish(mcm::exp_core(pa0), a0, pa0), ea1 = mcm::exp10_finish(mcm::exp_core(pa1), a1, pa1), ea2 = mcm::exp10_finish(mcm::exp_core(pa2), a2, pa2);
        const double eb0 = mcm::exp10_finish(mcm::exp_core(pb0), b0, pb0), eb1 = mcm::exp10_finish(mcm::exp_core(pb1), b1, pb1), eb2 = mcm::exp10_finish(mcm::exp_core(pb2), b2, pb2);
        const double l0 = mcm::log_core(r0) * 4.34294481903251816668e-01, l1 = mcm::log_core(r1) * 4.34294481903251816668e-01, l2 = mcm::log_core(r2) * 4.34294481903251816668e-01;
        const double g0 = -7.90298 * (r0 - 1.0) + 5.02808 * l0 - 1.3816e-7 * (ea0 - 1.0) + 8.1328e-3 * (eb0 - 1.0) + 3.0057148979490314;
        const double g1 = -7.90298 * (r1 - 1.0) + 5.02808 * l1 - 1.3816e-7 * (ea1 - 1.0) + 8.1328e-3 * (eb1 - 1.0) + 3.0057148979490314;
        const double g2 = -7.90298 * (r2 - 1.0) + 5.02808 * l2 - 1.3816e-7 * (ea2 - 1.0) + 8.1328e-3 * (eb2 - 1.0) + 3.0057148979490314;
        const double p0 = mcm::exp10_arg(g0), p1 = mcm::exp10_arg(g1), p2 = mcm::exp10_arg(g2);
        o.em = 100.0 * mcm::exp10_finish(mcm::exp_core(p0), g0, p0);
        o.e0 = 100.0 * mcm::exp10_finish(mcm::exp_core(p1), g1, p1);
        o.ep = 100.0 * mcm::exp10_finish(mcm::exp_core(p2), g2, p2);
    } else {
        o.em = goff_gratch(T - 1.0); o.e0 = goff_gratch(T); o.ep = goff_gratch(T + 1.0);
    }
    return o;
}

struct WetAir { double vapour_pressure, vapour_density, specific_heat, density; };

#if MC_WETAIR_NOINLINE
MC_DEV_NOINLINE
#else
MC_DEV
#endif
WetAir wet_air(double T, double rh, double P)
{
    const DevModel& m = MC_CONST_MODEL;
    WetAir w;
    const double e = goff_gratch(T) * rh;
    w.vapour_pressure = e;
    const double mr = (m.eps * m.dep * e) / (P - m.dep * e);
    w.vapour_density = e * m.Mw / (0.998 * m.R * T);
    const double T_vir = T * ((1.0 + mr / m.eps) / (1.0 + mr));
    w.density = m.Ma_over_R * P / (0.999 * T_vir);
    w.specific_heat = (m.cpa + mr * m.cpv) / (1.0 + mr);
    return w;
}
MC_DEV double vapour_density_sat(double T) { const DevModel& m = MC_CONST_MODEL; return goff_gratch(T) * m.Mw / (0.998 * m.R * T); }
MC_DEV double latent_heat(double T) { const DevModel& m = MC_CONST_MODEL; return m.Lv_a - m.Lv_b * (T - T0C); }

/* atmospheric_radiation (Campbell-Norman / Swinbank), src/radiation/atmospheric */
MC_DEV double atmospheric_radiation(double e_Pa, double T_air)
{
    const DevModel& m = MC_CONST_MODEL;
    if (m.swinbank) return (9.2e-6 * T_air * T_air) * m.sigma * mcm::pow4(T_air);
    const double Tp = T_air + 0.01;
    return (1.72 * mcm::pow((e_Pa / 1000.0) / Tp, 1.0 / 7.0)) * m.sigma * mcm::pow4(Tp);
}

/* precompute_longwave_sky -> sky temperature (view_factor.jl:23-67) */
MC_DEV_NOINLINE double sky_temperature(double svf, double /*P*/, double rh, double T_ref, double eps_cloud, double cloud, double shade,
                              bool has_lw, double supplied_lw)
{
    const DevModel& m = MC_CONST_MODEL;
    const double cloud_rad = m.sigma * eps_cloud * mcm::pow4(T_ref - 2.0);
    const double hill_rad = m.sigma * eps_cloud * mcm::pow4(T_ref);
    double lw_sky;
    if (has_lw) lw_sky = supplied_lw * (1.0 - shade);
    else {
        const double e = goff_gratch(T_ref) * rh;
        lw_sky = (atmospheric_radiation(e, T_ref) * (1.0 - cloud) + cloud_rad * cloud) * (1.0 - shade);
    }
    const double incoming = (lw_sky + shade * hill_rad) * svf + hill_rad * (1.0 - svf);
    return mcm::pow(incoming / m.sigma, 0.25);
}

/* ------------------------------------------------------------------------------------------
 * Monin-Obukhov (src/boundary_layer/monin_obukhov.jl)
 * ------------------------------------------------------------------------------------------ */
MC_DEV double calc_rho_cp(double Tmean) { return (0.08472 / Tmean) * 4.184e6; }
MC_DEV double sublayer_stanton(double z0, double ustar) { return 0.62 / mcm::pow(z0 * 100.0 * (ustar * 6000.0) / 12.0, 0.45); }
MC_DEV double convective_flux(double rho_cp, double dT, double ustar, double Stb, double Sts) { return rho_cp * dT * ustar * Stb / (1.0 + Stb / Sts); }
MC_DEV double calc_phi_m(double z, double gamma, double L) { return sqrt(sqrt(1.0 - gamma * z / L)); }
MC_DEV double calc_psi_m(double x) { return 2.0 * mcm::log((1.0 + x) / 2.0) + mcm::log((1.0 + x * x) / 2.0) - 2.0 * mcm::atan(x) + PI_D / 2.0; }
MC_DEV double calc_psi_h(double x) { return 2.0 * mcm::log((1.0 + x * x) / 2.0); }

struct Obukhov { double L, Sts, Stb, ustar, psi_h, Q; int iters; };

/* calc_Obukhov_length fixed point (monin_obukhov.jl:527-577): <=30 iterations, 1 % relative change */
MC_DEV_NOINLINE Obukhov calc_obukhov(double v_ref, double z0, double z, double lz, double dT, double rho_cp, double L0)
{
    const DevModel& m = MC_CONST_MODEL;
    Obukhov o; o.L = L0; o.Q = 0.0; o.Stb = 0.0; o.Sts = 0.0; o.psi_h = 0.0; o.ustar = 0.0;
    double rel = 1.0; int count = 0;
    while (rel > 1.0e-2 && count < 30) {
        count++;
        const double x = calc_phi_m(z, 16.0, o.L);
        const double psi_m = calc_psi_m(x);
        o.psi_h = calc_psi_h(x);
        double lrc = lz - psi_m;
        if (lrc <= 0.0) lrc = 1.0e-6;
        o.ustar = m.kappa * v_ref / lrc;
        if (o.ustar < 1.0e-6) o.ustar = 1.0e-6;
        o.Sts = sublayer_stanton(z0, o.ustar);
        o.Stb = (0.64 / lrc) * (1.0 - 0.1 * z / o.L);
        o.Q = convective_flux(rho_cp, dT, o.ustar, o.Stb, o.Sts);
        const double Lnew = RHOCPTKG * o.ustar * o.ustar * o.ustar / o.Q;
        rel = fabs((Lnew - o.L) / o.L);
        o.L = Lnew;
    }
    o.iters = count;
    return o;
}

/* ------------------------------------------------------------------------------------------
 * Per-site context: what a thread keeps for its site across the whole chunk
 * ------------------------------------------------------------------------------------------ */
template <int NA>
struct HourSoil {           /* theta-only soil property terms, rebuilt once per hour (theta is constant within the ODE hour) */
    double cap[NA];         /* areal heat capacity of the node's control volume, J/m2/K            (heat_transport_1d.jl:175-190) */
    double theta[NA], phim[NA], phig[NA], fw[NA], lam_m[NA];
};

struct HourPar {            /* SoilEnergyInputs scalars (heat_transport_1d.jl:41-63) */
    double shade, eps_surf, eps_cloud, soil_wetness, albedo, L_prev;
    double z0, lz, slope, svf;          /* lz = mcm::log(zref/z0) */
    const double *fP, *fT, *fwind, *frh, *fcloud, *fsolar, *fdiff, *fzen, *fszen;   /* this site's day: base[k*S] is knot k */
    const double* ftab[7];              /* {P, T, wind, zen, solar, cloud, rh}: the seven always-used interpolants, for a rolled loop */
    int64_t S;
    const struct SnowCtx* sn;           /* SnowModel context (NULL without snow) */
};

MC_DEV double knot_lerp(const double* base, int64_t S, int i, double d)
{
    const int i1 = i + 1 > 23 ? 23 : i + 1;      /* 25th knot repeats hour 24 (simulation.jl:990) */
    return (1.0 - d) * MC_LDG(base + (int64_t)i * S) + d * MC_LDG(base + (int64_t)i1 * S);
}

/* Campbell-de Vries bulk conductivity of one node (campbell_devries.jl:79-142), T-dependent part only */
#ifndef MC_WEIGHT_NOINLINE
#define MC_WEIGHT_NOINLINE 0
#endif
#if MC_WEIGHT_NOINLINE
MC_DEV_NOINLINE
#else
MC_DEV
#endif
double devries_weight(double lam_c, double lam_f, double ga, double gc)
{
    const double r1 = MCM_DIV(lam_c, lam_f) - 1.0;
    return MCM_DIV(2.0, 3.0 * (1.0 + ga * r1)) + MCM_DIV(1.0, 3.0 * (1.0 + gc * r1));
}
MC_DEV_NOINLINE double soil_conductivity(double T, double P, double p0_over_p, double p_over_p0, double theta, double phim, double phig, double fw, double lam_m)
{
    const DevModel& m = MC_CONST_MODEL;
    const double Tc = T - T0C;
    const double lam_w = 0.554 + 2.24e-3 * Tc - 9.87e-6 * Tc * Tc;
    const double lam_a = 0.024 + 7.73e-5 * Tc - 2.6e-8 * Tc * Tc;
    const double Dv = 2.12e-5 * p0_over_p * mcm::pow(MCM_DIV(T, 273.15), 1.75);       /* campbell_devries.jl:109; p0_over_p = P0/P hoisted by the caller */
    const double rho_hat = 44.65 * p_over_p0 * MCM_DIV(273.15, T);                    /* :111 */
    const double Hv = m.Hv_a - m.Hv_b * (T - T0C);
#if MC_USE_GG3
    const GG3 gg = goff_gratch3(T);
    const double e0 = gg.e0 * 0.99;
    const double em = gg.em * 0.99;
    const double ep = gg.ep * 0.99;
#else
    const double e0 = goff_gratch(T) * 0.99;
    const double em = goff_gratch(T - 1.0) * 0.99;
    const double ep = goff_gratch(T + 1.0) * 0.99;
#endif
    const double dedT = (ep - em) / 2.0;
    const double lam_g = lam_a + MCM_DIV(Hv * dedT * fw * rho_hat * Dv, P - e0);
    const double lam_f = lam_g + fw * (lam_w - lam_g);
    const double ww = devries_weight(lam_w, lam_f, m.ga, m.gc);
    const double wm = devries_weight(lam_m, lam_f, m.ga, m.gc);
    const double wg = devries_weight(lam_g, lam_f, m.ga, m.gc);
    return MCM_DIV(theta * ww * lam_w + phim * wm * lam_m + phig * wg * lam_g, theta * ww + phim * wm + phig * wg);
}

/* full soil_properties for the output rows (update_soil_properties!, simulation.jl:905-913) */
MC_DEV void soil_properties_full(double T, double P, double theta, double bd, double md, double lam_m, double c_m,
                                 double& lam, double& cp, double& rho)
{
    const DevModel& m = MC_CONST_MODEL;
    cp = bd / md * c_m + theta * WATER_CP;
    rho = theta * WATER_RHO + bd;
    const double phim = bd / md;
    const double phig = jmax(0.0, 1.0 - theta - phim);
    const double fw = 1.0 / (1.0 + mcm::pow(theta / m.theta0, -m.recirc_pow));
    lam = soil_conductivity(T, P, m.P0 / P, P / m.P0, theta, phim, phig, fw, lam_m);
}

/* ------------------------------------------------------------------------------------------
 * soil_energy_balance (heat_transport_1d.jl:137-302), no-snow form: the surface is node 0.
 * u, du in K and K/min.
 * ------------------------------------------------------------------------------------------ */
template <int NA>
MC_DEV_NOINLINE void soil_rhs(const HourSoil<NA>& hs, const HourPar& hp, const double* __restrict__ u, double t,
                              double* __restrict__ du, Counters& cn)
{
    const DevModel& m = MC_CONST_MODEL;
    const int N = m.N;
    cn.n_rhs++;
    /* interpolate_forcings (forcing.jl:24-37) */
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

    /* conduction: node conductance G_i = lambda_i(T_i) / dz_i, built on the fly so only two live at a time; the next node's
     * hoisted terms are loaded BEFORE the long conductivity call so their latency hides behind it */
    const double Ts_node = jclamp(u[0], T_LO, m.T_hi);
    double T_up = Ts_node;
    double th = hs.theta[1], pm = hs.phim[1], pg = hs.phig[1], fw = hs.fw[1], lm = hs.lam_m[1], cap = hs.cap[1];
    double T_i = jclamp(u[1], T_LO, m.T_hi);
    const double p0p = m.P0 / P, pp0 = P / m.P0;
    double G_up = soil_conductivity(T_up, P, p0p, pp0, hs.theta[0], hs.phim[0], hs.phig[0], hs.fw[0], hs.lam_m[0]) / m.dzg[0];
    const double Q_cond = G_up * (T_i - Ts_node);
    MC_ROLLED for (int i = 1; i < N - 1; i++) {
        const int in = i + 1 < N - 1 ? i + 1 : i;
        const double T_dn = jclamp(u[i + 1], T_LO, m.T_hi);
        const double nth = hs.theta[in], npm = hs.phim[in], npg = hs.phig[in], nfw = hs.fw[in], nlm = hs.lam_m[in], ncap = hs.cap[in];
        const double G_i = soil_conductivity(T_i, P, p0p, pp0, th, pm, pg, fw, lm) / m.dzg[i];
        du[i] = 60.0 * (G_up * (T_up - T_i) + G_i * (T_dn - T_i)) / cap;
        T_up = T_i; T_i = T_dn; G_up = G_i;
        th = nth; pm = npm; pg = npg; fw = nfw; lm = nlm; cap = ncap;
    }
    du[N - 1] = 0.0;                                         /* Dirichlet bottom node (:291-293) */

    /* shortwave (:223-232) */
    double Q_solar = (1.0 - hp.albedo) * solar * (1.0 - hp.shade);
    if (hp.slope > 0.0 && zen < 90.0) {
        const double szen = jmin(90.0, knot_lerp(hp.fszen, hp.S, ik, dk));
        const double fd = jclamp(knot_lerp(hp.fdiff, hp.S, ik, dk), 0.0, 1.0);
        const double cz = mcm::cos(zen * PI_D / 180.0), csz = mcm::cos(szen * PI_D / 180.0);
        Q_solar = jmax(0.0, (((1.0 - fd) * Q_solar) / cz) * csz) + (fd * Q_solar) * hp.svf;
    }
    /* longwave (:237-250) */
    const WetAir wa = wet_air(T_air, rh, P);
    const double atm = atmospheric_radiation(wa.vapour_pressure, T_air);
    const double cloud_rad = m.sigma * hp.eps_cloud * mcm::pow4(T_air - 2.0);
    const double hill_rad = m.sigma * hp.eps_cloud * mcm::pow4(T_air);
    const double lw_sky = (atm * (1.0 - cloud) + cloud_rad * cloud) * (1.0 - hp.shade);
    const double incoming = (lw_sky + hp.shade * hill_rad) * hp.svf + hill_rad * (1.0 - hp.svf);
    const double Q_ir = incoming - ((1.0 - hp.shade) * m.sigma * hp.eps_surf) * mcm::pow4(Ts_node) - hp.shade * hill_rad;

    /* surface temperature guards (:260-266) */
    double Ts = Ts_node;
    if (fabs(Ts - T0C) > 100.0) Ts = T_air + 1.0;
    if (Ts == T_air) Ts = Ts + 0.1;

    /* surface_fluxes (monin_obukhov.jl:208-245) */
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
    /* surface_convection_evaporation + evaporation (bulk_transfer.jl:11-81) */
    const double hc = dT == 0.0 ? 0.5 : jmax(fabs(Qc / dT), 0.5);
    const double hd = (hc / (wa.specific_heat * wa.density)) * m.schmidt_factor;
    const double Tsc = Ts < T_LO ? T_LO : Ts;
    const double Qe = (hp.soil_wetness * hd * (vapour_density_sat(Tsc) - wa.vapour_density)) * latent_heat(Tsc);

    du[0] = 60.0 * (Q_solar + Q_ir + Q_cond + Qc + 0.0 /* Q_freeze: snow only */ - Qe) / hs.cap[0];
}
