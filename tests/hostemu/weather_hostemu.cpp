/* TEST HARNESS ONLY: the product's weather device code (microclimate.jl_b200/csrc/mc_weather.cuh) compiled for the CPU, so the no-GPU tier can
 * check the code the CUDA kernel executes against the oracle's literal restatement.  Mirrors weather_kernel's thread grid with two loops. */
#include "../../microclimate.jl_b200/csrc/mc_weather.cuh"
#include <cstring>

extern "C" int he_populate_weather(int64_t S, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                                   const mc_weather_in_soa* in, mc_weather_out_soa* out)
{
    mcw::WArgs a;
    std::memset(&a, 0, sizeof(a));
    a.S = S; a.ndays = ndays; a.doy = in->day_of_year;
    if (sites) { a.lat = sites->latitude; a.slope = sites->slope; a.aspect = sites->aspect; a.noon_off = sites->solar_noon_offset_h; }
    const double* const* series[4] = {in->reference_temperature_values, in->reference_wind_speed_values, in->reference_humidity_values, in->cloud_cover_values};
    const mc_diel_curve* curves[4] = {&w->reference_temperature, &w->reference_wind_speed, &w->reference_humidity, &w->cloud_cover};
    for (int q = 0; q < 4; q++) for (int i = 0; i < curves[q]->n_inputs; i++) a.vals[q][i] = series[q][i];
    a.noon_in = in->hour_solar_noon; a.ha_in = in->hour_angle_sunrise; a.direct = in->direct_horizontal; a.diffuse = in->diffuse_horizontal;
    a.gr_hourly = in->global_radiation_hourly; a.p_hourly = in->pressure_hourly; a.p_site = in->pressure_site;
    a.o_P = out->pressure; a.o_T = out->reference_temperature; a.o_wind = out->reference_wind_speed; a.o_rh = out->reference_humidity;
    a.o_cloud = out->cloud_cover; a.o_gr = out->global_radiation; a.o_df = out->diffuse_fraction; a.o_zen = out->zenith_angle; a.o_szen = out->zenith_slope_angle;
    a.o_noon = out->hour_solar_noon; a.o_ha = out->hour_angle_sunrise;
    for (int d = 0; d < ndays; d++)
        for (int64_t s = 0; s < S; s++) mcw::site_day(*w, a, s, d);
    return 0;
}
