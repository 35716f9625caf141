"""Device-side generator of the synthetic forcing (BASELINE.json config 5) - measurement tooling for long runs.

`synthetic.make_forcing` (NumPy, host) needs ~6 s of host CPU per 1 Mi-site day, more than the GPU needs to integrate that day,
so a 365-day run of the 1 Mi-site grid would be bounded by host forcing generation.  This module evaluates THE SAME definition
(same counter-based splitmix64 hash, same formulas, synthetic._forcing_block line by line) with torch element-wise ops on the
GPU and hands the arrays to `SoilEngine.step_days_device` without a host round trip.  The uniform variates are bit-identical
to the host generator's (integer hash + exact conversion); the transcendental steps (cos, acos, exp, log, ndtri) use the
device math library, so values agree with the host generator to a few ulp, not bit for bit.  Parity checks on such runs
therefore copy the forcing of the spot-check sites back from the device and give exactly those arrays to the oracle
(tools/run_year.py).  torch is plumbing here (device arrays for a benchmark input), not part of the product path.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi as A
from . import synthetic as syn

_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    """Python int (mod 2^64) -> the int64 with the same bit pattern."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(torch, x, k):
    """Logical shift right of an int64 tensor holding uint64 bit patterns."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(torch, x):
    x = (x ^ _lsr(torch, x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(torch, x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(torch, x, 31)


def u01(torch, stream: int, site, counter=0):
    """synthetic.u01 on the device (int64 arithmetic wraps like uint64): bit-identical variates."""
    key = site * _i64(int(syn._K_SITE)) + counter * _i64(int(syn._K_TIME)) + _i64((stream + 1) * int(syn._K_STREAM) + syn.SEED)
    h = _mix(torch, _mix(torch, key))
    return (_lsr(torch, h, 11).to(torch.float64) + 0.5) * (1.0 / 9007199254740992.0)


class DeviceForcing:
    """Forcing of sites [s0, s1) of an S_total-site grid for successive day chunks, generated into reusable device tensors."""

    def __init__(self, s0: int, s1: int, S_total: int, ndays: int, device):
        import torch
        self.torch, self.s0, self.s1, self.S_total, self.ndays, self.device = torch, s0, s1, S_total, ndays, device
        S, H = s1 - s0, 24 * ndays
        f64 = dict(dtype=torch.float64, device=device)
        self.t = {f: torch.empty((H, S), **f64) for f in A.FORCING_HOURLY if f != "longwave_radiation"}
        self.t.update({f: torch.empty((ndays, S), **f64) for f in A.FORCING_DAILY if f != "albedo_daily"})
        s = torch.arange(s0, s1, dtype=torch.int64, device=device)
        self.s = s
        lat = -60.0 + 130.0 * ((s % S_total).to(torch.float64) + u01(torch, syn.ST_LAT, s)) / S_total
        self.lat = lat
        self.t_mean = 27.0 - 0.55 * lat.abs() + 3.0 * torch.special.ndtri(u01(torch, syn.ST_TMEAN, s))
        self.a_ann = 2.0 + 0.25 * lat.abs()
        self.a_diu = 4.0 + 10.0 * u01(torch, syn.ST_ADIU, s)
        elev = 2500.0 * u01(torch, syn.ST_ELEV, s)
        self.pressure = 101325.0 * (1.0 - 2.25577e-5 * elev) ** 5.2553
        self.lai = 0.1 + 2.9 * u01(torch, syn.ST_LAI, s)
        self.peak = torch.where(lat >= 0.0, torch.full_like(lat, 200.0), torch.full_like(lat, 17.0))

    def _smooth_noise(self, stream, hours, block):
        torch = self.torch
        blk = hours // block
        f = (hours % block).to(torch.float64) / float(block)
        w = f * f * (3.0 - 2.0 * f)
        g0 = torch.special.ndtri(u01(torch, stream, self.s[None, :], blk[:, None]))
        g1 = torch.special.ndtri(u01(torch, stream, self.s[None, :], (blk + 1)[:, None]))
        return (1.0 - w)[:, None] * g0 + w[:, None] * g1

    def generate(self, day0: int):
        """Fill the tensors with days [day0, day0 + ndays); returns ({field: data_ptr}, day_of_year list)."""
        torch, nd, t = self.torch, self.ndays, self.t
        dev = self.device
        hours = torch.arange(24 * day0, 24 * (day0 + nd), dtype=torch.int64, device=dev)
        hr = (hours % 24).to(torch.float64)
        doy = ((hours // 24) % 365 + 1).to(torch.float64)
        seasonal = self.a_ann[None, :] * torch.cos(2.0 * math.pi * (doy[:, None] - self.peak[None, :]) / 365.0)
        t_air = (self.t_mean[None, :] + seasonal + self.a_diu[None, :] * torch.cos(2.0 * math.pi * (hr[:, None] - 15.0) / 24.0)
                 + 1.5 * self._smooth_noise(syn.ST_TNOISE, hours, 6))
        t_daymin = self.t_mean[None, :] + seasonal - self.a_diu[None, :]
        t["reference_temperature"].copy_(t_air + 273.15)
        rh_noise = 0.05 * self._smooth_noise(syn.ST_RH, hours, 3)
        t["reference_humidity"].copy_(torch.clamp(0.85 - 0.03 * (t_air - t_daymin) + rh_noise, 0.05, 1.0))
        blk6 = hours // 6
        wind = torch.clamp_min(torch.exp(math.log(2.5) + 0.5 * torch.special.ndtri(u01(torch, syn.ST_WIND, self.s[None, :], blk6[:, None]))), 0.1)
        t["reference_wind_speed"].copy_(wind)
        days = torch.arange(day0, day0 + nd, dtype=torch.int64, device=dev)
        uc = u01(torch, syn.ST_CLOUD, self.s[None, :], days[:, None])
        cloud_d = (1.0 - (1.0 - uc) ** (1.0 / 0.9)) ** (1.0 / 0.6)
        cloud = cloud_d.repeat_interleave(24, dim=0)
        t["cloud_cover"].copy_(cloud)
        decl = math.radians(23.45) * torch.sin(2.0 * math.pi * (284.0 + doy) / 365.0)
        ha = torch.deg2rad(15.0 * (hr - 12.0))
        latr = torch.deg2rad(self.lat)
        cosz = (torch.sin(latr)[None, :] * torch.sin(decl)[:, None] + torch.cos(latr)[None, :] * torch.cos(decl)[:, None] * torch.cos(ha)[:, None])
        cosz = torch.clamp(cosz, 0.0, 1.0)
        zen = torch.clamp_max(torch.rad2deg(torch.acos(cosz)), 90.0)
        t["zenith_angle"].copy_(zen)
        t["zenith_slope_angle"].copy_(zen)
        solar = torch.clamp_min(1000.0 * cosz * (0.36 + 0.64 * (1.0 - cloud)), 0.0)
        t["global_radiation"].copy_(solar)
        kt = torch.where(cosz > 1e-6, solar / torch.clamp_min(1367.0 * cosz, 1e-9), torch.zeros_like(solar))
        fd = torch.where(kt <= 0.22, 1.0 - 0.09 * kt,
                         torch.where(kt <= 0.80, 0.9511 - 0.1604 * kt + 4.388 * kt ** 2 - 16.638 * kt ** 3 + 12.336 * kt ** 4, torch.full_like(kt, 0.165)))
        t["diffuse_fraction"].copy_(torch.clamp(fd, 0.0, 1.0))
        t["pressure"].copy_(self.pressure[None, :].expand_as(solar))
        ur = u01(torch, syn.ST_RAIN, self.s[None, :], hours[:, None])
        rain = torch.where(ur < 0.04, -1.5 * torch.log(torch.clamp_min(ur, 1e-300) / 0.04), torch.zeros_like(ur))
        t["rainfall_hourly"].copy_(rain)
        S = self.s1 - self.s0
        t["leaf_area_index"].copy_(self.lai[None, :].expand(nd, S))
        t["shade"].zero_()
        t["surface_emissivity"].fill_(0.96)
        t["cloud_emissivity"].fill_(0.96)
        t["soil_wetness"].zero_()
        t["deep_soil_temperature"].copy_((self.t_mean + 273.15)[None, :].expand(nd, S))
        t["rainfall_daily"].copy_(rain.reshape(nd, 24, S).sum(dim=1))
        self.doy = [(d % 365) + 1 for d in range(day0, day0 + nd)]
        return {k: v.data_ptr() for k, v in t.items()}, self.doy

    def host_columns(self, cols) -> "syn.M.ForcingBatch":
        """The current chunk's forcing of the given local site columns as a host ForcingBatch (for the oracle spot check)."""
        torch = self.torch
        idx = torch.as_tensor(np.asarray(cols, dtype=np.int64), device=self.device)
        fb = syn.M.ForcingBatch(self.ndays, len(cols), hourly_rain=True, day_of_year=self.doy)
        for k, v in self.t.items():
            getattr(fb, k)[...] = v.index_select(1, idx).cpu().numpy()
        return fb
