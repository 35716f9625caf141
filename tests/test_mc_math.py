"""include/mc_math.h (the bit-reproducible math kernel shared by the CUDA engine and the oracle) against the host libm.
exp/log/atan/cos: <= 1 ulp; log10: <= 2 ulp; pow/exp10/exp2: <= 2e-15 relative over the ranges the soil path uses."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mathlib")
SO = os.path.join(HERE, "libmc_math_test.so")
PD = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def L():
    src = os.path.join(HERE, "mc_math_export.cpp")
    hdr = os.path.join(HERE, "..", "..", "include", "mc_math.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", SO, src])
    return C.CDLL(SO)


def ap(L, fn, x, y=None):
    x = np.ascontiguousarray(x, float)
    y = np.ascontiguousarray(x if y is None else y, float)
    o = np.empty_like(x)
    L.mcm_apply(fn, x.ctypes.data_as(PD), y.ctypes.data_as(PD), o.ctypes.data_as(PD), C.c_long(len(x)))
    return o


def ulps(a, ref):
    ref = np.asarray(ref, float)
    return np.abs(np.asarray(a) - ref) / np.spacing(np.abs(ref))


RNG = np.random.default_rng(1)
N = 40000


def test_exp_log(L):
    x = np.concatenate([RNG.uniform(-700, 700, N), RNG.uniform(-2, 2, N), RNG.uniform(-1e-3, 1e-3, 1000)])
    assert ulps(ap(L, 0, x), [math.exp(v) for v in x]).max() <= 1.0
    x = np.concatenate([np.exp(RNG.uniform(-700, 700, N)), RNG.uniform(0.5, 2, N), 1 + RNG.uniform(-1e-6, 1e-6, 1000), [5e-324, 1e-310]])
    assert ulps(ap(L, 1, x), [math.log(v) for v in x]).max() <= 1.0
    assert ulps(ap(L, 2, x), [math.log10(v) for v in x]).max() <= 2.0


def test_exp10_exp2_pow(L):
    x = RNG.uniform(-300, 300, N)
    assert ulps(ap(L, 3, x), [10.0 ** v for v in x]).max() <= 2.0
    x = RNG.uniform(-100, 100, N)
    assert ulps(ap(L, 7, x), [2.0 ** v for v in x]).max() <= 2.0
    x, y = np.exp(RNG.uniform(-20, 20, N)), RNG.uniform(-12, 12, N)
    r, ref = ap(L, 4, x, y), np.array([math.pow(a, b) for a, b in zip(x, y)])
    assert np.abs(r / ref - 1).max() < 2e-15
    # the exponents the path actually uses
    for yy in (1.75, 0.45, 1.0 / 7.0, 0.25, 0.666, -4.0, 10.0, 3.1, 1.0 / 3.1, 2.0 + 3.0 / 3.1):
        x = np.exp(RNG.uniform(-12, 12, 4000))
        r, ref = ap(L, 4, x, np.full_like(x, yy)), np.array([math.pow(a, yy) for a in x])
        assert np.abs(r / ref - 1).max() < 2e-15, yy


def test_atan_cos(L):
    x = np.concatenate([RNG.uniform(-50, 50, N), RNG.uniform(-3, 3, N), np.exp(RNG.uniform(-40, 60, 2000))])
    assert ulps(ap(L, 5, x), [math.atan(v) for v in x]).max() <= 1.0
    x = RNG.uniform(-2.3, 2.3, N)
    assert ulps(ap(L, 6, x), [math.cos(v) for v in x]).max() <= 1.0


def test_special_values(L):
    assert np.array_equal(ap(L, 0, [-np.inf, np.inf, 800.0, -800.0]), [0.0, np.inf, np.inf, 0.0])
    assert np.isnan(ap(L, 0, [np.nan])[0])
    r = ap(L, 1, [0.0, -1.0, np.inf])
    assert r[0] == -np.inf and np.isnan(r[1]) and r[2] == np.inf
    r = ap(L, 4, [0.0, 0.0, -2.0, -2.0, -2.0, 3.0], [-4.0, 2.0, 10.0, 3.0, 0.5, 0.0])
    assert r[0] == np.inf and r[1] == 0.0 and r[2] == pytest.approx(1024.0, rel=1e-15) and r[3] == pytest.approx(-8.0, rel=1e-15) and np.isnan(r[4]) and r[5] == 1.0
    r = ap(L, 5, [np.inf, -np.inf])
    assert r[0] == pytest.approx(math.pi / 2) and r[1] == pytest.approx(-math.pi / 2)


def test_solar_geometry_functions(L):
    """sin_r / cos_r (|x| <= 64), acos, sincosd: the functions the solar-geometry and weather kernels add (<= 1 ulp, exact at multiples of 90 degrees)."""
    x = np.concatenate([RNG.uniform(-64, 64, N), RNG.uniform(-7, 7, N), RNG.uniform(-1e-3, 1e-3, 1000), [0.0, math.pi / 2, math.pi, -math.pi, 2 * math.pi, 6.3]])
    # near the zeros of sin/cos a two-word pi/2 leaves an ABSOLUTE error ~1e-21 |n|: compare in ulps away from them, absolutely near them
    s, c = ap(L, 8, x), ap(L, 9, x)
    rs, rc = np.sin(x), np.cos(x)
    for got, ref in ((s, rs), (c, rc)):
        big = np.abs(ref) > 1e-3
        assert ulps(got[big], ref[big]).max() <= 1.0
        assert np.abs(got[~big] - ref[~big]).max() < 1e-18
    assert np.isnan(ap(L, 8, [65.0, np.nan])).all() and np.isnan(ap(L, 9, [-100.0])).all()
    x = np.concatenate([RNG.uniform(-1, 1, N), RNG.uniform(-1e-3, 1e-3, 1000), 1 - np.exp(RNG.uniform(-40, -1, 2000)), -1 + np.exp(RNG.uniform(-40, -1, 2000)), [1.0, -1.0, 0.0, 0.5, -0.5]])
    a = ap(L, 10, x)
    assert ulps(a, np.arccos(x)).max() <= 1.0 and a[-5] == 0.0 and a[-4] == math.pi
    assert np.isnan(ap(L, 10, [1.0000001, -2.0, np.nan])).all()
    d = np.concatenate([RNG.uniform(-720, 720, N), np.arange(-720.0, 721.0, 15.0)])
    sd, cd = ap(L, 11, d), ap(L, 12, d)
    rad = np.fmod(d, 360.0).astype(np.longdouble) * (np.longdouble("3.14159265358979323846264338327950288") / 180)      # 80-bit reference
    rs, rc = np.sin(rad).astype(float), np.cos(rad).astype(float)
    assert np.abs(sd - rs).max() < 2.3e-16 and np.abs(cd - rc).max() < 2.3e-16
    for deg, es, ec in ((0, 0, 1), (90, 1, 0), (180, 0, -1), (270, -1, 0), (360, 0, 1), (-90, -1, 0), (450, 1, 0), (30, None, None)):
        sv, cv = ap(L, 11, [float(deg)])[0], ap(L, 12, [float(deg)])[0]
        if es is not None:
            assert sv == es and cv == ec, deg
    assert abs(ap(L, 11, [30.0])[0] - 0.5) <= 1.2e-16 and abs(ap(L, 12, [60.0])[0] - 0.5) <= 1.2e-16
