"""include/tcr_libm.h (shared by the CUDA path and the oracle) against glibc/NumPy: accuracy on the
argument ranges the hot path uses, and exact special values.  fn codes: oracle.orc_libm_eval."""
import numpy as np
import pytest

from oracle import tcr_oracle as orc

FN = dict(exp=0, log=1, sin=2, cos=3, asin=4, tanh=5, sin2pi=6, cos2pi=7)


def _ulps(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


@pytest.mark.parametrize("name,ref,lo,hi", [
    ("exp", np.exp, -100.0, 5.0),            # alpha = 1 - 0.87 exp(-z), z in [0, 100]; pow via exp(p log x)
    ("log", np.log, 1e-6, 1e3),              # pow(t_strat, -0.4), pow(err, -0.2), Box-Muller
    ("sin", np.sin, -7.0, 7.0),              # haversine, genesis latitude
    ("cos", np.cos, -1.6, 1.6),              # cos(lat)
    ("asin", np.arcsin, -1.0, 1.0),
])
def test_accuracy_vs_glibc(name, ref, lo, hi):
    rng = np.random.default_rng(hash(name) % 1000 + 1)
    x = rng.uniform(lo, hi, 200000)
    got = orc.libm_eval(FN[name], x)
    want = ref(x)
    ok = want != 0
    assert _ulps(got[ok], want[ok]).max() < 2.0, name


def test_tanh_absolute_accuracy():
    """tcr_tanh is used only in an additive blend (wind/tc_wind.py:8): absolute accuracy."""
    x = np.random.default_rng(6).uniform(-25.0, 25.0, 200000)
    assert np.max(np.abs(orc.libm_eval(FN["tanh"], x) - np.tanh(x))) < 3e-16


def test_sincos2pi():
    rng = np.random.default_rng(5)
    u = rng.uniform(-4.0, 4.0, 200000)
    s = orc.libm_eval(FN["sin2pi"], u)
    c = orc.libm_eval(FN["cos2pi"], u)
    assert np.max(np.abs(s - np.sin(2 * np.pi * u))) < 4e-15
    assert np.max(np.abs(c - np.cos(2 * np.pi * u))) < 4e-15
    assert np.max(np.abs(s * s + c * c - 1.0)) < 5e-16
    # exact at quarter turns
    q = np.array([0.0, 0.25, 0.5, 0.75, 1.0, -0.25])
    assert np.array_equal(orc.libm_eval(FN["sin2pi"], q), [0.0, 1.0, 0.0, -1.0, 0.0, -1.0])
    assert np.array_equal(orc.libm_eval(FN["cos2pi"], q), [1.0, 0.0, -1.0, 0.0, 1.0, 0.0])


def test_special_values():
    assert orc.libm_eval(FN["exp"], [0.0])[0] == 1.0
    assert orc.libm_eval(FN["log"], [1.0])[0] == 0.0
    assert np.isnan(orc.libm_eval(FN["log"], [-1.0])[0])
    assert np.isnan(orc.libm_eval(FN["exp"], [np.nan])[0])
    assert orc.libm_eval(FN["tanh"], [0.0])[0] == 0.0
    assert np.array_equal(orc.libm_eval(FN["asin"], [1.0, -1.0, 0.0]), [np.pi / 2, -np.pi / 2, 0.0])
    assert np.isnan(orc.libm_eval(FN["asin"], [1.0000001])[0])


@pytest.mark.parametrize("name,ref,lo,hi", [
    ("exp", np.exp, -20.0, 10.0),            # Bolton's saturation vapour pressure exp(min(17.625 Tc / (Tc + 243.04), 10)); Lambert-W iterates
    ("log", np.log, 50.0, 1.1e5),            # log(T), log(p - es rh), log(p_env) of the thermodynamic kernel (thermo/thermo.py:50-76)
    ("log", np.log, 1e-5, 1.0),              # log(rh); log(-z) of the Lambert-W first guess
])
def test_accuracy_on_thermo_ranges(name, ref, lo, hi):
    rng = np.random.default_rng(int(abs(lo) * 10) + 3)
    x = rng.uniform(lo, hi, 200000)
    got = orc.libm_eval(FN[name], x)
    assert _ulps(got, ref(x)).max() < 2.0, name
