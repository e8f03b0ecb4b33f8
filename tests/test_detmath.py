"""The product's bit-reproducible math and samplers (csrc/ser_detmath.h; ser_exp_weight of ser_chain_core.h),
validated INDEPENDENTLY of the oracle.  The oracle's free-running mode includes the same header, so the
GPU == oracle parity tests cannot see an error in these functions themselves; here they are checked against
libm (ulp distance) and against scipy's distributions (Kolmogorov-Smirnov), with the header compiled for the
host (tests/emul/chain_emul.cpp).  The GPU executes the same correctly-rounded basic operations, and
tests/test_gpu_parity.py::test_free_running_equals_oracle_philox ties its bits to these.
Reference call sites of what is sampled: gsl_ran_beta in mcmc_samplebeta (mcmc.c:751-765)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def L():
    so = os.path.join(EMUL_DIR, "libchain_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                    os.path.join(EMUL_DIR, "chain_emul.cpp")], check=True)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    for name in ("emul_detmath_log", "emul_detmath_exp", "emul_exp_weight"):
        getattr(lib, name).argtypes = [dp, C.c_int, dp]
    lib.emul_gamma.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.c_int, dp]
    lib.emul_beta.argtypes = [C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_int, dp]
    lib.emul_uniform.argtypes = [C.c_uint32] * 4 + [C.c_int, dp]
    return lib


def _map(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    fn(x.ctypes.data_as(C.POINTER(C.c_double)), x.size, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _ulps(a, b):
    """distance in units of the last place between two arrays of finite doubles of the same sign"""
    ia, ib = a.view(np.int64), b.view(np.int64)
    return np.abs(ia - ib)


def test_ser_log_within_one_ulp_of_libm_over_1e6_points(L):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(400_000),                                   # (0, 1): Beta variates, uniforms
                        np.exp(rng.uniform(-700, 700, 300_000)),               # the whole exponent range
                        1.0 + rng.uniform(-1e-3, 1e-3, 150_000),               # around 1 (cancellation in log1p-like use)
                        1.0 - np.exp(rng.uniform(-7, -0.2, 150_000))])         # 1 - e^c, 1 - e^d of mcmc.c:847-848
    x = x[x > 0]
    assert x.size >= 999_000
    d = _ulps(_map(L.emul_detmath_log, x), np.log(x))
    assert d.max() <= 1, (d.max(), x[d.argmax()])
    assert (d == 0).mean() > 0.7


def test_ser_exp_within_one_ulp_of_libm_over_1e6_points(L):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-700, 700, 400_000), rng.uniform(-40, 1, 400_000), rng.uniform(-1e-3, 1e-3, 200_000)])
    d = _ulps(_map(L.emul_detmath_exp, x), np.exp(x))
    assert d.max() <= 1, (d.max(), x[d.argmax()])


def test_ser_exp_weight_within_two_ulp_of_libm(L):
    """exp of the Gibbs log-weights relative to their maximum (mcmc_logtop's exp, mcmc.c:734): arguments in
    [LOGEPSILON, 0].  FMA Horner form; the decisions only need ~1e-13 relative (tests assert the margins)."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-32.3, 0, 900_000), rng.uniform(-708, -32.3, 100_000), [0.0, -1e-300]])
    got, want = _map(L.emul_exp_weight, x), np.exp(x)
    d = _ulps(got, want)
    assert d.max() <= 2, (d.max(), x[d.argmax()])
    assert _map(L.emul_exp_weight, np.array([-709.0, -1e4]))[0] == 0.0


def test_stream_uniforms_are_uniform_and_53_bit(L):
    u = np.empty(200_000)
    L.emul_uniform(20060206, 5, 9, 5, u.size, u.ctypes.data_as(C.POINTER(C.c_double)))
    assert 0.0 <= u.min() and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert np.all(u * 2.0 ** 53 == np.floor(u * 2.0 ** 53)) and len(np.unique(u)) == u.size
    # serial correlation of consecutive draws (pairs come from one Philox block)
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 0.01


@pytest.mark.parametrize("shape", [1.0, 1.5, 4.0, 77.0, 3001.0, 152000.0])
def test_gamma_sampler_matches_scipy_distribution(L, shape):
    """ser_gamma_ge1 (Marsaglia-Tsang over polar normals), the shapes 1 + count the sweep uses: from 1 (empty
    class) to 1 + t0a ~ 1.5e5 true zeros on g2s2"""
    n = 60_000
    g = np.empty(n)
    L.emul_gamma(shape, 12345, 1000, n, g.ctypes.data_as(C.POINTER(C.c_double)))
    assert g.min() > 0
    assert stats.kstest(g, "gamma", args=(shape,)).pvalue > 1e-3
    assert abs(g.mean() - shape) < 5 * np.sqrt(shape / n) and abs(g.var() / shape - 1) < 0.05


@pytest.mark.parametrize("a,b", [(1.0, 1.0), (3.0, 5.0), (1.0 + 120, 1.0 + 151000), (1.0 + 2500, 1.0 + 1700), (40.0, 2.0)])
def test_beta_sampler_matches_scipy_distribution(L, a, b):
    """Beta(1 + f1a, 1 + t0a) / Beta(1 + f0a, 1 + t1a) of mcmc_samplec / mcmc_sampled (mcmc.c:790, :820) as the ratio
    of two Gammas, incl. the g2s2-sized parameters"""
    n = 60_000
    y = np.empty(n)
    L.emul_beta(a, b, 777, 31, n, y.ctypes.data_as(C.POINTER(C.c_double)))
    assert 0 < y.min() and y.max() < 1
    assert stats.kstest(y, "beta", args=(a, b)).pvalue > 1e-3
    assert abs(y.mean() - a / (a + b)) < 5 * np.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)) / n)
