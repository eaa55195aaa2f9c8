"""The device's Dirichlet fit (thrifty_b200/csrc/dirichlet_lm.cuh) compiled for the host and compared with what the
reference calls: scipy.optimize.curve_fit -> leastsq -> MINPACK lmdif (thrifty/carrier_sync.py:185-189).

The header is plain scalar C++; tests/native/lm_harness.cpp wraps it for ctypes.  The CUDA kernel runs the same source
(one lane per point for the sines), so agreement here is agreement of the iteration itself: the same iterates, the same
number of function evaluations, the same stopping test firing."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.optimize import curve_fit, leastsq

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "lm_harness.cpp")
OUT = os.path.join(HERE, "native", "_build", "liblm_harness.so")


@pytest.fixture(scope="module")
def lm():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(HERE, "..", "thrifty_b200", "csrc", "dirichlet_lm.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        # no -march flags: without FMA contraction the host build rounds exactly like NumPy / MINPACK
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, SRC], check=True)
    lib = ctypes.CDLL(OUT)
    lib.lm_fit_host.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 3
    lib.lm_fit_host.restype = ctypes.c_int

    def fit(y, carrier_len, block_len):
        y = np.ascontiguousarray(y, dtype=np.float64)
        a, d, nf = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        info = lib.lm_fit_host(y.ctypes.data, float(carrier_len), float(block_len), ctypes.byref(a), ctypes.byref(d),
                               ctypes.byref(nf))
        return a.value, d.value, nf.value, info
    return fit


XD = np.arange(-3, 4)


def kern(x, w, n):          # thrifty/carrier_sync.py:121-132
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.sin(np.pi * w * x / n) / np.sin(np.pi * x / n) / w
        v[np.isnan(v)] = 1
    return v


def random_case(rng):
    n = int(rng.choice([1024, 4096, 16384, 32768]))
    w = int(rng.integers(32, n // 2))
    d_true = rng.uniform(-0.7, 0.7)
    amp = rng.uniform(1, 1000)
    y = amp * np.abs(kern(XD - d_true, w, n)) + rng.normal(0, amp * rng.choice([1e-6, 1e-3, 3e-2, 0.2]), 7)
    return n, w, np.abs(y).astype(np.float32)        # float32 magnitudes, as Signal.mag delivers them


def test_lm_matches_minpack_iterate_for_iterate(lm):
    rng = np.random.default_rng(20161125)
    for _ in range(3000):
        n, w, y = random_case(rng)
        y64 = y.astype(np.float64)
        ref = leastsq(lambda p: p[0] * np.abs(kern(XD - p[1], w, n)) - y64, (y64[3], 0.0), full_output=1)
        a, d, nfev, info = lm(y, w, n)
        assert nfev == ref[2]["nfev"] and info == ref[4], (n, w, y)
        assert a == ref[0][0] and d == ref[0][1], (n, w, y, ref[0], (a, d))


def test_lm_matches_curve_fit_on_reference_vectors(lm):
    """The reference's own interpolator test (tests/test_carrier_sync.py:44-65: 10 offsets, N=8192, W=2085) through
    curve_fit with the reference's model function."""
    n, w = 8192, 2085
    for offset in np.linspace(-0.5, 0.5, 10):
        y = (25.0 * np.abs(kern(XD - offset, w, n))).astype(np.float32)

        def model(xdata, amplitude, time_offset):          # carrier_sync.py:180-183
            return amplitude * np.abs(kern(np.array(xdata, dtype=np.float64) - time_offset, w, n))
        popt, _ = curve_fit(model, XD, y, p0=(y[3], 0))
        _, d, _, _ = lm(y, w, n)
        assert d == popt[1]
        assert abs(d - offset) < 1e-6


def test_lm_degenerate_inputs(lm):
    # all-equal magnitudes, zeros around a lone peak, huge dynamic range: must terminate like MINPACK does
    for y in ([1, 1, 1, 1, 1, 1, 1], [0, 0, 0, 5, 0, 0, 0], [1e-6, 1e-3, 1, 1e4, 1, 1e-3, 1e-6], [3, 2, 1, 4, 9, 1, 0.5]):
        y = np.asarray(y, dtype=np.float32)
        y64 = y.astype(np.float64)
        n, w = 4096, 1226
        ref = leastsq(lambda p: p[0] * np.abs(kern(XD - p[1], w, n)) - y64, (y64[3], 0.0), full_output=1)
        a, d, nfev, info = lm(y, w, n)
        assert (nfev, info) == (ref[2]["nfev"], ref[4])
        assert d == ref[0][1] or (np.isnan(d) and np.isnan(ref[0][1]))


@pytest.fixture(scope="module")
def quick(lm):
    lib = ctypes.CDLL(OUT)
    lib.lm_quick_host.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 3
    lib.lm_quick_host.restype = ctypes.c_int

    def run(y, carrier_len, block_len):
        y = np.ascontiguousarray(y, dtype=np.float64)
        a, d, s = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        ok = lib.lm_quick_host(y.ctypes.data, float(carrier_len), float(block_len), ctypes.byref(a), ctypes.byref(d),
                               ctypes.byref(s))
        return bool(ok), d.value, s.value
    return run


def test_short_cut_only_stands_in_where_it_agrees_with_minpack(quick):
    """quick_fit (Gauss-Newton to the least-squares minimum) may replace lmdif only when it reports ok; wherever it does,
    its offset must sit within a small fraction of the 1e-4 parity bar of what scipy returns."""
    rng = np.random.default_rng(7)
    n_ok, worst = 0, 0.0
    for _ in range(4000):
        n = int(rng.choice([1024, 4096, 16384, 32768]))
        u = rng.random()
        w = int(rng.integers(32, n // 2)) if u < 0.4 else int(n * (rng.uniform(0.02, 0.15) if u < 0.7 else rng.uniform(0.25, 0.35)))
        amp = rng.uniform(1, 1000)
        y = amp * np.abs(kern(XD - rng.uniform(-0.7, 0.7), w, n)) + rng.normal(0, amp * rng.choice([1e-6, 1e-4, 1e-3, 1e-2, 3e-2, 0.1]), 7)
        y = np.abs(y).astype(np.float32).astype(np.float64)
        ok, d, slack = quick(y, w, n)
        if not ok:
            continue
        assert n / w <= 24.0 and slack < 3e-5
        n_ok += 1
        ref = leastsq(lambda p: p[0] * np.abs(kern(XD - p[1], w, n)) - y, (y[3], 0.0), full_output=1)
        worst = max(worst, abs(d - ref[0][1]))
    assert n_ok > 1500
    assert worst < 2e-5, worst


def test_short_cut_refuses_points_next_to_a_null(quick):
    """A magnitude next to a null of the kernel gives |D| a kink with a local minimum on either side; the short cut must
    either land where lmdif lands or hand the block over."""
    n, w = 16384, 4914                       # nulls at +-3.334 bins: offsets near -+0.334 put a point on one
    rng = np.random.default_rng(3)
    for _ in range(1500):
        d_true = rng.choice([-1, 1]) * rng.uniform(0.30, 0.37)
        amp = rng.uniform(100, 1000)
        y = amp * np.abs(kern(XD - d_true, w, n)) + rng.normal(0, amp * 0.01, 7)
        y = np.abs(y).astype(np.float32).astype(np.float64)
        ok, d, _ = quick(y, w, n)
        if ok:
            ref = leastsq(lambda p: p[0] * np.abs(kern(XD - p[1], w, n)) - y, (y[3], 0.0), full_output=1)
            assert abs(d - ref[0][1]) < 2e-5, (y, d, ref[0])


def test_short_cut_refuses_a_start_on_a_null(quick):
    """Carrier as long as the block (or half / a third of it): at lmdif's starting guess the off-peak points sit ON nulls of
    the kernel, where the slope of |D| is one-sided.  The short cut must hand such blocks over (the reference's own
    freq_shift vectors, tests/test_carrier_sync.py:12-39, are of this kind) -- and wherever it does answer, agree."""
    rng = np.random.default_rng(0)
    n = 4096
    for ratio, may_answer in [(1, False), (1.5, False), (2, False), (3, False), (1.0001, False), (2.999, False), (4, True), (6, True)]:
        n_ok = 0
        for _ in range(200):
            w = int(round(n / ratio))
            y = 100 * np.abs(kern(XD - rng.uniform(-0.6, 0.6), w, n)) + rng.normal(0, 100 * rng.choice([1e-6, 1e-3, 1e-2]), 7)
            y = np.abs(y).astype(np.float32).astype(np.float64)
            ok, d, _ = quick(y, w, n)
            if ok:
                n_ok += 1
                ref = leastsq(lambda p: p[0] * np.abs(kern(XD - p[1], w, n)) - y, (y[3], 0.0), full_output=1)
                assert abs(d - ref[0][1]) < 2e-5, (ratio, y, d, ref[0])
        assert (n_ok > 100) == may_answer, (ratio, n_ok)
