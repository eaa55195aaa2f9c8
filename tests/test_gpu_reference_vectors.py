"""The reference's own known-answer tests for the hot path, run through the CUDA path (C ABI).

The reference pins its functions one by one; the fused kernel has no such seams, so every
vector is fed to a whole detector and the pinned quantity is read from the record / the debug outputs:

  tests/test_carrier_detect.py:25-72   19-row window table (negative and zero-straddling windows)
  tests/test_carrier_sync.py:12-39     freq_shift on pure tones incl. fractional shifts (|FFT#2|)
  tests/test_carrier_sync.py:50-65     Dirichlet interpolator recovers sub-bin offsets
  tests/test_soa_estimator.py:13-67    despreader == scipy.signal.correlate(..., 'valid'), peak index == position
  tests/test_soa_estimator.py:86-109   get_peak honours the half-open window

Sizes are scaled to what the kernel supports (N >= 1024); tolerances are float32 ones and are stated per test
(the reference's 1e-8 ... 1e-12 are float64 tolerances of its float64 path).
"""
import numpy as np
import pytest
import scipy.signal

from test_oracle import WINDOW_CASES, window_case_block
from thrifty_b200 import synth
from thrifty_b200._native import NativeDetector

pytestmark = pytest.mark.gpu

NEVER = (1e30, 0., 0.)        # a threshold nothing exceeds
ALWAYS = (0., 0., 0.)         # ... and one everything exceeds


@pytest.mark.parametrize("as_raw", [False, True])
def test_reference_window_table(as_raw):
    """carrier_detect.detect: detected iff the carrier lies inside [bin_min, bin_max] (closed, signed bins)."""
    block_len, sample_rate = 8192, 2.2e6
    bin_freq = sample_rate / block_len
    tpl = synth.gold_template(10)
    by_window = {}
    for fmin, fmax, fc, expected in WINDOW_CASES:
        by_window.setdefault((int(fmin / bin_freq), int(fmax / bin_freq)), []).append((fc, expected))
    assert len(by_window) == 3
    for window, cases in by_window.items():
        amp = 0.5 if as_raw else 1.0               # |X| peak = amp * 2085, threshold sqrt(500^2) = 500 (amp = 1) or 250
        thresh = ((500.0 * amp) ** 2, 0., 0.)
        det = NativeDetector(block_len, len(tpl) + 6, tpl, 2085, window, thresh, NEVER, max_batch=8)
        blocks = np.stack([amp * window_case_block(fc) for fc, _ in cases])
        if as_raw:
            rec = det.detect_raw(np.stack([synth.complex_to_raw(b) for b in blocks]))[:, 0]
        else:
            rec = det.detect_c64(blocks.astype(np.complex64))[:, 0]
        got = (rec["flags"] & 1) != 0
        assert list(got) == [e for _, e in cases], (window, cases, rec["carrier_bin"])
        for (fc, e), r in zip(cases, rec):
            if e:                                  # the reported bin is the signed-window bin wrapped to [0, N)
                d = (r["carrier_bin"] - fc / bin_freq + block_len / 2) % block_len - block_len / 2
                assert abs(d) <= 0.6, (fc, r["carrier_bin"])
        det.close()


@pytest.mark.parametrize("offset", [-0.51, -0.5, -0.25, -0.1263, -0.1, 0., 0.001, 0.2, 0.4995, 0.56])
@pytest.mark.parametrize("window", [(7, 110), (7, 300)])          # pruned and full FFT#1
def test_reference_dirichlet_offsets(offset, window):
    """make_dirichlet_interpolator: bin + offset == true carrier position.  The reference fixes peak_idx = 10 and gets
    1e-8 in float64; here the arg-max picks the peak (bin 9 / 11 for |offset| > 0.5) and the spectrum is float32."""
    peak_idx, block_len, carrier_len = 10, 8192, 2024
    tpl = synth.gold_template(10)
    freq = (1. * offset + peak_idx) * carrier_len / block_len
    carrier = 0.9 * np.exp(2j * np.pi * np.arange(carrier_len) / carrier_len * freq)
    block = np.concatenate([carrier, np.zeros(block_len - carrier_len)]).astype(np.complex64)
    det = NativeDetector(block_len, len(tpl) + 6, tpl, carrier_len, window, ALWAYS, NEVER, max_batch=2)
    rec = det.detect_c64(block[None])[0, 0]
    det.close()
    assert rec["flags"] & 1
    assert abs(rec["carrier_bin"] - (peak_idx + offset)) <= 0.5 + 1e-3
    assert abs(rec["carrier_bin"] + rec["carrier_offset"] - (peak_idx + offset)) <= 5e-5


@pytest.mark.parametrize("freq,shift", [(0, 0), (-32, 32), (32, 16), (-10.5, 0.5), (8.3, -8.3), (57.25, 0), (100.5, 0),
                                        (-700.9, 0)])
def test_reference_freq_shift_tones(freq, shift):
    """carrier_sync.freq_shift: |FFT(x * mix)| equals the spectrum of the ideally shifted tone.  Through the detector the
    shift is -(bin + offset) of the tone itself, so FFT#2 must put the tone into bin 0.  A lone tone is *not* detected by the
    reference (next test), so a stronger second tone on an integer bin outside the carrier window (no leakage) keeps the
    noise estimate positive without disturbing the seven magnitudes of the fit."""
    n = 4096
    tpl = synth.gold_template(9)
    f0 = (freq + shift) % n
    t = np.arange(n) / n
    x = (0.4 * np.exp(2j * np.pi * t * f0) + 0.5 * np.exp(2j * np.pi * t * 2000)).astype(np.complex64)
    det = NativeDetector(n, len(tpl) + 6, tpl, n, (-800, 300), ALWAYS, NEVER, max_batch=2)
    rec, sfft, _, _ = det.detect_block_data(iq=x)
    det.close()
    r = rec[0]
    assert r["flags"] & 1
    pos = (r["carrier_bin"] + float(r["carrier_offset"]))
    assert abs((pos - f0 + n / 2) % n - n / 2) <= 1e-4
    expected = np.abs(np.fft.fft(x.astype(np.complex128) * np.exp(-2j * np.pi * pos * np.arange(n) / n)))
    np.testing.assert_allclose(np.abs(sfft), expected, atol=1e-5 * expected.max())
    assert np.abs(sfft[0]) >= 0.999 * 0.4 * n and np.abs(sfft[1:1000]).max() < 1e-3 * n


@pytest.mark.parametrize("f0", [48, 57.25, 100.5])
@pytest.mark.parametrize("cthresh,window", [((0., 0., 0.), (0, -1)), ((0., 15., 0.), (7, 110))])
def test_lone_tone_noise_estimate_quirk(f0, cthresh, window):
    """carrier_detect.py:116-126: noise^2 = (sum mag^2 - 2 peak^2)/(N-1) is negative when one bin holds more than half of
    the energy; the reference's sqrt gives NaN, the threshold becomes NaN and `peak > NaN` is False -- a lone tone on (or
    near) a bin centre is not a carrier.  The kernel must reproduce that, not clamp (fastdet semantics would)."""
    from oracle import thrifty_oracle as orc
    n = 4096
    tpl = synth.gold_template(9)
    x = (0.8 * np.exp(2j * np.pi * np.arange(n) / n * f0)).astype(np.complex64)
    with np.errstate(invalid="ignore"):
        want, want_bin, want_peak, _ = orc.carrier_detect(np.abs(np.fft.fft(x)), cthresh, window)
    assert want == (f0 == 100.5)                     # energy split over two bins: the estimate stays positive
    det = NativeDetector(n, len(tpl) + 6, tpl, n, window, cthresh, NEVER, max_batch=2)
    r = det.detect_c64(x[None])[0, 0]
    det.close()
    assert bool(r["flags"] & 1) == bool(want) and r["carrier_bin"] == want_bin
    assert abs(r["carrier_energy"] - want_peak) <= 1e-5 * want_peak
    if not want:
        assert np.isnan(r["carrier_noise"]) and r["carrier_offset"] == 0


@pytest.mark.parametrize("pos", [0, 1, 10, 1500, 2869, 2870, 2871, 3500, 4095])
def test_reference_despreader_vs_scipy_correlate(pos):
    """SoaEstimator.despread == scipy.signal.correlate(block, template, 'valid') and the peak sits at the burst position
    (tests/test_soa_estimator.py:13-67, scaled from N=64 / 31 chips to N=4096 / 1226 samples).  The block carries a
    DC 'carrier' (OOK chips in {0, 1}), so the carrier bin is 0 and the mix is the identity up to its constant phase."""
    n = 4096
    tpl = synth.gold_template(9)                     # +-1 chips, L = 1226
    L = len(tpl)
    hist = L - 1                                     # smallest legal history: the peak window is all of corr
    corr_len = n - L + 1
    block = np.zeros(n)
    ook = 0.4 * (tpl + 1) / 2
    end = min(n, pos + L)
    block[pos:end] += ook[:end - pos]
    det = NativeDetector(n, hist, tpl, L, (-3, 3), ALWAYS, ALWAYS, max_batch=2)
    rec, sfft, corr, _ = det.detect_block_data(iq=block.astype(np.complex64))
    det.close()
    r = rec[0]
    assert len(corr) == corr_len
    if pos < corr_len:       # (the code's own spectrum perturbs the 7-point Dirichlet fit of the DC line a little)
        assert r["carrier_bin"] == 0 and abs(r["carrier_offset"]) < 0.1
    mixed = block * np.exp(-2j * np.pi * (r["carrier_bin"] + float(r["carrier_offset"])) * (np.arange(n) / n - 0.5))
    expected = scipy.signal.correlate(mixed, tpl, mode="valid")
    scale = 0.4 * L / 2
    assert np.abs(corr - expected).max() <= 2e-5 * scale            # float32 transforms; reference: 1e-12 in float64
    mag = np.abs(expected)
    if pos < corr_len:                               # whole burst inside the block: the peak is its position
        assert r["corr_sample"] == pos == int(np.argmax(mag))
        assert abs(r["corr_energy"] - mag[pos]) <= 1e-4 * mag[pos] and mag[pos] >= 0.99 * scale
        side = mag.copy()                            # 2.4 samples per chip: the neighbours belong to the peak
        side[max(0, pos - 3):pos + 4] = 0
        assert side.max() < 0.2 * scale
    else:                                            # truncated burst: only sidelobes, still a maximum of |corr|
        assert mag.max() < 0.8 * scale
        assert mag[r["corr_sample"]] >= mag.max() * (1 - 1e-4)


@pytest.mark.parametrize("hist,burst_at", [(1232, 1), (1232, 2868), (2000, 100), (2000, 2700), (2000, 1200)])
def test_reference_get_peak_window(hist, burst_at):
    """get_peak / calculate_window: the correlation window [p//2, corr_len - (p - p//2)), p = H - L + 1, is half-open;
    a burst outside it is not returned -- the arg-max of the window is (tests/test_soa_estimator.py:70-109)."""
    n = 4096
    tpl = synth.gold_template(9)
    L = len(tpl)
    corr_len = n - L + 1
    p = hist - L + 1
    start, stop = p // 2, corr_len - (p - p // 2)
    block = np.zeros(n)
    block[burst_at:burst_at + L] += 0.4 * (tpl + 1) / 2
    det = NativeDetector(n, hist, tpl, L, (-3, 3), ALWAYS, ALWAYS, max_batch=2)
    rec, _, corr, _ = det.detect_block_data(iq=block.astype(np.complex64))
    det.close()
    r = rec[0]
    mag = np.abs(corr.astype(np.complex128))
    assert int(np.argmax(mag)) == burst_at           # the burst is where it was put ...
    assert start <= r["corr_sample"] < stop          # ... but only reported when it lies inside the window
    if start <= burst_at < stop:
        assert r["corr_sample"] == burst_at
    else:
        assert r["corr_sample"] != burst_at
        assert r["corr_sample"] == start + int(np.argmax(mag[start:stop]))
