"""Pin the oracle: the reference's own known-answer tests for the detect path (restated from
/root/reference/tests/*.py, cited per test) and the golden outputs of the real reference."""
import io

import numpy as np
import pytest
import scipy.signal

import parity_util as parity
from oracle import thrifty_oracle as orc


# ---- tests/test_block_data.py:13-37
def test_raw_to_complex():
    raw = np.array([0, 0, 127, 128, 255, 255], dtype=np.uint8)
    expected = np.array([-0.9953 - 0.9953j, -0.0031 + 0.0047j, 0.9969 + 0.9969j], dtype=np.complex64)
    np.testing.assert_allclose(orc.raw_to_complex(raw), expected, rtol=1e-2)


def test_complex_to_raw_and_inverse():
    cplx = np.array([-0.9953 - 0.9953j, -0.0031 + 0.0047j, 0.9969 + 0.9969j], dtype=np.complex64)
    np.testing.assert_array_equal(orc.complex_to_raw(cplx), [0, 0, 127, 128, 255, 255])
    every = np.arange(256, dtype=np.uint8)
    np.testing.assert_array_equal(orc.complex_to_raw(orc.raw_to_complex(every)), every)


# ---- tests/test_block_data.py:59-71
def test_card_reader():
    stream = io.StringIO("# Some comments\n# more comments\n1000.5425 10 r0+Om5==\n1000.5442 20 aaaaaa==")
    blocks = list(orc.card_reader(stream))
    assert [b[0] for b in blocks] == [1000.5425, 1000.5442]
    assert [b[1] for b in blocks] == [10, 20]
    assert [tuple(b[2]) for b in blocks] == [(175, 79, 142, 155), (105, 166, 154, 105)]


# ---- tests/test_block_data.py:40-56
def test_block_reader():
    stream = io.BytesIO(bytes(range(14)))
    blocks = list(orc.block_reader(stream, 3, 1))
    raw = [list(orc.complex_to_raw(d)) for _, d in blocks]
    assert raw == [[0x7f, 0x7f, 0, 1, 2, 3], [2, 3, 4, 5, 6, 7], [6, 7, 8, 9, 10, 11]]
    assert [i for i, _ in blocks] == [0, 1, 2]


# ---- tests/test_carrier_detect.py:11-22
@pytest.mark.parametrize("start,stop,length,expected", [
    (50, 100, 1024, (50, 100)), (0, -1, 1024, (0, 1023)),
    (-10, 10, 1024, (1014, 1034)), (-1, 0, 1024, (1023, 1024))])
def test_fft_range_index(start, stop, length, expected):
    assert orc.fft_range_index(start, stop, length) == expected


# ---- tests/test_carrier_detect.py:25-72
WINDOW_CASES = [
    (-81.0e3, -79.0e3, -80.0e3, True), (-81.0e3, -79.0e3, -79.1e3, True),
    (-81.0e3, -79.0e3, -80.9e3, True), (-81.0e3, -79.0e3, -82.0e3, False),
    (-81.0e3, -79.0e3, -78.0e3, False), (-81.0e3, -79.0e3, 0.0e3, False),
    (79.0e3, 81.0e3, 80.0e3, True), (79.0e3, 81.0e3, 79.1e3, True),
    (79.0e3, 81.0e3, 80.9e3, True), (79.0e3, 81.0e3, 82.0e3, False),
    (79.0e3, 81.0e3, 78.0e3, False), (79.0e3, 81.0e3, -80.0e3, False),
    (79.0e3, 81.0e3, 0.0e3, False),
    (-10.0e3, 5.0e3, 0.0e3, True), (-10.0e3, 5.0e3, -9.9e3, True), (-10.0e3, 5.0e3, 4.9e3, True),
    (-10.0e3, 5.0e3, 6.0e3, False), (-10.0e3, 5.0e3, -11.0e3, False)]


def window_case_block(carrier_freq, block_len=8192, carrier_len=2085, sample_rate=2.2e6):
    carrier = np.exp(2j * np.pi * carrier_freq * np.arange(carrier_len) / sample_rate)
    return np.concatenate([carrier, np.zeros(block_len - carrier_len)])


@pytest.mark.parametrize("freq_min,freq_max,carrier_freq,expected", WINDOW_CASES)
def test_detect_window(freq_min, freq_max, carrier_freq, expected):
    block_len, sample_rate = 8192, 2.2e6
    bin_freq = sample_rate / block_len
    window = (int(freq_min / bin_freq), int(freq_max / bin_freq))
    fft_mag = np.abs(np.fft.fft(window_case_block(carrier_freq)))
    detected, _, _, _ = orc.carrier_detect(fft_mag, (500.0 ** 2, 0.0, 0.0), window)
    assert detected == expected


# ---- tests/test_carrier_sync.py:12-39
@pytest.mark.parametrize("size,freq,shift", [(128, 0, 0), (128, -32, 32), (128, 32, 16),
                                             (128, -10.5, 0.5), (128, 8.3, -8.3)])
def test_freq_shift(size, freq, shift):
    signal = np.exp(2j * np.pi * np.arange(size) / size * freq)
    expected = np.fft.fft(np.exp(2j * np.pi * np.arange(size) / size * (freq + shift)))
    got = orc.freq_shift(signal, shift)
    np.testing.assert_allclose(np.abs(got), np.abs(expected), atol=1e-6, rtol=1e-6)


# ---- tests/test_carrier_sync.py:42-47
def test_dirichlet_kernel():
    expected = np.array([-0.1711, 0.0164, 0.3164, 0.6468, 0.9034, 1., 0.9034, 0.6468, 0.3164, 0.0164, -0.1711])
    np.testing.assert_allclose(orc.dirichlet_kernel(np.arange(-5, 6), 8192, 2015), expected, rtol=2e-3)


# ---- tests/test_carrier_sync.py:50-65
@pytest.mark.parametrize("offset", [-0.51, -0.5, -0.25, -0.1263, -0.1, 0., 0.001, 0.2, 0.4995, 0.56])
def test_dirichlet_interpolator(offset):
    peak_idx, block_len, carrier_len = 10, 8192, 2024
    freq = (1. * offset + peak_idx) * carrier_len / block_len
    carrier = np.exp(2j * np.pi * np.arange(carrier_len) / carrier_len * freq)
    signal_fft = np.abs(np.fft.fft(np.concatenate([carrier, np.zeros(block_len - carrier_len)])))
    got = orc.dirichlet_interpolate(signal_fft, peak_idx, block_len, carrier_len, width=6)
    np.testing.assert_allclose(got, offset, atol=1e-8, rtol=1e-8)


# ---- tests/test_soa_estimator.py:13-67
GOLD31 = np.array([1, 1, 1, 1, 1, -1, -1, -1, 1, 1, -1, 1, 1, 1, -1, 1,
                   -1, 1, -1, -1, -1, -1, 1, -1, -1, 1, -1, 1, 1, -1, -1])


def gen_block(pos, block_len=64):
    block = np.zeros(block_len)
    ook = (GOLD31 + 1) / 2
    end = min(block_len, pos + len(ook))
    block[pos:end] += ook[:end - pos]
    return block


@pytest.mark.parametrize("pos", [0, 1, 10, 33, 34, 63])
def test_despreader(pos):
    est = orc.SoaEstimator(GOLD31, (0., 0., 0.), 64, len(GOLD31))
    block = gen_block(pos)
    corr = est.despread(np.fft.fft(block))
    assert len(corr) == 64 - 31 + 1
    np.testing.assert_allclose(corr, scipy.signal.correlate(block, GOLD31, mode="valid"), atol=1e-12, rtol=1e-12)
    mag = np.abs(corr)
    if pos <= 33:
        peak = int(np.argmax(mag))
        assert peak == pos and mag[peak] >= 15.9
        np.testing.assert_array_less(np.delete(mag, peak), 5.1)
    else:
        np.testing.assert_array_less(mag, 5.1)


# ---- tests/test_soa_estimator.py:70-83
@pytest.mark.parametrize("params,expected", [((64, 31, 32), (0, 33)), ((64, 32, 32), (0, 32)),
                                             ((64, 33, 32), (1, 32)), ((64, 63, 32), (16, 17))])
def test_calculate_window(params, expected):
    assert orc.calculate_window(*params) == expected


# ---- golden outputs of the real reference (oracle/make_golden.py)
@pytest.mark.parametrize("name", parity.GOLDEN_NAMES)
def test_oracle_matches_reference_golden(name):
    cfg, raw, block_idx, ref, lines = parity.load_golden(name)
    limit = 16 if cfg["block_len"] >= 16384 else 24
    st = orc.DetectorSettings(cfg["block_len"], cfg["history_len"], len(cfg["template"]), cfg["cthresh"],
                              cfg["window"], cfg["template"], cfg["kthresh"])
    got = orc.detect_blocks(st, raw[:limit], block_idx[:limit])
    for f in ref.dtype.names:
        if f.endswith("margin"):
            continue
        if ref[f].dtype.kind == "f":
            np.testing.assert_allclose(got[f], ref[f][:limit], rtol=1e-10, atol=1e-10, equal_nan=True)
        else:
            np.testing.assert_array_equal(got[f], ref[f][:limit])


@pytest.mark.parametrize("name", ["n16384_gold11x4", "n32768_gold11x3"])
def test_oracle_matches_reference_multi_template_golden(name):
    """BASELINE config 5 (4 Gold-11 templates at N=16384) and 3 templates at N=32768 with a wide carrier window: golden =
    one independent reference Detector per template."""
    cfg, tpls, raw, block_idx, ref, _ = parity.load_multi_golden(name)
    assert ref.shape == (len(tpls), cfg["n_blocks"])
    if name == "n16384_gold11x4":
        assert len(tpls) == 4 and cfg["n_blocks"] > 2 * 148
    limit = 10 if cfg["block_len"] <= 16384 else 6
    for t in (0, len(tpls) - 1):
        st = orc.DetectorSettings(cfg["block_len"], cfg["history_len"], tpls.shape[1], cfg["cthresh"], cfg["window"],
                                  tpls[t], cfg["kthresh"])
        got = orc.detect_blocks(st, raw[:limit], block_idx[:limit])
        for f in ref.dtype.names:
            if ref[f].dtype.kind == "f":
                np.testing.assert_allclose(got[f], ref[t][f][:limit], rtol=1e-10, atol=1e-10, equal_nan=True)
            else:
                np.testing.assert_array_equal(got[f], ref[t][f][:limit])


def test_golden_arrays(golden_dir):
    import os
    from thrifty_b200 import synth
    g = np.load(os.path.join(golden_dir, "arrays_n4096_gold9.npz"))
    tpl = synth.gold_template(9)
    raw, _ = synth.make_blocks(1, 4096, len(tpl) + 6, tpl, 1.0, seed=int(g["seed"]))
    st = orc.DetectorSettings(4096, len(tpl) + 6, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.))
    res, sfft, corr = orc.Detector(st, 0).detect_raw(0.0, 5, raw[0], True)
    np.testing.assert_allclose(sfft, g["shifted_fft"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(corr, g["corr"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(res.soa, float(g["soa"]), rtol=0, atol=1e-9)
