"""The two halves of the reference's plug-in seam as drop-in classes on the GPU:

    thrifty_b200.carrier_sync.DefaultSynchronizer(thresh, window, block_len, carrier_len)(block) -> (shifted_fft, CarrierSyncInfo)
    thrifty_b200.soa_estimator.SoaEstimator(template, thresh, block_len, history_len)(fft)       -> (detected, CorrDetectionInfo, corr)

(thrifty/carrier_sync.py:82-118, thrifty/soa_estimator.py:42-92; C entry points thr_sync_batch / thr_soa_batch).
The reference's own vectors for these classes (tests/test_carrier_sync.py:12-65, tests/test_soa_estimator.py:13-109) are
run through them, scaled to block lengths the kernel supports (N >= 1024) with float32 tolerances stated per test, and
both halves are checked block by block against the oracle's restatement of the same classes on the golden inputs."""
import os

import numpy as np
import pytest
import scipy.signal

import parity_util as parity
from oracle import thrifty_oracle as orc
from thrifty_b200 import synth
from thrifty_b200.carrier_sync import DefaultSynchronizer
from thrifty_b200.soa_estimator import SoaEstimator, calculate_window

pytestmark = pytest.mark.gpu

ALWAYS = (0., 0., 0.)


# ---------------------------------------------------------------- tests/test_carrier_sync.py:12-39
@pytest.mark.parametrize("freq,shift", [(0, 0), (-32, 32), (32, 16), (-10.5, 0.5), (8.3, -8.3)])
def test_synchronizer_freq_shift_tones(freq, shift):
    """freq_shift on pure tones: the synchronizer shifts by -(bin + offset) of the tone it finds, so |shifted_fft| must be
    the spectrum of the tone moved to bin 0 (reference tolerance 1e-6 in float64; here 1e-5 of the peak in float32).  A
    second, stronger tone on an integer bin outside the window keeps the reference's noise estimate positive."""
    n = 4096
    f0 = (freq + shift) % n
    t = np.arange(n) / n
    x = (0.4 * np.exp(2j * np.pi * t * f0) + 0.5 * np.exp(2j * np.pi * t * 2000)).astype(np.complex64)
    sync = DefaultSynchronizer(ALWAYS, (-800, 300), n, n)
    sfft, info = sync(x)
    sync.close()
    assert sfft is not None and sfft.shape == (n,)
    pos = info.bin + info.offset
    assert abs((pos - f0 + n / 2) % n - n / 2) <= 1e-4
    expected = np.abs(np.fft.fft(x.astype(np.complex128) * np.exp(-2j * np.pi * pos * np.arange(n) / n)))
    np.testing.assert_allclose(np.abs(sfft), expected, atol=1e-5 * expected.max())


# ---------------------------------------------------------------- tests/test_carrier_sync.py:42-65
@pytest.mark.parametrize("offset", [-0.51, -0.5, -0.25, -0.1263, -0.1, 0., 0.001, 0.2, 0.4995, 0.56])
def test_synchronizer_dirichlet_offsets(offset):
    """make_dirichlet_interpolator's vectors (peak_idx 10, block_len 8192, carrier_len 2024) through the class: the
    reference gets 1e-8 from float64 magnitudes; from a float32 spectrum 5e-5."""
    peak_idx, block_len, carrier_len = 10, 8192, 2024
    freq = (1. * offset + peak_idx) * carrier_len / block_len
    carrier = 0.9 * np.exp(2j * np.pi * np.arange(carrier_len) / carrier_len * freq)
    block = np.concatenate([carrier, np.zeros(block_len - carrier_len)])
    sync = DefaultSynchronizer(ALWAYS, (7, 110), block_len, carrier_len)
    sfft, info = sync(block)
    sync.close()
    assert sfft is not None
    assert abs(info.bin + info.offset - (peak_idx + offset)) <= 5e-5
    # ... and against the reference's own fit of the same (float32) magnitudes
    mag = np.abs(np.fft.fft(block.astype(np.complex64)))
    want = orc.dirichlet_interpolate(mag, info.bin, block_len, carrier_len)
    assert abs(info.offset - want) <= 2e-5


def test_synchronizer_no_carrier_returns_none():
    n = 4096
    rng = np.random.default_rng(5)
    x = (0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    sync = DefaultSynchronizer((0., 15., 0.), (7, 110), n, 1226)
    sfft, info = sync(x)
    mag = np.abs(np.fft.fft(x))
    det, k, pk, nz = orc.carrier_detect(mag, (0., 15., 0.), (7, 110))
    assert not det and sfft is None
    assert info.bin == k and info.offset == 0
    assert abs(info.energy - pk) <= 1e-5 * pk and abs(info.noise - nz) <= 1e-5 * nz
    with pytest.raises(ValueError):                      # carrier_detect.py:47-49
        DefaultSynchronizer(ALWAYS, (-5000, 10), n, 1226)
    sync.close()


@pytest.mark.parametrize("name", ["n4096_gold9", "n16384_example", "n8192_gold10", "n32768_example"])
def test_synchronizer_matches_reference_per_block(name):
    """Every block of a golden configuration: CarrierSyncInfo against the reference's records, the shifted spectrum against
    the oracle's freq_shift (float64) -- error relative to the largest bin, float32 transforms."""
    cfg, raw, block_idx, ref, _ = parity.load_golden(name)
    n = cfg["block_len"]
    nb = min(len(raw), 24)
    sync = DefaultSynchronizer(cfg["cthresh"], cfg["window"], n, len(cfg["template"]), batch=32)
    out = sync.sync_many(list(raw[:nb]))                 # raw uint8 blocks: conversion on the GPU
    out_c = sync.sync_many([orc.raw_to_complex(r) for r in raw[:4]])
    sync.close()
    odet = orc.Detector(orc.DetectorSettings(n, cfg["history_len"], len(cfg["template"]), cfg["cthresh"], cfg["window"],
                                             cfg["template"], cfg["kthresh"]))
    n_car = 0
    for i in range(nb):
        sfft, info = out[i]
        r = ref[i]
        assert (sfft is not None) == bool(r["carrier_detected"])
        assert info.bin == r["carrier_bin"]
        np.testing.assert_allclose(info.energy, r["carrier_energy"], rtol=1e-4)
        np.testing.assert_allclose(info.noise, r["carrier_noise"], rtol=1e-4)
        if sfft is None:
            assert info.offset == 0
            continue
        n_car += 1
        assert abs(info.offset - r["carrier_offset"]) <= 1e-4
        want, _ = odet.sync(orc.raw_to_complex(raw[i]))
        err = np.abs(sfft - want).max() / np.abs(want).max()
        assert err <= 2e-5, (i, err)
        if i < 4:
            np.testing.assert_array_equal(out_c[i][0], sfft)         # complex64 input path: same kernel, same bits
    assert n_car >= 5


# ---------------------------------------------------------------- tests/test_soa_estimator.py:13-67
@pytest.mark.parametrize("pos", [0, 1, 10, 2870, 2871, 4095])
def test_estimator_despreader_vs_scipy_correlate(pos):
    """despread(fft) == scipy.signal.correlate(block, template, 'valid'); the peak is the burst position.  Scaled from
    N=64 / 31 chips to N=4096 / 1226 samples; reference tolerance 1e-12 (float64), here 2e-5 of the peak (float32)."""
    n = 4096
    tpl = synth.gold_template(9)
    L = len(tpl)
    block = np.zeros(n)
    end = min(n, pos + L)
    block[pos:end] += ((tpl + 1) / 2)[:end - pos]
    est = SoaEstimator(tpl, None, n, L)                  # history = len(template), as the reference's test
    corr = est.despread(np.fft.fft(block))
    est.close()
    assert len(corr) == n - L + 1
    want = scipy.signal.correlate(block, tpl, mode="valid")
    assert np.abs(corr - want).max() <= 2e-5 * L / 2
    if pos <= n - L:
        assert int(np.argmax(np.abs(corr))) == pos
        assert np.abs(corr[pos]) >= 0.99 * L / 2


# ---------------------------------------------------------------- tests/test_soa_estimator.py:70-109
@pytest.mark.parametrize("params,expected", [((64, 31, 32), (0, 33)), ((64, 32, 32), (0, 32)), ((64, 33, 32), (1, 32)),
                                             ((64, 63, 32), (16, 17))])
def test_calculate_window(params, expected):
    assert calculate_window(*params) == expected


@pytest.mark.parametrize("hist,burst_at", [(1232, 1), (1232, 2868), (2000, 100), (2000, 2700), (2000, 1200)])
def test_estimator_peak_window_is_half_open(hist, burst_at):
    n = 4096
    tpl = synth.gold_template(9)
    L = len(tpl)
    start, stop = calculate_window(n, hist, L)
    block = np.zeros(n)
    block[burst_at:burst_at + L] += 0.4 * (tpl + 1) / 2
    est = SoaEstimator(tpl, ALWAYS, n, hist)
    detected, info, corr = est(np.fft.fft(block))
    est.close()
    mag = np.abs(corr.astype(np.complex128))
    assert detected and int(np.argmax(mag)) == burst_at and start <= info.sample < stop
    if start <= burst_at < stop:
        assert info.sample == burst_at
    else:
        assert info.sample == start + int(np.argmax(mag[start:stop]))


@pytest.mark.parametrize("name", ["n4096_gold9", "n4096_gold9_wrapwin_std", "n16384_example", "n32768_example"])
def test_estimator_matches_reference_per_block(name):
    """The reference's float64 shifted spectra (oracle.sync) fed to the GPU estimator: every field of CorrDetectionInfo
    against the oracle's SoaEstimator on the same input, and the correlation itself."""
    cfg, raw, block_idx, ref, _ = parity.load_golden(name)
    n, h, tpl = cfg["block_len"], cfg["history_len"], cfg["template"]
    odet = orc.Detector(orc.DetectorSettings(n, h, len(tpl), cfg["cthresh"], cfg["window"], tpl, cfg["kthresh"]))
    ffts, rows = [], []
    for i in range(min(len(raw), 24)):
        sfft, _ = odet.sync(orc.raw_to_complex(raw[i]))
        if sfft is not None:
            ffts.append(sfft)
            rows.append(i)
    assert len(ffts) >= 5
    est = SoaEstimator(tpl, cfg["kthresh"], n, h, batch=32)
    got = est.estimate_many(np.stack(ffts))
    est.close()
    for (detected, info, corr), fft, i in zip(got, ffts, rows):
        with np.errstate(all="ignore"):
            w_det, w_info, w_corr = odet.soa_estimate(fft)
        r = ref[i]
        assert info.sample == w_info.sample == r["corr_sample"]
        marginal = np.isfinite(r["corr_margin"]) and abs(r["corr_margin"] - 1) < 1e-3
        assert detected == w_det or marginal
        np.testing.assert_allclose(info.energy, w_info.energy, rtol=1e-4)
        if np.isnan(w_info.noise):
            assert np.isnan(info.noise)
        else:
            np.testing.assert_allclose(info.noise, w_info.noise, rtol=1e-4)
        if detected and w_det:
            assert abs(info.offset - w_info.offset) <= 1e-4
        assert np.abs(corr - w_corr).max() <= 2e-5 * np.abs(w_corr).max()


def test_chip_rate_search_style_chain():
    """scripts/chip_rate_search.py:44-55,121-127: synchronize once, then match the shifted spectrum against several
    candidate templates; the chain sync -> estimate must equal the fused Detector on the same block."""
    from thrifty_b200.detect import Detector, DetectorSettings
    n = 4096
    tpls = [synth.gold_template(9, i) for i in range(3)]
    L = len(tpls[0])
    raw, _ = synth.make_blocks(6, n, L + 6, tpls[1], 1.0, seed=4321)
    block = orc.raw_to_complex(raw[2])
    sync = DefaultSynchronizer(thresh_coeffs=(100, 0, 0), window=None, block_len=len(block), carrier_len=L)
    shifted_fft, cinfo = sync(block)
    sync.close()
    assert shifted_fft is not None
    energies = []
    for t in tpls:
        est = SoaEstimator(template=t, thresh_coeffs=(0, 0, 0), block_len=len(shifted_fft), history_len=len(t) - 1)
        detected, corr_info, _ = est(shifted_fft)
        est.close()
        assert detected
        energies.append(corr_info.energy)
    assert int(np.argmax(energies)) == 1 and energies[1] > 2 * max(energies[0], energies[2])
    # fused detector, same settings as the chain with template 1
    st = DetectorSettings(n, L - 1, L, (100., 0., 0.), (0, -1), tpls[1], (0., 0., 0.))
    det = Detector(st, rxid=0)
    d, res = det.detect(0.0, 0, block)
    det.close()
    est = SoaEstimator(tpls[1], (0, 0, 0), n, L - 1)
    _, ci, _ = est(shifted_fft)
    est.close()
    assert d and res.corr_info.sample == ci.sample
    assert res.carrier_info.bin == cinfo.bin and abs(res.carrier_info.offset - cinfo.offset) < 1e-6
    np.testing.assert_allclose(res.corr_info.energy, ci.energy, rtol=1e-5)
    assert abs(res.corr_info.offset - ci.offset) <= 1e-5
