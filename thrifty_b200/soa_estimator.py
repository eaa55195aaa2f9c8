"""Estimate sample-of-arrival by correlating against a template -- B200 drop-in for thrifty.soa_estimator's
SoaEstimator (thrifty/soa_estimator.py:42-134).

    soa_estimate = SoaEstimator(template, thresh_coeffs, block_len, history_len)
    detected, corr_info, corr = soa_estimate(fft)       # soa_estimator.py:78-92

`fft` is a (carrier-synchronized) spectrum of block_len bins, e.g. the first output of
thrifty_b200.carrier_sync.DefaultSynchronizer; `corr_info` is a toads_data.CorrDetectionInfo(sample, offset, energy,
noise); `corr` the block_len - len(template) + 1 correlation values (complex64).  Multiply by conj(FFT(template)), IFFT,
windowed first-maximum peak, noise estimate, threshold and Gaussian interpolation run in the fused CUDA kernel, entered
at the stage boundary (thr_soa_batch); `estimate_many` does a whole batch in one launch.  thresh_coeffs=None (as the
reference's despreader tests pass) means (0, 0, 0).  There is no CPU fallback.
"""

from __future__ import print_function

import numpy as np

from thrifty_b200 import toads_data
from thrifty_b200._native import FLAG_CORR, NativeDetector


def calculate_window(block_len, history_len, template_len):
    """Half-open window of unique correlation lags (soa_estimator.py:20-39)."""
    corr_len = block_len - template_len + 1
    assert history_len >= template_len - 1
    assert history_len < block_len
    padding = history_len - template_len + 1
    left_pad = padding // 2
    right_pad = padding - left_pad
    return left_pad, corr_len - right_pad


class SoaEstimator(object):
    """Despreader (correlate using FFT), threshold detector and Gaussian interpolator (soa_estimator.py:42-62)."""

    def __init__(self, template, thresh_coeffs, block_len, history_len, device=0, batch=256):
        template = np.asarray(template, dtype=np.float64)
        self.template = template
        self.template_energy = float(np.sum(template ** 2))
        self.block_len = int(block_len)
        self.history_len = int(history_len)
        self.corr_len = self.block_len - len(template) + 1
        self.window = calculate_window(self.block_len, self.history_len, len(template))
        self.thresh_coeffs = (0., 0., 0.) if thresh_coeffs is None else thresh_coeffs
        self.native = NativeDetector(self.block_len, self.history_len, template, len(template), None, (0., 0., 0.),
                                     self.thresh_coeffs, device=device, max_batch=max(1, int(batch)))

    def estimate_many(self, ffts, want_corr=True):
        """Array [B, block_len] of spectra -> list of (detected, CorrDetectionInfo, corr), one launch."""
        ffts = np.asarray(ffts)
        if ffts.ndim == 1:
            ffts = ffts[None, :]
        assert ffts.shape[1] == self.block_len
        recs, corr = self.native.soa_batch(ffts.astype(np.complex64), want_corr=want_corr)
        out = []
        for i in range(len(recs)):
            detected = bool(recs["flags"][i] & FLAG_CORR)
            info = toads_data.CorrDetectionInfo(int(recs["corr_sample"][i]), float(recs["corr_offset"][i]) if detected else 0,
                                                float(recs["corr_energy"][i]), float(recs["corr_noise"][i]))
            out.append((detected, info, corr[i] if want_corr else None))
        return out

    def soa_estimate(self, fft):
        """Estimate the SoA of the given signal (soa_estimator.py:78-92)."""
        return self.estimate_many(np.asarray(fft)[None, :])[0]

    def __call__(self, fft):
        return self.soa_estimate(fft)

    def despread(self, fft):
        """Correlate / despread using FFT (soa_estimator.py:97-102)."""
        return self.soa_estimate(fft)[2]

    def close(self):
        self.native.close()
