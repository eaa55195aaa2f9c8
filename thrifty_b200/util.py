"""Small helpers used by the summary line (thrifty/util.py:6-22)."""
import numpy as np


def snr(peak_ampl, noise_rms):
    """SNR in dB from an amplitude and a noise RMS (thrifty/util.py:6-8)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return 20 * np.log10(np.divide(peak_ampl, noise_rms))


def fft_bin(idx, fft_len):
    """FFT array index -> signed frequency bin, == np.fft.fftfreq(n, 1/n)[idx] (thrifty/util.py:11-22)."""
    if idx < 0 or idx <= (2 * fft_len - 1) / 4:
        return idx
    return idx - fft_len
