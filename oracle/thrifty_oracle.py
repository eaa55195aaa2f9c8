"""CPU oracle for the Thrifty `detect` hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the reference's per-block detect
chain.  It is the checker the CUDA path is compared against; it is *not* part
of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
``thrifty_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

Parity status: PINNED.  ``oracle/make_golden.py`` runs the *real* reference
(imported from /root/reference in the build container) on seeded synthetic
blocks and (a) asserts this restatement reproduces every field to 1e-9 and
(b) stores the reference's outputs under ``tests/golden/``.  The reference's own
known-answer tests for this path are re-stated in ``tests/test_oracle.py``.

Numeric types follow what the reference does under numpy >= 2:
rawconv, FFT #1, |X| and the carrier decision in float32; Dirichlet fit, mix,
FFT #2, correlation, IFFT and the correlation decision in float64.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""

from __future__ import annotations

import base64
from collections import namedtuple

import numpy as np
from scipy.optimize import curve_fit

# --------------------------------------------------------------------------
# record types (thrifty/toads_data.py:8-19, thrifty/detect.py:24-31)
# --------------------------------------------------------------------------
CarrierSyncInfo = namedtuple("CarrierSyncInfo", "bin offset energy noise")
CorrDetectionInfo = namedtuple("CorrDetectionInfo", "sample offset energy noise")
DetectorSettings = namedtuple(
    "DetectorSettings",
    "block_len history_len carrier_len carrier_thresh carrier_window "
    "template corr_thresh")
OracleResult = namedtuple(
    "OracleResult", "detected timestamp block soa carrier_info corr_info rxid")


# --------------------------------------------------------------------------
# block I/O  (thrifty/block_data.py)
# --------------------------------------------------------------------------
def raw_to_complex(data):
    """uint8 I/Q pairs -> complex64, (b - 127.4) / 128 per component.

    thrifty/block_data.py:38-52 (native twin: fastcard/rawconv.c:5-28)."""
    values = np.asarray(data, dtype=np.uint8).astype(np.float32).view(np.complex64)
    values = values - np.complex64(127.4 + 127.4j)
    values = values / np.float32(128)
    return values.astype(np.complex64)


def complex_to_raw(array):
    """Inverse of raw_to_complex: uint8(x*128 + 127.4), truncating.

    thrifty/block_data.py:55-67."""
    scaled = np.asarray(array).astype(np.complex64).view(np.float32) * 128 + 127.4
    return scaled.astype(np.uint8)


def card_reader(stream):
    """Yield (timestamp, block_idx, raw uint8[2N]) for each .card data line.

    thrifty/block_data.py:101-131, restated for py3 text or binary streams.
    Skips '#' comments, blank lines and the 'Using Volk machine:' / 'linux;'
    noise lines.  Yields the raw bytes (callers apply raw_to_complex)."""
    for line in stream:
        if isinstance(line, bytes):
            line = line.decode("ascii")
        if len(line) == 0:
            break
        if line[0] == "#" or line[0] == "\n":
            continue
        if line.startswith("Using Volk machine:") or line.startswith("linux;"):
            continue
        timestamp, idx, encoded = line.rstrip("\n").split(" ")
        raw = np.frombuffer(base64.b64decode(encoded), dtype=np.uint8)
        yield float(timestamp), int(idx), raw


def block_reader(stream, size, history):
    """Raw uint8 I/Q stream -> overlapping complex blocks.

    thrifty/block_data.py:70-98: each block has `size` samples of which the
    first `history` repeat the end of the previous block; the first block's
    history is complex zeros; a trailing partial block is dropped."""
    new = size - history
    data = np.zeros(size, dtype=np.complex64)
    block_idx = 0
    while True:
        chunk = stream.read(new * 2)
        if len(chunk) < new * 2:
            break
        new_data = raw_to_complex(np.frombuffer(chunk, dtype=np.uint8))
        data = np.concatenate([data[-history:], new_data]) if history else new_data
        yield block_idx, data
        block_idx += 1


# --------------------------------------------------------------------------
# carrier detection (thrifty/carrier_detect.py)
# --------------------------------------------------------------------------
def fft_range_index(start, stop, length):
    """Closed signed-bin interval -> FFT index interval (stop may be >= length).

    thrifty/carrier_detect.py:17-58."""
    if abs(start) >= length or abs(stop) >= length:
        raise ValueError("Frequency window out of range: {} - {}".format(start, stop))
    if start < 0 and stop >= 0:
        start, stop = length + start, length + stop
    if start < 0:
        start = length + start
    if stop < 0:
        stop = length + stop
    if stop < start:
        start, stop = stop, start
    return start, stop


def carrier_detect(fft_mag, thresh_coeffs, window=None):
    """Windowed spectral peak + threshold test (no peak filter).

    thrifty/carrier_detect.py:61-96 with helpers :99-154.  Returns
    (detected, peak_idx, peak_mag, noise_rms); arithmetic stays in fft_mag's
    dtype (float32 on the reference path)."""
    n = len(fft_mag)
    start, stop = (0, -1) if window is None else window
    start_idx, stop_idx = fft_range_index(start, stop, n)
    sel = np.take(fft_mag, range(start_idx, stop_idx + 1), mode="wrap")
    max_idx = int(np.argmax(sel))
    peak_mag = sel[max_idx]
    peak_idx = max_idx + start_idx
    if peak_idx > n:                       # sic: '>' (carrier_detect.py:151)
        peak_idx -= n
    fft_energy = np.sum(fft_mag ** 2)
    noise_power = (fft_energy - 2 * peak_mag ** 2) / (n - 1)
    noise_rms = np.sqrt(noise_power)
    t_const, t_snr, t_std = thresh_coeffs
    stddev = np.std(fft_mag) if t_std else 0
    thresh = np.sqrt(t_const + t_snr * noise_rms ** 2 + t_std * stddev ** 2)
    detected = bool(peak_mag > thresh)
    return detected, peak_idx, peak_mag, noise_rms


# --------------------------------------------------------------------------
# carrier sync (thrifty/carrier_sync.py)
# --------------------------------------------------------------------------
def dirichlet_kernel(xdata, block_len, carrier_len):
    """sin(pi W x / N) / (W sin(pi x / N)), 1 at x == 0.

    thrifty/carrier_sync.py:121-132."""
    n, w = block_len, carrier_len
    xdata = np.array(xdata, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        weights = np.sin(np.pi * w * xdata / n) / np.sin(np.pi * xdata / n) / w
        weights[np.isnan(weights)] = 1
    return weights


def dirichlet_interpolate(fft_mag, peak_idx, block_len, carrier_len, width=6):
    """Sub-bin carrier offset: LM fit of A*|dirichlet(x - delta)| to 7 bins.

    thrifty/carrier_sync.py:150-196 (scipy curve_fit, default 'lm', p0 =
    (mag[peak], 0))."""
    def _model(xdata, amplitude, offset):
        xdata = np.array(xdata, dtype=np.float64)
        return amplitude * np.abs(dirichlet_kernel(xdata - offset, block_len, carrier_len))

    xdata = np.arange(-(width // 2), width // 2 + 1)
    ydata = fft_mag[peak_idx + xdata]
    popt, _ = curve_fit(_model, xdata, ydata, p0=(fft_mag[peak_idx], 0))
    return popt[1]


def freq_shift(signal, shift):
    """Time-domain fractional frequency shift followed by an FFT.

    thrifty/carrier_sync.py:222-238: x * exp(2j pi shift (n/N - 0.5)); the
    product promotes complex64 to complex128, so the FFT runs in double."""
    n = len(signal)
    freqs = np.arange(n) * 1.0 / n - 0.5
    shift_signal = np.exp(2j * np.pi * shift * freqs)
    return np.fft.fft(signal * shift_signal)


# --------------------------------------------------------------------------
# SoA estimation (thrifty/soa_estimator.py)
# --------------------------------------------------------------------------
def calculate_window(block_len, history_len, template_len):
    """Half-open interval of correlation lags unique to one block.

    thrifty/soa_estimator.py:20-39."""
    assert history_len >= template_len - 1
    corr_len = block_len - template_len + 1
    padding = history_len - template_len + 1
    left = padding // 2
    right = padding - left
    return left, corr_len - right


def gaussian_interpolation(corr_mag, peak_idx):
    """3-point log-parabola peak offset.  thrifty/soa_estimator.py:159-170."""
    if peak_idx == 0 or peak_idx == len(corr_mag) - 1:
        return 0
    a, b, c = np.log(corr_mag[peak_idx - 1]), np.log(corr_mag[peak_idx]), np.log(corr_mag[peak_idx + 1])
    return 0.5 * (c - a) / (2 * b - a - c)


def _clip(offset, max_=0.6):
    """thrifty/soa_estimator.py:16-17."""
    return -max_ if offset < -max_ else max_ if offset > max_ else offset


class SoaEstimator(object):
    """FFT cross-correlation + windowed peak + threshold + interpolation.

    thrifty/soa_estimator.py:42-124."""

    def __init__(self, template, thresh_coeffs, block_len, history_len):
        template = np.asarray(template)
        self.template_energy = np.sum(np.abs(template) ** 2)
        tlen = len(template)
        self.corr_len = block_len - tlen + 1
        padded = np.concatenate([template, np.zeros(self.corr_len - 1)])
        self.template_fft = np.fft.fft(padded)
        self.window = calculate_window(block_len, history_len, tlen)
        self.thresh_coeffs = thresh_coeffs

    def despread(self, fft):
        """soa_estimator.py:97-102."""
        return np.fft.ifft(fft * np.conj(self.template_fft))[:self.corr_len]

    def __call__(self, fft):
        """soa_estimator.py:78-92 -> (detected, CorrDetectionInfo, corr)."""
        corr = self.despread(fft)
        corr_mag = np.abs(corr)
        start, stop = self.window
        peak_idx = int(np.argmax(corr_mag[start:stop])) + start          # :137-143
        peak_mag = corr_mag[peak_idx]
        signal_energy = np.sqrt(np.mean(np.abs(fft) ** 2)) ** 2           # fft.rms**2, :111
        with np.errstate(invalid="ignore"):
            noise_power = (signal_energy * self.template_energy - peak_mag ** 2) / len(fft)
            noise_rms = np.sqrt(noise_power)                              # may be NaN, :118-119
            t_const, t_snr, t_std = self.thresh_coeffs
            stddev = np.std(corr_mag) if t_std else 0
            thresh = np.sqrt(t_const + t_snr * noise_rms ** 2 + t_std * stddev ** 2)
            detected = bool(peak_mag > thresh)
        offset = 0 if not detected else gaussian_interpolation(corr_mag, peak_idx)
        offset = _clip(offset)
        return detected, CorrDetectionInfo(peak_idx, offset, peak_mag, noise_rms), corr


# --------------------------------------------------------------------------
# Detector (thrifty/detect.py:34-91)
# --------------------------------------------------------------------------
class Detector(object):
    """Per-block detect chain; `detect` mirrors thrifty/detect.py:60-78."""

    def __init__(self, settings, rxid=-1):
        self.settings = settings
        self.rxid = rxid
        self.soa_estimate = SoaEstimator(settings.template, settings.corr_thresh,
                                         settings.block_len, settings.history_len)
        self.new_len = settings.block_len - settings.history_len

    def sync(self, block):
        """thrifty/carrier_sync.py:52-76 with the default algorithms (:103-118)."""
        s = self.settings
        fft_mag = np.abs(np.fft.fft(block))        # complex64 in -> float32 out (numpy >= 2)
        detected, peak_idx, peak_mag, noise_rms = carrier_detect(
            fft_mag, s.carrier_thresh, s.carrier_window)
        offset = 0
        shifted_fft = None
        if detected:
            offset = dirichlet_interpolate(fft_mag, peak_idx, s.block_len, s.carrier_len)
            shifted_fft = freq_shift(block, -(peak_idx + offset))
        return shifted_fft, CarrierSyncInfo(peak_idx, offset, peak_mag, noise_rms)

    def detect(self, timestamp, block_idx, block, yield_data=False):
        assert len(block) == self.settings.block_len
        shifted_fft, carrier_info = self.sync(block)
        if shifted_fft is not None:
            detected, corr_info, corr = self.soa_estimate(shifted_fft)
            soa = self.new_len * block_idx + corr_info.sample + corr_info.offset
        else:
            detected, corr_info, soa, corr = False, None, None, None
        res = OracleResult(detected, timestamp, block_idx, soa, carrier_info, corr_info, self.rxid)
        if yield_data:
            return res, shifted_fft, corr
        return res

    def detect_raw(self, timestamp, block_idx, raw, yield_data=False):
        return self.detect(timestamp, block_idx, raw_to_complex(raw), yield_data)


def serialize(res):
    """.toad line.  thrifty/toads_data.py:47-61."""
    corr, carr = res.corr_info, res.carrier_info
    s = ("{t:.6f} {b} {s:.8f} {ps} {po} {pe} {pn} {cb} {co} {ce} {cn}".format(
        t=res.timestamp, b=res.block, s=res.soa,
        ps=corr.sample, po=corr.offset, pe=corr.energy, pn=corr.noise,
        cb=carr.bin, co=carr.offset, ce=carr.energy, cn=carr.noise))
    if res.rxid is not None:
        s = str(res.rxid) + " " + s
    return s


# --------------------------------------------------------------------------
# structured-array view used by the parity tests and the CPU baseline
# --------------------------------------------------------------------------
RECORD_DTYPE = np.dtype([
    ("block_idx", "<i8"), ("soa", "<f8"),
    ("carrier_bin", "<i4"), ("carrier_offset", "<f8"),
    ("carrier_energy", "<f8"), ("carrier_noise", "<f8"),
    ("corr_sample", "<i4"), ("corr_offset", "<f8"),
    ("corr_energy", "<f8"), ("corr_noise", "<f8"),
    ("carrier_detected", "?"), ("corr_detected", "?"),
    ("carrier_margin", "<f8"), ("corr_margin", "<f8"),
])


def result_to_row(res):
    ci, co = res.carrier_info, res.corr_info
    carrier_detected = co is not None
    return (res.block, np.nan if res.soa is None else res.soa,
            ci.bin, ci.offset, ci.energy, ci.noise,
            -1 if co is None else co.sample,
            np.nan if co is None else co.offset,
            np.nan if co is None else co.energy,
            np.nan if co is None else co.noise,
            carrier_detected, bool(res.detected), np.nan, np.nan)


def detect_blocks(settings, raw_blocks, block_indices=None, rxid=0):
    """Run the oracle over uint8 blocks [B, 2N]; returns a RECORD_DTYPE array.

    carrier_margin / corr_margin = peak / threshold (values within 1e-3 of 1
    are 'marginal': float32-vs-float64 rounding may flip the verdict)."""
    det = Detector(settings, rxid=rxid)
    raw_blocks = np.asarray(raw_blocks, dtype=np.uint8)
    nblk = raw_blocks.shape[0]
    if block_indices is None:
        block_indices = np.arange(nblk)
    out = np.zeros(nblk, dtype=RECORD_DTYPE)
    for i in range(nblk):
        res = det.detect_raw(0.0, int(block_indices[i]), raw_blocks[i])
        out[i] = result_to_row(res)
        ci = res.carrier_info
        t_c, t_s, _ = settings.carrier_thresh
        with np.errstate(invalid="ignore", divide="ignore"):
            out[i]["carrier_margin"] = float(ci.energy) / np.sqrt(
                t_c + t_s * float(ci.noise) ** 2) if not settings.carrier_thresh[2] else np.nan
            if res.corr_info is not None and not settings.corr_thresh[2]:
                k_c, k_s, _ = settings.corr_thresh
                out[i]["corr_margin"] = float(res.corr_info.energy) / np.sqrt(
                    k_c + k_s * float(res.corr_info.noise) ** 2)
    return out
