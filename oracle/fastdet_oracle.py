"""CPU oracle for the reference's NATIVE detect path (fastcard + fastdet)  --  TEST INFRASTRUCTURE ONLY.

`fastdet` is the C/C++ twin of `thrifty detect` with different semantics:
single precision throughout, decisions on POWERS, integer-bin carrier shift (a circular
roll of the spectrum, no fractional mix), 3-point parabolic carrier offset (reporting only),
Gaussian correlation offset clipped to +-0.5, noise clamped at 0.

Two checkers live here:

* ``run_reference(...)``: the reference's OWN sources compiled by ``oracle/Makefile`` into
  ``oracle/_ref/libfastdet_ref.so`` and driven through ``fastdet_ref_driver.cpp`` exactly like
  ``fastdet/fastdet.cpp:113-186`` (FFTW/VOLK replaced by the stand-ins in ``oracle/shim``).
* ``detect_blocks(...)``: a NumPy restatement, each step citing the reference file:line.

Parity status: PINNED against the compiled reference sources -- ``oracle/make_golden_fastdet.py``
runs both on the same seeded `.card` / raw files, asserts they agree (indices exact, powers to
2e-5 relative: NumPy's pocketfft vs the shim FFT) and stores the compiled reference's outputs in
``tests/golden/fastdet_*.npz``.  The reference itself has no tests for this path.

Only ``tests/``, ``oracle/`` scripts and ``bench.py``'s CPU legs may import this module.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libfastdet_ref.so")

# mirrors struct ref_fastdet_record in fastdet_ref_driver.cpp
REF_RECORD_DTYPE = np.dtype([
    ("block_idx", "<i8"), ("ts_sec", "<i8"), ("ts_usec", "<i8"),
    ("soa", "<f8"), ("corr_offset", "<f8"), ("carrier_offset", "<f8"),
    ("carrier_max", "<f4"), ("carrier_noise", "<f4"), ("carrier_threshold", "<f4"), ("fft_sum", "<f4"),
    ("corr_peak_power", "<f4"), ("corr_noise_power", "<f4"), ("corr_threshold", "<f4"),
    ("carrier_detected", "<i4"), ("carrier_argmax", "<i4"),
    ("corr_detected", "<i4"), ("corr_peak_idx", "<i4"), ("pad", "<i4"),
])
assert REF_RECORD_DTYPE.itemsize == 96

_ref = None


def have_reference():
    return os.path.exists(REF_LIB)


def _load_ref():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(REF_LIB)
        lib.ref_fastdet_run.restype = ctypes.c_int
        lib.ref_fastdet_run.argtypes = [
            ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        lib.ref_rawconv.restype = None
        lib.ref_rawconv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _ref = lib
    return _ref


def run_reference(path, input_card, block_len, history_len, thresh, window, template, corr_thresh,
                  max_blocks=1 << 20):
    """Run the compiled reference (fastcard reader -> CarrierDetector -> CorrDetector) over a file.

    thresh / corr_thresh = (constant, snr) in the POWER domain (fastcard/parse.c:54-99 '<c>c<s>s')."""
    lib = _load_ref()
    tpl = np.ascontiguousarray(template, dtype=np.float32)
    out = np.zeros(max_blocks, dtype=REF_RECORD_DTYPE)
    err = ctypes.create_string_buffer(256)
    n = lib.ref_fastdet_run(os.fsencode(path), int(bool(input_card)), block_len, history_len,
                            float(thresh[0]), float(thresh[1]), int(window[0]), int(window[1]),
                            tpl.ctypes.data, len(tpl), float(corr_thresh[0]), float(corr_thresh[1]),
                            out.ctypes.data, max_blocks, err, len(err))
    if n < 0:
        raise RuntimeError("reference fastdet failed: %s" % err.value.decode())
    return out[:min(n, max_blocks)].copy()


def reference_rawconv(raw):
    lib = _load_ref()
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.zeros(len(raw) // 2, dtype=np.complex64)
    lib.ref_rawconv(raw.ctypes.data, len(raw) // 2, out.ctypes.data)
    return out


# ------------------------------------------------------------------------------------------
# NumPy restatement
# ------------------------------------------------------------------------------------------
def rawconv(raw):
    """fastcard/rawconv.c:5-28: LUT of ((float)b - 127.4f) * (1.0f/128.0f) per component."""
    v = np.asarray(raw, dtype=np.uint8).astype(np.float32)
    v = (v - np.float32(127.4)) * np.float32(1.0 / 128.0)
    return v.view(np.complex64)


def normalize_window(win_min, win_max, n):
    """fastcard/cardet.c:43-69 cardet_normalize_window: no zero-straddling window, closed interval."""
    if win_min < 0 and win_max >= 0:
        raise ValueError("Carrier frequency window range not supported.")
    if win_min < 0:
        win_min += n
    if win_max < 0:
        win_max += n
    if win_min >= n or win_max >= n:
        raise ValueError("Carrier frequency window out of range.")
    if win_max < win_min:
        win_min, win_max = win_max, win_min
    return win_min, win_max


def calculate_window(block_len, history_len, template_len):
    """fastdet/corr_detector.cpp:73-86 set_window (== soa_estimator.py:20-39)."""
    assert history_len >= template_len - 1
    padding = history_len - template_len + 1
    left = padding // 2
    right = padding - left
    corr_len = block_len - template_len + 1
    return left, corr_len - right


def _interp(a, b, c):
    """(c - a) / (4b - 2a - 2c) clipped to +-0.5, in double (corr_detector.cpp:88-116)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        off = (c - a) / (4 * b - 2 * a - 2 * c)
    if off < -0.5:
        off = -0.5
    if off > 0.5:
        off = 0.5
    return float(off)


class FastDetector(object):
    """fastcard_process (fastcard/fastcard.c:177-189) + cardet_detect (cardet.c:7-41) +
    CorrDetector (fastdet/corr_detector.cpp:31-197) on one block of uint8 I/Q."""

    def __init__(self, block_len, history_len, thresh, window, template, corr_thresh):
        self.n = block_len
        self.h = history_len
        self.thresh = (np.float32(thresh[0]), np.float32(thresh[1]))
        self.corr_thresh = (np.float32(corr_thresh[0]), np.float32(corr_thresh[1]))
        self.wmin, self.wmax = normalize_window(window[0], window[1], block_len)
        tpl = np.asarray(template, dtype=np.float32)
        self.tpl_len = len(tpl)
        self.corr_len = block_len - len(tpl) + 1
        padded = np.zeros(block_len, dtype=np.complex64)
        padded[:len(tpl)] = tpl
        # corr_detector.cpp:51-71 set_template: conj(FFT(template || 0)) and sum(t^2) in float32
        self.template_fft_conj = np.conj(np.fft.fft(padded)).astype(np.complex64)
        e = np.float32(0)
        for v in tpl:                       # sequential float accumulation like the C loop
            e = np.float32(e + v * v)
        self.template_energy = e
        self.start, self.stop = calculate_window(block_len, history_len, len(tpl))

    def detect(self, block_idx, raw):
        n = self.n
        rec = np.zeros((), dtype=REF_RECORD_DTYPE)
        rec["block_idx"] = block_idx
        x = rawconv(raw)
        fft = np.fft.fft(x).astype(np.complex64)                               # fastcard.c:179
        power = (fft.real * fft.real + fft.imag * fft.imag).astype(np.float32)  # fastcard.c:180
        fsum = np.float32(0)
        fsum = np.float32(np.sum(power, dtype=np.float32))                     # cardet.c:12 (order differs)
        argmax = self.wmin + int(np.argmax(power[self.wmin:self.wmax + 1]))     # cardet.c:15-19
        mx = power[argmax]
        noise = np.float32(0)
        if fsum != 0:
            noise = np.float32((fsum - np.float32(2) * mx) / np.float32(n - 1))  # cardet.c:22-25
        threshold = np.float32(self.thresh[0] + self.thresh[1] * noise)
        if not (mx > threshold):
            return rec                                                          # cardet.c:29-40
        rec["carrier_detected"] = 1
        rec["carrier_argmax"] = argmax
        rec["carrier_max"] = mx
        rec["carrier_noise"] = noise
        rec["carrier_threshold"] = threshold
        rec["fft_sum"] = fsum
        # corr_detector.cpp:177-184: roll by -argmax, signal_energy = fft_sum / len
        shifted = np.roll(fft, -argmax)
        signal_energy = np.float32(fsum / np.float32(n))
        # corr_detector.cpp:127-141: multiply, backward FFT, /len on the first corr_len values
        corr_fft = (shifted * self.template_fft_conj).astype(np.complex64)
        corr = (np.fft.ifft(corr_fft) * n).astype(np.complex64)[:self.corr_len]
        corr = (corr / np.float32(n)).astype(np.complex64)
        cp = (corr.real * corr.real + corr.imag * corr.imag).astype(np.float32)  # :144-146
        peak_idx = self.start + int(np.argmax(cp[self.start:self.stop]))          # :149-155
        peak_power = cp[peak_idx]
        # :118-125 estimate_noise -- the peak power is passed as size_t (truncated towards zero)
        trunc_peak = np.float32(int(peak_power))
        noise_power = np.float32((np.float32(signal_energy * self.template_energy) - trunc_peak) / np.float32(n))
        if noise_power < 0:
            noise_power = np.float32(0)
        cthr = np.float32(self.corr_thresh[0] + self.corr_thresh[1] * noise_power)   # :158
        detected = bool(peak_power > cthr)
        offset = 0.0
        if detected:                                                             # :103-116, :164
            a, b, c = (np.log(np.sqrt(np.float64(cp[peak_idx + d]))) for d in (-1, 0, 1))
            offset = _interp(a, b, c)
        # :88-101 parabolic interpolation of the carrier peak on sqrt(power), reporting only
        pa, pb, pc = (np.sqrt(np.float64(power[(argmax + d) % n])) if 0 <= argmax + d < n else np.nan
                      for d in (-1, 0, 1))
        rec["carrier_offset"] = _interp(pa, pb, pc)
        rec["corr_detected"] = int(detected)
        rec["corr_peak_idx"] = peak_idx
        rec["corr_offset"] = offset
        rec["corr_peak_power"] = peak_power
        rec["corr_noise_power"] = noise_power
        rec["corr_threshold"] = cthr
        rec["soa"] = float((n - self.h) * int(block_idx) + peak_idx) + offset      # fastdet.cpp:184-186
        return rec


def detect_blocks(block_len, history_len, thresh, window, template, corr_thresh, raw_blocks, block_indices=None):
    det = FastDetector(block_len, history_len, thresh, window, template, corr_thresh)
    raw_blocks = np.asarray(raw_blocks, dtype=np.uint8)
    if block_indices is None:
        block_indices = np.arange(len(raw_blocks))
    out = np.zeros(len(raw_blocks), dtype=REF_RECORD_DTYPE)
    for i in range(len(raw_blocks)):
        out[i] = det.detect(int(block_indices[i]), raw_blocks[i])
    return out


def toad_line(rec, rxid=0):
    """fastdet/fastdet.cpp:191-207: sqrt of the powers, same 12 columns as toads_data.py:47-61."""
    return ("%d %d.%06d %d %.8f %u %.12f %f %f %u %f %f %f" % (
        rxid, rec["ts_sec"], rec["ts_usec"], rec["block_idx"], rec["soa"], rec["corr_peak_idx"],
        rec["corr_offset"], np.sqrt(rec["corr_peak_power"]), np.sqrt(rec["corr_noise_power"]),
        rec["carrier_argmax"], rec["carrier_offset"], np.sqrt(rec["carrier_max"]),
        np.sqrt(rec["carrier_noise"])))
