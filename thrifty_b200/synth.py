"""Deterministic synthetic inputs for the detect path.

The reference ships no sample ``.card`` capture, so every parity test and
benchmark runs on synthetic blocks: an OOK/DSSS positioning burst (template
gated carrier) at a random fractional carrier bin plus complex Gaussian noise,
quantised to RTL-SDR style unsigned 8-bit I/Q with the reference's
``complex_to_raw`` rule (thrifty/block_data.py:55-67).  Also provides a py3
Gold-code template generator following thrifty/gold.py:26-82 and
thrifty/template_generate.py:39-45 (the reference versions are py2-only).
"""

from __future__ import annotations

import numpy as np

SEED0 = 20161125

# Preferred LFSR tap pairs (thrifty/gold.py:14-23)
_TAPS = {
    5: [[2], [1, 2, 3]],
    6: [[5], [1, 4, 5]],
    7: [[4], [4, 5, 6]],
    8: [[1, 2, 3, 6, 7], [1, 2, 7]],
    9: [[5], [3, 5, 6]],
    10: [[2, 5, 9], [3, 4, 6, 8, 9]],
    11: [[9], [3, 6, 9]],
}


def lfsr(taps, nbits):
    """Maximal-length sequence from an all-ones seed (thrifty/gold.py:54-82)."""
    seq_len = (1 << nbits) - 1
    seq = np.zeros(seq_len, dtype=bool)
    seq[:nbits] = True
    for i in range(nbits, seq_len):
        bit = seq[i - nbits]
        for tap in taps:
            bit ^= seq[i - nbits + tap]
        seq[i] = bit
    return seq


def gold(bits, idx):
    """idx-th Gold code of length 2**bits - 1 (thrifty/gold.py:26-51)."""
    if bits not in _TAPS:
        raise ValueError("Preferred pairs for %d bits unknown." % bits)
    seq1 = lfsr(_TAPS[bits][0], bits)
    seq2 = lfsr(_TAPS[bits][1], bits)
    if idx == 0:
        return seq1
    if idx == 1:
        return seq2
    return np.logical_xor(seq1, np.roll(seq2, -idx + 2))


def resample(code, sps):
    """Integer sampler, +-1 symbols (thrifty/template_generate.py:39-45)."""
    length = int(sps * len(code))
    indices = np.arange(length) * len(code) // length
    return np.where(code, 1, -1)[indices]


def gold_template(bits, idx=0, sps=2.4e6 / 0.999707e6):
    """Real +-1 template: Gold code `idx` sampled at `sps` samples per chip."""
    return resample(gold(bits, idx), sps).astype(np.float64)


def default_geometry(block_len):
    """(gold bits, template_len, history_len) per block_len as in SURVEY 8d cfg 3."""
    bits = {4096: 9, 8192: 10}.get(block_len)
    if bits is None:
        raise ValueError("no default Gold geometry for block_len=%d" % block_len)
    tpl = gold_template(bits)
    return bits, len(tpl), len(tpl) + 6


def complex_to_raw(x):
    """uint8(x*128 + 127.4) per component (thrifty/block_data.py:55-67)."""
    scaled = np.asarray(x).astype(np.complex64).view(np.float32) * 128 + 127.4
    return scaled.astype(np.uint8)


def peak_window(block_len, history_len, template_len):
    """Half-open unique-lag window (thrifty/soa_estimator.py:20-39)."""
    corr_len = block_len - template_len + 1
    padding = history_len - template_len + 1
    left = padding // 2
    return left, corr_len - (padding - left)


def make_block(rng, block_len, history_len, template, p_signal=0.5,
               bin_range=(8.0, 109.0), force_pos=None):
    """One synthetic block.  Returns (raw uint8[2N], truth dict)."""
    n = block_len
    tlen = len(template)
    start, stop = peak_window(n, history_len, tlen)
    has_signal = rng.random() < p_signal
    pos = int(rng.integers(start, stop))
    if len(bin_range) == 4:      # two disjoint ranges (e.g. either side of DC)
        lo, hi = bin_range[:2] if rng.random() < 0.5 else bin_range[2:]
        fbin = rng.uniform(lo, hi)
    else:
        fbin = rng.uniform(*bin_range)
    amp = rng.uniform(0.1, 0.5)
    phi = rng.uniform(0, 2 * np.pi)
    sigma = rng.uniform(0.005, 0.05)
    if force_pos is not None:
        pos = force_pos
    x = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    if has_signal:
        idx = np.arange(n)
        carrier = np.exp(1j * (2 * np.pi * fbin * idx / n + phi))
        gate = np.zeros(n)
        end = min(n, pos + tlen)
        gate[pos:end] = (template[:end - pos] + 1) / 2
        x = x + amp * gate * carrier
    x = np.clip(x.real, -0.989, 0.989) + 1j * np.clip(x.imag, -0.989, 0.989)
    truth = dict(signal=has_signal, pos=pos, bin=fbin, amp=amp, sigma=sigma)
    return complex_to_raw(x), truth


def make_blocks(n_blocks, block_len, history_len, template, p_signal=0.5,
                seed=SEED0, bin_range=(8.0, 109.0)):
    """`n_blocks` blocks; block b uses default_rng(seed + b).  -> (uint8[B,2N], truths)."""
    raw = np.empty((n_blocks, 2 * block_len), dtype=np.uint8)
    truths = []
    for b in range(n_blocks):
        rng = np.random.default_rng(seed + b)
        raw[b], t = make_block(rng, block_len, history_len, template, p_signal, bin_range)
        truths.append(t)
    return raw, truths
