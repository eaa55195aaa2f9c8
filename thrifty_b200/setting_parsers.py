"""String -> value parsers for detector settings (grammar of thrifty/setting_parsers.py).

`detector.cfg` feeds thresholds and the carrier window to the hot path through these, so the
accepted grammar is the reference's: metric floats ('2.4M'), frequency ranges ('7 - 110',
'-80k - -79kHz') and threshold formulas ('15*snr', '100c+2s', '1+2s+3d').
"""
import re

_FLOAT = r"[-+]?(?:\d+(?:\.\d*)?|\.\d+)(?:[eE][-+]?\d+)?"
_RANGE_RE = re.compile(r"^({0})(?:\s*-\s*({0}))?\s*([kKmM]?)([hH][zZ])?$".format(_FLOAT))
_TERM_RE = re.compile(r"^\s*(?=\S)(?:({0})\s*\*?\s*)?(constant|c|snr|s|stddev|d|)\s*$".format(_FLOAT))

_SI = {"y": 1e-24, "z": 1e-21, "a": 1e-18, "f": 1e-15, "p": 1e-12, "n": 1e-9, "u": 1e-6,
       "m": 1e-3, "c": 1e-2, "d": 1e-1, "k": 1e3, "M": 1e6, "G": 1e9, "T": 1e12, "P": 1e15,
       "E": 1e18, "Z": 1e21, "Y": 1e24}


def metric_float(string):
    """'1.2M' -> 1200000.0 (setting_parsers.py:44-61)."""
    string = string.strip()
    if string and string[-1] in _SI:
        return float(string[:-1]) * _SI[string[-1]]
    return float(string)


def freq_range(string):
    """'a-b [k|M][Hz]' -> (start, stop, unit_is_hz) (setting_parsers.py:64-114)."""
    match = _RANGE_RE.match(string)
    if not match:
        raise ValueError("Invalid range: {}".format(string))
    start_s, stop_s, prefix, unit = match.groups()
    if stop_s is None:
        stop_s = start_s
    start, stop = float(start_s), float(stop_s)
    scale = {"k": 1e3, "m": 1e6}.get(prefix.lower(), 1.0)
    if scale != 1.0:
        start, stop = start * scale, stop * scale
    return start, stop, unit is not None


def normalize_freq_range(range_, bin_freq):
    """(start, stop, is_hz) -> integer bins, truncating (setting_parsers.py:117-138)."""
    start, stop, is_hz = range_
    if not is_hz:
        return int(start), int(stop)
    return int(start / bin_freq), int(stop / bin_freq)


def threshold(string):
    """'5 + 3*snr + stddev' -> (constant, snr, stddev) (setting_parsers.py:141-185)."""
    if not string:
        raise ValueError("Empty string")
    acc = {"c": 0.0, "s": 0.0, "d": 0.0}
    for term in string.split("+"):
        match = _TERM_RE.match(term)
        if not match:
            raise ValueError("Invalid threshold term: {}".format(term))
        qty_s, symbol = match.groups()
        qty = 1.0 if qty_s is None else float(qty_s)
        key = {"constant": "c", "c": "c", "": "c", "snr": "s", "s": "s", "stddev": "d", "d": "d"}[symbol]
        acc[key] += qty
    return acc["c"], acc["s"], acc["d"]
