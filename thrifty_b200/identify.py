"""Merge RX detections, identify transmitter IDs, filter duplicates -- the step after `detect`, on the GPU.

Drop-in for thrifty/identify.py:26-257 (same function names and results), organised the other way round: the reference
walks Python lists of DetectionResult; here the detections become columns (rxid, txid, block, timestamp, corr energy,
carrier bin, carrier offset) and the three data-parallel steps run as CUDA kernels behind the C ABI
(thrifty_b200/csrc/identify.cu): frequency-map classification (one thread per detection), the carrier-bin histogram of the
automatic classification, and the duplicate mask -- np.argsort over (rxid, txid, block, timestamp) as a device bitonic sort
followed by neighbour compares.  Only what is sequential or I/O stays on the host: the scan of the ~100 histogram bins for
peaks, the frequency-map / .toad file formats, and the final ordering of the survivors by timestamp.  There is no CPU
fallback for the kernels.  `columns_from_records` feeds thr_record arrays (one receiver's detect output, e.g. gathered from
several GPUs) in without building result objects.

    python -m thrifty_b200 identify rx0.toad rx1.toad -o data.toads [-m freqmap.cfg]
"""

from __future__ import print_function

import argparse
import ctypes
import glob

import numpy as np

from thrifty_b200 import toads_data
from thrifty_b200._native import FLAG_CORR, NativeError, load_library
from thrifty_b200.settings import parse_kvconfig

UNIDENTIFIED = -1     # identify.py:112 "FIXME: don't use magic number"


class Columns(object):
    """Detections as columns (the layout the kernels take)."""

    def __init__(self, rxid, block, timestamp, energy, carrier_bin, carrier_offset, txid=None):
        self.rxid = np.ascontiguousarray(rxid, dtype=np.int32)
        self.block = np.ascontiguousarray(block, dtype=np.int32)          # toads_array: 'block' is i4
        self.timestamp = np.ascontiguousarray(timestamp, dtype=np.float64)
        self.energy = np.ascontiguousarray(energy, dtype=np.float64)      # correlation peak magnitude
        self.carrier_bin = np.ascontiguousarray(carrier_bin, dtype=np.int32)
        self.carrier_offset = np.ascontiguousarray(carrier_offset, dtype=np.float64)
        n = len(self.rxid)
        self.txid = (np.full(n, UNIDENTIFIED, dtype=np.int32) if txid is None
                     else np.ascontiguousarray(txid, dtype=np.int32))

    def __len__(self):
        return len(self.rxid)


def columns_from_detections(detections):
    n = len(detections)
    return Columns([d.rxid for d in detections], [d.block for d in detections], [d.timestamp for d in detections],
                   [d.corr_info.energy for d in detections], [d.carrier_info.bin for d in detections],
                   [d.carrier_info.offset for d in detections],
                   [UNIDENTIFIED if getattr(d, "txid", None) is None else d.txid for d in detections] if n else [])


def columns_from_records(records, timestamps, rxid):
    """thr_record array (one template) + timestamps -> (Columns of the DETECTED blocks, their positions in `records`):
    what a .toad file of that receiver would hold (detect.py:218-219 writes detected blocks only)."""
    records = np.asarray(records)
    sel = np.nonzero((records["flags"] & FLAG_CORR) != 0)[0]
    r = records[sel]
    ts = np.broadcast_to(np.asarray(timestamps, dtype=np.float64), records.shape)[sel]
    return Columns(np.full(len(sel), rxid), r["block_idx"], ts, r["corr_energy"], r["carrier_bin"], r["carrier_offset"]), sel


def _ptr(a):
    return a.ctypes.data if len(a) else None


def _check(lib, rc, what):
    if rc != 0:
        raise NativeError("%s failed (%d): %s" % (what, rc, lib.thr_identify_last_error().decode()))


def _bind(lib):
    if getattr(lib, "_identify_bound", False):
        return lib
    i32, i64, f64, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    lib.thr_identify_classify.argtypes = [i32, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.thr_identify_bin_histogram.argtypes = [i32, i64, vp, vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), vp, i32]
    lib.thr_identify_digitize.argtypes = [i32, i64, vp, vp, i32, i32, vp, vp]
    lib.thr_identify_duplicates.argtypes = [i32, i64, vp, vp, vp, vp, vp, vp]
    lib.thr_identify_last_error.restype = ctypes.c_char_p
    for f in ("thr_identify_classify", "thr_identify_bin_histogram", "thr_identify_digitize", "thr_identify_duplicates"):
        getattr(lib, f).restype = ctypes.c_int
    lib._identify_bound = True
    return lib


def _lib():
    return _bind(load_library())


# ---------------------------------------------------------------------------------------------------------------------
def bin_histogram(cols, rxid, device=0):
    """(first_bin, counts) of the carrier bins of one receiver -- np.bincount(freqs - min(freqs)) on the device."""
    lib = _lib()
    first, nb = ctypes.c_int32(0), ctypes.c_int32(0)
    _check(lib, lib.thr_identify_bin_histogram(device, len(cols), _ptr(cols.rxid), _ptr(cols.carrier_bin), int(rxid),
                                               ctypes.byref(first), ctypes.byref(nb), None, 0), "thr_identify_bin_histogram")
    counts = np.zeros(nb.value, dtype=np.uint32)
    if nb.value:
        _check(lib, lib.thr_identify_bin_histogram(device, len(cols), _ptr(cols.rxid), _ptr(cols.carrier_bin), int(rxid),
                                                   ctypes.byref(first), ctypes.byref(nb), counts.ctypes.data, len(counts)),
               "thr_identify_bin_histogram")
    return first.value, counts.astype(np.int64)


def windows_from_histogram(first_bin, cnts, verbose=False):
    """Histogram -> edges of the transmitter frequency windows: the sequential part of identify.py:26-77 (hysteresis
    thresholds at 0.4 / 1.25 standard deviations of the counts, window edges half-way between neighbouring peaks)."""
    last_bin = first_bin + len(cnts)
    spread = np.std(cnts)
    low_thresh, high_thresh = spread * 0.4, spread * 1.25
    peaks, start, quiet = [], None, True
    for i, cnt in enumerate(cnts):
        if not quiet and cnt < low_thresh:
            peaks.append((start, i))
            start, quiet = None, True
        if quiet and cnt > high_thresh:
            start, quiet = i, False
    if not quiet:
        peaks.append((start, len(cnts) - 1))
    inner = [(peaks[i][1] + peaks[i + 1][0]) // 2 + first_bin for i in range(len(peaks) - 1)]
    edges = np.array([first_bin] + inner + [last_bin], dtype=np.int64)
    if verbose:
        print("Window threshold: low = {}; high = {}:".format(low_thresh, high_thresh))
        print("Detected {} transmitter(s):".format(len(edges) - 1))
    return edges


def detect_transmitter_windows(freqs, verbose=False, device=0):
    """Carrier bins of one receiver's detections -> window edges (identify.py:26-77)."""
    freqs = np.ascontiguousarray(freqs, dtype=np.int32)
    cols = Columns(np.zeros(len(freqs)), np.zeros(len(freqs)), np.zeros(len(freqs)), np.zeros(len(freqs)), freqs,
                   np.zeros(len(freqs)))
    first, cnts = bin_histogram(cols, 0, device)
    return windows_from_histogram(first, cnts, verbose)


def auto_classify_columns(cols, verbose=True, device=0):
    """txid from the carrier bin, windows detected per receiver (identify.py:80-103)."""
    lib = _lib()
    txid = np.full(len(cols), UNIDENTIFIED, dtype=np.int32)
    for rxid in sorted(set(cols.rxid.tolist())):
        first, cnts = bin_histogram(cols, rxid, device)
        rx_edges = windows_from_histogram(first, cnts)
        if verbose:
            print("Detected {} transmitter(s) at RX {}:".format(len(rx_edges) - 1, rxid)
                  + "".join(" {}-{}".format(rx_edges[i], rx_edges[i + 1] - 1) for i in range(len(rx_edges) - 1)))
        edges = np.ascontiguousarray(rx_edges[:-1], dtype=np.int64)
        _check(lib, lib.thr_identify_digitize(device, len(cols), _ptr(cols.rxid), _ptr(cols.carrier_bin), int(rxid),
                                              len(edges), _ptr(edges), _ptr(txid)), "thr_identify_digitize")
    return txid


def classify_columns(cols, freqmap, device=0):
    """txid = the nominal frequency range holding bin + offset (identify.py:106-118; a receiver that is missing from the
    map raises KeyError like the reference's dict lookup)."""
    lib = _lib()
    for rxid in set(cols.rxid.tolist()):
        freqmap[rxid]                                   # KeyError as in the reference
    rows = [(rx, tx, lo, hi) for rx, ranges in freqmap.items() for tx, (lo, hi) in ranges.items()]
    m_rx = np.array([r[0] for r in rows], dtype=np.int32)
    m_tx = np.array([r[1] for r in rows], dtype=np.int32)
    m_lo = np.array([r[2] for r in rows], dtype=np.float64)
    m_hi = np.array([r[3] for r in rows], dtype=np.float64)
    txid = np.full(len(cols), UNIDENTIFIED, dtype=np.int32)
    _check(lib, lib.thr_identify_classify(device, len(cols), _ptr(cols.rxid), _ptr(cols.carrier_bin), _ptr(cols.carrier_offset),
                                          len(rows), _ptr(m_rx), _ptr(m_tx), _ptr(m_lo), _ptr(m_hi), _ptr(txid)),
           "thr_identify_classify")
    return txid


def duplicates_mask_columns(cols, device=0):
    """keep-mask (identify.py:134-164) for columns whose txid is set."""
    lib = _lib()
    keep = np.zeros(len(cols), dtype=np.uint8)
    _check(lib, lib.thr_identify_duplicates(device, len(cols), _ptr(cols.rxid), _ptr(cols.txid), _ptr(cols.block),
                                            _ptr(cols.timestamp), _ptr(cols.energy), _ptr(keep)), "thr_identify_duplicates")
    return keep.astype(bool)


# ---- the reference's function names, on lists of DetectionResult ----------------------------------------------------
def auto_classify_transmitters(detections, verbose=True, device=0):
    return auto_classify_columns(columns_from_detections(detections), verbose, device).tolist()


def classify_transmitters(detections, freqmap, device=0):
    return classify_columns(columns_from_detections(detections), freqmap, device).tolist()


def identify_transmitters(detections, freqmap=None, verbose=True, device=0):
    """Set ``txid`` on every detection, in place (identify.py:121-133)."""
    txids = (auto_classify_transmitters(detections, verbose, device) if freqmap is None
             else classify_transmitters(detections, freqmap, device))
    for det, txid in zip(detections, txids):
        det.txid = txid


def identify_duplicates(detections, device=0):
    """Mask that drops the weaker of two detections of one transmitter in adjacent blocks, and unidentified ones."""
    return duplicates_mask_columns(columns_from_detections(detections), device)


def filter_duplicates(detections, device=0):
    """Detections without duplicates / unidentified ones, sorted by timestamp (identify.py:169-175)."""
    mask = identify_duplicates(detections, device)
    kept = [d for d, k in zip(detections, mask) if k]
    kept.sort(key=lambda d: d.timestamp)
    return kept


def integrate(detections, freqmap=None, verbose=True, device=0):
    """Identify and filter (identify.py:216-220)."""
    identify_transmitters(detections, freqmap, verbose, device)
    return filter_duplicates(detections, device)


def integrate_records(records, timestamps, rxid, freqmap=None, device=0):
    """thr_record array of one receiver (e.g. gathered from several GPUs) -> (positions of the detections that survive,
    their txids), without building result objects."""
    cols, sel = columns_from_records(records, timestamps, rxid)
    cols.txid = auto_classify_columns(cols, False, device) if freqmap is None else classify_columns(cols, freqmap, device)
    keep = duplicates_mask_columns(cols, device)
    order = np.argsort(cols.timestamp[keep], kind="stable")
    return sel[keep][order], cols.txid[keep][order]


# ---- file formats (host) ---------------------------------------------------------------------------------------------
def load_toad_files(toad_globs):
    """identify.py:178-188."""
    filenames = []
    for pattern in toad_globs:
        filenames.extend(sorted(glob.glob(pattern)))
    detections = []
    for filename in filenames:
        with open(filename, "r") as file_:
            detections.extend(toads_data.load_toad(file_))
    return detections, filenames


def load_freqmap(file_):
    """'txid: start - stop' and '@rxid: offset' lines -> {rxid: {txid: (start, stop)}} (identify.py:191-213)."""
    if file_ is None:
        return None
    tx_ranges, rx_offset = {}, {}
    for key, value in parse_kvconfig(file_).items():
        if key[0] == "@":
            rx_offset[int(key[1:])] = float(value)
        else:
            start, stop = [float(x.strip()) for x in value.split("-")]
            tx_ranges[int(key)] = (start, stop)
    return {rxid: {txid: (start + offset, stop + offset) for txid, (start, stop) in tx_ranges.items()}
            for rxid, offset in rx_offset.items()}


def generate_toads(output, toad_globs, freqmap, verbose=True, device=0):
    """identify.py:223-234."""
    detections, filenames = load_toad_files(toad_globs)
    output.write("# source_files: [%s]\n" % (" ".join(filenames)))
    filtered = integrate(detections, freqmap, verbose, device)
    if verbose:
        print("Removed {} duplicates / unidentified transmisisons from {} detections."
              .format(len(detections) - len(filtered), len(detections)))
    for det in filtered:
        output.write(det.serialize() + "\n")
    return filtered


def _main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("toad_file", type=str, nargs="*", default=["*.toad"],
                        help="toad file(s) from receivers [default: *.toad]")
    parser.add_argument("-o", "--output", type=argparse.FileType("w"), default="data.toads",
                        help="output file [default: data.toads]")
    parser.add_argument("-m", "--map", type=argparse.FileType("r"),
                        help="schema for mapping DFT index to transmitter ID [default: auto-detect]")
    parser.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    args = parser.parse_args(argv)
    generate_toads(args.output, args.toad_file, load_freqmap(args.map), device=args.device)
    args.output.flush()


if __name__ == "__main__":
    _main()
