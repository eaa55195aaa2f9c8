"""CPU restatement of thrifty/identify.py:26-257 -- TEST INFRASTRUCTURE ONLY.

Python-3 restatement of the reference's `identify` step on lists of DetectionResult, pinned by
tests/golden/identify_cases.npz (written by oracle/make_golden_identify.py with the reference's own functions).  It is
the checker for the CUDA path (thrifty_b200/csrc/identify.cu behind thrifty_b200/identify.py); only tests/ and bench.py's
checking legs may import it.
"""

from __future__ import print_function

import argparse
import glob
import itertools
from collections import defaultdict

import numpy as np

from thrifty_b200 import toads_data
from thrifty_b200.settings import parse_kvconfig

UNIDENTIFIED = -1     # identify.py:112 "FIXME: don't use magic number"


def detect_transmitter_windows(freqs, verbose=False):
    """Carrier-bin histogram -> edges of the transmitter frequency windows (identify.py:26-77)."""
    freqs = np.asarray(freqs)
    first_bin = np.min(freqs)
    cnts = np.bincount(freqs - first_bin)
    last_bin = first_bin + len(cnts)
    low_thresh = np.std(cnts) * 0.4
    high_thresh = np.std(cnts) * 1.25

    peaks = []
    below_thresh = True
    above_thresh_start = None
    for i, cnt in enumerate(cnts):
        if not below_thresh and cnt < low_thresh:
            peaks.append((above_thresh_start, i))
            above_thresh_start = None
            below_thresh = True
        if below_thresh and cnt > high_thresh:
            above_thresh_start = i
            below_thresh = False
    if not below_thresh:
        peaks.append((above_thresh_start, len(cnts) - 1))

    edges = [(peaks[i][1] + peaks[i + 1][0]) // 2 for i in range(len(peaks) - 1)]
    edges = np.concatenate([[first_bin], np.array(edges, dtype=np.int64) + first_bin, [last_bin]])
    if verbose:
        print("Window threshold: low = {}; high = {}:".format(low_thresh, high_thresh))
        print("Detected {} transmitter(s):".format(len(edges) - 1))
    return edges


def auto_classify_transmitters(detections, verbose=True):
    """txid from the carrier bin, windows detected per receiver (identify.py:80-103)."""
    by_rx = defaultdict(list)
    for det in detections:
        by_rx[det.rxid].append(det.carrier_info.bin)
    edges = {}
    for rxid, bins in by_rx.items():
        rx_edges = detect_transmitter_windows(np.array(bins))
        if verbose:
            print("Detected {} transmitter(s) at RX {}:".format(len(rx_edges) - 1, rxid)
                  + "".join(" {}-{}".format(rx_edges[i], rx_edges[i + 1] - 1) for i in range(len(rx_edges) - 1)))
        edges[rxid] = rx_edges[:-1]
    return [int(np.digitize(d.carrier_info.bin, edges[d.rxid]) - 1) for d in detections]


def classify_transmitters(detections, freqmap):
    """txid = the nominal frequency range holding bin + offset (identify.py:106-118)."""
    txids = []
    for det in detections:
        freq = det.carrier_info.bin + det.carrier_info.offset
        this_txid = UNIDENTIFIED
        for txid, (start, stop) in freqmap[det.rxid].items():
            if start <= freq <= stop:
                this_txid = txid
        txids.append(this_txid)
    return txids


def identify_transmitters(detections, freqmap=None, verbose=True):
    """Set ``txid`` on every detection, in place (identify.py:121-133)."""
    txids = (auto_classify_transmitters(detections, verbose) if freqmap is None
             else classify_transmitters(detections, freqmap))
    for det, txid in zip(detections, txids):
        det.txid = txid


def identify_duplicates(detections):
    """Mask that drops the weaker of two detections of one transmitter in adjacent blocks, and unidentified
    ones (identify.py:136-166: a burst straddling two blocks is detected in both)."""
    array = toads_data.toads_array(detections, with_ids=True)
    idx = np.argsort(array[["rxid", "txid", "block", "timestamp"]])
    cur = array[idx]
    prev = np.roll(cur, 1)
    next_ = np.roll(cur, -1)
    mask_unidentified = (cur["txid"] == UNIDENTIFIED)
    mask_prev = ((cur["block"] == prev["block"] + 1) & (cur["energy"] < prev["energy"]))
    mask_next = ((cur["block"] == next_["block"] - 1) & (cur["energy"] < next_["energy"]))
    mask = ~(mask_prev | mask_next | mask_unidentified)
    return mask[np.argsort(idx)]


def filter_duplicates(detections):
    """Detections without duplicates / unidentified ones, sorted by timestamp (identify.py:169-175)."""
    mask = identify_duplicates(detections)
    filtered = list(itertools.compress(detections, mask))
    filtered.sort(key=lambda d: d.timestamp)
    return filtered


def integrate(detections, freqmap=None, verbose=True):
    """Identify and filter (identify.py:216-220)."""
    identify_transmitters(detections, freqmap, verbose)
    return filter_duplicates(detections)


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


def generate_toads(output, toad_globs, freqmap, verbose=True):
    """identify.py:223-234."""
    detections, filenames = load_toad_files(toad_globs)
    output.write("# source_files: [%s]\n" % (" ".join(filenames)))
    filtered = integrate(detections, freqmap, verbose)
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
    args = parser.parse_args(argv)
    generate_toads(args.output, args.toad_file, load_freqmap(args.map))
    args.output.flush()


if __name__ == "__main__":
    _main()
