"""Host-side `identify` (thrifty/identify.py) against goldens produced by the reference's own functions."""
import io
import os

import numpy as np
import pytest

from thrifty_b200 import identify, toads_data

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "identify_cases.npz")


def _cases():
    g = np.load(GOLDEN)
    for c in range(int(g["n_cases"])):
        dets = [toads_data.DetectionResult.deserialize(str(l), with_rxid=True, with_txid=True) for l in g["lines_%d" % c]]
        yield dets, g["mask_%d" % c], g["bins_%d" % c], g["edges_%d" % c]


def test_duplicates_mask_matches_reference():
    # identify.py:136-166
    for dets, mask, _, _ in _cases():
        assert np.array_equal(identify.identify_duplicates(dets), mask)
        kept = identify.filter_duplicates(dets)
        assert len(kept) == int(mask.sum())
        assert all(a.timestamp <= b.timestamp for a, b in zip(kept, kept[1:]))
        assert all(d.txid != -1 for d in kept)


def test_transmitter_windows_match_reference():
    # identify.py:26-77
    for _, _, bins, edges in _cases():
        assert np.array_equal(identify.detect_transmitter_windows(bins), edges)


def test_auto_classification_groups_by_carrier_bin():
    # identify.py:80-103: three transmitters around bins 20 / 55 / 90 at each receiver
    dets, _, _, _ = next(_cases())
    truth = [d.txid for d in dets]
    identify.identify_transmitters(dets, None, verbose=False)
    for d, t in zip(dets, truth):
        if t != -1:
            assert d.txid == t


def test_freqmap_classification_and_toads_file(tmp_path):
    # identify.py:106-118,191-234
    freqmap = identify.load_freqmap(io.StringIO("0: 17 - 23\n1: 52 - 58\n2: 87 - 93.5\n@0: 0\n@1: 0.25\n"))
    assert freqmap[1][2] == (87.25, 93.75) and set(freqmap) == {0, 1}
    dets, _, _, _ = next(_cases())
    by_rx = {0: [], 1: []}
    for d in dets:
        by_rx[d.rxid].append(d)
    for rx, lst in by_rx.items():
        with open(str(tmp_path / ("rx%d.toad" % rx)), "w") as f:
            for d in lst:
                d.txid = None
                f.write(d.serialize() + "\n")
    out = io.StringIO()
    kept = identify.generate_toads(out, [str(tmp_path / "rx*.toad")], freqmap, verbose=False)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("# source_files: [") and len(lines) == len(kept) + 1
    back = toads_data.load_toads(io.StringIO(out.getvalue()))
    assert [d.txid for d in back] == [d.txid for d in kept]
    for d in back:
        lo, hi = freqmap[d.rxid][d.txid]
        assert lo <= d.carrier_info.bin + d.carrier_info.offset <= hi
