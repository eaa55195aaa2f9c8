#!/usr/bin/env python
"""Golden vectors for thrifty_b200/identify.py from the reference's own functions.

`thrifty.identify.detect_transmitter_windows` and `identify_duplicates` (identify.py:26-77,136-166) run
unmodified under Python 3 (the `iteritems` users `auto_classify_transmitters`, `classify_transmitters` and
`load_freqmap` do not; they are pinned through these two plus hand-checked cases in tests/test_identify.py).
Runs only in the build container (needs /root/reference):   python oracle/make_golden_identify.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from thrifty import identify as ref_identify  # noqa: E402
from thrifty import toads_data as ref_toads  # noqa: E402


def make_detections(rng, n, n_rx=2, tx_bins=(20, 55, 90)):
    dets = []
    for i in range(n):
        rx = int(rng.integers(0, n_rx))
        tx = int(rng.integers(0, len(tx_bins)))
        block = int(rng.integers(0, n // 3))
        cbin = int(tx_bins[tx] + rng.integers(-2, 3))
        dets.append(ref_toads.DetectionResult(
            timestamp=1000.0 + block * 0.0047767 + i * 1e-6, block=block, soa=11464.0 * block + rng.uniform(3, 11000),
            carrier_info=ref_toads.CarrierSyncInfo(cbin, float(rng.uniform(-0.5, 0.5)), float(rng.uniform(500, 5000)), float(rng.uniform(5, 20))),
            corr_info=ref_toads.CorrDetectionInfo(int(rng.integers(3, 11000)), float(rng.uniform(-0.5, 0.5)), float(rng.uniform(100, 900)), float(rng.uniform(1, 9))),
            rxid=rx, txid=tx if rng.random() > 0.05 else -1))
    return dets


def main():
    rng = np.random.default_rng(20161125)
    cases = {}
    for c in range(4):
        dets = make_detections(rng, 300 + 100 * c)
        # both implementations see exactly what the text file holds
        lines = [d.serialize() for d in dets]
        dets = [ref_toads.DetectionResult.deserialize(l, with_rxid=True, with_txid=True) for l in lines]
        mask = ref_identify.identify_duplicates(dets)
        bins = np.array([d.carrier_info.bin for d in dets if d.rxid == 0])
        edges = ref_identify.detect_transmitter_windows(bins)
        cases["lines_%d" % c] = np.array([d.serialize() for d in dets])
        cases["mask_%d" % c] = mask
        cases["bins_%d" % c] = bins
        cases["edges_%d" % c] = np.asarray(edges)
        print("case %d: %d detections, %d kept, edges %s" % (c, len(dets), int(mask.sum()), list(edges)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "identify_cases.npz"), n_cases=4, **cases)


if __name__ == "__main__":
    main()
