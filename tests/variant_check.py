#!/usr/bin/env python
"""Kernel experiment helper: parity of the N=16384 golden + device-resident timing for the library
selected by THRIFTY_B200_LIB (see tools/variants.sh).  One JSON line per call."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import parity_util as parity  # noqa: E402
import sweep  # noqa: E402
from thrifty_b200._native import NativeDetector  # noqa: E402


def main():
    label = os.environ.get("THRIFTY_B200_LIB", "default")
    ok = True
    msg = ""
    for window in [None, (7, 300)]:           # pruned and full FFT#1 paths
        cfg, raw, block_idx, ref, _ = parity.load_golden("n16384_example")
        if window is not None:
            from oracle import thrifty_oracle as orc
            st = orc.DetectorSettings(cfg["block_len"], cfg["history_len"], len(cfg["template"]), cfg["cthresh"],
                                      window, cfg["template"], cfg["kthresh"])
            ref = orc.detect_blocks(st, raw, block_idx)
        det = NativeDetector(cfg["block_len"], cfg["history_len"], cfg["template"], len(cfg["template"]),
                             window or cfg["window"], cfg["cthresh"], cfg["kthresh"], device=0, max_batch=64)
        got = det.detect_raw(raw, block_idx)[:, 0]
        det.close()
        try:
            st = parity.compare_records(got, ref, what=label)
            msg += " %s" % (st,)
        except AssertionError as e:
            ok = False
            msg += " PARITY FAIL: %s" % (str(e)[:300],)
    print(json.dumps(dict(variant=label, parity_ok=ok, detail=msg)), flush=True)
    example = np.load(os.path.join(ROOT, "tests", "golden", "template_example.npy"))
    sweep.run(16384, example, 4920, 4096, 1.0, steps=128, warmup=8, label=label + " zoom")
    sweep.run(16384, example, 4920, 4096, 1.0, steps=64, warmup=8, window=(7, 300), label=label + " full")


if __name__ == "__main__":
    main()
