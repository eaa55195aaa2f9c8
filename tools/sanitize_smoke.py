#!/usr/bin/env python
"""Tiny workloads for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thrifty_b200 import synth  # noqa: E402
from thrifty_b200._native import NativeDetector  # noqa: E402

example = np.load(os.path.join(ROOT, "tests", "golden", "template_example.npy"))
cases = [(16384, example, 4920, (7, 110), 10), (16384, example, 4920, (7, 300), 6),
         (4096, synth.gold_template(9), None, (7, 110), 12), (32768, example, 4920, (7, 110), 3)]
for n, tpl, hist, win, nblk in cases:
    hist = hist or len(tpl) + 6
    raw, _ = synth.make_blocks(nblk, n, hist, tpl, 0.7, seed=5)
    det = NativeDetector(n, hist, tpl, len(tpl), win, (0., 15., 0.), (0., 15., 0.), max_batch=4)
    rec = det.detect_raw(raw)
    print(n, win, "carrier", int((rec["flags"] & 1).sum()), "detected", int(((rec["flags"] & 2) != 0).sum()), flush=True)
    det.close()
