#!/usr/bin/env python
"""One sweep configuration (for ncu captures and A/B runs):  sweep_one.py <block_len> [fastdet|multi|wide]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sweep  # noqa: E402
from thrifty_b200 import synth  # noqa: E402

n = int(sys.argv[1])
fast = len(sys.argv) > 2 and sys.argv[2] == "fastdet"
multi = len(sys.argv) > 2 and sys.argv[2] == "multi"
wide = len(sys.argv) > 2 and sys.argv[2] == "wide"      # carrier window 7-300: FFT#1 cannot be pruned
if n >= 16384:
    tpl, hist = np.load(os.path.join(ROOT, "tests", "golden", "template_example.npy")), 4920
else:
    tpl = synth.gold_template({4096: 9, 8192: 10}[n])
    hist = len(tpl) + 6
if multi:      # four Gold codes of the block length's family
    tpl = np.stack([synth.gold_template({4096: 9, 8192: 10}.get(n, 11), i) for i in range(4)])
    hist = tpl.shape[1] + 6
sweep.run(n, tpl, hist, 2048 if n > 16384 else 4096, 1.0, steps=8, warmup=3,
          label="N=%d%s" % (n, " fastdet" if fast else " 4 templates" if multi else " window 7-300" if wide else ""),
          fastdet=fast, **(dict(window=(7, 300)) if wide else {}))
