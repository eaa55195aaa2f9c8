#!/usr/bin/env python
"""Tiny workloads for compute-sanitizer (memcheck / racecheck / synccheck):

    THRIFTY_B200_MAX_GRID=2 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py [block_len ...]

With the grid capped at 2 CTAs every CTA walks several blocks, so the software pipeline (stage A of block
i+1 next to the fit / tail of earlier blocks, raw-tile ring, mailboxes) is exercised, not just its prologue.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thrifty_b200 import synth  # noqa: E402
from thrifty_b200._native import NativeDetector  # noqa: E402

example = np.load(os.path.join(ROOT, "tests", "golden", "template_example.npy"))
t9, t10 = synth.gold_template(9), synth.gold_template(10)
t11 = np.stack([synth.gold_template(11, i) for i in range(2)])
# (block_len, template(s), history, window, blocks, detector kwargs)
cases = [
    (16384, example, 4920, (7, 110), 14, {}),                          # pruned FFT#1, service warpgroup, 7 blocks per CTA:
                                                                       # the 4-deep fit pipeline in steady state
    (16384, example, 4920, (7, 300), 8, {}),                           # full FFT#1
    (16384, example, 4920, (7, 110), 8, dict(fastdet=True)),           # fastdet semantics (shifted-template table)
    (16384, t11, None, (7, 110), 6, {}),                               # two templates
    (8192, t10, None, (7, 110), 10, {}),                               # 2 CTAs/SM variant, service warpgroup
    (8192, np.stack([synth.gold_template(10, i) for i in range(2)]), None, (7, 300), 8, {}),   # ... two templates, full FFT#1
    (4096, t9, None, (7, 110), 12, {}),                                # 4 CTAs/SM variant
    (4096, t9, None, (1, 2047), 8, dict(fastdet=True)),                # fastdet gather fall-back
    (32768, example, 4920, (7, 110), 7, {}),                           # 2 x 16384 kernel
    (32768, example, 4920, (7, 300), 7, {}),                           # ... with FFT#1 in full (powers parked in the scratch)
    (32768, t11, None, (7, 110), 6, {}),                               # ... two templates (E', O' and B parked per template)
    (32768, example, 4920, (7, 110), 4, dict(generic_kernel=True)),    # global-scratch variant
]
only = [int(a) for a in sys.argv[1:]]                                 # optional: block lengths to run
cases.append((16384, example, 4920, (7, 110), 12, dict(carrier_len=600)))   # N/W = 27: every fit takes the full lmdif path
for n, tpl, hist, win, nblk, kw in cases:
    if only and n not in only:
        continue
    kw = dict(kw)
    carrier_len = kw.pop("carrier_len", None)
    tpl0 = tpl[0] if tpl.ndim == 2 else tpl
    hist = hist or len(tpl0) + 6
    raw, _ = synth.make_blocks(nblk, n, hist, tpl0, 0.7, seed=5)
    det = NativeDetector(n, hist, tpl, carrier_len or len(tpl0), win, (0., 15., 0.), (0., 15., 0.), max_batch=16, **kw)
    rec = det.detect_raw(raw)
    print(n, win, kw, det.info()["kernel"], "grid", det.info()["grid"], "carrier", int((rec["flags"] & 1).sum()),
          "detected", int(((rec["flags"] & 2) != 0).sum()), flush=True)
    det.close()

# stage-boundary kernels (thr_sync_batch / thr_soa_batch) and the identify kernels
if not only or 4096 in only:
    from thrifty_b200 import identify
    hist = len(t9) + 6
    raw, _ = synth.make_blocks(10, 4096, hist, t9, 0.8, seed=9)
    det = NativeDetector(4096, hist, t9, len(t9), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=16)
    recs, sfft = det.sync_batch(raw=raw)
    recs2, corr = det.soa_batch(sfft)
    print(4096, "stage kernels: carrier", int((recs["flags"] & 1).sum()), "detected", int(((recs2["flags"] & 2) != 0).sum()), flush=True)
    full = det.detect_raw(raw)[:, 0]
    sel, tx = identify.integrate_records(full, 1000.0 + np.arange(len(full)), 0)
    print(4096, "identify: kept", len(sel), flush=True)
    det.close()
