#!/usr/bin/env python
"""How fast is the oracle port next to the reference it restates?  (build container only: imports /root/reference)

bench.py's CPU legs time oracle/thrifty_oracle.py because the reference itself cannot travel to the GPU box.  This script
times both on the same blocks of the headline workload, single thread, best of 3, and writes
profiles/port_vs_reference.json, which bench.py quotes in `cpu_baseline.sample` / `cpu_baseline.port_vs_reference`.

    OMP_NUM_THREADS=1 python oracle/port_vs_reference.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from thrifty.detect import Detector as RefDetector, DetectorSettings as RefSettings  # noqa: E402
from thrifty.block_data import raw_to_complex as ref_raw_to_complex  # noqa: E402
from thrifty.signal_utils import Signal  # noqa: E402

from oracle import thrifty_oracle as orc  # noqa: E402
from thrifty_b200 import synth  # noqa: E402

N, H, NBLK = 16384, 4920, 192


def main():
    import scipy
    tpl = np.load(os.path.join(ROOT, "tests", "golden", "template_example.npy"))
    raw, _ = synth.make_blocks(NBLK, N, H, tpl, 1.0, seed=synth.SEED0)
    ref = RefDetector(RefSettings(N, H, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.)), rxid=0)
    port = orc.Detector(orc.DetectorSettings(N, H, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.)), rxid=0)
    t_ref, t_port = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        for i in range(NBLK):
            ref.detect(0.0, i, Signal(ref_raw_to_complex(raw[i])))
        t_ref.append((time.perf_counter() - t0) / NBLK)
        t0 = time.perf_counter()
        for i in range(NBLK):
            port.detect_raw(0.0, i, raw[i])
        t_port.append((time.perf_counter() - t0) / NBLK)
    out = {"reference_ms_per_block": round(min(t_ref) * 1e3, 4), "port_ms_per_block": round(min(t_port) * 1e3, 4),
           "reference_over_port": round(min(t_ref) / min(t_port), 4), "blocks": NBLK, "block_len": N,
           "workload": "headline (example template, window 7-110, 15*snr, every block a burst)", "threads": 1,
           "host": "build container, %d cpus" % (os.cpu_count() or 0), "numpy": np.__version__, "scipy": scipy.__version__,
           "note": "the port skips the reference's Signal bookkeeping (cached properties on an ndarray subclass); same arithmetic"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "port_vs_reference.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
