#!/usr/bin/env python
"""Golden records for BASELINE config 5: four Gold-11 templates correlated jointly at block_len = 16384.

The reference has no multi-template detector; SURVEY.md 8d defines parity for this configuration as FOUR INDEPENDENT
runs of the reference's own ``thrifty.detect.Detector`` (one per template) on the same blocks.  This script (build
container only: imports /root/reference) does exactly that, checks the oracle restatement against it and stores the
reference's outputs as tests/golden/detect_n16384_gold11x4.npz.  Inputs are regenerated from seeds by
tests/parity_util.make_multi_blocks (CRC32 stored).

    python oracle/make_golden_multi.py [name ...]      # all of CONFIGS, or only the named ones
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from thrifty.detect import Detector as RefDetector, DetectorSettings as RefSettings  # noqa: E402
from thrifty.block_data import raw_to_complex as ref_raw_to_complex  # noqa: E402
from thrifty.signal_utils import Signal  # noqa: E402

from oracle import thrifty_oracle as orc  # noqa: E402
from oracle.make_golden import compare  # noqa: E402
from thrifty_b200 import synth  # noqa: E402
import parity_util  # noqa: E402

# name -> (block_len, Gold bits, template indices, blocks, window, carrier / correlation thresholds, p_signal, seed)
CONFIGS = {
    # BASELINE config 5; > 2 x 148 blocks: every persistent CTA of a B200 walks more than one block
    "n16384_gold11x4": (16384, 11, (0, 1, 2, 3), 320, (7, 110), (0., 15., 0.), (0., 15., 0.), 0.8, synth.SEED0 + 555),
    # block_len 32768 (transformed as two halves, detect_kernel_2x.cuh) with a window too wide for the pruned FFT#1
    "n32768_gold11x3": (32768, 11, (0, 1, 2), 48, (7, 300), (0., 15., 0.), (0., 15., 0.), 0.8, synth.SEED0 + 556),
}


def main(name):
    NAME = name
    N, BITS, IDX, N_BLOCKS, WINDOW, CTH, KTH, P_SIGNAL, SEED = CONFIGS[name]
    tpls = np.stack([synth.gold_template(BITS, i) for i in IDX])
    hist = tpls.shape[1] + 6                 # 4914 + 6 = 4920, as example/detector.cfg
    assert tpls.shape[1] == 4914
    raw, which = parity_util.make_multi_blocks(N_BLOCKS, N, hist, tpls, P_SIGNAL, SEED)
    recs = np.zeros((len(IDX), N_BLOCKS), dtype=orc.RECORD_DTYPE)
    for t in range(len(IDX)):
        st = RefSettings(block_len=N, history_len=hist, carrier_len=tpls.shape[1], carrier_thresh=CTH,
                         carrier_window=WINDOW, template=tpls[t], corr_thresh=KTH)
        det = RefDetector(st, rxid=0)
        for i in range(N_BLOCKS):
            detected, res = det.detect(1000.0 + i * 0.0047767, 10 + 3 * i, Signal(ref_raw_to_complex(raw[i])))
            recs[t, i] = orc.result_to_row(orc.OracleResult(detected, res.timestamp, res.block, res.soa,
                                                            res.carrier_info, res.corr_info, 0))
        ost = orc.DetectorSettings(N, hist, tpls.shape[1], CTH, WINDOW, tpls[t], KTH)
        orows = orc.detect_blocks(ost, raw, 10 + 3 * np.arange(N_BLOCKS))
        compare(recs[t], orows, "%s template %d" % (NAME, t))
        for f in ("carrier_margin", "corr_margin"):      # peak / threshold, to recognise knife-edge decisions
            recs[t][f] = orows[f]
        own = recs[t]["corr_detected"][which == t].sum()
        other = recs[t]["corr_detected"][which != t].sum()
        print("template %d: carrier=%d detected=%d (own bursts %d, other templates' bursts %d)  oracle==reference OK"
              % (t, recs[t]["carrier_detected"].sum(), recs[t]["corr_detected"].sum(), own, other))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "detect_%s.npz" % NAME), records=recs,
                        raw_crc32=np.uint32(zlib.crc32(raw.tobytes())), block_len=N, history_len=hist,
                        gold_bits=BITS, gold_idx=np.array(IDX), window=np.array(WINDOW), n_blocks=N_BLOCKS,
                        p_signal=P_SIGNAL, cthresh=np.array(CTH), kthresh=np.array(KTH), seed=SEED, which=which)


if __name__ == "__main__":
    for cfg_name in (sys.argv[1:] or list(CONFIGS)):
        main(cfg_name)
