import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import parity_util as parity
from thrifty_b200._native import NativeDetector
cfg, tpls, raw, block_idx, ref, which = parity.load_multi_golden()
for nt in (4, 1):
    det = NativeDetector(cfg["block_len"], cfg["history_len"], tpls[:nt] if nt > 1 else tpls[0], tpls.shape[1], cfg["window"], cfg["cthresh"], cfg["kthresh"], max_batch=512)
    base = det.detect_raw(raw, block_idx)
    bad = 0
    for rep in range(30):
        got = det.detect_raw(raw, block_idx)
        if got.tobytes() != base.tobytes():
            bad += 1
            diff = np.nonzero((got.view(np.uint8).reshape(len(raw), nt, 64) != base.view(np.uint8).reshape(len(raw), nt, 64)).any(axis=2))
            print("nt", nt, "rep", rep, "differs at (block, tpl):", list(zip(diff[0][:8].tolist(), diff[1][:8].tolist())), "of", len(diff[0]))
            i, t = int(diff[0][0]), int(diff[1][0])
            print("   base", base[i, t]); print("   got ", got[i, t])
    print("templates", nt, "nondeterministic runs:", bad, "of 30", det.info()["kernel"])
    det.close()
