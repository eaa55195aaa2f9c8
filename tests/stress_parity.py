#!/usr/bin/env python
"""Randomised parity stress: random detector geometries / thresholds / signal levels, CUDA path vs oracle.

    python tests/stress_parity.py [n_configs] [seed]

Prints one line per configuration and a summary of any mismatch (the parity bar of tests/parity_util.py: one bar for
every block).  A carrier-offset mismatch on a geometry whose Dirichlet main lobe is far wider than the 7 fitted bins can be
checked with tests/fit_sensitivity.py: it shows how far the REFERENCE's own value moves under 1 ulp on its inputs."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))      # a checker: lives with the tests, the only users of oracle/

import parity_util as parity  # noqa: E402
from oracle import thrifty_oracle as orc  # noqa: E402
from thrifty_b200 import synth  # noqa: E402
from thrifty_b200._native import NativeDetector  # noqa: E402


SIZES = [1024, 2048, 4096, 8192, 16384]      # `... <n> <seed> big` adds 32768 (2 x 16384 kernel and its fall-back)


def random_config(rng):
    n = int(rng.choice(SIZES))
    bits = {1024: 7, 2048: 8, 4096: 9, 8192: 10, 16384: 11, 32768: 11}[n]
    full = synth.gold_template(bits)
    tlen = int(rng.integers(max(64, len(full) // 4), min(len(full), n - 64)))
    tpl = full[:tlen]
    hist = int(rng.integers(tlen - 1, min(n - 1, tlen + n // 3)))
    kind = rng.integers(0, 4)
    if kind == 0:
        lo = int(rng.integers(3, 40)); window = (lo, lo + int(rng.integers(10, 100)))
    elif kind == 1:
        hi = -int(rng.integers(5, 40)); window = (hi - int(rng.integers(10, 100)), hi)
    elif kind == 2:
        lo = int(rng.integers(100, n // 2 - 200)); window = (lo, lo + int(rng.integers(10, 300)))
    else:
        window = (int(rng.integers(5, 20)), int(rng.integers(200, n // 2 - 1)))
    cth = (float(rng.choice([0., 1., 50.])), float(rng.uniform(5, 25)), float(rng.choice([0., 0., 1.5])))
    kth = (float(rng.choice([0., 0.5])), float(rng.uniform(5, 25)), float(rng.choice([0., 0., 2.0])))
    carrier_len = int(rng.choice([tlen, tlen, max(32, tlen // 2)]))
    return dict(n=n, tpl=tpl, hist=hist, window=window, cth=cth, kth=kth, carrier_len=carrier_len)


def make_blocks(rng, cfg, nblk):
    n, hist, tpl, (w0, w1) = cfg["n"], cfg["hist"], cfg["tpl"], cfg["window"]
    raws = []
    for _ in range(nblk):
        brng = np.random.default_rng(int(rng.integers(0, 2**31)))
        raw, _ = synth.make_block(brng, n, hist, tpl, 0.75, bin_range=(w0 + 0.6, w1 - 0.6))
        mode = rng.integers(0, 6)
        if mode == 0:       # weak signal: scale towards the noise floor (re-quantise around mid-scale)
            x = (raw.astype(np.float32) - 127.4) * float(rng.uniform(0.05, 0.4)) + 127.4
            raw = np.clip(x, 0, 255).astype(np.uint8)
        elif mode == 1:     # hard clipping
            x = (raw.astype(np.float32) - 127.4) * float(rng.uniform(2, 6)) + 127.4
            raw = np.clip(x, 0, 255).astype(np.uint8)
        raws.append(raw)
    return np.stack(raws)


def main_fastdet(n_cfg, seed):
    """Same, for the fastdet-semantics kernels against oracle/fastdet_oracle.py."""
    from oracle import fastdet_oracle as fo
    rng = np.random.default_rng(seed)
    bad = 0
    for c in range(n_cfg):
        cfg = random_config(rng)
        nblk = 32 if cfg["n"] >= 16384 else 64
        raw = make_blocks(rng, cfg, nblk)
        thresh, kthresh = (cfg["cth"][0] * 100, cfg["cth"][1]), (cfg["kth"][0] * 100, cfg["kth"][1])
        tpl32 = np.asarray(cfg["tpl"], dtype=np.float32)
        tag = "fastdet cfg %d: N=%d L=%d H=%d window=%s t=%s u=%s" % (
            c, cfg["n"], len(tpl32), cfg["hist"], cfg["window"], thresh, kthresh)
        try:
            with np.errstate(all="ignore"):
                ref = fo.detect_blocks(cfg["n"], cfg["hist"], thresh, cfg["window"], tpl32, kthresh, raw)
            det = NativeDetector(cfg["n"], cfg["hist"], tpl32.astype(np.float64), len(tpl32), cfg["window"],
                                 (thresh[0], thresh[1], 0.0), (kthresh[0], kthresh[1], 0.0), max_batch=nblk, fastdet=True)
            got = det.detect_raw(raw)[:, 0]
            det.close()
            stats = parity.compare_fastdet(got, ref, what=tag)
            print("ok  ", tag, {k: (round(v, 7) if isinstance(v, float) else v) for k, v in stats.items()}, flush=True)
        except Exception as e:      # noqa: BLE001
            bad += 1
            print("FAIL", tag, "\n    ", str(e).replace("\n", " ")[:600], flush=True)
            if not isinstance(e, AssertionError):
                traceback.print_exc()
    print("fastdet configs", n_cfg, "failed", bad)
    return bad


def main(n_cfg=None, seed=None):
    if n_cfg is None:
        n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    bad = 0
    for c in range(n_cfg):
        cfg = random_config(rng)
        nblk = 24 if cfg["n"] > 16384 else (48 if cfg["n"] >= 16384 else 64)
        raw = make_blocks(rng, cfg, nblk)
        tag = "cfg %d: N=%d L=%d H=%d W=%d window=%s cth=%s kth=%s" % (
            c, cfg["n"], len(cfg["tpl"]), cfg["hist"], cfg["carrier_len"], cfg["window"], cfg["cth"], cfg["kth"])
        try:
            st = orc.DetectorSettings(cfg["n"], cfg["hist"], cfg["carrier_len"], cfg["cth"], cfg["window"], cfg["tpl"],
                                      cfg["kth"])
            with np.errstate(all="ignore"):
                ref = orc.detect_blocks(st, raw)
            det = NativeDetector(cfg["n"], cfg["hist"], cfg["tpl"], cfg["carrier_len"], cfg["window"], cfg["cth"],
                                 cfg["kth"], max_batch=nblk)
            got = det.detect_raw(raw)[:, 0]
            kern = det.info()["kernel"]
            det.close()
            # ill-conditioned Dirichlet fits (main lobe much wider than the 7 fitted bins) are sensitive to the
            # last bits of the magnitudes: widen the carrier-offset bar by 10 sigma of that sensitivity
            stats = parity.compare_records(got, ref, what=tag)
            print("ok  ", tag, kern, {k: (round(v, 7) if isinstance(v, float) else v) for k, v in stats.items()}, flush=True)
        except Exception as e:      # noqa: BLE001
            bad += 1
            print("FAIL", tag, "\n    ", str(e).replace("\n", " ")[:600], flush=True)
            if not isinstance(e, AssertionError):
                traceback.print_exc()
    print("configs", n_cfg, "failed", bad)
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "big":
        SIZES[:] = [16384, 32768, 32768]
    if len(sys.argv) > 3 and sys.argv[3] == "fastdet":
        sys.exit(1 if main_fastdet(int(sys.argv[1]), int(sys.argv[2])) else 0)
    sys.exit(1 if main() else 0)
