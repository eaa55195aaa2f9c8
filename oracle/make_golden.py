#!/usr/bin/env python
"""Pin the oracle against the REAL reference and write tests/golden/.

Runs only in the build container (needs /root/reference).  For a set of seeded
synthetic block configurations it
  1. runs the reference's own ``thrifty.detect.Detector.detect`` (unmodified
     arithmetic: carrier_detect / carrier_sync / soa_estimator / signal_utils),
  2. asserts that ``oracle/thrifty_oracle.py`` reproduces every field, and
  3. stores the reference outputs as ``tests/golden/detect_<name>.npz``
     (inputs are regenerated from seeds by ``thrifty_b200.synth``; a CRC32 of
     the raw bytes is stored to catch generator drift).
It also copies the canonical template ``example/template.npy`` into
``tests/golden/template_example.npy`` (a data fixture, 39 KB).

    python oracle/make_golden.py [name ...]      # all configurations, or only the named ones
"""

import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from thrifty.detect import Detector as RefDetector, DetectorSettings as RefSettings  # noqa: E402
from thrifty.block_data import raw_to_complex as ref_raw_to_complex  # noqa: E402
from thrifty.signal_utils import Signal  # noqa: E402

from oracle import thrifty_oracle as orc  # noqa: E402
from thrifty_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def configs():
    tpl = np.load(os.path.join(REF, "example", "template.npy"))
    yield dict(name="n16384_example", block_len=16384, history_len=4920, template=tpl,
               template_id="example", window=(7, 110), n_blocks=64, p_signal=0.6,
               bin_range=(8.0, 109.0), cthresh=(0., 15., 0.), kthresh=(0., 15., 0.))
    t10 = synth.gold_template(10)
    yield dict(name="n8192_gold10", block_len=8192, history_len=len(t10) + 6, template=t10,
               template_id="gold10_0", window=(7, 110), n_blocks=48, p_signal=0.6,
               bin_range=(8.0, 109.0), cthresh=(0., 15., 0.), kthresh=(0., 15., 0.))
    t9 = synth.gold_template(9)
    yield dict(name="n4096_gold9", block_len=4096, history_len=len(t9) + 6, template=t9,
               template_id="gold9_0", window=(7, 110), n_blocks=48, p_signal=0.6,
               bin_range=(8.0, 109.0), cthresh=(0., 15., 0.), kthresh=(0., 15., 0.))
    # zero-straddling window + constant/stddev threshold terms.  Carriers within 4 bins
    # of DC are excluded: the reference raises IndexError there (fft_mag[peak_idx + xdata]
    # runs past N, carrier_sync.py:187, after the '>' quirk at carrier_detect.py:151), and
    # every block carries a burst because in a noise-only block the truncation DC bias at
    # bin 0 wins the window and triggers the same IndexError.
    yield dict(name="n4096_gold9_wrapwin_std", block_len=4096, history_len=len(t9) + 6, template=t9,
               template_id="gold9_0", window=(-40, 60), n_blocks=32, p_signal=1.0,
               bin_range=(-38.0, -5.0, 2.0, 58.0), cthresh=(1., 12., 2.), kthresh=(0.5, 10., 3.))
    # unmodulated tone bursts (no spreading code): carrier found, correlation peak rejected
    yield dict(name="n4096_gold9_tone", block_len=4096, history_len=len(t9) + 6, template=t9,
               gen_template=np.ones(len(t9)),
               template_id="gold9_0", window=(7, 110), n_blocks=24, p_signal=0.8,
               bin_range=(8.0, 109.0), cthresh=(0., 15., 0.), kthresh=(0., 15., 0.))
    yield dict(name="n32768_example", block_len=32768, history_len=4920, template=tpl,
               template_id="example", window=(7, 110), n_blocks=16, p_signal=0.7,
               bin_range=(8.0, 109.0), cthresh=(0., 15., 0.), kthresh=(0., 15., 0.))
    # carrier window wider than the 128-bin zoom band + constant / stddev threshold terms: FFT#1 in full at the block
    # length that is transformed as two halves (detect_kernel_2x.cuh)
    yield dict(name="n32768_example_wide_std", block_len=32768, history_len=4920, template=tpl,
               template_id="example", window=(7, 300), n_blocks=16, p_signal=0.7,
               bin_range=(9.0, 298.0), cthresh=(1., 12., 2.), kthresh=(0.5, 10., 3.))


def run_reference(cfg, raw):
    st = RefSettings(block_len=cfg["block_len"], history_len=cfg["history_len"],
                     carrier_len=len(cfg["template"]), carrier_thresh=cfg["cthresh"],
                     carrier_window=cfg["window"], template=cfg["template"],
                     corr_thresh=cfg["kthresh"])
    det = RefDetector(st, rxid=0)
    rows = np.zeros(len(raw), dtype=orc.RECORD_DTYPE)
    lines = []
    for i in range(len(raw)):
        detected, res = det.detect(1000.0 + i * 0.0047767, 10 + 3 * i,
                                   Signal(ref_raw_to_complex(raw[i])))
        rows[i] = orc.result_to_row(orc.OracleResult(
            detected, res.timestamp, res.block, res.soa, res.carrier_info, res.corr_info, 0))
        lines.append(res.serialize() if detected else "")
    return rows, lines


def run_oracle(cfg, raw):
    st = orc.DetectorSettings(block_len=cfg["block_len"], history_len=cfg["history_len"],
                              carrier_len=len(cfg["template"]), carrier_thresh=cfg["cthresh"],
                              carrier_window=cfg["window"], template=cfg["template"],
                              corr_thresh=cfg["kthresh"])
    det = orc.Detector(st, rxid=0)
    rows = np.zeros(len(raw), dtype=orc.RECORD_DTYPE)
    lines = []
    for i in range(len(raw)):
        res = det.detect_raw(1000.0 + i * 0.0047767, 10 + 3 * i, raw[i])
        rows[i] = orc.result_to_row(res)
        lines.append(orc.serialize(res) if res.detected else "")
    return rows, lines


def compare(a, b, name):
    for f in a.dtype.names:
        if f.endswith("margin"):
            continue
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            np.testing.assert_allclose(x, y, rtol=1e-10, atol=1e-10, equal_nan=True,
                                       err_msg="%s field %s" % (name, f))
        else:
            np.testing.assert_array_equal(x, y, err_msg="%s field %s" % (name, f))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    tpl = np.load(os.path.join(REF, "example", "template.npy"))
    np.save(os.path.join(GOLDEN, "template_example.npy"), tpl)
    only = sys.argv[1:]
    for cfg in configs():
        if only and cfg["name"] not in only:
            continue
        raw, truths = synth.make_blocks(cfg["n_blocks"], cfg["block_len"], cfg["history_len"],
                                        cfg.get("gen_template", cfg["template"]),
                                        cfg["p_signal"], seed=synth.SEED0,
                                        bin_range=cfg["bin_range"])
        ref_rows, ref_lines = run_reference(cfg, raw)
        orc_rows, orc_lines = run_oracle(cfg, raw)
        compare(ref_rows, orc_rows, cfg["name"])
        # .toad lines: same fields; floats may differ in the last ulp because numpy's
        # SIMD complex multiply rounds differently depending on buffer alignment.
        for la, lb in zip(ref_lines, orc_lines):
            assert (la == "") == (lb == ""), cfg["name"]
            if la:
                np.testing.assert_allclose([float(v) for v in la.split()],
                                           [float(v) for v in lb.split()], rtol=1e-10)
        np.savez_compressed(
            os.path.join(GOLDEN, "detect_%s.npz" % cfg["name"]),
            records=ref_rows, toad_lines=np.array(ref_lines),
            raw_crc32=np.uint32(zlib.crc32(raw.tobytes())),
            block_len=cfg["block_len"], history_len=cfg["history_len"],
            template_id=cfg["template_id"], template_len=len(cfg["template"]),
            gen_template_id="ones" if "gen_template" in cfg else cfg["template_id"],
            window=np.array(cfg["window"]), n_blocks=cfg["n_blocks"],
            p_signal=cfg["p_signal"], bin_range=np.array(cfg["bin_range"]),
            cthresh=np.array(cfg["cthresh"]), kthresh=np.array(cfg["kthresh"]),
            seed=synth.SEED0, truth_signal=np.array([t["signal"] for t in truths]),
            truth_pos=np.array([t["pos"] for t in truths]),
            truth_bin=np.array([t["bin"] for t in truths]))
        ncar = int(ref_rows["carrier_detected"].sum())
        ndet = int(ref_rows["corr_detected"].sum())
        nsig = sum(t["signal"] for t in truths)
        print("%-28s blocks=%d signal=%d carrier=%d detected=%d  oracle==reference OK"
              % (cfg["name"], cfg["n_blocks"], nsig, ncar, ndet))

    if only:
        return
    # one block with full intermediate arrays (yield_data=True) at N=4096
    cfg = [c for c in configs() if c["name"] == "n4096_gold9"][0]
    raw, truths = synth.make_blocks(8, cfg["block_len"], cfg["history_len"], cfg["template"],
                                    1.0, seed=synth.SEED0 + 1000)
    st = RefSettings(block_len=cfg["block_len"], history_len=cfg["history_len"],
                     carrier_len=len(cfg["template"]), carrier_thresh=cfg["cthresh"],
                     carrier_window=cfg["window"], template=cfg["template"],
                     corr_thresh=cfg["kthresh"])
    det = RefDetector(st, rxid=0, yield_data=True)
    detected, res, sfft, corr = det.detect(0.0, 5, Signal(ref_raw_to_complex(raw[0])))
    assert detected
    ores, osfft, ocorr = orc.Detector(orc.DetectorSettings(*st), 0).detect_raw(0.0, 5, raw[0], True)
    np.testing.assert_allclose(osfft, np.asarray(sfft), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ocorr, np.asarray(corr), rtol=1e-9, atol=1e-9)
    np.savez_compressed(os.path.join(GOLDEN, "arrays_n4096_gold9.npz"),
                        shifted_fft=np.asarray(sfft), corr=np.asarray(corr),
                        raw_crc32=np.uint32(zlib.crc32(raw[0].tobytes())),
                        seed=synth.SEED0 + 1000, soa=res.soa)
    print("arrays_n4096_gold9: yield_data arrays pinned")


if __name__ == "__main__":
    main()
