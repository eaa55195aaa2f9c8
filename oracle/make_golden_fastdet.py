#!/usr/bin/env python
"""Pin the native-path (fastdet) oracle against the reference's own compiled sources and write
tests/golden/fastdet_*.npz.

Runs only in the build container: needs oracle/_ref/libfastdet_ref.so (``make -C oracle``, which
compiles /root/reference/fastcard + fastdet in place against the FFTW/VOLK stand-ins).  For each
seeded configuration it
  1. writes the synthetic blocks as a `.card` file (and, for the stream case, as a raw uint8 file),
  2. runs the compiled reference over the file exactly like fastdet's main loop,
  3. asserts that oracle/fastdet_oracle.py (NumPy restatement) agrees -- flags and indices exactly,
     powers to 5e-5 relative (different float32 FFTs), offsets to 2e-4 --
  4. stores the compiled reference's records (inputs are regenerated from seeds; CRC32 stored).

    make -C oracle && python oracle/make_golden_fastdet.py
"""

import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fastdet_oracle as fo  # noqa: E402
from thrifty_b200 import block_data, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def configs():
    tpl = np.load(os.path.join(GOLDEN, "template_example.npy"))
    yield dict(name="n16384_example", block_len=16384, history_len=4920, template=tpl, template_id="example",
               window=(7, 110), n_blocks=64, p_signal=0.6, bin_range=(8.0, 109.0), thresh=(0., 15.),
               corr_thresh=(0., 15.), mode="card")
    t10 = synth.gold_template(10)
    yield dict(name="n8192_gold10_const", block_len=8192, history_len=len(t10) + 6, template=t10,
               template_id="gold10_0", window=(7, 110), n_blocks=48, p_signal=0.6, bin_range=(8.0, 109.0),
               thresh=(100., 2.), corr_thresh=(50., 12.), mode="card")
    t9 = synth.gold_template(9)
    # negative-frequency window (cardet_normalize_window maps it to [N-110, N-7])
    yield dict(name="n4096_gold9_negwin", block_len=4096, history_len=len(t9) + 6, template=t9,
               template_id="gold9_0", window=(-110, -7), n_blocks=48, p_signal=0.6, bin_range=(-109.0, -8.0),
               thresh=(0., 15.), corr_thresh=(0., 15.), mode="card")
    # raw stream through the reference's raw_reader (history overlap, first block padded with 127/0)
    yield dict(name="n4096_gold9_stream", block_len=4096, history_len=len(t9) + 6, template=t9,
               template_id="gold9_0", window=(7, 110), n_blocks=48, p_signal=0.6, bin_range=(8.0, 109.0),
               thresh=(0., 15.), corr_thresh=(0., 15.), mode="raw")
    yield dict(name="n32768_example", block_len=32768, history_len=4920, template=tpl, template_id="example",
               window=(7, 110), n_blocks=16, p_signal=0.6, bin_range=(8.0, 109.0), thresh=(0., 15.),
               corr_thresh=(0., 15.), mode="card")


def stream_blocks(raw, block_len, history_len):
    """Blocks the reference's raw_reader forms from the concatenated NEW parts of `raw`
    (fastcard/raw_reader.c:15-46; initial history = uint16 127 per sample, reader.c:56-59)."""
    new = block_len - history_len
    stream = np.concatenate([r[2 * history_len:] for r in raw])
    blocks = np.zeros((len(raw), 2 * block_len), dtype=np.uint8)
    cur = np.zeros(2 * block_len, dtype=np.uint8)
    cur[0::2] = 127
    for b in range(len(raw)):
        cur = np.concatenate([cur[2 * new:], stream[2 * new * b:2 * new * (b + 1)]])
        blocks[b] = cur
    return stream, blocks


def check(ref, mine, what):
    assert len(ref) == len(mine), what
    for i in range(len(ref)):
        r, m = ref[i], mine[i]
        tag = "%s block %d" % (what, i)
        assert r["block_idx"] == m["block_idx"], tag
        assert r["carrier_detected"] == m["carrier_detected"], tag
        if not r["carrier_detected"]:
            continue
        assert r["carrier_argmax"] == m["carrier_argmax"], tag
        assert r["corr_peak_idx"] == m["corr_peak_idx"], tag
        margin = abs(r["corr_peak_power"] / r["corr_threshold"] - 1) if r["corr_threshold"] > 0 else 1
        if margin > 1e-3:
            assert r["corr_detected"] == m["corr_detected"], tag
        for f in ("carrier_max", "carrier_noise", "fft_sum", "corr_peak_power", "corr_noise_power"):
            assert abs(r[f] - m[f]) <= 5e-5 * abs(r[f]) + 1e-12, (tag, f, r[f], m[f])
        assert abs(r["corr_offset"] - m["corr_offset"]) <= 2e-4, tag
        assert abs(r["carrier_offset"] - m["carrier_offset"]) <= 2e-4, tag
        assert abs(r["soa"] - m["soa"]) <= 2e-4, tag


def main():
    if not fo.have_reference():
        raise SystemExit("build oracle/_ref first: make -C oracle")
    for cfg in configs():
        n, h = cfg["block_len"], cfg["history_len"]
        raw, _ = synth.make_blocks(cfg["n_blocks"], n, h, cfg["template"], cfg["p_signal"],
                                   seed=synth.SEED0 + 7000 + n, bin_range=cfg["bin_range"])
        block_idx = 10 + 3 * np.arange(cfg["n_blocks"], dtype=np.int64)
        with tempfile.TemporaryDirectory() as tmp:
            if cfg["mode"] == "card":
                path = os.path.join(tmp, "in.card")
                with open(path, "w") as f:
                    block_data.write_card(f, raw, block_idx)
                blocks = raw
            else:
                stream, blocks = stream_blocks(raw, n, h)
                block_idx = np.arange(cfg["n_blocks"], dtype=np.int64)       # raw_reader counts from 0
                path = os.path.join(tmp, "in.dat")
                stream.tofile(path)
            ref = fo.run_reference(path, cfg["mode"] == "card", n, h, cfg["thresh"], cfg["window"],
                                   cfg["template"], cfg["corr_thresh"])
        mine = fo.detect_blocks(n, h, cfg["thresh"], cfg["window"], cfg["template"], cfg["corr_thresh"],
                                blocks, block_idx)
        check(ref, mine, cfg["name"])
        # rawconv: the reference LUT and the restatement are bit-identical
        assert np.array_equal(fo.reference_rawconv(blocks[0]).view(np.uint32), fo.rawconv(blocks[0]).view(np.uint32))
        ncar, ndet = int(ref["carrier_detected"].sum()), int(ref["corr_detected"].sum())
        print("%-24s blocks %3d carrier %3d detected %3d  (oracle == compiled reference)"
              % (cfg["name"], len(ref), ncar, ndet))
        np.savez_compressed(
            os.path.join(GOLDEN, "fastdet_%s.npz" % cfg["name"]),
            block_len=n, history_len=h, template_id=cfg["template_id"], window=np.array(cfg["window"]),
            n_blocks=cfg["n_blocks"], p_signal=cfg["p_signal"], bin_range=np.array(cfg["bin_range"]),
            thresh=np.array(cfg["thresh"]), corr_thresh=np.array(cfg["corr_thresh"]), mode=cfg["mode"],
            seed=synth.SEED0 + 7000 + n, raw_crc32=np.uint32(zlib.crc32(raw.tobytes())),
            records=ref, toad_lines=np.array([fo.toad_line(r) for r in ref if r["corr_detected"]]))


if __name__ == "__main__":
    main()
