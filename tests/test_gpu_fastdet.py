"""GPU parity of the fastdet-semantics kernels (THR_CFG_FASTDET_SEMANTICS) against the goldens produced
by the reference's own compiled native sources, and against the native-path oracle on fresh seeds."""
import numpy as np
import pytest

import parity_util as parity

pytestmark = pytest.mark.gpu


def _detector(cfg, **kw):
    from thrifty_b200._native import NativeDetector
    tpl32 = np.asarray(cfg["template"], dtype=np.float32).astype(np.float64)     # .tpl holds float32
    return NativeDetector(cfg["block_len"], cfg["history_len"], tpl32, len(tpl32), cfg["window"],
                          (cfg["thresh"][0], cfg["thresh"][1], 0.0), (cfg["corr_thresh"][0], cfg["corr_thresh"][1], 0.0),
                          device=0, max_batch=kw.get("max_batch", 64), fastdet=True)


@pytest.mark.parametrize("name", parity.FASTDET_GOLDEN_NAMES)
def test_fastdet_golden(name):
    cfg, raw, block_idx, ref, _, stream = parity.load_fastdet_golden(name)
    det = _detector(cfg)
    assert "fastdet" in det.info()["kernel"]
    got = det.detect_raw(raw, block_idx)[:, 0]
    stats = parity.compare_fastdet(got, ref, what="fastdet/" + name)
    print(name, stats)
    assert stats["carrier"] == int(ref["carrier_detected"].sum())
    if stream is not None:
        # same blocks read in place from the contiguous stream (raw_reader.c semantics: block 0's history
        # is uint16 127 per sample)
        hist = np.zeros(2 * cfg["history_len"], dtype=np.uint8)
        hist[0::2] = 127
        got2 = det.detect_stream(np.concatenate([hist, stream]), 0)[:, 0]
        assert got2.tobytes() == got.tobytes()
    det.close()


@pytest.mark.parametrize("window", [(7, 110), (7, 300), (1, -2), (-200, -20)])
def test_fastdet_fresh_seeds_vs_oracle(window):
    from oracle import fastdet_oracle as fo
    from thrifty_b200 import synth
    tpl = parity.template_by_id("example")
    n, h = 16384, 4920
    neg = window[0] < 0
    raw, _ = synth.make_blocks(96, n, h, tpl, 0.7, seed=777 + abs(window[0]),
                               bin_range=(-190.0, -25.0) if neg else (8.0, 109.0))
    cfg = dict(block_len=n, history_len=h, template=tpl, window=window, thresh=(0., 15.), corr_thresh=(0., 15.))
    ref = fo.detect_blocks(n, h, cfg["thresh"], window, np.asarray(tpl, dtype=np.float32), cfg["corr_thresh"], raw)
    det = _detector(cfg, max_batch=96)
    got = det.detect_raw(raw)[:, 0]
    stats = parity.compare_fastdet(got, ref, what="fastdet/fresh %s" % (window,))
    assert stats["carrier"] > 20
    det.close()


def test_fastdet_gather_fallback_matches_table():
    # whole-spectrum window: the per-bin shifted template table would be N x N -> gather path
    from thrifty_b200 import synth
    tpl = parity.template_by_id("gold9_0")
    n, h = 4096, len(tpl) + 6
    raw, _ = synth.make_blocks(64, n, h, tpl, 0.8, seed=4321)
    a = _detector(dict(block_len=n, history_len=h, template=tpl, window=(7, 110), thresh=(0., 15.),
                       corr_thresh=(0., 15.)))
    b = _detector(dict(block_len=n, history_len=h, template=tpl, window=(1, 2047), thresh=(0., 15.),
                       corr_thresh=(0., 15.)))
    ra, rb = a.detect_raw(raw)[:, 0], b.detect_raw(raw)[:, 0]
    both = ((ra["flags"] & 1) != 0) & ((rb["flags"] & 1) != 0) & (ra["carrier_bin"] == rb["carrier_bin"])
    assert both.sum() > 20
    assert np.array_equal(ra["corr_sample"][both], rb["corr_sample"][both])
    np.testing.assert_allclose(ra["corr_energy"][both], rb["corr_energy"][both], rtol=1e-6)
    a.close()
    b.close()


def test_fastdet_invalid_settings():
    from thrifty_b200._native import NativeDetector, NativeError
    tpl = parity.template_by_id("gold9_0")
    with pytest.raises(ValueError):        # fastcard/cardet.c:44-48
        NativeDetector(4096, len(tpl) + 6, tpl, len(tpl), (-10, 10), (0, 15, 0), (0, 15, 0), fastdet=True)
    with pytest.raises(NativeError):       # no stddev term natively
        NativeDetector(4096, len(tpl) + 6, tpl, len(tpl), (7, 110), (0, 15, 1), (0, 15, 0), fastdet=True)
    with pytest.raises(NativeError):
        NativeDetector(4096, len(tpl) + 6, np.stack([tpl, tpl]), len(tpl), (7, 110), (0, 15, 0), (0, 15, 0),
                       fastdet=True)


def test_fastdet_cli_card_to_toad(tmp_path):
    """`python -m thrifty_b200 fastdet --card ...` writes the `.toad` lines the reference's fastdet writes
    (fastdet.cpp:191-206; golden lines come from the compiled reference sources)."""
    import subprocess
    import sys
    from thrifty_b200 import block_data, fastdet
    cfg, raw, block_idx, ref, toads, _ = parity.load_fastdet_golden("n4096_gold9_negwin")
    card, tpl, out = str(tmp_path / "in.card"), str(tmp_path / "t.tpl"), str(tmp_path / "out.toad")
    with open(card, "w") as f:
        block_data.write_card(f, raw, block_idx)
    fastdet.save_template(tpl, cfg["template"])
    cmd = [sys.executable, "-m", "thrifty_b200", "fastdet", "--card", "-i", card, "-z", tpl, "-o", out,
           "-b", str(cfg["block_len"]), "-h", str(cfg["history_len"]), "-w", "%d-%d" % cfg["window"],
           "-t", "%gc%gs" % cfg["thresh"], "-u", "%gc%gs" % cfg["corr_thresh"], "-r", "0", "--batch", "16"]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=parity.ROOT)
    assert res.returncode == 0, res.stderr
    lines = open(out).read().splitlines()
    assert len(lines) == len(toads) == int(ref["corr_detected"].sum())
    for mine, gold in zip(lines, toads):
        a, b = mine.split(" "), gold.split(" ")
        assert len(a) == 12 and a[0] == b[0] and a[2] == b[2] and a[4] == b[4] and a[8] == b[8]
        assert abs(float(a[3]) - float(b[3])) <= 1e-4                      # soa
        for x, y in zip(a[5:], b[5:]):
            assert abs(float(x) - float(y)) <= 1e-4 * max(1.0, abs(float(y)))
    assert "carrier @" in res.stdout and "Read %d blocks." % len(raw) in res.stdout


def test_fastdet_cli_raw_stream(tmp_path):
    """`fastdet -i capture.dat` (raw uint8 I/Q, no --card): blocks are formed like fastcard's raw_reader
    (raw_reader.c:15-46, first history = uint16 127) and read in place on the GPU."""
    import subprocess
    import sys
    from thrifty_b200 import fastdet
    cfg, raw, block_idx, ref, toads, stream = parity.load_fastdet_golden("n4096_gold9_stream")
    dat, tpl, out = str(tmp_path / "in.dat"), str(tmp_path / "t.tpl"), str(tmp_path / "out.toad")
    stream.tofile(dat)
    fastdet.save_template(tpl, cfg["template"])
    cmd = [sys.executable, "-m", "thrifty_b200", "fastdet", "-i", dat, "-z", tpl, "-o", out, "-k", "0", "-q",
           "-b", str(cfg["block_len"]), "-h", str(cfg["history_len"]), "-w", "%d-%d" % cfg["window"],
           "-t", "%gc%gs" % cfg["thresh"], "-u", "%gc%gs" % cfg["corr_thresh"], "-r", "3", "--batch", "10"]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=parity.ROOT)
    assert res.returncode == 0, res.stderr
    lines = open(out).read().splitlines()
    det = ref[ref["corr_detected"] != 0]
    assert len(lines) == len(det)
    for mine, r in zip(lines, det):
        f = mine.split(" ")
        assert int(f[0]) == 3 and int(f[2]) == int(r["block_idx"]) and int(f[4]) == int(r["corr_peak_idx"])
        assert int(f[8]) == int(r["carrier_argmax"]) and abs(float(f[3]) - float(r["soa"])) <= 1e-4
