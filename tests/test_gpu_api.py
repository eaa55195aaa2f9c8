"""GPU tests of the reference-facing Python seam (Detector / detector_cli) and of the C-ABI
entry points, against golden reference outputs and the oracle."""
import ctypes
import io
import os
import subprocess
import sys

import numpy as np
import pytest

import parity_util as parity
from oracle import thrifty_oracle as orc
from thrifty_b200 import block_data, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _settings(cfg):
    from thrifty_b200.detect import DetectorSettings
    return DetectorSettings(block_len=cfg["block_len"], history_len=cfg["history_len"],
                            carrier_len=len(cfg["template"]), carrier_thresh=cfg["cthresh"],
                            carrier_window=cfg["window"], template=cfg["template"],
                            corr_thresh=cfg["kthresh"])


def _results_to_rows(results):
    rows = np.zeros(len(results), dtype=orc.RECORD_DTYPE)
    for i, (detected, res) in enumerate(results):
        rows[i] = orc.result_to_row(orc.OracleResult(detected, res.timestamp, res.block, res.soa,
                                                     res.carrier_info, res.corr_info, res.rxid))
    return rows


def _rows_as_records(rows):
    """Oracle-row layout -> thr_record-like array so compare_records can be reused."""
    from thrifty_b200._native import RECORD_DTYPE
    rec = np.zeros(len(rows), dtype=RECORD_DTYPE)
    for f in ("block_idx", "soa", "carrier_bin", "carrier_offset", "carrier_energy", "carrier_noise",
              "corr_sample", "corr_offset", "corr_energy", "corr_noise"):
        rec[f] = rows[f]
    rec["flags"] = rows["carrier_detected"].astype(np.uint32) + 2 * rows["corr_detected"].astype(np.uint32)
    return rec


@pytest.mark.parametrize("name,mode", [("n4096_gold9", "complex"), ("n4096_gold9", "raw"),
                                       ("n16384_example", "raw"), ("n4096_gold9_tone", "complex")])
def test_detector_iterator_dropin(name, mode):
    """Detector(settings, blocks) yields one (detected, result) per block, in order (detect.py:80-91)."""
    from thrifty_b200.detect import Detector
    cfg, raw, block_idx, ref, _ = parity.load_golden(name)
    text = io.StringIO()
    block_data.write_card(text, raw, block_indices=block_idx)
    text.seek(0)
    blocks = block_data.card_reader(text, raw=(mode == "raw"))
    det = Detector(_settings(cfg), blocks, rxid=0, batch=7)
    results = list(det)
    assert len(results) == len(raw)
    for (detected, res), r in zip(results, ref):
        assert res.block == r["block_idx"] and res.rxid == 0
        assert (res.corr_info is None) == (not r["carrier_detected"])
        assert (res.soa is None) == (not r["carrier_detected"])
        if res.corr_info is None:
            assert detected is False and res.carrier_info.offset == 0
    parity.compare_records(_rows_as_records(_results_to_rows(results)), ref, what=name)
    det.close()


def test_detector_single_and_yield_data(golden_dir):
    """detect() on one block; yield_data=True returns shifted_fft and corr (detect.py:75-76)."""
    from thrifty_b200.detect import Detector
    g = np.load(os.path.join(golden_dir, "arrays_n4096_gold9.npz"))
    tpl = synth.gold_template(9)
    cfg = dict(block_len=4096, history_len=len(tpl) + 6, template=tpl, cthresh=(0., 15., 0.),
               window=(7, 110), kthresh=(0., 15., 0.))
    raw, _ = synth.make_blocks(1, 4096, len(tpl) + 6, tpl, 1.0, seed=int(g["seed"]))
    det = Detector(_settings(cfg), rxid=5, yield_data=True)
    detected, res, sfft, corr = det.detect(12.5, 5, block_data.raw_to_complex(raw[0]))
    assert detected and res.rxid == 5 and res.timestamp == 12.5 and res.block == 5
    assert abs(res.soa - float(g["soa"])) < 1e-4
    ref_sfft, ref_corr = g["shifted_fft"], g["corr"]
    assert sfft.shape == ref_sfft.shape and corr.shape == ref_corr.shape
    # complex arrays: error relative to the largest magnitude (float32 FFT vs float64 reference)
    assert np.abs(sfft - ref_sfft).max() <= 2e-5 * np.abs(ref_sfft).max()
    assert np.abs(corr - ref_corr).max() <= 2e-5 * np.abs(ref_corr).max()
    det.close()
    det2 = Detector(_settings(cfg), rxid=5)
    d2, r2 = det2.detect(12.5, 5, raw[0])          # raw uint8 block, no yield_data
    # the batched path prunes FFT#1 for this window, the yield_data path computes it in full: the carrier
    # offset (hence the mix) differs in the last float32 bits
    assert d2 and abs(r2.soa - res.soa) < 1e-6 and r2.corr_info.sample == res.corr_info.sample
    det2.close()


def test_fft_mag_debug_output():
    """|FFT(block)| from the kernel equals numpy's float32 FFT magnitude (Signal.fft.mag)."""
    from thrifty_b200._native import NativeDetector
    tpl = synth.gold_template(10)
    raw, _ = synth.make_blocks(2, 8192, len(tpl) + 6, tpl, 1.0, seed=77)
    det = NativeDetector(8192, len(tpl) + 6, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=4)
    _, _, _, mag = det.detect_block_data(raw=raw[1], block_idx=0)
    ref = np.abs(np.fft.fft(orc.raw_to_complex(raw[1]))).astype(np.float64)
    assert np.abs(mag - ref).max() <= 2e-6 * ref.max()
    det.close()


def test_card_ingest_gpu_decode():
    """thr_detect_card: `.card` text scanned on the host, base64 decoded on the GPU == detect_raw."""
    from thrifty_b200._native import NativeDetector, NativeError
    from thrifty_b200.detect import Detector
    for name in ("n4096_gold9", "n16384_example"):
        cfg, raw, block_idx, ref, _ = parity.load_golden(name)
        text = io.StringIO()
        block_data.write_card(text, raw, block_indices=block_idx, t0=1000.0, dt=0.0047767)
        data = ("linux; GNU C++ version\n" + text.getvalue()).encode()
        det = NativeDetector(cfg["block_len"], cfg["history_len"], cfg["template"], len(cfg["template"]),
                             cfg["window"], cfg["cthresh"], cfg["kthresh"], max_batch=20)
        want = det.detect_raw(raw, block_idx)
        ts, idx, got, consumed = det.detect_card(data)
        assert consumed == len(data) and len(ts) == len(raw)
        np.testing.assert_array_equal(idx, block_idx)
        np.testing.assert_allclose(ts, 1000.0 + 0.0047767 * np.arange(len(raw)), atol=2e-6)
        assert got.tobytes() == want.tobytes()            # decoded payloads are bit-identical
        parity.compare_records(got[:, 0], ref, what=name + "/card")
        # CRLF line ends and a chunk boundary in the middle of a line
        crlf = data.replace(b"\n", b"\r\n")
        cut = len(crlf) // 2 + 123
        t1, i1, r1, used = det.detect_card(crlf[:cut], final=False)
        t2, i2, r2, used2 = det.detect_card(crlf[used:], final=True)
        assert used <= cut and used + used2 == len(crlf)
        assert np.concatenate([r1, r2]).tobytes() == want.tobytes()
        # a non-base64 character inside a payload is an error, not silent garbage
        bad = bytearray(data)
        bad[len(bad) // 2] = ord("!")
        with pytest.raises(NativeError):
            det.detect_card(bytes(bad))
        # a short payload in a later chunk (lines are scanned chunk by chunk): same line number as a scan of the whole text
        lines = data.split(b"\n")
        victim = [i for i, ln in enumerate(lines) if ln and ln[:1].isdigit()][len(raw) - 2]
        lines[victim] = lines[victim][:-5]
        with pytest.raises(NativeError, match="data line %d is malformed" % (victim + 1)):
            det.detect_card(b"\n".join(lines))
        ts3, _, r3, _ = det.detect_card(data)                # the detector is usable after both errors
        assert r3.tobytes() == want.tobytes()
        det.close()
    # the Detector front end over a binary stream
    cfg, raw, block_idx, ref, _ = parity.load_golden("n4096_gold9")
    text = io.StringIO()
    block_data.write_card(text, raw, block_indices=block_idx)
    d = Detector(_settings(cfg), rxid=0, batch=16)
    results = list(d.detect_card_stream(io.BytesIO(text.getvalue().encode()), chunk_bytes=50000))
    assert len(results) == len(raw)
    parity.compare_records(_rows_as_records(_results_to_rows(results)), ref, what="card stream")
    # the two page-locked staging buffers are pinned once per process: a second stream of the same chunk size gets the
    # same memory back (and the same results)
    from thrifty_b200 import _native
    cached = {b.ptr for b in _native._staging_pool.get(50000 + 1, [])}
    assert cached
    again = list(d.detect_card_stream(io.BytesIO(text.getvalue().encode()), chunk_bytes=50000))
    assert {b.ptr for b in _native._staging_pool.get(50000 + 1, [])} == cached
    assert [(ok, r.serialize() if ok else r.block) for ok, r in again] == [(ok, r.serialize() if ok else r.block) for ok, r in results]
    d.close()


def test_raw_stream_front_end():
    """`--raw` streams: overlapping windows read in place on the GPU == the reference block_reader
    (block_data.py:70-98) followed by Detector.detect on every block, including block 0."""
    from thrifty_b200.detect import Detector
    tpl = synth.gold_template(9)
    n, hist = 4096, len(tpl) + 6
    new = n - hist
    nblk = 41
    rng = np.random.default_rng(2024)
    total = nblk * new + 77                              # trailing partial block is dropped
    x = 0.02 * (rng.standard_normal(total) + 1j * rng.standard_normal(total))
    pos = 500
    while pos + len(tpl) < total:                        # a burst every ~1.7 blocks
        f = rng.uniform(8, 109)
        t = np.arange(len(tpl))
        x[pos:pos + len(tpl)] += rng.uniform(0.15, 0.4) * (tpl + 1) / 2 * np.exp(2j * np.pi * f * (pos + t) / n)
        pos += int(rng.uniform(1.2, 2.2) * new)
    stream = synth.complex_to_raw(x).tobytes()
    cfg = dict(block_len=n, history_len=hist, template=tpl, cthresh=(0., 15., 0.), window=(7, 110),
               kthresh=(0., 15., 0.))
    st = orc.DetectorSettings(n, hist, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.))
    odet = orc.Detector(st, rxid=0)
    ref = np.zeros(nblk, dtype=orc.RECORD_DTYPE)
    blocks = list(orc.block_reader(io.BytesIO(stream), n, hist))
    assert len(blocks) == nblk
    for i, (bi, data) in enumerate(blocks):
        ref[i] = orc.result_to_row(odet.detect(0.0, bi, data))
    det = Detector(_settings(cfg), rxid=0, batch=16)
    results = list(det.detect_raw_stream(io.BytesIO(stream), chunk_blocks=7))
    assert len(results) == nblk
    assert [r.block for _, r in results] == list(range(nblk))
    stats = parity.compare_records(_rows_as_records(_results_to_rows(results)), ref, what="raw stream")
    assert stats["detected"] >= 10
    det.close()


def test_cli_card_to_toad(tmp_path):
    """`thrifty_b200 detect x.card -o x.toad` reproduces the reference's .toad lines."""
    cfg, raw, block_idx, ref, lines = parity.load_golden("n4096_gold9")
    np.save(str(tmp_path / "template.npy"), cfg["template"])
    (tmp_path / "detector.cfg").write_text(
        "rxid: 0\nsample_rate: 2.4M\nblock_size: 4096\nblock_history: %d\ncarrier_window: 7 - 110\n"
        "carrier_threshold: 15 * snr\ncorr_threshold: 15 * snr\ntemplate: template.npy\n" % cfg["history_len"])
    with open(str(tmp_path / "rx.card"), "w") as f:
        block_data.write_card(f, raw, block_indices=block_idx, t0=1000.0, dt=0.0047767)
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "thrifty_b200", "detect", "rx.card", "-o", "rx.toad", "--batch", "16"],
                         cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    summary = out.stdout.strip().split("\n")
    assert len(summary) == len(raw) and summary[0].startswith("blk=10; carrier:")
    got = [l for l in open(str(tmp_path / "rx.toad")).read().split("\n") if l]
    want = [l for l in lines if l]
    assert len(got) == len(want)
    for lg, lw in zip(got, want):
        fg, fw = lg.split(), lw.split()
        assert len(fg) == len(fw) == 12
        assert [fg[0], fg[2], fg[4], fg[8]] == [fw[0], fw[2], fw[4], fw[8]]      # rxid block sample bin
        assert abs(float(fg[1]) - float(fw[1])) < 2e-6                              # timestamp
        assert abs(float(fg[3]) - float(fw[3])) < 1e-4                              # soa
        assert abs(float(fg[5]) - float(fw[5])) < 1e-4 and abs(float(fg[9]) - float(fw[9])) < 1e-4
        for i in (6, 7, 10, 11):
            assert abs(float(fg[i]) / float(fw[i]) - 1) < 1e-4
    # --quiet: no result objects, native .toad text (thr_format_toad), reader thread -- the same file, byte for byte
    out = subprocess.run([sys.executable, "-m", "thrifty_b200", "detect", "rx.card", "-o", "rx_quiet.toad", "--batch", "16",
                          "--quiet"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout == "", out.stderr
    assert open(str(tmp_path / "rx_quiet.toad")).read() == open(str(tmp_path / "rx.toad")).read()
    # ... and from a pipe (no seekable file: sequential reads, small pieces)
    with open(str(tmp_path / "rx.card"), "rb") as f:
        out = subprocess.run([sys.executable, "-m", "thrifty_b200", "detect", "-", "-o", "rx_pipe.toad", "--batch", "16",
                              "--quiet"], cwd=str(tmp_path), env=env, stdin=f, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert open(str(tmp_path / "rx_pipe.toad")).read() == open(str(tmp_path / "rx.toad")).read()


@pytest.mark.parametrize("block_len,n_blocks,seed", [(16384, 192, 1), (8192, 128, 2), (4096, 128, 3),
                                                     (2048, 64, 4), (1024, 64, 5), (32768, 24, 6)])
def test_fresh_seeds_vs_oracle(block_len, n_blocks, seed):
    """Parity on seeds that are not in the goldens, every supported block length."""
    from thrifty_b200._native import NativeDetector
    if block_len >= 16384:
        tpl = np.load(os.path.join(parity.GOLDEN, "template_example.npy"))
        hist = 4920
    else:
        bits = {8192: 10, 4096: 9, 2048: 8, 1024: 7}[block_len]
        tpl = synth.gold_template(bits)
        hist = len(tpl) + 6
    window = (7, 110) if block_len >= 2048 else (7, 60)
    bins = (8.0, 109.0) if block_len >= 2048 else (8.0, 59.0)
    raw, _ = synth.make_blocks(n_blocks, block_len, hist, tpl, 0.5, seed=900000 + 1000 * seed, bin_range=bins)
    st = orc.DetectorSettings(block_len, hist, len(tpl), (0., 15., 0.), window, tpl, (0., 15., 0.))
    idx = np.arange(n_blocks, dtype=np.int64) * 5 + 100
    ref = orc.detect_blocks(st, raw, idx)
    det = NativeDetector(block_len, hist, tpl, len(tpl), window, (0., 15., 0.), (0., 15., 0.), max_batch=50)
    got = det.detect_raw(raw, idx)[:, 0]                 # n_blocks > max_batch: chunked inside the call
    stats = parity.compare_records(got, ref, what="N=%d" % block_len)
    print(block_len, stats)
    assert stats["carrier"] >= n_blocks // 4
    det.close()


@pytest.mark.parametrize("window,cthresh,kthresh", [
    ((7, 300), (0., 15., 0.), (0., 15., 0.)),         # window beyond bin 124: full FFT#1 path
    ((7, 110), (1., 12., 2.), (0.5, 10., 3.)),        # stddev terms: full FFT#1 path, extra reductions
    ((-300, -7), (0., 15., 0.), (0., 15., 0.)),       # negative-frequency window
    ((3, 124), (0., 15., 0.), (0., 15., 0.)),         # widest window the pruned FFT#1 path accepts without pre-shift
    ((-110, -7), (0., 15., 0.), (0., 15., 0.)),       # narrow negative-frequency window: pruned path, band pre-shifted
    ((300, 400), (0., 15., 0.), (0., 15., 0.)),       # narrow window far from DC: pruned path, band pre-shifted
    ((16000, 16100), (0., 15., 0.), (0., 15., 0.)),   # same, given as unsigned bins above N/2
])
def test_n16384_carrier_paths(window, cthresh, kthresh):
    """N=16384: the pruned ('zoom') and the full FFT#1 carrier paths both match the oracle."""
    from thrifty_b200._native import NativeDetector
    tpl = np.load(os.path.join(parity.GOLDEN, "template_example.npy"))
    n, hist, nblk = 16384, 4920, 96
    lo, hi = (window[0] + 1.0, window[1] - 1.0)
    raw, _ = synth.make_blocks(nblk, n, hist, tpl, 0.6, seed=31337 + abs(window[0]), bin_range=(lo, hi))
    st = orc.DetectorSettings(n, hist, len(tpl), cthresh, window, tpl, kthresh)
    ref = orc.detect_blocks(st, raw)
    det = NativeDetector(n, hist, tpl, len(tpl), window, cthresh, kthresh, max_batch=128)
    got = det.detect_raw(raw)[:, 0]
    stats = parity.compare_records(got, ref, what="N=16384 window %s" % (window,))
    print(window, stats)
    assert stats["carrier"] >= nblk // 3
    det.close()


def test_whole_spectrum_window_and_degenerate_blocks():
    """Default window '0--1' (all bins) and degenerate inputs: constant, saturated, DC-only."""
    from thrifty_b200._native import NativeDetector
    tpl = synth.gold_template(9)
    n, hist = 4096, len(tpl) + 6
    raw, _ = synth.make_blocks(12, n, hist, tpl, 1.0, seed=555)
    rng = np.random.default_rng(9)
    raw[3] = 127                                       # constant block -> only the DC bin: NaN noise
    raw[5] = rng.integers(0, 2, size=2 * n) * 255      # saturated random
    raw[7] = 0
    st = orc.DetectorSettings(n, hist, len(tpl), (0., 15., 0.), (0, -1), tpl, (0., 15., 0.))
    with np.errstate(all="ignore"):
        ref = orc.detect_blocks(st, raw)
    det = NativeDetector(n, hist, tpl, len(tpl), (0, -1), (0., 15., 0.), (0., 15., 0.), max_batch=16)
    got = det.detect_raw(raw)[:, 0]
    for i in (3, 7):
        assert not ref["carrier_detected"][i] and not (got["flags"][i] & 1)
        assert got["carrier_bin"][i] == ref["carrier_bin"][i] == 0
    keep = np.array([i for i in range(12) if i not in (3, 7)])
    parity.compare_records(got[keep], ref[keep], what="whole-spectrum")
    det.close()


def test_empty_and_single_block():
    from thrifty_b200._native import NativeDetector
    tpl = synth.gold_template(9)
    n, hist = 4096, len(tpl) + 6
    det = NativeDetector(n, hist, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=8)
    assert det.detect_raw(np.zeros((0, 2 * n), dtype=np.uint8)).shape == (0, 1)
    raw, _ = synth.make_blocks(1, n, hist, tpl, 1.0, seed=31)
    got = det.detect_raw(raw, [123456789012])
    assert got.shape == (1, 1) and got["block_idx"][0, 0] == 123456789012
    assert got["flags"][0, 0] == 3
    # SoA keeps integer precision for huge block indices (detect.py:67)
    expect = (n - hist) * 123456789012 + int(got["corr_sample"][0, 0]) + float(got["corr_offset"][0, 0])
    assert abs(got["soa"][0, 0] - expect) <= 1.0
    with pytest.raises(ValueError):
        det.detect_raw(np.zeros((2, 100), dtype=np.uint8))
    det.close()


def test_invalid_settings_raise():
    from thrifty_b200._native import NativeDetector, NativeError
    tpl = synth.gold_template(9)
    with pytest.raises(ValueError):                       # carrier_detect.py:47-49
        NativeDetector(4096, len(tpl) + 6, tpl, len(tpl), (-5000, 10), (0., 15., 0.), (0., 15., 0.))
    with pytest.raises(NativeError):                      # soa_estimator.py:33
        NativeDetector(4096, 100, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.))
    with pytest.raises(NativeError):
        NativeDetector(5000, 4000, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.))


def test_full_size_properties():
    """BASELINE full size (N=16384, batch 4096+): size-independent properties.

    * determinism / batch invariance: every copy of a block yields the same record wherever it
      sits in the batch, across chunk boundaries and CTAs;
    * SoA linearity in block_idx: soa - (N-H)*block_idx == sample + offset;
    * u8 and complex64 inputs agree; host-buffer and device-buffer entry points agree."""
    from thrifty_b200._native import NativeDetector, RECORD_DTYPE, load_library
    tpl = np.load(os.path.join(parity.GOLDEN, "template_example.npy"))
    n, hist, uniq_n, total = 16384, 4920, 96, 4096 + 160
    uniq, _ = synth.make_blocks(uniq_n, n, hist, tpl, 0.75, seed=424242)
    rng = np.random.default_rng(3)
    pick = rng.integers(0, uniq_n, size=total)
    raw = uniq[pick]
    idx = rng.integers(0, 1 << 33, size=total).astype(np.int64)
    det = NativeDetector(n, hist, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=4096)
    got = det.detect_raw(raw, idx)[:, 0]
    fields = ["carrier_bin", "carrier_offset", "carrier_energy", "carrier_noise", "corr_sample",
              "corr_offset", "corr_energy", "corr_noise", "flags", "signal_energy"]
    first = {}
    for i in range(total):
        key = tuple(got[f][i].tobytes() for f in fields)
        assert first.setdefault(pick[i], key) == key, "block copy %d differs" % i
    car = (got["flags"] & 1) != 0
    assert car.sum() > total // 2
    lin = got["soa"][car] - (n - hist) * idx[car].astype(np.float64)
    np.testing.assert_allclose(lin, got["corr_sample"][car] + got["corr_offset"][car].astype(np.float64),
                               rtol=0, atol=1e-2)     # float64 ulp at (N-H)*2^33 ~ 1e14 is 0.016
    # oracle on the unique blocks
    st = orc.DetectorSettings(n, hist, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.))
    ref = orc.detect_blocks(st, uniq)
    firsts = np.array([np.nonzero(pick == u)[0][0] for u in range(uniq_n)])
    g = got[firsts].copy()
    g["block_idx"] = np.arange(uniq_n)
    # re-base the SoA on small block indices (at idx ~ 2^33 a float64 SoA only resolves 0.016 samples;
    # linearity in block_idx was checked above)
    g["soa"] = (n - hist) * np.arange(uniq_n) + g["corr_sample"] + g["corr_offset"].astype(np.float64)
    parity.compare_records(g, ref, what="full-size")
    # complex64 input path, first 300 blocks
    iq = np.stack([orc.raw_to_complex(r) for r in raw[:300]])
    got_c = det.detect_c64(iq, idx[:300])[:, 0]
    for f in fields + ["soa", "block_idx"]:
        np.testing.assert_array_equal(got_c[f], got[f][:300], err_msg=f)
    # device-buffer entry point
    lib = load_library()
    nb = 1000
    d_raw = lib.thr_device_alloc(0, nb * 2 * n)
    d_idx = lib.thr_device_alloc(0, nb * 8)
    d_out = lib.thr_device_alloc(0, nb * 64)
    assert d_raw and d_idx and d_out
    chunk = np.ascontiguousarray(raw[:nb])
    cidx = np.ascontiguousarray(idx[:nb])
    assert lib.thr_memcpy_h2d(0, d_raw, chunk.ctypes.data, chunk.nbytes) == 0
    assert lib.thr_memcpy_h2d(0, d_idx, cidx.ctypes.data, cidx.nbytes) == 0
    det.detect_device(d_raw, d_idx, nb, d_out)
    det.synchronize()
    out = np.zeros(nb, dtype=RECORD_DTYPE)
    assert lib.thr_memcpy_d2h(0, out.ctypes.data, d_out, out.nbytes) == 0
    for f in fields + ["soa", "block_idx"]:
        np.testing.assert_array_equal(out[f], got[f][:nb], err_msg=f)
    for p in (d_raw, d_idx, d_out):
        lib.thr_device_free(0, p)
    assert det.info()["launches"] >= 3
    det.close()


def test_time_shift_property():
    """Moving the burst by d samples moves the SoA by d (noise-free, N=16384)."""
    from thrifty_b200._native import NativeDetector
    tpl = np.load(os.path.join(parity.GOLDEN, "template_example.npy"))
    n, hist = 16384, 4920
    idx = np.arange(n)
    blocks, positions = [], [3, 100, 1001, 5000, 11466]
    for pos in positions:
        gate = np.zeros(n)
        gate[pos:pos + len(tpl)] = (tpl[:n - pos] + 1) / 2 if pos + len(tpl) > n else (tpl + 1) / 2
        x = 0.3 * gate * np.exp(2j * np.pi * 40.25 * idx / n)
        blocks.append(synth.complex_to_raw(x))
    det = NativeDetector(n, hist, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=8)
    got = det.detect_raw(np.stack(blocks), np.zeros(len(positions), dtype=np.int64))[:, 0]
    assert np.all(got["flags"] == 3)
    np.testing.assert_array_equal(got["corr_sample"], positions)
    np.testing.assert_array_equal(got["carrier_bin"], 40)
    assert np.ptp(got["corr_offset"]) < 0.02 and np.ptp(got["carrier_offset"]) < 0.01
    det.close()


def test_multi_template_vs_independent_oracles():
    """4 Gold templates jointly == 4 independent reference detectors (BASELINE config 5)."""
    from thrifty_b200.detect import DetectorSettings, MultiTemplateDetector
    n, bits = 4096, 9
    tpls = np.stack([synth.gold_template(bits, i) for i in range(4)])
    hist = tpls.shape[1] + 6
    rng = np.random.default_rng(11)
    raws, which = [], []
    for b in range(40):
        t = int(rng.integers(0, 4))
        r, _ = synth.make_blocks(1, n, hist, tpls[t], 0.8, seed=7000 + b)
        raws.append(r[0])
        which.append(t)
    raws = np.stack(raws)
    settings = DetectorSettings(n, hist, tpls.shape[1], (0., 15., 0.), (7, 110), tpls[0], (0., 15., 0.))
    det = MultiTemplateDetector(settings, tpls, rxid=1, batch=64)
    out = det.detect_many([(0.0, i, raws[i]) for i in range(len(raws))])
    for t in range(4):
        st = orc.DetectorSettings(n, hist, tpls.shape[1], (0., 15., 0.), (7, 110), tpls[t], (0., 15., 0.))
        ref = orc.detect_blocks(st, raws)
        rows = _results_to_rows([o[t] for o in out])
        parity.compare_records(_rows_as_records(rows), ref, what="template %d" % t)
        assert all(o[t][1].txid == t for o in out)
    det.close()


@pytest.mark.parametrize("window", [(7, 110), (7, 300)])
def test_multi_template_n32768_two_half_kernel(window):
    """Three Gold-11 templates jointly at block_len 32768: the 2 x 16384 kernel (E' and O' parked for the template loop)
    == three independent oracle detectors == the generic global-scratch kernel; several blocks per CTA."""
    from thrifty_b200._native import NativeDetector
    n = 32768
    tpls = np.stack([synth.gold_template(11, i) for i in range(3)])
    hist = tpls.shape[1] + 6
    rng = np.random.default_rng(12)
    raws = []
    for b in range(30):
        t = int(rng.integers(0, 3))
        r, _ = synth.make_blocks(1, n, hist, tpls[t], 0.8, seed=7100 + b, bin_range=(9.0, window[1] - 2.0))
        raws.append(r[0])
    raws = np.stack(raws)
    idx = 7 + 3 * np.arange(len(raws), dtype=np.int64)
    two = NativeDetector(n, hist, tpls, tpls.shape[1], window, (0., 15., 0.), (0., 15., 0.), max_batch=600)
    gen = NativeDetector(n, hist, tpls, tpls.shape[1], window, (0., 15., 0.), (0., 15., 0.), max_batch=64, generic_kernel=True)
    assert "detect2x" in two.info()["kernel"] and "gmem" in gen.info()["kernel"]
    got2 = two.detect_raw(raws, idx)
    gotg = gen.detect_raw(raws, idx)
    assert got2.shape == (len(raws), 3)
    n_det = 0
    for t in range(3):
        st = orc.DetectorSettings(n, hist, tpls.shape[1], (0., 15., 0.), window, tpls[t], (0., 15., 0.))
        ref = orc.detect_blocks(st, raws, idx)
        stats = parity.compare_records(got2[:, t], ref, what="n32768/2x template %d" % t)
        parity.compare_records(gotg[:, t], ref, what="n32768/generic template %d" % t)
        assert np.all(got2[:, t]["template_idx"] == t)
        n_det += stats["detected"]
        for f in ("flags", "carrier_bin", "corr_sample"):
            assert np.array_equal(got2[:, t][f], gotg[:, t][f]), f
    assert n_det > 10
    big = two.detect_raw(raws[np.arange(600) % 30])                    # 4 blocks per CTA through the pipeline
    for f in ("flags", "carrier_bin", "corr_sample", "corr_energy", "corr_offset", "carrier_offset"):
        assert np.array_equal(big[f][:30], big[f][30 * 19:30 * 20], equal_nan=True), f
        assert np.array_equal(big[f][:30], got2[f], equal_nan=True), f
    two.close()
    gen.close()


def test_multi_template_n16384_gold11x4_vs_reference_golden():
    """BASELINE config 5 at its stated size: four Gold-11 templates (L=4914), block_len 16384, history 4920, 320 blocks
    (> 2 per persistent CTA) through MultiTemplateDetector == four independent runs of the reference's own Detector
    (tests/golden/detect_n16384_gold11x4.npz, written by oracle/make_golden_multi.py from /root/reference)."""
    from thrifty_b200.detect import DetectorSettings, MultiTemplateDetector
    cfg, tpls, raw, block_idx, ref, which = parity.load_multi_golden()
    settings = DetectorSettings(cfg["block_len"], cfg["history_len"], tpls.shape[1], cfg["cthresh"], cfg["window"],
                                tpls[0], cfg["kthresh"])
    det = MultiTemplateDetector(settings, tpls, rxid=2, batch=512)
    assert "multi" in det.native.info()["kernel"] and "16384" in det.native.info()["kernel"]
    out = det.detect_many([(0.0, int(block_idx[i]), raw[i]) for i in range(len(raw))])
    for t in range(len(tpls)):
        rows = _results_to_rows([o[t] for o in out])
        stats = parity.compare_records(_rows_as_records(rows), ref[t], what="gold11x4 template %d" % t)
        assert stats["carrier"] > 200 and stats["detected"] > 200
        assert all(o[t][1].txid == t for o in out)
    # the raw record path (no result objects): same launch, records [B, T]
    recs = det.native.detect_raw(raw, block_idx)
    for t in range(len(tpls)):
        assert np.all(recs[:, t]["template_idx"] == t)
        parity.compare_records(recs[:, t], ref[t], what="gold11x4 records template %d" % t)
    det.close()


def test_multi_template_n32768_gold11x3_vs_reference_golden():
    """Three Gold-11 templates at block_len 32768 with a carrier window too wide for the pruned FFT#1 (2 x 16384 kernel,
    FFT#1 in full, template loop over the parked half spectra) == three independent runs of the reference's own Detector
    (tests/golden/detect_n32768_gold11x3.npz, oracle/make_golden_multi.py)."""
    from thrifty_b200._native import NativeDetector
    cfg, tpls, raw, block_idx, ref, which = parity.load_multi_golden("n32768_gold11x3")
    det = NativeDetector(cfg["block_len"], cfg["history_len"], tpls, tpls.shape[1], cfg["window"], cfg["cthresh"],
                         cfg["kthresh"], max_batch=64)
    assert "detect2x" in det.info()["kernel"] and "multi" in det.info()["kernel"]
    recs = det.detect_raw(raw, block_idx)
    for t in range(len(tpls)):
        assert np.all(recs[:, t]["template_idx"] == t)
        stats = parity.compare_records(recs[:, t], ref[t], what="n32768 gold11x3 template %d" % t)
        assert stats["carrier"] == int(ref[t]["carrier_detected"].sum()) and stats["detected"] > 30
    det.close()


def test_raw_stream_card_export(tmp_path):
    """`detect --raw --card-out`: every carrier-positive block of the stream is re-exported as a .card line
    (fastcard/fastcard_cli.c:171-193); detecting that .card gives the same records as the stream did."""
    from thrifty_b200.detect import Detector, DetectorSettings
    tpl = synth.gold_template(9)
    n, hist = 4096, len(tpl) + 6
    new = n - hist
    nblk = 60
    rng = np.random.default_rng(77)
    total = nblk * new
    x = 0.02 * (rng.standard_normal(total) + 1j * rng.standard_normal(total))
    pos = 900
    while pos + len(tpl) < total:
        t = np.arange(len(tpl))
        x[pos:pos + len(tpl)] += 0.3 * (tpl + 1) / 2 * np.exp(2j * np.pi * rng.uniform(8, 109) * (t + pos) / n)
        pos += int(rng.uniform(2.2, 4.0) * new)
    stream = synth.complex_to_raw(x).tobytes()
    st = DetectorSettings(n, hist, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.))
    det = Detector(st, rxid=0, batch=16)
    card = io.StringIO()
    from_stream = list(det.detect_raw_stream(io.BytesIO(stream), chunk_blocks=16, card_out=card))
    det.close()
    assert len(from_stream) == nblk
    with_carrier = [r for _, r in from_stream if r.corr_info is not None]
    assert 10 <= len(with_carrier) < nblk
    lines = card.getvalue().splitlines()
    assert len(lines) == len(with_carrier)
    det2 = Detector(st, block_data.card_reader(io.StringIO(card.getvalue()), raw=True), rxid=0, batch=16)
    from_card = list(det2)
    det2.close()
    assert [r.block for _, r in from_card] == [r.block for r in with_carrier]
    n_zero_hist = -(-hist // new)                    # blocks whose history precedes the stream (complex zeros there;
    for (d1, r1), r0 in zip(from_card, with_carrier):   # the exported bytes say 127 = -0.003: equal only to ~1e-6)
        assert r1.corr_info is not None and r1.carrier_info.bin == r0.carrier_info.bin
        if r0.block >= n_zero_hist:
            assert r1.soa == r0.soa and r1.corr_info == r0.corr_info
        else:
            assert abs(r1.soa - r0.soa) < 1e-3 and r1.corr_info.sample == r0.corr_info.sample


def test_pipelined_chunks_do_not_share_scratch():
    """A host-buffer call is cut into chunks that run on two streams and overlap on the device; kernels with per-CTA
    global scratch (multi-template X' save area, block_len 32768) must not share it between the two chunks in flight
    (found in round 2: records of a 4-template detector changed from run to run once a call spanned two chunks)."""
    from thrifty_b200._native import NativeDetector
    tpls = np.stack([synth.gold_template(11, i) for i in range(3)])
    n, h = 16384, tpls.shape[1] + 6
    raw, _ = synth.make_blocks(64, n, h, tpls[1], 0.8, seed=99)
    raw = raw[np.arange(1400) % 64]
    det = NativeDetector(n, h, tpls, tpls.shape[1], (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=1024)
    want = np.concatenate([det.detect_raw(raw[a:a + 200], np.arange(a, a + 200)) for a in range(0, 1400, 200)])   # one chunk each
    for _ in range(5):
        assert det.detect_raw(raw, np.arange(1400)).tobytes() == want.tobytes()
    det.close()
    tpl = np.load(os.path.join(parity.GOLDEN, "template_example.npy"))
    raw, _ = synth.make_blocks(32, 32768, 4920, tpl, 0.8, seed=98)
    raw = raw[np.arange(700) % 32]
    det = NativeDetector(32768, 4920, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=512)
    want = np.concatenate([det.detect_raw(raw[a:a + 100], np.arange(a, a + 100)) for a in range(0, 700, 100)])
    for _ in range(3):
        assert det.detect_raw(raw, np.arange(700)).tobytes() == want.tobytes()
    det.close()


def test_smoke_entry():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.smoke()


def test_n32768_two_half_kernel_shifted_band():
    """2 x 16384 kernel with the zoom band pre-shifted (negative-frequency window) against the oracle."""
    from thrifty_b200._native import NativeDetector
    tpl = np.load(os.path.join(os.path.dirname(__file__), "golden", "template_example.npy"))
    n, h = 32768, 4920
    raw, _ = synth.make_blocks(40, n, h, tpl, 0.7, seed=9002, bin_range=(-108.0, -9.0))
    st = orc.DetectorSettings(n, h, len(tpl), (0., 15., 0.), (-110, -7), tpl, (0., 15., 0.))
    ref = orc.detect_blocks(st, raw)
    det = NativeDetector(n, h, tpl, len(tpl), (-110, -7), (0., 15., 0.), (0., 15., 0.), max_batch=64)
    assert "detect2x" in det.info()["kernel"]
    stats = parity.compare_records(det.detect_raw(raw)[:, 0], ref, what="n32768/2x shifted band")
    assert stats["carrier"] > 15
    det.close()


@pytest.mark.parametrize("window,cthresh,bins", [
    ((7, 300), (0., 15., 0.), (9.0, 298.0)),             # wide window: FFT#1 in full
    ((-400, -5), (0., 15., 0.), (-398.0, -7.0)),         # negative frequencies: bins of the upper half spectrum
    ((5, -5), (2.0, 12., 0.), (-16000.0, 16000.0)),      # every bin but the DC leak of rawconv (a peak there is noise on
                                                         # both sides: the fit is ill-conditioned), constant term
    ((7, 110), (0., 12., 2.5), (9.0, 108.0)),            # narrow window, but a stddev term needs every magnitude
    ((16000, 16800), (0., 15., 1.0), (16010.0, 16790.0)),  # window across the seam between the two half spectra
])
def test_n32768_two_half_kernel_full_fft1(window, cthresh, bins):
    """2 x 16384 kernel with FFT#1 computed in full (carrier window wider than the 128-bin zoom band, or a carrier stddev
    threshold term): oracle parity, agreement with the generic kernel, and several blocks per CTA."""
    from thrifty_b200._native import NativeDetector
    tpl = np.load(os.path.join(os.path.dirname(__file__), "golden", "template_example.npy"))
    n, h = 32768, 4920
    raw, _ = synth.make_blocks(40, n, h, tpl, 0.7, seed=9003, bin_range=bins)
    idx = 3 + 2 * np.arange(40, dtype=np.int64)
    st = orc.DetectorSettings(n, h, len(tpl), cthresh, window, tpl, (0., 15., 0.))
    ref = orc.detect_blocks(st, raw, idx)
    two = NativeDetector(n, h, tpl, len(tpl), window, cthresh, (0., 15., 0.), max_batch=700)
    gen = NativeDetector(n, h, tpl, len(tpl), window, cthresh, (0., 15., 0.), max_batch=64, generic_kernel=True)
    assert "detect2x" in two.info()["kernel"] and "gmem" in gen.info()["kernel"]
    got2 = two.detect_raw(raw, idx)[:, 0]
    gotg = gen.detect_raw(raw, idx)[:, 0]
    stats = parity.compare_records(got2, ref, what="n32768/2x full FFT#1")
    parity.compare_records(gotg, ref, what="n32768/generic")
    assert stats["carrier"] > 10
    for f in ("flags", "carrier_bin", "corr_sample"):
        assert np.array_equal(got2[f], gotg[f]), f
    gotb = two.detect_raw(raw[np.arange(700) % 40])[:, 0]          # 4-5 blocks per CTA through the pipeline
    for f in ("flags", "carrier_bin", "corr_sample", "corr_energy", "corr_offset", "carrier_offset", "carrier_energy"):
        assert np.array_equal(gotb[f][:40], gotb[f][40 * 16:40 * 17], equal_nan=True), f
        assert np.array_equal(gotb[f][:40], got2[f], equal_nan=True), f
    two.close()
    gen.close()


@pytest.mark.parametrize("kthresh", [(0., 15., 0.), (0.5, 10., 3.)])
def test_n32768_two_half_kernel_vs_generic_and_oracle(kthresh):
    """block_len 32768 runs as two interleaved 16384-point transforms (detect_kernel_2x.cuh) when FFT#1 can
    be pruned; the generic global-scratch kernel and the oracle must agree with it (raw and complex64 input)."""
    from oracle import thrifty_oracle as orc
    from thrifty_b200._native import NativeDetector
    tpl = np.load(os.path.join(os.path.dirname(__file__), "golden", "template_example.npy"))
    n, h = 32768, 4920
    raw, _ = synth.make_blocks(40, n, h, tpl, 0.7, seed=9001)
    idx = 5 + 2 * np.arange(40, dtype=np.int64)
    st = orc.DetectorSettings(n, h, len(tpl), (0., 15., 0.), (7, 110), tpl, kthresh)
    ref = orc.detect_blocks(st, raw, idx)
    two = NativeDetector(n, h, tpl, len(tpl), (7, 110), (0., 15., 0.), kthresh, max_batch=64)
    gen = NativeDetector(n, h, tpl, len(tpl), (7, 110), (0., 15., 0.), kthresh, max_batch=64, generic_kernel=True)
    assert "detect2x" in two.info()["kernel"] and "gmem" in gen.info()["kernel"]
    got2 = two.detect_raw(raw, idx)[:, 0]
    gotg = gen.detect_raw(raw, idx)[:, 0]
    stats = parity.compare_records(got2, ref, what="n32768/2x")
    parity.compare_records(gotg, ref, what="n32768/generic")
    assert stats["carrier"] > 15
    assert np.array_equal(got2["corr_sample"], gotg["corr_sample"]) and np.array_equal(got2["flags"], gotg["flags"])
    iq = np.stack([orc.raw_to_complex(r) for r in raw])
    got2c = two.detect_c64(iq, idx)[:, 0]
    assert got2c.tobytes() == got2.tobytes()
    # more blocks than CTAs: every CTA walks several blocks (software pipeline, single raw stage)
    big = raw[np.arange(700) % 40]
    two.close()
    two = NativeDetector(n, h, tpl, len(tpl), (7, 110), (0., 15., 0.), kthresh, max_batch=700)
    gotb = two.detect_raw(big)[:, 0]
    for f in ("flags", "carrier_bin", "corr_sample", "corr_energy", "corr_offset", "carrier_offset"):
        assert np.array_equal(gotb[f][:40], gotb[f][40 * 16:40 * 17], equal_nan=True), f
        assert np.array_equal(gotb[f][:40], got2[f], equal_nan=True), f
    two.close()
    gen.close()


def test_shifted_zoom_band_small_blocks():
    """Pruned FFT#1 with a pre-shifted band at N=4096 and 8192 (negative-frequency windows)."""
    from thrifty_b200._native import NativeDetector
    for n, tpl in ((4096, synth.gold_template(9)), (8192, synth.gold_template(10))):
        hist = len(tpl) + 6
        raw, _ = synth.make_blocks(64, n, hist, tpl, 0.7, seed=2468 + n, bin_range=(-108.0, -9.0))
        st = orc.DetectorSettings(n, hist, len(tpl), (0., 15., 0.), (-110, -7), tpl, (0., 15., 0.))
        ref = orc.detect_blocks(st, raw)
        det = NativeDetector(n, hist, tpl, len(tpl), (-110, -7), (0., 15., 0.), (0., 15., 0.), max_batch=64)
        stats = parity.compare_records(det.detect_raw(raw)[:, 0], ref, what="N=%d window -110..-7" % n)
        assert stats["carrier"] >= 30
        det.close()


@pytest.mark.parametrize("tpl_len,hist", [(1226, 1225), (1226, 1232), (1226, 3000), (700, 699), (333, 1000)])
def test_template_and_history_geometries_with_edge_peaks(tpl_len, hist):
    """Ragged template lengths, minimal history (H = L-1: every lag is in the peak window) and long history;
    bursts forced onto the first / last lag of the peak window and just outside it (soa_estimator.py:20-39,137-143)."""
    from thrifty_b200._native import NativeDetector
    n = 4096
    tpl = synth.gold_template(9)[:tpl_len]
    start, stop = synth.peak_window(n, hist, tpl_len)
    positions = [start, start + 1, stop - 2, stop - 1, max(start - 1, 0), min(stop, n - tpl_len)]
    raws = []
    for b in range(48):
        rng = np.random.default_rng(8800 + 97 * b + tpl_len)
        raw, _ = synth.make_block(rng, n, hist, tpl, 1.0, force_pos=positions[b % len(positions)] if b < 36 else None)
        raws.append(raw)
    raw = np.stack(raws)
    st = orc.DetectorSettings(n, hist, tpl_len, (0., 15., 0.), (7, 110), tpl, (0., 15., 0.))
    with np.errstate(all="ignore"):
        ref = orc.detect_blocks(st, raw)
    det = NativeDetector(n, hist, tpl, tpl_len, (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=64)
    got = det.detect_raw(raw)[:, 0]
    stats = parity.compare_records(got, ref, what="L=%d H=%d" % (tpl_len, hist))
    assert stats["carrier"] >= 40
    edge = (ref["corr_sample"] == start) | (ref["corr_sample"] == stop - 1)
    assert edge.sum() >= 8            # peaks on the window edges were exercised
    det.close()


def test_random_geometries_stress():
    """30 random detector configurations (block length 1024..16384, ragged template / history / carrier lengths,
    positive / negative / far / wide windows, constant + snr + stddev thresholds, weak and hard-clipped blocks)
    against the oracle (tests/stress_parity.py) at the one parity bar of parity_util: 1e-4 on every field of every
    block, short-carrier geometries included."""
    import stress_parity
    assert stress_parity.main(30, 2024) == 0


def test_pageable_inputs_large():
    """Pageable (NumPy / bytes) inputs above the 4-thread staging threshold, with sizes that are not multiples of the
    64-byte copy granule: `.card` text whose lines change length (block index 999 -> 1000), a raw stream, raw blocks."""
    from thrifty_b200._native import NativeDetector
    tpl = np.load(os.path.join(parity.GOLDEN, "template_example.npy"))
    n, h = 16384, 4920
    base, _ = synth.make_blocks(48, n, h, tpl, 0.8, seed=606)
    det = NativeDetector(n, h, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=256)
    # .card text: lines 768..1279 cross the 3 -> 4 digit block index boundary
    idx = np.arange(768, 1280, dtype=np.int64)
    raw = base[np.arange(len(idx)) % 48]
    ref = det.detect_raw(raw, idx)[:, 0]
    buf = io.StringIO()
    block_data.write_card(buf, raw, idx)
    text = buf.getvalue().encode()
    for a, b in ((0, 256), (0, 512), (3, 301)):
        lines = [l + b"\n" for l in text.split(b"\n") if l and not l.startswith(b"#")]
        ts, got_idx, recs, used = det.detect_card(b"".join(lines[a:b]), final=True)
        assert np.array_equal(got_idx, idx[a:b]) and recs[:, 0].tobytes() == ref[a:b].tobytes()
    # raw stream (pageable): 450 blocks, (450 - 1) * 22928 + 32768 bytes
    nblk = 450
    new = 2 * (n - h)
    stream = np.concatenate([base[0][:2 * h]] + [base[b % 48][2 * h:] for b in range(nblk)])
    blocks = np.stack([stream[b * new:b * new + 2 * n] for b in range(nblk)])
    assert len(stream) % 64 != 0
    got = det.detect_stream(stream, 0)[:, 0]
    want = det.detect_raw(blocks, np.arange(nblk))[:, 0]
    assert got.tobytes() == want.tobytes()
    det.close()
