"""Host-side logic of the drop-in (no GPU): parsers, settings, block I/O, record text."""
import io
import os

import numpy as np
import pytest

from thrifty_b200 import block_data, setting_parsers, settings, stripe, synth, toads_data, util


# grammar tables of /root/reference/tests/test_setting_parsers.py:10-99
def test_freq_range():
    cases = [('100', (100.0, 100.0, False)), ('-123.4', (-123.4, -123.4, False)),
             ('100-200', (100.0, 200.0, False)), ('10e1 - 20e1', (100.0, 200.0, False)),
             ('-100-100', (-100.0, 100.0, False)), ('-200--100', (-200.0, -100.0, False)),
             ('100hz', (100.0, 100.0, True)), ('100-200 Hz', (100.0, 200.0, True)),
             ('10-20 khz', (10000.0, 20000.0, True)), ('1.2345-2.3456 KHZ', (1.2345e3, 2.3456e3, True)),
             ('433-435Mhz', (433e6, 435e6, True)), ('1.2345-2.3456 mhz', (1.2345e6, 2.3456e6, True))]
    for string, expected in cases:
        assert setting_parsers.freq_range(string) == expected
    with pytest.raises(ValueError):
        setting_parsers.freq_range('garbage')


def test_metric_float():
    for string, expected in [('1337.15', 1337.15), ('15.2M', 15200000.0), ('987k', 987000.0),
                             ('55m', 0.055), ('164u ', 164e-6)]:
        assert setting_parsers.metric_float(string) == expected
    for bad in ['garbage', 'x53m', '500A']:
        with pytest.raises(ValueError):
            setting_parsers.metric_float(bad)


def test_threshold():
    cases = [('0', (0.0, 0.0, 0.0)), ('10.2', (10.2, 0.0, 0.0)), ('c', (1.0, 0.0, 0.0)),
             ('11c', (11.0, 0.0, 0.0)), ('100 * constant', (100.0, 0.0, 0.0)), ('snr', (0.0, 1.0, 0.0)),
             ('5.2*snr', (0.0, 5.2, 0.0)), (' 8s ', (0.0, 8.0, 0.0)), ('stddev', (0.0, 0.0, 1.0)),
             ('2.1stddev', (0.0, 0.0, 2.1)), ('8.7*d', (0.0, 0.0, 8.7)), ('10 + 4*snr', (10.0, 4.0, 0)),
             ('40 + 3.8*snr + 2 stddev', (40.0, 3.8, 2.0)), ('1+2s+3d+4+5s+6d', (5.0, 7.0, 9.0)),
             ('c + s + d', (1.0, 1.0, 1.0)), ('15 * snr', (0.0, 15.0, 0.0))]
    for string, expected in cases:
        assert setting_parsers.threshold(string) == expected
    for bad in ['', ' ', '5+', 'junk', '+5*stddev', '*snr', 'stddev*snr', '5 * stdde', '2 * sn', 'const']:
        with pytest.raises(ValueError):
            setting_parsers.threshold(bad)


def test_normalize_freq_range():
    assert setting_parsers.normalize_freq_range((7.0, 110.0, False), 146.48) == (7, 110)
    bin_freq = 2.2e6 / 8192
    assert setting_parsers.normalize_freq_range((-81e3, -79e3, True), bin_freq) == (
        int(-81e3 / bin_freq), int(-79e3 / bin_freq))


@pytest.mark.parametrize("num", [15, 16])
def test_fft_bin(num):
    got = np.array([util.fft_bin(i, num) for i in range(num)])
    np.testing.assert_array_equal(got, np.fft.fftfreq(num, 1. / num))


def test_settings_config_overlay(tmp_path):
    cfg = tmp_path / "detector.cfg"
    cfg.write_text("rxid: 3\nblock_size: 8192   # comment\ncarrier_window: 7 - 110\n"
                   "carrier_threshold: 15 * snr\n\ntemplate: t.npy\n")
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("input")
    keys = ["sample_rate", "block_size", "block_history", "carrier_window", "carrier_threshold",
            "corr_threshold", "template", "rxid"]
    config, args = settings.load_args(parser, keys, argv=["x.card", "-c", str(cfg), "--history", "2461"])
    assert config.rxid == 3 and config.block_size == 8192 and config.block_history == 2461
    assert config.carrier_window == (7.0, 110.0, False)
    assert config.carrier_threshold == (0.0, 15.0, 0.0) and config.corr_threshold == (0.0, 15.0, 0.0)
    assert config.sample_rate == 2.4e6 and config.template == "t.npy"
    assert args.input == "x.card"
    with pytest.raises(settings.SettingKeyError):
        settings.load(config_file=io.StringIO("nonsense: 1\n"))
    with pytest.raises(settings.ConfigSyntaxError):
        settings.parse_kvconfig(io.StringIO("no delimiter here\n"))


def test_card_reader_reference_vector():
    # /root/reference/tests/test_block_data.py:59-71 (non-canonical base64 payloads)
    text = "# Some comments\n# more comments\n1000.5425 10 r0+Om5==\n1000.5442 20 aaaaaa=="
    for stream in (io.StringIO(text), io.BytesIO(text.encode())):
        blocks = list(block_data.card_reader(stream, raw=True))
        assert [b[0] for b in blocks] == [1000.5425, 1000.5442]
        assert [b[1] for b in blocks] == [10, 20]
        assert [tuple(b[2]) for b in blocks] == [(175, 79, 142, 155), (105, 166, 154, 105)]
    cplx = list(block_data.card_reader(io.StringIO(text)))
    assert cplx[0][2].dtype == np.complex64
    assert tuple(block_data.complex_to_raw(cplx[0][2])) == (175, 79, 142, 155)


def test_card_reader_skips_noise_lines():
    text = "Using Volk machine: avx2\nlinux; GNU C++\n\n# c\n1.5 7 AAAA\n"
    blocks = list(block_data.card_reader(io.StringIO(text), raw=True))
    assert len(blocks) == 1 and blocks[0][1] == 7 and tuple(blocks[0][2]) == (0, 0, 0)


def test_card_write_read_roundtrip():
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, size=(5, 2 * 64), dtype=np.uint8)
    buf = io.StringIO()
    block_data.write_card(buf, raw, block_indices=[3, 4, 9, 10, 20])
    buf.seek(0)
    blocks = list(block_data.card_reader(buf, raw=True))
    assert [b[1] for b in blocks] == [3, 4, 9, 10, 20]
    for (_, _, got), want in zip(blocks, raw):
        np.testing.assert_array_equal(got, want)
    # line grammar of fastcard_cli.c:187-192: "<sec>.<usec> <idx> <base64>\n"
    line = block_data.card_line(1480000000.25, 12, raw[0])
    ts, idx, payload = line.rstrip("\n").split(" ")
    assert ts == "1480000000.250000" and idx == "12" and len(payload) == (2 * 64 + 2) // 3 * 4


def test_block_reader_overlap():
    # /root/reference/tests/test_block_data.py:40-56
    blocks = list(block_data.block_reader(io.BytesIO(bytes(range(14))), 3, 1))
    raw = [list(block_data.complex_to_raw(b[2])) for b in blocks]
    assert raw == [[0x7f, 0x7f, 0, 1, 2, 3], [2, 3, 4, 5, 6, 7], [6, 7, 8, 9, 10, 11]]
    assert [b[1] for b in blocks] == [0, 1, 2]
    assert blocks[0][2][0] == 0                       # first block's history is exactly zero
    rawmode = list(block_data.block_reader(io.BytesIO(bytes(range(14))), 3, 1, raw=True))
    assert rawmode[0][2].dtype == np.complex64 and rawmode[0][2][0] == 0
    assert [list(b[2]) for b in rawmode[1:]] == [[2, 3, 4, 5, 6, 7], [6, 7, 8, 9, 10, 11]]


def test_raw_complex_roundtrip():
    every = np.arange(256, dtype=np.uint8)
    np.testing.assert_array_equal(block_data.complex_to_raw(block_data.raw_to_complex(every)), every)
    np.testing.assert_allclose(block_data.raw_to_complex(np.array([0, 0, 127, 128, 255, 255], dtype=np.uint8)),
                               [-0.9953 - 0.9953j, -0.0031 + 0.0047j, 0.9969 + 0.9969j], rtol=1e-2)


def test_toad_serialize_roundtrip():
    res = toads_data.DetectionResult(
        1000.019107, 22, 258721.99332702, toads_data.CarrierSyncInfo(30, 0.4434, np.float32(590.445), np.float32(11.75)),
        toads_data.CorrDetectionInfo(6514, -0.00667, 445.5628, 5.2213), rxid=0)
    line = res.serialize()
    assert line.split()[:4] == ["0", "1000.019107", "22", "258721.99332702"]
    assert len(line.split()) == 12
    back = toads_data.DetectionResult.deserialize(line, with_rxid=True)
    assert back.block == 22 and back.corr_info.sample == 6514 and back.carrier_info.bin == 30
    assert abs(back.soa - res.soa) < 1e-8 and back.rxid == 0
    assert toads_data.DetectionResult.deserialize("1 2 3") is None
    arr = toads_data.toads_array([back], with_ids=False)
    assert arr["sample"][0] == 6514 and arr["carrier_bin"][0] == 30


def test_toad_line_matches_reference_format(golden_dir):
    # a golden .toad line produced by the real reference parses field-for-field
    g = np.load(os.path.join(golden_dir, "detect_n4096_gold9.npz"))
    line = [str(s) for s in g["toad_lines"] if str(s)][0]
    res = toads_data.DetectionResult.deserialize(line, with_rxid=True)
    np.testing.assert_allclose([float(v) for v in res.serialize().split()],
                               [float(v) for v in line.split()], rtol=1e-12)


def test_stripe_bounds():
    for n, w in [(10, 2), (11, 2), (4096, 8), (7, 8), (0, 4), (64 * 2**20, 8)]:
        spans = [stripe.stripe_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        assert max(hi - lo for lo, hi in spans) <= (n + w - 1) // w


def test_gold_templates():
    # Gold code properties (thrifty/gold.py): length 2^n-1, two-valued-ish cross-correlation
    c0, c2 = synth.gold(5, 0), synth.gold(5, 2)
    assert len(c0) == 31 and c0.dtype == bool
    b0, b2 = np.where(c0, 1, -1), np.where(c2, 1, -1)
    auto = [int(np.sum(b0 * np.roll(b0, s))) for s in range(1, 31)]
    assert set(auto) == {-1}                                   # m-sequence autocorrelation
    cross = {int(np.sum(b0 * np.roll(b2, s))) for s in range(31)}
    assert cross <= {-9, -1, 7}                                # Gold three-valued bound
    t = synth.gold_template(9)
    assert len(t) == 1226 and set(np.unique(t)) == {-1.0, 1.0}
    assert len(synth.gold_template(10)) == 2455 and len(synth.gold_template(11)) == 4914


def test_card_scan_native():
    """thr_card_scan (host line scanner of the GPU ingest path) against the reference grammar."""
    import ctypes
    from thrifty_b200 import _native
    lib = _native.load_library()
    n = 48                                        # 2N = 96 bytes -> 128 base64 characters
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, size=(4, 2 * n), dtype=np.uint8)
    buf = io.StringIO()
    block_data.write_card(buf, raw, block_indices=[7, 8, 20, -3], t0=1000.25, dt=0.5)
    text = ("Using Volk machine: avx2\n\n" + buf.getvalue()).encode()
    ts = np.zeros(8); idx = np.zeros(8, dtype=np.int64); off = np.zeros(8, dtype=np.int64)
    nf, used, bad = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()

    def scan(data, final, maxb=8):
        return lib.thr_card_scan(data, len(data), n, final, maxb, ts.ctypes.data, idx.ctypes.data, off.ctypes.data,
                                 ctypes.byref(nf), ctypes.byref(used), ctypes.byref(bad))
    assert scan(text, 1) == 0 and nf.value == 4 and used.value == len(text)
    assert list(idx[:4]) == [7, 8, 20, -3]
    np.testing.assert_allclose(ts[:4], 1000.25 + 0.5 * np.arange(4))
    import base64
    for i in range(4):
        payload = text[off[i]:off[i] + 128]
        np.testing.assert_array_equal(np.frombuffer(base64.b64decode(payload), dtype=np.uint8), raw[i])
    # unterminated last line is left for the next chunk unless final
    cut = text[:-40]
    assert scan(cut, 0) == 0 and nf.value == 3 and cut[used.value:].startswith(b"1001.75")
    assert scan(cut, 1) == -1 and bad.value > 0          # truncated payload at EOF is malformed
    assert scan(text, 1, maxb=2) == 0 and nf.value == 2   # max_blocks respected
    assert scan(b"1.0 5 AAAA\n", 1) == -1                 # wrong payload length (card_reader.c:58-66)
    assert scan(b"garbage line\n", 1) == -1


# ---------------------------------------------------------------- fastdet-compatible front end (host parts)
def test_fastdet_threshold_parser():
    # fastcard/parse.c:54-99
    from thrifty_b200 import fastdet
    assert fastdet.parse_threshold("100c2s") == (100.0, 2.0)
    assert fastdet.parse_threshold("15s") == (0.0, 15.0)
    assert fastdet.parse_threshold("2s100c") == (100.0, 2.0)
    assert fastdet.parse_threshold("42") == (42.0, 0.0)
    assert fastdet.parse_threshold("1.5e2c") == (150.0, 0.0)
    for bad in ("1c2c", "1s2s", "3x", "c"):
        with pytest.raises(ValueError):
            fastdet.parse_threshold(bad)


def test_fastdet_window_parser():
    # fastcard/parse.c:38-52 (sscanf "%d-%d")
    from thrifty_b200 import fastdet
    assert fastdet.parse_carrier_window("7-110") == (7, 110)
    assert fastdet.parse_carrier_window("0--1") == (0, -1)
    assert fastdet.parse_carrier_window("-110--7") == (-110, -7)
    assert fastdet.parse_carrier_window("55") == (55, 55)
    with pytest.raises(ValueError):
        fastdet.parse_carrier_window("a-b")


def test_fastdet_tpl_round_trip(tmp_path):
    # fastdet/corr_detector.cpp:200-228, scripts/npy_to_tpl.py:18-22
    from thrifty_b200 import fastdet
    tpl = np.load(os.path.join(os.path.dirname(__file__), "golden", "template_example.npy"))
    path = str(tmp_path / "t.tpl")
    fastdet.save_template(path, tpl)
    back = fastdet.load_template(path)
    assert back.dtype == np.float32 and len(back) == len(tpl)
    assert np.array_equal(back, tpl.astype(np.float32))
    assert os.path.getsize(path) == 2 + 4 * len(tpl)
    with open(path, "r+b") as f:
        f.truncate(100)
    with pytest.raises(RuntimeError):
        fastdet.load_template(path)


def test_fastdet_toad_line_matches_native_format():
    # fastdet/fastdet.cpp:191-206 vs the compiled reference's records (goldens)
    import parity_util as parity
    from thrifty_b200 import fastdet
    from thrifty_b200._native import RECORD_DTYPE
    cfg, raw, block_idx, ref, toads, _ = parity.load_fastdet_golden("n16384_example")
    r = ref[ref["corr_detected"] != 0][0]
    rec = np.zeros((), dtype=RECORD_DTYPE)
    rec["block_idx"], rec["soa"] = r["block_idx"], r["soa"]
    rec["corr_sample"], rec["corr_offset"] = r["corr_peak_idx"], r["corr_offset"]
    rec["corr_energy"], rec["corr_noise"] = np.sqrt(r["corr_peak_power"]), np.sqrt(r["corr_noise_power"])
    rec["carrier_bin"], rec["carrier_offset"] = r["carrier_argmax"], r["carrier_offset"]
    rec["carrier_energy"], rec["carrier_noise"] = np.sqrt(r["carrier_max"]), np.sqrt(r["carrier_noise"])
    rec["flags"] = 3
    line = fastdet.toad_line(rec, r["ts_sec"] + r["ts_usec"] * 1e-6, 0)
    mine, gold = line.split(" "), toads[0].split(" ")
    assert len(mine) == 12 and mine[0] == gold[0] and mine[2] == gold[2] and mine[4] == gold[4] and mine[8 + 0] == gold[8]
    for a, b in zip(mine[5:], gold[5:]):
        assert abs(float(a) - float(b)) <= 1e-4 * max(1.0, abs(float(b)))
    assert "carrier @" in fastdet.info_line(rec, (0., 15.), (0., 15.))


def test_records_to_results_matches_single_record_conversion():
    """The batch conversion of thr_records (used by every iterator of the Detector seam) yields the same values
    and field types as the per-record one: carrier energy / noise numpy float32 (reference: float32 spectrum),
    the rest Python numbers, corr_info / soa None without a carrier, offsets 0 below threshold (detect.py:60-78)."""
    from thrifty_b200 import detect
    from thrifty_b200._native import RECORD_DTYPE
    rng = np.random.default_rng(5)
    n = 500
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    f = rng.integers(0, 3, n)
    recs["flags"] = np.where(f == 2, 3, f)
    recs["soa"] = rng.random(n) * 1e7
    for k in ("carrier_offset", "carrier_energy", "carrier_noise", "corr_offset", "corr_energy", "corr_noise"):
        recs[k] = rng.random(n).astype(np.float32)
    recs["carrier_noise"][::7] = np.nan
    recs["block_idx"] = np.arange(n) + 10
    recs["carrier_bin"] = rng.integers(0, 16384, n)
    recs["corr_sample"] = rng.integers(0, 11000, n)
    ts = rng.random(n)

    def key(detected, r):
        def tv(x):
            return type(x).__name__, repr(x)
        return (detected, tv(r.timestamp), tv(r.block), tv(r.soa), tuple(tv(x) for x in r.carrier_info),
                None if r.corr_info is None else tuple(tv(x) for x in r.corr_info), r.rxid, r.txid)

    one = [detect.record_to_result(recs[i], float(ts[i]), 3) for i in range(n)]
    many = detect.records_to_results(recs, ts, 3)
    assert [key(*p) for p in one] == [key(*p) for p in many]
    assert detect.records_to_results(recs[:0], ts[:0], 3) == []
    assert [r.timestamp for _, r in detect.records_to_results(recs[:3], 1.5, 3)] == [1.5] * 3
    assert [r.serialize() for d, r in one if d] == [r.serialize() for d, r in many if d]


def test_bench_numa_binding_is_a_no_op_without_locality_information():
    """bench.py binds multi-GPU ranks to their GPU's NUMA node on a best-effort basis: with no GPU / a single node /
    no sysfs entry it must return None and leave the CPU affinity alone."""
    import bench
    before = os.sched_getaffinity(0)
    assert bench.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


# ---- Detector.detect_raw_stream host logic with a fake native layer (no GPU): block count, block indices and the exact
# ---- byte windows handed to thr_detect_stream, including history_len > block_len / 2 (negative slice)
class _FakeStreamNative(object):
    def __init__(self, block_len, history_len, expected_blocks):
        self.block_len, self.history_len, self.n_templates = block_len, history_len, 1
        self.expected = expected_blocks          # block index -> complex64 block (reference block_reader semantics)
        self.calls = []

    def detect_stream(self, stream, first_block):
        from thrifty_b200._native import RECORD_DTYPE
        from thrifty_b200.block_data import raw_to_complex
        n, h = self.block_len, self.history_len
        new = 2 * (n - h)
        nblk = (len(stream) - 2 * n) // new + 1
        assert nblk >= 1 and (len(stream) - 2 * n) % new == 0
        out = np.zeros((nblk, 1), dtype=RECORD_DTYPE)
        for k in range(nblk):
            got = raw_to_complex(stream[k * new:k * new + 2 * n])
            np.testing.assert_array_equal(got, self.expected[first_block + k])
            out[k, 0]["block_idx"] = first_block + k
        self.calls.append((first_block, nblk))
        return out


@pytest.mark.parametrize("n,h,nblk,chunk", [(64, 16, 10, 4), (64, 40, 10, 3), (64, 50, 12, 4), (64, 63, 9, 2), (64, 0, 5, 2),
                                            (64, 32, 7, 100)])
@pytest.mark.parametrize("read_size", [None, 37])
def test_raw_stream_windows_and_indices(n, h, nblk, chunk, read_size):
    import io
    from collections import namedtuple
    from oracle import thrifty_oracle as orc
    from thrifty_b200 import detect as det_mod
    rng = np.random.default_rng(n * 1000 + h)
    new = n - h
    data = rng.integers(0, 256, size=2 * (nblk * new) + min(11, 2 * new - 1), dtype=np.uint8).tobytes()   # + a partial block
    expected = {i: b for i, b in orc.block_reader(io.BytesIO(data), n, h)}
    assert len(expected) == nblk
    d = object.__new__(det_mod.Detector)
    St = namedtuple("St", "block_len history_len")
    d.settings, d.rxid, d.batch = St(n, h), 7, chunk
    d.native = _FakeStreamNative(n, h, expected)
    seen = []

    def fake_detect(timestamp, block_idx, block):       # the complex path (zero history)
        np.testing.assert_array_equal(np.asarray(block, dtype=np.complex64), expected[block_idx])
        seen.append(block_idx)
        return False, det_mod.toads_data.DetectionResult(timestamp, block_idx, None, det_mod.toads_data.CarrierSyncInfo(0, 0, 0., 0.), None, 7)
    d.detect = fake_detect

    class Slow(io.BytesIO):                              # a pipe that hands out a few bytes at a time
        def read1(self, size=-1):
            return io.BytesIO.read(self, min(size, read_size) if read_size else size)
    results = list(d.detect_raw_stream(Slow(data), chunk_blocks=chunk))
    assert [r.block for _, r in results] == list(range(nblk))
    n_zero_hist = min(nblk, -(-h // new)) if h else 0
    assert seen == list(range(n_zero_hist))              # only blocks reaching before the stream take the complex path
    assert sum(c[1] for c in d.native.calls) == nblk - n_zero_hist
    assert all(c[1] <= chunk for c in d.native.calls)


def test_card_stream_short_prefix_lines_are_not_dropped():
    """lines with short '<t> <i>' prefixes ('1.0 3 <payload>') used to overflow the max_blocks estimate and the
    rest of the chunk was dropped on the final call.  Host logic only: a fake native layer that consumes a limited number
    of lines per call."""
    import io
    from collections import namedtuple
    from thrifty_b200 import detect as det_mod
    from thrifty_b200._native import RECORD_DTYPE
    n = 12
    pay = "A" * (((2 * n + 2) // 3) * 4)
    lines = ["%d.0 %d %s\n" % (i, i, pay) for i in range(57)]
    text = ("# header\n" + "".join(lines)).encode()

    class FakeNative(object):
        n_templates = 1

        def __init__(self):
            self.calls = 0

        def detect_card_ptr(self, ptr, length, final=True, max_blocks=None):
            import ctypes
            raw = ctypes.string_at(ptr, length)
            self.calls += 1
            done, ts, idx = 0, [], []
            for ln in raw.split(b"\n")[:-1] if not raw.endswith(b"\n") or True else []:
                if done + len(ln) + 1 > length:
                    break
                if len(idx) >= 5:                        # consumes at most 5 lines per call
                    break
                done += len(ln) + 1
                if ln.startswith(b"#") or not ln:
                    continue
                t, i, _ = ln.split(b" ")
                ts.append(float(t))
                idx.append(int(i))
            recs = np.zeros((len(idx), 1), dtype=RECORD_DTYPE)
            recs["block_idx"][:, 0] = idx
            return np.array(ts), np.array(idx, dtype=np.int64), recs, done

    d = object.__new__(det_mod.Detector)
    St = namedtuple("St", "block_len history_len")
    d.settings, d.rxid, d.batch = St(n, 4), 0, 4
    d.native = FakeNative()
    orig = det_mod.__dict__.get("PinnedBuffer")

    class FakePinned(object):
        def __init__(self, nbytes):
            self.array = np.zeros(nbytes, dtype=np.uint8)
            self.ptr = self.array.ctypes.data

        def close(self):
            pass
    import thrifty_b200._native as nat
    saved = nat.PinnedBuffer
    nat.PinnedBuffer = FakePinned
    try:
        out = list(d.detect_card_stream(io.BytesIO(text), chunk_bytes=1))      # clamps to 4 lines' worth
    finally:
        nat.PinnedBuffer = saved
    assert [r.block for _, r in out] == list(range(57))


def test_native_toad_text_equals_python_serialize():
    """thr_format_toad (host code of the library, no GPU involved) writes, character for character, what
    DetectionResult.serialize() (toads_data.py:47-61) gives for the result objects built from the same records --
    including numpy.float32's own switch to exponent notation at 1e6, nan / inf, zero, and .toads lines with a txid."""
    from thrifty_b200 import _native
    from thrifty_b200.detect import records_to_results
    rng = np.random.default_rng(5)
    n = 30000
    recs = np.zeros(n, dtype=_native.RECORD_DTYPE)
    recs["flags"] = rng.choice([0, 1, 3], n, p=[0.1, 0.1, 0.8])
    recs["block_idx"] = rng.integers(-5, 1 << 40, n)

    def mix(dt):
        v = np.exp(rng.uniform(-25, 40, n)) * rng.choice([-1, 1], n)
        sel = rng.random(n) < 0.1
        v[sel] = np.round(v[sel])
        v[rng.random(n) < 0.02] = 0
        return v.astype(dt)
    recs["soa"] = mix(np.float64) * 1e3
    for f in ("carrier_offset", "carrier_energy", "carrier_noise", "corr_offset", "corr_energy", "corr_noise"):
        recs[f] = mix(np.float32)
    edge = np.array([1e-4, 1e6, 999999.94, 1e-5, 1e16, 9.999999e15, 123.0, 0.5, 1e22, 2.5e-5, np.inf, -np.inf, np.nan])
    for f in ("carrier_offset", "carrier_energy", "corr_energy", "corr_noise"):
        recs[f][:len(edge)] = edge.astype(np.float32)
    recs["soa"][:len(edge)] = edge
    recs["flags"][:len(edge)] = 3
    recs["carrier_bin"] = rng.integers(0, 16384, n)
    recs["corr_sample"] = rng.integers(0, 16384, n)
    ts = 1.48e9 + np.cumsum(rng.uniform(0, 0.01, n))
    results = [r for d, r in records_to_results(recs, ts, 7) if d]
    want = "".join(r.serialize() + "\n" for r in results).encode()
    assert _native.format_toad(recs, ts, 7) == want
    assert _native.format_toad(np.stack([recs, recs], axis=1), ts, 7) == want        # [B, T]: template 0
    tx = rng.integers(-1, 5, n).astype(np.int32)
    for r, t in zip(results, tx[(recs["flags"] & 2) != 0]):
        r.txid = int(t)
    assert _native.format_toad(recs, ts, 7, txids=tx) == "".join(r.serialize() + "\n" for r in results).encode()
    assert _native.format_toad(recs[:0], ts[:0], 7) == b""


def test_staging_cache_reuses_buffers_and_is_bounded(monkeypatch):
    """The stream readers' page-locked staging buffers are handed back to a per-process cache (pinning is slow and
    cudaFreeHost occasionally stalls): same size -> same buffer again; at most _STAGING_KEEP per size and
    _STAGING_MAX_BYTES in total stay cached, the rest is freed."""
    import thrifty_b200._native as nat

    class Fake(object):
        def __init__(self, nbytes):
            self.nbytes, self.ptr, self.array, self.closed = nbytes, 1, None, False

        def close(self):
            self.closed, self.ptr = True, None

    monkeypatch.setattr(nat, "PinnedBuffer", Fake)
    monkeypatch.setattr(nat, "_staging_pool", {})
    a = nat.acquire_staging(100)
    nat.release_staging(a)
    assert nat.acquire_staging(100) is a and not a.closed
    assert nat.acquire_staging(101) is not a
    bufs = [nat.acquire_staging(64) for _ in range(nat._STAGING_KEEP + 2)]
    for b in bufs:
        nat.release_staging(b)
    assert sum(b.closed for b in bufs) == 2 and len(nat._staging_pool[64]) == nat._STAGING_KEEP
    monkeypatch.setattr(nat, "_STAGING_MAX_BYTES", 64 * nat._STAGING_KEEP + 150)
    big = [nat.acquire_staging(100) for _ in range(2)]
    for b in big:
        nat.release_staging(b)
    assert [b.closed for b in big] == [False, True]             # the second one would exceed the byte bound
    nat.release_staging(None)                                   # harmless
