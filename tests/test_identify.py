"""`identify` (thrifty/identify.py:26-257): the CUDA path (thrifty_b200/identify.py over csrc/identify.cu, -m gpu) and its
checker (oracle/identify_oracle.py, CPU) against goldens produced by the reference's own functions."""
import io
import os

import numpy as np
import pytest

from oracle import identify_oracle
from thrifty_b200 import toads_data

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "identify_cases.npz")


def _cuda():
    from thrifty_b200 import identify as cuda_identify
    return cuda_identify


IMPLS = [pytest.param(lambda: identify_oracle, id="oracle"), pytest.param(_cuda, id="cuda", marks=pytest.mark.gpu)]


def _cases():
    g = np.load(GOLDEN)
    for c in range(int(g["n_cases"])):
        dets = [toads_data.DetectionResult.deserialize(str(l), with_rxid=True, with_txid=True) for l in g["lines_%d" % c]]
        yield dets, g["mask_%d" % c], g["bins_%d" % c], g["edges_%d" % c]


@pytest.mark.parametrize("impl", IMPLS)
def test_duplicates_mask_matches_reference(impl):
    identify = impl()
    # identify.py:136-166
    for dets, mask, _, _ in _cases():
        assert np.array_equal(identify.identify_duplicates(dets), mask)
        kept = identify.filter_duplicates(dets)
        assert len(kept) == int(mask.sum())
        assert all(a.timestamp <= b.timestamp for a, b in zip(kept, kept[1:]))
        assert all(d.txid != -1 for d in kept)


@pytest.mark.parametrize("impl", IMPLS)
def test_transmitter_windows_match_reference(impl):
    identify = impl()
    # identify.py:26-77
    for _, _, bins, edges in _cases():
        assert np.array_equal(identify.detect_transmitter_windows(bins), edges)


@pytest.mark.parametrize("impl", IMPLS)
def test_auto_classification_groups_by_carrier_bin(impl):
    identify = impl()
    # identify.py:80-103: three transmitters around bins 20 / 55 / 90 at each receiver
    dets, _, _, _ = next(_cases())
    truth = [d.txid for d in dets]
    identify.identify_transmitters(dets, None, verbose=False)
    for d, t in zip(dets, truth):
        if t != -1:
            assert d.txid == t


@pytest.mark.parametrize("impl", IMPLS)
def test_freqmap_classification_and_toads_file(impl, tmp_path):
    identify = impl()
    # identify.py:106-118,191-234
    freqmap = identify.load_freqmap(io.StringIO("0: 17 - 23\n1: 52 - 58\n2: 87 - 93.5\n@0: 0\n@1: 0.25\n"))
    assert freqmap[1][2] == (87.25, 93.75) and set(freqmap) == {0, 1}
    dets, _, _, _ = next(_cases())
    by_rx = {0: [], 1: []}
    for d in dets:
        by_rx[d.rxid].append(d)
    for rx, lst in by_rx.items():
        with open(str(tmp_path / ("rx%d.toad" % rx)), "w") as f:
            for d in lst:
                d.txid = None
                f.write(d.serialize() + "\n")
    out = io.StringIO()
    kept = identify.generate_toads(out, [str(tmp_path / "rx*.toad")], freqmap, verbose=False)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("# source_files: [") and len(lines) == len(kept) + 1
    back = toads_data.load_toads(io.StringIO(out.getvalue()))
    assert [d.txid for d in back] == [d.txid for d in kept]
    for d in back:
        lo, hi = freqmap[d.rxid][d.txid]
        assert lo <= d.carrier_info.bin + d.carrier_info.offset <= hi


@pytest.mark.gpu
def test_cuda_identify_equals_oracle_on_detect_output():
    """Records of one receiver straight from the detect kernel (a burst that straddles two blocks is detected in both:
    the duplicate filter's reason to exist) -> identify on the device, without result objects, == the oracle on the
    DetectionResult list built from the same records; random multi-receiver columns as well."""
    from thrifty_b200 import identify, synth
    from thrifty_b200._native import NativeDetector
    from thrifty_b200.detect import records_to_results
    tpl = synth.gold_template(9)
    n, h = 4096, len(tpl) + 6
    new = n - h
    rng = np.random.default_rng(12)
    nblk = 300
    total = nblk * new + h
    x = 0.02 * (rng.standard_normal(total) + 1j * rng.standard_normal(total))
    pos = 700
    while pos + len(tpl) < total:
        f = rng.choice([20.3, 55.1, 90.4]) + rng.uniform(-1, 1)
        t = np.arange(len(tpl))
        x[pos:pos + len(tpl)] += 0.3 * (tpl + 1) / 2 * np.exp(2j * np.pi * f * (t + pos) / n)
        pos += int(rng.uniform(1.2, 2.5) * new)
    stream = synth.complex_to_raw(x)
    det = NativeDetector(n, h, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 10., 0.), max_batch=512)
    recs = det.detect_stream(stream, 1)[:, 0]
    det.close()
    ts = 1000.0 + 0.0047767 * np.arange(len(recs))
    results = [r for d, r in records_to_results(recs, ts, 3) if d]
    assert len(results) > 100
    for freqmap in (None, {3: {0: (17.0, 23.5), 1: (52.0, 58.5), 2: (87.0, 93.9)}}):
        want = [d.txid for d in _integrate(identify_oracle, results, freqmap)]
        want_blocks = [d.block for d in _integrate(identify_oracle, results, freqmap)]
        sel, txids = identify.integrate_records(recs, ts, 3, freqmap)
        assert recs["block_idx"][sel].tolist() == want_blocks and txids.tolist() == want
        assert len(sel) < len(results)               # some duplicates were dropped
        got = _integrate(identify, results, freqmap)
        assert [(d.block, d.txid) for d in got] == list(zip(want_blocks, want))
    # random columns: several receivers, shuffled order, unidentified entries, ties in block
    for seed in range(3):
        r = np.random.default_rng(seed)
        m = 5000 + 777 * seed
        cols = identify.Columns(r.integers(0, 3, m), r.integers(0, m // 3, m), 1000 + r.permutation(m) * 1e-3,
                                r.uniform(1, 100, m), r.integers(10, 100, m), r.uniform(-0.5, 0.5, m), r.integers(-1, 4, m))
        keep = identify.duplicates_mask_columns(cols)
        arr = np.zeros(m, dtype=[("rxid", "i4"), ("txid", "i4"), ("block", "i4"), ("timestamp", "f8"), ("energy", "f8")])
        for f in ("rxid", "txid", "block", "timestamp", "energy"):
            arr[f] = getattr(cols, f)
        idx = np.argsort(arr[["rxid", "txid", "block", "timestamp"]])
        cur = arr[idx]
        prev, nxt = np.roll(cur, 1), np.roll(cur, -1)
        mask = ~(((cur["block"] == prev["block"] + 1) & (cur["energy"] < prev["energy"]))
                 | ((cur["block"] == nxt["block"] - 1) & (cur["energy"] < nxt["energy"])) | (cur["txid"] == -1))
        assert np.array_equal(keep, mask[np.argsort(idx)])


def _integrate(impl, results, freqmap):
    import copy
    dets = copy.deepcopy(results)
    return impl.integrate(dets, freqmap, verbose=False)
