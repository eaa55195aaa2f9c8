"""Several GPUs behind one handle (thr_group_*, thrifty_b200._native.NativeGroup, Detector(devices=[...])): every batch is
cut into contiguous stripes, one per device; the records must come back in input order and be byte-identical to what one
GPU returns.  Runs with however many devices the box has (the 1-GPU box exercises the group plumbing with one member; the
driver's multi-GPU boxes exercise ragged stripes across 2..8 members)."""
import io

import numpy as np
import pytest

from thrifty_b200 import block_data, synth
from thrifty_b200._native import NativeDetector, NativeGroup, load_library

pytestmark = pytest.mark.gpu

N, BITS = 4096, 9


def _setup():
    tpl = synth.gold_template(BITS)
    return tpl, len(tpl) + 6


def _device_sets():
    n = load_library().thr_device_count()
    sets = [[0]]
    if n >= 2:
        sets.append([0, 1])
    if n >= 3:
        sets.append(list(range(n)))
        sets.append([n - 1, 0, 1])                     # any order, any subset
    return sets


@pytest.mark.parametrize("devices", _device_sets())
@pytest.mark.parametrize("nblk", [1, 7, 333])
def test_group_blocks_equal_single_gpu(devices, nblk):
    tpl, h = _setup()
    raw, _ = synth.make_blocks(min(nblk, 48), N, h, tpl, 0.7, seed=31)
    raw = raw[np.arange(nblk) % len(raw)]
    idx = 5 + 3 * np.arange(nblk, dtype=np.int64)
    single = NativeDetector(N, h, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=128)
    group = NativeGroup(devices, N, h, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=128)
    want = single.detect_raw(raw, idx)
    got = group.detect_raw(raw, idx)
    assert got.tobytes() == want.tobytes()
    assert group.detect_raw(raw).tobytes() == single.detect_raw(raw).tobytes()        # default indices are global
    assert group.info()["devices"] == devices and len(group.numa_nodes()) == len(devices)
    group.close()
    single.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_group_stream_and_card_equal_single_gpu(devices):
    tpl, h = _setup()
    nblk = 101
    base, _ = synth.make_blocks(24, N, h, tpl, 0.8, seed=77)
    stream = np.concatenate([base[0][:2 * h]] + [base[b % 24][2 * h:] for b in range(nblk)])
    single = NativeDetector(N, h, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=64)
    group = NativeGroup(devices, N, h, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=64)
    assert group.detect_stream(stream, 3).tobytes() == single.detect_stream(stream, 3).tobytes()
    buf = io.StringIO()
    idx = 1000 + 7 * np.arange(nblk)
    block_data.write_card(buf, base[np.arange(nblk) % 24], block_indices=idx)
    text = buf.getvalue().encode()
    ts1, i1, r1, c1 = single.detect_card(text)
    ts2, i2, r2, c2 = group.detect_card_ptr(text, len(text))
    assert c1 == c2 == len(text) and np.array_equal(i1, i2) and np.array_equal(ts1, ts2)
    assert r1.tobytes() == r2.tobytes()
    group.close()
    single.close()


def test_detector_devices_argument_and_cli(tmp_path):
    """Detector(devices=[...]) and `detect --devices`: same results, same order as one GPU."""
    from thrifty_b200.detect import Detector, DetectorSettings, detector_cli, parse_devices
    assert parse_devices("0-3") == [0, 1, 2, 3] and parse_devices("0,2") == [0, 2] and parse_devices("1") == [1]
    n_dev = load_library().thr_device_count()
    devices = list(range(min(n_dev, 4)))
    tpl, h = _setup()
    raw, _ = synth.make_blocks(37, N, h, tpl, 0.7, seed=5)
    st = DetectorSettings(N, h, len(tpl), (0., 15., 0.), (7, 110), tpl, (0., 15., 0.))
    items = [(1.0 + i, i, raw[i]) for i in range(len(raw))]
    one = Detector(st, rxid=1, batch=16)
    many = Detector(st, rxid=1, batch=16, devices=devices)
    a, b = one.detect_many(items), many.detect_many(items)
    assert [(d, r.serialize() if d else r.block) for d, r in a] == [(d, r.serialize() if d else r.block) for d, r in b]
    one.close()
    many.close()
    # command line
    np.save(tmp_path / "t.npy", tpl)
    with open(tmp_path / "x.card", "w") as f:
        block_data.write_card(f, raw)
    (tmp_path / "d.cfg").write_text("block_size: %d\nblock_history: %d\ncarrier_window: 7 - 110\ncarrier_threshold: 15*snr\n"
                                    "corr_threshold: 15*snr\ntemplate: %s\n" % (N, h, tmp_path / "t.npy"))
    outs = []
    for extra in ([], ["--devices", ",".join(str(d) for d in devices)]):
        out = tmp_path / ("o%d.toad" % len(outs))
        detector_cli(Detector, argv=[str(tmp_path / "x.card"), "-c", str(tmp_path / "d.cfg"), "-o", str(out), "--quiet"] + extra)
        outs.append(out.read_text())
    assert outs[0] == outs[1] and outs[0].count("\n") > 10
