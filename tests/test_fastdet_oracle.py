"""CPU: the native-path (fastdet) oracle against the goldens produced by the reference's own compiled
sources (oracle/make_golden_fastdet.py), and -- when oracle/_ref is present -- against that library."""
import os
import tempfile

import numpy as np
import pytest

import parity_util as parity
from oracle import fastdet_oracle as fo
from thrifty_b200 import block_data


def _oracle_records(cfg, raw, block_idx):
    return fo.detect_blocks(cfg["block_len"], cfg["history_len"], cfg["thresh"], cfg["window"], cfg["template"],
                            cfg["corr_thresh"], raw, block_idx)


@pytest.mark.parametrize("name", parity.FASTDET_GOLDEN_NAMES)
def test_restatement_matches_compiled_reference_goldens(name):
    cfg, raw, block_idx, ref, toads, _ = parity.load_fastdet_golden(name)
    if cfg["block_len"] > 16384:
        raw, block_idx, ref = raw[:6], block_idx[:6], ref[:6]
    mine = _oracle_records(cfg, raw, block_idx)
    assert np.array_equal(mine["carrier_detected"], ref["carrier_detected"])
    car = ref["carrier_detected"] != 0
    assert car.any()
    assert np.array_equal(mine["carrier_argmax"][car], ref["carrier_argmax"][car])
    assert np.array_equal(mine["corr_peak_idx"][car], ref["corr_peak_idx"][car])
    assert np.array_equal(mine["corr_detected"][car], ref["corr_detected"][car])
    for f in ("carrier_max", "carrier_noise", "fft_sum", "corr_peak_power", "corr_noise_power"):
        np.testing.assert_allclose(mine[f][car], ref[f][car], rtol=5e-5)
    np.testing.assert_allclose(mine["corr_offset"][car], ref["corr_offset"][car], atol=2e-4)
    np.testing.assert_allclose(mine["carrier_offset"][car], ref["carrier_offset"][car], atol=2e-4)
    np.testing.assert_allclose(mine["soa"][car], ref["soa"][car], atol=2e-4)


def test_toad_line_format():
    # fastdet/fastdet.cpp:191-206: 12 columns, sqrt of the powers
    cfg, raw, block_idx, ref, toads, _ = parity.load_fastdet_golden("n16384_example")
    det = ref[ref["corr_detected"] != 0]
    assert len(toads) == len(det)
    f = toads[0].split(" ")
    assert len(f) == 12
    assert int(f[2]) == int(det[0]["block_idx"]) and int(f[4]) == int(det[0]["corr_peak_idx"])
    assert abs(float(f[6]) - np.sqrt(det[0]["corr_peak_power"])) < 1e-3 * np.sqrt(det[0]["corr_peak_power"])
    assert abs(float(f[3]) - det[0]["soa"]) < 1e-6


def test_window_rules():
    # fastcard/cardet.c:43-69
    assert fo.normalize_window(7, 110, 4096) == (7, 110)
    assert fo.normalize_window(-110, -7, 4096) == (3986, 4089)
    assert fo.normalize_window(110, 7, 4096) == (7, 110)
    with pytest.raises(ValueError):
        fo.normalize_window(-10, 10, 4096)
    with pytest.raises(ValueError):
        fo.normalize_window(0, 4096, 4096)
    # fastdet/corr_detector.cpp:73-86 == soa_estimator.py:20-39
    assert fo.calculate_window(64, 31, 32) == (0, 33)
    assert fo.calculate_window(64, 33, 32) == (1, 32)


def test_rawconv_matches_python_path():
    # fastcard/rawconv.c:5-28 vs thrifty/block_data.py:38-52: bit-identical
    from oracle import thrifty_oracle as orc
    raw = np.arange(512, dtype=np.uint16).astype(np.uint8)
    raw = np.concatenate([raw, raw[::-1]])
    assert np.array_equal(fo.rawconv(raw).view(np.uint32), orc.raw_to_complex(raw).view(np.uint32))


@pytest.mark.skipif(not fo.have_reference(), reason="oracle/_ref not built (make -C oracle; needs /root/reference)")
def test_compiled_reference_end_to_end_card_and_stream():
    # the reference's own card_reader / raw_reader / cardet / CorrDetector, run here
    for name in ("n4096_gold9_negwin", "n4096_gold9_stream"):
        cfg, raw, block_idx, ref, _, stream = parity.load_fastdet_golden(name)
        with tempfile.TemporaryDirectory() as tmp:
            if cfg["mode"] == "card":
                path = os.path.join(tmp, "in.card")
                with open(path, "w") as f:
                    block_data.write_card(f, raw, block_idx)
            else:
                path = os.path.join(tmp, "in.dat")
                stream.tofile(path)
            got = fo.run_reference(path, cfg["mode"] == "card", cfg["block_len"], cfg["history_len"], cfg["thresh"],
                                   cfg["window"], cfg["template"], cfg["corr_thresh"])
        assert len(got) == len(ref)
        for f in ("block_idx", "carrier_detected", "carrier_argmax", "corr_detected", "corr_peak_idx"):
            assert np.array_equal(got[f], ref[f]), f
        np.testing.assert_array_equal(got["corr_peak_power"], ref["corr_peak_power"])
        assert np.array_equal(fo.reference_rawconv(raw[0]).view(np.uint32), fo.rawconv(raw[0]).view(np.uint32))
