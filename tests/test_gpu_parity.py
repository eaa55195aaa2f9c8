"""GPU parity: CUDA path (through the C ABI) vs golden outputs of the real reference."""
import numpy as np
import pytest

import parity_util as parity

pytestmark = pytest.mark.gpu


def _detector(cfg, **kw):
    from thrifty_b200._native import NativeDetector
    return NativeDetector(cfg["block_len"], cfg["history_len"], cfg["template"], len(cfg["template"]),
                          cfg["window"], cfg["cthresh"], cfg["kthresh"], device=0,
                          max_batch=kw.get("max_batch", 64))


@pytest.mark.parametrize("name", parity.GOLDEN_NAMES)
def test_golden_u8(name):
    cfg, raw, block_idx, ref, _ = parity.load_golden(name)
    det = _detector(cfg)
    got = det.detect_raw(raw, block_idx)[:, 0]
    stats = parity.compare_records(got, ref, what=name)
    print(name, stats)
    assert stats["carrier"] == int(ref["carrier_detected"].sum())
    det.close()


@pytest.mark.parametrize("name", ["n16384_example", "n4096_gold9"])
def test_golden_c64(name):
    from oracle.thrifty_oracle import raw_to_complex
    cfg, raw, block_idx, ref, _ = parity.load_golden(name)
    det = _detector(cfg)
    iq = np.stack([raw_to_complex(r) for r in raw])
    got = det.detect_c64(iq, block_idx)[:, 0]
    parity.compare_records(got, ref, what=name + "/c64")
    det.close()
