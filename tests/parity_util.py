"""Shared parity helpers: golden-config loading and record comparison.

Parity definition (BASELINE.json north_star): flags equal; carrier bin and
correlation peak sample bit-exact; magnitudes / noise within 1e-4 relative; sub-sample
offsets within 1e-4 absolute; SoA within 1e-4 samples.  Blocks whose reference
|peak/threshold - 1| < 1e-3 are 'marginal' (float32-vs-float64 rounding may flip the
verdict) and are excluded from flag equality -- none of the committed goldens has one.
"""
import os
import zlib

import numpy as np

from thrifty_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

RTOL_MAG = 1e-4
ATOL_OFFSET = 1e-4


def template_by_id(tid):
    if tid == "example":
        return np.load(os.path.join(GOLDEN, "template_example.npy"))
    assert tid.startswith("gold")
    bits, idx = tid[4:].split("_")
    return synth.gold_template(int(bits), int(idx))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, "detect_%s.npz" % name))
    cfg = dict(
        name=name, block_len=int(g["block_len"]), history_len=int(g["history_len"]),
        template=template_by_id(str(g["template_id"])), window=tuple(int(v) for v in g["window"]),
        n_blocks=int(g["n_blocks"]), p_signal=float(g["p_signal"]),
        bin_range=tuple(float(v) for v in g["bin_range"]),
        cthresh=tuple(float(v) for v in g["cthresh"]), kthresh=tuple(float(v) for v in g["kthresh"]),
        seed=int(g["seed"]))
    gen_id = str(g["gen_template_id"])
    gen_tpl = np.ones(len(cfg["template"])) if gen_id == "ones" else template_by_id(gen_id)
    raw, _ = synth.make_blocks(cfg["n_blocks"], cfg["block_len"], cfg["history_len"], gen_tpl,
                               cfg["p_signal"], seed=cfg["seed"], bin_range=cfg["bin_range"])
    assert np.uint32(zlib.crc32(raw.tobytes())) == g["raw_crc32"], "synthetic generator drifted"
    block_idx = 10 + 3 * np.arange(cfg["n_blocks"], dtype=np.int64)
    return cfg, raw, block_idx, g["records"], [str(s) for s in g["toad_lines"]]


GOLDEN_NAMES = ["n16384_example", "n8192_gold10", "n4096_gold9", "n4096_gold9_wrapwin_std",
                "n4096_gold9_tone", "n32768_example", "n32768_example_wide_std"]


def compare_records(got, ref, what=""):
    """got: thr_record array [B] (one template); ref: oracle RECORD_DTYPE array [B].  One bar for every block:
    RTOL_MAG on magnitudes, ATOL_OFFSET on offsets and SoA (no per-block widening: the carrier offset comes out of
    the same float64 lmdif iteration the reference runs, thrifty_b200/csrc/dirichlet_lm.cuh)."""
    assert len(got) == len(ref)
    stats = dict(n=len(ref), carrier=0, detected=0, marginal=0, max_rel_corr_energy=0.0,
                 max_abs_corr_offset=0.0, max_abs_carrier_offset=0.0, max_abs_soa=0.0)
    for i in range(len(ref)):
        r, g = ref[i], got[i]
        tag = "%s block %d" % (what, i)
        g_car = bool(g["flags"] & 1)
        g_det = bool(g["flags"] & 2)
        cm, km = r["carrier_margin"], r["corr_margin"]
        marg_c = np.isfinite(cm) and abs(cm - 1) < 1e-3
        marg_k = np.isfinite(km) and abs(km - 1) < 1e-3
        if marg_c or marg_k:
            stats["marginal"] += 1
        assert g["block_idx"] == r["block_idx"], tag
        if not marg_c:
            assert g_car == bool(r["carrier_detected"]), tag + " carrier flag"
        if not r["carrier_detected"] or not g_car:
            continue
        stats["carrier"] += 1
        assert g["carrier_bin"] == r["carrier_bin"], tag + " carrier bin"
        np.testing.assert_allclose(g["carrier_energy"], r["carrier_energy"], rtol=RTOL_MAG, err_msg=tag)
        np.testing.assert_allclose(g["carrier_noise"], r["carrier_noise"], rtol=RTOL_MAG, err_msg=tag)
        np.testing.assert_allclose(g["carrier_offset"], r["carrier_offset"], err_msg=tag + " carrier offset",
                                   atol=ATOL_OFFSET)
        stats["max_abs_carrier_offset"] = max(stats["max_abs_carrier_offset"],
                                              abs(float(g["carrier_offset"]) - r["carrier_offset"]))
        if not marg_k:
            assert g_det == bool(r["corr_detected"]), tag + " corr flag"
        assert g["corr_sample"] == r["corr_sample"], tag + " corr sample"
        rtol_k = RTOL_MAG
        np.testing.assert_allclose(g["corr_energy"], r["corr_energy"], rtol=rtol_k, err_msg=tag + " corr energy")
        if np.isnan(r["corr_noise"]):
            assert np.isnan(g["corr_noise"]), tag
        else:
            np.testing.assert_allclose(g["corr_noise"], r["corr_noise"], rtol=rtol_k, err_msg=tag + " corr noise")
        stats["max_rel_corr_energy"] = max(stats["max_rel_corr_energy"],
                                           abs(float(g["corr_energy"]) / r["corr_energy"] - 1))
        if g_det and r["corr_detected"]:
            stats["detected"] += 1
            np.testing.assert_allclose(g["corr_offset"], r["corr_offset"], atol=ATOL_OFFSET, err_msg=tag)
            stats["max_abs_corr_offset"] = max(stats["max_abs_corr_offset"],
                                               abs(float(g["corr_offset"]) - r["corr_offset"]))
        if g_det == bool(r["corr_detected"]):
            np.testing.assert_allclose(g["soa"], r["soa"], rtol=0, atol=ATOL_OFFSET, err_msg=tag + " soa")
            stats["max_abs_soa"] = max(stats["max_abs_soa"], abs(float(g["soa"]) - r["soa"]))
    return stats


# ------------------------------------------------------------------ fastdet (native twin) semantics
FASTDET_GOLDEN_NAMES = ["n16384_example", "n8192_gold10_const", "n4096_gold9_negwin", "n4096_gold9_stream",
                        "n32768_example"]


def fastdet_stream_blocks(raw, block_len, history_len):
    """(stream bytes, blocks) the reference's raw_reader forms from the NEW parts of `raw`
    (fastcard/raw_reader.c:15-46; the first block's history is uint16 127 per sample, reader.c:56-59)."""
    new = block_len - history_len
    stream = np.concatenate([r[2 * history_len:] for r in raw])
    blocks = np.zeros((len(raw), 2 * block_len), dtype=np.uint8)
    cur = np.zeros(2 * block_len, dtype=np.uint8)
    cur[0::2] = 127
    for b in range(len(raw)):
        cur = np.concatenate([cur[2 * new:], stream[2 * new * b:2 * new * (b + 1)]])
        blocks[b] = cur
    return stream, blocks


def load_fastdet_golden(name):
    """-> (cfg, blocks uint8[B,2N], block_idx, reference records, toad lines, stream or None)."""
    g = np.load(os.path.join(GOLDEN, "fastdet_%s.npz" % name))
    cfg = dict(name=name, block_len=int(g["block_len"]), history_len=int(g["history_len"]),
               template=template_by_id(str(g["template_id"])), window=tuple(int(v) for v in g["window"]),
               n_blocks=int(g["n_blocks"]), thresh=tuple(float(v) for v in g["thresh"]),
               corr_thresh=tuple(float(v) for v in g["corr_thresh"]), mode=str(g["mode"]))
    raw, _ = synth.make_blocks(cfg["n_blocks"], cfg["block_len"], cfg["history_len"], cfg["template"],
                               float(g["p_signal"]), seed=int(g["seed"]),
                               bin_range=tuple(float(v) for v in g["bin_range"]))
    assert np.uint32(zlib.crc32(raw.tobytes())) == g["raw_crc32"], "synthetic generator drifted"
    stream = None
    if cfg["mode"] == "raw":
        stream, raw = fastdet_stream_blocks(raw, cfg["block_len"], cfg["history_len"])
        block_idx = np.arange(cfg["n_blocks"], dtype=np.int64)
    else:
        block_idx = 10 + 3 * np.arange(cfg["n_blocks"], dtype=np.int64)
    return cfg, raw, block_idx, g["records"], [str(s) for s in g["toad_lines"]], stream


def compare_fastdet(got, ref, what="", rtol=RTOL_MAG, atol_off=ATOL_OFFSET):
    """got: thr_record array [B] from a THR_CFG_FASTDET_SEMANTICS detector; ref: fastdet_oracle
    REF_RECORD_DTYPE array [B] (powers; thr_record carries their square roots)."""
    assert len(got) == len(ref)
    stats = dict(n=len(ref), carrier=0, detected=0, marginal=0, max_rel_power=0.0, max_abs_offset=0.0)
    for i in range(len(ref)):
        r, g = ref[i], got[i]
        tag = "%s block %d" % (what, i)
        assert int(g["block_idx"]) == int(r["block_idx"]), tag
        assert bool(g["flags"] & 1) == bool(r["carrier_detected"]), tag + ": carrier flag"
        if not r["carrier_detected"]:
            assert not (g["flags"] & 2), tag
            continue
        stats["carrier"] += 1
        assert int(g["carrier_bin"]) == int(r["carrier_argmax"]), tag + ": carrier bin"
        assert int(g["corr_sample"]) == int(r["corr_peak_idx"]), tag + ": corr sample"
        marginal = r["corr_threshold"] > 0 and abs(r["corr_peak_power"] / r["corr_threshold"] - 1) < 1e-3
        if marginal:
            stats["marginal"] += 1
        else:
            assert bool(g["flags"] & 2) == bool(r["corr_detected"]), tag + ": corr flag"
        for gf, rf in (("carrier_energy", "carrier_max"), ("carrier_noise", "carrier_noise"),
                       ("corr_energy", "corr_peak_power"), ("corr_noise", "corr_noise_power")):
            gv, rv = float(g[gf]) ** 2, float(r[rf])
            rel = abs(gv - rv) / max(abs(rv), 1e-30) if rv != 0 else abs(gv)
            assert rel <= 2 * rtol, "%s: %s^2 %g vs %g" % (tag, gf, gv, rv)      # squares: twice the magnitude bar
            stats["max_rel_power"] = max(stats["max_rel_power"], rel)
        assert abs(float(g["carrier_offset"]) - float(r["carrier_offset"])) <= atol_off, tag + ": carrier offset"
        if bool(g["flags"] & 2) == bool(r["corr_detected"]):
            d = abs(float(g["corr_offset"]) - float(r["corr_offset"]))
            assert d <= atol_off, tag + ": corr offset %g" % d
            stats["max_abs_offset"] = max(stats["max_abs_offset"], d)
            assert abs(float(g["soa"]) - float(r["soa"])) <= atol_off, tag + ": soa"
            stats["detected"] += int(bool(r["corr_detected"]))
    return stats


# ------------------------------------------------------------------ several templates jointly (BASELINE config 5)
def make_multi_blocks(n_blocks, block_len, history_len, templates, p_signal, seed):
    """Block b carries (with probability p_signal) a burst spread with ONE of the templates, chosen by
    default_rng(seed).  -> (raw uint8[B, 2N], which int[B])."""
    rng = np.random.default_rng(seed)
    which = rng.integers(0, len(templates), n_blocks)
    raw = np.empty((n_blocks, 2 * block_len), dtype=np.uint8)
    for b in range(n_blocks):
        raw[b] = synth.make_blocks(1, block_len, history_len, templates[which[b]], p_signal, seed=seed + 1 + b)[0][0]
    return raw, which


def load_multi_golden(name="n16384_gold11x4"):
    """-> (cfg, templates [T, L], raw, block_idx, reference records [T, B])."""
    g = np.load(os.path.join(GOLDEN, "detect_%s.npz" % name))
    tpls = np.stack([synth.gold_template(int(g["gold_bits"]), int(i)) for i in g["gold_idx"]])
    cfg = dict(name=name, block_len=int(g["block_len"]), history_len=int(g["history_len"]),
               window=tuple(int(v) for v in g["window"]), n_blocks=int(g["n_blocks"]), p_signal=float(g["p_signal"]),
               cthresh=tuple(float(v) for v in g["cthresh"]), kthresh=tuple(float(v) for v in g["kthresh"]),
               seed=int(g["seed"]))
    raw, which = make_multi_blocks(cfg["n_blocks"], cfg["block_len"], cfg["history_len"], tpls, cfg["p_signal"], cfg["seed"])
    assert np.uint32(zlib.crc32(raw.tobytes())) == g["raw_crc32"], "synthetic generator drifted"
    block_idx = 10 + 3 * np.arange(cfg["n_blocks"], dtype=np.int64)
    return cfg, tpls, raw, block_idx, g["records"], which
