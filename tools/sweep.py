#!/usr/bin/env python
"""Secondary measurements (BASELINE.json configs 3 and 5, signal mixes), device-resident inputs.

    python tools/sweep.py            # prints one JSON line per configuration

Uses only the C ABI (device pool allocated through thr_device_alloc, timing with the library's
CUDA-event timer on the launch stream).  The headline number is bench.py's; this is the table
quoted in DESIGN.md section 7."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from thrifty_b200 import synth  # noqa: E402
from thrifty_b200._native import NativeDetector, RECORD_DTYPE, load_library  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(block_len, templates, history, batch, p_signal, steps=64, warmup=4, window=(7, 110), unique=256,
        label="", fastdet=False, bin_range=(8.0, 109.0), generic=False):
    lib = load_library()
    tpl0 = templates[0] if templates.ndim == 2 else templates
    raw, _ = synth.make_blocks(unique, block_len, history, tpl0, p_signal, seed=424242, bin_range=bin_range)
    n_tpl = templates.shape[0] if templates.ndim == 2 else 1
    det = NativeDetector(block_len, history, templates, len(tpl0), window, (0., 15., 0.), (0., 15., 0.),
                         max_batch=batch, overlap_launches=True, fastdet=fastdet, generic_kernel=generic)
    # pool > L2 (126 MB): at least 160 MiB of raw blocks, a multiple of the batch
    pool_blocks = max(2 * batch, ((160 << 20) // (2 * block_len) + batch - 1) // batch * batch)
    host = np.ascontiguousarray(raw[np.arange(pool_blocks) % unique])
    d_raw = lib.thr_device_alloc(0, host.nbytes)
    d_out = [lib.thr_device_alloc(0, batch * n_tpl * 64) for _ in range(2)]
    assert d_raw and all(d_out)
    assert lib.thr_memcpy_h2d(0, d_raw, host.ctypes.data, host.nbytes) == 0
    windows = pool_blocks // batch

    def step(i):
        off = (i % windows) * batch * 2 * block_len
        det.detect_device(d_raw + off, None, batch, d_out[i & 1])

    for i in range(warmup):
        step(i)
    det.synchronize()
    det.timer_start()
    for i in range(steps):
        step(i)
    ms = det.timer_stop() / steps
    out = np.zeros(batch * n_tpl, dtype=RECORD_DTYPE)
    lib.thr_memcpy_d2h(0, out.ctypes.data, d_out[(steps - 1) & 1], out.nbytes)
    info = det.info()
    blocks_per_s = batch / (ms * 1e-3)
    alg = batch * (2 * block_len + 64 * n_tpl)
    peak = 6533.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    line = dict(label=label, block_len=block_len, n_templates=n_tpl, batch=batch, p_signal=p_signal,
                ms_per_launch=ms, blocks_per_s=blocks_per_s, msamples_per_s=blocks_per_s * block_len / 1e6,
                hbm_gbs_algorithmic=alg / (ms * 1e-3) / 1e9, hbm_frac_of_measured_peak=alg / (ms * 1e-3) / 1e9 / peak,
                kernel=info["kernel"], grid=info["grid"], threads=info["threads"], ctas_per_sm=info["ctas_per_sm"],
                smem_bytes=info["smem_bytes"], carrier=int(((out["flags"] & 1) != 0).sum()),
                detected=int(((out["flags"] & 2) != 0).sum()))
    print(json.dumps(line), flush=True)
    for ptr in [d_raw] + d_out:
        lib.thr_device_free(0, ptr)
    det.close()


def run_card(n_blocks=2048, reps=3):
    """`.card` text -> records: host scan + GPU base64 decode + detect (thr_detect_card) vs the
    host-side decode a Python card_reader does (base64.b64decode + detect_raw)."""
    import io
    import time
    from thrifty_b200 import block_data
    example = np.load(os.path.join(GOLDEN, "template_example.npy"))
    n, hist = 16384, 4920
    raw, _ = synth.make_blocks(128, n, hist, example, 1.0, seed=99)
    raw = raw[np.arange(n_blocks) % 128]
    buf = io.StringIO()
    block_data.write_card(buf, raw)
    text = buf.getvalue().encode()
    det = NativeDetector(n, hist, example, len(example), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=4096)
    from thrifty_b200._native import PinnedBuffer
    pin = PinnedBuffer(len(text))
    pin.array[:] = np.frombuffer(text, dtype=np.uint8)
    det.detect_card_ptr(pin.ptr, len(text))
    t0 = time.perf_counter()
    for _ in range(reps):
        ts, idx, recs, used = det.detect_card_ptr(pin.ptr, len(text))
    dt_gpu = (time.perf_counter() - t0) / reps
    det.detect_card(text)                       # warm-up: page-locked staging is allocated on first use
    t0 = time.perf_counter()
    for _ in range(reps):
        det.detect_card(text)
    dt_pageable = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    blocks = list(block_data.card_reader(io.BytesIO(text), raw=True))
    recs2 = det.detect_raw(np.stack([b[2] for b in blocks]), np.array([b[1] for b in blocks]))
    dt_host = time.perf_counter() - t0
    assert recs2.tobytes() == recs.tobytes()
    print(json.dumps(dict(label="card ingest N=16384 (text -> records)", n_blocks=n_blocks, text_mb=len(text) / 1e6,
                          gpu_decode_blocks_per_s=n_blocks / dt_gpu, gpu_decode_msamples_per_s=n_blocks * n / dt_gpu / 1e6,
                          gpu_decode_text_gbs=len(text) / dt_gpu / 1e9,
                          gpu_decode_pageable_blocks_per_s=n_blocks / dt_pageable,
                          host_decode_blocks_per_s=n_blocks / dt_host,
                          host_decode_msamples_per_s=n_blocks * n / dt_host / 1e6)), flush=True)
    det.close()


def run_stream(n_blocks=16384, reps=3):
    """Raw uint8 I/Q stream in pinned host memory -> records (thr_detect_stream): the kernel reads the
    overlapping windows in place, so only N-H new samples per block cross PCIe (block_reader semantics,
    thrifty/block_data.py:70-98)."""
    import ctypes
    import time
    from thrifty_b200._native import PinnedBuffer
    example = np.load(os.path.join(GOLDEN, "template_example.npy"))
    n, hist = 16384, 4920
    new = 2 * (n - hist)
    raw, _ = synth.make_blocks(128, n, hist, example, 1.0, seed=77)
    det = NativeDetector(n, hist, example, len(example), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=4096)
    nbytes = 2 * hist + n_blocks * new
    pin = PinnedBuffer(nbytes)
    pin.array[:2 * hist] = raw[0][:2 * hist]
    body = np.concatenate([raw[b % 128][2 * hist:] for b in range(n_blocks)])
    pin.array[2 * hist:] = body
    out = PinnedBuffer(n_blocks * 64)
    got = ctypes.c_int64(0)
    lib = det._lib
    det._check(lib.thr_detect_stream(det.handle, pin.ptr, nbytes, 0, out.ptr, ctypes.byref(got)))
    t0 = time.perf_counter()
    for _ in range(reps):
        det._check(lib.thr_detect_stream(det.handle, pin.ptr, nbytes, 0, out.ptr, ctypes.byref(got)))
    dt = (time.perf_counter() - t0) / reps
    recs = out.array.view(RECORD_DTYPE)
    print(json.dumps(dict(label="raw stream N=16384 (pinned host stream -> records, windows read in place)",
                          n_blocks=int(got.value), stream_mb=nbytes / 1e6, blocks_per_s=got.value / dt,
                          msamples_per_s=got.value * n / dt / 1e6, new_msamples_per_s=got.value * (n - hist) / dt / 1e6,
                          h2d_gbs=nbytes / dt / 1e9, carrier=int(((recs["flags"] & 1) != 0).sum()))), flush=True)
    pin.close()
    out.close()
    det.close()


def run_pageable(n_blocks=8192, reps=3):
    """NumPy (pageable) uint8 blocks -> records through thr_detect_batch: what `Detector.detect_many` costs."""
    import time
    example = np.load(os.path.join(GOLDEN, "template_example.npy"))
    n, hist = 16384, 4920
    raw, _ = synth.make_blocks(128, n, hist, example, 1.0, seed=78)
    blocks = np.ascontiguousarray(raw[np.arange(n_blocks) % 128])
    det = NativeDetector(n, hist, example, len(example), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=n_blocks)
    det.detect_raw(blocks[:1024])
    t0 = time.perf_counter()
    for _ in range(reps):
        recs = det.detect_raw(blocks)
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps(dict(label="pageable NumPy blocks N=16384 -> records (thr_detect_batch, internal page-locked staging)",
                          n_blocks=n_blocks, blocks_per_s=n_blocks / dt, msamples_per_s=n_blocks * n / dt / 1e6,
                          h2d_gbs=blocks.nbytes / dt / 1e9, detected=int(((recs["flags"] & 2) != 0).sum()))), flush=True)
    det.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "pageable":
        run_pageable()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "card":
        run_card(n_blocks=4096)
        run_stream()
        run_pageable()
        return
    example = np.load(os.path.join(GOLDEN, "template_example.npy"))
    t9, t10 = synth.gold_template(9), synth.gold_template(10)
    # config 3: block_len sweep, batch 4096, 100 % burst blocks
    run(4096, t9, len(t9) + 6, 4096, 1.0, label="cfg3 N=4096")
    run(8192, t10, len(t10) + 6, 4096, 1.0, label="cfg3 N=8192")
    run(16384, example, 4920, 4096, 1.0, label="cfg3 N=16384 (headline, pruned FFT#1)")
    run(16384, example, 4920, 4096, 1.0, window=(7, 300), label="N=16384 full FFT#1 (window 7-300)")
    run(16384, example, 4920, 4096, 1.0, window=(-110, -7), bin_range=(-108.0, -9.0),
        label="N=16384 window -110..-7 (pruned FFT#1, pre-shifted band)")
    run(32768, example, 4920, 2048, 1.0, steps=16, label="cfg3 N=32768 (2 x 16384 kernel)")
    run(32768, example, 4920, 2048, 1.0, window=(7, 300), label="N=32768 full FFT#1 (window 7-300, 2 x 16384 kernel)")
    g11 = np.stack([synth.gold_template(11, i) for i in range(4)])
    run(32768, g11, g11.shape[1] + 6, 2048, 1.0, steps=8, label="N=32768, 4 Gold-11 templates (2 x 16384 kernel)")
    run(32768, g11, g11.shape[1] + 6, 2048, 1.0, steps=4, generic=True,
        label="N=32768, 4 Gold-11 templates on the generic global-scratch kernel (debug launches only)")
    run(32768, example, 4920, 2048, 1.0, window=(7, 300), steps=8, generic=True,
        label="N=32768 full FFT#1 on the generic global-scratch kernel (debug launches only)")
    # signal mixes at N=16384
    run(16384, example, 4920, 4096, 0.5, label="N=16384 50% burst blocks")
    run(16384, example, 4920, 4096, 0.0, label="N=16384 noise only")
    # fastdet semantics (native twin): 2 transforms per block
    run(16384, example, 4920, 4096, 1.0, label="N=16384 fastdet semantics (2 FFTs/block)", fastdet=True)
    # config 5: 4 Gold templates jointly
    t11 = np.stack([synth.gold_template(11, i) for i in range(4)])
    run(16384, t11, t11.shape[1] + 6, 4096, 1.0, steps=32, label="cfg5 N=16384, 4 Gold templates")


if __name__ == "__main__":
    main()
