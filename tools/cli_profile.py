"""Where does the wall clock of `detect x.card -o x.toad --quiet` go?  Times the pieces of the CLI fast path one by one
(page-locked allocation, parallel pread out of the page cache, thr_detect_card, thr_format_toad, handle create / destroy)
and the whole command on a synthetic .card in /dev/shm.   python tools/cli_profile.py [blocks]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from thrifty_b200 import block_data, synth  # noqa: E402
from thrifty_b200._native import NativeDetector, PinnedBuffer, format_toad  # noqa: E402
from thrifty_b200.detect import Detector, _pread_full, detector_cli  # noqa: E402

tpl = np.load("tests/golden/template_example.npy")
raw, _ = synth.make_blocks(256, 16384, 4920, tpl, 1.0, seed=1)
tmp = "/dev/shm/clip"
os.makedirs(tmp, exist_ok=True)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
t0 = time.perf_counter()
with open(tmp + "/x.card", "w") as f:
    block_data.write_card(f, raw[np.arange(nb) % 256])
print("wrote %d lines in %.2f s" % (nb, time.perf_counter() - t0))
np.save(tmp + "/t.npy", tpl)
open(tmp + "/d.cfg", "w").write("block_size: 16384\nblock_history: 4920\ncarrier_window: 7 - 110\ncarrier_threshold: 15*snr\n"
                                "corr_threshold: 15*snr\ntemplate: %s/t.npy\n" % tmp)
argv = [tmp + "/x.card", "-c", tmp + "/d.cfg", "-o", tmp + "/x.toad", "--quiet", "--batch", "4096"]
detector_cli(Detector, argv=argv)
for _ in range(2):
    t0 = time.perf_counter()
    detector_cli(Detector, argv=argv)
    dt = time.perf_counter() - t0
    print("whole command: %.3f s for %d blocks = %.0f blocks/s" % (dt, nb, nb / dt))

chunk = 32 << 20
t0 = time.perf_counter(); buf = PinnedBuffer(chunk + 1); print("PinnedBuffer(32 MiB): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
fd = os.open(tmp + "/x.card", os.O_RDONLY)
for threads in (1, 4, 8, 16):
    pool = ThreadPoolExecutor(threads)
    part = -(-chunk // threads)
    t0 = time.perf_counter()
    futs = [pool.submit(_pread_full, fd, memoryview(buf.array)[o:min(o + part, chunk)], o) for o in range(0, chunk, part)]
    got = sum(f.result() for f in futs)
    dt = time.perf_counter() - t0
    print("pread 32 MiB with %2d threads: %.1f ms = %.1f GB/s" % (threads, dt * 1e3, got / dt / 1e9))
    pool.shutdown()
t0 = time.perf_counter()
det = NativeDetector(16384, 4920, tpl, len(tpl), (7, 110), (0., 15., 0.), (0., 15., 0.), max_batch=4096)
print("thr_create: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
cut = int(np.flatnonzero(buf.array[:chunk] == 10)[-1]) + 1
for _ in range(3):
    t0 = time.perf_counter()
    ts, idx, recs, consumed = det.detect_card_ptr(buf.ptr, cut, final=True)
    dt = time.perf_counter() - t0
    print("thr_detect_card on %d lines (%.1f MB): %.2f ms = %.0f blocks/s" % (len(idx), cut / 1e6, dt * 1e3, len(idx) / dt))
t0 = time.perf_counter(); text = format_toad(recs, ts, 0); dt = time.perf_counter() - t0
print("format_toad %d lines: %.2f ms" % (len(idx), dt * 1e3))
t0 = time.perf_counter(); det.close(); print("thr_destroy: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); buf.close(); print("PinnedBuffer.close: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
