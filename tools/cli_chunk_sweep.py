"""Wall clock of `detect x.card -o x.toad --quiet` (in-process, 16384 lines of N=16384 in /dev/shm) against the size of the
two page-locked staging buffers (THRIFTY_B200_CARD_CHUNK_MB):   python tools/cli_chunk_sweep.py [lines]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from thrifty_b200 import block_data, synth  # noqa: E402
from thrifty_b200.detect import Detector, detector_cli  # noqa: E402

tpl = np.load("tests/golden/template_example.npy")
raw, _ = synth.make_blocks(256, 16384, 4920, tpl, 1.0, seed=1)
tmp = "/dev/shm/clip"
os.makedirs(tmp, exist_ok=True)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
with open(tmp + "/x.card", "w") as f:
    block_data.write_card(f, raw[np.arange(nb) % 256])
np.save(tmp + "/t.npy", tpl)
open(tmp + "/d.cfg", "w").write("block_size: 16384\nblock_history: 4920\ncarrier_window: 7 - 110\ncarrier_threshold: 15*snr\n"
                                "corr_threshold: 15*snr\ntemplate: %s/t.npy\n" % tmp)
argv = [tmp + "/x.card", "-c", tmp + "/d.cfg", "-o", tmp + "/x.toad", "--quiet", "--batch", "4096"]
for _ in range(5):
    detector_cli(Detector, argv=argv)
for mb in ("32", "16", "8", "4", "32", "16", "8"):
    os.environ["THRIFTY_B200_CARD_CHUNK_MB"] = mb
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        detector_cli(Detector, argv=argv)
        ts.append(time.perf_counter() - t0)
    print("chunk %2s MiB: min %.1f ms  median %.1f ms  max %.1f ms  -> %.0f blocks/s (median)"
          % (mb, min(ts) * 1e3, float(np.median(ts)) * 1e3, max(ts) * 1e3, nb / float(np.median(ts))), flush=True)
