"""Phase times of 40 consecutive `detect x.card -o x.toad --quiet` runs (THRIFTY_B200_CLI_TIMING): which phase do the
occasional 0.1-0.5 s stalls sit in?   python tools/cli_stalls.py [lines]"""
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
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
with open(tmp + "/x.card", "w") as f:
    block_data.write_card(f, raw[np.arange(nb) % 256])
np.save(tmp + "/t.npy", tpl)
open(tmp + "/d.cfg", "w").write("block_size: 16384\nblock_history: 4920\ncarrier_window: 7 - 110\ncarrier_threshold: 15*snr\n"
                                "corr_threshold: 15*snr\ntemplate: %s/t.npy\n" % tmp)
argv = [tmp + "/x.card", "-c", tmp + "/d.cfg", "-o", tmp + "/x.toad", "--quiet", "--batch", "4096"]
os.environ["THRIFTY_B200_CLI_TIMING"] = "1"
for i in range(40):
    t0 = time.perf_counter()
    detector_cli(Detector, argv=argv)
    print("run %2d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
