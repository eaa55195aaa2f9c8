import sys, time, os, io, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np
from thrifty_b200 import block_data, synth
from thrifty_b200.detect import Detector, detector_cli
tpl = np.load("tests/golden/template_example.npy")
raw, _ = synth.make_blocks(256, 16384, 4920, tpl, 1.0, seed=1)
tmp = "/dev/shm/clip"; os.makedirs(tmp, exist_ok=True)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
with open(tmp + "/x.card", "w") as f:
    block_data.write_card(f, raw[np.arange(nb) % 256])
np.save(tmp + "/t.npy", tpl)
open(tmp + "/d.cfg", "w").write("block_size: 16384\nblock_history: 4920\ncarrier_window: 7 - 110\ncarrier_threshold: 15*snr\ncorr_threshold: 15*snr\ntemplate: %s/t.npy\n" % tmp)
argv = [tmp + "/x.card", "-c", tmp + "/d.cfg", "-o", tmp + "/x.toad", "--quiet", "--batch", "4096"]
detector_cli(Detector, argv=argv)
t0 = time.perf_counter(); detector_cli(Detector, argv=argv); print("wall", time.perf_counter() - t0, "blocks", nb)
pr = cProfile.Profile(); pr.enable(); detector_cli(Detector, argv=argv); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
