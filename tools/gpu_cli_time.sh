#!/bin/bash
# wall-clock of the `detect` / `fastdet` command lines on a synthetic 4096-block .card (N=16384)
mkdir -p gpurun_out /tmp/cli
export PYTHONPATH=$PWD${PYTHONPATH:+:$PYTHONPATH}
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from thrifty_b200 import synth, block_data, fastdet
tpl = np.load('tests/golden/template_example.npy')
raw, _ = synth.make_blocks(256, 16384, 4920, tpl, 0.9, seed=11)
raw = raw[np.arange(4096) % 256]
with open('/tmp/cli/in.card', 'w') as f:
    block_data.write_card(f, raw)
np.save('/tmp/cli/template.npy', tpl)
fastdet.save_template('/tmp/cli/template.tpl', tpl)
open('/tmp/cli/detector.cfg', 'w').write("sample_rate: 2.4M\nblock_size: 16384\nblock_history: 4920\ncarrier_window: 7 - 110\ncarrier_threshold: 15 * snr\ncorr_threshold: 15 * snr\ntemplate: /tmp/cli/template.npy\nrxid: 0\n")
PY
cd /tmp/cli
python - <<'PY'
import subprocess, sys, time
def run(name, cmd):
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.time() - t0
    out = cmd[cmd.index("-o") + 1]
    n = sum(1 for _ in open(out)) if r.returncode == 0 else -1
    print("%-22s %6.2f s wall, %5d toad lines, rc %d %s" % (name, dt, n, r.returncode, r.stderr[-300:] if r.returncode else ""), flush=True)
py = [sys.executable, "-m", "thrifty_b200"]
run("detect --quiet", py + ["detect", "in.card", "-o", "out.toad", "--quiet"])
run("detect", py + ["detect", "in.card", "-o", "out.toad"])
run("detect --host-decode -q", py + ["detect", "in.card", "-o", "out3.toad", "--quiet", "--host-decode"])
run("fastdet -q", py + ["fastdet", "--card", "-i", "in.card", "-z", "template.tpl", "-o", "out2.toad", "-w", "7-110", "-t", "15s", "-u", "15s", "-q"])
run("fastdet", py + ["fastdet", "--card", "-i", "in.card", "-z", "template.tpl", "-o", "out2.toad", "-w", "7-110", "-t", "15s", "-u", "15s"])
t0 = time.time(); subprocess.run([sys.executable, "-c", "import thrifty_b200, numpy"]); print("python + imports alone %.2f s" % (time.time() - t0))
PY
