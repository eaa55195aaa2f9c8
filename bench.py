#!/usr/bin/env python
"""Benchmark of the fused detect path (BASELINE.json metric: detect Msamples/s, block_len=16384).

    python bench.py --gpus 1 --steps 256 --warmup 8            # our arm (CUDA, via the C ABI)
    python bench.py --impl reference --steps K --warmup W     # reference algorithm on host cores
    torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling

A "step" is one pass of the detect hot path over one batch (4096 blocks of 16384 complex
samples) of synthetic .card payloads.  `value` = blocks/s x block_len / 1e6 with the inputs
resident in HBM; `e2e` = the same through the product API with pinned HOST buffers (H2D of raw
blocks + D2H of records inside the timed region): thr_detect_batch() on one GPU, and with N GPUs
the multi-GPU handle (thr_group_detect_batch: rank 0 drives all N GPUs, one host thread and one
contiguous stripe per GPU).  `sustained` repeats `value` over >= 1.2 s of launches; `cli_e2e` is
the wall clock of the `detect` command line on a synthetic .card; with N GPUs `stripe_parity`
says that a probe striped over the ranks and all-gathered over NCCL equals one GPU's records.

The oracle (oracle/thrifty_oracle.py, NumPy restatement of the reference) is executed here ONLY
for the `cpu_baseline` leg and the `--impl reference` arm.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from thrifty_b200 import synth  # noqa: E402

BLOCK_LEN = 16384
HISTORY = 4920
WINDOW = (7, 110)
THRESH = (0.0, 15.0, 0.0)
TEMPLATE_PATH = os.path.join(ROOT, "tests", "golden", "template_example.npy")
GATHER_EVERY = 8     # steps per all-gather of the record ring (multi-GPU)


def workload_name(args):
    return ("detect: synthetic .card payloads, block_len=%d, history=%d, example template L=4914, "
            "window 7-110, thresholds 15*snr, batch=%d, %d%% burst blocks"
            % (args.block_len, HISTORY, args.batch, round(100 * args.p_signal)))


def bench_config(args):
    """`config` of the JSON line: the same keys and values on both arms (ours and --impl reference)."""
    return {"workload": workload_name(args), "block_len": args.block_len, "history_len": HISTORY, "batch": args.batch,
            "p_signal": args.p_signal, "template": "example (L=4914)", "window": list(WINDOW), "thresholds": "15*snr / 15*snr"}


def port_vs_reference():
    """profiles/port_vs_reference.json (written in the build container by oracle/port_vs_reference.py): how the NumPy
    restatement timed here compares with the reference's own Detector on the same blocks."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "port_vs_reference.json")))
    except Exception:
        return None


def make_unique_blocks(n_unique, p_signal, seed):
    tpl = np.load(TEMPLATE_PATH)
    raw, _ = synth.make_blocks(n_unique, BLOCK_LEN, HISTORY, tpl, p_signal, seed=seed)
    return tpl, raw


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:   # timed region shorter than one sample: use everything we have
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    smmax.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smmax)) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU legs
_W = {}


def _worker_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    from oracle import thrifty_oracle as orc
    tpl = np.load(TEMPLATE_PATH)
    st = orc.DetectorSettings(block_len=BLOCK_LEN, history_len=HISTORY, carrier_len=len(tpl),
                              carrier_thresh=THRESH, carrier_window=WINDOW, template=tpl,
                              corr_thresh=THRESH)
    _W["det"] = orc.Detector(st, rxid=0)


def _worker_run(raw_chunk):
    det = _W["det"]
    n = 0
    for i in range(len(raw_chunk)):
        res = det.detect_raw(0.0, i, raw_chunk[i])
        n += bool(res.detected)
    return n


def cpu_oracle_rate_single(raw, max_seconds=20.0):
    """Single-thread oracle port over a bounded sample.  Returns (blocks/s, blocks done)."""
    _worker_init()
    det = _W["det"]
    det.detect_raw(0.0, 0, raw[0])   # warm-up (FFT plan caches, imports)
    t0 = time.perf_counter()
    done = 0
    i = 0
    while time.perf_counter() - t0 < max_seconds:      # bounded by time, cycling the sample blocks
        det.detect_raw(0.0, i, raw[i % len(raw)])
        done += 1
        i += 1
    dt = time.perf_counter() - t0
    return done / dt, done


def run_reference_arm(args):
    """--impl reference: the reference algorithm (oracle port of the NumPy path) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    _, uniq = make_unique_blocks(64, args.p_signal, synth.SEED0)
    # calibrate on one core, then size the per-step sample so the whole run takes ~90 s
    rate1, _ = cpu_oracle_rate_single(uniq[:16], max_seconds=5.0)
    total_steps = args.steps + args.warmup
    budget_s = 90.0
    sample = int(budget_s * rate1 * cores * 0.8 / total_steps)
    sample = max(cores, min(args.batch, sample))
    sample -= sample % cores
    sample = max(sample, cores)
    blocks = uniq[np.arange(sample) % len(uniq)]
    chunks = np.array_split(blocks, cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init) as pool:
        for _ in range(args.warmup):
            pool.map(_worker_run, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_run, chunks)
        dt = time.perf_counter() - t0
    ms_per_step = dt / args.steps * 1e3
    value = sample * BLOCK_LEN / (ms_per_step / 1e3) / 1e6
    pvr = port_vs_reference()
    line = {
        "impl": "reference",
        "metric": "detect Msamples/s (block_len=16384)", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (numpy)",
        "data": "synthetic",
        "config": bench_config(args),
        "details": {"sample_blocks_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "sample": "%d blocks per step x %d steps of the same synthetic workload, "
                                   "NumPy/SciPy restatement of thrifty.detect.Detector.detect "
                                   "(numpy %s pocketfft), %d worker processes; port vs the reference's own Detector in the "
                                   "build container: %s"
                                   % (sample, args.steps, np.__version__, cores,
                                      ("%.2f ms vs %.2f ms per block (port %.0f %% %s)" % (
                                          pvr["port_ms_per_block"], pvr["reference_ms_per_block"],
                                          abs(pvr["reference_ms_per_block"] / pvr["port_ms_per_block"] - 1) * 100,
                                          "faster" if pvr["port_ms_per_block"] < pvr["reference_ms_per_block"] else "slower"))
                                      if pvr else "not recorded"),
                         "port_vs_reference": pvr},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU hosts: run this rank (and so allocate its page-locked buffers, first touch) on the CPUs of the NUMA node
    its GPU hangs off, so that the e2e leg's H2D traffic does not cross the socket interconnect.  Best effort: returns the
    node, or None when the host exposes a single node / no locality (then nothing is changed)."""
    try:
        import glob
        import torch
        nodes = glob.glob("/sys/devices/system/node/node[0-9]*")
        if len(nodes) < 2:
            return None
        pr = torch.cuda.get_device_properties(local_rank)
        dev = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(dev).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


PROBE_BLOCKS = 4096        # fixed probe of the stripe-parity check (multi-GPU)


def load_profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def h2d_ceiling(n_gpus):
    """profiles/r02_h2d_concurrent.jsonl (tools/microbench/h2d_concurrent.cu on the scaling node): best aggregate pinned
    host->device bandwidth measured with n_gpus GPUs copying at once, or None."""
    best = None
    try:
        for line in open(os.path.join(ROOT, "profiles", "r02_h2d_concurrent.jsonl")):
            line = line.strip()
            if line.startswith("{"):
                row = json.loads(line)
                if row.get("gpus") == n_gpus:
                    best = max(best or 0.0, float(row["aggregate_h2d_gbs"]))
    except Exception:
        return None
    return best


def cli_e2e(tpl, raw_unique, n_blocks, device):
    """Wall-clock blocks/s of the command line (`python -m thrifty_b200 detect file.card -o out.toad --quiet`) on a
    synthetic .card of n_blocks lines; interpreter start-up and imports excluded (detector_cli is called in-process),
    everything else -- opening the file, reading text, GPU base64 decode + detect, .toad text -- included."""
    import tempfile
    from thrifty_b200 import block_data
    from thrifty_b200.detect import Detector, detector_cli
    tmp = tempfile.mkdtemp(prefix="thrifty_b200_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        card, toad, cfg, tplf = (os.path.join(tmp, f) for f in ("x.card", "x.toad", "d.cfg", "t.npy"))
        np.save(tplf, tpl)
        with open(card, "w") as f:
            block_data.write_card(f, raw_unique[np.arange(n_blocks) % len(raw_unique)])
        with open(cfg, "w") as f:
            f.write("block_size: %d\nblock_history: %d\ncarrier_window: %d - %d\ncarrier_threshold: 15*snr\n"
                    "corr_threshold: 15*snr\ntemplate: %s\n" % (BLOCK_LEN, HISTORY, WINDOW[0], WINDOW[1], tplf))
        argv = [card, "-c", cfg, "-o", toad, "--quiet", "--batch", "4096", "--device", str(device)]
        def timed(av, reps=5, warm=2):
            # wall clock of the whole command, median of `reps` runs after `warm` untimed ones (the first runs on a fresh
            # box also pay for the CUDA context, the page cache of the file and first-touch of the staging buffers:
            # 0.74 / 0.12 / 0.09 s before it settles at ~0.075 s)
            for _ in range(warm):
                detector_cli(Detector, argv=av)
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                detector_cli(Detector, argv=av)
                ts.append(time.perf_counter() - t0)
            return float(np.median(ts)), ts

        dt, dt_all = timed(argv)
        n_lines = sum(1 for _ in open(toad))
        # the same on the first quarter of the file: the difference quotient is the rate once the fixed costs of a run
        # (handle creation, pinning two 16 MiB buffers, opening files: ~50-100 ms) are paid
        quarter = os.path.join(tmp, "q.card")
        with open(card, "rb") as src, open(quarter, "wb") as dst:
            lines_q = 0
            for line in src:
                dst.write(line)
                lines_q += not line.startswith(b"#")
                if lines_q >= n_blocks // 4:
                    break
        argv_q = [quarter] + argv[1:]
        dt_q, dt_q_all = timed(argv_q, warm=1)
        # host-side noise is one-sided (stalls of 0.1-0.5 s on shared hosts): the difference quotient uses the best runs
        best, best_q = min(dt_all), min(dt_q_all)
        steady = (n_blocks - lines_q) / (best - best_q) if best > best_q else None
        return {"value": n_blocks / dt, "unit": "blocks/s", "msamples_per_s": n_blocks * BLOCK_LEN / dt / 1e6,
                "blocks": n_blocks, "seconds": dt, "card_bytes": os.path.getsize(card), "toad_lines": n_lines,
                "seconds_all_runs": dt_all, "quarter_file_seconds": dt_q, "quarter_file_seconds_all_runs": dt_q_all,
                "steady_blocks_per_s": steady,
                "fixed_cost_seconds": (best_q - lines_q / steady) if steady else None,
                "best_seconds": best, "best_blocks_per_s": n_blocks / best,
                "what": "thrifty_b200.detect.detector_cli(Detector) in-process on a synthetic .card in /dev/shm: file read + "
                        "GPU base64 decode + detect + .toad text; interpreter start-up, imports, CUDA context and the two page-locked staging buffers (pinned once per process) "
                        "excluded; seconds = median of 5 runs after 2 warm-up runs; steady_blocks_per_s = extra blocks / extra seconds between the best quarter-file run and the best whole-file run"}
    except Exception as e:      # noqa: BLE001  (the bench line must still come out)
        return {"error": "%s: %s" % (type(e).__name__, e)}
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from thrifty_b200._native import NativeDetector, NativeGroup, PinnedBuffer, RECORD_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the detect path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL / c10d print their version banner on stdout when the communicator is created:
        # send it to stderr so that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
            cpu_group = dist.new_group(backend="gloo")      # host-side barriers (no kernel spinning on the GPUs)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus

    n, batch = args.block_len, args.batch
    tpl, uniq = make_unique_blocks(args.unique, args.p_signal, synth.SEED0 + 100003 * rank)
    det = NativeDetector(n, HISTORY, tpl, len(tpl), WINDOW, THRESH, THRESH, device=local_rank,
                         max_batch=batch, overlap_launches=True)
    info = det.info()

    # device-resident pool of raw blocks, larger than L2 (126 MB): pool_blocks * 32 KiB
    pool_blocks = args.pool
    assert pool_blocks % batch == 0
    uniq_d = torch.from_numpy(uniq).to(dev)
    reps = (pool_blocks + len(uniq) - 1) // len(uniq)
    pool = uniq_d.repeat(reps, 1)[:pool_blocks].contiguous()
    idx = torch.arange(pool_blocks, dtype=torch.int64, device=dev) + rank * (1 << 40)
    # ring of record buffers: consecutive launches may overlap at their edges (PDL), so each step
    # writes its own slot; with several ranks the ring is all-gathered once per GATHER_EVERY steps
    # (one 2 MiB collective instead of eight 256 KiB ones: NCCL launch latency, not bandwidth, is
    # what a 64-byte-per-block gather costs)
    ring = torch.zeros(GATHER_EVERY * batch * 64, dtype=torch.uint8, device=dev)
    recs2 = [ring[k * batch * 64:(k + 1) * batch * 64] for k in range(GATHER_EVERY)]
    gathered = torch.zeros(world * GATHER_EVERY * batch * 64, dtype=torch.uint8, device=dev) if world > 1 else None
    # a real (non-legacy) stream: handle 0 would mean "the detector's own stream" to thr_set_stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    det.set_stream(stream.cuda_stream)
    n_windows = pool_blocks // batch

    def step(i, gather=True, last=False):
        w = i % n_windows
        rec = recs2[i % GATHER_EVERY]
        det.detect_device(pool[w * batch].data_ptr(), idx[w * batch].data_ptr(), batch, rec.data_ptr())
        if gather and world > 1 and (i % GATHER_EVERY == GATHER_EVERY - 1 or last):
            dist.all_gather_into_tensor(gathered, ring)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def timed(k, gather=True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(k):
            step(i, gather, last=(i == k - 1))
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = det.info()["launches"]
    t_wall0 = time.time()
    ms_total = timed(args.steps, gather=True)
    t_wall1 = time.time()
    launches = det.info()["launches"] - launches0
    # the timed region may be shorter than the sampler period: keep the GPU busy a little longer
    # (untimed) so that at least a few clock samples are taken under the same load
    t_extra0 = time.time()
    while time.time() - t_extra0 < 1.0:
        for i in range(8):
            step(i, gather=False)
        torch.cuda.synchronize()
    clocks = sampler.stop(t_wall0, time.time())
    ms_per_step = ms_total / args.steps
    value = world * batch * n / (ms_per_step * 1e-3) / 1e6

    # kernel-only timing for the roofline (same launches, no gather)
    ms_kernel = timed(args.steps, gather=False) / args.steps if world > 1 else ms_per_step

    # ---- sustained: at least args.sustained_seconds of back-to-back launches with its own clock record
    sustained = None
    if args.sustained_seconds > 0:
        k_sus = max(args.steps, int(args.sustained_seconds * 1e3 / ms_per_step) + 1)
        s2 = ClockSampler(local_rank)
        s2.start()
        time.sleep(0.25)
        t0w = time.time()
        ms_sus = timed(k_sus, gather=True)
        t1w = time.time()
        clk2 = s2.stop(t0w, t1w)
        sustained = {"value": world * batch * n * k_sus / (ms_sus * 1e-3) / 1e6, "unit": "Msamples/s", "steps": k_sus,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / k_sus, "clocks": clk2}

    # records sanity: every block of the last batch must carry a decision
    torch.cuda.synchronize()
    recs = np.frombuffer(recs2[(args.steps - 1) % GATHER_EVERY].cpu().numpy().tobytes(), dtype=RECORD_DTYPE)
    n_det = int(((recs["flags"] & 2) != 0).sum())
    n_car = int(((recs["flags"] & 1) != 0).sum())

    # ---- stripe parity (multi-GPU): one fixed probe striped over the ranks, records all-gathered over NCCL, compared
    # byte for byte with rank 0's single-GPU records of the whole probe
    stripe_parity = None
    identify_leg = None
    if world > 1:
        _, probe_u = make_unique_blocks(256, 0.7, synth.SEED0 + 424242)           # same probe on every rank
        probe = probe_u[np.arange(PROBE_BLOCKS) % len(probe_u)]
        from thrifty_b200 import stripe                 # stripe bounds + record gather: the code tests/test_stripe_gloo.py
        lo, hi = stripe.stripe_bounds(PROBE_BLOCKS, world, rank)                    # covers with gloo on CPU ranks
        part = torch.from_numpy(np.ascontiguousarray(probe[lo:hi])).to(dev)
        pidx = torch.arange(lo, hi, dtype=torch.int64, device=dev)
        mine = torch.zeros(max(hi - lo, 1) * 64, dtype=torch.uint8, device=dev)
        if hi > lo:
            det.detect_device(part.data_ptr(), pidx.data_ptr(), hi - lo, mine.data_ptr())
        torch.cuda.synchronize()
        local = np.frombuffer(mine.cpu().numpy().tobytes(), dtype=RECORD_DTYPE)[:hi - lo].reshape(hi - lo, 1)
        got = stripe.gather_records(local, PROBE_BLOCKS)[:, 0]                      # NCCL all-gather of the 64-byte records
        if rank == 0:
            whole = det.detect_raw(probe, np.arange(PROBE_BLOCKS))[:, 0]
            stripe_parity = bool(got.tobytes() == whole.tobytes())
            # the consumer of the gathered records: `identify` on the device (txid by carrier-bin window, duplicate
            # filter over adjacent blocks) must give the same survivors from the gathered records as from one GPU's
            from thrifty_b200 import identify as dev_identify
            ts_probe = 1000.0 + 0.0047767 * np.arange(PROBE_BLOCKS)
            sel_g, tx_g = dev_identify.integrate_records(got, ts_probe, 0, None, device=local_rank)
            sel_1, tx_1 = dev_identify.integrate_records(whole, ts_probe, 0, None, device=local_rank)
            identify_leg = {"detections": int(((got["flags"] & 2) != 0).sum()), "kept": int(len(sel_g)),
                            "transmitters": int(len(set(tx_g.tolist()))),
                            "equal_single_gpu": bool(np.array_equal(sel_g, sel_1) and np.array_equal(tx_g, tx_1))}
        del part, pidx, mine

    # ---- e2e: HOST buffers through the product API, copies inside the timed region.
    # 1 GPU: thr_detect_batch with pinned buffers.  N GPUs: rank 0 drives all N through the multi-GPU handle
    # (thr_group_detect_batch: one host thread + stripe per GPU, NUMA-placed pinned input), the other ranks idle at a
    # host-side barrier -- this is the product's multi-GPU path, not N independent benchmarks.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_extra = {}
    host_barrier()
    if world == 1:
        hbuf = [PinnedBuffer(batch * 2 * n) for _ in range(2)]
        hidx = [PinnedBuffer(batch * 8) for _ in range(2)]
        hout = [PinnedBuffer(batch * 64) for _ in range(2)]
        lib = det._lib
        for b in range(2):
            src = uniq[(np.arange(batch) + b * 7) % len(uniq)]
            hbuf[b].array[:] = src.reshape(-1)
            hidx[b].array.view(np.int64)[:] = np.arange(batch) + b * batch
        for b in range(2):   # warm-up
            det._check(lib.thr_detect_batch(det.handle, hbuf[b].ptr, hidx[b].ptr, batch, hout[b].ptr))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            b = i & 1
            det._check(lib.thr_detect_batch(det.handle, hbuf[b].ptr, hidx[b].ptr, batch, hout[b].ptr))
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_recs = hout[(e2e_steps - 1) & 1].array.view(RECORD_DTYPE)
        e2e_det = int(((e2e_recs["flags"] & 2) != 0).sum())
        e2e_api = "thr_detect_batch (pinned host buffers)"
        for b in hbuf + hidx + hout:
            b.close()
        # the same number of blocks as ONE contiguous raw stream (the reference's block_reader input, block_data.py:70-98):
        # the windows overlap by `history`, so only N - H new samples per block cross PCIe (thr_detect_stream)
        new = 2 * (n - HISTORY)
        sbytes = 2 * HISTORY + batch * new
        spin = [PinnedBuffer(sbytes) for _ in range(2)]
        sout = [PinnedBuffer(batch * 64) for _ in range(2)]
        for b in range(2):
            src = uniq[(np.arange(batch) + b * 7) % len(uniq)]
            spin[b].array[:2 * HISTORY] = src[0][:2 * HISTORY]
            spin[b].array[2 * HISTORY:] = src[:, 2 * HISTORY:].reshape(-1)
        got = ctypes.c_int64(0)
        for b in range(2):   # warm-up
            det._check(lib.thr_detect_stream(det.handle, spin[b].ptr, sbytes, b * batch, sout[b].ptr, ctypes.byref(got)))
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            b = i & 1
            det._check(lib.thr_detect_stream(det.handle, spin[b].ptr, sbytes, b * batch, sout[b].ptr, ctypes.byref(got)))
        torch.cuda.synchronize()
        s_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        s_recs = sout[(e2e_steps - 1) & 1].array.view(RECORD_DTYPE)
        e2e_extra["raw_stream"] = {
            "value": int(got.value) * n / (s_ms * 1e-3) / 1e6, "unit": "Msamples/s", "api": "thr_detect_stream (pinned host stream)",
            "blocks_per_step": int(got.value), "ms_per_step": s_ms, "h2d_bytes_per_step": sbytes,
            "h2d_gbs": sbytes / (s_ms * 1e-3) / 1e9, "carrier_detected_last_batch": int(((s_recs["flags"] & 1) != 0).sum()),
            "what": "the same blocks as one contiguous uint8 I/Q stream whose windows overlap by `history` (what "
                    "`detect --raw` reads): N - H new samples per block cross PCIe instead of N"}
        for b in spin + sout:
            b.close()
    elif rank == 0:
        grp = NativeGroup(list(range(world)), n, HISTORY, tpl, len(tpl), WINDOW, THRESH, THRESH, max_batch=batch)
        lib = grp._lib
        total = world * batch
        bufs = []
        for b in range(2):
            ptr = lib.thr_group_host_alloc(grp._g, batch * 2 * n)
            assert ptr, "thr_group_host_alloc failed"
            arr = np.ctypeslib.as_array((ctypes_u8() * (total * 2 * n)).from_address(ptr))
            arr[:] = uniq[(np.arange(total) + b * 7) % len(uniq)].reshape(-1)
            bufs.append((ptr, arr))
        hidx = [PinnedBuffer(total * 8) for _ in range(2)]
        hout = [PinnedBuffer(total * 64) for _ in range(2)]
        for b in range(2):
            hidx[b].array.view(np.int64)[:] = np.arange(total) + b * total
        for b in range(2):   # warm-up
            grp.detect_raw_ptr(bufs[b][0], total, hidx[b].ptr, hout[b].ptr)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            b = i & 1
            grp.detect_raw_ptr(bufs[b][0], total, hidx[b].ptr, hout[b].ptr)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_recs = hout[(e2e_steps - 1) & 1].array.view(RECORD_DTYPE)
        e2e_det = int(((e2e_recs["flags"] & 2) != 0).sum())
        # the striped records must be what one GPU returns for the same blocks
        chk = det.detect_raw(bufs[(e2e_steps - 1) & 1][1].reshape(total, 2 * n)[:batch],
                             hidx[(e2e_steps - 1) & 1].array.view(np.int64)[:batch])[:, 0]
        e2e_extra["group_records_equal_single_gpu"] = bool(chk.tobytes() == e2e_recs[:batch].tobytes())
        e2e_extra["numa_nodes"] = grp.numa_nodes()
        e2e_api = "thr_group_detect_batch (rank 0 drives %d GPUs: one host thread + stripe per GPU, pinned host buffers)" % world
        for ptr, arr in bufs:
            del arr
            lib.thr_group_host_free(grp._g, ptr, batch * 2 * n)
        for b in hidx + hout:
            b.close()
        grp.close()
    host_barrier()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback from B200_PROFILING.md"
        alg_bytes = batch * (2 * n + 64)                   # raw u8 in + 64-B record out, per launch
        achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
        traffic = (load_profile_json("ncu_traffic.json") or {}).get("dram_bytes_per_launch")
        # the binding resource is the FP32 pipe, not HBM.  Two flop counts per launch:
        #   algorithmic = 3 FFTs x 5 N log2 N + 40 N pointwise per block (the SURVEY's convention; FFT#1 counted in full
        #                 although this configuration prunes it)
        #   executed    = what the kernel's instruction stream really performs (ncu opcode mix of the same command:
        #                 32 lanes x (4 per FFMA2, 2 per FADD2/FMUL2/FFMA/DFMA, 1 per FADD/FMUL), profiles/r02_opcode_mix.json)
        # against 148 SMs x 128 lanes x 2 flop x the SM clock sampled during the timed region
        alg_flops = batch * (3 * 5 * n * int(np.log2(n)) + 40 * n)
        fp32_peak = 148 * 128 * 2 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12
        fp32_ach = alg_flops / (ms_kernel * 1e-3) / 1e12
        mix = load_profile_json("r02_opcode_mix.json")
        exe_flops = batch * mix["executed_flops_per_block"] if mix else None
        fma_pipe = None
        try:
            m = load_profile_json("r02_final_ncu_summary.json")["metrics"]
            fma_pipe = float(m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]["value"]) / 100.0
        except Exception:
            pass
        e2e_bytes = world * batch * (2 * n + 8)
        ceiling = h2d_ceiling(world)
        e2e_gbs = e2e_bytes / (e2e_ms * 1e-3) / 1e9
        line = {
            "metric": "detect Msamples/s (block_len=16384)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "precision_note": "float32 transforms and decisions throughout (the reference: float32 FFT#1, complex128 from the "
                              "mix on), float64 Dirichlet fit (MINPACK lmdif, as the reference); within the 1e-4 parity bar",
            "data": "synthetic",
            "config": bench_config(args),
            "details": {"pool_blocks_per_gpu": pool_blocks,
                        "l2_policy": "inputs larger than L2: %d MiB raw pool cycled per GPU" % (pool_blocks * 2 * n >> 20),
                        "kernel": info["kernel"], "grid": info["grid"], "threads": info["threads"],
                        "smem_bytes": info["smem_bytes"], "parallelism": "stripe%d" % world, "numa_node_rank0": numa_node,
                        "record_gather": ("all_gather of the 64-B record ring every %d steps" % GATHER_EVERY) if world > 1 else "none (1 GPU)",
                        "carrier_detected_last_batch": n_car, "corr_detected_last_batch": n_det,
                        "blocks_per_s": world * batch / (ms_per_step * 1e-3)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_per_launch": ms_kernel,
                         "fp32": {"achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak,
                                  "algorithmic_flops_per_launch": alg_flops,
                                  "executed_flops_per_launch": exe_flops,
                                  "executed_achieved": (exe_flops / (ms_kernel * 1e-3) / 1e12) if exe_flops else None,
                                  "executed_frac": (exe_flops / (ms_kernel * 1e-3) / 1e12 / fp32_peak) if exe_flops else None,
                                  "executed_source": "profiles/r02_opcode_mix.json (tools/ncu_opcodes.py on the ncu capture of this command)",
                                  "fma_pipe_cycles_active_ncu": fma_pipe},
                         "note": "fused kernel is FP32-issue/shared-memory bound (~125 FLOP/B), see DESIGN.md"},
            "e2e": dict({"value": world * batch * n / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                         "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": world * batch * 64,
                         "ms_per_step": e2e_ms, "steps": e2e_steps, "api": e2e_api, "corr_detected_last_batch": e2e_det,
                         "h2d_gbs": e2e_gbs, "h2d_ceiling_gbs": ceiling,
                         "frac_of_h2d_ceiling": (e2e_gbs / ceiling) if ceiling else None,
                         "h2d_ceiling_source": "profiles/r02_h2d_concurrent.jsonl (tools/microbench/h2d_concurrent.cu, "
                                               "%d GPUs copying at once)" % world}, **e2e_extra),
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if sustained:
            line["sustained"] = sustained
        if stripe_parity is not None:
            line["stripe_parity"] = stripe_parity
            line["identify_on_gathered_records"] = identify_leg
        if args.cli and world == 1:
            line["cli_e2e"] = cli_e2e(tpl, uniq, args.cli_blocks, local_rank)
        if args.cpu_baseline and world == 1:
            rate, done = cpu_oracle_rate_single(uniq[np.arange(4096) % len(uniq)], max_seconds=args.cpu_seconds)
            pvr = port_vs_reference()
            line["cpu_baseline"] = {
                "value": rate * n / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                "sample": "%d blocks (~%.0f s) of the same workload, single process, NumPy/SciPy restatement of "
                          "thrifty.detect.Detector.detect (numpy %s pocketfft, scipy curve_fit); the reference's own Detector "
                          "takes %s x the port's time per block (profiles/port_vs_reference.json)"
                          % (done, args.cpu_seconds, np.__version__, ("%.2f" % pvr["reference_over_port"]) if pvr else "?"),
                "blocks_per_s": rate, "host_cores_available": os.cpu_count(), "port_vs_reference": pvr}
        print(json.dumps(line))
    det.close()
    if world > 1:
        dist.destroy_process_group()


def ctypes_u8():
    import ctypes
    return ctypes.c_uint8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--block-len", type=int, default=BLOCK_LEN)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--pool", type=int, default=16384, help="device-resident raw blocks per GPU")
    ap.add_argument("--unique", type=int, default=512, help="distinct synthetic blocks generated on the host")
    ap.add_argument("--p-signal", type=float, default=1.0,
                    help="fraction of blocks carrying a burst (a .card holds carrier-positive blocks: 1.0)")
    ap.add_argument("--e2e-steps", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--sustained-seconds", type=float, default=1.2,
                    help="length of the extra back-to-back run reported as `sustained` (0: skip)")
    ap.add_argument("--no-cli", dest="cli", action="store_false", help="skip the command-line wall-clock leg (cli_e2e)")
    ap.add_argument("--cli-blocks", type=int, default=16384,
                    help="lines of the synthetic .card of the cli_e2e leg (44 KB of text each)")
    args = ap.parse_args()
    if args.block_len != BLOCK_LEN:
        raise SystemExit("bench.py measures the headline config (block_len=16384); use tools/sweep.py for others")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
