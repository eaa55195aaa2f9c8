"""Detect positioning signals and estimate sample-of-arrival -- B200 drop-in for thrifty.detect.

Same public surface as thrifty/detect.py:24-227:

  * ``DetectorSettings`` namedtuple (detect.py:24-31),
  * ``Detector(settings, blocks, rxid, yield_data)``: iterator yielding one
    ``(detected, DetectionResult)`` per input block, in input order, including blocks
    without a carrier (detect.py:80-91); ``detect(timestamp, block_idx, block)`` for
    single blocks (detect.py:60-78); with ``yield_data=True`` the shifted spectrum and the
    correlation are returned too (detect.py:75-76),
  * ``SummaryLineFormatter`` (detect.py:103-158) and ``detector_cli(detector_class, ...)``
    (detect.py:161-223), the plug-in seam used by the reference's experimental detectors.

All arithmetic runs in the fused CUDA kernel behind the C ABI (include/thrifty_b200.h);
this module only batches blocks (read-ahead of ``batch`` blocks per launch) and converts
the 64-byte records back into the reference's result types.  There is no CPU fallback.
"""

from __future__ import print_function

import argparse
import gc
import os
import sys
from collections import namedtuple

import numpy as np

from thrifty_b200 import toads_data, util
from thrifty_b200._native import FLAG_CARRIER, FLAG_CORR, NativeDetector, NativeGroup
from thrifty_b200.block_data import block_reader, card_reader
from thrifty_b200.setting_parsers import normalize_freq_range
from thrifty_b200.settings import load_args

DetectorSettings = namedtuple("DetectorSettings", [
    "block_len", "history_len", "carrier_len", "carrier_thresh", "carrier_window",
    "template", "corr_thresh"])


def _pread_full(fd, view, offset):
    """pread until `view` is full or the file ends; returns the number of bytes read (the GIL is released in pread)."""
    import os
    got = 0
    while got < len(view):
        n = os.preadv(fd, [view[got:]], offset + got)
        if n <= 0:
            break
        got += n
    return got


def record_to_result(rec, timestamp, rxid):
    """thr_record -> (detected, DetectionResult) with the reference's field types."""
    carrier_found = bool(rec["flags"] & FLAG_CARRIER)
    detected = bool(rec["flags"] & FLAG_CORR)
    carrier_info = toads_data.CarrierSyncInfo(
        int(rec["carrier_bin"]), float(rec["carrier_offset"]) if carrier_found else 0,
        rec["carrier_energy"], rec["carrier_noise"])
    if carrier_found:
        corr_info = toads_data.CorrDetectionInfo(
            int(rec["corr_sample"]), float(rec["corr_offset"]) if detected else 0,
            float(rec["corr_energy"]), float(rec["corr_noise"]))
        soa = float(rec["soa"])
    else:
        corr_info, soa = None, None
    result = toads_data.DetectionResult(timestamp, int(rec["block_idx"]), soa, carrier_info,
                                        corr_info, rxid)
    return detected, result


def records_to_results(recs, timestamps, rxid):
    """Batch form of record_to_result: 1-D thr_record array + timestamps -> list of (detected, DetectionResult).

    Same values and field types as record_to_result (carrier energy / noise stay numpy float32 scalars like the
    reference's, everything else Python numbers), but the columns are pulled out of the structured array once per batch:
    ~1 us instead of ~12 us per block, which matters once the GPU side ingests a million blocks a second."""
    n = len(recs)
    if n == 0:
        return []
    flags = recs["flags"].tolist()
    block_idx = recs["block_idx"].tolist()
    soa = recs["soa"].tolist()
    cbin = recs["carrier_bin"].tolist()
    coff = recs["carrier_offset"].tolist()
    cen = list(recs["carrier_energy"])
    cno = list(recs["carrier_noise"])
    ksam = recs["corr_sample"].tolist()
    koff = recs["corr_offset"].tolist()
    ken = recs["corr_energy"].tolist()
    kno = recs["corr_noise"].tolist()
    if np.ndim(timestamps) == 0:
        timestamps = [timestamps] * n
    elif isinstance(timestamps, np.ndarray):
        timestamps = timestamps.tolist()
    carrier_cls, corr_cls, result_cls = toads_data.CarrierSyncInfo, toads_data.CorrDetectionInfo, toads_data.DetectionResult
    out = []
    # 3 n container objects without reference cycles: the cyclic collector would only re-scan them (2-3x slower)
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        _fill_results(out, n, flags, timestamps, block_idx, soa, cbin, coff, cen, cno, ksam, koff, ken, kno, rxid,
                      carrier_cls, corr_cls, result_cls)
    finally:
        if gc_was_on:
            gc.enable()
    return out


def _fill_results(out, n, flags, timestamps, block_idx, soa, cbin, coff, cen, cno, ksam, koff, ken, kno, rxid,
                  carrier_cls, corr_cls, result_cls):
    for i in range(n):
        f = flags[i]
        detected = bool(f & FLAG_CORR)
        if f & FLAG_CARRIER:
            carrier_info = carrier_cls(cbin[i], coff[i], cen[i], cno[i])
            corr_info = corr_cls(ksam[i], koff[i] if detected else 0, ken[i], kno[i])
            out.append((detected, result_cls(timestamps[i], block_idx[i], soa[i], carrier_info, corr_info, rxid)))
        else:
            out.append((detected, result_cls(timestamps[i], block_idx[i], None, carrier_cls(cbin[i], 0, cen[i], cno[i]),
                                             None, rxid)))


class Detector(object):
    """All-in-one carrier detect / sync / correlate / SoA estimate on the GPU.

    Parameters follow thrifty/detect.py:40; extras: ``batch`` = blocks per kernel launch
    (read-ahead of the iterator), ``device`` = CUDA ordinal, ``devices`` = list of CUDA ordinals:
    every batch is then cut into contiguous stripes, one per GPU (one host thread each, no
    collective), and the results come back in input order exactly as from one GPU.  ``block``
    arguments may be complex arrays of length block_len (reference behaviour) or uint8 arrays of
    length 2*block_len (raw I/Q, conversion then runs on the GPU)."""

    def __init__(self, settings, blocks=None, rxid=-1, yield_data=False, batch=256, device=0, devices=None):
        self.settings = settings
        self.blocks = iter(blocks) if blocks is not None else None
        self.rxid = rxid
        self.yield_data = yield_data
        self.batch = max(1, int(batch))
        template = np.asarray(settings.template, dtype=np.float64)
        if template.ndim != 1:
            raise ValueError("Detector takes a single 1-D template; see MultiTemplateDetector")
        assert settings.history_len >= len(template) - 1     # soa_estimator.py:33
        if devices is not None and len(devices) > 1:
            self.native = NativeGroup(
                devices, settings.block_len, settings.history_len, template, settings.carrier_len,
                settings.carrier_window, settings.carrier_thresh, settings.corr_thresh,
                max_batch=self.batch)
        else:
            self.native = NativeDetector(
                settings.block_len, settings.history_len, template, settings.carrier_len,
                settings.carrier_window, settings.carrier_thresh, settings.corr_thresh,
                device=device if not devices else devices[0], max_batch=self.batch)
        self.new_len = settings.block_len - settings.history_len
        self._pending = []

    # ---- single block (detect.py:60-78)
    def detect(self, timestamp, block_idx, block):
        """Process the given block of data."""
        block = np.asarray(block)
        is_raw = block.dtype == np.uint8
        assert len(block) == (2 if is_raw else 1) * self.settings.block_len
        if self.yield_data:
            recs, sfft, corr, _ = self.native.detect_block_data(
                raw=block if is_raw else None, iq=None if is_raw else block, block_idx=block_idx)
            detected, result = record_to_result(recs[0], timestamp, self.rxid)
            if result.corr_info is None:
                return detected, result, None, None
            return detected, result, sfft, corr
        if is_raw:
            recs = self.native.detect_raw(block[None, :], [block_idx])
        else:
            recs = self.native.detect_c64(block[None, :], [block_idx])
        return record_to_result(recs[0, 0], timestamp, self.rxid)

    def __call__(self, timestamp, block_idx, block):
        return self.detect(timestamp, block_idx, block)

    # ---- batch of blocks: list of (timestamp, block_idx, block)
    def detect_many(self, items):
        """One launch for a list of (timestamp, block_idx, block); returns a list of
        (detected, DetectionResult) in input order."""
        if not items:
            return []
        n = self.settings.block_len
        idx = np.array([it[1] for it in items], dtype=np.int64)
        blocks = [np.asarray(it[2]) for it in items]
        if all(b.dtype == np.uint8 for b in blocks):
            for b in blocks:
                assert len(b) == 2 * n
            recs = self.native.detect_raw(np.stack(blocks), idx)
        else:
            from thrifty_b200.block_data import raw_to_complex
            conv = []
            for b in blocks:
                if b.dtype == np.uint8:
                    assert len(b) == 2 * n
                    b = raw_to_complex(b)
                assert len(b) == n                                  # detect.py:62
                conv.append(np.asarray(b, dtype=np.complex64))
            recs = self.native.detect_c64(np.stack(conv), idx)
        return records_to_results(recs[:, 0], [it[0] for it in items], self.rxid)

    # ---- whole `.card` streams: scan on the host, base64 decode + detect on the GPU
    def iter_card_records(self, stream, chunk_bytes=None, min_lines=None, read_threads=8):
        """Yield (timestamps f8[B], block_idx i8[B], records [B, T]) per chunk of a binary `.card` stream.

        The text goes to the GPU as is (thr_detect_card: host jump-scan of the line headers, base64 decode and detect on
        the device).  A reader thread fills one of two page-locked buffers while the GPU works on the other; a regular
        file is read with `read_threads` parallel preads (one thread copies ~6 GB/s out of the page cache, eight ~21 on the
        test hosts, the GPU path takes ~45; mapping the file instead costs a page fault per 4 KB and is 2x slower), a pipe
        with whatever each read returns -- accumulated until `min_lines` lines' worth of text
        (default min(batch, 8)) or the end of the stream is buffered, so a pipe that hands out 64 KB at a time does not
        cost one launch per read; lower it (or --batch) for latency on live inputs.  Chunks are cut at line ends."""
        import os
        import queue
        import stat
        import threading
        from concurrent.futures import ThreadPoolExecutor
        from thrifty_b200._native import acquire_staging, release_staging
        line_len = ((2 * self.settings.block_len + 2) // 3) * 4 + 64
        if chunk_bytes is None:                  # size of each staging buffer; THRIFTY_B200_CARD_CHUNK_MB overrides
            chunk_bytes = int(float(os.environ.get("THRIFTY_B200_CARD_CHUNK_MB", "16")) * (1 << 20))
        chunk_bytes = max(int(chunk_bytes), 4 * line_len)
        if min_lines is None:
            min_lines = max(1, min(self.batch, 8))
        want = min(chunk_bytes, min_lines * line_len)
        fd, regular = None, False
        try:
            fd = stream.fileno()
            regular = stat.S_ISREG(os.fstat(fd).st_mode)
        except (AttributeError, OSError, ValueError):
            pass
        read_into = getattr(stream, "readinto1", None) or getattr(stream, "readinto", None)
        # two page-locked buffers (one spare byte keeps the C number parser inside the allocation), taken from the
        # per-process cache or allocated by the reader thread when it first needs them: pinning costs 0.6-1 ms per MiB (the fixed cost of a short run), and while the
        # reader bounds the rate 16 MiB chunks (~380 lines of 16384 samples) keep the GPU side above it
        bufs = []
        free_q, ready_q = queue.Queue(), queue.Queue()
        stop = threading.Event()

        def next_buffer():
            try:
                return free_q.get_nowait()
            except queue.Empty:
                if len(bufs) < 2:
                    bufs.append(acquire_staging(chunk_bytes + 1))
                    return bufs[-1]
                return free_q.get()

        def fill_regular(view, start, pool, pos):
            """Parallel pread of [pos, pos + room) into view[start:]; returns the number of bytes read."""
            room = chunk_bytes - start
            part = max(1 << 20, -(-room // read_threads))
            jobs = []
            for off in range(0, room, part):
                n = min(part, room - off)
                jobs.append(pool.submit(_pread_full, fd, memoryview(view)[start + off:start + off + n], pos + off))
            got = 0
            for n_req, fut in zip([min(part, room - off) for off in range(0, room, part)], jobs):
                n = fut.result()
                got += n
                if n < n_req:
                    break
            return got

        def reader():
            carry = b""
            pos = 0
            pool = ThreadPoolExecutor(read_threads) if regular else None
            try:
                if regular:
                    pos = stream.tell()
                eof = False
                while not eof and not stop.is_set():
                    buf = next_buffer()
                    if buf is None:
                        break
                    view = buf.array
                    fill = len(carry)
                    if fill:
                        view[:fill] = np.frombuffer(carry, dtype=np.uint8)
                    if regular:
                        got = fill_regular(view, fill, pool, pos)
                        pos += got
                        eof = got < chunk_bytes - fill
                        fill += got
                    else:
                        target = max(want, fill + 1)
                        while not eof and fill < target:
                            if read_into is not None:
                                got = read_into(memoryview(view)[fill:chunk_bytes]) or 0
                            else:
                                data = stream.read(chunk_bytes - fill)
                                if isinstance(data, str):
                                    data = data.encode("ascii")
                                got = len(data)
                                view[fill:fill + got] = np.frombuffer(data, dtype=np.uint8)
                            eof = got == 0
                            fill += got
                    cut = fill
                    if not eof:                  # cut behind the last complete line
                        lo = max(0, fill - 2 * line_len)
                        nl = np.flatnonzero(view[lo:fill] == 10)
                        if len(nl) == 0 and lo > 0:
                            nl, lo = np.flatnonzero(view[:fill] == 10), 0
                        if len(nl) == 0:
                            if fill >= chunk_bytes:
                                raise ValueError(".card line longer than the %d-byte chunk buffer" % chunk_bytes)
                            cut = 0              # not one complete line yet: read on
                        else:
                            cut = lo + int(nl[-1]) + 1
                    carry = view[cut:fill].tobytes()
                    view[cut] = 0
                    ready_q.put((buf, cut, eof))
            except BaseException as exc:         # noqa: BLE001  (handed to the consumer)
                ready_q.put(exc)
            finally:
                if pool is not None:
                    pool.shutdown()

        thread = threading.Thread(target=reader, daemon=True)
        thread.start()
        try:
            while True:
                item = ready_q.get()
                if isinstance(item, BaseException):
                    raise item
                buf, cut, eof = item
                if cut:
                    done = 0
                    while done < cut:            # (a short max_blocks estimate only costs another call)
                        ts, idx, recs, consumed = self.native.detect_card_ptr(buf.ptr + done, cut - done, final=True)
                        if len(idx):
                            yield ts, idx, recs
                        if consumed == 0:
                            break
                        done += consumed
                free_q.put(buf)
                if eof:
                    break
        finally:
            stop.set()
            free_q.put(None)
            thread.join(timeout=5)
            if not thread.is_alive():            # (a reader stuck in a blocking read keeps its buffers: never recycle
                for b in bufs:                   # page-locked memory a thread may still write to)
                    release_staging(b)           # stays pinned for the next stream of this process

    def detect_card_stream(self, stream, chunk_bytes=None, min_lines=None):
        """Yield (detected, DetectionResult) for every data line of a binary `.card` stream.

        Same results as iterating Detector(settings, card_reader(stream)) (block_data.py:101-131), through
        iter_card_records (no host-side base64 or rawconv)."""
        for ts, _, recs in self.iter_card_records(stream, chunk_bytes, min_lines):
            for pair in records_to_results(recs[:, 0], ts, self.rxid):
                yield pair

    # ---- raw sample streams (`thrifty detect --raw`): no host-side re-blocking
    def detect_raw_stream(self, stream, chunk_blocks=4096, card_out=None):
        """Yield (detected, DetectionResult) for every block of a raw uint8 I/Q stream
        (block_data.py:70-98 semantics: block b = H samples of history + N-H new samples; the history
        that precedes the stream is complex zeros; a trailing partial block is dropped).  The
        overlapping windows are read in place on the GPU (thr_detect_stream), only new samples are
        copied.  The first ceil(H / (N-H)) blocks reach back before the start of the stream: their
        zero history has no uint8 representation, so they go through the complex64 entry point.
        Whatever a read returns is processed (at most `chunk_blocks` blocks per launch): on a live
        pipe the latency is one block, not one batch.
        `card_out`: text stream; every carrier-positive block is also written to it as a `.card` line
        ("<sec>.<usec> <block> <base64 of the 2N raw bytes>", fastcard/fastcard_cli.c:171-193 /
        fastdet/fastdet.cpp:210-219), so the capture can be re-processed later -- the fastcard half of
        the raw-stream front end."""
        import time
        from thrifty_b200.block_data import raw_to_complex
        n, h = self.settings.block_len, self.settings.history_len
        new_s = n - h                                # new samples per block
        read = getattr(stream, "read1", None) or stream.read
        buf = np.zeros(0, dtype=np.uint8)            # stream bytes from sample `pos` on
        pos = 0                                      # global index of the first sample held in buf
        blk = 0                                      # next block to emit
        eof = False
        while not eof:
            data = read(2 * new_s * chunk_blocks)
            if not data:
                eof = True
            else:
                buf = np.concatenate([buf, np.frombuffer(data, dtype=np.uint8)])
            n_avail = (pos + len(buf) // 2) // new_s - blk          # complete blocks not yet emitted
            # blocks whose history starts before the stream: explicit zero history, complex path
            while n_avail > 0 and blk * new_s - h < 0:
                assert pos == 0
                real = raw_to_complex(buf[:2 * (blk + 1) * new_s])
                block = np.concatenate([np.zeros(n - len(real), dtype=np.complex64), real])
                now = time.time()
                pair = self.detect(now, blk, block)[:2]
                if card_out is not None and pair[1].corr_info is not None:      # carrier found
                    from thrifty_b200.block_data import card_line, complex_to_raw
                    card_out.write(card_line(now, blk, complex_to_raw(block)))
                yield pair
                blk += 1
                n_avail -= 1
            while n_avail > 0:
                nb = min(n_avail, chunk_blocks)
                first = blk * new_s - h                              # first sample of block blk
                sub = buf[2 * (first - pos):2 * ((blk + nb) * new_s - pos)]
                recs = self.native.detect_stream(sub, blk)
                now = time.time()
                if card_out is not None:
                    from thrifty_b200.block_data import card_line
                    for k in np.nonzero(recs[:nb, 0]["flags"] & FLAG_CARRIER)[0]:
                        card_out.write(card_line(now, blk + int(k), sub[2 * new_s * k:2 * new_s * k + 2 * n]))
                for pair in records_to_results(recs[:nb, 0], now, self.rxid):
                    yield pair
                blk += nb
                n_avail -= nb
            keep = max(pos, blk * new_s - h)                         # history of the next block stays
            if keep > pos:
                buf = buf[2 * (keep - pos):]
                pos = keep

    # ---- iterator protocol (detect.py:80-91)
    def next(self):
        """Process the next block of data."""
        if self.yield_data:
            return self.detect(*next(self.blocks))
        if not self._pending:
            items = []
            for item in self.blocks:
                items.append(item)
                if len(items) >= self.batch:
                    break
            if not items:
                raise StopIteration
            self._pending = self.detect_many(items)
            self._pending.reverse()
        return self._pending.pop()

    def __iter__(self):
        return self

    def __next__(self):
        return self.next()

    def close(self):
        self.native.close()


class MultiTemplateDetector(object):
    """Joint correlation against several templates (BASELINE.json config 5).

    FFT #1, the carrier fit, the mix and FFT #2 are shared; the multiply / IFFT / peak stage
    runs once per template inside the same kernel pass.  ``detect_many`` returns, per block,
    a list with one (detected, DetectionResult) per template (result.txid = template index).
    Equivalent to running one reference Detector per template on the same block."""

    def __init__(self, settings, templates, rxid=-1, batch=256, device=0):
        templates = np.asarray(templates, dtype=np.float64)
        assert templates.ndim == 2
        self.settings = settings
        self.rxid = rxid
        self.n_templates = templates.shape[0]
        self.native = NativeDetector(
            settings.block_len, settings.history_len, templates, settings.carrier_len,
            settings.carrier_window, settings.carrier_thresh, settings.corr_thresh,
            device=device, max_batch=batch)

    def detect_many(self, items):
        idx = np.array([it[1] for it in items], dtype=np.int64)
        raw = np.stack([np.asarray(it[2], dtype=np.uint8) for it in items])
        recs = self.native.detect_raw(raw, idx)
        out = []
        for i, it in enumerate(items):
            per_tpl = []
            for t in range(self.n_templates):
                detected, res = record_to_result(recs[i, t], it[0], self.rxid)
                res.txid = t
                per_tpl.append((detected, res))
            out.append(per_tpl)
        return out

    def close(self):
        self.native.close()


def _carrier_freq(carrier_info, block_len, sample_rate):
    """Carrier bin + offset -> Hz (detect.py:94-100)."""
    pos = util.fft_bin(carrier_info.bin, block_len) + carrier_info.offset
    return pos * sample_rate / block_len


class SummaryLineFormatter(object):
    """One-line human-readable summary per block (detect.py:103-158)."""

    def __init__(self, sample_rate, block_len, add_dt=False):
        self.sample_rate = sample_rate
        self.block_len = block_len
        self.add_dt = add_dt

    def __call__(self, detected, result):
        carrier_detect = result.corr_info is not None
        ci = result.carrier_info
        info = ("blk={blk}; carrier: {det} @ {freq:.3f} kHz / {idx:>3.0f}:{offset:+.2f}, "
                "SNR = {ampl:>4.0f} / {noise:>2.0f} = {snr:>5.2f} dB".format(
                    blk=result.block, det="yes" if carrier_detect else "no ",
                    freq=_carrier_freq(ci, self.block_len, self.sample_rate) / 1e3,
                    idx=ci.bin, offset=ci.offset, ampl=ci.energy, noise=ci.noise,
                    snr=util.snr(ci.energy, ci.noise)))
        if carrier_detect:
            co = result.corr_info
            info += ("; corr: {det} @ {idx:>4}{offset:+.3f}, "
                     "SNR = {ampl:>4.0f}/{noise:>2.0f} = {snr:>5.2f} dB".format(
                         det="yes" if detected else "no ", idx=co.sample, offset=co.offset,
                         ampl=co.energy, noise=co.noise, snr=util.snr(co.energy, co.noise)))
        return info


def detector_cli(detector_class, parser=None, extra_args=None, argv=None):
    """`thrifty detect` command line (detect.py:161-223); `detector_class` is the plug-in seam."""
    if parser is None:
        parser = argparse.ArgumentParser(description=__doc__,
                                         formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("input", type=argparse.FileType("rb"), default="-",
                        help="input data ('-' streams from stdin)")
    parser.add_argument("--raw", dest="raw", action="store_true", help="input data is raw binary data")
    parser.add_argument("--quiet", dest="quiet", action="store_true",
                        help="do not write anything to standard output")
    parser.add_argument("--batch", dest="batch", type=int, default=256,
                        help="blocks per GPU launch (read-ahead) [default: 256]")
    parser.add_argument("--device", dest="device", type=int, default=0, help="CUDA device ordinal")
    parser.add_argument("--devices", dest="devices", type=parse_devices, default=None,
                        help="several GPUs, e.g. 0-7 or 0,2,3: every batch is striped across them "
                             "(same output, same order)")
    parser.add_argument("--host-decode", dest="host_decode", action="store_true",
                        help="decode the .card base64 payloads on the host (reference behaviour) "
                             "instead of on the GPU")
    parser.add_argument("--card-out", dest="card_out", type=argparse.FileType("w"), default=None,
                        help="with --raw: also write every carrier-positive block as a .card line to this file "
                             "(what fastcard does while capturing)")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("-o", "--output", dest="output", type=argparse.FileType("w"),
                       help="Output file (.toad) ('-' for stdout)")
    group.add_argument("-a", "--append", dest="append", type=argparse.FileType("a"),
                       help="Output file to append to (.toad)")
    setting_keys = ["sample_rate", "block_size", "block_history", "carrier_window",
                    "carrier_threshold", "corr_threshold", "template", "rxid"]
    # `--template a.npy --template b.npy ...`: the setting itself takes one file (the first); repeating the option
    # correlates against all of them jointly (one .toad line per template that detects, txid = position of the template)
    argv = list(sys.argv[1:] if argv is None else argv)
    templates, kept, i = [], [], 0
    while i < len(argv):
        arg = argv[i]
        if arg in ("--template", "-z") and i + 1 < len(argv):
            templates.append(argv[i + 1])
            if len(templates) == 1:
                kept += argv[i:i + 2]
            i += 2
            continue
        if arg.startswith("--template="):
            templates.append(arg.split("=", 1)[1])
            if len(templates) == 1:
                kept.append(arg)
            i += 1
            continue
        kept.append(arg)
        i += 1
    config, args = load_args(parser, setting_keys, argv=kept)

    kwargs = {}
    if extra_args is not None:
        kwargs = {arg: args[arg] for arg in extra_args}
    output_file = args.output if args.append is None else args.append
    info_out = sys.stderr if output_file == sys.stdout else sys.stdout
    bin_freq = config.sample_rate / config.block_size
    window = normalize_freq_range(config.carrier_window, bin_freq)
    if args.raw:
        blocks = block_reader(args.input, config.block_size, config.block_history, raw=True)
    else:
        blocks = card_reader(args.input, raw=True)
    template = np.load(config.template)
    settings = DetectorSettings(block_len=config.block_size, history_len=config.block_history,
                                carrier_len=len(template), carrier_thresh=config.carrier_threshold,
                                carrier_window=window, template=template,
                                corr_thresh=config.corr_threshold)
    if issubclass(detector_class, Detector):
        kwargs.setdefault("batch", args.batch)
        kwargs.setdefault("device", args.device)
        if args.devices:
            kwargs.setdefault("devices", args.devices)
    if detector_class is Detector and len(templates) > 1:
        _multi_template_cli(settings, [np.load(t) for t in templates], blocks, config, args, output_file, info_out)
        return
    import time as _time
    marks = [("start", _time.perf_counter())]            # THRIFTY_B200_CLI_TIMING=1: phase times of the quiet path on stderr
    detections = detector_class(settings, blocks, rxid=config.rxid, **kwargs)
    marks.append(("detector created", _time.perf_counter()))
    if detector_class is Detector and not args.raw and not args.host_decode and args.quiet:
        # fastest path: nothing is printed per block, so no result objects are built either -- records go straight to
        # .toad text (thr_format_toad: the same characters DetectionResult.serialize() produces)
        from thrifty_b200._native import format_toad
        import queue
        import threading
        sink = None
        if output_file is not None:
            output_file.flush()
            sink = getattr(output_file, "buffer", None)
        todo = queue.Queue(maxsize=8)
        failed = []

        def writer():                            # formats and writes chunk c while the GPU works on chunk c+1
            try:
                while True:
                    item = todo.get()
                    if item is None:
                        return
                    text = format_toad(item[0], item[1], config.rxid)
                    if sink is not None:
                        sink.write(text)
                    else:
                        output_file.write(text.decode("ascii"))
            except BaseException as exc:         # noqa: BLE001
                failed.append(exc)

        wthread = threading.Thread(target=writer, daemon=True) if output_file is not None else None
        if wthread is not None:
            wthread.start()
        try:
            gaps = []
            for ts, _, recs in detections.iter_card_records(args.input):
                gaps.append(_time.perf_counter())
                if wthread is not None and not failed:
                    todo.put((recs, ts))
            marks.append(("%d chunks (first after %.1f ms, longest gap %.1f ms)"
                          % (len(gaps), (gaps[0] - marks[-1][1]) * 1e3 if gaps else 0.0,
                             max(np.diff(gaps)) * 1e3 if len(gaps) > 1 else 0.0), _time.perf_counter()))
        finally:
            if wthread is not None:
                todo.put(None)
                wthread.join()
        marks.append(("writer joined", _time.perf_counter()))
        if failed:
            raise failed[0]
        if output_file is not None:
            if sink is not None:
                sink.flush()
            output_file.flush()
        detections.close()
        marks.append(("closed", _time.perf_counter()))
        if os.environ.get("THRIFTY_B200_CLI_TIMING"):
            print("detect --quiet: " + ", ".join("%s +%.1f ms" % (name, (t - marks[i][1]) * 1e3)
                                                 for i, (name, t) in enumerate(marks[1:])), file=sys.stderr)
        return
    if detector_class is Detector and not args.raw and not args.host_decode:
        # fast path: the `.card` text is decoded on the GPU (same records, same order)
        detections = detections.detect_card_stream(args.input)
    elif detector_class is Detector and args.raw and not args.host_decode \
            and (2 * (config.block_size - config.block_history)) % 16 == 0:
        # fast path: overlapping windows are read in place from the contiguous stream
        detections = detections.detect_raw_stream(args.input, chunk_blocks=max(args.batch, 1), card_out=args.card_out)
    summary_liner = SummaryLineFormatter(config.sample_rate, config.block_size, add_dt=True)
    for detected, result in detections:
        if detected and output_file is not None:
            print(result.serialize(), file=output_file)
        if not args.quiet:
            print(summary_liner(detected, result), file=info_out)
    if output_file is not None:
        output_file.flush()
    if args.card_out is not None:
        args.card_out.flush()


def parse_devices(text):
    """'0-7' / '0,2,3' / '1' -> list of CUDA ordinals."""
    out = []
    for part in str(text).split(","):
        lo, sep, hi = part.strip().partition("-")
        out.extend(range(int(lo), int(hi) + 1) if sep else [int(lo)])
    if not out or len(set(out)) != len(out):
        raise argparse.ArgumentTypeError("bad device list %r" % (text,))
    return out


def _multi_template_cli(settings, templates, blocks, config, args, output_file, info_out):
    """`detect --template a.npy --template b.npy ...`: joint correlation (MultiTemplateDetector); every template that
    detects writes its own .toad line with txid = position of the template on the command line."""
    lens = set(len(t) for t in templates)
    if len(lens) != 1:
        raise SystemExit("all templates must have the same length (got %s)" % sorted(lens))
    det = MultiTemplateDetector(settings, np.stack(templates), rxid=config.rxid, batch=args.batch,
                                device=args.devices[0] if args.devices else args.device)
    summary_liner = SummaryLineFormatter(config.sample_rate, config.block_size, add_dt=True)
    items = []

    def flush():
        for per_tpl in det.detect_many(items):
            for detected, result in per_tpl:
                if detected and output_file is not None:
                    print(result.serialize(), file=output_file)
                if not args.quiet:
                    print("tpl=%d; %s" % (result.txid, summary_liner(detected, result)), file=info_out)
        del items[:]

    for item in blocks:
        items.append(item)
        if len(items) >= args.batch:
            flush()
    if items:
        flush()
    if output_file is not None:
        output_file.flush()
    det.close()


def _main(argv=None):
    detector_cli(Detector, argv=argv)


if __name__ == "__main__":
    _main()
