"""`fastdet`-compatible front end: the semantics and command line of the reference's native twin.

The reference ships a second implementation of the detect path, ``fastdet`` (C++11 on top of the
``fastcard`` C library), with its own semantics -- decisions on powers, integer-bin carrier shift,
parabolic carrier offset, +-0.5 clip.  This module exposes the same behaviour on the
GPU (``THR_CFG_FASTDET_SEMANTICS`` kernels: two transforms per block instead of three):

  * ``FastDetector``      -- ``CarrierDetector`` + ``CorrDetector`` of fastdet/corr_detector.h:24-63 and
                            fastdet/fastcard_wrappers.h:28-46 rolled into one batched object,
  * ``parse_threshold``   -- '<constant>c<snr>s' strings (fastcard/parse.c:54-99),
  * ``parse_carrier_window`` -- '<min>-<max>' (fastcard/parse.c:38-52),
  * ``load_template`` / ``save_template`` -- the `.tpl` format (fastdet/corr_detector.cpp:200-228,
                            scripts/npy_to_tpl.py:18-22),
  * ``toad_line`` / ``info_line`` -- fastdet's output formats (fastdet/fastdet.cpp:191-206, 222-247),
  * ``main``              -- ``python -m thrifty_b200 fastdet`` with fastdet's option letters
                            (fastdet/fastdet.cpp:30-44, fastcard/fargs.c:29-77).

There is no CPU fallback: everything numeric runs in the CUDA kernel behind the C ABI.
"""

from __future__ import print_function

import argparse
import re
import struct
import sys

import numpy as np

from thrifty_b200._native import FLAG_CARRIER, FLAG_CORR, NativeDetector

# defaults of fastcard/fargs.c:6-19 and fastdet/fastdet.cpp:46-52
DEFAULT_BLOCK_LEN = 16384
DEFAULT_HISTORY_LEN = 4920
DEFAULT_THRESHOLD = (100.0, 2.0)
DEFAULT_CORR_THRESHOLD = (0.0, 15.0)
DEFAULT_WINDOW = (0, -1)


def parse_threshold(arg):
    """'<constant>c<snr>s' -> (constant, snr); a bare number is the constant (fastcard/parse.c:54-99)."""
    constant = snr = 0.0
    got_c = got_s = False
    pos = 0
    num = re.compile(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)")
    while True:
        m = num.match(arg, pos)
        if not m:
            break
        value = float(m.group(0))
        pos = m.end()
        nxt = arg[pos:pos + 1]
        if nxt == "c" or nxt == "":
            if got_c:
                raise ValueError("Argument '--threshold' contains more than one value for constant.")
            constant, got_c = value, True
            pos += 1 if nxt == "c" else 0
        elif nxt == "s":
            if got_s:
                raise ValueError("Argument '--threshold' contains more than one value for SNR.")
            snr, got_s = value, True
            pos += 1
        # any other character: like the C parser, stop at the next failed number scan
    if pos != len(arg):
        raise ValueError("Argument '--threshold' contains an invalid value.")
    return constant, snr


def parse_carrier_window(arg):
    """'<min>-<max>' (signed integers, e.g. '7-110', '0--1', '-110--7') -> (min, max); a single
    number means min == max (fastcard/parse.c:38-52, sscanf "%d-%d")."""
    m = re.match(r"\s*([-+]?\d+)(?:-([-+]?\d+))?", arg)
    if not m:
        raise ValueError("Argument '--carrier' contains an invalid value.")
    lo = int(m.group(1))
    hi = int(m.group(2)) if m.group(2) is not None else lo
    return lo, hi


def load_template(filename):
    """`.tpl`: native-endian uint16 length followed by that many float32 samples
    (fastdet/corr_detector.cpp:200-228)."""
    with open(filename, "rb") as f:
        head = f.read(2)
        if len(head) != 2:
            raise RuntimeError("Failed to load template: short read")
        (length,) = struct.unpack("=H", head)
        payload = f.read(4 * length)
    if len(payload) != 4 * length:
        raise RuntimeError("Failed to load template: short read")
    return np.frombuffer(payload, dtype=np.float32).copy()


def save_template(filename, samples):
    """Inverse of load_template (scripts/npy_to_tpl.py:18-22)."""
    samples = np.asarray(samples, dtype=np.float32)
    if len(samples) > 0xFFFF:
        raise ValueError("template too long for the .tpl format")
    with open(filename, "wb") as f:
        f.write(struct.pack("=H", len(samples)))
        f.write(samples.tobytes())


def toad_line(rec, timestamp, rxid):
    """One `.toad` line as fastdet prints it (fastdet/fastdet.cpp:191-206)."""
    sec = int(timestamp)
    usec = int(round((timestamp - sec) * 1e6))
    if usec >= 1000000:
        sec, usec = sec + 1, usec - 1000000
    return "%d %d.%06d %d %.8f %u %.12f %f %f %u %f %f %f" % (
        rxid, sec, usec, rec["block_idx"], rec["soa"], rec["corr_sample"], rec["corr_offset"],
        rec["corr_energy"], rec["corr_noise"], rec["carrier_bin"], rec["carrier_offset"],
        rec["carrier_energy"], rec["carrier_noise"])


def info_line(rec, thresh, corr_thresh):
    """Per-block summary as fastdet prints it for carrier-positive blocks (fastdet/fastdet.cpp:222-247)."""
    cmax, cnoise = float(rec["carrier_energy"]) ** 2, float(rec["carrier_noise"]) ** 2
    cthr = thresh[0] + thresh[1] * cnoise
    with np.errstate(divide="ignore", invalid="ignore"):
        line = "block #%d: carrier @ %3u %+.1f = %4.0f / %2.0f [>%2.0f] = %2.0f dB" % (
            rec["block_idx"], rec["carrier_bin"], rec["carrier_offset"], np.sqrt(cmax), np.sqrt(cnoise),
            np.sqrt(cthr), 10 * np.log10(np.float64(cmax) / np.float64(cnoise)))
        if rec["flags"] & FLAG_CORR:
            pk, pn = float(rec["corr_energy"]) ** 2, float(rec["corr_noise"]) ** 2
            kthr = corr_thresh[0] + corr_thresh[1] * pn
            line += "; corr = %4.0f / %2.0f [>%2.0f] = %2.0f dB" % (
                np.sqrt(pk), np.sqrt(pn), np.sqrt(kthr), 10 * np.log10(np.float64(pk) / np.float64(pn)))
    return line


class FastDetector(object):
    """Batched GPU twin of fastdet's CarrierDetector + CorrDetector.

    template: float32 samples (a `.tpl`); thresh / corr_thresh: (constant, snr) on POWERS;
    carrier_window: (min, max) signed bins, closed, must not straddle zero (fastcard/cardet.c:43-69)."""

    def __init__(self, template, block_len=DEFAULT_BLOCK_LEN, history_len=DEFAULT_HISTORY_LEN,
                 thresh=DEFAULT_THRESHOLD, carrier_window=DEFAULT_WINDOW, corr_thresh=DEFAULT_CORR_THRESHOLD,
                 rxid=-1, batch=256, device=0):
        if history_len > block_len:                       # fastcard/fastcard.c:14-17
            raise ValueError("History length cannot be larger than block length.")
        tpl = np.asarray(template, dtype=np.float32).astype(np.float64)
        self.block_len, self.history_len = int(block_len), int(history_len)
        self.thresh = (float(thresh[0]), float(thresh[1]))
        self.corr_thresh = (float(corr_thresh[0]), float(corr_thresh[1]))
        self.rxid = rxid
        self.batch = max(1, int(batch))
        self.native = NativeDetector(self.block_len, self.history_len, tpl, len(tpl), carrier_window,
                                     (self.thresh[0], self.thresh[1], 0.0),
                                     (self.corr_thresh[0], self.corr_thresh[1], 0.0),
                                     device=device, max_batch=self.batch, fastdet=True)

    def detect_raw(self, raw_blocks, block_idx=None):
        """uint8 [B, 2N] -> thr_record array [B]."""
        return self.native.detect_raw(raw_blocks, block_idx)[:, 0]

    def detect_card(self, stream):
        """Yield (timestamp, record) per data line of a binary `.card` stream, GPU-side base64 decode."""
        text = stream.read()
        if isinstance(text, str):
            text = text.encode("ascii")
        pos = 0
        line_len = ((2 * self.block_len + 2) // 3) * 4 + 64
        step = max(self.batch, 1) * line_len
        while pos < len(text):
            chunk = text[pos:pos + step]
            final = pos + step >= len(text)
            ts, _, recs, consumed = self.native.detect_card(chunk, final=final)
            for i in range(len(ts)):
                yield float(ts[i]), recs[i, 0]
            if consumed == 0 and not final:
                step *= 2
                continue
            pos += consumed if consumed else len(chunk)

    def detect_stream(self, stream_bytes, first_block=0):
        """Contiguous raw uint8 I/Q that starts with the history of block `first_block`."""
        return self.native.detect_stream(stream_bytes, first_block)[:, 0]

    def close(self):
        self.native.close()


def main(argv=None):
    parser = argparse.ArgumentParser(prog="fastdet", add_help=False,
                                     description="FastDet: Fast Detector -- like Thrifty, but faster (on a B200).")
    parser.add_argument("--help", action="help")
    parser.add_argument("-i", "--input", default="-", help="input file with samples ('-' for stdin)")
    parser.add_argument("--card", action="store_true", help="input is a .card file instead of binary data")
    parser.add_argument("-o", "--output", default=None, help="output toad file ('-' for stdout)")
    parser.add_argument("-b", "--block-len", type=int, default=DEFAULT_BLOCK_LEN)
    parser.add_argument("-h", "--history", type=int, default=DEFAULT_HISTORY_LEN)
    parser.add_argument("-k", "--skip", type=int, default=1, help="blocks to skip (raw input only)")
    parser.add_argument("-w", "--carrier-window", default="0--1")
    parser.add_argument("-t", "--threshold", default="100c2s")
    parser.add_argument("-u", "--corr-threshold", default="15s")
    parser.add_argument("-z", "--template", default="template.tpl")
    parser.add_argument("-r", "--rxid", type=int, default=-1)
    parser.add_argument("-q", "--quiet", action="store_true")
    parser.add_argument("--batch", type=int, default=256)
    parser.add_argument("--device", type=int, default=0)
    # argp takes "-w -110--7"; argparse would read the negative window as an option: glue such values with '='
    argv = list(sys.argv[1:] if argv is None else argv)
    glued, i = [], 0
    while i < len(argv):
        if argv[i] in ("-w", "--carrier-window") and i + 1 < len(argv) and re.match(r"-\d", argv[i + 1]):
            glued.append("--carrier-window=" + argv[i + 1])
            i += 2
        else:
            glued.append(argv[i])
            i += 1
    args = parser.parse_args(glued)

    thresh = parse_threshold(args.threshold)
    corr_thresh = parse_threshold(args.corr_threshold)
    window = parse_carrier_window(args.carrier_window)
    det = FastDetector(load_template(args.template), args.block_len, args.history, thresh, window, corr_thresh,
                       rxid=args.rxid, batch=args.batch, device=args.device)
    out = None
    if args.output is not None:
        out = sys.stdout if args.output == "-" else open(args.output, "w")
    info = None if args.quiet else (sys.stderr if out is sys.stdout else sys.stdout)
    inp = sys.stdin.buffer if args.input == "-" else open(args.input, "rb")
    count = 0

    def emit(timestamp, rec):
        if not (rec["flags"] & FLAG_CARRIER):
            return
        if (rec["flags"] & FLAG_CORR) and out is not None:
            print(toad_line(rec, timestamp, args.rxid), file=out)
        if info is not None:
            print(info_line(rec, thresh, corr_thresh), file=info)

    if args.card:
        # fastcard.c:70-72: no blocks are skipped when reading a .card
        for timestamp, rec in det.detect_card(inp):
            count += 1
            emit(timestamp, rec)
    else:
        import time
        n, h = args.block_len, args.history
        new = 2 * (n - h)
        hist = np.zeros(2 * h, dtype=np.uint8)
        hist[0::2] = 127                                   # reader.c:56-59: uint16 127 per sample
        # fastcard.c:108-110: with --skip the first blocks get negative indices, block 0 is the first kept one
        tail, first_block, skip = hist, -args.skip, args.skip
        while True:
            data = inp.read(new * args.batch)
            nblk = len(data) // new
            if nblk == 0:
                break
            buf = np.concatenate([tail, np.frombuffer(data[:nblk * new], dtype=np.uint8)])
            if (new % 16) == 0:
                recs = det.detect_stream(buf, first_block)
            else:   # TMA tiles need 16-byte aligned block starts: re-block on the host
                blocks = np.stack([buf[b * new:b * new + 2 * n] for b in range(nblk)])
                recs = det.detect_raw(blocks, first_block + np.arange(nblk))
            now = time.time()
            for rec in recs:
                if skip > 0:
                    skip -= 1
                    continue
                count += 1
                emit(now, rec)
            tail = buf[len(buf) - 2 * h:] if h else buf[:0]
            first_block += nblk
    if info is not None:
        print("\nRead %d blocks." % count, file=info)
    if out is not None and out is not sys.stdout:
        out.close()
    det.close()


def _main():
    main()


if __name__ == "__main__":
    main()
