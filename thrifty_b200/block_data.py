"""Block I/O for the detect path: `.card` text and raw RTL-SDR byte streams.

Python-3 counterpart of thrifty/block_data.py:38-131 (the reference module is py2-only).
Readers can yield the raw uint8 payload (`raw=True`) so that rawconv runs on the GPU.
"""
import base64
import time

import numpy as np


def raw_to_complex(data):
    """uint8 I/Q pairs -> complex64: (b - 127.4)/128 per component (block_data.py:38-52)."""
    values = np.asarray(data, dtype=np.uint8).astype(np.float32).view(np.complex64)
    values = values - np.complex64(127.4 + 127.4j)
    return (values / np.float32(128)).astype(np.complex64)


def complex_to_raw(array):
    """complex -> uint8 I/Q pairs: uint8(x*128 + 127.4) (block_data.py:55-67)."""
    scaled = np.asarray(array).astype(np.complex64).view(np.float32) * 128 + 127.4
    return scaled.astype(np.uint8)


def card_reader(stream, raw=False):
    """Yield (timestamp, block_idx, block) per `.card` data line (block_data.py:101-131).

    Accepts text or binary streams.  Comment lines ('#'), blank lines and the
    'Using Volk machine:' / 'linux;' noise lines are skipped.  base64 decoding is lenient
    like the reference's b64decode (tests/test_block_data.py:63 uses a non-canonical payload).
    block is complex64[N] (reference behaviour) or, with raw=True, uint8[2N]."""
    for line in stream:
        if isinstance(line, bytes):
            line = line.decode("ascii", errors="replace")
        if len(line) == 0:
            break
        if line[0] == "#" or line[0] == "\n":
            continue
        if line.startswith("Using Volk machine:") or line.startswith("linux;"):
            continue
        timestamp, idx, encoded = line.rstrip("\n").split(" ")
        payload = np.frombuffer(base64.b64decode(encoded), dtype=np.uint8)
        yield float(timestamp), int(idx), (payload if raw else raw_to_complex(payload))


def card_line(timestamp, block_idx, raw_block):
    """Format one `.card` line like fastcard (fastcard/fastcard_cli.c:184-192)."""
    sec = int(timestamp)
    usec = int(round((timestamp - sec) * 1e6))
    if usec >= 1000000:
        sec, usec = sec + 1, usec - 1000000
    payload = base64.b64encode(np.asarray(raw_block, dtype=np.uint8).tobytes()).decode("ascii")
    return "%d.%06d %d %s\n" % (sec, usec, block_idx, payload)


def write_card(stream, raw_blocks, block_indices=None, t0=1480000000.0, dt=0.0047767, header=None):
    """Write a fastcard-style `.card` file (header per fastcard/fargs.c:194-214)."""
    hdr = header or {}
    stream.write("# arguments: { carrier_bin: '%s', threshold: '%s', block_size: %d, history_size: %d }\n"
                 % (hdr.get("carrier_bin", "7-110"), hdr.get("threshold", "100c2s"),
                    len(raw_blocks[0]) // 2, hdr.get("history_size", 4920)))
    stream.write("# tool: 'thrifty_b200 synth'\n")
    stream.write("# start_time: %.6f\n" % t0)
    for i, blk in enumerate(raw_blocks):
        idx = i if block_indices is None else int(block_indices[i])
        stream.write(card_line(t0 + i * dt, idx, blk))


def block_reader(stream, size, history, raw=False):
    """Re-block a raw uint8 I/Q stream with `history` overlap (block_data.py:70-98).

    The first block's history is complex zeros (reference behaviour).  No byte value maps to
    exactly 0, so in raw mode block 0 is still yielded as complex64 (the detector accepts
    either type); later blocks are uint8[2*size].  A trailing partial block is dropped."""
    new = size - history
    if raw:
        data = np.full(2 * size, 127, dtype=np.uint8)
    else:
        data = np.zeros(size, dtype=np.complex64)
    block_idx = 0
    while True:
        chunk = stream.read(new * 2)
        if len(chunk) < new * 2:
            break
        payload = np.frombuffer(chunk, dtype=np.uint8)
        if raw:
            data = np.concatenate([data[2 * new:], payload])
        else:
            data = np.concatenate([data[new:], raw_to_complex(payload)])
        if raw and block_idx == 0 and history > 0:
            first = raw_to_complex(data)
            first[:history] = 0
            yield time.time(), block_idx, first
        else:
            yield time.time(), block_idx, data
        block_idx += 1
