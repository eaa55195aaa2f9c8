"""ctypes binding of libthrifty_b200.so (the C ABI in include/thrifty_b200.h).

There is deliberately no fallback: if the shared object is missing or no sm_100
device is usable, constructing a detector raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C thrifty_b200/csrc``.
"""

from __future__ import annotations

import atexit
import ctypes
import os
import threading
from ctypes import (POINTER, Structure, byref, c_char, c_char_p, c_double, c_float, c_int,
                    c_int32, c_int64, c_size_t, c_uint32, c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libthrifty_b200.so")

THR_OK = 0
FLAG_CARRIER = 1
FLAG_CORR = 2


class NativeError(RuntimeError):
    """A thr_* call returned a negative status."""


class ThrConfig(Structure):
    _fields_ = [
        ("block_len", c_int32), ("history_len", c_int32), ("template_len", c_int32),
        ("n_templates", c_int32), ("templates", POINTER(c_double)), ("carrier_len", c_int32),
        ("window_start", c_int32), ("window_stop", c_int32),
        ("carrier_thresh", c_double * 3), ("corr_thresh", c_double * 3),
        ("device", c_int32), ("max_batch", c_int32), ("flags", c_uint32), ("reserved", c_int32),
    ]


class ThrInfo(Structure):
    _fields_ = [
        ("abi_version", c_int32), ("device", c_int32), ("sm_count", c_int32), ("grid", c_int32),
        ("threads", c_int32), ("smem_bytes", c_int32), ("ctas_per_sm", c_int32),
        ("buffer_in_smem", c_int32), ("launches", c_int64),
        ("device_name", c_char * 64), ("kernel", c_char * 64),
    ]


# numpy view of thr_record (64 bytes)
RECORD_DTYPE = np.dtype([
    ("block_idx", "<i8"), ("soa", "<f8"),
    ("carrier_bin", "<i4"), ("carrier_offset", "<f4"), ("carrier_energy", "<f4"), ("carrier_noise", "<f4"),
    ("corr_sample", "<i4"), ("corr_offset", "<f4"), ("corr_energy", "<f4"), ("corr_noise", "<f4"),
    ("flags", "<u4"), ("template_idx", "<i4"), ("signal_energy", "<f4"), ("reserved", "<f4"),
])
assert RECORD_DTYPE.itemsize == 64

EXPORTS = [
    "thr_create", "thr_destroy", "thr_last_error", "thr_get_info", "thr_device_count",
    "thr_detect_batch", "thr_detect_batch_c64", "thr_detect_batch_device",
    "thr_detect_batch_device_c64", "thr_detect_block_data", "thr_set_stream", "thr_synchronize",
    "thr_timer_start", "thr_timer_stop", "thr_host_alloc", "thr_host_free", "thr_device_alloc",
    "thr_device_free", "thr_memcpy_h2d", "thr_memcpy_d2h", "thr_card_scan", "thr_detect_card",
    "thr_detect_stream", "thr_detect_stream_device", "thr_sync_batch", "thr_soa_batch",
    "thr_group_create", "thr_group_destroy", "thr_group_last_error", "thr_group_size", "thr_group_member",
    "thr_group_numa_node", "thr_group_detect_batch", "thr_group_detect_stream", "thr_group_detect_card",
    "thr_group_host_alloc", "thr_group_host_free",
    "thr_identify_classify", "thr_identify_bin_histogram", "thr_identify_digitize", "thr_identify_duplicates",
    "thr_identify_last_error", "thr_format_toad",
]

_lib = None


def load_library(path=None):
    """dlopen the shared object and declare prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # THRIFTY_B200_LIB: alternative build of the same library (kernel experiments, tools/variants.sh)
    path = path or os.environ.get("THRIFTY_B200_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise NativeError(
            "libthrifty_b200.so not found at %s: build it (python -c 'import __graft_entry__ as g; "
            "g.build()'); there is no CPU fallback for the detect path" % path)
    lib = ctypes.CDLL(path)
    lib.thr_create.argtypes = [POINTER(ThrConfig), POINTER(c_void_p)]
    lib.thr_create.restype = c_int
    lib.thr_destroy.argtypes = [c_void_p]
    lib.thr_destroy.restype = None
    lib.thr_last_error.argtypes = [c_void_p]
    lib.thr_last_error.restype = c_char_p
    lib.thr_get_info.argtypes = [c_void_p, POINTER(ThrInfo)]
    lib.thr_get_info.restype = c_int
    lib.thr_device_count.argtypes = []
    lib.thr_device_count.restype = c_int
    lib.thr_detect_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.thr_detect_batch.restype = c_int
    lib.thr_detect_batch_c64.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.thr_detect_batch_c64.restype = c_int
    lib.thr_detect_batch_device.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.thr_detect_batch_device.restype = c_int
    lib.thr_detect_batch_device_c64.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.thr_detect_batch_device_c64.restype = c_int
    lib.thr_detect_block_data.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_void_p]
    lib.thr_detect_block_data.restype = c_int
    lib.thr_card_scan.argtypes = [c_char_p, c_size_t, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                  POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    lib.thr_card_scan.restype = c_int
    lib.thr_detect_card.argtypes = [c_void_p, c_char_p, c_size_t, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                    POINTER(c_int64), POINTER(c_int64)]
    lib.thr_detect_card.restype = c_int
    lib.thr_detect_stream.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_void_p, POINTER(c_int64)]
    lib.thr_detect_stream.restype = c_int
    lib.thr_detect_stream_device.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p]
    lib.thr_detect_stream_device.restype = c_int
    try:
        lib.thr_format_toad.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_size_t,
                                        POINTER(c_size_t)]
        lib.thr_format_toad.restype = c_int
    except AttributeError:
        if path == LIB_PATH:
            raise
    try:        # entry points added in round 2 (an older experiment build selected by THRIFTY_B200_LIB may lack them)
        lib.thr_sync_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
        lib.thr_sync_batch.restype = c_int
        lib.thr_soa_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
        lib.thr_soa_batch.restype = c_int
        lib.thr_group_create.argtypes = [POINTER(ThrConfig), POINTER(c_int32), c_int32, POINTER(c_void_p)]
        lib.thr_group_create.restype = c_int
        lib.thr_group_destroy.argtypes = [c_void_p]
        lib.thr_group_destroy.restype = None
        lib.thr_group_last_error.argtypes = [c_void_p]
        lib.thr_group_last_error.restype = c_char_p
        lib.thr_group_size.argtypes = [c_void_p]
        lib.thr_group_size.restype = c_int
        lib.thr_group_member.argtypes = [c_void_p, c_int32]
        lib.thr_group_member.restype = c_void_p
        lib.thr_group_numa_node.argtypes = [c_void_p, c_int32, POINTER(c_int32)]
        lib.thr_group_numa_node.restype = c_int
        lib.thr_group_detect_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
        lib.thr_group_detect_batch.restype = c_int
        lib.thr_group_detect_stream.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_void_p, POINTER(c_int64)]
        lib.thr_group_detect_stream.restype = c_int
        lib.thr_group_detect_card.argtypes = [c_void_p, c_char_p, c_size_t, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                              POINTER(c_int64), POINTER(c_int64)]
        lib.thr_group_detect_card.restype = c_int
        lib.thr_group_host_alloc.argtypes = [c_void_p, c_size_t]
        lib.thr_group_host_alloc.restype = c_void_p
        lib.thr_group_host_free.argtypes = [c_void_p, c_void_p, c_size_t]
        lib.thr_group_host_free.restype = None
    except AttributeError:
        if path == LIB_PATH:
            raise
    lib.thr_set_stream.argtypes = [c_void_p, c_void_p]
    lib.thr_set_stream.restype = c_int
    lib.thr_synchronize.argtypes = [c_void_p]
    lib.thr_synchronize.restype = c_int
    lib.thr_timer_start.argtypes = [c_void_p]
    lib.thr_timer_start.restype = c_int
    lib.thr_timer_stop.argtypes = [c_void_p, POINTER(c_float)]
    lib.thr_timer_stop.restype = c_int
    lib.thr_host_alloc.argtypes = [c_size_t]
    lib.thr_host_alloc.restype = c_void_p
    lib.thr_host_free.argtypes = [c_void_p]
    lib.thr_host_free.restype = None
    lib.thr_device_alloc.argtypes = [c_int, c_size_t]
    lib.thr_device_alloc.restype = c_void_p
    lib.thr_device_free.argtypes = [c_int, c_void_p]
    lib.thr_device_free.restype = None
    lib.thr_memcpy_h2d.argtypes = [c_int, c_void_p, c_void_p, c_size_t]
    lib.thr_memcpy_h2d.restype = c_int
    lib.thr_memcpy_d2h.argtypes = [c_int, c_void_p, c_void_p, c_size_t]
    lib.thr_memcpy_d2h.restype = c_int
    if path == LIB_PATH:
        _lib = lib
    return lib


def format_toad(records, timestamps, rxid, txids=None):
    """thr_record array [B] or [B, T] (template 0 is used) + timestamps [B] -> bytes of .toad lines for the detected
    blocks (thr_format_toad: the text DetectionResult.serialize() would give, formatted natively)."""
    lib = load_library()
    records = np.ascontiguousarray(records)
    stride = 1 if records.ndim == 1 else records.shape[1]
    n = records.shape[0]
    ts = np.ascontiguousarray(np.broadcast_to(np.asarray(timestamps, dtype=np.float64), (n,)))
    tx = None if txids is None else np.ascontiguousarray(txids, dtype=np.int32)
    if n == 0:
        return b""
    n_det = int(((records.reshape(n, -1)[:, 0]["flags"] & FLAG_CORR) != 0).sum())
    buf = ctypes.create_string_buffer(max(1, n_det * 1024))
    used = c_size_t(0)
    rc = lib.thr_format_toad(records.ctypes.data, ts.ctypes.data, n, stride, int(rxid),
                             None if tx is None else tx.ctypes.data, buf, len(buf), byref(used))
    if rc != THR_OK:
        raise NativeError("thr_format_toad failed (%d)" % rc)
    return buf.raw[:used.value]


class PinnedBuffer(object):
    """Page-locked host memory exposed as a numpy uint8 array (thr_host_alloc)."""

    def __init__(self, nbytes):
        self._lib = load_library()
        self.ptr = self._lib.thr_host_alloc(nbytes)
        if not self.ptr:
            raise NativeError("thr_host_alloc(%d) failed" % nbytes)
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            self._lib.thr_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# Page-locked staging buffers of the stream readers, kept for the life of the process: pinning costs 0.6-1 ms per MiB and
# cudaFreeHost / cudaMallocHost now and then stall for 0.2-0.7 s (measured: tools/cli_stalls.py), so a process that
# reads many `.card` streams pins its two buffers once.  At most _STAGING_KEEP buffers per size and _STAGING_MAX_BYTES in
# total stay cached (page-locked memory is not swappable).
_STAGING_KEEP = 4
_STAGING_MAX_BYTES = 256 << 20
_staging_pool = {}
_staging_lock = threading.Lock()


def acquire_staging(nbytes):
    """A PinnedBuffer of exactly `nbytes` from the per-process cache, or a new one."""
    with _staging_lock:
        cached = _staging_pool.get(nbytes)
        if cached:
            return cached.pop()
    return PinnedBuffer(nbytes)


def release_staging(buf):
    """Hand a buffer from acquire_staging back (it stays page-locked for the next reader)."""
    if buf is None or not buf.ptr:
        return
    if getattr(buf, "nbytes", None) is None:     # (a test double)
        buf.close()
        return
    with _staging_lock:
        cached = _staging_pool.setdefault(buf.nbytes, [])
        total = sum(size * len(bufs) for size, bufs in _staging_pool.items())
        if len(cached) < _STAGING_KEEP and total + buf.nbytes <= _STAGING_MAX_BYTES:
            cached.append(buf)
            return
    buf.close()


def _forget_staging():
    # interpreter exit: the driver unmaps everything with the process; an explicit cudaFreeHost per buffer only adds
    # wall clock to the end of a command-line run
    with _staging_lock:
        for cached in _staging_pool.values():
            for buf in cached:
                buf.array = None
                buf.ptr = None
        _staging_pool.clear()


atexit.register(_forget_staging)


def _make_config(self, block_len, history_len, templates, carrier_len, carrier_window, carrier_thresh, corr_thresh,
                 device, max_batch, overlap_launches, fastdet, generic_kernel):
    """Fill a ThrConfig (and the matching attributes of `self`); returns (cfg, template array to keep alive)."""
    tpl = np.ascontiguousarray(np.atleast_2d(np.asarray(templates, dtype=np.float64)))
    self.n_templates, self.template_len = tpl.shape
    self.block_len = int(block_len)
    self.history_len = int(history_len)
    self.max_batch = int(max_batch)
    self.device = int(device)
    cfg = ThrConfig()
    cfg.block_len = self.block_len
    cfg.history_len = self.history_len
    cfg.template_len = self.template_len
    cfg.n_templates = self.n_templates
    cfg.templates = tpl.ctypes.data_as(POINTER(c_double))
    cfg.carrier_len = int(carrier_len)
    win = (0, -1) if carrier_window is None else carrier_window
    cfg.window_start, cfg.window_stop = int(win[0]), int(win[1])
    cfg.carrier_thresh = (c_double * 3)(*[float(v) for v in carrier_thresh])
    cfg.corr_thresh = (c_double * 3)(*[float(v) for v in corr_thresh])
    cfg.device = self.device
    cfg.max_batch = self.max_batch
    # THR_CFG_OVERLAP_LAUNCHES | THR_CFG_FASTDET_SEMANTICS (the native twin's semantics:
    # thresholds (constant, snr) apply to POWERS, fastcard/parse.c:54-99 '<c>c<s>s')
    # THR_CFG_GENERIC_KERNEL: block_len 32768 on the generic global-scratch kernel (comparisons)
    cfg.flags = (1 if overlap_launches else 0) | (2 if fastdet else 0) | (4 if generic_kernel else 0)
    self.fastdet = bool(fastdet)
    return cfg, tpl


def _raise_create_error(what, rc, msg):
    if ("out of range" in msg and "window" in msg) or "window range not supported" in msg:
        raise ValueError(msg)            # carrier_detect.py:47-49 raises ValueError
    raise NativeError("%s failed (%d): %s" % (what, rc, msg))


class NativeDetector(object):
    """Thin owner of a thr_detector handle."""

    def __init__(self, block_len, history_len, templates, carrier_len, carrier_window,
                 carrier_thresh, corr_thresh, device=0, max_batch=4096, overlap_launches=False,
                 fastdet=False, generic_kernel=False):
        self._lib = load_library()
        self._h = c_void_p()
        self._owned = True
        cfg, _tpl = _make_config(self, block_len, history_len, templates, carrier_len, carrier_window, carrier_thresh,
                                 corr_thresh, device, max_batch, overlap_launches, fastdet, generic_kernel)
        rc = self._lib.thr_create(byref(cfg), byref(self._h))
        if rc != THR_OK:
            msg = self._lib.thr_last_error(None).decode()
            self._h = c_void_p()
            _raise_create_error("thr_create", rc, msg)

    # -- helpers
    def _check(self, rc):
        if rc != THR_OK:
            raise NativeError("thrifty_b200 call failed (%d): %s"
                              % (rc, self._lib.thr_last_error(self._h).decode()))

    @property
    def handle(self):
        return self._h

    def info(self):
        info = ThrInfo()
        self._check(self._lib.thr_get_info(self._h, byref(info)))
        return {k: (getattr(info, k).decode() if isinstance(getattr(info, k), bytes) else getattr(info, k))
                for k, _ in ThrInfo._fields_}

    def close(self):
        if self._h:
            if self._owned:
                self._lib.thr_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host-buffer API
    def detect_raw(self, raw, block_idx=None):
        """raw: uint8 [B, 2N] (C-contiguous).  Returns records [B, n_templates]."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        if raw.ndim != 2 or raw.shape[1] != 2 * self.block_len:
            raise ValueError("raw must have shape [B, %d]" % (2 * self.block_len))
        nblk = raw.shape[0]
        out = np.zeros((nblk, self.n_templates), dtype=RECORD_DTYPE)
        idx_ptr = None
        if block_idx is not None:
            idx = np.ascontiguousarray(block_idx, dtype=np.int64)
            assert idx.shape == (nblk,)
            idx_ptr = idx.ctypes.data
        self._check(self._lib.thr_detect_batch(self._h, raw.ctypes.data, idx_ptr, nblk, out.ctypes.data))
        return out

    def detect_c64(self, iq, block_idx=None):
        """iq: complex64 [B, N].  Returns records [B, n_templates]."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim != 2 or iq.shape[1] != self.block_len:
            raise ValueError("iq must have shape [B, %d]" % self.block_len)
        nblk = iq.shape[0]
        out = np.zeros((nblk, self.n_templates), dtype=RECORD_DTYPE)
        idx_ptr = None
        if block_idx is not None:
            idx = np.ascontiguousarray(block_idx, dtype=np.int64)
            assert idx.shape == (nblk,)
            idx_ptr = idx.ctypes.data
        self._check(self._lib.thr_detect_batch_c64(self._h, iq.ctypes.data, idx_ptr, nblk, out.ctypes.data))
        return out

    def detect_block_data(self, raw=None, iq=None, block_idx=0):
        """One block with intermediates -> (record[n_templates], shifted_fft, corr, fft_mag)."""
        n = self.block_len
        corr_len = n - self.template_len + 1
        out = np.zeros(self.n_templates, dtype=RECORD_DTYPE)
        sfft = np.zeros(n, dtype=np.complex64)
        corr = np.zeros(corr_len, dtype=np.complex64)
        mag = np.zeros(n, dtype=np.float32)
        raw_ptr = iq_ptr = None
        if raw is not None:
            raw = np.ascontiguousarray(raw, dtype=np.uint8)
            assert raw.shape == (2 * n,)
            raw_ptr = raw.ctypes.data
        else:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
            assert iq.shape == (n,)
            iq_ptr = iq.ctypes.data
        self._check(self._lib.thr_detect_block_data(self._h, raw_ptr, iq_ptr, int(block_idx), out.ctypes.data,
                                                    sfft.ctypes.data, corr.ctypes.data, mag.ctypes.data))
        return out, sfft, corr, mag

    # -- stage boundaries (carrier_sync.Synchronizer / soa_estimator.SoaEstimator)
    def sync_batch(self, raw=None, iq=None, block_idx=None):
        """Blocks -> (records [B] with the carrier fields, shifted spectra complex64 [B, N]) (thr_sync_batch)."""
        n = self.block_len
        raw_ptr = iq_ptr = None
        if raw is not None:
            raw = np.ascontiguousarray(raw, dtype=np.uint8)
            assert raw.ndim == 2 and raw.shape[1] == 2 * n
            nblk, raw_ptr = raw.shape[0], raw.ctypes.data
        else:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
            assert iq.ndim == 2 and iq.shape[1] == n
            nblk, iq_ptr = iq.shape[0], iq.ctypes.data
        idx_ptr = None
        if block_idx is not None:
            idx = np.ascontiguousarray(block_idx, dtype=np.int64)
            assert idx.shape == (nblk,)
            idx_ptr = idx.ctypes.data
        out = np.zeros(nblk, dtype=RECORD_DTYPE)
        sfft = np.zeros((nblk, n), dtype=np.complex64)
        self._check(self._lib.thr_sync_batch(self._h, raw_ptr, iq_ptr, idx_ptr, nblk, out.ctypes.data, sfft.ctypes.data))
        return out, sfft

    def soa_batch(self, fft, block_idx=None, want_corr=True):
        """Shifted spectra complex64 [B, N] -> (records [B] with the corr fields, corr complex64 [B, N-L+1] or None)
        (thr_soa_batch)."""
        n = self.block_len
        fft = np.ascontiguousarray(fft, dtype=np.complex64)
        assert fft.ndim == 2 and fft.shape[1] == n
        nblk = fft.shape[0]
        idx_ptr = None
        if block_idx is not None:
            idx = np.ascontiguousarray(block_idx, dtype=np.int64)
            assert idx.shape == (nblk,)
            idx_ptr = idx.ctypes.data
        out = np.zeros(nblk, dtype=RECORD_DTYPE)
        corr = np.zeros((nblk, n - self.template_len + 1), dtype=np.complex64) if want_corr else None
        self._check(self._lib.thr_soa_batch(self._h, fft.ctypes.data, idx_ptr, nblk, out.ctypes.data,
                                            corr.ctypes.data if want_corr else None))
        return out, corr

    def detect_stream(self, stream, first_block):
        """Contiguous uint8 I/Q stream that starts with the history of block `first_block`.
        Returns records [B, n_templates] for the B whole blocks it contains (thr_detect_stream)."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        new = 2 * (self.block_len - self.history_len)
        nblk = (len(stream) - 2 * self.block_len) // new + 1 if len(stream) >= 2 * self.block_len else 0
        out = np.zeros((max(nblk, 0), self.n_templates), dtype=RECORD_DTYPE)
        got = c_int64(0)
        self._check(self._lib.thr_detect_stream(self._h, stream.ctypes.data, len(stream), int(first_block),
                                                out.ctypes.data, byref(got)))
        assert got.value == len(out)
        return out

    def detect_card(self, text, final=True, max_blocks=None):
        """`.card` text (bytes) -> (timestamps f8[B], block_idx i8[B], records [B, n_templates], consumed).

        Lines are scanned on the host, the base64 payloads are decoded on the GPU (thr_detect_card).
        With final=False an unterminated last line is left for the next call (see `consumed`)."""
        if not isinstance(text, (bytes, bytearray)):
            raise TypeError("text must be bytes")
        if max_blocks is None:
            max_blocks = len(text) // self._min_card_line() + 1
        return self.detect_card_ptr(bytes(text), len(text), final, max_blocks)

    def _min_card_line(self):
        """Shortest possible data line: '<t> <i> <payload>\\n' with one-character time and index."""
        return ((2 * self.block_len + 2) // 3) * 4 + 5

    def detect_card_ptr(self, ptr, length, final=True, max_blocks=None):
        """Same as detect_card for text at a raw address (e.g. a PinnedBuffer: full-speed H2D copies)."""
        if max_blocks is None:
            max_blocks = length // self._min_card_line() + 1
        ts = np.zeros(max_blocks, dtype=np.float64)
        idx = np.zeros(max_blocks, dtype=np.int64)
        out = np.zeros((max_blocks, self.n_templates), dtype=RECORD_DTYPE)
        nblk, consumed = c_int64(0), c_int64(0)
        if not isinstance(ptr, (bytes, bytearray)):
            ptr = ctypes.cast(ptr, c_char_p)
        self._check(self._lib.thr_detect_card(self._h, ptr, length, 1 if final else 0, max_blocks,
                                              ts.ctypes.data, idx.ctypes.data, out.ctypes.data,
                                              byref(nblk), byref(consumed)))
        n = nblk.value
        return ts[:n], idx[:n], out[:n], consumed.value

    # -- device-buffer API (pointers are integers, e.g. torch.Tensor.data_ptr())
    def detect_device(self, d_raw, d_block_idx, n_blocks, d_out):
        self._check(self._lib.thr_detect_batch_device(self._h, d_raw, d_block_idx, int(n_blocks), d_out))

    def detect_device_c64(self, d_iq, d_block_idx, n_blocks, d_out):
        self._check(self._lib.thr_detect_batch_device_c64(self._h, d_iq, d_block_idx, int(n_blocks), d_out))

    def set_stream(self, cuda_stream):
        self._check(self._lib.thr_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._check(self._lib.thr_synchronize(self._h))

    def timer_start(self):
        self._check(self._lib.thr_timer_start(self._h))

    def timer_stop(self):
        ms = c_float()
        self._check(self._lib.thr_timer_stop(self._h, byref(ms)))
        return ms.value


class NativeGroup(NativeDetector):
    """Several GPUs behind one handle (thr_group_*): every batch is cut into contiguous stripes, one per device, and the
    records come back in input order, byte-identical to one GPU's.  Same host-buffer methods as NativeDetector
    (detect_raw / detect_stream / detect_card[_ptr]); complex64 blocks, single-block debug outputs and the device-buffer
    entry points go to the first member."""

    def __init__(self, devices, block_len, history_len, templates, carrier_len, carrier_window, carrier_thresh,
                 corr_thresh, max_batch=4096, fastdet=False):
        self._lib = load_library()
        self._g = c_void_p()
        self._h = c_void_p()
        self._owned = False
        self.devices = [int(d) for d in devices]
        cfg, _tpl = _make_config(self, block_len, history_len, templates, carrier_len, carrier_window, carrier_thresh,
                                 corr_thresh, self.devices[0], max_batch, False, fastdet, False)
        dev = (c_int32 * len(self.devices))(*self.devices)
        rc = self._lib.thr_group_create(byref(cfg), dev, len(self.devices), byref(self._g))
        if rc != THR_OK:
            msg = self._lib.thr_group_last_error(None).decode()
            self._g = c_void_p()
            _raise_create_error("thr_group_create", rc, msg)
        self._h = c_void_p(self._lib.thr_group_member(self._g, 0))      # borrowed: first member

    def _gcheck(self, rc):
        if rc != THR_OK:
            raise NativeError("thrifty_b200 group call failed (%d): %s" % (rc, self._lib.thr_group_last_error(self._g).decode()))

    def close(self):
        if self._g:
            self._lib.thr_group_destroy(self._g)
            self._g = c_void_p()
            self._h = c_void_p()

    def numa_nodes(self):
        out = []
        for i in range(len(self.devices)):
            bound = c_int32(0)
            out.append((self._lib.thr_group_numa_node(self._g, i, byref(bound)), bool(bound.value)))
        return out

    def info(self):
        info = NativeDetector.info(self)
        info["devices"] = list(self.devices)
        launches = 0
        for i in range(len(self.devices)):
            m = ThrInfo()
            self._check(self._lib.thr_get_info(c_void_p(self._lib.thr_group_member(self._g, i)), byref(m)))
            launches += m.launches
        info["launches"] = launches
        return info

    def detect_raw(self, raw, block_idx=None):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        if raw.ndim != 2 or raw.shape[1] != 2 * self.block_len:
            raise ValueError("raw must have shape [B, %d]" % (2 * self.block_len))
        nblk = raw.shape[0]
        out = np.zeros((nblk, self.n_templates), dtype=RECORD_DTYPE)
        idx_ptr = None
        if block_idx is not None:
            idx = np.ascontiguousarray(block_idx, dtype=np.int64)
            assert idx.shape == (nblk,)
            idx_ptr = idx.ctypes.data
        self._gcheck(self._lib.thr_group_detect_batch(self._g, raw.ctypes.data, idx_ptr, nblk, out.ctypes.data))
        return out

    def detect_raw_ptr(self, ptr, n_blocks, idx_ptr, out_ptr):
        """Raw addresses (page-locked buffers): no NumPy copies."""
        self._gcheck(self._lib.thr_group_detect_batch(self._g, ptr, idx_ptr, int(n_blocks), out_ptr))

    def detect_stream(self, stream, first_block):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        new = 2 * (self.block_len - self.history_len)
        nblk = (len(stream) - 2 * self.block_len) // new + 1 if len(stream) >= 2 * self.block_len else 0
        out = np.zeros((max(nblk, 0), self.n_templates), dtype=RECORD_DTYPE)
        got = c_int64(0)
        self._gcheck(self._lib.thr_group_detect_stream(self._g, stream.ctypes.data, len(stream), int(first_block),
                                                       out.ctypes.data, byref(got)))
        assert got.value == len(out)
        return out

    def detect_card_ptr(self, ptr, length, final=True, max_blocks=None):
        line_len = ((2 * self.block_len + 2) // 3) * 4 + 5
        if max_blocks is None:
            max_blocks = length // line_len + 1
        ts = np.zeros(max_blocks, dtype=np.float64)
        idx = np.zeros(max_blocks, dtype=np.int64)
        out = np.zeros((max_blocks, self.n_templates), dtype=RECORD_DTYPE)
        nblk, consumed = c_int64(0), c_int64(0)
        if not isinstance(ptr, (bytes, bytearray)):
            ptr = ctypes.cast(ptr, c_char_p)
        self._gcheck(self._lib.thr_group_detect_card(self._g, ptr, length, 1 if final else 0, max_blocks,
                                                     ts.ctypes.data, idx.ctypes.data, out.ctypes.data,
                                                     byref(nblk), byref(consumed)))
        n = nblk.value
        return ts[:n], idx[:n], out[:n], consumed.value
