"""Detect presence of a carrier and synchronize to the carrier's frequency -- B200 drop-in for
thrifty.carrier_sync's DefaultSynchronizer (thrifty/carrier_sync.py:82-118).

    sync = DefaultSynchronizer(thresh_coeffs, window, block_len, carrier_len)
    shifted_fft, info = sync(block)          # carrier_sync.py:52-76: None + info when no carrier is found

`block` is a complex array of block_len samples (or raw uint8 I/Q of 2 * block_len bytes); `info` is a
toads_data.CarrierSyncInfo(bin, offset, energy, noise); `shifted_fft` is the spectrum of the block mixed down by
bin + offset (complex64, natural bin order, including the reference's block-constant phase exp(+j pi (bin + offset))).
The carrier decision, the float64 Dirichlet fit, the mix and FFT #2 run in the fused CUDA kernel, stopped at the stage
boundary (thr_sync_batch); `sync_many` does a whole batch in one launch.  The reference's building blocks
(Synchronizer with replaceable detector / interpolator / shifter callables) are not re-exposed: on the GPU the three
steps are one kernel.  There is no CPU fallback.
"""

from __future__ import print_function

import numpy as np

from thrifty_b200 import toads_data
from thrifty_b200._native import FLAG_CARRIER, NativeDetector


class DefaultSynchronizer(object):
    """Carrier detector (threshold on the windowed spectral peak, carrier_detect.py:61-154), Dirichlet-kernel
    sub-bin interpolator (carrier_sync.py:150-196) and time-domain frequency shifter (carrier_sync.py:222-238).

    Parameters as thrifty/carrier_sync.py:82-101: thresh_coeffs = (constant, snr, stddev), window = (start, stop)
    closed interval in signed bins or None, block_len, carrier_len.  Extras: device, batch (blocks per launch)."""

    def __init__(self, thresh_coeffs, window, block_len, carrier_len, device=0, batch=256):
        self.thresh_coeffs = thresh_coeffs
        self.window = window
        self.block_len = int(block_len)
        self.carrier_len = int(carrier_len)
        # the synchronizer has no template: a one-sample dummy keeps the handle's template checks happy
        self.native = NativeDetector(self.block_len, 0, np.ones(1), self.carrier_len, window, thresh_coeffs,
                                     (0., 0., 0.), device=device, max_batch=max(1, int(batch)))

    def sync_many(self, blocks):
        """List / array of blocks -> list of (shifted_fft or None, CarrierSyncInfo), one launch."""
        blocks = [np.asarray(b) for b in blocks]
        if not blocks:
            return []
        if all(b.dtype == np.uint8 for b in blocks):
            recs, sfft = self.native.sync_batch(raw=np.stack(blocks))
        else:
            for b in blocks:
                assert len(b) == self.block_len
            recs, sfft = self.native.sync_batch(iq=np.stack([np.asarray(b, dtype=np.complex64) for b in blocks]))
        out = []
        for i in range(len(blocks)):
            found = bool(recs["flags"][i] & FLAG_CARRIER)
            info = toads_data.CarrierSyncInfo(int(recs["carrier_bin"][i]), float(recs["carrier_offset"][i]) if found else 0,
                                              recs["carrier_energy"][i], recs["carrier_noise"][i])
            out.append((sfft[i] if found else None, info))
        return out

    def sync(self, signal):
        """Detect presence of carrier, estimate frequency, and compensate (carrier_sync.py:52-76)."""
        return self.sync_many([signal])[0]

    def __call__(self, signal):
        return self.sync(signal)

    def close(self):
        self.native.close()
