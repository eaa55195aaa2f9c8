"""Detection record types and `.toad` text (de)serialisation.

Same field names, column order and text format as thrifty/toads_data.py:8-90 so that
`.toad` files written here feed `thrifty identify / match / tdoa / pos` unchanged.
"""
from collections import namedtuple

import numpy as np

CarrierSyncInfo = namedtuple("CarrierSyncInfo", ["bin", "offset", "energy", "noise"])
CorrDetectionInfo = namedtuple("CorrDetectionInfo", ["sample", "offset", "energy", "noise"])


class DetectionResult(object):
    """One block's detection outcome (thrifty/toads_data.py:22-45)."""

    __slots__ = ("timestamp", "block", "soa", "carrier_info", "corr_info", "rxid", "txid")

    def __init__(self, timestamp, block, soa, carrier_info, corr_info, rxid=None, txid=None):
        self.timestamp = timestamp
        self.block = block
        self.soa = soa
        self.carrier_info = carrier_info
        self.corr_info = corr_info
        self.rxid = rxid
        self.txid = txid

    def serialize(self):
        """`.toad(s)` line: [rxid] [txid] t block soa  corr(4)  carrier(4)  (toads_data.py:47-61)."""
        corr, carr = self.corr_info, self.carrier_info
        fields = ["{:.6f}".format(self.timestamp), str(self.block), "{:.8f}".format(self.soa),
                  str(corr.sample), str(corr.offset), str(corr.energy), str(corr.noise),
                  str(carr.bin), str(carr.offset), str(carr.energy), str(carr.noise)]
        if self.txid is not None:
            fields.insert(0, str(self.txid))
        if self.rxid is not None:
            fields.insert(0, str(self.rxid))
        return " ".join(fields)

    @classmethod
    def deserialize(cls, string, with_rxid=False, with_txid=False):
        """Inverse of serialize (toads_data.py:63-90); None for malformed lines."""
        fields = string.split()
        if len(fields) < 11 + with_rxid + with_txid:
            return None
        rxid = int(fields.pop(0)) if with_rxid else None
        txid = int(fields.pop(0)) if with_txid else None
        t, b, s, ps, po, pe, pn, cb, co, ce, cn = [float(v) for v in fields[:11]]
        return cls(timestamp=t, block=int(b), soa=float(s),
                   carrier_info=CarrierSyncInfo(bin=int(cb), offset=co, energy=ce, noise=cn),
                   corr_info=CorrDetectionInfo(sample=int(ps), offset=po, energy=pe, noise=pn),
                   rxid=rxid, txid=txid)


def _load(stream, with_rxid, with_txid):
    if isinstance(stream, str):
        stream = open(stream, "r")
    out = []
    for i, line in enumerate(stream):
        if len(line) == 0 or line[0] == "#":
            continue
        det = DetectionResult.deserialize(line, with_rxid=with_rxid, with_txid=with_txid)
        if det is None:
            print("WARNING: skipped line #{}: line's formatting is invalid".format(i + 1))
            continue
        out.append(det)
    return out


def load_toad(stream):
    """Single receiver `.toad` (toads_data.py:113-115)."""
    return _load(stream, True, False)


def load_toads(stream):
    """Multi-receiver `.toads` with txid column (toads_data.py:118-120)."""
    return _load(stream, True, True)


def toads_array(detections, with_ids=True):
    """Structured array view (toads_data.py:123-143)."""
    rows = [(i, d.rxid if with_ids else -1, d.txid if with_ids else -1, d.timestamp, d.block, d.soa,
             d.corr_info.sample, d.corr_info.offset, d.corr_info.energy, d.corr_info.noise,
             d.carrier_info.bin, d.carrier_info.offset, d.carrier_info.energy, d.carrier_info.noise)
            for i, d in enumerate(detections)]
    return np.array(rows, dtype=[
        ("idx", "i4"), ("rxid", "i4"), ("txid", "i4"), ("timestamp", "f8"), ("block", "i4"),
        ("soa", "f8"), ("sample", "i4"), ("offset", "f8"), ("energy", "f8"), ("noise", "f8"),
        ("carrier_bin", "i4"), ("carrier_offset", "f8"), ("carrier_energy", "f8"),
        ("carrier_noise", "f8")])
