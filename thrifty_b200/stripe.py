"""Striping of a block stream across ranks (one process per GPU) and gathering of records.

`.card` blocks are self-contained (each carries its own history, fastcard/card_reader.c:69-75)
and the detector keeps no cross-block state (thrifty/detect.py:40-58), so the stream shards
into contiguous stripes with no data-path exchange; the only collective is the gather of the
64-byte TOAD records (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def stripe_bounds(n_blocks, world, rank):
    """Contiguous stripe [lo, hi) of rank `rank`: ceil(n/world) blocks per rank, so that the
    rank-order concatenation of stripes is the input order."""
    per = (n_blocks + world - 1) // world
    lo = min(n_blocks, rank * per)
    hi = min(n_blocks, lo + per)
    return lo, hi


def gather_records(local_records, n_blocks, group=None):
    """All-gather per-rank record arrays (numpy, RECORD_DTYPE [n_local, T]) into the full
    [n_blocks, T] array on every rank.  Stripes are padded to equal length for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (n_blocks + world - 1) // world
    n_tpl = local_records.shape[1]
    item = local_records.dtype.itemsize
    buf = np.zeros((per, n_tpl), dtype=local_records.dtype)
    buf[:len(local_records)] = local_records
    send = torch.from_numpy(buf.view(np.uint8).reshape(-1).copy())
    backend = dist.get_backend(group)
    if backend == "nccl":
        send = send.cuda()
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    full = np.frombuffer(recv.cpu().numpy().tobytes(), dtype=local_records.dtype)
    full = full.reshape(world * per, n_tpl)
    assert item * n_tpl * per == send.numel()
    return full[:n_blocks]
