"""world_size=2 gloo test of the N>1 host path: contiguous stripes + gather of TOAD records."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_blocks, n_tpl, out_dir):
    import torch.distributed as dist
    from thrifty_b200 import stripe
    from thrifty_b200._native import RECORD_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = stripe.stripe_bounds(n_blocks, world, rank)
    local = np.zeros((hi - lo, n_tpl), dtype=RECORD_DTYPE)
    for t in range(n_tpl):
        local["block_idx"][:, t] = np.arange(lo, hi)
        local["soa"][:, t] = np.arange(lo, hi) * 11464.0 + t
        local["template_idx"][:, t] = t
        local["flags"][:, t] = (np.arange(lo, hi) % 3).astype(np.uint32)
    full = stripe.gather_records(local, n_blocks)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), full)
    dist.destroy_process_group()


def _run(n_blocks, n_tpl, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_blocks, n_tpl, str(tmp_path)), nprocs=world, join=True)
    results = [np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(world)]
    for full in results:
        assert full.shape == (n_blocks, n_tpl)
        for t in range(n_tpl):
            np.testing.assert_array_equal(full["block_idx"][:, t], np.arange(n_blocks))
            np.testing.assert_array_equal(full["soa"][:, t], np.arange(n_blocks) * 11464.0 + t)
            np.testing.assert_array_equal(full["flags"][:, t], np.arange(n_blocks) % 3)
    np.testing.assert_array_equal(results[0], results[1])


def test_gather_even(tmp_path):
    _run(64, 1, tmp_path)


def test_gather_ragged_multi_template(tmp_path):
    _run(37, 4, tmp_path)
