# -*- coding: utf-8 -*-
"""N>1 host logic on CPU: world_size-2 gloo process group exercising the exchange helpers of george_b200.parallel
(shard ranges, padded all-gather, partial log-det reduction).  The arithmetic itself needs GPUs (tests/test_gpu_*.py
and bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, min_size, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from george_b200.parallel import allgather_padded, shard_ranges
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ranges = shard_ranges(n, world, min_size)
    start, size = ranges[rank]
    rows_pad = max(sz for _, sz in ranges)
    cols = 5
    # a "panel": element (row i, col c) = i + 1000*c ; each rank owns its rows
    full = (np.arange(n)[None, :] + 1000.0 * np.arange(cols)[:, None])
    local = torch.from_numpy(full[:, start:start + size].copy())
    out = allgather_padded(local, rows_pad)
    rebuilt = np.zeros_like(full)
    for s, (st, sz) in enumerate(ranges):
        rebuilt[:, st:st + sz] = out[s, :, :sz].numpy()
    ok = np.array_equal(rebuilt, full)
    part = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(part)
    ok = ok and part.item() == sum(range(1, world + 1))
    np.save(os.path.join(tmp, "ok{0}.npy".format(rank)), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,min_size", [(1001, 100), (4096, 256)])
def test_gloo_world2_exchange(tmp_path, n, min_size):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + n % 7
    mp.spawn(_worker, args=(2, port, n, min_size, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(os.path.join(str(tmp_path), "ok{0}.npy".format(r)))[0]


def test_shard_ranges():
    from george_b200.parallel import shard_ranges
    assert shard_ranges(1000, 1, 100) == [(0, 1000)]
    assert shard_ranges(1001, 2, 100) == [(0, 500), (500, 501)]
    assert shard_ranges(1001, 4, 100) == [(0, 250), (250, 250), (500, 250), (750, 251)]
    assert shard_ranges(300, 4, 100) is None
    with pytest.raises(ValueError):
        shard_ranges(100, 3, 10)
