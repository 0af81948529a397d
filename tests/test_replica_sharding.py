"""Host-side multi-rank logic on CPU: world_size-2 gloo (the N>1 path of bench.py / ReplicaSharder)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ann_force_b200.replica import ReplicaSharder, shard_bounds


def test_shard_bounds_cover_everything_once():
    for R in [1, 2, 7, 8, 1024, 1025, 8192]:
        for G in [1, 2, 3, 4, 8]:
            blocks = [shard_bounds(R, G, r) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == R
            for (a0, a1), (b0, b1) in zip(blocks[:-1], blocks[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, R, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sharder = ReplicaSharder(R)
        energies = torch.arange(R, dtype=torch.float64) * 1.5 + 0.25          # the "truth", same on every rank
        gathered = sharder.gather_energies(sharder.local_slice(energies).clone())
        first = gathered.numpy().copy()
        # a second exchange re-uses the pre-allocated buffers and must see the new values
        gathered = sharder.gather_energies(sharder.local_slice(energies + 1.0).clone())
        assert np.array_equal(gathered.numpy(), first + 1.0)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), first)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("R", [8, 9])
def test_energy_allgather_two_ranks_gloo(tmp_path, R):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, R, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(R) * 1.5 + 0.25
    for rank in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / ("rank%d.npy" % rank)), want)
