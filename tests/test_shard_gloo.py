"""Host-side logic of the multi-GPU sharding, exercised on CPU with gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from paresis_b200 import shard
    # a detector bin of 5 energies spread over 2 ranks; per-rank partial images are rank-dependent
    indices = [3, 4, 5, 6, 7]
    mine = shard.energies_of(indices, rank, world)
    partial = torch.zeros((2, 4, 6), dtype=torch.float32)
    for e in mine:
        partial += float(e)
    shard.reduce_images(partial, owner=0)
    means = shard.combine_means(torch.tensor([10.0 * e for e in mine], dtype=torch.float64), mine, 9)
    q.put((rank, mine, partial.numpy().copy(), means.numpy().copy(), shard.positions_of(7, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_energy_reduce_and_position_partition():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, mine, partial, means, positions = q.get(timeout=120)
        got[rank] = (mine, partial, means, positions)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == [3, 5, 7] and got[1][0] == [4, 6]            # round-robin, disjoint, complete
    assert np.all(got[0][1] == 3 + 4 + 5 + 6 + 7)                      # the owner holds the sum of every energy
    want = np.zeros(9); want[3:8] = [30, 40, 50, 60, 70]
    assert np.array_equal(got[0][2], want) and np.array_equal(got[1][2], want)
    assert got[0][3] == [0, 2, 4, 6] and got[1][3] == [1, 3, 5]       # positions: no overlap, rank 0 owns position 0


def test_single_process_paths_need_no_group():
    from paresis_b200 import shard
    t = torch.ones((2, 3, 3))
    assert shard.reduce_images(t) is t
    m = shard.combine_means(torch.tensor([1.0, 2.0], dtype=torch.float64), [1, 4], 6)
    assert m.tolist() == [0, 1, 0, 0, 2, 0]
    assert shard.positions_of(5, 0, 1) == [0, 1, 2, 3, 4]
