"""Row-sharded item tables, host-side logic on CPU (world_size 2, gloo): the id routing of
tlsan_b200/sharded.py -- partition maps, the all-to-all of distinct ids to their owners and the return trip
of rows -- with a plain tensor standing in for the table (the CUDA pack / unpack kernels are covered by
tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tlsan_b200.sharded import exchange, local_of, owner_of, route_ids, shard_ids


@pytest.mark.parametrize("partition", ["mod", "block"])
@pytest.mark.parametrize("NI,world", [(1, 1), (10, 3), (1583, 8), (4096, 8), (7, 8)])
def test_partition_maps_are_a_bijection(partition, NI, world):
    ids = torch.arange(NI)
    own, loc = owner_of(ids, world, NI, partition), local_of(ids, world, NI, partition)
    assert int(own.max()) < world
    seen = np.zeros(NI, bool)
    for r in range(world):
        mine = shard_ids(r, world, NI, partition)
        assert np.array_equal(ids[own == r].numpy(), mine)                       # shard order = ascending ids
        assert np.array_equal(loc[own == r].numpy(), np.arange(mine.size))       # local index = position in shard
        seen[mine] = True
    assert seen.all()


def _worker(rank, world, port, partition, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    NI = 1000
    table = torch.arange(NI, dtype=torch.float32)[:, None] * torch.tensor([1.0, 10.0])   # row id -> (id, 10 id)
    mine = torch.from_numpy(shard_ids(rank, world, NI, partition))
    shard = table[mine]
    rng = np.random.default_rng(100 + rank)
    uniq = torch.from_numpy(np.unique(rng.integers(0, NI, 300)))
    order, recv_ids, in_splits, out_splits = route_ids(uniq, world, NI, partition, dist.group.WORLD)
    ok = bool((owner_of(recv_ids, world, NI, partition) == rank).all())          # only ids this rank owns arrive
    rows_out = shard[local_of(recv_ids, world, NI, partition)]
    rows_in = exchange(rows_out, out_splits, in_splits, dist.group.WORLD, world)
    compact = torch.empty(uniq.numel(), 2)
    compact[order] = rows_in                                                     # what tlsan_shard_unpack_rows does
    ok = ok and torch.equal(compact, table[uniq])
    # gradients travel the reverse way with the same splits and land on the owner's local rows
    grads_in = exchange(compact[order], in_splits, out_splits, dist.group.WORLD, world)
    ok = ok and torch.equal(grads_in, shard[local_of(recv_ids, world, NI, partition)])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["mod", "block"])
def test_rows_reach_requesters_and_gradients_reach_owners(partition):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, partition, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
