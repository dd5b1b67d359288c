"""Data-parallel decomposition on CPU: world_size 2, gloo.  Each rank runs the oracle on its
row block with the global-mean scaling, the flat gradients are all-reduced (sum), and the result
must equal the single-process full-batch gradients -- the identity the GPU path relies on when
it all-reduces the buffer of tlsan_step_grads."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tlsan_b200.parallel import row_block, shard_rows


def test_row_block_partitions():
    for n in (1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            blocks = [row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import tlsan_oracle as O
    from tests.util import synth_batch
    rng = np.random.default_rng(11)                       # same data on every rank
    NU, NI, NC, L, S, B = 30, 80, 5, 10, 4, 51
    cfg = O.default_config(NU, NI, NC, Ls=L, regulation_rate=0.0)
    params = O.randomize_params(O.init_params(cfg), seed=5)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    batch = synth_batch(rng, B, L, S, NI, NU, NC)
    local, (lo, hi) = shard_rows(batch, rank, world)
    assert local[4].shape[1] == batch[4].shape[1]
    r = O.train_step(params, icl, local, 1.0, cfg, dtype=torch.float64)
    w = (hi - lo) / B                                     # local mean -> share of the global mean
    flat = torch.cat([torch.as_tensor(v).reshape(-1) * w for v in r["grads"].values()] +
                     [torch.tensor([r["bce"] * w], dtype=torch.float64)])
    dist.all_reduce(flat)
    if rank == 0:
        full = O.train_step(params, icl, batch, 1.0, cfg, dtype=torch.float64)
        ref = torch.cat([torch.as_tensor(v).reshape(-1) for v in full["grads"].values()] +
                        [torch.tensor([full["bce"]], dtype=torch.float64)])
        q.put(float((flat - ref).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduced_shard_gradients_equal_full_batch():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12
