"""Data-parallel train step on 2 GPUs: every rank trains on its contiguous row block, the gradients are summed
and the update applied either by the fused NVLink peer-memory exchange (tlsan_dp_exchange, "p2p") or by NCCL
all-reduce + tlsan_apply_flat ("nccl"); the weights must equal (up to fp32
reduction order) those of one GPU training on the whole batch, and be identical across ranks."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist
    from oracle import tlsan_oracle as O
    from tests.util import load_digital_music, model_from_params
    from tlsan_b200.parallel import shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dm = load_digital_music()
    cfg = O.default_config(*dm.counts)
    params = O.randomize_params(O.init_params(cfg, seed=1234), seed=7)
    model = model_from_params(params, dm.icl, cfg, process_group=dist.group.WORLD, dp_mode=mode)
    losses = []
    for step in range(3):
        batch = O.collate_train(dm.train_set[step * 301:(step + 1) * 301], 10)      # odd size: uneven shards
        local, _ = shard_rows(batch, rank, world)
        db = model.stage_batch(local)
        stats = model.train_staged(db, 1.0, global_batch=len(batch[0]))
        losses.append(float(stats[0].item()))
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    if rank == 0:
        ref = model_from_params(params, dm.icl, cfg)
        ref_losses = []
        for step in range(3):
            ref_losses.append(ref.train(None, O.collate_train(dm.train_set[step * 301:(step + 1) * 301], 10), 1.0))
        rsd = {k: v.numpy() for k, v in ref.state_dict().items()}
        err = max(float(np.max(np.abs(sd[k] - rsd[k]))) / (float(np.max(np.abs(rsd[k]))) + 1e-12) for k in sd)
        lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
        q.put(("cmp", err, lerr))
    flat = torch.cat([torch.as_tensor(v).reshape(-1) for v in sd.values()]).cuda()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put(("same", all(torch.equal(gathered[0], g) for g in gathered)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_two_gpu_data_parallel_matches_single_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict()
    for _ in range(2):
        item = q.get(timeout=240)
        out[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    err, lerr = out["cmp"]
    assert err < 1e-5 and lerr < 1e-5, (err, lerr)
    assert out["same"][0] is True
