"""Row-sharded item tables (SURVEY 8e / BASELINE config 5) on the GPU: ShardedModel must train to the same
weights as the replicated Model on the same global batch -- one rank (every collective degenerate, all kernels
exercised) and two ranks over NCCL, both partitions -- and score identically."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=7):
    from oracle import tlsan_oracle as O
    from tests.util import load_digital_music
    dm = load_digital_music()
    cfg = O.default_config(*dm.counts)
    params = O.randomize_params(O.init_params(cfg, seed=1234), seed=seed)
    return O, dm, cfg, params


def _max_rel(sd, rsd):
    return max(float(np.max(np.abs(np.asarray(sd[k]) - np.asarray(rsd[k])))) /
               (float(np.max(np.abs(np.asarray(rsd[k])))) + 1e-12) for k in rsd)


@pytest.mark.parametrize("partition", ["mod", "block"])
def test_one_rank_sharded_equals_replicated(partition):
    from tests.util import model_from_params
    from tlsan_b200.sharded import ShardedModel
    O, dm, cfg, params = _setup()
    ref = model_from_params(params, dm.icl, cfg)
    sm = ShardedModel(cfg, dm.icl, partition=partition)
    sm.load_full_state({k: np.asarray(v) for k, v in params.items()})
    tb = O.collate_test(dm.test_set[:256], 10)
    lg_ref, _ = ref.score_staged(ref.stage_batch(tb, is_test=True), 2)
    lg = sm.score_staged(sm.stage_batch(tb, is_test=True), 2)
    assert torch.equal(lg, lg_ref)                      # same kernels on the same rows: bit-identical logits
    for step in range(3):
        batch = O.collate_train(dm.train_set[step * 300:(step + 1) * 300], 10)
        l_ref = ref.train(None, batch, 1.0)
        l = sm.train(None, batch, 1.0)
        assert abs(l - l_ref) <= 1e-6 * abs(l_ref), (step, l, l_ref)
    err = _max_rel(sm.gather_full_state(), {k: v.numpy() for k, v in ref.state_dict().items()})
    assert err < 1e-6, err
    assert sm.last_unique <= sm.cap and sm.global_step == 3


def _worker(rank, world, port, partition, q):
    import torch.distributed as dist
    from tests.util import model_from_params
    from tlsan_b200.parallel import shard_rows
    from tlsan_b200.sharded import ShardedModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    O, dm, cfg, params = _setup()
    sm = ShardedModel(cfg, dm.icl, process_group=dist.group.WORLD, partition=partition)
    sm.load_full_state({k: np.asarray(v) for k, v in params.items()})
    losses = []
    for step in range(3):
        batch = O.collate_train(dm.train_set[step * 301:(step + 1) * 301], 10)      # odd size: uneven shards
        local, _ = shard_rows(batch, rank, world)
        stats = sm.train_staged(sm.stage_batch(local), 1.0, global_batch=len(batch[0]))
        losses.append(float(stats[0].item()))
    assert int(sm._bad.item()) == 0
    tb = O.collate_test(dm.test_set[:200], 10)
    local_t, _ = shard_rows(tb, rank, world)
    lg = sm.score_staged(sm.stage_batch(local_t, is_test=True), 2).cpu().numpy()
    sd = sm.gather_full_state()
    if rank == 0:
        ref = model_from_params(params, dm.icl, cfg)
        ref_losses = [ref.train(None, O.collate_train(dm.train_set[s * 301:(s + 1) * 301], 10), 1.0) for s in range(3)]
        err = _max_rel(sd, {k: v.numpy() for k, v in ref.state_dict().items()})
        lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
        lg_ref = ref.score_staged(ref.stage_batch(local_t, is_test=True), 2)[0].cpu().numpy()
        q.put((err, lerr, float(np.max(np.abs(lg - lg_ref)) / np.max(np.abs(lg_ref)))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["mod", "block"])
def test_two_rank_sharded_matches_replicated(partition):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, partition, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, lerr, serr = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert err < 1e-5 and lerr < 1e-5 and serr < 1e-5, (err, lerr, serr)
