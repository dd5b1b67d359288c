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


def test_one_rank_sharded_prec_recall_equal_replicated():
    """Sharded full-catalogue P@k / R@k (tlsan_label_rank_shard: label rows gathered, per-shard counts summed) against
    the replicated Model's tcgen05 rank kernel on the same rows: identical ranks, hence identical metrics."""
    from tests.util import model_from_params
    from tlsan_b200.sharded import ShardedModel
    O, dm, cfg, params = _setup()
    ref = model_from_params(params, dm.icl, cfg)
    for partition in ("mod", "block"):
        sm = ShardedModel(cfg, dm.icl, partition=partition)
        sm.load_full_state({k: np.asarray(v) for k, v in params.items()})
        ref.reset_metrics()
        for lo in (0, 128):
            tb = O.collate_test(dm.test_set[lo:lo + 128], 10)
            assert np.array_equal(sm._label_ranks(tb), ref._update_topk(tb))
            assert np.allclose(sm.eval_prec(None, tb), ref.eval_prec(None, tb), atol=0, rtol=0)
            assert np.allclose(sm.eval_recall(None, tb), ref.eval_recall(None, tb), atol=0, rtol=0)


def _worker_rank(rank, world, port, q):
    import torch.distributed as dist
    from tests.util import model_from_params
    from tlsan_b200.parallel import shard_rows
    from tlsan_b200.sharded import ShardedModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    O, dm, cfg, params = _setup()
    sm = ShardedModel(cfg, dm.icl, process_group=dist.group.WORLD, partition="mod")
    sm.load_full_state({k: np.asarray(v) for k, v in params.items()})
    tb = O.collate_test(dm.test_set[:201], 10)                       # odd: uneven row blocks
    local_t, _ = shard_rows(tb, rank, world)
    ranks = sm._label_ranks(local_t)
    ref = model_from_params(params, dm.icl, cfg)
    want = ref._update_topk(local_t)
    q.put((rank, bool(np.array_equal(ranks, want))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_label_ranks_equal_replicated():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == {0: True, 1: True}


def test_sharded_step_with_item_ids_above_2_pow_24():
    """BASELINE config 5 territory: item ids that do not fit fp32's 24-bit integer range.  A catalogue of 2^24 + 50 000
    items, a batch that only uses ids above 2^24; the touched rows must follow the fp64 oracle run on the touched rows
    alone (renumbered), an untouched row must see pure L2 decay, and loss = bce + reg * l2 over the WHOLE catalogue."""
    from oracle import tlsan_oracle as O
    from tests.util import synth_batch
    from tlsan_b200.sharded import ShardedModel
    rng = np.random.default_rng(24)
    NI, NU, NC, L, S, B = (1 << 24) + 50000, 40, 7, 10, 3, 96
    cfg = O.default_config(NU, NI, NC, Ls=L)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    sm = ShardedModel(cfg, icl, partition="mod")
    used = np.sort(rng.choice(np.arange(1 << 24, NI), 300, replace=False))     # the batch's item universe
    small = synth_batch(rng, B, L, S, len(used), NU, NC)                          # ids in [0, 300)
    big = list(small)
    for k in (1, 3, 4):
        big[k] = used[np.asarray(small[k])]
    col = np.arange(L)[None, :]
    big[3] = np.where(col < np.asarray(small[6])[:, None], big[3], 0)             # padding stays id 0
    big[4] = np.where(np.arange(S)[None, :] < np.asarray(small[7])[:, None], big[4], 0)
    small = list(small)
    small[3] = np.where(col < np.asarray(small[6])[:, None], np.asarray(small[3]) + 1, 0)   # small id 0 = global id 0
    small[4] = np.where(np.arange(S)[None, :] < np.asarray(small[7])[:, None], np.asarray(small[4]) + 1, 0)
    small[1] = np.asarray(small[1]) + 1
    rows = np.concatenate([[0], used])
    # oracle parameters = the touched rows of the sharded model's own weights
    p = O.randomize_params(O.init_params(O.default_config(NU, len(rows), NC, Ls=L), seed=1234), seed=5)
    dev_rows = torch.from_numpy(rows).cuda()
    p["item_emb"] = sm.item_emb_shard[dev_rows].cpu().numpy()
    p["item_b"] = sm.item_b_shard[dev_rows].cpu().numpy()
    full = {k: np.asarray(v) for k, v in p.items()}
    sm.cate_emb.copy_(torch.as_tensor(full["cate_emb"])); sm.user_emb.copy_(torch.as_tensor(full["user_emb"]))
    sm.usert_emb.copy_(torch.as_tensor(full["usert_emb"]))
    from tlsan_b200.model import DENSE_LAYOUT
    dense = sm.dense.cpu().clone()
    for name, (off, shape) in DENSE_LAYOUT.items():
        n = int(np.prod(shape)) if shape else 1
        dense[off:off + n] = torch.as_tensor(full[name]).reshape(-1)
    sm.dense.copy_(dense)
    l2_all = 0.5 * (float(sm.item_emb_shard.double().pow(2).sum()) + sum(float(np.sum(full[k].astype(np.float64) ** 2))
                                                                         for k in ("user_emb", "cate_emb", "usert_emb")))
    untouched = int((1 << 24) + 7 if (1 << 24) + 7 not in set(used.tolist()) else (1 << 24) + 8)
    before_untouched = sm.item_emb_shard[untouched].cpu().numpy().copy()
    ref = O.train_step(p, icl[rows], tuple(small), 0.5, O.default_config(NU, len(rows), NC, Ls=L), dtype=torch.float64)
    loss = sm.train(None, tuple(big), 0.5)
    stats = sm._stats.cpu().numpy()
    assert stats[3] == 1.0
    assert abs(stats[1] - ref["bce"]) <= 1e-4 * abs(ref["bce"])
    assert abs(loss - (ref["bce"] + cfg["regulation_rate"] * l2_all)) <= 1e-4 * abs(loss)
    got = sm.item_emb_shard[dev_rows].cpu().numpy()
    step = np.abs(p["item_emb"] - ref["new_params"]["item_emb"])
    assert np.max(np.abs(got - ref["new_params"]["item_emb"])) <= 1e-4 * np.max(step) + 2e-7
    got_b = sm.item_b_shard[dev_rows].cpu().numpy()
    assert np.max(np.abs(got_b - ref["new_params"]["item_b"])) <= 1e-4 * np.max(np.abs(p["item_b"] - ref["new_params"]["item_b"])) + 2e-7
    exp = before_untouched * np.float32(1.0 - 0.5 * cfg["regulation_rate"])
    assert np.max(np.abs(sm.item_emb_shard[untouched].cpu().numpy() - exp)) < 1e-7
