"""CPU oracle for the TLSAN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  ``tlsan_b200`` never does.

What it is: an op-for-op restatement, on torch-CPU tensors, of the TensorFlow-1.8 graph
built by the reference ``TLSAN/model.py`` (forward ``:56-137``, loss ``:164-172``,
gradients / clip / SGD ``:185-205``, attention helpers ``:316-483``, AUC ``:237-263``,
P@k / R@k ``:140-156,265-299``).  The arithmetic lives in the un-vendored third-party
dependency **tensorflow == 1.8.0** (``README.md:6``), which cannot be installed in this
image (python 3.12, no network); its published op semantics are restated here and each
function cites the reference call site it follows.

PARITY STATUS
  * model arithmetic: the reference has no tests, golden vectors or fixtures for this path
     and TF cannot execute here.  **Pinned to the reference's own graph
    code**: ``oracle/make_model_golden.py`` imports the UNMODIFIED ``TLSAN/model.py`` with
    ``tensorflow`` bound to ``oracle/tf1_shim.py`` (an eager torch restatement of the ~50 public
    TF-1.8 API entries the file calls) and executes build_model / attention_net / init_optimizer
    on fixed Digital-Music batches; this oracle reproduces the recorded loss, logits, every
    gradient and the updated weights to 1e-10 in float64
    (``tests/golden/model_ref_graph.npz``, ``tests/test_reference_graph.py``).  **Still
    unpinned**: TF's own kernels and TF-internal gradient plumbing (items 1-3 below) -- no
    TensorFlow output exists to compare with; external anchor README.md:35 (Digital-Music
    AUC 0.9753).
  * data path (batch layout, time buckets): **pinned** -- ``tests/golden/`` holds
    outputs of the *unmodified* reference ``TLSAN/input.py`` and ``TLSAN/build_dataset.py``
    executed in the build container by ``oracle/make_golden.py``.

TF-internal semantics that are decided here:
  1. clip norm: ``clip_mode='tf'`` (default) squares the *un-aggregated* IndexedSlices
     values (one slice per gather occurrence + one dense reg*W slice per table), which is
     what ``tf.clip_by_global_norm`` does to the output of ``tf.gradients`` in TF 1.8;
     ``clip_mode='aggregated'`` uses the norm of the summed dense gradient.
  2. softmax = exp(x - max) / sum over the sequence axis; additive -1e30 mask.
  3. top-k ties resolve to the lower index.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

VERY_BIG_NUMBER = 1e30                      # model.py:9
VERY_NEGATIVE_NUMBER = -VERY_BIG_NUMBER     # model.py:10

_L = "all/long_term/num_blocks0_0/long_term_layer/feature_wise_attention1/"
_S = "all/short_term/num_blocks1_0/short_term_layer/feature_wise_attention2/"
_D = "all/long_term/num_blocks0_0/dense/"

# trainable variables in creation order (model.py:58-81, 443-454, 347)
PARAM_NAMES = [
    "gamma_parameter", "item_emb", "item_b", "user_emb", "usert_emb", "cate_emb",
    _L + "bn_dense_map1/linear_map/W", _L + "bn_dense_map1/linear_map/bias",
    _L + "bn_dense_map2/linear_map/W", _L + "bn_dense_map2/linear_map/bias",
    _D + "kernel", _D + "bias",
    _S + "bn_dense_map1/linear_map/W", _S + "bn_dense_map1/linear_map/bias",
    _S + "bn_dense_map2/linear_map/W", _S + "bn_dense_map2/linear_map/bias",
]
TABLE_NAMES = ("user_emb", "item_emb", "cate_emb", "usert_emb")   # l2_loss terms, model.py:164-169


def default_config(user_count, item_count, cate_count, **over):
    """Flag defaults of TLSAN/train.py:26-49 plus the dataset counts (train.py:149-154)."""
    cfg = OrderedDict(
        hidden_units=64, num_blocks=1, num_heads=8, Ls=10, dropout=0.0,
        regulation_rate=0.00005, itemid_embedding_size=32, userid_embedding_size=32,
        cateid_embedding_size=32, optimizer="sgd", learning_rate=1.0,
        max_gradient_norm=5.0, train_batch_size=32, test_batch_size=128,
        model_dir="save_path", user_count=int(user_count), item_count=int(item_count),
        cate_count=int(cate_count))
    cfg.update(over)
    return cfg


def _glorot(rng, shape):
    """tf.get_variable default initializer = glorot_uniform (distribution parity only)."""
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_params(config, seed=1234):
    """Variables of model.py:56-81 + attention weights (:443-454, :347), as fp32 numpy."""
    rng = np.random.default_rng(seed)
    ni, nu, nc, L = config["item_count"], config["user_count"], config["cate_count"], config["Ls"]
    dh = config["hidden_units"] // config["num_heads"]
    hid = config["hidden_units"]
    p = OrderedDict()
    p["gamma_parameter"] = np.float32(1.0) * np.ones((), np.float32)          # :58-60
    p["item_emb"] = _glorot(rng, (ni, config["itemid_embedding_size"]))       # :62-64
    p["item_b"] = np.zeros((ni,), np.float32)                                 # :65-68
    p["user_emb"] = _glorot(rng, (nu, config["userid_embedding_size"]))       # :70-72
    p["usert_emb"] = np.full((nu, L), -1.0, np.float32)                       # :74-77
    p["cate_emb"] = _glorot(rng, (nc, config["cateid_embedding_size"]))       # :79-81
    for pre in (_L, None, _S):
        if pre is None:
            p[_D + "kernel"] = _glorot(rng, (hid, hid))                      # tf.layers.dense :347
            p[_D + "bias"] = np.zeros((hid,), np.float32)
            continue
        p[pre + "bn_dense_map1/linear_map/W"] = _glorot(rng, (dh, dh))       # :447
        p[pre + "bn_dense_map1/linear_map/bias"] = np.zeros((dh,), np.float32)
        p[pre + "bn_dense_map2/linear_map/W"] = _glorot(rng, (dh, dh))
        p[pre + "bn_dense_map2/linear_map/bias"] = np.zeros((dh,), np.float32)
    return OrderedDict((k, p[k]) for k in PARAM_NAMES)


def randomize_params(params, seed=7, scale=0.3):
    """Non-degenerate weights for parity tests (biases, gamma, usert_emb, item_b != init)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, v in params.items():
        if k == "gamma_parameter":
            out[k] = np.float32(0.8) * np.ones((), np.float32)
        elif k == "usert_emb":
            out[k] = (-1.0 + 0.5 * rng.standard_normal(v.shape)).astype(np.float32)
        elif k in ("item_emb", "user_emb", "cate_emb"):
            out[k] = (scale * rng.standard_normal(v.shape)).astype(np.float32)
        elif k == "item_b":
            out[k] = (0.1 * rng.standard_normal(v.shape)).astype(np.float32)
        elif v.ndim == 1:
            out[k] = (0.1 * rng.standard_normal(v.shape)).astype(np.float32)
        else:
            out[k] = (v + 0.2 * rng.standard_normal(v.shape) / np.sqrt(v.shape[0])).astype(np.float32)
    return out


# ----------------------------------------------------------------------------- graph helpers
def _linear(x, W, b):
    """model.py:443-454  out = matmul(x, W) + bias on the flattened [-1, d] tensor (:421-440)."""
    flat = x.reshape(-1, x.shape[-1])                                       # flatten :457-464
    out = flat @ W + b
    return out.reshape(*x.shape[:-1], W.shape[1])                          # reconstruct :467-477


def _bn_dense_layer(x, W, b, activation):
    """model.py:397-418 with enable_bn=False, keep_prob=1."""
    y = _linear(x, W, b)
    return torch.relu(y) if activation == "relu" else y


def _exp_mask_for_high_rank(val, mask):
    """model.py:480-483."""
    return val + (1 - mask.unsqueeze(-1).to(val.dtype)) * VERY_NEGATIVE_NUMBER


def feature_wise_attention(rep, rep_length, num_heads, W1, b1, W2, b2):
    """model.py:370-394."""
    rep = torch.cat(torch.split(rep, rep.shape[2] // num_heads, dim=2), dim=0)      # :374
    sl = rep.shape[1]
    mask = torch.arange(sl)[None, :] < rep_length[:, None]                          # :376
    mask = mask.repeat(num_heads, 1)                                                # :377
    map1 = _bn_dense_layer(rep, W1, b1, "relu")                                     # :380-381
    map2 = _bn_dense_layer(map1, W2, b2, "linear")                                  # :382-383
    map2_masked = _exp_mask_for_high_rank(map2, mask)                               # :384
    soft = torch.softmax(map2_masked, 1)                                            # :386
    out = torch.sum(soft * rep, 1)                                                  # :387
    out = torch.cat(torch.split(out, out.shape[0] // num_heads, dim=0), dim=1)      # :388
    return out, soft


def attention_net(enc, enc_new, sl, sl_new, P, num_heads):
    """model.py:316-366 (num_blocks == 1)."""
    enc, att0 = feature_wise_attention(
        enc, sl, num_heads,
        P[_L + "bn_dense_map1/linear_map/W"], P[_L + "bn_dense_map1/linear_map/bias"],
        P[_L + "bn_dense_map2/linear_map/W"], P[_L + "bn_dense_map2/linear_map/bias"])
    enc = (enc @ P[_D + "kernel"] + P[_D + "bias"]).unsqueeze(1)                    # :347
    enc = torch.cat([enc, enc_new], 1)                                              # :350
    enc_new, att1 = feature_wise_attention(
        enc, sl_new + 1, num_heads,                                                 # :356
        P[_S + "bn_dense_map1/linear_map/W"], P[_S + "bn_dense_map1/linear_map/bias"],
        P[_S + "bn_dense_map2/linear_map/W"], P[_S + "bn_dense_map2/linear_map/bias"])
    return enc_new, att0, att1


def _feed(batch, cand_index=1):
    """feed_dict mapping of model.py:210-222 / :239-261 (int64 -> int32 happens at the feed)."""
    u = torch.as_tensor(np.asarray(batch[0], dtype=np.int64))
    c = torch.as_tensor(np.asarray(batch[8], dtype=np.int64))
    i = torch.as_tensor(np.asarray(batch[cand_index], dtype=np.int64))
    hist_i = torch.as_tensor(np.asarray(batch[3], dtype=np.int64))
    hist_i_new = torch.as_tensor(np.asarray(batch[4], dtype=np.int64))
    hist_t = torch.as_tensor(np.asarray(batch[5], dtype=np.float32))
    sl = torch.as_tensor(np.asarray(batch[6], dtype=np.int64))
    sl_new = torch.as_tensor(np.asarray(batch[7], dtype=np.int64))
    return u, c, i, hist_i, hist_i_new, hist_t, sl, sl_new


def build_forward(P, icl, batch, cand_index=1, num_heads=8, hidden_units=64, keep=None):
    """model.py:83-137.  ``P`` maps variable name -> torch tensor.  Returns (logits, u_t).

    ``keep`` (dict) receives every gather output so the caller can read the per-occurrence
    (IndexedSlices) gradient values.
    """
    u, c, i, hist_i, hist_i_new, hist_t, sl, sl_new = _feed(batch, cand_index)
    icl = torch.as_tensor(np.asarray(icl, dtype=np.int64))
    dt = P["item_emb"].dtype
    g = {}
    g["i_item"] = P["item_emb"][i]                                                   # :84
    g["i_cate"] = P["cate_emb"][icl[i]]                                              # :85
    i_emb = torch.cat([g["i_item"], g["i_cate"]], -1)                                # :86
    g["i_b"] = P["item_b"][i]                                                        # :87
    g["u_user"] = P["user_emb"][u]                                                   # :93
    g["u_cate"] = P["cate_emb"][c]                                                   # :94
    u_emb = torch.cat([g["u_user"], g["u_cate"]], -1)                                # :95
    g["ut"] = P["usert_emb"][u]                                                      # :98
    ut_emb = (g["ut"] * hist_t.to(dt)).unsqueeze(-1).repeat(1, 1, hidden_units)      # :99-102
    g["h_item"] = P["item_emb"][hist_i]                                              # :105
    g["h_cate"] = P["cate_emb"][icl[hist_i]]                                         # :106
    h_emb = torch.cat([g["h_item"], g["h_cate"]], -1) * (P["gamma_parameter"] * ut_emb)  # :107-109
    g["hn_item"] = P["item_emb"][hist_i_new]                                         # :111
    g["hn_cate"] = P["cate_emb"][icl[hist_i_new]]                                    # :112
    h_emb_new = torch.cat([g["hn_item"], g["hn_cate"]], -1)                          # :113
    if keep is not None:
        for v in g.values():
            if v.requires_grad:
                v.retain_grad()
        keep.update(g)
    u_t, _, _ = attention_net(h_emb, h_emb_new, sl, sl_new, P, num_heads)            # :122-134
    u_t = u_t + u_emb                                                                # :135
    logits = torch.sum(u_t * i_emb, -1) + g["i_b"]                                   # :137
    return logits, u_t


def _to_torch(params, dtype, requires_grad=False):
    P = OrderedDict()
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        P[k] = t
    return P


def forward_logits(params, icl, batch, cand_index=1, dtype=torch.float32, config=None):
    """sess.run(self.logits, ...) of model.py:239-261.  numpy in, numpy out."""
    nh = 8 if config is None else config["num_heads"]
    hu = 64 if config is None else config["hidden_units"]
    with torch.no_grad():
        logits, u_t = build_forward(_to_torch(params, dtype), icl, batch, cand_index, nh, hu)
    return logits.numpy(), u_t.numpy()


def bce_with_logits(x, y):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*y + log1p(exp(-|x|))."""
    return torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-torch.abs(x)))


# tf.train.*Optimizer constructor defaults of TF 1.8 (model.py:188-193 passes learning_rate only)
OPT_DEFAULTS = dict(adam=dict(beta1=0.9, beta2=0.999, epsilon=1e-8),
                    rmsprop=dict(decay=0.9, momentum=0.0, epsilon=1e-10),
                    adadelta=dict(rho=0.95, epsilon=1e-8))


def init_opt_state(params, optimizer, dtype=torch.float64):
    """Slot variables as TF 1.8 creates them: adam m, v = 0; rmsprop rms = ONES, momentum = 0; adadelta
    accum, accum_update = 0.  ``t`` counts apply_gradients calls (adam's beta powers)."""
    z = lambda v: torch.zeros(np.shape(v), dtype=dtype)
    s1 = OrderedDict((k, torch.ones(np.shape(v), dtype=dtype) if optimizer == "rmsprop" else z(v))
                     for k, v in params.items())
    return dict(t=0, s1=s1, s2=OrderedDict((k, z(v)) for k, v in params.items()))


def _opt_update(optimizer, w, g, lr, s1, s2, t, touched=None):
    """One apply op.  g = the aggregated, clipped gradient.  Every table with an L2 term has a gradient whose
    IndexedSlices cover all rows (tf.gradients concatenates the gather slices with the dense reg*W slice), and
    Optimizer._apply_sparse_duplicate_indices sums duplicates first, so its sparse apply equals the dense formula.
    item_b (no L2 term) is truly sparse: ``touched`` = bool mask of the rows in its slices; TF's
    sparse_apply_rms_prop / sparse_apply_adadelta update only those rows (weights and slots), while
    AdamOptimizer._apply_sparse (non-lazy) decays m, v and steps every row."""
    o = OPT_DEFAULTS[optimizer]
    if optimizer == "adam":
        b1, b2, eps = o["beta1"], o["beta2"], o["epsilon"]
        m = b1 * s1 + (1 - b1) * g
        v = b2 * s2 + (1 - b2) * g * g
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        return w - lr_t * m / (torch.sqrt(v) + eps), m, v
    if optimizer == "rmsprop":
        rho, mom_c, eps = o["decay"], o["momentum"], o["epsilon"]
        ms = rho * s1 + (1 - rho) * g * g
        mom = mom_c * s2 + lr * g / torch.sqrt(ms + eps)
        nw = w - mom
    else:
        rho, eps = o["rho"], o["epsilon"]
        ms = rho * s1 + (1 - rho) * g * g
        upd = torch.sqrt(s2 + eps) / torch.sqrt(ms + eps) * g
        mom = rho * s2 + (1 - rho) * upd * upd
        nw = w - lr * upd
    if touched is not None:
        nw, ms, mom = torch.where(touched, nw, w), torch.where(touched, ms, s1), torch.where(touched, mom, s2)
    return nw, ms, mom


def train_step(params, icl, batch, lr, config=None, dtype=torch.float32, clip_mode="tf", opt_state=None):
    """One ``Model.train`` call (model.py:208-234): loss, tf.gradients, clip_by_global_norm, then the optimizer of
    ``config['optimizer']`` (:188-195; default GradientDescentOptimizer; the others need ``opt_state`` from
    ``init_opt_state``, updated in place).  Returns dict(loss, bce, grads (dense, un-clipped, incl. reg),
    occ (per-gather slice gradients), norm_tf, norm_agg, scale, new_params)."""
    config = config or {}
    reg = config.get("regulation_rate", 0.00005)
    clip = config.get("max_gradient_norm", 5.0)
    nh, hu = config.get("num_heads", 8), config.get("hidden_units", 64)
    P = _to_torch(params, dtype, requires_grad=True)
    keep = {}
    logits, _ = build_forward(P, icl, batch, 1, nh, hu, keep=keep)
    y = torch.as_tensor(np.asarray(batch[2], dtype=np.float32)).to(dtype)
    l2_norm = sum(torch.sum(P[k] ** 2) / 2 for k in TABLE_NAMES)                    # :164-169
    bce = torch.mean(bce_with_logits(logits, y))
    loss = bce + reg * l2_norm                                                       # :171-172
    loss.backward()                                                                  # :198
    grads = OrderedDict((k, P[k].grad.detach().clone()) for k in PARAM_NAMES)
    occ = {k: v.grad.detach().clone() for k, v in keep.items() if v.grad is not None}
    # --- tf.clip_by_global_norm (:201)
    sq_tf = sum(torch.sum(v ** 2) for v in occ.values())                             # gather slices
    sq_tf = sq_tf + sum(torch.sum((reg * P[k].detach()) ** 2) for k in TABLE_NAMES)  # dense reg slice
    dense_names = [k for k in PARAM_NAMES if k not in TABLE_NAMES and k != "item_b"]
    sq_tf = sq_tf + sum(torch.sum(grads[k] ** 2) for k in dense_names)
    sq_agg = sum(torch.sum(v ** 2) for v in grads.values())
    norm_tf, norm_agg = torch.sqrt(sq_tf), torch.sqrt(sq_agg)
    norm = norm_tf if clip_mode == "tf" else norm_agg
    one = torch.ones((), dtype=dtype)
    scale = clip * torch.minimum(one / norm, one / clip)                             # clip_ops.py
    new_params = OrderedDict()
    optimizer = config.get("optimizer", "sgd")
    if optimizer in OPT_DEFAULTS:
        if opt_state is None:
            raise ValueError("optimizer %r needs opt_state=init_opt_state(params, optimizer)" % optimizer)
        opt_state["t"] += 1
        touched = torch.zeros(P["item_b"].shape, dtype=torch.bool)
        touched[torch.as_tensor(np.asarray(batch[1], dtype=np.int64))] = True        # i_b = gather(item_b, self.i), :87
    for k in PARAM_NAMES:                                                            # :204
        if optimizer in OPT_DEFAULTS:
            w, s1, s2 = _opt_update(optimizer, P[k].detach(), grads[k] * scale, lr, opt_state["s1"][k].to(dtype),
                                    opt_state["s2"][k].to(dtype), opt_state["t"], touched if k == "item_b" else None)
            opt_state["s1"][k], opt_state["s2"][k] = s1, s2
            new_params[k] = w.numpy()
        else:
            new_params[k] = (P[k].detach() - lr * (grads[k] * scale)).numpy()
    return dict(loss=float(loss.detach()), bce=float(bce.detach()), logits=logits.detach().numpy(),
                grads=OrderedDict((k, v.numpy()) for k, v in grads.items()),
                occ={k: v.numpy() for k, v in occ.items()},
                norm_tf=float(norm_tf), norm_agg=float(norm_agg), scale=float(scale),
                new_params=new_params)


def eval_auc(params, icl, batch, dtype=torch.float32, config=None):
    """Model.eval_auc, model.py:237-263: two forwards (batch[1] pos, batch[2] neg)."""
    res1, _ = forward_logits(params, icl, batch, 1, dtype, config)
    res2, _ = forward_logits(params, icl, batch, 2, dtype, config)
    return float(np.mean(res1 - res2 > 0)), res1, res2


def eval_logits_all(params, icl, batch, dtype=torch.float32, config=None):
    """self.eval_logits, model.py:89-91,140: u_t @ all_emb^T + item_b  -> [B, NI]."""
    _, u_t = forward_logits(params, icl, batch, 1, dtype, config)
    item = np.asarray(params["item_emb"]).astype(u_t.dtype)
    cate = np.asarray(params["cate_emb"]).astype(u_t.dtype)[np.asarray(icl)]
    all_emb = np.concatenate([item, cate], -1)
    return u_t @ all_emb.T + np.asarray(params["item_b"]).astype(u_t.dtype)


KS = (1, 10, 20, 30, 40, 50)                                                         # model.py:144-156


def label_ranks(scores, labels):
    """0-based rank of the label item under tf.nn.top_k ordering (ties -> lower index)."""
    labels = np.asarray(labels)
    s_lab = scores[np.arange(scores.shape[0]), labels][:, None]
    idx = np.arange(scores.shape[1])[None, :]
    ahead = (scores > s_lab) | ((scores == s_lab) & (idx < labels[:, None]))
    return ahead.sum(1)


class StreamingTopK:
    """tf.metrics.precision_at_k / recall_at_k accumulators (model.py:142-156); never reset
    by the reference driver (train.py:75-76,82 initialise them once)."""

    def __init__(self):
        self.tp = np.zeros(len(KS)); self.fp = np.zeros(len(KS)); self.fn = np.zeros(len(KS))

    def update(self, scores, labels):
        r = label_ranks(scores, labels)
        for n, k in enumerate(KS):
            hit = float(np.sum(r < k))
            self.tp[n] += hit
            self.fp[n] += scores.shape[0] * k - hit
            self.fn[n] += scores.shape[0] - hit
        return self.precision(), self.recall()

    def precision(self):
        return list(self.tp / (self.tp + self.fp))

    def recall(self):
        return list(self.tp / (self.tp + self.fn))


# ----------------------------------------------------------------------------- data path
GAP = np.array([2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])             # build_dataset.py:16


def time_bucket(d):
    """n = sum(d >= gap), build_dataset.py:18-21 (d = cur_t - t + 1, in days)."""
    return int(np.sum(d >= GAP))


def time_weight(d):
    """hist_t value as it reaches the graph: float64 1/n (build_dataset.py:20) stored into
    an np.float32 array by input.py:36,45."""
    return np.float32(1 / np.sum(d >= GAP))


def bucket_lut():
    """13-entry LUT n -> float32(1/n) (n = 0 is unreachable, kept as 0)."""
    lut = np.zeros(13, np.float32)
    for n in range(1, 13):
        lut[n] = np.float32(1 / np.float64(n))
    return lut


def collate_train(ts, k):
    """DataInput.__next__, input.py:17-54 (same loops, same dtypes)."""
    u, i, y, sl, new_sl, c = [], [], [], [], [], []
    for t in ts:
        u.append(t[0]); i.append(t[4]); y.append(t[5]); c.append(t[6])
        sl.append(min(len(t[1]), k)); new_sl.append(len(t[2]))
    return (u, i, y) + _pad(ts, k, max(new_sl)) + (sl, new_sl, c)


def collate_test(ts, k):
    """DataInputTest.__next__, input.py:70-107."""
    u, i, j, sl, new_sl, c = [], [], [], [], [], []
    for t in ts:
        u.append(t[0]); i.append(t[4][0]); j.append(t[4][1]); c.append(t[5])
        sl.append(min(len(t[1]), k)); new_sl.append(len(t[2]))
    return (u, i, j) + _pad(ts, k, max(new_sl)) + (sl, new_sl, c)


def _pad(ts, k, max_new_sl):
    hist_i = np.zeros([len(ts), k], np.int64)                                       # input.py:35
    hist_t = np.zeros([len(ts), k], np.float32)                                     # :36
    hist_i_new = np.zeros([len(ts), max_new_sl], np.int64)                          # :37
    for kk, t in enumerate(ts):
        length = len(t[1])
        lo = length - k if length > k else 0                                        # :41-49
        n = min(length, k)
        hist_i[kk, :n] = t[1][lo:lo + n]
        hist_t[kk, :n] = t[3][lo:lo + n]
        hist_i_new[kk, :len(t[2])] = t[2]                                           # :50-51
    return hist_i, hist_i_new, hist_t
