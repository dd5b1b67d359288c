"""Data-parallel helpers: samples are independent, so the global batch
is split into contiguous row blocks, one per rank; tables and dense parameters are replicated;
the only collective of a train step is ONE all-reduce(sum) of the flat gradient buffer produced
by ``tlsan_step_grads`` (sparse-part table gradients, dense gradients, sum-of-squares and loss
partials).  Scoring shards rows with no collective at all."""
import numpy as np


def row_block(n_rows, rank, world):
    """[lo, hi) of the contiguous block owned by ``rank`` (first ``n_rows % world`` ranks get +1)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(batch, rank, world):
    """Slice the input.py 9-tuple to this rank's rows.  ``hist_i_new`` keeps the global width
    (its padding columns are never read), so every rank sees the same S."""
    n = len(batch[0])
    lo, hi = row_block(n, rank, world)
    return tuple(np.asarray(f)[lo:hi] for f in batch), (lo, hi)
