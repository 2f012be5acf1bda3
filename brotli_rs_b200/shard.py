"""Multi-GPU sharding of a stream batch.

Streams are independent, so the batch is partitioned by stream with no data-path collective: every rank decodes
its shard with its own context.  The partition balances compressed AND expected uncompressed bytes: streams are
sorted by cost (descending) and dealt round-robin in a serpentine order, which keeps every rank's byte totals
within one stream of each other and gives each rank the same mix of large and small streams.
"""
import numpy as np


def shard_streams(in_lens, out_lens, world_size, rank=None):
    """Return the stream indices of every rank (list of int64 arrays), or of `rank` only.

    in_lens / out_lens: per-stream compressed sizes and expected (or capacity) uncompressed sizes."""
    in_lens = np.asarray(in_lens, dtype=np.int64)
    out_lens = np.asarray(out_lens, dtype=np.int64)
    n = len(in_lens)
    cost = in_lens + out_lens
    order = np.argsort(-cost, kind="stable")
    pos = np.arange(n)
    rnd, k = pos // world_size, pos % world_size
    owner = np.where(rnd % 2 == 0, k, world_size - 1 - k)          # serpentine deal
    shards = [np.sort(order[owner == r]) for r in range(world_size)]
    return shards if rank is None else shards[rank]
