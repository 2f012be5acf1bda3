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


# ---- batch scatter / gather over torch.distributed (NCCL between GPUs; gloo in the CPU tests) ----
#
# The decode itself needs no exchange.  These two helpers cover the one situation SURVEY.md section 8e names in which
# bytes cross GPUs at all: the whole compressed batch is resident on one rank (it arrived there from the network or a
# file) and the decoded streams are wanted back on that rank.  Every transfer is point-to-point and the transfers of a
# call are posted together (dist.batch_isend_irecv -> one ncclGroupStart/End), so over NVSwitch each peer link carries
# one message of ~total/N bytes.  torch tensors only: the buffers stay on whatever device the caller put them.

def _p2p(ops):
    import torch.distributed as dist
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def scatter_batch(buf, in_lens, caps, src=0, device=None, group=None):
    """Rank `src` holds the batch: `buf` (uint8 tensor, all compressed streams back to back), `in_lens` and `caps`
    (per-stream compressed sizes and output capacities; int64 numpy arrays).  Other ranks pass None for all three.
    Every rank gets its shard as chosen by shard_streams: -> (indices into the batch, uint8 tensor with the shard's
    streams back to back, their sizes, their capacities)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = buf.device if rank == src else torch.device("cpu")
    n_t = torch.tensor([len(in_lens) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n_t, src, group=group)
    n = int(n_t.item())
    meta = torch.empty((2, n), dtype=torch.int64, device=device)
    if rank == src:
        meta[0] = torch.from_numpy(np.asarray(in_lens, dtype=np.int64)).to(device)
        meta[1] = torch.from_numpy(np.asarray(caps, dtype=np.int64)).to(device)
    dist.broadcast(meta, src, group=group)            # 16 bytes per stream; every rank derives the same partition
    in_lens, caps = meta[0].cpu().numpy(), meta[1].cpu().numpy()
    shards = shard_streams(in_lens, caps, world)
    mine = shards[rank]
    if rank == src:
        off = np.concatenate([[0], np.cumsum(in_lens)]).astype(np.int64)
        ops, keep, my_buf = [], [], None
        for r in range(world):
            # the shard's streams, back to back: one gather on the device
            idx = shards[r]
            starts = torch.from_numpy(off[idx]).to(device)
            lens = torch.from_numpy(in_lens[idx]).to(device)
            new_off = torch.cumsum(lens, 0) - lens
            pos = torch.repeat_interleave(starts - new_off, lens) + torch.arange(int(in_lens[idx].sum()), device=device)
            piece = buf[pos].contiguous()
            if r == rank:
                my_buf = piece
            elif piece.numel():
                keep.append(piece)
                ops.append(dist.P2POp(dist.isend, piece, r, group=group))
        _p2p(ops)
    else:
        my_buf = torch.empty(int(in_lens[mine].sum()), dtype=torch.uint8, device=device)
        _p2p([dist.P2POp(dist.irecv, my_buf, src, group=group)] if my_buf.numel() else [])
    return mine, my_buf, in_lens[mine], caps[mine]


def gather_outputs(indices, out, out_off, out_len, status, n_total, dst=0, group=None):
    """The inverse of scatter_batch.  Every rank passes its shard: `indices` (into the batch, as scatter_batch returned
    them), the output buffer `out` (uint8 tensor), its slot offsets `out_off` (n_shard + 1 integers, host), and the
    decoded lengths and statuses (tensors on out's device).  Rank `dst` -> (status[n_total], out_len[n_total],
    rank_of[n_total], slot_off[n_total], [output buffer of rank 0, 1, ...]) with the first four as numpy arrays in
    batch order: stream i's bytes are buffers[rank_of[i]][slot_off[i]: slot_off[i] + out_len[i]].  Other ranks -> None.
    The buffers are not re-interleaved: a slot layout is per shard, and moving 26 GB a second time buys nothing."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = out.device
    k = len(indices)
    assert out_len.numel() == k and status.numel() == k and len(out_off) == k + 1
    sizes = [torch.empty(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([k, out.numel()], dtype=torch.int64, device=device), group=group)
    sizes = [tuple(int(x) for x in t.tolist()) for t in sizes]
    assert sum(t[0] for t in sizes) == n_total
    # per stream: index, slot offset, decoded length, status
    table = torch.stack([torch.from_numpy(np.asarray(indices, dtype=np.int64)).to(device),
                         torch.from_numpy(np.asarray(out_off[:-1]).astype(np.int64)).to(device),
                         out_len.to(torch.int64), status.to(torch.int64)]).contiguous()
    if rank != dst:
        ops = [dist.P2POp(dist.isend, table, dst, group=group)] if k else []
        if out.numel():
            ops.append(dist.P2POp(dist.isend, out, dst, group=group))
        _p2p(ops)
        return None
    tables, buffers, ops = [], [], []
    for r in range(world):
        if r == rank:
            tables.append(table)
            buffers.append(out)
            continue
        tables.append(torch.empty((4, sizes[r][0]), dtype=torch.int64, device=device))
        buffers.append(torch.empty(sizes[r][1], dtype=torch.uint8, device=device))
        if sizes[r][0]:
            ops.append(dist.P2POp(dist.irecv, tables[r], r, group=group))
        if sizes[r][1]:
            ops.append(dist.P2POp(dist.irecv, buffers[r], r, group=group))
    _p2p(ops)
    st_all = np.full(n_total, -1, dtype=np.int32)
    len_all = np.zeros(n_total, dtype=np.int64)
    rank_of = np.full(n_total, -1, dtype=np.int32)
    slot_off = np.zeros(n_total, dtype=np.int64)
    for r in range(world):
        t = tables[r].cpu().numpy()
        idx = t[0]
        slot_off[idx], len_all[idx], st_all[idx], rank_of[idx] = t[1], t[2], t[3].astype(np.int32), r
    assert (rank_of >= 0).all(), "some stream of the batch belongs to no shard"
    return st_all, len_all, rank_of, slot_off, buffers
