"""The N > 1 host logic on CPU: two gloo ranks shard a batch with brotli_rs_b200.shard_streams, decode their shards
(the oracle stands in for the device -- this test exercises sharding and reduction, not the decoder), and the
all-reduced totals and the union of shards equal the single-rank result.  No data-path collective is involved."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, corpus_files


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brotli_rs_b200 import shard_streams
    from oracle import oracle
    files = corpus_files()
    streams = [c for _, c, _ in files] * 3
    caps = [len(e) if e is not None else 70000 for _, _, e in files] * 3
    mine = shard_streams([len(s) for s in streams], caps, world, rank)
    total = 0
    statuses = np.full(len(streams), -1, dtype=np.int64)
    for i in mine:
        st, out = oracle.decode(streams[i])
        statuses[i] = st
        if st == 0:
            total += len(out)
    t = torch.tensor([float(total), float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s = torch.from_numpy(statuses)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)      # "max over ranks" timing reduction used by bench.py
    if rank == 0:
        q.put((t.tolist(), s.numpy().tolist(), tmax.item()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    (total, count), statuses, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = corpus_files()
    assert count == 3 * len(files)
    assert total == 3 * 3094120
    assert tmax == 2.0
    want = [0 if e is not None else None for _, _, e in files] * 3
    for st, w in zip(statuses, want):
        assert st >= 0 and (w is None or st == 0)


def _scatter_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brotli_rs_b200 import gather_outputs, scatter_batch
    from brotli_rs_b200.batch import slot_offsets
    from oracle import oracle
    files = corpus_files()
    streams = [c for _, c, _ in files] * 2
    caps = np.array([len(e) if e is not None else 70000 for _, _, e in files] * 2, dtype=np.int64)
    if rank == 0:      # the batch is resident on rank 0 only
        buf = torch.from_numpy(np.frombuffer(b"".join(streams), dtype=np.uint8).copy())
        mine, shard, lens, mycaps = scatter_batch(buf, np.array([len(s) for s in streams]), caps, src=0)
    else:
        mine, shard, lens, mycaps = scatter_batch(None, None, None, src=0)
    # decode the shard (the oracle stands in for the device) into slots
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    out_off = slot_offsets(mycaps)
    o, ol, st = oracle.decode_batch(shard.numpy(), off, out_off)
    res = gather_outputs(mine, torch.from_numpy(o), out_off, torch.from_numpy(ol.astype(np.int64)), torch.from_numpy(st.astype(np.int32)),
                         len(streams), dst=0)
    if rank == 0:
        st_all, len_all, rank_of, slot_off, buffers = res
        bad = 0
        for i, s in enumerate(streams):
            want_st, want = oracle.decode(s)
            got = buffers[rank_of[i]][slot_off[i]: slot_off[i] + len_all[i]].numpy().tobytes()
            bad += not (st_all[i] == want_st and (want_st != 0 or got == want))
        q.put((bad, len(streams), sorted(set(rank_of.tolist())), [int(b.numel()) for b in buffers]))
    else:
        assert res is None
    dist.destroy_process_group()


def test_two_rank_scatter_decode_gather():
    """batch resident on rank 0 -> scatter_batch -> every rank decodes its shard -> gather_outputs on rank 0: every
    stream of the batch comes back with the oracle's status and bytes, half of them via the other rank"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_scatter_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    bad, n, ranks, sizes = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert bad == 0 and n == 2 * len(corpus_files()) and ranks == [0, 1]
    assert abs(sizes[0] - sizes[1]) < 0.2 * max(sizes)       # the serpentine deal balances the shards
