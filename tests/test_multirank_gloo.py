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
