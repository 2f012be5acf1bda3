#!/usr/bin/env python3
"""Decode time of every corpus stream ALONE (one stream per call: the fused kernel's latency build, one warp), and of the same
stream on a parse-kernel thread (two-phase forced) -- what bounds a batch that holds the stream.
usage: python tools/time_single.py [min compressed bytes [name substring]]   (BRO_SINGLE_MODES=fused,two-phase)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brotli_rs_b200 import BatchDecoder
from brotli_rs_b200.batch import pack_streams, slot_offsets
from brotli_rs_b200.workloads import corpus_workload

min_bytes = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
only = sys.argv[2] if len(sys.argv) > 2 else ""          # substring of the file name
modes = os.environ.get("BRO_SINGLE_MODES", "fused,two-phase").split(",")
names, streams, raws, status = corpus_workload(os.path.join(ROOT, "tests", "golden", "data"))
for mode, label in ((BatchDecoder.MODE_WARP, "fused"), (BatchDecoder.MODE_TWOPHASE, "two-phase")):
    if label not in modes:
        continue
    dec = BatchDecoder(0, mode=mode)
    for nm, s, raw, st in sorted(zip(names, streams, raws, status), key=lambda t: -len(t[1])):
        if len(s) < min_bytes or st != 0 or only not in nm:
            continue
        in_buf, in_off = pack_streams([s])
        out_off = slot_offsets([len(raw)])
        d_in = torch.from_numpy(np.concatenate([in_buf, np.zeros(16, dtype=np.uint8)])).cuda()
        d_in_off = torch.from_numpy(in_off.astype(np.int64)).cuda()
        d_out_off = torch.from_numpy(out_off.astype(np.int64)).cuda()
        d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d_len, d_st = dec.decode_device(d_in, d_in_off, d_out, d_out_off)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        ok = int(d_st[0]) == 0 and d_out[: len(raw)].cpu().numpy().tobytes() == raw
        print("%-10s %-36s in %7d out %7d  %8.2f ms  %s" % (label, nm, len(s), len(raw), best, "ok" if ok else "MISMATCH"), flush=True)
    dec.close()
