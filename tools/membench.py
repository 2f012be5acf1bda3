#!/usr/bin/env python3
"""What bounds a copy whose sources are L2 hits?  Times (CUDA events, best of 5) on one GPU:
  fill      write-only: N bytes written
  copy      plain copy: N bytes read from DRAM + N written (the MEASURED_PEAKS.json method)
  bcast     N bytes written, sources read from a small buffer that stays in L2 (the copy kernel's situation)
"""
import json, sys, torch
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 24 * (1 << 30)
dev = torch.device("cuda", 0)
big = torch.empty(N, dtype=torch.uint8, device=dev)
src = torch.empty(N, dtype=torch.uint8, device=dev).random_(0, 255)
res = {}
def timeit(name, fn, moved):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    res[name] = {"ms": best, "GB/s (bytes moved)": moved / best / 1e6, "GB/s (bytes written)": N / best / 1e6}
timeit("fill", lambda: big.fill_(7), N)
timeit("copy", lambda: big.copy_(src), 2 * N)
for small_mb in (16, 48):
    small = src[: small_mb << 20]
    rows = N // small.numel()
    v = big[: rows * small.numel()].view(rows, small.numel())
    timeit("bcast_%dMB" % small_mb, lambda: v.copy_(small.unsqueeze(0).expand(rows, -1)), 2 * rows * small.numel())
print(json.dumps(res, indent=1))
