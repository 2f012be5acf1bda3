#!/usr/bin/env python3
"""Mutation check of the 32-lane host simulation (tests/warpsim.py): every bro_syncwarp() of bro_decoder_core.h that the fused
kernel's code executes on the test data, and every __syncwarp() of the copy kernel's product path, is left out in turn; the table
says which means notices -- wrong bytes / status under one of the lane orders (ascending, descending, shuffled; several slot
alignments), or a ThreadSanitizer report under the CUDA memory model (only __syncwarp orders memory between lanes).  A barrier
nobody notices is either redundant on this data (another barrier follows before the data is used) or protects something the data
does not exercise; it is NOT evidence that it can go.

    python tools/warpsim_barriers.py > profiles/r02_warpsim_barriers.md
"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


_S = {}


def _fused_line(line):
    """worker: one barrier of the fused code left out -> (line, what the lane orders notice, race detector's reports)"""
    import warpsim
    S = _state()
    with warpsim.drop_sync(line):
        o = S["fused_orders"]()
    warpsim.set_alignment(0, 0)
    return line, o, S["fused_tsan"](line)


def _state():
    if not _S:
        main(build_only=True)
    return _S


def main(build_only=False):
    import fuzzgen
    import hostsim
    import warpsim
    from conftest import corpus_files
    from oracle import oracle
    import test_warpsim_parity as T

    streams = [(n, c) for n, c, _ in corpus_files() if len(c) <= 170000]
    enc = fuzzgen.libbrotli_enc()
    if enc is not None:
        k = 0
        for kind in ("words", "skewed", "runs", "repeat2k"):
            for q, lgwin, size in ((5, 16, 60000), (9, 18, 40000), (11, 16, 30000), (10, 22, 30000)):
                k += 1
                streams.append(("fresh%02d" % k, fuzzgen.compress(enc, fuzzgen.synthetic_raw(kind, 900 + k, size), q, lgwin)))
    exp = [oracle.decode(c) for _, c in streams]
    tmp = tempfile.mkdtemp()
    files = []
    for (n, c), (st, out) in zip(streams, exp):
        p = os.path.join(tmp, n)
        open(p, "wb").write(c)
        files.append((p, len(out)))
    exe = T._tsan_binary()

    def fused_orders():
        for (n, c), (st, out) in zip(streams, exp):
            for latency in (False, True):
                for order in warpsim.ORDERS:
                    for om in (0, 1):
                        warpsim.set_alignment(8 * om + 3, om)
                        try:
                            got = warpsim.decode(c, cap=len(out), latency=latency, order=order)
                        except AssertionError:
                            return "%s (%s build, order %d, slot alignment %d): the warp diverged" % (n, "latency" if latency else "throughput", order, om)
                        if got[0] != st or (st == 0 and got[1] != out):
                            return "%s (%s build, order %d, slot alignment %d)" % (n, "latency" if latency else "throughput", order, om)
        return None

    def fused_tsan(line):
        total = 0
        for latency in (0, 1):
            try:
                races, _, _ = T._tsan_run(exe, files, latency=latency, order=0, align=(3, 1), drop=line, halt=False)
            except AssertionError:
                return "the warp diverged"
            total += races
        return total

    _S["fused_orders"] = fused_orders
    _S["fused_tsan"] = fused_tsan
    if build_only:
        return
    import multiprocessing
    warpsim.sync_hits()
    assert fused_orders() is None
    hits = warpsim.sync_hits()
    lines = open(os.path.join(ROOT, "brotli_rs_b200", "csrc", "bro_decoder_core.h")).read().split("\n")
    hits = {ln: c for ln, c in hits.items() if "bro_syncwarp();" in lines[ln - 1]}      # (the simulation's own driver has one barrier, on a line of its file)
    print("# Barriers of the warp code, left out one at a time (tools/warpsim_barriers.py)\n")
    print("Data: %d streams (the corpus up to 170 KB compressed + 16 fresh libbrotli streams, qualities 5 - 11).  With every barrier in place: no wrong"
          % len(streams))
    print("answer under any lane order, 0 ThreadSanitizer reports.\n")
    print("## Fused kernel's code (`bro_decoder_core.h`, 32-lane form): %d barrier sites executed\n" % len(hits))
    print("| line | executions | source | lane orders notice | race detector (reports) |")
    print("|---|---|---|---|---|")
    n_any = n_ord = n_ts = 0
    with multiprocessing.Pool(min(8, os.cpu_count() or 1)) as pool:
        found = {ln: (o, ts) for ln, o, ts in pool.map(_fused_line, sorted(hits), chunksize=1)}
    for line, cnt in sorted(hits.items()):
        ctx = ""
        for j in range(line - 2, max(line - 8, 0), -1):
            t = lines[j].strip()
            if t and not t.startswith("//") and not t.startswith("#"):
                ctx = t
                break
        o, ts = found[line]
        n_ord += o is not None
        n_ts += ts != 0
        n_any += (o is not None) or ts != 0
        print("| %d | %d | after `%s` | %s | %s |" % (line, cnt, ctx.replace("|", "\\|")[:90], o or "-", ts or "-"), flush=True)
    print("\n%d of %d sites noticed (%d by the lane orders, %d by the race detector).\n" % (n_any, len(hits), n_ord, n_ts))

    # the copy kernel
    cs = [(n, c) for n, c, _ in corpus_files()]
    cexp = [oracle.decode(c) for _, c in cs]
    warpsim.two_phase([c for _, c in cs], [len(o) for _, o in cexp])
    L = warpsim._LIB_COPY
    h = np.zeros(4096, dtype=np.uint64)
    L.bro_warpsim_copy_sync_hits(ctypes.c_void_p(h.ctypes.data), 4096, 1)
    klines = open(os.path.join(ROOT, "brotli_rs_b200", "csrc", "bro_kernels_copy.cu")).read().split("\n")
    cexe = os.path.join(ROOT, "tests", "_build", "warpsim_copy_tsan")
    cfiles = []
    for (n, c), (st, out) in zip(cs, cexp):
        p = os.path.join(tmp, "c_" + n)
        open(p, "wb").write(c)
        cfiles.append("%s:%d" % (p, len(out)))
    print("## Copy kernel (`bro_kernels_copy.cu`, the product's path): %d barrier sites executed\n" % len(h.nonzero()[0]))
    print("| line | executions | source | lane orders notice | race detector (reports) |")
    print("|---|---|---|---|---|")
    for line in [int(x) for x in h.nonzero()[0]]:
        L.bro_warpsim_copy_drop_sync(line)
        o = None
        try:
            for order in warpsim.ORDERS:
                for om in (0, 5, 11):
                    try:
                        res, _ = warpsim.two_phase([c for _, c in cs], [len(x) for _, x in cexp], order=order, out_mis=om, in_mis=om)
                    except AssertionError:
                        o = "the warp diverged (order %d)" % order
                        break
                    for (n, _), e, r in zip(cs, cexp, res):
                        if not (r[0] in hostsim.RETRY or (r[0] == e[0] and (e[0] != 0 or r[1] == e[1]))):
                            o = "%s (order %d, slot alignment %d)" % (n, order, om)
                            break
                    if o:
                        break
                if o:
                    break
        finally:
            L.bro_warpsim_copy_drop_sync(-1)
        env = dict(os.environ, BRO_WS_ALIGN="3,5", BRO_WS_DROP_SYNC=str(line),
                   TSAN_OPTIONS="exitcode=66 suppressions=" + os.path.join(ROOT, "tests", "warpsim_tsan.supp"))
        r = subprocess.run([cexe, "0", "0", "1", "0"] + cfiles, env=env, capture_output=True, text=True)
        ts = r.stderr.count("WARNING: ThreadSanitizer")
        print("| %d | %d | `%s` | %s | %s |" % (line, int(h[line]), klines[line - 1].strip().replace("|", "\\|")[:100], o or "-", ts or "-"), flush=True)


if __name__ == "__main__":
    main()
