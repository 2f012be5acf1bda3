#!/usr/bin/env python3
"""Fuzz harness standing in for the reference's AFL workflow (docs/notes_afl.txt, src/main.rs:49-70): mutated corpus
streams are decoded on the GPU by both paths, without size hints, and (--streaming N: the first N of them) through the
streaming Read-struct, and compared with the oracle (status class and bytes).

    python tools/fuzz_gpu.py [--count 4000] [--seed 1] [--slack exact|tight|generous]
    compute-sanitizer --tool memcheck python tools/fuzz_gpu.py --count 600      (memory safety of the kernels)

Exit status 0 = every stream agreed."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--slack", default="tight", choices=["exact", "tight", "generous"])
    ap.add_argument("--streaming", type=int, default=200, help="streams also decoded through Decompressor(streaming=...) (resumable kernel)")
    ap.add_argument("--blocktypes", type=int, default=0,
                    help="seed the mutations with this many fresh libbrotli streams of heterogeneous payloads (quality 5-9: literal / "
                         "insert&copy / distance block types and block switches without literal context modelling) instead of the corpus")
    args = ap.parse_args()
    import fuzzgen
    from brotli_rs_b200 import BatchDecoder
    from oracle import oracle
    data = os.path.join(ROOT, "tests", "golden", "data")
    corpus = [open(os.path.join(data, f), "rb").read() for f in sorted(os.listdir(data)) if ".compressed" in f]
    if args.blocktypes:
        enc = fuzzgen.libbrotli_enc()
        if enc is None:
            print("libbrotlienc.so.1 not present: --blocktypes needs it")
            return 2
        g = np.random.default_rng(args.seed + 77)
        kinds = ["words", "skewed", "small_alpha", "random", "runs", "repeat2k"]
        corpus = []
        for i in range(args.blocktypes):
            raw = bytearray()
            size = int(g.integers(20000, 260000))
            while len(raw) < size:
                raw += fuzzgen.synthetic_raw(kinds[int(g.integers(len(kinds)))], int(g.integers(1 << 30)), int(g.integers(2000, 40000)))
            corpus.append(fuzzgen.compress(enc, bytes(raw[:size]), int(g.integers(5, 10)), int(g.integers(16, 23))))
    streams = list(fuzzgen.mutations(corpus, seed=args.seed, count=args.count)) + [c for c in corpus if len(c) < 70000 or args.blocktypes]
    rng = np.random.default_rng(args.seed)
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    caps = []
    for s in streams:
        st, out = oracle.decode(s)
        caps.append(len(out) if args.slack == "exact" else len(out) + 4096 if args.slack == "generous" else int(rng.integers(0, len(out) + 64)))
    in_buf, in_off = pack_streams(streams)
    out_off = slot_offsets(caps)              # slot i = [out_off[i], out_off[i+1]): the same slots for the GPU and the oracle
    ref, ref_len, ref_st = oracle.decode_batch(in_buf, in_off, out_off, nthreads=8)
    bad = 0
    for mode, name in ((BatchDecoder.MODE_WARP, "fused"), (BatchDecoder.MODE_TWOPHASE, "two-phase")):
        dec = BatchDecoder(0, mode=mode)
        out, out_len, status = dec.decode_host(in_buf, in_off, out_off)
        for i in range(len(streams)):
            b, n = int(out_off[i]), int(ref_len[i])
            if int(status[i]) != int(ref_st[i]) or (int(ref_st[i]) == 0 and (int(out_len[i]) != n or not np.array_equal(out[b: b + n], ref[b: b + n]))):
                bad += 1
                print("MISMATCH %s stream %d: status %d vs %d, %d bytes vs %d; %s" % (name, i, int(status[i]), int(ref_st[i]), int(out_len[i]), n, streams[i][:16].hex()))
        res = dec.decode_unsized(streams[: max(1, len(streams) // 4)])
        for i, (st, out) in enumerate(res):
            wst, wout = oracle.decode(streams[i])
            if st != wst or (st == 0 and out != wout):
                bad += 1
                print("MISMATCH %s unsized stream %d: status %d vs %d" % (name, i, st, wst))
        dec.close()
        print("%s: %d streams, %d status classes" % (name, len(streams), len(set(int(x) for x in status))), flush=True)
    if args.streaming:
        from brotli_rs_b200 import BroError, Decompressor
        dec = BatchDecoder(0)
        k = min(args.streaming, len(streams))
        for i in range(k):
            wst, wout = oracle.decode(streams[i])
            got, st = b"", 0
            r = Decompressor(streams[i], decoder=dec, streaming=int(rng.integers(1, 4000)))
            try:
                got = r.read()
            except BroError as e:
                st = e.status
            r.close()
            if st != wst or (st == 0 and got != wout):
                bad += 1
                print("MISMATCH streaming stream %d: status %d vs %d; %s" % (i, st, wst, streams[i][:16].hex()))
        dec.close()
        print("streaming reader: %d streams" % k, flush=True)
    print("fuzz_gpu: %d streams x 2 paths, %d mismatches" % (len(streams), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
