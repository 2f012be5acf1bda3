#!/usr/bin/env python3
"""CPU fuzz campaign over the KERNELS' OWN CODE as compiled for the host (tests/warpsim.py: 32 lanes as fibers): mutated streams
-- seeded with the corpus and with fresh libbrotli streams of every quality band, heterogeneous payloads included -- go through
(a) the fused kernel's per-warp code, both builds, and (b) the three launches of the two-phase call as one batch -- parse kernel,
copy kernel, the fused kernel's retry pass; CTAs of 1, 2 or 8 warps -- under a random lane order, stream / slot alignment, slot capacity (exact, too small, generous) and warp occupancy (`lanes`), and are compared
with the oracle (status class and bytes).  The GPU twin of this harness is tools/fuzz_gpu.py.

    python tools/fuzz_warpsim.py [--count 4000] [--seed 1] [--jobs 8] [--fresh 24]

Exit status 0 = every stream agreed."""
import argparse
import multiprocessing
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def seeds(fresh, seed):
    import fuzzgen
    data = os.path.join(ROOT, "tests", "golden", "data")
    corpus = [open(os.path.join(data, f), "rb").read() for f in sorted(os.listdir(data)) if ".compressed" in f]
    enc = fuzzgen.libbrotli_enc()
    if enc is not None and fresh:
        g = np.random.default_rng(seed + 77)
        kinds = ["words", "skewed", "small_alpha", "random", "runs", "repeat2k"]
        for i in range(fresh):
            raw = bytearray()
            size = int(g.integers(3000, 120000))
            while len(raw) < size:
                raw += fuzzgen.synthetic_raw(kinds[int(g.integers(len(kinds)))], int(g.integers(1 << 30)), int(g.integers(1000, 40000)))
            corpus.append(fuzzgen.compress(enc, bytes(raw[:size]), int(g.integers(1, 12)), int(g.integers(10, 23))))
    return corpus


def worker(args):
    job, count, seed, fresh = args
    import fuzzgen
    import hostsim
    import warpsim
    from oracle import oracle
    rng = np.random.default_rng(1000 * seed + job)
    corpus = seeds(fresh, seed)
    streams = list(fuzzgen.mutations(corpus, seed=1000 * seed + job, count=count, max_len=60000))
    bad, classes, handed = [], set(), 0
    caps, exp = [], []
    quirks = job & 1          # odd workers decode in BRO_QUIRKS_SPEC mode, against the oracle in the same mode
    for s in streams:
        st, out = oracle.decode(s, quirks=quirks)
        r = rng.random()
        cap = len(out) if r < 0.4 else len(out) + 4096 if r < 0.6 else int(rng.integers(0, len(out) + 64))
        o, ol, sts = oracle.decode_batch(np.frombuffer(s, dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64), np.array([0, cap], dtype=np.uint64),
                                         quirks=quirks)
        caps.append(cap)
        exp.append((int(sts[0]), o[: int(ol[0])].tobytes()))
        classes.add(int(sts[0]))
    # (a) the fused kernel's code, stream by stream
    for i, s in enumerate(streams):
        latency, order = bool(rng.integers(2)), int(rng.integers(3))
        warpsim.set_alignment(int(rng.integers(128)), int(rng.integers(16)))
        try:
            got = warpsim.decode(s, cap=caps[i], quirks=quirks, latency=latency, order=order, seed=i + 1)
        except AssertionError as e:
            got = ("sim", str(e))
        if got[0] != exp[i][0] or (exp[i][0] == 0 and got[1] != exp[i][1]):
            bad.append(("fused", job, i, exp[i][0], got[0], s.hex()[:64], len(s), caps[i], latency, order))
    # (a') one stream in sixteen through the streaming reader's loop over the resume kernel's code, input in random pieces
    warpsim.set_alignment(0, 0)
    for i in range(0, len(streams), 16):
        s = streams[i]
        chunks = [int(x) for x in rng.integers(1, max(2, len(s)), 4)]
        st, out = oracle.decode(s, quirks=quirks)
        try:
            st1, served, _, _, _ = warpsim.stream_decode(s, chunks, quirks=quirks, order=int(rng.integers(3)), seed=i + 1)
        except AssertionError as e:
            st1, served = "sim", b""
        if st1 != st or (st == 0 and served != out) or (st != 0 and not out.startswith(served)):
            bad.append(("streaming", job, i, st, st1, s.hex()[:64], len(s), chunks))
    # (b) both kernels of the two-phase path, in batches of up to 96 streams
    k = 0
    while k < len(streams):
        m = int(rng.integers(1, 97))
        part = list(range(k, min(k + m, len(streams))))
        k += m
        lanes = int(rng.choice([32, 32, 32, 17, 4, 1]))
        order, copy_order = int(rng.integers(3)), int(rng.integers(3))
        ho = [int(x) for x in rng.permutation(len(part))] if rng.random() < 0.5 else None
        try:
            res, nretry, _ = warpsim.two_phase_kernels([streams[i] for i in part], [caps[i] for i in part], quirks=quirks, lanes=lanes, hand_out=ho, order=order,
                                                       seed=k, in_mis=int(rng.integers(16)), out_mis=int(rng.integers(16)), copy_shape=int(rng.integers(2)),
                                                       copy_order=copy_order, retry_pass=True, retry_latency=bool(rng.integers(2)),
                                                       threads=int(rng.choice([32, 32, 64, 256])))
            handed += nretry
        except AssertionError as e:
            bad.append(("two-phase", job, part[0], "sim", str(e), "", len(part), lanes, order, copy_order))
            continue
        for i, got in zip(part, res):
            if got[0] != exp[i][0] or (exp[i][0] == 0 and got[1] != exp[i][1]):
                bad.append(("two-phase", job, i, exp[i][0], got[0], streams[i].hex()[:64], len(streams[i]), caps[i], lanes, order))
    return bad, classes, handed, len(streams)


def asan_worker(args):
    """one chunk of mutated streams through the address-sanitizer builds of the three simulated kernels (tests/_build/warpsim_*_asan,
    built by tests/test_warpsim_parity.py::test_kernels_under_address_sanitizer) -> (problems, streams)"""
    job, count, seed, fresh, kind = args
    import subprocess
    import tempfile
    import fuzzgen
    import hostsim
    from oracle import oracle
    rng = np.random.default_rng(7000 * seed + job)
    streams = list(fuzzgen.mutations(seeds(fresh, seed), seed=7000 * seed + job, count=count, max_len=60000))
    build = os.path.join(ROOT, "tests", "_build")
    bad = []
    with tempfile.TemporaryDirectory() as tmp:
        for k0 in range(0, len(streams), 300):
            files, expect = [], {}
            for i, s in enumerate(streams[k0: k0 + 300]):
                st, out = oracle.decode(s)
                cap = len(out) if rng.random() < 0.6 else int(rng.integers(0, len(out) + 64))
                o, ol, sts = oracle.decode_batch(np.frombuffer(s, dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64), np.array([0, cap], dtype=np.uint64))
                p = os.path.join(tmp, "s%05d" % (k0 + i))
                open(p, "wb").write(s)
                files.append("%s:%d" % (p, cap))
                expect[p] = (int(sts[0]), int(ol[0]))
            env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:exitcode=55",
                       TSAN_OPTIONS="exitcode=66 suppressions=" + os.path.join(ROOT, "tests", "warpsim_tsan.supp"))
            sfx = "_" + kind
            runs = (([os.path.join(build, "warpsim" + sfx), str(int(rng.integers(2))), str(int(rng.integers(3))), str(k0 + 1), "0", "0"],
                     dict(env, BRO_WS_BATCH=str(int(rng.choice([32, 64, 256]))))),
                    ([os.path.join(build, "warpsim" + sfx), str(int(rng.integers(2))), str(int(rng.integers(3))), str(k0 + 1), "0", "0"],
                     dict(env, BRO_WS_ALIGN="%d,%d" % (int(rng.integers(128)), int(rng.integers(16))))),
                    ([os.path.join(build, "warpsim_parse" + sfx), str(int(rng.choice([32, 32, 9, 1]))), str(int(rng.integers(3))), str(k0 + 1)],
                     dict(env, BRO_WS_THREADS=str(int(rng.choice([32, 64]))))),
                    ([os.path.join(build, "warpsim_copy" + sfx), str(int(rng.integers(2))), str(int(rng.integers(3))), str(k0 + 1), str(k0 + 5)],
                     dict(env, BRO_WS_ALIGN="%d,%d" % (int(rng.integers(16)), int(rng.integers(16))))))
            for cmd, e in runs:
                r = subprocess.run(cmd + files, env=e, capture_output=True, text=True)
                if r.returncode != 0 or "ERROR: AddressSanitizer" in r.stderr or "runtime error" in r.stderr or "WARNING: ThreadSanitizer" in r.stderr:
                    bad.append((os.path.basename(cmd[0]), job, k0, r.returncode, r.stderr[:1500]))
                    continue
                for ln in r.stdout.splitlines():
                    name, st1, n1, _, err = ln.rsplit(" ", 4)
                    st, n = expect[name]
                    if err != "0" or not (int(st1) in hostsim.RETRY or (int(st1) == st and (st != 0 or int(n1) == n))):
                        bad.append((os.path.basename(cmd[0]), job, k0, ln, st))
    return bad, len(streams)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--fresh", type=int, default=24, help="fresh libbrotli streams (qualities 1-11, windows 10-22, heterogeneous payloads) added to the seeds")
    ap.add_argument("--tsan", action="store_true", help="... through the race-detector builds (ThreadSanitizer over the lanes, only __syncwarp / "
                                                        "__syncthreads order memory) instead")
    ap.add_argument("--asan", action="store_true", help="run the streams through the address-sanitizer builds of the simulated kernels instead "
                                                        "(memory safety; statuses and sizes against the oracle)")
    args = ap.parse_args()
    if args.asan or args.tsan:
        per = (args.count + args.jobs - 1) // args.jobs
        with multiprocessing.Pool(args.jobs) as pool:
            results = pool.map(asan_worker, [(j, per, args.seed, args.fresh, "tsan" if args.tsan else "asan") for j in range(args.jobs)])
        bad = [b for r in results for b in r[0]]
        for b in bad[:20]:
            print("PROBLEM", b)
        print("fuzz_warpsim --%s: %d streams x (fused kernel as a batch and stream by stream, parse kernel, copy kernel) under -fsanitize=%s,undefined, %d problems"
              % ("tsan" if args.tsan else "asan", sum(r[1] for r in results), "thread" if args.tsan else "address", len(bad)))
        return 1 if bad else 0
    import warpsim
    warpsim.lib(); warpsim.two_phase([], []); warpsim._lib_parse()       # build once, before the workers start
    per = (args.count + args.jobs - 1) // args.jobs
    with multiprocessing.Pool(args.jobs) as pool:
        results = pool.map(worker, [(j, per, args.seed, args.fresh) for j in range(args.jobs)])
    bad = [b for r in results for b in r[0]]
    classes = set().union(*[r[1] for r in results])
    for b in bad[:40]:
        print("MISMATCH", b)
    print("fuzz_warpsim: %d streams x (fused code; parse + copy kernels + retry pass; 1 in 16 through the streaming reader), both quirk modes, %d status classes, %d handed to the fused kernel by phase one, %d mismatches"
          % (sum(r[3] for r in results), len(classes), sum(r[2] for r in results), len(bad)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
