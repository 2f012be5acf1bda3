"""CPU-side check of the FUSED kernel's 32-lane code (bro_decoder_core.h with BRO_W = 32, exactly what
bro_decode_warp_kernel and bro_decode_resume_kernel compile) against the oracle: tests/warpsim.py runs the 32 lanes as
fibers that meet at the warp intrinsics, in ascending, descending and shuffled lane order, at every alignment of the
compressed stream and of the output slot.  The second half is a race check under the CUDA memory model (ThreadSanitizer
over the same fibers; only __syncwarp orders memory between lanes), including the global-memory buffers that
compute-sanitizer's racecheck does not see.  The race-check and guarded-buffer binaries are also built with
-fsanitize=undefined (no recovery): a misaligned vector access -- a fault on the device -- a shift by 32 or more, or a signed
overflow in the kernels' code ends the run.  The GPU parity tests proper are tests/test_gpu_parity.py."""
import os
import re
import subprocess

import numpy as np
import pytest

import fuzzgen
import warpsim
from conftest import ROOT, corpus_files, stream_vectors
from oracle import oracle

pytestmark = pytest.mark.skipif(not warpsim.available(), reason="the fiber switch of bro_warpsim.cpp is x86-64 only")

CORE = os.path.join(ROOT, "brotli_rs_b200", "csrc", "bro_decoder_core.h")


def _check(stream, label, cap=None, quirks=0, configs=None):
    st, out = oracle.decode(stream, quirks=quirks)
    if cap is None:
        cap = len(out) if st == 0 else len(out) + (1 << 16)
    for latency, order, in_mis, out_mis in configs or [(False, warpsim.ASCENDING, 0, 0)]:
        warpsim.set_alignment(in_mis, out_mis)
        st1, out1 = warpsim.decode(stream, cap=cap, quirks=quirks, latency=latency, order=order, seed=in_mis + 1)
        assert st1 == st and (st != 0 or out1 == out), (label, latency, order, in_mis, out_mis, st, st1)
    warpsim.set_alignment(0, 0)
    return st


def _configs(rng, n):
    """n (build, lane order, stream alignment mod 128, slot alignment mod 16) tuples; both builds and all orders appear"""
    out = []
    for k in range(n):
        out.append((bool(k & 1), warpsim.ORDERS[(k >> 1) % 3], int(rng.integers(128)), int(rng.integers(16))))
    return out


def test_corpus_and_vectors():
    rng = np.random.default_rng(2)
    for name, comp, _ in corpus_files():
        big = len(comp) > 100000
        _check(comp, name, configs=_configs(rng, 2 if big else 6))
    for name, inp, _, _ in stream_vectors():
        if len(inp) > 100000:
            continue
        _check(inp, name, configs=_configs(rng, 6))


def test_every_slot_alignment_of_the_copy_streams():
    """the warp copies (bro_lz_copy / bro_copy_far: ragged head, 16-byte vectors, ragged tail; periodic fills) at all 16
    alignments of the slot, in all three lane orders"""
    names = ("backward65536.compressed", "zeros.compressed", "quickfox_repeated.compressed", "compressed_file.compressed",
             "random_org_10k.bin.compressed", "64x.compressed", "ukkonooa.compressed")
    for name, comp, _ in corpus_files():
        if name in names:
            _check(comp, name, configs=[(False, order, 8 * m + 3, m) for m in range(16) for order in warpsim.ORDERS])


def test_mutation_fuzz():
    corpus = [c for _, c, _ in corpus_files()]
    rng = np.random.default_rng(17)
    seen = set()
    for m in fuzzgen.mutations(corpus, seed=21, count=1500, max_len=52000):
        st, out = oracle.decode(m)
        cap = len(out) if st == 0 else len(out) + (1 << 16)
        if rng.random() < 0.25:
            cap = int(rng.integers(0, len(out) + 100))
        o, ol, sts = oracle.decode_batch(np.frombuffer(m, dtype=np.uint8), np.array([0, len(m)], dtype=np.uint64),
                                         np.array([0, cap], dtype=np.uint64))
        st0, out0 = int(sts[0]), o[: int(ol[0])].tobytes()
        latency, order, in_mis, out_mis = _configs(rng, int(rng.integers(1, 7)))[-1]
        warpsim.set_alignment(in_mis, out_mis)
        st1, out1 = warpsim.decode(m, cap=cap, latency=latency, order=order, seed=len(seen))
        assert st1 == st0 and (st0 != 0 or out1 == out0), (st0, st1, m[:16].hex(), len(m), cap, latency, order, in_mis, out_mis)
        seen.add(st0)
    warpsim.set_alignment(0, 0)
    assert len(seen) >= 15


def test_fresh_streams():
    """every quality band of libbrotli (1: one code of each kind; 5-9: block types; 10-11: context modelling -> the general loop
    with its on-chip tables and, in the latency build, the 10-bit literal root)"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    rng = np.random.default_rng(4)
    k = 0
    for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
        for q, lgwin, size in ((1, 18, 30000), (5, 16, 70000), (9, 10, 20000), (11, 22, 40000), (11, 16, 9000)):
            raw = fuzzgen.synthetic_raw(kind, 100 + k, size)
            k += 1
            comp = fuzzgen.compress(enc, raw, q, lgwin)
            for latency, order, in_mis, out_mis in _configs(rng, 4):
                warpsim.set_alignment(in_mis, out_mis)
                assert warpsim.decode(comp, cap=len(raw), latency=latency, order=order) == (0, raw), (kind, q, lgwin, latency, order)
    # heterogeneous streams: literal block types without context modelling (the lane-parallel chunks of the general loop)
    kinds = ["words", "skewed", "small_alpha", "random", "runs", "repeat2k"]
    for i in range(6):
        raw = bytearray()
        while len(raw) < 90000:
            raw += fuzzgen.synthetic_raw(kinds[int(rng.integers(len(kinds)))], int(rng.integers(1 << 30)), int(rng.integers(2000, 30000)))
        raw = bytes(raw)
        comp = fuzzgen.compress(enc, raw, int(rng.integers(5, 10)), int(rng.integers(16, 23)))
        for latency, order, in_mis, out_mis in _configs(rng, 4):
            warpsim.set_alignment(in_mis, out_mis)
            assert warpsim.decode(comp, cap=len(raw), latency=latency, order=order) == (0, raw), (i, latency, order)
    warpsim.set_alignment(0, 0)


def test_quirk_vectors_and_dictionary_kats():
    """SURVEY appendix D in both quirk modes, and every transform id x word length (tests/dictgen.py) through the 32-lane
    bro_dict_word (reference src/transformation/mod.rs:84-209, src/lib.rs:1506-1540)"""
    import dictgen
    for hx in ("82000000445008122001", "02000000445008122b0106", "02000000445008122a0102", "02000000445008122a0108",
               "e200000044501812a6fb01", "4c8000" + "00" * 257 + "03"):
        for quirks in (0, 1):
            _check(bytes.fromhex(hx), hx, cap=1024, quirks=quirks, configs=[(False, o, 5, 3) for o in warpsim.ORDERS])
    seen = set()
    k = 0
    for quirks in (0, 1):
        for label, s, st, out in dictgen.kat_batch(oracle, quirks, indices_per_length=2):
            k += 1
            warpsim.set_alignment(k % 128, k % 16)
            st1, out1 = warpsim.decode(s, cap=64, quirks=quirks, order=warpsim.ORDERS[k % 3], latency=bool(k & 1))
            assert st1 == st and (st != 0 or out1 == out), (label, quirks, st, st1)
            seen.add(st)
    warpsim.set_alignment(0, 0)
    assert 0 in seen and len(seen) >= 2


def test_stream_resume():
    """the resume kernel's 32-lane code behind the streaming reader's loop: input in pieces, bounded buffers"""
    rng = np.random.default_rng(6)
    for name, comp, _ in corpus_files():
        if len(comp) > 60000:
            continue
        st, out = oracle.decode(comp)
        for chunks, order in (([7], warpsim.DESCENDING), ([int(x) for x in rng.integers(1, 3000, 16)], warpsim.SHUFFLED)):
            if chunks == [7] and len(comp) > 5000:
                chunks = [997]
            st1, served, calls, _, _ = warpsim.stream_decode(comp, chunks, order=order)
            assert st1 == st and (served == out if st == 0 else out.startswith(served)), (name, chunks[:3], st, st1)
    corpus = [c for _, c, _ in corpus_files()]
    for m in fuzzgen.mutations(corpus, seed=31, count=250, max_len=30000):
        st, out = oracle.decode(m)
        chunks = [int(x) for x in rng.integers(1, max(2, len(m)), 4)]
        st1, served, _, _, _ = warpsim.stream_decode(m, chunks, order=warpsim.ORDERS[len(m) % 3])
        assert st1 == st and (served == out if st == 0 else out.startswith(served)), (m[:16].hex(), chunks, st, st1)


# ---- the barriers ----

def _barrier_line(pattern):
    """line number of the first bro_syncwarp() at or after the line matching `pattern` in bro_decoder_core.h"""
    lines = open(CORE).read().split("\n")
    for i, text in enumerate(lines):
        if re.search(pattern, text):
            for j in range(i, len(lines)):
                if "bro_syncwarp();" in lines[j]:
                    return j + 1
    raise AssertionError("pattern not found: " + pattern)


def test_lane_orders_notice_a_missing_barrier():
    """mutation check of the simulation itself: with the bro_syncwarp() between the prefix of a periodic fill and the copy
    that reads it left out (the read-after-write round 1's review found by inspection), one of the lane orders must produce
    wrong bytes -- and with it in place none does (test_every_slot_alignment_of_the_copy_streams)"""
    line = _barrier_line(r"i < n0; i \+= BRO_W\) dst\[i\] = src\[i % dist\]")
    comp = [c for n, c, _ in corpus_files() if n == "backward65536.compressed"][0]
    st, out = oracle.decode(comp)
    wrong = 0
    with warpsim.drop_sync(line):
        for m in range(16):
            for order in warpsim.ORDERS:
                warpsim.set_alignment(0, m)
                try:
                    wrong += warpsim.decode(comp, cap=len(out), order=order) != (st, out)
                except AssertionError:
                    wrong += 1
    warpsim.set_alignment(0, 0)
    assert wrong >= 1
    assert warpsim.decode(comp, cap=len(out), order=warpsim.DESCENDING) == (st, out)


def _tsan_binary():
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "warpsim_tsan")
    csrc = os.path.join(ROOT, "brotli_rs_b200", "csrc")
    srcs = [os.path.join(csrc, "bro_warpsim.cpp"), os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(csrc, f) for f in ("bro_warpsim.h", "bro_decoder_core.h", "bro_records.h", "bro_status.h", "bro_tables_generated.h")]
    if not (os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps)):
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread,undefined", "-fno-sanitize-recover=undefined", "-DBRO_WARPSIM_MAIN", "-Wno-unknown-pragmas", "-o", exe] + srcs + \
              ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("g++ -fsanitize=thread does not build here: " + r.stderr[-300:])
    probe = subprocess.run([exe], capture_output=True, text=True)
    if "usage" not in probe.stderr:
        pytest.skip("ThreadSanitizer does not start here: " + probe.stderr[-300:])
    return exe


def _tsan_run(exe, files, latency=0, order=0, quirks=0, align=(0, 0), drop=None, batch_threads=None, halt=True):
    """-> (ThreadSanitizer reports, {file: (status, out_len, fnv1a64 hex)}, stderr).  batch_threads: the files as one batch through
    bro_decode_warp_kernel itself with a CTA of that many threads"""
    env = dict(os.environ, BRO_WS_ALIGN="%d,%d" % align, TSAN_OPTIONS="exitcode=66 history_size=4")
    if drop is not None:
        env["BRO_WS_DROP_SYNC"] = str(drop)
        if halt:
            env["TSAN_OPTIONS"] += " halt_on_error=1"     # (a mutation run only has to produce its first report)
    if batch_threads:
        env["BRO_WS_BATCH"] = str(batch_threads)
    r = subprocess.run([exe, str(latency), str(order), "1", str(quirks), "0"] + ["%s:%d" % f for f in files], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode in (0, 66), r.stderr[-2000:]
    res = {}
    for ln in r.stdout.splitlines():
        name, st, n, h, err = ln.rsplit(" ", 4)
        assert err == "0", ln
        res[name] = (int(st), int(n), h)
    return r.stderr.count("WARNING: ThreadSanitizer"), res, r.stderr


def _fnv(b):
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xffffffffffffffff
    return "%016x" % h


def test_no_data_race_between_lanes(tmp_path):
    """the corpus, context-modelled and block-typed fresh streams and mutated streams under the race detector, both builds, several
    alignments: no byte is stored by one lane and touched by another without a __syncwarp in between -- shared-memory scratch, table
    arena and output slot alike -- and a left-out barrier IS reported (the detector detects)"""
    exe = _tsan_binary()
    files, expect = [], {}

    def add(name, comp):
        st, out = oracle.decode(comp)
        p = str(tmp_path / name)
        open(p, "wb").write(comp)
        files.append((p, len(out)))
        expect[p] = (st, out)

    for name, comp, _ in corpus_files():
        if len(comp) <= 170000:
            add(name, comp)
    corpus = [c for _, c, _ in corpus_files()]
    for i, m in enumerate(fuzzgen.mutations(corpus, seed=77, count=60, max_len=30000)):
        add("mut%03d" % i, m)
    enc = fuzzgen.libbrotli_enc()
    if enc is not None:
        k = 0
        for kind in ("words", "skewed", "runs", "repeat2k"):
            for q, lgwin, size in ((5, 16, 60000), (9, 18, 40000), (11, 16, 30000), (10, 22, 30000)):
                k += 1
                add("fresh%02d" % k, fuzzgen.compress(enc, fuzzgen.synthetic_raw(kind, 900 + k, size), q, lgwin))
    for latency, order, align in ((0, 0, (0, 0)), (1, 1, (77, 5))):
        races, res, err = _tsan_run(exe, files, latency=latency, order=order, align=align)
        assert races == 0, err[:6000]
        for p, (st, out) in expect.items():
            st1, n1, h1 = res[p]
            assert st1 == st and (st != 0 or (n1, h1) == (len(out), _fnv(out))), (os.path.basename(p), latency, order, align, st, st1)
    # the same streams as ONE batch through bro_decode_warp_kernel itself, eight warps taking streams from the work queue side by side
    # (each warp its own scratch block in shared memory and its own table arena): no report, same results
    races, res, err = _tsan_run(exe, files, latency=1, order=2, batch_threads=256)
    assert races == 0, err[:6000]
    for p, (st, out) in expect.items():
        st1, n1, h1 = res[p]
        assert st1 == st and (st != 0 or (n1, h1) == (len(out), _fnv(out))), (os.path.basename(p), "batch", st, st1)
    # the detector detects: the barrier in front of pass 2 of the table build (its stores to the per-length positions are read by
    # other lanes), and the one behind the staging of the general loop's on-chip tables
    small = [f for f in files if os.path.basename(f[0]) in ("alice29.txt.compressed", "10x10y.compressed", "fresh03", "fresh04")]
    for pattern in (r"T\[BRO_T_MAXDEPTH \+ 1u\] = 0;", r"if \(hot\.modes\) for \(uint32_t i = lane; i < nl"):
        races, _, err = _tsan_run(exe, small, drop=_barrier_line(pattern))
        assert races >= 1, pattern


# ---- the copy kernel (phase two of the two-phase path): bro_kernels_copy.cu itself, run by a simulated warp ----

def _two_phase_check(streams, label, shapes=(0, 1), aligns=((0, 0), (3, 5), (9, 15)), queue_seed=0):
    import hostsim
    exp = [oracle.decode(s) for s in streams]
    caps = [len(out) for _, out in exp]
    handed, k = 0, 0
    for shape in shapes:
        for in_mis, out_mis in aligns:
            order = warpsim.ORDERS[k % 3]
            k += 1
            res, stats = warpsim.two_phase(streams, caps, shape=shape, order=order, seed=k, in_mis=in_mis, out_mis=out_mis, queue_seed=queue_seed)
            for i, ((st, out), (st1, out1)) in enumerate(zip(exp, res)):
                if st1 in hostsim.RETRY:
                    handed += 1
                    continue
                assert st1 == st and (st != 0 or out1 == out), (label, i, shape, order, in_mis, out_mis, st, st1)
    return handed


def test_copy_kernel_corpus_batch():
    """the corpus as ONE batch (completion queue in shuffled order): periodic fills, long records (lane groups of 8), short records
    (segmented copy through shared memory), stored blocks; both instantiations of the kernel, all lane orders, several alignments"""
    streams = [c for _, c, _ in corpus_files()]
    assert _two_phase_check(streams, "corpus", queue_seed=7) == 0
    assert _two_phase_check(streams, "corpus", shapes=(0,), aligns=[(m, m) for m in range(16)]) == 0


def test_copy_kernel_fresh_and_mutated_streams():
    enc = fuzzgen.libbrotli_enc()
    streams = []
    if enc is not None:
        k = 0
        for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
            for q, lgwin, size in ((1, 18, 30000), (5, 16, 70000), (9, 10, 20000), (6, 22, 120000)):
                k += 1
                streams.append(fuzzgen.compress(enc, fuzzgen.synthetic_raw(kind, 300 + k, size), q, lgwin))
        _two_phase_check(streams, "fresh", queue_seed=3)
    corpus = [c for _, c, _ in corpus_files()]
    muts = list(fuzzgen.mutations(corpus, seed=55, count=400, max_len=40000))
    _two_phase_check(muts, "mutated", shapes=(0,), aligns=((5, 11),), queue_seed=9)


def _copy_barrier_line(pattern):
    lines = open(os.path.join(ROOT, "brotli_rs_b200", "csrc", "bro_kernels_copy.cu")).read().split("\n")
    for i, text in enumerate(lines):
        if re.search(pattern, text):
            for j in range(i, len(lines)):
                if "__syncwarp();" in lines[j]:
                    return j + 1
    raise AssertionError("pattern not found: " + pattern)


def test_copy_kernel_no_data_race_between_lanes(tmp_path):
    """the copy kernel under the race detector: between the groups of records a warp executes one after the other, only the
    __syncwarp at the top of the group loop orders the stores of one group before the loads of the next -- in GLOBAL memory, where
    compute-sanitizer's racecheck does not look.  No report with the barriers in place; each of the two barriers of the product's
    path, left out, is reported (the one between groups is NOT noticed by the lane orders alone: the shuffles around it act as
    barriers in a sequential simulation, which is exactly why the check is done under the memory model)"""
    build = os.path.join(ROOT, "tests", "_build")
    exe = os.path.join(build, "warpsim_copy_tsan")
    csrc = os.path.join(ROOT, "brotli_rs_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in warpsim.COPY_SOURCES] + [os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(csrc, f) for f in warpsim.COPY_DEPS]
    if not (os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps)):
        os.makedirs(build, exist_ok=True)
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread,undefined", "-fno-sanitize-recover=undefined", "-DBRO_WARPSIM_MAIN", "-Wno-unknown-pragmas", "-o", exe] + srcs +
                           ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("g++ -fsanitize=thread does not build here: " + r.stderr[-300:])
    if "usage" not in subprocess.run([exe], capture_output=True, text=True).stderr:
        pytest.skip("ThreadSanitizer does not start here")
    files, expect = [], {}

    def add(name, comp):
        st, out = oracle.decode(comp)
        p = str(tmp_path / name)
        open(p, "wb").write(comp)
        files.append("%s:%d" % (p, len(out)))
        expect[p] = (st, out)

    for name, comp, _ in corpus_files():
        add(name, comp)
    enc = fuzzgen.libbrotli_enc()
    if enc is not None:
        k = 0
        for kind in ("skewed", "repeat2k", "runs", "words"):
            for q, lgwin, size in ((1, 18, 30000), (5, 16, 70000), (9, 22, 90000)):
                k += 1
                add("fresh%02d" % k, fuzzgen.compress(enc, fuzzgen.synthetic_raw(kind, 700 + k, size), q, lgwin))

    def run(shape, order, align, drop=None, queue_seed=0):
        env = dict(os.environ, BRO_WS_ALIGN="%d,%d" % align,
                   TSAN_OPTIONS="exitcode=66 suppressions=" + os.path.join(ROOT, "tests", "warpsim_tsan.supp"))
        if drop is not None:
            env["BRO_WS_DROP_SYNC"] = str(drop)
            env["TSAN_OPTIONS"] += " halt_on_error=1"
        r = subprocess.run([exe, str(shape), str(order), "1", str(queue_seed)] + files, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode in (0, 66), r.stderr[-2000:]
        return r.stderr.count("WARNING: ThreadSanitizer"), r

    import hostsim
    for shape, order, align, qs in ((0, 0, (0, 0), 0), (1, 2, (3, 5), 4)):
        races, r = run(shape, order, align, queue_seed=qs)
        assert races == 0, r.stderr[:6000]
        for ln in r.stdout.splitlines():
            name, st1, n1, h1, err = ln.rsplit(" ", 4)
            st, out = expect[name]
            assert err == "0" and (int(st1) in hostsim.RETRY or (int(st1) == st and (st != 0 or (int(n1), h1) == (len(out), _fnv(out))))), ln
    all_files = list(files)
    del files[:]
    files.extend(f for f in all_files if any(k in f for k in ("64x", "backward65536", "quickfox", "alice29", "zeros", "monkey", "ukkonooa")))
    for pattern in (r"stores of earlier groups are visible", r"done \+= m;"):
        races, _ = run(0, 0, (3, 5), drop=_copy_barrier_line(pattern))
        assert races >= 1, pattern


# ---- the parse kernel (phase one) and both kernels of the two-phase path back to back ----

def _kernels_check(streams, label, configs, quirks=0):
    import hostsim
    exp = [oracle.decode(s, quirks=quirks) for s in streams]
    caps = [len(out) for _, out in exp]
    handed = 0
    for k, (lanes, order, in_mis, out_mis, hand_out) in enumerate(configs):
        ho = None if hand_out is None else hand_out(len(streams))
        res, retry, queue = warpsim.two_phase_kernels(streams, caps, quirks=quirks, lanes=lanes, hand_out=ho, order=order, seed=k + 1, in_mis=in_mis,
                                                      out_mis=out_mis, copy_shape=k & 1, copy_order=warpsim.ORDERS[(k + 1) % 3])
        assert sorted(queue) == list(range(len(streams)))
        for i, ((st, out), (st1, out1)) in enumerate(zip(exp, res)):
            if st1 in hostsim.RETRY:
                handed += 1
                continue
            assert st1 == st and (st != 0 or out1 == out), (label, i, lanes, order, in_mis, out_mis, st, st1)
    return handed


def test_parse_and_copy_kernels_corpus_batch():
    """the corpus as one batch through the parse kernel (32 streams to a warp: boundary protocol, hand-out, completion queue, lanes
    in immediate mode next to lanes that write records, 65,537 empty meta-blocks next to text) and then the copy kernel, which takes
    the streams in the order the parse kernel finished them"""
    streams = [c for _, c, _ in corpus_files()]
    rev = lambda n: list(range(n - 1, -1, -1))
    assert _kernels_check(streams, "corpus", [(32, warpsim.ASCENDING, 0, 0, None), (32, warpsim.DESCENDING, 5, 9, rev),
                                              (32, warpsim.SHUFFLED, 13, 3, None), (7, warpsim.SHUFFLED, 2, 15, rev),
                                              (1, warpsim.ASCENDING, 1, 1, None)]) == 0


def test_parse_and_copy_kernels_fresh_mutated_and_sizing():
    import hostsim
    rng = np.random.default_rng(12)
    corpus = [c for _, c, _ in corpus_files()]
    muts = list(fuzzgen.mutations(corpus, seed=91, count=600, max_len=30000))
    seen = set(oracle.decode(m)[0] for m in muts)
    assert len(seen) >= 15
    _kernels_check(muts, "mutated", [(32, warpsim.SHUFFLED, 7, 11, None), (32, warpsim.DESCENDING, 0, 4, None)])
    enc = fuzzgen.libbrotli_enc()
    if enc is not None:
        streams, k = [], 0
        for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
            for q, lgwin, size in ((1, 18, 30000), (5, 16, 70000), (9, 10, 20000), (11, 22, 40000), (10, 16, 9000), (6, 22, 120000)):
                k += 1
                streams.append(fuzzgen.compress(enc, fuzzgen.synthetic_raw(kind, 1300 + k, size), q, lgwin))
        _kernels_check(streams, "fresh", [(32, warpsim.ASCENDING, 3, 6, None), (32, warpsim.SHUFFLED, 12, 1, None), (4, warpsim.DESCENDING, 0, 0, None)])
    # quirk vectors and dictionary words in spec mode
    import dictgen
    for quirks in (0, 1):
        kats = [s for _, s, _, _ in dictgen.kat_batch(oracle, quirks, indices_per_length=1)]
        _kernels_check(kats, "kats", [(32, warpsim.SHUFFLED, 1, 2, None)], quirks=quirks)
    # sizing mode (bro_batch_sizes): nothing is written, every stream's decoded size is reported
    streams = [c for _, c, _ in corpus_files()]
    res, retry, _ = warpsim.two_phase_kernels(streams, [0] * len(streams), sizing=True, order=warpsim.SHUFFLED)
    for s, (st1, size) in zip(streams, res):
        st, out = oracle.decode(s)
        assert st1 in hostsim.RETRY or (st1 == st and (st != 0 or size == len(out))), (st, st1, size, len(out))


def test_parse_kernel_ring_wait_is_exact():
    """the compressed words reach a lane through asynchronous copies into a ring; the simulation lets every copy land as late as the
    kernel's cp.async.wait_group allows.  With the kernel's own count the batch decodes; let one more group stay in flight than the
    kernel asks for and the window takes a word that has not arrived -- the batch must come out wrong (the simulation would notice
    a wait that is one group short)"""
    import hostsim
    streams = [c for n, c, _ in corpus_files() if n in ("alice29.txt.compressed", "asyoulik.txt.compressed", "ukkonooa.compressed")]
    exp = [oracle.decode(s) for s in streams]
    caps = [len(o) for _, o in exp]
    res, _, _ = warpsim.two_phase_kernels(streams, caps)
    assert [r for r in res] == exp
    L = warpsim._lib_parse()
    L.bro_warpsim_parse_ring_slack(1)
    try:
        res, _, _ = warpsim.two_phase_kernels(streams, caps)
    finally:
        L.bro_warpsim_parse_ring_slack(0)
    assert [r for r in res] != exp


def test_fused_code_stays_inside_the_granules_of_its_buffers(tmp_path):
    """how far beyond what a caller provides does the fused kernel's code READ (or write)?  The compressed stream and the output slot
    are placed so that the first byte behind the 16-byte granule that holds their last byte -- or the byte in front of their first
    -- lies in an unmapped page: the corpus decodes without touching it.  (include/brotli_b200.h asks for 16 readable bytes behind
    the compressed batch; aligned granules that hold a wanted byte are all the code reads, which no real allocation ends inside.)"""
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "warpsim_guard")
    csrc = os.path.join(ROOT, "brotli_rs_b200", "csrc")
    srcs = [os.path.join(csrc, "bro_warpsim.cpp"), os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(csrc, f) for f in ("bro_warpsim.h", "bro_decoder_core.h", "bro_records.h", "bro_status.h", "bro_tables_generated.h")]
    if not (os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-DBRO_WARPSIM_MAIN",
                               "-Wno-unknown-pragmas", "-o", exe] + srcs +
                              ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")])
    files, expect = [], {}
    for name, comp, _ in corpus_files():
        st, out = oracle.decode(comp)
        p = str(tmp_path / name)
        open(p, "wb").write(comp)
        files.append("%s:%d" % (p, len(out)))
        expect[p] = (st, out)
    for k, guard in enumerate(("back,16,0", "back,1,1", "back,15,15", "front")):
        r = subprocess.run([exe, str(k & 1), str(k % 3), "1", "0", "0"] + files, env=dict(os.environ, BRO_WS_GUARD=guard), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, (guard, r.returncode, r.stderr[-300:])
        for ln in r.stdout.splitlines():
            name, st1, n1, h1, err = ln.rsplit(" ", 4)
            st, out = expect[name]
            assert err == "0" and int(st1) == st and (st != 0 or (int(n1), h1) == (len(out), _fnv(out))), (guard, ln)


def test_two_phase_path_with_its_retry_pass():
    """the product's whole two-phase call on the CPU: parse kernel, copy kernel, then bro_decode_warp_kernel (bro_kernels.cu itself:
    work queue, retry mode) over the same buffers for the streams phase one handed over -- every status final and equal to the
    oracle's, slots too small included, and streams that were not handed over untouched by the third launch"""
    import hostsim
    rng = np.random.default_rng(23)
    corpus = [c for _, c, _ in corpus_files()]
    enc = fuzzgen.libbrotli_enc()
    fresh = []
    if enc is not None:
        # 4-symbol data: copies so short that they outnumber the stream's share of the record arena (-> RecordsFull); heterogeneous
        # payloads at quality 9-11: more prefix codes than a thread's arena holds (-> ArenaTooSmall)
        for k, (kind, q, lgwin, size) in enumerate((("small_alpha", 5, 16, 30000), ("small_alpha", 9, 18, 20000), ("small_alpha", 7, 22, 40000),
                                                    ("words", 11, 22, 60000), ("skewed", 10, 16, 50000))):
            fresh.append(fuzzgen.compress(enc, fuzzgen.synthetic_raw(kind, 2000 + k, size), q, lgwin))
    muts = list(fuzzgen.mutations(corpus + fresh * 4, seed=123, count=900, max_len=30000)) + fresh
    handed = 0
    for part in range(0, len(muts), 300):
        streams = muts[part: part + 300]
        caps, exp = [], []
        for s in streams:
            st, out = oracle.decode(s)
            cap = len(out) if rng.random() < 0.6 else int(rng.integers(0, len(out) + 64))
            o, ol, sts = oracle.decode_batch(np.frombuffer(s, dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64), np.array([0, cap], dtype=np.uint64))
            caps.append(cap)
            exp.append((int(sts[0]), o[: int(ol[0])].tobytes()))
        res, retry, _ = warpsim.two_phase_kernels(streams, caps, order=warpsim.ORDERS[(part // 300) % 3], in_mis=part % 16, out_mis=(part // 7) % 16,
                                                  retry_pass=True, retry_latency=bool(part & 256))
        handed += retry
        for i, (e, r) in enumerate(zip(exp, res)):
            assert r[0] == e[0] and (e[0] != 0 or r[1] == e[1]), (part + i, e[0], r[0], streams[i][:16].hex())
    assert enc is None or handed >= 3, handed


def test_parse_kernel_lanes_never_meet(tmp_path):
    """the parse kernel under the race detector: 32 unrelated streams to a warp, and nothing one lane touches is touched by another
    -- its stream, slot, records, table arena, and its block of shared memory, which is interleaved WORD BY WORD with the blocks
    of the other 31 lanes (bro_decoder_core.h, BroTl).  A lane whose shared-memory accesses land one word further on (in its
    neighbour's words) is reported."""
    import hostsim
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "warpsim_parse_tsan")
    csrc = os.path.join(ROOT, "brotli_rs_b200", "csrc")
    srcs = [os.path.join(csrc, "bro_warpsim_parse.cpp"), os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(csrc, f) for f in warpsim.PARSE_DEPS]
    if not (os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps)):
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread,undefined", "-fno-sanitize-recover=undefined", "-DBRO_WARPSIM_MAIN", "-Wno-unknown-pragmas", "-o", exe] + srcs +
                           ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("g++ -fsanitize=thread does not build here: " + r.stderr[-300:])
    if "usage" not in subprocess.run([exe], capture_output=True, text=True).stderr:
        pytest.skip("ThreadSanitizer does not start here")
    files, expect = [], {}
    streams = [(n, c) for n, c, _ in corpus_files()]
    corpus = [c for _, c in streams]
    streams += [("mut%03d" % i, m) for i, m in enumerate(fuzzgen.mutations(corpus, seed=201, count=150, max_len=30000))]
    for name, comp in streams:
        st, out = oracle.decode(comp)
        p = str(tmp_path / name)
        open(p, "wb").write(comp)
        files.append("%s:%d" % (p, len(out)))
        expect[p] = (st, len(out))
    for lanes, order, threads in ((32, 0, 32), (32, 2, 64), (6, 1, 32)):         # (64 threads: two warps side by side, one insert/copy table)
        r = subprocess.run([exe, str(lanes), str(order), "1"] + files, env=dict(os.environ, TSAN_OPTIONS="exitcode=66", BRO_WS_THREADS=str(threads)),
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:5000]
        for ln in r.stdout.splitlines():
            name, st1, n1, nrec, err = ln.rsplit(" ", 4)
            st, size = expect[name]
            assert err == "0" and (int(st1) in hostsim.RETRY or (int(st1) == st and (st != 0 or int(n1) == size))), ln
    r = subprocess.run([exe, "32", "0", "1"] + files[:20], env=dict(os.environ, TSAN_OPTIONS="exitcode=66", BRO_WS_SMEM_SKEW="5"),
                       capture_output=True, text=True, timeout=900)
    assert "WARNING: ThreadSanitizer" in r.stderr


def test_eight_warps_side_by_side_and_the_ordering_kernels():
    """every launch of the two-phase call with a CTA of eight warps (256 threads, run interleaved thread by thread): the warps compete
    for the streams of the work queues -- parse kernel, copy kernel, the fused kernel's retry pass -- and interleave their entries
    in the completion queue; the streams are handed out in the order the size-class ordering kernels produce, which must be a
    permutation that puts larger classes first, with the batch's longest stream and AUTO's verdict in the gate"""
    rng = np.random.default_rng(77)
    corpus = [c for _, c, _ in corpus_files()]
    streams = corpus + list(fuzzgen.mutations(corpus, seed=311, count=500, max_len=20000))
    in_off = np.zeros(len(streams) + 1, dtype=np.uint64)
    in_off[1:] = np.cumsum([len(s) for s in streams])
    lens = np.diff(in_off).astype(np.int64)

    def size_class(l):
        b = int(l).bit_length() - 1 if l else 0
        sub = (int(l) >> (b - 3)) & 7 if b >= 3 else 0
        return 255 - (b * 8 + sub)

    for order in warpsim.ORDERS:
        ho, gate = warpsim.order_streams(in_off, order=order, seed=5)
        assert sorted(ho.tolist()) == list(range(len(streams)))
        classes = [size_class(lens[i]) for i in ho]
        assert classes == sorted(classes)
        assert int(gate[0]) == int(lens.max()) and int(gate[1]) == int(int(lens.max()) * 4500 > int(in_off[-1]))
    # bro_batch_sizes' last launch: the internal hand-over codes (105..107) become SizeUnknown (103), nothing else changes
    st = np.array(list(range(0, 25)) + [100, 101, 102, 103, 104, 105, 106, 107] * 40, dtype=np.int32)
    want = np.where((st >= 105) & (st <= 107), 103, st)
    assert (warpsim.sizes_finish(st) == want).all()
    # a batch of equal streams is not bound by its longest one
    eq = np.arange(0, 6001, dtype=np.uint64) * np.uint64(3900)
    ho, gate = warpsim.order_streams(eq)
    assert sorted(ho.tolist()) == list(range(6000)) and int(gate[1]) == 0 and int(gate[0]) == 3900
    exp, caps = [], []
    for s in streams:
        st, out = oracle.decode(s)
        cap = len(out) if rng.random() < 0.7 else int(rng.integers(0, len(out) + 64))
        o, ol, sts = oracle.decode_batch(np.frombuffer(s, dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64), np.array([0, cap], dtype=np.uint64))
        caps.append(cap)
        exp.append((int(sts[0]), o[: int(ol[0])].tobytes()))
    ho, _ = warpsim.order_streams(in_off)
    for k, (order, threads) in enumerate(((warpsim.SHUFFLED, 256), (warpsim.DESCENDING, 128), (warpsim.ASCENDING, 64))):
        res, _, queue = warpsim.two_phase_kernels(streams, caps, hand_out=ho.tolist(), order=order, seed=k + 3, in_mis=3 * k, out_mis=5 * k,
                                                  copy_shape=k & 1, retry_pass=True, threads=threads)
        assert sorted(queue) == list(range(len(streams)))
        for i, (e, r) in enumerate(zip(exp, res)):
            assert r[0] == e[0] and (e[0] != 0 or r[1] == e[1]), (threads, i, e[0], r[0])


def test_kernels_under_address_sanitizer(tmp_path):
    """memory safety of the kernels' code on valid and mutated streams (the CPU twin of compute-sanitizer memcheck over
    tools/fuzz_gpu.py): the fused kernel as a batch of eight warps, the parse kernel, and the copy kernel behind phase one, built with
    -fsanitize=address,undefined -- table arenas, scratch blocks, shared-memory arrays, records, streams and slots all have red zones"""
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    csrc = os.path.join(ROOT, "brotli_rs_b200", "csrc")
    blob = os.path.join(ROOT, "oracle", "dict_blob.c")
    targets = (("warpsim_asan", ["bro_warpsim.cpp"], ("bro_warpsim.h", "bro_kernels.cu", "bro_kernels_resume.cu", "bro_decoder_core.h")),
               ("warpsim_parse_asan", ["bro_warpsim_parse.cpp"], warpsim.PARSE_DEPS),
               ("warpsim_copy_asan", list(warpsim.COPY_SOURCES), warpsim.COPY_DEPS))
    exes = {}
    for name, sources, dep_names in targets:
        exe = os.path.join(build, name)
        srcs = [os.path.join(csrc, f) for f in sources] + [blob]
        deps = srcs + [os.path.join(csrc, f) for f in dep_names]
        if not (os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps)):
            r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-DBRO_WARPSIM_MAIN",
                                "-Wno-unknown-pragmas", "-o", exe] + srcs + ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")],
                               capture_output=True, text=True)
            if r.returncode != 0:
                pytest.skip("g++ -fsanitize=address does not build here: " + r.stderr[-300:])
        exes[name] = exe
    import hostsim
    rng = np.random.default_rng(5)
    files, expect = [], {}
    streams = [(n, c) for n, c, _ in corpus_files() if len(c) <= 170000]
    corpus = [c for _, c in streams]
    streams += [("mut%03d" % i, m) for i, m in enumerate(fuzzgen.mutations(corpus, seed=401, count=250, max_len=30000))]
    for name, comp in streams:
        st, out = oracle.decode(comp)
        cap = len(out) if rng.random() < 0.7 else int(rng.integers(0, len(out) + 64))
        o, ol, sts = oracle.decode_batch(np.frombuffer(comp, dtype=np.uint8), np.array([0, len(comp)], dtype=np.uint64), np.array([0, cap], dtype=np.uint64))
        p = str(tmp_path / name)
        open(p, "wb").write(comp)
        files.append("%s:%d" % (p, cap))
        expect[p] = (int(sts[0]), o[: int(ol[0])].tobytes())
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:exitcode=55")
    runs = (([exes["warpsim_asan"], "1", "2", "1", "0", "0"], dict(env, BRO_WS_BATCH="256"), True),
            ([exes["warpsim_asan"], "0", "1", "3", "0", "0"], dict(env, BRO_WS_ALIGN="101,7"), True),
            ([exes["warpsim_parse_asan"], "32", "2", "1"], dict(env, BRO_WS_THREADS="64"), False),
            ([exes["warpsim_copy_asan"], "0", "2", "1", "5"], dict(env, BRO_WS_ALIGN="3,5"), True))
    for cmd, e, has_bytes in runs:
        r = subprocess.run(cmd + files, env=e, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, (cmd[0], r.returncode, r.stderr[:4000])
        for ln in r.stdout.splitlines():
            name, st1, n1, tail, err = ln.rsplit(" ", 4)
            st, out = expect[name]
            assert err == "0", ln
            if int(st1) in hostsim.RETRY:
                continue
            assert int(st1) == st, (cmd[0], ln, st)
            if st == 0:
                assert int(n1) == len(out) and (not has_bytes or tail == _fnv(out)), (cmd[0], ln)
