"""CPU-side check of the warp decoder's LOGIC (brotli_rs_b200/csrc/bro_decoder_core.h compiled with a 1-lane
warp, tests/hostsim.py) against the oracle.  The GPU parity tests proper are tests/test_gpu_parity.py."""
import numpy as np
import pytest

import fuzzgen
import hostsim
from conftest import corpus_files, stream_vectors
from oracle import oracle


def _oracle_slot(stream, cap):
    o, ol, st = oracle.decode_batch(np.frombuffer(stream, dtype=np.uint8), np.array([0, len(stream)], dtype=np.uint64),
                                    np.array([0, cap], dtype=np.uint64))
    return int(st[0]), o[: int(ol[0])].tobytes()


def test_corpus_and_vectors():
    for name, comp, _ in corpus_files():
        st, out = oracle.decode(comp)
        assert hostsim.decode(comp, cap=len(out)) == ((st, out) if st == 0 else (st, hostsim.decode(comp, cap=len(out))[1])), name
    for name, inp, _, _ in stream_vectors():
        st, out = oracle.decode(inp)
        st1, out1 = hostsim.decode(inp, cap=len(out) + 64)
        assert st1 == st and (st != 0 or out1 == out), name


def test_mutation_fuzz():
    corpus = [c for _, c, _ in corpus_files()]
    rng = np.random.default_rng(99)
    seen = set()
    for m in fuzzgen.mutations(corpus, seed=5, count=4000):
        st, out = oracle.decode(m)
        cap = len(out) if st == 0 else len(out) + (1 << 20)
        if rng.random() < 0.25:
            cap = int(rng.integers(0, len(out) + 100))
        st0, out0 = _oracle_slot(m, cap)
        st1, out1 = hostsim.decode(m, cap=cap)
        assert st1 == st0 and (st0 != 0 or out1 == out0), (st0, st1, m[:16].hex(), len(m), cap)
        seen.add(st0)
    assert len(seen) >= 15


def test_fresh_streams():
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    k = 0
    for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
        for q, lgwin, size in ((1, 18, 30000), (5, 16, 70000), (9, 10, 20000), (11, 22, 40000), (11, 16, 9000)):
            raw = fuzzgen.synthetic_raw(kind, 100 + k, size)
            k += 1
            comp = fuzzgen.compress(enc, raw, q, lgwin)
            assert hostsim.decode(comp, cap=len(raw)) == (0, raw), (kind, q, lgwin)


def test_quirk_vectors():
    """SURVEY appendix D (reference mode and spec mode agree with the oracle's)."""
    for hx in ("82000000445008122001", "02000000445008122b0106", "02000000445008122a0102", "02000000445008122a0108",
               "e200000044501812a6fb01", "4c8000" + "00" * 257 + "03"):
        s = bytes.fromhex(hx)
        for quirks in (0, 1):
            st, out = oracle.decode(s, quirks=quirks)
            st1, out1 = hostsim.decode(s, cap=1024, quirks=quirks)
            assert st1 == st and (st != 0 or out1 == out), (hx, quirks, st, st1)


def test_dictionary_transform_kats():
    """every transform id x every word length x a few word indices as one-command streams (tests/dictgen.py), both quirk
    modes: the fused loops' bro_dict_word and phase one's bro_parse_dict_* against the oracle (reference
    src/transformation/mod.rs:84-209, src/lib.rs:1506-1540)"""
    import dictgen
    seen = set()
    for quirks in (0, 1):
        for label, s, st, out in dictgen.kat_batch(oracle, quirks, indices_per_length=2):
            st1, out1 = hostsim.decode(s, cap=64, quirks=quirks)
            assert st1 == st and (st != 0 or out1 == out), (label, quirks, st, st1)
            st2, out2, _, _ = hostsim.parse_decode(s, cap=64, quirks=quirks)
            assert st2 == st and (st != 0 or out2 == out), (label, quirks, st, st2)
            seen.add(st)
    assert seen == {0, 24, 102}


# ---- the two-phase path: phase one (bro_parse.h, the parse kernel's per-lane code) + a byte loop over its copy records ----

def _parse_check(stream, cap, want_st, want_out, label, mis=None):
    st1, out1, nrec, steps = hostsim.parse_decode(stream, cap=cap, mis=mis)
    if st1 in hostsim.RETRY:
        return False           # the product re-runs such a stream with the fused warp kernel
    assert st1 == want_st and (want_st != 0 or out1 == want_out), (label, want_st, st1, nrec, steps)
    return True


def test_parse_corpus_and_vectors():
    handled = 0
    for name, comp, _ in corpus_files():
        st, out = oracle.decode(comp)
        handled += _parse_check(comp, len(out), st, out, name)
    for name, inp, _, _ in stream_vectors():
        st, out = oracle.decode(inp)
        handled += _parse_check(inp, len(out) + 64, st, out, name)
    assert handled >= 40


def test_parse_mutation_fuzz():
    corpus = [c for _, c, _ in corpus_files()]
    rng = np.random.default_rng(98)
    seen, handled = set(), 0
    for m in fuzzgen.mutations(corpus, seed=6, count=4000):
        st, out = oracle.decode(m)
        cap = len(out) if st == 0 else len(out) + (1 << 20)
        if rng.random() < 0.25:
            cap = int(rng.integers(0, len(out) + 100))
        st0, out0 = _oracle_slot(m, cap)
        if _parse_check(m, cap, st0, out0, (m[:16].hex(), len(m), cap)):
            handled += 1
            seen.add(st0)
    assert handled >= 1500 and len(seen) >= 12


def test_parse_fresh_streams():
    """libbrotli streams of every quality: qualities <= 9 (no literal context modelling) must be handled by the
    two-phase path itself, block switches and distance context maps included"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    k = 0
    for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
        for q, lgwin, size in ((1, 18, 30000), (3, 22, 200000), (5, 16, 70000), (7, 22, 300000), (9, 10, 20000),
                               (9, 22, 400000), (11, 22, 40000), (11, 16, 9000)):
            raw = fuzzgen.synthetic_raw(kind, 300 + k, size)
            k += 1
            comp = fuzzgen.compress(enc, raw, q, lgwin)
            st, out, nrec, steps = hostsim.parse_decode(comp, cap=len(raw))
            if q <= 9 and kind != "words":
                # very short copies (4-symbol data) can outnumber the stream's share of the record arena
                assert st == 0 or (st == hostsim.RECORDS_FULL and kind == "small_alpha"), (kind, q, lgwin, st)
            assert st in hostsim.RETRY or (st, out) == (0, raw), (kind, q, lgwin, st)
    # text at quality 5..9: block switches for commands and distances, distance context map, no literal contexts
    import os
    from conftest import DATA
    for name in ("alice29.txt", "plrabn12.txt"):
        raw = open(os.path.join(DATA, name), "rb").read()
        for q in (5, 9):
            st, out, nrec, steps = hostsim.parse_decode(fuzzgen.compress(enc, raw, q, 22), cap=len(raw))
            assert st in hostsim.RETRY or (st, out) == (0, raw), (name, q, st)


def test_parse_immediate_mode():
    """literal context modelling (libbrotli quality >= 10) inside phase one: the stream's thread executes its copies itself
    (no records) and reads the two context bytes back -- text at quality 10 / 11, and mutations of those streams against the
    oracle, status by status"""
    import os
    from conftest import DATA
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    streams = []
    for name, q, lgwin in (("alice29.txt", 11, 16), ("alice29.txt", 10, 22), ("asyoulik.txt", 11, 22), ("lcet10.txt", 11, 18)):
        raw = open(os.path.join(DATA, name), "rb").read()[:120000]
        comp = fuzzgen.compress(enc, raw, q, lgwin)
        st, out, nrec, steps = hostsim.parse_decode(comp, cap=len(raw))
        assert (st, out, nrec) == (0, raw, 0), (name, q, lgwin, st, nrec)
        streams.append(comp)
        # a slot that is too small, by one byte and by a lot
        for cap in (len(raw) - 1, len(raw) // 3):
            st0, out0 = _oracle_slot(comp, cap)
            st1, out1, _, _ = hostsim.parse_decode(comp, cap=cap)
            assert st1 == st0, (name, cap, st0, st1)
    rng = np.random.default_rng(97)
    seen, handled = set(), 0
    for m in fuzzgen.mutations(streams, seed=8, count=1200, max_len=60000):
        st, out = oracle.decode(m)
        cap = len(out) if st == 0 else len(out) + (1 << 20)
        if rng.random() < 0.25:
            cap = int(rng.integers(0, len(out) + 100))
        st0, out0 = _oracle_slot(m, cap)
        if _parse_check(m, cap, st0, out0, (m[:16].hex(), len(m), cap)):
            handled += 1
            seen.add(st0)
    assert handled >= 900 and len(seen) >= 8, (handled, sorted(seen))


def test_parse_record_arena_overflow():
    """a stream with more copies than its share of the record arena is handed to the fused kernel"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    raw = fuzzgen.synthetic_raw("runs", 7, 60000)
    comp = fuzzgen.compress(enc, raw, 5, 16)
    st, out, nrec, _ = hostsim.parse_decode(comp, cap=len(raw))
    assert (st, out) == (0, raw) and nrec > 8
    st, _, _, _ = hostsim.parse_decode(comp, cap=len(raw), rec_cap=8)
    assert st == hostsim.RECORDS_FULL


def test_parse_piece_geometry():
    """every copy record is a piece of at most 32 aligned 16-byte vectors (+ < 16 ragged bytes at each end) unless it is
    a periodic fill; records are in stream order, do not overlap, stay inside the slot and, together with the literals,
    tile the output"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    seen_fill = seen_split = seen_stored = 0
    for kind, q, lgwin, size in (("repeat2k", 5, 16, 262144), ("runs", 5, 16, 90000), ("random", 5, 16, 10000), ("words", 5, 18, 50000),
                                 ("small_alpha", 9, 22, 30000), ("skewed", 1, 16, 40000)):
        raw = fuzzgen.synthetic_raw(kind, 900, size)
        st, recs, out_mis, out = hostsim.parse_records(fuzzgen.compress(enc, raw, q, lgwin), cap=len(raw))
        if st in hostsim.RETRY:
            continue
        assert (st, out) == (0, raw), kind
        end = 0
        for dst, ln, rk, a in recs:
            assert dst >= end and ln >= 1 and dst + ln <= len(raw), (kind, dst, ln)
            end = dst + ln
            fill = rk == 0 and a < ln and a < 16 * 32 + 32
            if fill:
                seen_fill += 1
                continue
            head = (16 - ((dst + out_mis) & 15)) & 15
            head = min(head, ln)
            assert (ln - head) >> 4 <= 32, (kind, dst, ln, a)
            seen_stored += rk == 1
            seen_split += ln >= 512
            if rk == 0:
                assert a <= dst      # a back-reference never reaches before the stream (max_allowed = min(window, pos))
    assert seen_fill and seen_split and seen_stored


def test_parse_sizing_mode():
    """bro_batch_sizes' mode of phase one: the decoded size (or the stream's error) without writing a byte"""
    corpus = [c for _, c, _ in corpus_files()]
    known = 0
    for m in list(fuzzgen.mutations(corpus, seed=12, count=1500)) + corpus:
        st, out = oracle.decode(m)
        st1, size = hostsim.parse_size(m)
        if st1 in hostsim.RETRY:
            continue
        known += 1
        assert st1 == st and (st != 0 or size == len(out)), (st, st1, size, len(out), m[:12].hex())
    assert known >= 600


@pytest.mark.parametrize("group", [308, 208, 203, 32, 16, 8, 4, 108, 116, 132])
def test_copy_phase_lane_code(group):
    """phase two as the copy kernel executes it: 32 records at a time, groups of independent records, long records piece
    by piece through the kernel's own lane code (bro_copy_piece.h) with `group` lanes per piece, all loads of a step
    before its first store (group > 100: the staged form, issue / consume through slots with two steps in flight; group >
    200: the bulk form, whole pieces fetched into group - 200 slots, then consumed by 32 lanes each; 308: the product's window
    form -- the last 4 KiB of output in a ring, sources taken from the ring where they lie in it); short groups last record
    first.  Bytes must equal the oracle's."""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    pieces = shorts = fills = handled = 0
    with hostsim.copy_group(group) as cg:
        for name, comp, _ in corpus_files():
            st, out = oracle.decode(comp)
            if st == 0:
                handled += _parse_check(comp, len(out), st, out, name, mis=len(comp) % 16)
                p, s, f = cg.stats()
                pieces, shorts, fills = pieces + p, shorts + s, fills + f
        k = 0
        for kind, q, lgwin, size in (("repeat2k", 5, 16, 262144), ("repeat2k", 1, 18, 150001), ("runs", 5, 16, 90000),
                                     ("random", 5, 16, 10000), ("random", 5, 16, 70001), ("words", 5, 18, 50000),
                                     ("small_alpha", 9, 22, 30000), ("skewed", 1, 16, 40000), ("repeat2k", 9, 22, 400000)):
            for seed in range(3):
                raw = fuzzgen.synthetic_raw(kind, 700 + k, size + 13 * seed)
                k += 1
                comp = fuzzgen.compress(enc, raw, q, lgwin)
                # destination alignments 0..15: the slot starts at an address with (address & 15) == mis
                handled += _parse_check(comp, len(raw), 0, raw, (kind, q, lgwin, seed), mis=(5 * k + 3) % 16)
                p, s, f = cg.stats()
                pieces, shorts, fills = pieces + p, shorts + s, fills + f
    assert handled >= 40 and pieces >= 100 and shorts >= 100 and fills >= 10, (handled, pieces, shorts, fills)
    if group == 308:
        hits, seen = cg.win_stats()
        assert hits >= 300 and seen > hits, (hits, seen)      # the ring served sources, and not all of them


# ---- the resumable decode behind the streaming reader (bro_decode_stream_resume + the reader's loop, tests/hostsim.py) ----

def _stream_check(stream, chunks, label, out_cap=1 << 16):
    st, out = oracle.decode(stream)
    st1, served, calls, max_in, max_out = hostsim.stream_decode(stream, chunks, out_cap=out_cap)
    assert st1 == st, (label, chunks[:4], st, st1)
    if st == 0:
        assert served == out, (label, chunks[:4], len(served), len(out))
    else:
        # what was served before the error are whole meta-blocks the reference decoded too
        assert out.startswith(served), (label, chunks[:4], st, len(served), len(out))
    return calls, max_in, max_out


def test_stream_resume_corpus():
    """input in pieces of 1 byte .. the whole stream: status and bytes equal the oracle's for every corpus file and test
    vector, with the input and output buffers bounded by one meta-block (+ window), not by the stream"""
    rng = np.random.default_rng(3)
    for name, comp, _ in corpus_files():
        for chunks in ([1], [7], [4096], [len(comp) + 1], [int(x) for x in rng.integers(1, 3000, 16)]):
            if chunks == [1] and len(comp) > 60000:
                continue
            _stream_check(comp, chunks, name)
    for name, inp, _, _ in stream_vectors():
        _stream_check(inp, [3], name)
        _stream_check(inp, [len(inp) + 1], name)
    # bounded memory: 106 meta-blocks, 405,808 -> 912,868 bytes, window 2^22: neither buffer ever holds the stream
    comp = [c for n, c, _ in corpus_files() if n == "metablock_reset.compressed"][0]
    calls, max_in, max_out = _stream_check(comp, [2048], "metablock_reset", out_cap=1 << 14)
    assert max_in < len(comp) // 2 and max_out <= 1 << 18 and calls >= 20, (calls, max_in, max_out)
    # 65,537 empty meta-blocks: progress without output
    comp = [c for n, c, _ in corpus_files() if n == "empty.compressed.18"][0]
    calls, max_in, max_out = _stream_check(comp, [1000], "empty.18")
    assert max_in <= 2100, max_in


def test_stream_resume_fuzz():
    corpus = [c for _, c, _ in corpus_files()]
    rng = np.random.default_rng(41)
    seen = set()
    for m in fuzzgen.mutations(corpus, seed=8, count=2500):
        chunks = [int(x) for x in rng.integers(1, max(2, len(m)), 5)] if rng.random() < 0.7 else [int(rng.integers(1, 64))]
        if chunks == [1] and len(m) > 20000:
            chunks = [977]
        _stream_check(m, chunks, m[:16].hex())
        seen.add(oracle.decode(m)[0])
    assert len(seen) >= 15


def test_stream_resume_fresh_streams():
    """streams with small windows (the history slides many times) and many meta-blocks, all qualities"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    k = 0
    for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
        for q, lgwin, size in ((1, 10, 60000), (5, 16, 300000), (9, 10, 50000), (11, 12, 40000), (5, 22, 500000)):
            raw = fuzzgen.synthetic_raw(kind, 500 + k, size)
            k += 1
            comp = fuzzgen.compress(enc, raw, q, lgwin)
            st, served, calls, max_in, max_out = hostsim.stream_decode(comp, [1500], out_cap=1 << 12)
            assert (st, served) == (0, raw), (kind, q, lgwin, st, len(served))


def test_largest_window_far_reference():
    """WBITS = 24, back-references 15,700,000 bytes back (the window's limit is 16 MiB - 16): fused logic, phase one +
    the copy kernel's lane code, and the streaming loop (whose history is then a whole window)"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    raw = fuzzgen.far_reference_raw()
    comp = fuzzgen.compress(enc, raw, 9, 24)
    assert oracle.decode(comp) == (0, raw)
    assert hostsim.decode(comp, cap=len(raw)) == (0, raw)
    st, recs, _, out = hostsim.parse_records(comp, cap=len(raw))
    assert (st, out) == (0, raw) and max(a for _, _, k, a in recs if k == 0) == 15_700_000
    with hostsim.copy_group(8):
        st, out, _, _ = hostsim.parse_decode(comp, cap=len(raw), mis=9)
    assert (st, out) == (0, raw)
    st, served, calls, max_in, max_out = hostsim.stream_decode(comp, [1 << 16], out_cap=1 << 20)
    assert (st, served) == (0, raw) and max_out <= 1 << 25


# ---- runs of metadata meta-blocks (src/lib.rs:1617-1683): the byte-at-a-time fast path of bro_next_metablock ----

def _metadata_stream(rng, n_blocks, tail=b"\x03", corrupt=None, desync=False):
    """WBITS 16 + an empty metadata block (one byte, 0x0c: the stream is byte aligned behind it), then n_blocks metadata blocks on
    byte boundaries and the empty last meta-block (0x03).  A metadata block: ISLAST 0, MNIBBLES 3, reserved 0, MSKIPBYTES k (six
    bits), then AT ONCE k bytes of skip length minus one, then zero fill bits up to the byte boundary, then the bytes skipped:
    0x06 for k = 0; 16 bits 0x16 | (L - 1) << 6 for k = 1; 24 bits 0x26 | (L - 1) << 6 for k = 2 (the reference combines the two
    length bytes its own way, quirk Q1: whatever it makes of them, both decoders must agree).
    corrupt = (block index, byte): that block's first byte is replaced.  desync: the skip length is written byte aligned (as a
    first version of the fast path read it -- wrongly): the decoder then skips too little or too much and takes the random bytes that
    follow for meta-block headers, which is the best test of those there is."""
    out = bytearray(b"\x0c")
    plain = corrupt is None and not desync          # -> the reference decodes it to nothing, cleanly
    for k in range(n_blocks):
        kind = int(rng.integers(0, 6))
        if corrupt is not None and corrupt[0] == k:
            out.append(corrupt[1])
            continue
        if kind <= 1:
            out.append(0x06)
        elif kind <= 4:
            skip = int(rng.integers(1, 6))                      # 1..5 bytes skipped: the fast path takes up to two, the general route the rest
            hdr = bytes([0x16, skip - 1]) if desync else (0x16 | ((skip - 1) << 6)).to_bytes(2, "little")
            out += hdr + rng.integers(0, 256, skip, dtype=np.uint8).tobytes()
        else:
            skip = int(rng.integers(257, 700))
            plain = False                                       # (quirk Q1: the reference skips another number of bytes)
            hdr = bytes([0x26, (skip - 1) & 0xff, (skip - 1) >> 8]) if desync else (0x26 | ((skip - 1) << 6)).to_bytes(3, "little")
            out += hdr + rng.integers(0, 256, skip, dtype=np.uint8).tobytes()
    return bytes(out) + tail, plain


def test_metadata_block_runs():
    """status (and the empty output) of streams that are runs of metadata blocks -- intact, truncated, with bytes behind their end,
    with a corrupt header byte, and desynchronised -- equal the oracle's on the fused logic and on phase one of the two-phase path"""
    rng = np.random.default_rng(8)
    seen = set()
    ok_intact = 0
    for trial in range(600):
        n = int(rng.integers(1, 40))
        corrupt = None
        r = int(rng.integers(0, 4))
        desync = trial % 2 == 1
        if r == 0:
            corrupt = (int(rng.integers(0, n)), int(rng.choice([0x0e, 0x46, 0x86, 0x07, 0x36, 0x17, 0x00, 0xff])))
        s, plain = _metadata_stream(rng, n, corrupt=corrupt, desync=desync)
        if r == 1:
            s = s[: int(rng.integers(1, len(s)))]                # truncated
        elif r == 2:
            s = s + rng.integers(0, 256, int(rng.integers(1, 4)), dtype=np.uint8).tobytes()     # bytes behind the end of the stream
        st, out = oracle.decode(s)
        seen.add(st)
        if r == 3 and plain:
            assert (st, out) == (0, b""), (trial, st)
            ok_intact += 1
        st0, out0 = hostsim.decode(s, cap=64)
        assert st0 == st and (st != 0 or out0 == out), (trial, st, st0, s.hex())
        st1, out1, _, _ = hostsim.parse_decode(s, cap=64)
        if st1 in hostsim.RETRY:
            continue
        assert st1 == st and (st != 0 or out1 == out), (trial, st, st1, s.hex())
    assert ok_intact >= 10 and 0 in seen and len(seen) >= 6, (ok_intact, seen)
