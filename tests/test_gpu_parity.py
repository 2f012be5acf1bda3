"""GPU parity tests: the CUDA decoder, called through the C ABI (libbrotli_b200.so), against the oracle.

Bit-exact bar: identical bytes for every valid stream, identical status class for every stream.  Sizes the oracle
finishes in seconds are compared directly; BASELINE.json's full-size configurations are checked through
replica-equality (every replica's slot equals replica 0's, and replica 0 equals the oracle)."""
import os

import numpy as np
import pytest

import fuzzgen
from conftest import DATA, FREWSXCV_STATUS, corpus_files, stream_vectors
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["warp", "twophase"])
def dec(request):
    """Every parity test runs against both paths: the fused warp-per-stream kernel (small batches, retry pass) and
    the two-phase path (parse kernel, copy kernel, fused retry pass; large batches)."""
    import torch
    assert torch.cuda.is_available()
    from brotli_rs_b200 import BatchDecoder
    d = BatchDecoder(0, mode=BatchDecoder.MODE_WARP if request.param == "warp" else BatchDecoder.MODE_TWOPHASE)
    yield d
    d.close()


def oracle_slots(streams, caps):
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    in_buf, in_off = pack_streams(streams)
    out_off = slot_offsets(caps)
    out, out_len, status = oracle.decode_batch(in_buf, in_off, out_off, nthreads=8)
    return in_buf, in_off, out_off, out, out_len, status


def check_batch(dec, streams, caps, label=""):
    """Decode on the GPU through bro_batch_decode_host and compare with the oracle run on the same slots."""
    in_buf, in_off, out_off, ref, ref_len, ref_st = oracle_slots(streams, caps)
    out, out_len, status = dec.decode_host(in_buf, in_off, out_off)
    bad = np.nonzero(status != ref_st)[0]
    assert len(bad) == 0, (label, "status", [(int(i), int(status[i]), int(ref_st[i]), streams[int(i)][:12].hex()) for i in bad[:5]])
    ok = np.nonzero(ref_st == 0)[0]
    assert (out_len[ok] == ref_len[ok]).all(), (label, "length")
    for i in ok:
        b, n = int(out_off[i]), int(ref_len[i])
        if not np.array_equal(out[b: b + n], ref[b: b + n]):
            k = int(np.nonzero(out[b: b + n] != ref[b: b + n])[0][0])
            raise AssertionError((label, "bytes", int(i), "first diff at", k, "of", n))
    return status


def test_corpus_exact_slots(dec):
    """bit-exact on every file in data/ (BASELINE north_star); slots sized exactly to the expected output."""
    files = corpus_files()
    streams = [c for _, c, _ in files]
    want = [oracle.decode(c) for c in streams]
    st = check_batch(dec, streams, [len(w[1]) for w in want], "corpus")
    for (name, _, exp), s, w in zip(files, st, want):
        if exp is None:
            assert int(s) == FREWSXCV_STATUS[name]
        else:
            assert int(s) == 0 and w[1] == exp


def test_corpus_expected_files(dec):
    """independently of the oracle: GPU output equals the reference's expected-output files"""
    files = [f for f in corpus_files() if f[2] is not None]
    res = dec.decode_streams([c for _, c, _ in files], [len(e) + 32 for _, _, e in files])
    for (name, _, exp), (st, out) in zip(files, res):
        assert st == 0 and out == exp, name


def test_stream_vectors(dec):
    """the reference's tests/lib.rs:4-605 + doc-test, through the batch ABI"""
    from brotli_rs_b200 import status_description
    vecs = stream_vectors()
    res = dec.decode_streams([v[1] for v in vecs], [len(v[2]) + 64 if v[2] is not None else 1 << 17 for v in vecs])
    for (name, _, exp, err), (st, out) in zip(vecs, res):
        if err is not None:
            assert st != 0 and err in status_description(st), name
        else:
            assert st == 0 and out == exp, name


def test_decompressor_read_struct(dec):
    """the Read-struct boundary: Decompressor::new(r).read_to_end (src/lib.rs:361-376 doc-test shape)"""
    from brotli_rs_b200 import BroError, Decompressor
    for name in ("64x", "alice29.txt", "quickfox_repeated", "empty", "metablock_reset"):
        with open(os.path.join(DATA, name + ".compressed"), "rb") as f:
            got = Decompressor(f, decoder=dec).read()
        assert got == open(os.path.join(DATA, name), "rb").read(), name
    # chunked reads fill the caller's buffer completely until the end, then return 0 repeatedly
    r = Decompressor(open(os.path.join(DATA, "alice29.txt.compressed"), "rb").read(), decoder=dec)
    exp = open(os.path.join(DATA, "alice29.txt"), "rb").read()
    chunks = []
    while True:
        b = bytearray(10007)
        n = r.readinto(b)
        if n == 0:
            break
        assert n == 10007 or len(b"".join(chunks)) + n == len(exp)
        chunks.append(bytes(b[:n]))
    assert b"".join(chunks) == exp and r.readinto(bytearray(8)) == 0
    # an invalid stream raises with the reference's description (tests/lib.rs:36-53)
    with pytest.raises(BroError, match="non-zero bit"):
        Decompressor(bytes([0xa1, 0x03]), decoder=dec).read()
    # a reader without a shared context creates its own
    assert Decompressor(bytes([0x06])).read() == b""


def test_mutation_fuzz(dec):
    """stand-in for the reference's AFL workflow: mutated corpus streams, status + bytes vs the oracle, with exact,
    generous and too-small slots; invalid streams must never write outside their slot"""
    corpus = [c for _, c, _ in corpus_files()]
    streams = list(fuzzgen.mutations(corpus, seed=11, count=6000))
    rng = np.random.default_rng(3)
    caps = []
    for s in streams:
        st, out = oracle.decode(s)
        r = rng.random()
        caps.append(len(out) if (st == 0 and r < 0.5) else int(rng.integers(0, len(out) + 100)) if r < 0.7 else len(out) + 4096)
    status = check_batch(dec, streams, caps, "fuzz")
    assert len(set(int(s) for s in status)) >= 15


def test_metadata_block_runs(dec):
    """runs of metadata meta-blocks (src/lib.rs:1617-1683; the byte-at-a-time fast path of bro_next_metablock): intact, truncated,
    with bytes behind their end, with a corrupt header byte and desynchronised streams (tests/test_hostsim_parity.py builds them)"""
    from test_hostsim_parity import _metadata_stream
    rng = np.random.default_rng(18)
    streams = []
    for trial in range(1200):
        n = int(rng.integers(1, 40))
        r = int(rng.integers(0, 4))
        corrupt = (int(rng.integers(0, n)), int(rng.choice([0x0e, 0x46, 0x86, 0x07, 0x36, 0x17, 0x00, 0xff]))) if r == 0 else None
        s, _ = _metadata_stream(rng, n, corrupt=corrupt, desync=trial % 2 == 1)
        if r == 1:
            s = s[: int(rng.integers(1, len(s)))]
        elif r == 2:
            s = s + rng.integers(0, 256, int(rng.integers(1, 4)), dtype=np.uint8).tobytes()
        streams.append(s)
    st = check_batch(dec, streams, [64] * len(streams), "metadata runs")
    assert len(set(int(x) for x in st)) >= 6


def test_slot_guard_bytes(dec):
    """no stream writes past its slot: sentinel bytes between slots survive"""
    import torch
    from brotli_rs_b200.batch import pack_streams
    corpus = [c for _, c, _ in corpus_files()]
    streams = list(fuzzgen.mutations(corpus, seed=21, count=1500)) + corpus
    want = [oracle.decode(s) for s in streams]
    caps = np.array([len(w[1]) if w[0] == 0 else max(0, len(w[1]) - 3) for w in want], dtype=np.uint64)
    guard = 24
    off = np.zeros(len(streams) + 1, dtype=np.uint64)
    # slots of exact capacity followed by a guard gap: out_off[i+1] is the END of slot i, so give every stream
    # its own [start, end) by interleaving dummy zero-length streams that own the guard gaps
    starts = np.cumsum(np.concatenate([[0], caps[:-1] + guard])).astype(np.uint64)
    in_buf, in_off = pack_streams(streams)
    d_in = torch.from_numpy(in_buf.copy()).cuda()
    total = int(starts[-1] + caps[-1] + guard)
    d_out = torch.full((total,), 0xA5, dtype=torch.uint8, device="cuda")
    # decode stream by stream ranges using a 2-entry offset table per stream would be slow; instead build a batch
    # of 2n "streams" where the odd ones are empty inputs with the guard gap as their slot (they fail with EOF and
    # must write nothing)
    n = len(streams)
    in_off2 = np.zeros(2 * n + 1, dtype=np.uint64)
    out_off2 = np.zeros(2 * n + 1, dtype=np.uint64)
    for i in range(n):
        in_off2[2 * i] = in_off[i]
        in_off2[2 * i + 1] = in_off[i + 1]
        out_off2[2 * i] = starts[i]
        out_off2[2 * i + 1] = starts[i] + caps[i]
    in_off2[2 * n] = in_off[n]
    out_off2[2 * n] = total
    d_len, d_st = dec.decode_device(d_in, torch.from_numpy(in_off2.astype(np.int64)).cuda(), d_out,
                                    torch.from_numpy(out_off2.astype(np.int64)).cuda())
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    st = d_st.cpu().numpy()
    ln = d_len.cpu().numpy()
    for i in range(n):
        s, e = int(starts[i]), int(starts[i] + caps[i])
        assert (out[e: e + guard] == 0xA5).all(), ("guard overwritten after stream", i)
        assert int(st[2 * i + 1]) == 24 and int(ln[2 * i + 1]) == 0
        if want[i][0] == 0:
            assert int(st[2 * i]) == 0 and out[s:e].tobytes() == want[i][1]


def test_fresh_streams_all_kinds(dec):
    """fresh streams from the system libbrotlienc over payload kinds, qualities and window sizes"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    raws, streams = [], []
    k = 0
    for kind in ("random", "skewed", "repeat2k", "runs", "words", "small_alpha"):
        for q, lgwin, size in ((0, 22, 50000), (1, 18, 30000), (2, 16, 100000), (5, 16, 262144), (6, 24, 300000), (9, 10, 20000),
                               (10, 16, 30000), (11, 22, 60000), (11, 16, 9000), (5, 16, 1), (5, 16, 0), (4, 12, 5000)):
            raw = fuzzgen.synthetic_raw(kind, 500 + k, size) if size else b""
            k += 1
            raws.append(raw)
            streams.append(fuzzgen.compress(enc, raw, q, lgwin))
    res = dec.decode_streams(streams, [len(r) for r in raws])
    for i, ((st, out), raw) in enumerate(zip(res, raws)):
        assert st == 0 and out == raw, i


def test_unaligned_device_buffers(dec):
    """device API with input and output buffers at odd byte offsets (the bit window and the vector copies must not
    assume alignment)"""
    import torch
    from brotli_rs_b200.batch import pack_streams
    files = [f for f in corpus_files() if f[2] is not None]
    streams = [c for _, c, _ in files]
    exps = [e for _, _, e in files]
    in_buf, in_off = pack_streams(streams)
    for shift_in, shift_out in ((1, 3), (2, 5), (3, 7), (5, 1), (7, 13)):
        d_in = torch.zeros(len(in_buf) + 64, dtype=torch.uint8, device="cuda")
        d_in[shift_in: shift_in + len(in_buf)] = torch.from_numpy(in_buf.copy()).cuda()
        caps = np.array([len(e) + 1 for e in exps], dtype=np.uint64)      # odd slot sizes -> odd slot starts
        out_off = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
        d_out = torch.zeros(int(out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
        d_len, d_st = dec.decode_device(d_in[shift_in:], torch.from_numpy(in_off.astype(np.int64)).cuda(), d_out[shift_out:],
                                        torch.from_numpy(out_off.astype(np.int64)).cuda())
        torch.cuda.synchronize()
        out = d_out.cpu().numpy()[shift_out:]
        assert (d_st.cpu().numpy() == 0).all()
        for i, e in enumerate(exps):
            assert out[int(out_off[i]): int(out_off[i]) + len(e)].tobytes() == e, (files[i][0], shift_in, shift_out)


def _replica_check(dec, streams_unique, replicas, order_seed, caps_unique):
    """Full-size check by replica equality: decode `replicas` copies of each unique stream (shuffled), verify
    replica 0 of each against the oracle and every other replica against replica 0 on the GPU."""
    import torch
    from brotli_rs_b200.batch import slot_offsets
    nu = len(streams_unique)
    want = [oracle.decode(s) for s in streams_unique]
    idx = np.tile(np.arange(nu), replicas)
    np.random.default_rng(order_seed).shuffle(idx)
    lens = np.array([len(s) for s in streams_unique], dtype=np.uint64)
    in_off = np.concatenate([[0], np.cumsum(lens[idx])]).astype(np.uint64)
    ubuf = np.frombuffer(b"".join(streams_unique), dtype=np.uint8)
    uoff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    # gather the replicated input on the GPU
    d_u = torch.from_numpy(ubuf.copy()).cuda()
    pieces = [d_u[uoff[j]: uoff[j + 1]] for j in idx]
    d_in = torch.cat(pieces) if pieces else d_u[:0]
    out_off = slot_offsets(np.asarray(caps_unique, dtype=np.uint64)[idx])
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
    d_len, d_st = dec.decode_device(d_in, torch.from_numpy(in_off.astype(np.int64)).cuda(), d_out,
                                    torch.from_numpy(out_off.astype(np.int64)).cuda())
    torch.cuda.synchronize()
    st, ln = d_st.cpu().numpy(), d_len.cpu().numpy()
    first = {}
    for pos, j in enumerate(idx):
        j = int(j)
        assert int(st[pos]) == want[j][0], (pos, j, int(st[pos]), want[j][0])
        if want[j][0] != 0:
            continue
        assert int(ln[pos]) == len(want[j][1])
        b = int(out_off[pos])
        sl = d_out[b: b + len(want[j][1])]
        if j not in first:
            first[j] = sl
            assert sl.cpu().numpy().tobytes() == want[j][1], ("replica 0 of", j)
        else:
            assert torch.equal(sl, first[j]), ("replica differs", pos, j)
    return int(ln[st == 0].sum())


def test_config_c2_quickfox_repeated_x10k(dec):
    """BASELINE config 2: 10,000 copies of data/quickfox_repeated.compressed (58 B -> 176,128 B each)"""
    s = open(os.path.join(DATA, "quickfox_repeated.compressed"), "rb").read()
    total = _replica_check(dec, [s], 10000, 0, [176128])
    assert total == 10000 * 176128


def test_config_c3_corpus_x1000(dec):
    """BASELINE config 3: every data/*compressed* stream x 1000 replicas, shuffled with default_rng(0)"""
    files = corpus_files()
    streams = [c for _, c, _ in files]
    caps = [len(e) if e is not None else 70000 for _, _, e in files]
    total = _replica_check(dec, streams, 1000, 0, caps)
    assert total == 1000 * 3094120


def test_config_c4_c5_samples(dec):
    """BASELINE configs 4/5 at a size the oracle finishes in seconds: 64 high-ratio WBITS=16 streams (SURVEY C4
    recipe), 64 stored streams and 64 skewed-literal streams (C5, C5b), each replicated 16x"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    streams, caps = [], []
    for i in range(64):
        for kind, size in (("repeat2k", 262144), ("random", 10000), ("skewed", 10000)):
            raw = fuzzgen.synthetic_raw(kind, 1000 + i, size)
            streams.append(fuzzgen.compress(enc, raw, 5, 16))
            caps.append(len(raw))
    total = _replica_check(dec, streams, 16, 1, caps)
    assert total == 16 * 64 * (262144 + 10000 + 10000)


def test_parse_handover_is_retried_by_warp_kernel(dec):
    """streams the parse kernel hands over -- meta-blocks that need more table space than a thread's 64 KiB arena (many
    block types / trees), literal context modelling, more copies than the record share -- are finished by the fused
    kernel's retry pass; the caller never sees the internal statuses"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    parts = [open(os.path.join(DATA, n), "rb").read() for n in ("alice29.txt", "mapsdatazrh", "lcet10.txt", "random_chunks",
                                                                "plrabn12.txt", "asyoulik.txt")]
    rng = np.random.default_rng(0)
    chunks = []
    for k in range(200):
        p = parts[k % len(parts)]
        o = int(rng.integers(0, max(1, len(p) - 3000)))
        chunks.append(p[o: o + 3000])
        chunks.append(bytes(rng.integers(0, 256 >> (k % 7), 2000, dtype=np.uint8)))
    raw = b"".join(chunks)
    big = [fuzzgen.compress(enc, raw, q, lg) for q, lg in ((9, 18), (9, 22), (11, 22), (10, 18))]
    small = [c for _, c, e in corpus_files() if e is not None and len(c) < 60000]
    streams, raws = [], []
    for r in range(6):
        for b in big:
            streams.append(b); raws.append(raw)
        for c in small:
            streams.append(c); raws.append(oracle.decode(c)[1])
    res = dec.decode_streams(streams, [len(r) for r in raws])
    for i, ((st, out), r) in enumerate(zip(res, raws)):
        assert st == 0 and out == r, (i, st)


def test_context_modelled_streams_immediate_mode():
    """libbrotli quality 10 / 11 text (literal context modelling) is decoded by the two-phase path ITSELF: the parse kernel's
    thread executes the stream's copies and reads the context bytes back (immediate mode), nothing is handed to the fused
    kernel; mutations of the same streams against the oracle, status by status"""
    from brotli_rs_b200 import BatchDecoder
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    streams, raws = [], []
    rng = np.random.default_rng(3)
    for name in ("alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt"):
        text = open(os.path.join(DATA, name), "rb").read()
        for q, lgwin in ((11, 16), (10, 22), (11, 22), (11, 18)):
            for size in (3000, 40000, 140000):
                o = int(rng.integers(0, max(1, len(text) - size)))
                raw = text[o: o + size]
                streams.append(fuzzgen.compress(enc, raw, q, lgwin)); raws.append(raw)
    d = BatchDecoder(0, mode=BatchDecoder.MODE_TWOPHASE)
    res = d.decode_streams(streams * 8, [len(r) for r in raws] * 8)
    for i, ((st, out), r) in enumerate(zip(res, raws * 8)):
        assert st == 0 and out == r, (i, st)
    assert d.last_batch_stats()["retried_streams"] == 0
    muts = list(fuzzgen.mutations(streams, seed=12, count=1500, max_len=60000))
    caps = []
    for m in muts:
        st, out = oracle.decode(m)
        caps.append(len(out) if (st == 0 and rng.random() < 0.75) else int(rng.integers(0, len(out) + 100)) if rng.random() < 0.5 else len(out) + 4096)
    status = check_batch(d, muts, caps, "context-modelled mutations")
    assert len(set(int(x) for x in status)) >= 8
    d.close()


def test_auto_mode_launch_counts(monkeypatch):
    """default mode decodes a small batch of small streams with ONE launch of the warp kernel and a large one by the
    two-phase path: 3 ordering kernels, the parse kernel, the copy kernel and the fused kernel's retry pass (per slice of
    the host call: one slice here)"""
    from brotli_rs_b200 import BatchDecoder
    monkeypatch.setenv("BRO_B200_HOST_CHUNKS", "1")
    files = [f for f in corpus_files() if f[2] is not None and len(f[1]) < 2000]
    streams = [c for _, c, _ in files] * 100
    exps = [e for _, _, e in files] * 100
    import torch
    threshold = 160 * torch.cuda.get_device_properties(0).multi_processor_count      # AUTO takes the two-phase path from 160 streams per SM
    many = -(-threshold // len(streams))
    lat_wave = 16 * torch.cuda.get_device_properties(0).multi_processor_count      # resident warps of the fused kernel's latency build
    for mode, reps, want in ((None, 1, 1), (BatchDecoder.MODE_TWOPHASE, 1, 6), (None, many, 6)):
        d = BatchDecoder(0, mode=mode)
        assert len(streams) < threshold <= len(streams) * many
        before = d.launch_count
        res = d.decode_streams(streams * reps, [len(e) for e in exps] * reps)
        # behind the two-phase kernels both builds of the fused kernel are launched when the batch exceeds a latency wave
        # (they decide on the device whose job the pass is)
        assert d.launch_count - before == want + (1 if want == 6 and len(streams) * reps > lat_wave else 0)
        for (st, out), e in zip(res, exps * reps):
            assert st == 0 and out == e
        d.close()


def test_device_path_without_reservation():
    """bro_batch_decode without bro_ctx_reserve sizes the record arena by reading the end offsets back; with a bound
    that is too small the affected streams are decoded by the fused kernel -- same results either way"""
    import ctypes
    import torch
    from brotli_rs_b200 import BatchDecoder
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    raws = [fuzzgen.synthetic_raw("repeat2k", 40 + i, 65536) for i in range(48)]
    streams = [fuzzgen.compress(enc, r, 5, 16) for r in raws]
    in_buf, in_off = pack_streams(streams)
    out_off = slot_offsets([len(r) for r in raws])
    d = BatchDecoder(0, mode=BatchDecoder.MODE_TWOPHASE)
    for bound in (0, 64):
        d_in = torch.from_numpy(in_buf.copy()).cuda()
        d_in_off = torch.from_numpy(in_off.astype(np.int64)).cuda()
        d_out_off = torch.from_numpy(out_off.astype(np.int64)).cuda()
        d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
        d_len = torch.empty(len(streams), dtype=torch.int64, device="cuda")
        d_st = torch.empty(len(streams), dtype=torch.int32, device="cuda")
        d._check(d._lib.bro_ctx_reserve(d._ctx, bound, len(streams)))
        d._check(d._lib.bro_batch_decode(d._ctx, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                                         d_len.data_ptr(), d_st.data_ptr(), len(streams),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        out = d_out.cpu().numpy()
        assert (d_st.cpu().numpy() == 0).all()
        for i, r in enumerate(raws):
            assert out[int(out_off[i]): int(out_off[i]) + len(r)].tobytes() == r, (bound, i)
    d.close()


def test_dictionary_transform_kats_and_quirk_vectors():
    """GPU known answers for the static dictionary (reference src/lib.rs:1506-1540, src/transformation/mod.rs:84-209):
    every transform id 0..120 x every word length 4..24 x word indices {0, last, random} as one-command streams
    (tests/dictgen.py), plus the appendix D vectors for Q1 / Q3 / Q4 -- on BOTH decode paths (the
    fused kernel's 32-lane bro_dict_word, phase one's per-thread emit) in BOTH quirk modes (bro_ctx_set_quirks), against
    the oracle."""
    import dictgen
    from brotli_rs_b200 import BatchDecoder
    appendix_d = [bytes.fromhex(hx) for hx in ("82000000445008122001", "02000000445008122b0106", "02000000445008122a0102",
                                               "02000000445008122a0108", "e200000044501812a6fb01", "4c8000" + "00" * 257 + "03")]
    for quirks in (0, 1):
        kats = dictgen.kat_batch(oracle, quirks, indices_per_length=4)
        streams = [k[1] for k in kats] + appendix_d
        want = [(k[2], k[3]) for k in kats] + [oracle.decode(s, quirks) for s in appendix_d]
        assert {w[0] for w in want} >= ({0, 24, 102} if quirks == 0 else {0, 24})
        for mode in (BatchDecoder.MODE_WARP, BatchDecoder.MODE_TWOPHASE):
            d = BatchDecoder(0, quirks=quirks, mode=mode)
            res = d.decode_streams(streams, [64] * len(streams))
            d.close()
            for i, ((st, out), (wst, wout)) in enumerate(zip(res, want)):
                assert st == wst and (wst != 0 or out == wout), (quirks, mode, kats[i][0] if i < len(kats) else "appendix D %d" % (i - len(kats)), st, wst)


def test_default_context_readers_are_cheap():
    """A reader without a context -- what a drop-in Decompressor::new(r) creates, dozens of them in the reference's own
    test-suite -- uses the process-wide default context: 40 readers must not take 40 x (or even 1 x) the full-grid arenas."""
    import torch
    from brotli_rs_b200 import Decompressor
    files = [(c, e) for _, c, e in corpus_files() if e is not None and len(e) < 200000][:10]
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    readers = []
    for k in range(40):
        c, e = files[k % len(files)]
        r = Decompressor(c, streaming=(k % 2 == 1))
        assert r.read() == e
        readers.append(r)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < (512 << 20), (free0 - free1) >> 20
    for r in readers:
        r.close()


def test_mg_decode_host_matches_oracle():
    """bro_mg_* (one batch over every GPU of the box from one process): same results as the oracle, whatever the number
    of devices; the partition is contiguous and complete."""
    import torch
    from brotli_rs_b200 import MultiGpuDecoder, mg_partition
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    files = corpus_files()
    streams = [c for _, c, _ in files] * 6 + list(fuzzgen.mutations([c for _, c, _ in files], seed=77, count=200))
    free = [oracle.decode(s) for s in streams]
    caps = [len(o) if st == 0 else 4096 for st, o in free]
    in_buf, in_off = pack_streams(streams)
    out_off = slot_offsets(caps)
    # the oracle with the same slots (a stream that outgrows its slot before it fails is OUTPUT_TOO_SMALL on both sides)
    o_out, o_len, o_st = oracle.decode_batch(in_buf, in_off, out_off)
    want = [(int(o_st[i]), o_out[int(out_off[i]): int(out_off[i]) + int(o_len[i])].tobytes()) for i in range(len(streams))]
    mg = MultiGpuDecoder(0)
    assert mg.device_count == torch.cuda.device_count()
    first = mg_partition(in_off, out_off, mg.device_count)
    assert first[0] == 0 and first[-1] == len(streams)
    for _ in range(2):
        out, out_len, status = mg.decode_host(in_buf, in_off, out_off)
        for i, (st, o) in enumerate(want):
            assert int(status[i]) == st, (i, st, int(status[i]))
            if st == 0:
                assert out[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes() == o, i
    mg.close()


def test_copy_kernel_watchdog_is_loud(monkeypatch):
    """The copy kernel waits for the parse kernel through the completion queue.  If a slot is never filled its watchdog
    gives up -- and then the batch must FAIL (BRO_CUDA_ERROR from the host call and from bro_ctx_last_batch_stats), not
    come back with streams marked OK whose copies were never made.  BRO_B200_DEBUG_NO_PARSE=1 forces it: the parse
    kernel is not launched at all."""
    import torch
    from brotli_rs_b200 import BatchDecoder, BroError
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    raws = [fuzzgen.synthetic_raw("repeat2k", 70 + i, 32768) for i in range(64)]
    streams = [fuzzgen.compress(enc, r, 5, 16) for r in raws]
    monkeypatch.setenv("BRO_B200_DEBUG_NO_PARSE", "1")
    d = BatchDecoder(0, mode=BatchDecoder.MODE_TWOPHASE)
    monkeypatch.delenv("BRO_B200_DEBUG_NO_PARSE")
    with pytest.raises(BroError) as ei:
        d.decode_streams(streams, [len(r) for r in raws])
    assert ei.value.status == 101 and "watchdog" in str(ei.value)
    # the device entry point is asynchronous: the fault is reported by the statistics call
    in_buf, in_off = pack_streams(streams)
    out_off = slot_offsets([len(r) for r in raws])
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(in_buf).to(dev)
    d_in_off = torch.from_numpy(in_off.astype(np.int64)).to(dev)
    d_out_off = torch.from_numpy(out_off.astype(np.int64)).to(dev)
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(len(streams), dtype=torch.int64, device=dev)
    d_st = torch.zeros(len(streams), dtype=torch.int32, device=dev)
    d.decode_device(d_in, d_in_off, d_out, d_out_off, d_len, d_st)
    with pytest.raises(BroError) as ei:
        d.last_batch_stats()
    assert ei.value.status == 101
    d.close()
    # a healthy context decodes the same batch
    d = BatchDecoder(0, mode=BatchDecoder.MODE_TWOPHASE)
    for (st, out), r in zip(d.decode_streams(streams, [len(r) for r in raws]), raws):
        assert st == 0 and out == r
    d.close()


def test_decode_without_size_hints():
    """bro_batch_sizes / bro_batch_decode_unsized_host: the caller knows no uncompressed size (the reference's
    Decompressor does not either: src/lib.rs:2183-2189 grows a Vec)"""
    import torch
    from brotli_rs_b200 import BatchDecoder
    from brotli_rs_b200.batch import pack_streams
    files = corpus_files()
    streams = [c for _, c, _ in files] + list(fuzzgen.mutations([c for _, c, _ in files], seed=31, count=400))
    want = [oracle.decode(s) for s in streams]
    d = BatchDecoder(0)
    # measuring alone: exact sizes and final statuses wherever the size is knowable without a decode
    in_buf, in_off = pack_streams(streams)
    d_len, d_st = d.sizes_device(torch.from_numpy(in_buf.copy()).cuda(), torch.from_numpy(in_off.astype(np.int64)).cuda())
    torch.cuda.synchronize()
    ln, st = d_len.cpu().numpy(), d_st.cpu().numpy()
    known = 0
    for i, (wst, wout) in enumerate(want):
        if int(st[i]) == 103:
            continue
        known += 1
        assert int(st[i]) == wst and (wst != 0 or int(ln[i]) == len(wout)), (i, int(st[i]), wst)
    assert known >= len(streams) // 2
    # decoding without hints: bytes and statuses as with exact slots
    res = d.decode_unsized(streams)
    for i, ((gst, gout), (wst, wout)) in enumerate(zip(res, want)):
        assert gst == wst and (wst != 0 or gout == wout), (i, gst, wst, len(gout), len(wout))
    assert d.decode_unsized([]) == []
    d.close()


def test_largest_window_far_reference(dec):
    """WBITS = 24 with back-references 15,700,000 bytes back (the largest window), both paths,
    next to small streams in the same batch"""
    enc = fuzzgen.libbrotli_enc()
    if enc is None:
        pytest.skip("system libbrotlienc not present")
    raw = fuzzgen.far_reference_raw()
    comp = fuzzgen.compress(enc, raw, 9, 24)
    small = open(os.path.join(DATA, "quickfox_repeated.compressed"), "rb").read()
    check_batch(dec, [small, comp, small, comp], [176128, len(raw), 176128, len(raw)], "far reference")


# ---- streaming: the resumable decode and the streaming Read-struct ----

def test_streaming_reader_corpus():
    """Decompressor(r, streaming=chunk): input pulled `chunk` bytes at a time, decoded meta-block by meta-block; bytes and
    error class equal the oracle's for every corpus file, bytes decoded before an error come first"""
    from brotli_rs_b200 import BatchDecoder, BroError, Decompressor
    d = BatchDecoder(0)
    try:
        for name, comp, _ in corpus_files():
            st, out = oracle.decode(comp)
            for chunk in (True, 4096, 61):
                if chunk == 61 and len(comp) > 60000:
                    continue
                r = Decompressor(comp, decoder=d, streaming=chunk)
                got, err = bytearray(), 0
                try:
                    while True:
                        b = bytearray(30011)
                        n = r.readinto(b)
                        if n == 0:
                            break
                        got += b[:n]
                except BroError as e:
                    err = e.status
                assert err == st, (name, chunk, st, err)
                if st == 0:
                    assert bytes(got) == out and r.readinto(bytearray(8)) == 0, (name, chunk, len(got), len(out))
                else:
                    assert out.startswith(bytes(got)), (name, chunk)
                r.close()
        # trailing bytes behind a complete stream: every byte of the stream is delivered, then ExpectedEndOfStream
        comp = open(os.path.join(DATA, "64x.compressed"), "rb").read()
        r = Decompressor(comp + b"\x00" * 300, decoder=d, streaming=len(comp))
        b = bytearray(100)
        assert r.readinto(b) == 64 and bytes(b[:64]) == open(os.path.join(DATA, "64x"), "rb").read()
        with pytest.raises(BroError) as ei:
            r.readinto(b)
        assert ei.value.status == 2
        # a stream cut inside a meta-block: UnexpectedEOF once the input is exhausted
        comp = open(os.path.join(DATA, "alice29.txt.compressed"), "rb").read()
        with pytest.raises(BroError) as ei:
            Decompressor(comp[:30000], decoder=d, streaming=1000).read()
        assert ei.value.status == 24
    finally:
        d.close()


def test_streaming_reader_fuzz_and_small_windows():
    from brotli_rs_b200 import BatchDecoder, BroError, Decompressor
    d = BatchDecoder(0)
    rng = np.random.default_rng(17)
    try:
        corpus = [c for _, c, _ in corpus_files()]
        seen = set()
        for m in fuzzgen.mutations(corpus, seed=21, count=250, max_len=20000):
            st, out = oracle.decode(m)
            got, err = b"", 0
            r = Decompressor(m, decoder=d, streaming=int(rng.integers(1, 3000)))
            msg = ""
            try:
                got = r.read()
            except BroError as e:
                err, msg = e.status, str(e)
            assert err == st and (st != 0 or got == out), (m[:16].hex(), st, err, len(got), len(out), msg)
            seen.add(st)
            r.close()
        assert len(seen) >= 8
        enc = fuzzgen.libbrotli_enc()
        if enc is not None:
            for k, (kind, q, lgwin, size) in enumerate((("repeat2k", 5, 10, 200000), ("words", 11, 12, 60000), ("runs", 9, 10, 90000),
                                                        ("skewed", 5, 16, 300000), ("random", 5, 16, 150000))):
                raw = fuzzgen.synthetic_raw(kind, 800 + k, size)
                comp = fuzzgen.compress(enc, raw, q, lgwin)
                assert Decompressor(comp, decoder=d, streaming=1500).read() == raw, (kind, q, lgwin)
    finally:
        d.close()


def test_batch_decode_resume_device():
    """bro_batch_decode_resume on a batch: (1) from all-zero resume points with the whole input it is bro_batch_decode;
    (2) with the input cut in two, the second call continues from the resume points of the first"""
    import torch
    from brotli_rs_b200 import BatchDecoder
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    d = BatchDecoder(0)
    try:
        files = [(n, c, e) for n, c, e in corpus_files() if e is not None]
        streams = [c for _, c, _ in files]
        want = [e for _, _, e in files]
        n = len(streams)
        dev = torch.device("cuda", 0)
        out_off = slot_offsets([len(e) + 64 for e in want])
        d_out_off = torch.from_numpy(out_off.astype(np.int64)).to(dev)
        d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device=dev)
        rec = BatchDecoder.RESUME_DTYPE.itemsize

        def run(parts, d_resume):
            buf, off = pack_streams(parts)
            d_in = torch.from_numpy(np.concatenate([buf, np.zeros(16, dtype=np.uint8)])).to(dev)
            d_in_off = torch.from_numpy(off.astype(np.int64)).to(dev)
            ln, st = d.decode_resume_device(d_in, d_in_off, d_out, d_out_off, d_resume)
            torch.cuda.synchronize()
            return ln.cpu().numpy(), st.cpu().numpy(), d_resume.cpu().numpy().view(BatchDecoder.RESUME_DTYPE)

        ln, st, ck = run(streams, torch.zeros(n * rec, dtype=torch.uint8, device=dev))
        host = d_out.cpu().numpy()
        for i in range(n):
            assert st[i] == 0 and ln[i] == len(want[i]) and ck["flags"][i] == 7, (files[i][0], st[i], ln[i], ck["flags"][i])
            assert host[int(out_off[i]): int(out_off[i]) + len(want[i])].tobytes() == want[i], files[i][0]
        # two calls: the first sees 60 % of every stream
        d_out.zero_()
        cut = [int(0.6 * len(s)) for s in streams]
        d_resume = torch.zeros(n * rec, dtype=torch.uint8, device=dev)
        ln1, st1, ck1 = run([s[:c] for s, c in zip(streams, cut)], d_resume)
        assert set(st1.tolist()) <= {0, 24}
        rest, ck2 = [], ck1.copy()
        for i in range(n):
            drop = int(ck1["in_bits"][i]) >> 3
            rest.append(streams[i][drop:] if st1[i] != 0 else streams[i][len(streams[i]):])
            ck2["in_bits"][i] &= 7
        todo = [i for i in range(n) if st1[i] != 0]
        assert len(todo) >= 8
        ln2, st2, ck3 = run(rest, torch.from_numpy(ck2.view(np.uint8).copy()).to(dev))
        host = d_out.cpu().numpy()
        for i in todo:
            assert st2[i] == 0 and ln2[i] == len(want[i]), (files[i][0], st2[i], ln2[i], len(want[i]))
            assert host[int(out_off[i]): int(out_off[i]) + len(want[i])].tobytes() == want[i], files[i][0]
    finally:
        d.close()
