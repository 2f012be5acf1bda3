"""Pin the oracle against every golden vector the reference holds for the decode path."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import FREWSXCV_STATUS, GOLDEN, corpus_files, stream_vectors
from oracle import oracle


@pytest.mark.parametrize("name,comp,exp", corpus_files(), ids=[c[0] for c in corpus_files()])
def test_corpus(name, comp, exp):
    st, out = oracle.decode(comp)
    if exp is None:
        assert st == FREWSXCV_STATUS[name]
    else:
        assert st == 0 and out == exp


@pytest.mark.parametrize("name,inp,exp,err", stream_vectors(), ids=[v[0] for v in stream_vectors()])
def test_stream_vectors(name, inp, exp, err):
    """reference tests/lib.rs:4-605 and doc-test src/lib.rs:361-376"""
    st, out = oracle.decode(inp)
    if err is not None:
        assert st != 0 and err in oracle.description(st)
    else:
        # the reference's positive tests discard the Result and compare bytes only
        assert out == exp
        assert st == 0


def test_transform_kats():
    """reference src/transformation/mod.rs:215-1301"""
    kats = json.load(open(os.path.join(GOLDEN, "transform_kats.json")))
    assert sorted(k["id"] for k in kats) == list(range(121))
    for k in kats:
        assert oracle.transform(k["id"], bytes.fromhex(k["base_hex"])) == bytes.fromhex(k["expect_hex"]), k["name"]


def test_bitreader_kats():
    """reference src/bitreader/mod.rs:339-560"""
    kats = json.load(open(os.path.join(GOLDEN, "bitreader_kats.json")))
    assert len(kats) == 13
    for k in kats:
        res = oracle.bitreader_script(bytes.fromhex(k["data_hex"]), [(o[0], o[1]) for o in k["ops"]])
        for (op, arg, exp), got in zip(k["ops"], res):
            if exp is None:
                continue
            if op == "string":
                assert got == bytes.fromhex(exp), k["name"]
            else:
                assert got == exp, (k["name"], op, got, exp)


def test_imtf_kats():
    """reference tests/lib.rs:607-673 (the test file carries its own MTF; restated here)"""
    def mtf(v):
        alphabet = list(range(256))
        out = []
        for value in v:
            index = alphabet.index(value)
            alphabet.insert(0, alphabet.pop(index))
            out.append(index)
        return bytes(out)
    for k in json.load(open(os.path.join(GOLDEN, "imtf_kats.json"))):
        v = bytes(k["vector"])
        if k["name"] == "should_not_change":
            assert oracle.imtf(v) == v
        else:
            assert mtf(oracle.imtf(v)) == v


def test_tree_kats():
    """reference src/huffman/tree/mod.rs:96-212: codes are consumed first-bit-read = most significant; restated
    through canonical code lengths: lengths (1,2,2) give codes 0,10,11 for symbols 0,1,2."""
    # bits in read order (LSB first within the byte): 0 | 1,0 | 1,1 | 0  -> symbols 0,1,2,0
    data = bytes([0b0_11_01_0])
    assert oracle.tree_decode([1, 2, 2], data, 4) == [0, 1, 2, 0]
    # depth-2 complete code: 00,01,10,11
    data = bytes([0b11_01_10_00])  # read order: 0,0 | 0,1 | 1,0 | 1,1
    assert oracle.tree_decode([2, 2, 2, 2], data, 4) == [0, 1, 2, 3]
    # a single non-zero length decodes with zero bits (tree.len == 1, src/huffman/tree/mod.rs:87-91)
    assert oracle.tree_decode([0, 3, 0], b"", 3) == [1, 1, 1]


def test_status_strings():
    """reference src/lib.rs:331-354 (typos are part of the observable payload)"""
    assert oracle.description(13) == "Enocuntered non-zero fill bit"
    assert oracle.description(23) == "Run length excceeded declared length of context map"
    assert oracle.description(24) == "Encountered unexpected EOF"


def _libbrotli():
    try:
        dec = ctypes.CDLL("libbrotlidec.so.1")
        enc = ctypes.CDLL("libbrotlienc.so.1")
    except OSError:
        return None, None
    dec.BrotliDecoderDecompress.restype = ctypes.c_int
    dec.BrotliDecoderDecompress.argtypes = [ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
    enc.BrotliEncoderCompress.restype = ctypes.c_int
    enc.BrotliEncoderCompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p,
                                          ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
    return dec, enc


def _compress(enc, raw, q, lgwin):
    cap = len(raw) + (len(raw) >> 2) + 1024
    buf = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(cap)
    assert enc.BrotliEncoderCompress(q, lgwin, 0, len(raw), raw, ctypes.byref(n), buf) == 1
    return buf.raw[:n.value]


def test_crosscheck_libbrotli_roundtrip():
    """Secondary (non-reference) check: fresh streams from the system libbrotlienc decode to their input."""
    dec, enc = _libbrotli()
    if enc is None:
        pytest.skip("system libbrotli not present")
    rng = np.random.default_rng(7)
    text = open(os.path.join(GOLDEN, "data", "alice29.txt"), "rb").read()
    cases = []
    for q in (0, 1, 2, 4, 5, 6, 9, 10, 11):
        for lgwin in (10, 16, 22):
            cases.append((text[: 20000 + 3000 * q], q, lgwin))
    block = rng.integers(0, 256, 2048, dtype=np.uint8).tobytes()
    rep = bytearray(block * 64)
    for i in range(0, len(rep), 997):
        rep[i] = (rep[i] + 1) & 0xFF
    cases += [(bytes(rep), 5, 16), (bytes(rep), 11, 16), (rng.integers(0, 256, 10000, dtype=np.uint8).tobytes(), 5, 16),
              (bytes(rng.integers(0, 4, 50000, dtype=np.uint8)), 9, 18), (b"", 5, 16), (b"a", 11, 22),
              (np.minimum(255, rng.exponential(40, 10000)).astype(np.uint8).tobytes(), 5, 16)]
    for raw, q, lgwin in cases:
        comp = _compress(enc, raw, q, lgwin)
        st, out = oracle.decode(comp)
        assert st == 0 and out == raw, (q, lgwin, len(raw))


def test_quirk_vectors():
    """the 'parity unpinned' corners follow the reference's source text by default and the
    specification with quirks=1."""
    assert oracle.decode(bytes.fromhex("82000000445008122001")) == (0, b"time ")
    for hx in ("02000000445008122b0106", "02000000445008122a0102", "02000000445008122a0108"):
        assert oracle.decode(bytes.fromhex(hx)) == (0, b"e")          # Q3: OmitFirstN keeps the last byte
        assert oracle.decode(bytes.fromhex(hx), quirks=1)[0] != 0      # spec: empty word, stream then fails
    st, _ = oracle.decode(bytes.fromhex("e200000044501812a6fb01"))
    assert st == oracle.PANIC_UPPERCASE_ZERO                          # Q4: reference panics
    assert oracle.decode(bytes.fromhex("e200000044501812a6fb01"), quirks=1) == (0, b"\0" * 8)
    q1 = bytes.fromhex("4c8000") + b"\0" * 257 + b"\x03"    # 261 B
    assert oracle.decode(q1)[0] == 12                                 # Q1: MSKIPLEN assembled with << i
    assert oracle.decode(q1, quirks=1) == (0, b"")


def test_batch_matches_single():
    files = corpus_files()
    in_buf = np.frombuffer(b"".join(c for _, c, _ in files), dtype=np.uint8)
    in_off = np.cumsum([0] + [len(c) for _, c, _ in files]).astype(np.uint64)
    caps = [len(e) if e is not None else 70000 for _, _, e in files]
    out_off = np.cumsum([0] + caps).astype(np.uint64)
    for nthreads in (1, 4):
        out, out_len, status = oracle.decode_batch(in_buf, in_off, out_off, nthreads=nthreads)
        for i, (name, comp, exp) in enumerate(files):
            st, ref = oracle.decode(comp)
            assert status[i] == st, name
            if st == 0:
                assert out[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes() == ref
    # a slot that is too small reports OUTPUT_TOO_SMALL and never writes past it
    out_off2 = out_off.copy()
    out, out_len, status = oracle.decode_batch(in_buf[: int(in_off[2])], in_off[:3], np.array([0, 5, 10], dtype=np.uint64))
    assert status[0] == oracle.OUTPUT_TOO_SMALL
