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
