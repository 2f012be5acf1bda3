"""Single-command streams that are ONE static-dictionary reference, for every (word length, word
index, transform id): the known-answer batch for bro_dict_word / bro_parse_dict_* on the GPU.

Layout (bits LSB first): WBITS=16, ISLAST=1, MNIBBLES=4, MLEN, NBLTYPES 1/1/1, NPOSTFIX=0, NDIRECT=0, context mode 0,
NTREESL=NTREESD=1, three NSYM=1 simple codes (literal 'A'; one insert&copy symbol = insert 0 / copy `length`, explicit
distance; one distance symbol), then the copy-length extra bits and the distance extra bits.  With no output yet the
maximal back-reference distance is 0, so distance d addresses dictionary word id d - 1 (src/lib.rs:1506-1540)."""
import os

import numpy as np

NDBITS = [0, 0, 0, 0, 10, 10, 11, 11, 10, 10, 10, 10, 10, 9, 9, 8, 7, 7, 8, 7, 7, 6, 6, 5, 5]      # src/dictionary/mod.rs:1-11
COPY_BASE = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54, 70, 102, 134, 198, 326, 582, 1094, 2118]
COPY_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24]


class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, nbits):
        assert 0 <= value < (1 << nbits) or nbits == 0
        self.v |= value << self.n
        self.n += nbits

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def dictionary():
    here = os.path.dirname(os.path.abspath(__file__))
    return open(os.path.join(here, "..", "brotli_rs_b200", "data", "dictionary.bin"), "rb").read()


def word_offset(length):
    return sum(l << NDBITS[l] for l in range(4, length))


def base_word(length, index, dic=None):
    dic = dic or dictionary()
    o = word_offset(length) + index * length
    return dic[o: o + length]


def dict_ref_stream(length, index, transform, mlen):
    """-> the stream (bytes) whose single command emits transform(word[length][index]); mlen = expected output bytes"""
    assert 4 <= length <= 24 and 0 <= index < (1 << NDBITS[length]) and 1 <= mlen <= 65536
    b = Bits()
    b.put(0, 1)                       # WBITS = 16
    b.put(1, 1); b.put(0, 1)          # ISLAST, not empty
    b.put(0, 2); b.put(mlen - 1, 16)  # MNIBBLES = 4, MLEN - 1
    b.put(0, 1); b.put(0, 1); b.put(0, 1)      # NBLTYPES L, I, D = 1
    b.put(0, 2); b.put(0, 4)          # NPOSTFIX, NDIRECT
    b.put(0, 2)                       # context mode of literal block type 0
    b.put(0, 1); b.put(0, 1)          # NTREESL, NTREESD = 1
    b.put(1, 2); b.put(0, 2); b.put(65, 8)     # literal code: simple, NSYM = 1, 'A'
    code = max(c for c in range(24) if COPY_BASE[c] <= length)
    sym = (128 + code) if code < 8 else (192 + code - 8)     # insert code 0, explicit distance
    b.put(1, 2); b.put(0, 2); b.put(sym, 10)   # insert&copy code: simple, NSYM = 1
    dist = index + (transform << NDBITS[length]) + 1
    t = next(t for t in range(48) if ((2 + (t & 1)) << (1 + (t >> 1))) - 4 + 1 <= dist <= ((2 + (t & 1)) << (1 + (t >> 1))) - 4 + (1 << (1 + (t >> 1))))
    nbits = 1 + (t >> 1)
    b.put(1, 2); b.put(0, 2); b.put(16 + t, 6)  # distance code: simple, NSYM = 1 (alphabet 64: 6 bits)
    b.put(length - COPY_BASE[code], COPY_EXTRA[code])         # the command: copy-length extra bits ...
    b.put(dist - (((2 + (t & 1)) << nbits) - 4) - 1, nbits)   # ... and distance extra bits
    return b.bytes()


def kat_batch(oracle, quirks=0, indices_per_length=3, seed=5):
    """-> list of (label, stream, expected status, expected bytes): every transform id x every word length x a few word
    indices (index 0, the last one and random ones); MLEN is the transformed length the oracle's transform() gives, so the
    expected outcome is OK unless the reference panics (uppercase_first on a 0x00 byte) or the quirk changes the length."""
    rng = np.random.default_rng(seed)
    dic = dictionary()
    out = []
    for length in range(4, 25):
        nwords = 1 << NDBITS[length]
        idxs = sorted({0, nwords - 1} | {int(x) for x in rng.integers(0, nwords, max(0, indices_per_length - 2))})
        for index in idxs:
            w = base_word(length, index, dic)
            for tid in range(121):
                exp = oracle.transform(tid, w, quirks)
                mlen = len(exp) if exp else max(1, length)
                s = dict_ref_stream(length, index, tid, max(1, mlen))
                st, o = oracle.decode(s, quirks)
                out.append(((length, index, tid), s, st, o))
    return out
