"""Seeded stream generators shared by the CPU and GPU parity tests: mutated corpus streams (the stand-in for the
reference's AFL workflow, docs/notes_afl.txt) and fresh streams from the system libbrotlienc."""
import ctypes

import numpy as np


def mutations(corpus, seed, count, max_len=70000):
    """Yield `count` mutated streams: byte flips, bit flips, truncations, splices, insertions."""
    rng = np.random.default_rng(seed)
    small = [c for c in corpus if 0 < len(c) <= max_len]
    for _ in range(count):
        base = bytearray(small[int(rng.integers(len(small)))])
        kind = int(rng.integers(6))
        if kind == 0:      # flip a few bits, biased to the header region
            for _ in range(int(rng.integers(1, 4))):
                lim = len(base) if rng.random() < 0.5 else min(len(base), 64)
                i = int(rng.integers(lim))
                base[i] ^= 1 << int(rng.integers(8))
        elif kind == 1:    # overwrite random bytes
            for _ in range(int(rng.integers(1, 6))):
                base[int(rng.integers(len(base)))] = int(rng.integers(256))
        elif kind == 2:    # truncate
            base = base[: int(rng.integers(len(base) + 1))]
        elif kind == 3:    # splice two streams
            other = small[int(rng.integers(len(small)))]
            i = int(rng.integers(len(base) + 1))
            j = int(rng.integers(len(other) + 1))
            base = base[:i] + bytearray(other[j:])
        elif kind == 4:    # append garbage
            base += bytes(rng.integers(0, 256, int(rng.integers(1, 8)), dtype=np.uint8))
        else:              # pure random header + tail of a real stream
            k = int(rng.integers(1, 12))
            base = bytearray(rng.integers(0, 256, k, dtype=np.uint8).tobytes()) + base[k:]
        yield bytes(base[:max_len])


def libbrotli_enc():
    try:
        enc = ctypes.CDLL("libbrotlienc.so.1")
    except OSError:
        return None
    enc.BrotliEncoderCompress.restype = ctypes.c_int
    enc.BrotliEncoderCompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p,
                                          ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
    return enc


def compress(enc, raw, q, lgwin, mode=0):
    cap = len(raw) + (len(raw) >> 2) + 1024
    buf = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(cap)
    assert enc.BrotliEncoderCompress(q, lgwin, mode, len(raw), bytes(raw), ctypes.byref(n), buf) == 1
    return buf.raw[: n.value]


def synthetic_raw(kind, seed, size):
    """Raw payloads with different command mixes."""
    rng = np.random.default_rng(seed)
    if kind == "random":
        return rng.integers(0, 256, size, dtype=np.uint8).tobytes()
    if kind == "skewed":                      # entropy-coded literals, few matches
        return np.minimum(255, rng.exponential(40, size)).astype(np.uint8).tobytes()
    if kind == "repeat2k":                    # 2 KiB random block repeated with 4 mutations per repetition
        block = rng.integers(0, 256, 2048, dtype=np.uint8)
        reps = max(1, size // 2048)
        a = np.tile(block, reps)
        for r in range(reps):
            idx = rng.integers(0, 2048, 4)
            a[r * 2048 + idx] = rng.integers(0, 256, 4, dtype=np.uint8)
        return a.tobytes()
    if kind == "runs":                        # long runs and short periods (overlapping copies)
        out = bytearray()
        while len(out) < size:
            period = int(rng.integers(1, 70))
            pat = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
            out += pat * int(rng.integers(1, 4000 // period + 2))
        return bytes(out[:size])
    if kind == "words":                       # dictionary-friendly text
        words = [b"the ", b"of ", b"and ", b"time", b"number of different ", b"people ", b"information ", b"\n",
                 b"Government", b" which ", b"because", b"THE ", b"Search", b"http://", b"</div>", b"language"]
        out = bytearray()
        while len(out) < size:
            out += words[int(rng.integers(len(words)))]
        return bytes(out[:size])
    if kind == "small_alpha":
        return bytes(rng.integers(0, 4, size, dtype=np.uint8))
    raise ValueError(kind)
