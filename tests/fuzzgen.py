"""Seeded stream generators shared by the CPU and GPU parity tests: mutated corpus streams (the stand-in for the
reference's AFL workflow, docs/notes_afl.txt) and fresh streams from the system libbrotlienc."""
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


from brotli_rs_b200.workloads import compress, far_reference_raw, libbrotli_enc, synthetic_raw  # noqa: E402,F401
