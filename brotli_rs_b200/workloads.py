"""Synthetic workloads of BASELINE.json / SURVEY.md section 8d (bench and test harness support, not on the decode
path).  Streams are produced with the system libbrotlienc (same image on the GPU box)."""
import ctypes

import numpy as np


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
    if kind == "random":                      # incompressible -> stored meta-block
        return rng.integers(0, 256, size, dtype=np.uint8).tobytes()
    if kind == "skewed":                      # entropy-coded literals, few matches
        return np.minimum(255, rng.exponential(40, size)).astype(np.uint8).tobytes()
    if kind == "repeat2k":                    # 2 KiB random block repeated, 4 single-byte mutations per repetition
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
    if kind == "far_half":                    # round 2, C7: the second half repeats the first, size / 2 bytes back (far beyond
        half = size // 2                      # what L2 keeps per stream), with 4 single-byte mutations per 2 KiB
        a = rng.integers(0, 256, half, dtype=np.uint8)
        b = a.copy()
        idx = rng.integers(0, half, 4 * (half // 2048))
        b[idx] = rng.integers(0, 256, len(idx), dtype=np.uint8)
        return a.tobytes() + b.tobytes()
    raise ValueError(kind)


def far_reference_raw(seed=5, block=200000, gap=15_500_000):
    """A payload whose second half repeats a random block `block + gap` bytes (15.7 MB) back: with lgwin = 24 (window
    16 MiB - 16, the format's maximum) libbrotli at quality 9 encodes it as back-references at distance 15,700,000."""
    blk = np.random.default_rng(seed).integers(0, 256, block, dtype=np.uint8).tobytes()
    return blk + bytes(gap) + blk + blk[:5000]


WORKLOADS = {
    # name: (payload kind, raw bytes per stream, seed base, quality, lgwin, description)
    "c4_highratio_w16": ("repeat2k", 262144, 1000, 5, 16,
                         "synthetic 64 KiB-window high-ratio streams (2 KiB block x128, 4 mutations/rep, q5 lgwin16)"),
    "c5_stored_10k": ("random", 10000, 2000, 5, 16, "10,000 random bytes per stream -> stored meta-block (random_org_10k-like)"),
    "c5b_literals_10k": ("skewed", 10000, 2000, 5, 16, "10,000 skewed bytes per stream -> entropy-coded literals"),
    # round 2 (VERDICT "decide with data"): far references in a large window (WBITS 22: copies 512 KiB back, stored first half)
    "c7_far_w22": ("far_half", 1 << 20, 3000, 5, 22,
                   "1 MiB per stream, second half = first half 512 KiB back with mutations (q5 lgwin22): far-source copies"),
}


# Error classes of the 9 invalid corpus streams (reference tests/lib.rs:397-552 pin the messages)
CORPUS_INVALID_STATUS = {
    "frewsxcv_01.compressed": 24, "frewsxcv_02.compressed": 8, "frewsxcv_03.compressed": 12,
    "frewsxcv_04.compressed": 1, "frewsxcv_05.compressed": 24, "frewsxcv_06.compressed": 23,
    "frewsxcv_07.compressed": 1, "frewsxcv_08.compressed": 24, "frewsxcv_09.compressed": 10,
}


def corpus_workload(data_dir, only=None):
    """The reference's data/ corpus (vendored under tests/golden/data): -> (names, streams, expected bytes or None,
    expected status)."""
    import os
    names, streams, raws, status = [], [], [], []
    for fn in sorted(os.listdir(data_dir)):
        if ".compressed" not in fn or (only and fn not in only):
            continue
        names.append(fn)
        streams.append(open(os.path.join(data_dir, fn), "rb").read())
        if fn in CORPUS_INVALID_STATUS:
            raws.append(None)
            status.append(CORPUS_INVALID_STATUS[fn])
        else:
            raws.append(open(os.path.join(data_dir, fn.split(".compressed")[0]), "rb").read())
            status.append(0)
    return names, streams, raws, status


def make_unique_streams(name, n_unique):
    """-> (list of compressed streams, list of raw payloads) for workload `name`."""
    kind, size, seed, q, lgwin, _ = WORKLOADS[name]
    enc = libbrotli_enc()
    if enc is None:
        raise RuntimeError("libbrotlienc.so.1 is needed to synthesise the benchmark streams")
    raws = [synthetic_raw(kind, seed + i, size) for i in range(n_unique)]
    return [compress(enc, r, q, lgwin) for r in raws], raws
