"""Builds and binds the 1-lane host simulation of the warp decoder (brotli_rs_b200/csrc/bro_hostsim.cpp).

CPU TEST-SUITE ONLY: it lets `pytest -m "not gpu"` exercise the decoder logic of bro_decoder_core.h (table build,
canonical decode, EOF/hole semantics, every error path) against the oracle where no GPU exists.  The product
library never contains or calls it.
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "brotli_rs_b200", "csrc")
BUILD = os.path.join(ROOT, "tests", "_build")
_LIBS = {}


def _build(asan):
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libbro_hostsim_asan.so" if asan else "libbro_hostsim.so")
    srcs = [os.path.join(CSRC, "bro_hostsim.cpp"), os.path.join(CSRC, "bro_hostsim_parse.cpp"), os.path.join(CSRC, "bro_hostsim_copy.cpp"),
            os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("bro_decoder_core.h", "bro_parse.h", "bro_records.h", "bro_copy_piece.h", "bro_status.h", "bro_tables_generated.h")]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] if asan else ["-O2"]
    cmd = ["g++", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas"] + flags + ["-o", so] + srcs + \
          ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")]
    subprocess.check_call(cmd)
    return so


def lib(asan=False):
    if asan not in _LIBS:
        L = ctypes.CDLL(_build(asan))
        L.bro_hostsim_decode.restype = ctypes.c_int
        L.bro_hostsim_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_uint]
        L.bro_hostsim_parse_decode.restype = ctypes.c_int
        L.bro_hostsim_parse_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                               ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_uint, ctypes.c_uint,
                                               ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]
        _LIBS[asan] = L
    return _LIBS[asan]


ARENA_TOO_SMALL = 105
NEED_FUSED = 106
RECORDS_FULL = 107
RETRY = (ARENA_TOO_SMALL, NEED_FUSED, RECORDS_FULL)


def decode(data: bytes, cap: int = 1 << 20, quirks: int = 0, arena_u16: int = 0):
    """arena_u16 = 0: worst-case arena (warp kernel); thread_arena_u16(): the thread kernel's 64 KiB arena."""
    out = ctypes.create_string_buffer(max(cap, 1))
    n = ctypes.c_size_t()
    st = lib().bro_hostsim_decode(data, len(data), out, cap, ctypes.byref(n), quirks, arena_u16)
    return st, out.raw[: n.value]


def thread_arena_u16():
    return lib().bro_hostsim_thread_arena_u16()


def parse_decode(data: bytes, cap: int = 1 << 20, quirks: int = 0, arena_u16: int = 0, rec_cap: int = 0, asan=False, mis=None):
    """The two-phase path: phase one = the parse kernel's per-lane code (bro_parse.h: header step and lockstep rounds), phase two = a byte loop
    over its copy records.  -> (status, bytes, records, machine trips); status may be one of RETRY, which the product
    answers by re-running the stream with the fused warp kernel.  mis = 0..15: the output slot starts at an address with
    (address & 15) == mis (the pieces of the copy records are cut for the slot's alignment)."""
    buf = ctypes.create_string_buffer(max(cap, 1) + 48)
    base = ctypes.addressof(buf)
    off = ((-base) % 16 + 16 + mis) if mis is not None else 0
    n = ctypes.c_size_t()
    nrec, steps = ctypes.c_uint(), ctypes.c_uint()
    st = lib(asan).bro_hostsim_parse_decode(data, len(data), base + off, cap, ctypes.byref(n), quirks, arena_u16, rec_cap,
                                            ctypes.byref(nrec), ctypes.byref(steps))
    return st, buf.raw[off: off + n.value], nrec.value, steps.value


class copy_group:
    """with copy_group(G): parse_decode executes the records as the copy kernel does -- grouping, and for long records the
    kernel's own lane code with G (32, 16, 8) lanes per piece -- instead of the byte loop.  .stats() -> groups executed
    by the piece path, as short records, periodic fills (accumulated over the last decode)."""

    def __init__(self, group, asan=False):
        self.group, self.L = group, lib(asan)

    def __enter__(self):
        self.L.bro_hostsim_parse_set_copy_group(self.group)
        return self

    def __exit__(self, *exc):
        self.L.bro_hostsim_parse_set_copy_group(0)

    def stats(self):
        st = (ctypes.c_uint * 3)()
        self.L.bro_hostsim_parse_copy_stats(st)
        return tuple(st)

    def win_stats(self):
        """window form (group 308): (pieces served by the ring, pieces moved) since the last call"""
        st = (ctypes.c_uint64 * 2)()
        self.L.bro_hostsim_copy_win_stats(st)
        return int(st[0]), int(st[1])


def parse_records(data: bytes, cap: int = 1 << 20):
    """The copy records phase one writes for a stream: -> (status, [(dst, len, kind, a)], out_mis) where out_mis is the
    alignment (address & 15) of the output buffer the pieces were cut for."""
    import numpy as np
    L = lib()
    L.bro_hostsim_parse_export_records.restype = None
    L.bro_hostsim_parse_export_records.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    max_rec = len(data) // 2 + 32
    words = np.zeros(4 * max_rec, dtype=np.uint32)
    L.bro_hostsim_parse_export_records(words.ctypes.data, max_rec)
    st, out, nrec, _ = parse_decode(data, cap=cap)
    recs = [(int(words[4 * k]), int(words[4 * k + 1]) & 0x0fffffff, int(words[4 * k + 1]) >> 28, int(words[4 * k + 2])) for k in range(nrec)]
    return st, recs, (int(words[3]) if nrec else 0), out


def parse_size(data: bytes, quirks: int = 0):
    """bro_batch_sizes' mode of phase one: measure without writing.  -> (status, decoded size)"""
    L = lib()
    L.bro_hostsim_parse_set_sizing.restype = None
    L.bro_hostsim_parse_set_sizing.argtypes = [ctypes.c_uint]
    L.bro_hostsim_parse_set_sizing(1)
    try:
        out = ctypes.create_string_buffer(1)
        n = ctypes.c_size_t()
        nrec, steps = ctypes.c_uint(), ctypes.c_uint()
        st = L.bro_hostsim_parse_decode(data, len(data), ctypes.addressof(out), 0, ctypes.byref(n), quirks, 0, 0, ctypes.byref(nrec), ctypes.byref(steps))
        return st, n.value
    finally:
        L.bro_hostsim_parse_set_sizing(0)


class Resume(ctypes.Structure):
    """BroResume (brotli_rs_b200/csrc/bro_records.h)"""
    _fields_ = [("in_bits", ctypes.c_uint64), ("pos", ctypes.c_uint32), ("window", ctypes.c_uint32), ("dist", ctypes.c_uint32 * 4),
                ("p1", ctypes.c_uint32), ("p2", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


RESUME_HEADER, RESUME_LAST, RESUME_ENDED = 1, 2, 4
UNEXPECTED_EOF, EXPECTED_END_OF_STREAM, OUTPUT_TOO_SMALL = 24, 2, 100


def stream_decode(data: bytes, chunk_sizes, out_cap=1 << 16, quirks=0, max_calls=100000, step=None):
    """The streaming reader's loop (bro_reader_* in bro_abi.cu) over the host simulation of the resumable decode: the
    input arrives in pieces of chunk_sizes (cycled), the output buffer holds the history (at most a window) plus what one
    call produces.  -> (final status, bytes served before it, calls, largest input buffer, largest output buffer).
    step(buf, out, n_out, quirks, ck) -> status: another implementation of the one call (tests/warpsim.py: the 32-lane form)."""
    L = lib()
    L.bro_hostsim_decode_resume.restype = ctypes.c_int
    L.bro_hostsim_decode_resume.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.POINTER(Resume)]
    ck = Resume()
    served = bytearray()
    fed, k = 0, 0                       # bytes of `data` handed over so far; index into chunk_sizes
    buf = b""                           # unconsumed input
    out = ctypes.create_string_buffer(out_cap)
    hist = 0                            # history bytes at the start of `out`
    want = 0                            # read more input until the buffer holds this much
    calls = max_in = 0
    max_out = out_cap
    while True:
        while fed < len(data) and (len(buf) < max(want, 1)):
            n = chunk_sizes[k % len(chunk_sizes)]
            k += 1
            buf += data[fed: fed + n]
            fed += n
        eof = fed >= len(data)
        max_in = max(max_in, len(buf))
        n_out = ctypes.c_size_t()
        before = (ck.in_bits, ck.pos, ck.flags)
        if step is None:
            st = L.bro_hostsim_decode_resume(buf, len(buf), ctypes.addressof(out), len(out), ctypes.byref(n_out), quirks, ctypes.byref(ck))
        else:
            st = step(buf, out, n_out, quirks, ck)
        calls += 1
        assert calls <= max_calls, "the streaming loop does not terminate"
        progress = (ck.in_bits, ck.pos, ck.flags) != before
        final_pos = n_out.value if st == 0 else ck.pos      # bytes behind the last resume point are not final
        served += out.raw[hist: final_pos]
        if st == 0:
            if not eof:                  # the device saw the end of ITS input; bytes still to come are trailing garbage
                return EXPECTED_END_OF_STREAM, bytes(served), calls, max_in, max_out
            return 0, bytes(served), calls, max_in, max_out
        # consumed input goes, the history slides to the front
        drop = ck.in_bits >> 3
        buf = buf[drop:]
        ck.in_bits &= 7
        keep = min(ck.pos, ck.window) if ck.flags & RESUME_HEADER else 0
        raw = out.raw
        if st == OUTPUT_TOO_SMALL and not progress:
            out = ctypes.create_string_buffer(2 * len(out))
            max_out = max(max_out, len(out))
        ctypes.memmove(out, raw[ck.pos - keep: ck.pos], keep)
        ck.pos = hist = keep
        if st == UNEXPECTED_EOF and not eof:
            want = len(buf) + 1 if progress else 2 * len(buf) + 1      # no progress: this meta-block needs more input at once
            continue
        if st == OUTPUT_TOO_SMALL:
            continue
        return st, bytes(served), calls, max_in, max_out
