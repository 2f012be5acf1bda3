"""Builds and binds the 32-LANE host simulation of the fused warp-per-stream decoder (brotli_rs_b200/csrc/bro_warpsim.cpp).

CPU TEST-SUITE ONLY.  tests/hostsim.py runs bro_decoder_core.h with a 1-lane warp; this one compiles the very code the
fused kernel and the resume kernel run (BRO_W = 32) with g++ and executes its 32 lanes as fibers that meet at the warp
intrinsics, in a lane order the test chooses -- so the lane-parallel table build, the shuffle-fed bit window, the
lane-parallel literal rounds and the warp copies are checked against the oracle where no GPU exists, and a missing
__syncwarp shows up as a wrong answer in one of the orders.  The product library never contains or calls it.
"""
import ctypes
import os
import platform
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "brotli_rs_b200", "csrc")
BUILD = os.path.join(ROOT, "tests", "_build")
_LIB = None

ASCENDING, DESCENDING, SHUFFLED = 0, 1, 2
ORDERS = (ASCENDING, DESCENDING, SHUFFLED)
SIM_ERRORS = {1: "lanes met at different intrinsics / masks (divergent collective)", 2: "a lane returned while another waited for it",
              3: "status or output length differs between lanes", 4: "a lane called a collective whose mask does not name it",
              100: "bytes written outside the output slot"}


def available():
    return platform.machine() == "x86_64"


def _build():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libbro_warpsim.so")
    srcs = [os.path.join(CSRC, "bro_warpsim.cpp"), os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("bro_warpsim.h", "bro_decoder_core.h", "bro_records.h", "bro_status.h", "bro_tables_generated.h")]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    cmd = ["g++", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-O2", "-o", so] + srcs + \
          ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")]
    subprocess.check_call(cmd)
    return so


class BroResume(ctypes.Structure):
    """bro_records.h: BroResume == include/brotli_b200.h: bro_resume (layout pinned by tests/test_abi_library.py)"""
    _fields_ = [("in_bits", ctypes.c_uint64), ("pos", ctypes.c_uint32), ("window", ctypes.c_uint32), ("dist", ctypes.c_uint32 * 4),
                ("p1", ctypes.c_uint32), ("p2", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_build())
        L.bro_warpsim_decode.restype = ctypes.c_int
        L.bro_warpsim_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int)]
        L.bro_warpsim_decode_resume.restype = ctypes.c_int
        L.bro_warpsim_decode_resume.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                                ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(BroResume), ctypes.POINTER(ctypes.c_int)]
        L.bro_warpsim_last_rendezvous.restype = ctypes.c_uint64
        assert L.bro_warpsim_resume_bytes() == ctypes.sizeof(BroResume)
        _LIB = L
    return _LIB


def decode(data: bytes, cap: int = 1 << 20, quirks: int = 0, latency: bool = False, order: int = ASCENDING, seed: int = 1):
    """One stream through the fused kernel's per-warp code.  latency: the latency build (a 10-bit literal root on chip for the
    general loop's lane-parallel rounds).  -> (status, bytes); raises when the warp itself misbehaved (SIM_ERRORS)."""
    out = ctypes.create_string_buffer(max(cap, 1))
    n, err = ctypes.c_size_t(), ctypes.c_int()
    st = lib().bro_warpsim_decode(data, len(data), out, cap, ctypes.byref(n), quirks, int(latency), order, seed, ctypes.byref(err))
    if err.value:
        raise AssertionError("warp simulation: " + SIM_ERRORS.get(err.value, str(err.value)))
    return st, out.raw[: n.value]


def rendezvous():
    """warp intrinsics executed by the last decode"""
    return int(lib().bro_warpsim_last_rendezvous())


def stream_decode(data: bytes, chunk_sizes, out_cap=1 << 16, quirks=0, order=ASCENDING, seed=1):
    """tests/hostsim.py's streaming-reader loop with the resume kernel's 32-lane code as the one call"""
    import hostsim
    L = lib()

    def step(buf, out, n_out, q, ck):
        err = ctypes.c_int()
        st = L.bro_warpsim_decode_resume(buf, len(buf), ctypes.cast(out, ctypes.c_char_p), len(out), ctypes.byref(n_out), q, order, seed,
                                         ctypes.cast(ctypes.byref(ck), ctypes.POINTER(BroResume)), ctypes.byref(err))
        if err.value:
            raise AssertionError("warp simulation: " + SIM_ERRORS.get(err.value, str(err.value)))
        return st

    return hostsim.stream_decode(data, chunk_sizes, out_cap=out_cap, quirks=quirks, step=step)


def sync_hits(reset=True):
    """{line of bro_decoder_core.h: executions of the bro_syncwarp() there} since the last reset"""
    import numpy as np
    hits = np.zeros(4096, dtype=np.uint64)
    L = lib()
    L.bro_warpsim_sync_hits.restype = None
    L.bro_warpsim_sync_hits.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.bro_warpsim_sync_hits(hits.ctypes.data, 4096, int(reset))
    return {int(i): int(hits[i]) for i in hits.nonzero()[0]}


class drop_sync:
    """with drop_sync(line): the bro_syncwarp() on that line of bro_decoder_core.h is left out (a mutation the lane orders must notice)"""

    def __init__(self, line):
        self.line = line

    def __enter__(self):
        lib().bro_warpsim_drop_sync(self.line)
        return self

    def __exit__(self, *exc):
        lib().bro_warpsim_drop_sync(-1)


def set_alignment(in_mis=0, out_mis=0):
    """address of the stream's first byte mod 128 and of the output slot's first byte mod 16 for the decodes that follow"""
    lib().bro_warpsim_set_alignment(in_mis, out_mis)


# ---- the copy kernel (bro_kernels_copy.cu compiled for the host, bro_warpsim_copy.cpp) behind phase one's host simulation ----
_LIB_COPY = None
COPY_SOURCES = ("bro_warpsim_copy.cpp", "bro_hostsim_parse.cpp", "bro_hostsim_copy.cpp")
COPY_DEPS = ("bro_warpsim.h", "bro_kernels_copy.cu", "bro_copy_piece.h", "bro_kernels.h", "bro_decoder_core.h", "bro_parse.h", "bro_records.h",
             "bro_status.h", "bro_tables_generated.h")


def _build_copy():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libbro_warpsim_copy.so")
    srcs = [os.path.join(CSRC, f) for f in COPY_SOURCES] + [os.path.join(ROOT, "oracle", "dict_blob.c")]
    deps = srcs + [os.path.join(CSRC, f) for f in COPY_DEPS]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    cmd = ["g++", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-O2", "-o", so] + srcs + \
          ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")]
    subprocess.check_call(cmd)
    return so


def two_phase(streams, caps, quirks=0, shape=0, order=ASCENDING, seed=1, in_mis=0, out_mis=0, queue_seed=0):
    """A batch through the two-phase path: phase one = the parse kernel's per-lane code on the host, phase two = ONE launch of the
    copy kernel's own code by a simulated warp (shape 0 / 1: the kernel's two instantiations).  -> [(status, bytes)], (bytes moved by
    records, records executed); a status in hostsim.RETRY means the product would hand that stream to the fused kernel."""
    import numpy as np
    global _LIB_COPY
    if _LIB_COPY is None:
        L = ctypes.CDLL(_build_copy())
        L.bro_warpsim_two_phase.restype = ctypes.c_int
        L.bro_warpsim_two_phase.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                            ctypes.c_uint, ctypes.c_uint, ctypes.c_uint64, ctypes.c_void_p]
        _LIB_COPY = L
    n = len(streams)
    in_off = np.zeros(n + 1, dtype=np.uint64)
    in_off[1:] = np.cumsum([len(s) for s in streams])
    out_off = np.zeros(n + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)
    inb = np.frombuffer(b"".join(streams) + b"\0", dtype=np.uint8).copy()
    out = np.zeros(int(out_off[n]) + 1, dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    stats = np.zeros(2, dtype=np.uint64)
    err = _LIB_COPY.bro_warpsim_two_phase(inb.ctypes.data, in_off.ctypes.data, out.ctypes.data, out_off.ctypes.data, out_len.ctypes.data,
                                          status.ctypes.data, n, quirks, shape, order, seed, in_mis, out_mis, queue_seed, stats.ctypes.data)
    if err:
        raise AssertionError("warp simulation (copy kernel): " + SIM_ERRORS.get(err, str(err)))
    res = []
    for i in range(n):
        b = int(out_off[i])
        res.append((int(status[i]), out[b: b + int(out_len[i])].tobytes()))
    return res, (int(stats[0]), int(stats[1]))


# ---- the parse kernel (bro_kernels_parse.cu compiled for the host, bro_warpsim_parse.cpp), and both kernels back to back ----
_LIB_PARSE = None
PARSE_DEPS = ("bro_warpsim.h", "bro_kernels_parse.cu", "bro_parse.h", "bro_kernels.h", "bro_decoder_core.h", "bro_records.h", "bro_status.h",
              "bro_tables_generated.h")


def _lib_parse():
    global _LIB_PARSE
    if _LIB_PARSE is None:
        os.makedirs(BUILD, exist_ok=True)
        so = os.path.join(BUILD, "libbro_warpsim_parse.so")
        srcs = [os.path.join(CSRC, "bro_warpsim_parse.cpp"), os.path.join(ROOT, "oracle", "dict_blob.c")]
        deps = srcs + [os.path.join(CSRC, f) for f in PARSE_DEPS]
        if not (os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps)):
            subprocess.check_call(["g++", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-O2", "-o", so] + srcs +
                                  ["-Wa,-I" + os.path.join(ROOT, "brotli_rs_b200", "data")])
        L = ctypes.CDLL(so)
        L.bro_warpsim_parse_launch.restype = ctypes.c_int
        L.bro_warpsim_parse_launch.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _LIB_PARSE = L
    return _LIB_PARSE


PARSE_ERRORS = {102: "a stream was never reported", 103: "the completion queue does not hold every stream exactly once"}


def two_phase_kernels(streams, caps, quirks=0, lanes=32, hand_out=None, order=ASCENDING, seed=1, in_mis=0, out_mis=0, ring_late=True,
                      copy_shape=0, copy_order=None, sizing=False, retry_pass=False, retry_latency=False, threads=32):
    """A batch through BOTH kernels of the two-phase path as compiled for the host: ONE launch of the parse kernel (a warp holding
    `lanes` streams at a time, streams handed out in `hand_out` order) and ONE launch of the copy kernel over the completion queue
    the parse kernel left.  -> [(status, bytes)], streams handed to the fused kernel, completion order.  sizing: bro_batch_sizes'
    mode (only the parse kernel; bytes are empty, out_len = decoded size -> [(status, size)]).  retry_pass: the third launch of the
    product's two-phase path as well -- bro_decode_warp_kernel in retry mode over the same buffers decodes exactly the streams phase
    one handed over, so that every status is final.  threads: the CTA of every launch (32 ... 256: up to eight warps side by side,
    competing for the streams of the work queues and interleaving their entries in the completion queue)."""
    import numpy as np
    import hostsim
    two_phase([], [])            # (loads the copy library)
    LP, LC = _lib_parse(), _LIB_COPY
    LC.bro_warpsim_copy_launch.restype = ctypes.c_int
    LC.bro_warpsim_copy_launch.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                           ctypes.c_void_p, ctypes.c_int]
    n = len(streams)
    PAD = 256
    in_off = np.zeros(n + 1, dtype=np.uint64)
    in_off[1:] = np.cumsum([len(s) for s in streams])
    out_off = np.zeros(n + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)

    def padded(nbytes, mis, fill):
        raw = np.full(nbytes + 2 * PAD + 16, fill, dtype=np.uint8)
        start = (-raw.ctypes.data) % 16 + PAD + mis
        return raw, start

    in_raw, in_s = padded(int(in_off[n]), in_mis, 0xee)
    in_raw[in_s: in_s + int(in_off[n])] = np.frombuffer(b"".join(streams), dtype=np.uint8)
    out_raw, out_s = padded(int(out_off[n]), out_mis, 0xdd)
    rec_total = (int(in_off[n]) >> 1) + 32 * n
    rec_raw = np.full(4 * (rec_total + 8), 0xabababab, dtype=np.uint32)
    rec_s = ((-rec_raw.ctypes.data) % 16) // 4
    out_len = np.zeros(n + 1, dtype=np.uint64)
    status = np.zeros(n + 1, dtype=np.int32)
    nrec = np.zeros(n + 1, dtype=np.uint32)
    done_q = np.zeros(n + 1, dtype=np.uint32)
    retry = np.zeros(1, dtype=np.uint32)
    ho = None if hand_out is None else np.asarray(hand_out, dtype=np.uint32)
    err = LP.bro_warpsim_parse_launch(in_raw.ctypes.data + in_s, in_off.ctypes.data, out_raw.ctypes.data + out_s, out_off.ctypes.data,
                                      out_len.ctypes.data, status.ctypes.data, nrec.ctypes.data, rec_raw.ctypes.data + 4 * rec_s, rec_total, n,
                                      None if ho is None else ho.ctypes.data, lanes, quirks, int(sizing), order, seed, int(ring_late),
                                      done_q.ctypes.data, retry.ctypes.data, threads)
    if err:
        raise AssertionError("warp simulation (parse kernel): " + (PARSE_ERRORS.get(err) or SIM_ERRORS.get(err, str(err))))
    assert int(retry[0]) == sum(int(s) in hostsim.RETRY for s in status[:n])
    if sizing:
        return [(int(status[i]), int(out_len[i])) for i in range(n)], int(retry[0]), done_q[:n].tolist()
    stats = np.zeros(2, dtype=np.uint64)
    err = LC.bro_warpsim_copy_launch(in_raw.ctypes.data + in_s, in_off.ctypes.data, out_raw.ctypes.data + out_s, out_off.ctypes.data,
                                     status.ctypes.data, nrec.ctypes.data, rec_raw.ctypes.data + 4 * rec_s, n, done_q.ctypes.data, copy_shape,
                                     order if copy_order is None else copy_order, seed, stats.ctypes.data, threads)
    if err:
        raise AssertionError("warp simulation (copy kernel): " + SIM_ERRORS.get(err, str(err)))
    if retry_pass and int(retry[0]):
        LF = lib()
        LF.bro_warpsim_fused_launch.restype = ctypes.c_int
        LF.bro_warpsim_fused_launch.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
        before = status[:n].copy()
        err = LF.bro_warpsim_fused_launch(in_raw.ctypes.data + in_s, in_off.ctypes.data, out_raw.ctypes.data + out_s, out_off.ctypes.data,
                                          out_len.ctypes.data, status.ctypes.data, n, None, 1, int(retry_latency), quirks, order, seed, threads)
        if err:
            raise AssertionError("warp simulation (fused kernel, retry pass): " + SIM_ERRORS.get(err, str(err)))
        keep = np.array([int(x) not in hostsim.RETRY for x in before])
        assert (status[:n][keep] == before[keep]).all(), "the retry pass touched a stream that was not handed over"
        assert not any(int(x) in hostsim.RETRY for x in status[:n])
    assert (out_raw[:out_s] == 0xdd).all() and (out_raw[out_s + int(out_off[n]):] == 0xdd).all(), "bytes written outside the batch's output"
    res = []
    for i in range(n):
        b = out_s + int(out_off[i])
        res.append((int(status[i]), out_raw[b: b + int(out_len[i])].tobytes()))
    return res, int(retry[0]), done_q[:n].tolist()


def order_streams(in_off, order=ASCENDING, seed=1):
    """the size-class ordering kernels (bro_order_*: what bro_order_launch launches in front of the parse kernel), CTA by CTA
    -> (order[n], gate[2])"""
    import numpy as np
    L = _lib_parse()
    L.bro_warpsim_order.restype = ctypes.c_int
    L.bro_warpsim_order.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64]
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    n = len(in_off) - 1
    out = np.full(n + 1, 0xffffffff, dtype=np.uint32)
    gate = np.zeros(2, dtype=np.uint32)
    err = L.bro_warpsim_order(in_off.ctypes.data, n, out.ctypes.data, gate.ctypes.data, order, seed)
    if err:
        raise AssertionError("warp simulation (ordering kernels): " + SIM_ERRORS.get(err, str(err)))
    return out[:n], gate


def sizes_finish(status):
    """bro_sizes_finish_kernel over a status array (bro_batch_sizes' last launch: hand-over codes -> SizeUnknown) -> new array"""
    import numpy as np
    L = _lib_parse()
    L.bro_warpsim_sizes_finish.restype = ctypes.c_int
    L.bro_warpsim_sizes_finish.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    st = np.ascontiguousarray(status, dtype=np.int32).copy()
    assert L.bro_warpsim_sizes_finish(st.ctypes.data, len(st)) == 0
    return st
