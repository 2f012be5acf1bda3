"""ctypes binding of the parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
--impl reference legs.  The product package (brotli_rs_b200) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK = 0
OUTPUT_TOO_SMALL = 100
PANIC_UPPERCASE_ZERO = 102


def build(force=False):
    """Compile liboracle.so with the committed recipe (oracle/Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force and os.path.exists(so):
        os.remove(so)
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.bro_oracle_decode.restype = ctypes.c_int
        L.bro_oracle_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_size_t), ctypes.c_int]
        L.bro_oracle_free.argtypes = [ctypes.c_void_p]
        L.bro_oracle_status_description.restype = ctypes.c_char_p
        L.bro_oracle_status_description.argtypes = [ctypes.c_int]
        L.bro_oracle_decode_batch.restype = ctypes.c_int
        L.bro_oracle_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.bro_oracle_transform.restype = ctypes.c_long
        L.bro_oracle_transform.argtypes = [ctypes.c_uint, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int]
        L.bro_oracle_imtf.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.bro_oracle_bitreader_script.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_size_t]
        L.bro_oracle_tree_decode.restype = ctypes.c_int
        L.bro_oracle_tree_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                             ctypes.c_void_p, ctypes.c_size_t]
        _LIB = L
    return _LIB


def decode(data: bytes, quirks: int = 0):
    """Decode one stream -> (status, bytes).  On error the bytes are whatever was produced before it."""
    L = lib()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    st = L.bro_oracle_decode(data, len(data), ctypes.byref(out), ctypes.byref(n), quirks)
    res = ctypes.string_at(out, n.value) if out.value else b""
    L.bro_oracle_free(out)
    return st, res


def description(status: int) -> str:
    return lib().bro_oracle_status_description(status).decode()


def decode_batch(in_buf: np.ndarray, in_off: np.ndarray, out_off: np.ndarray, nthreads: int = 1, quirks: int = 0,
                 out: np.ndarray = None):
    """Batch decode into caller-sized slots -> (out, out_len, status).  Same argument meaning as
    bro_batch_decode_host."""
    n = len(in_off) - 1
    in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    if out is None:
        out = np.empty(int(out_off[-1]), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    lib().bro_oracle_decode_batch(in_buf.ctypes.data, in_off.ctypes.data, out.ctypes.data, out_off.ctypes.data,
                                  out_len.ctypes.data, status.ctypes.data, n, nthreads, quirks)
    return out, out_len, status


def transform(tid: int, word: bytes, quirks: int = 0):
    buf = ctypes.create_string_buffer(128)
    n = lib().bro_oracle_transform(tid, word, len(word), buf, quirks)
    return None if n < 0 else buf.raw[:n]


def imtf(v: bytes) -> bytes:
    a = np.frombuffer(v, dtype=np.uint8).copy()
    lib().bro_oracle_imtf(a.ctypes.data, len(a))
    return a.tobytes()


_OPS = {"u8": 0, "bit": 1, "bits": 2, "tail": 3, "nibble": 4, "nibbles": 5}


def bitreader_script(data: bytes, ops):
    """ops: list of (name, arg); 'string n' is expanded to n u8 reads.  Returns list of values (ints / bytes)."""
    flat, shape = [], []
    for name, arg in ops:
        if name == "string":
            shape.append(arg)
            flat += [(0, 0)] * arg
        else:
            shape.append(None)
            flat.append((_OPS[name], arg or 0))
    o = np.array([f[0] for f in flat], dtype=np.int32)
    a = np.array([f[1] for f in flat], dtype=np.int32)
    r = np.zeros(len(flat), dtype=np.int64)
    lib().bro_oracle_bitreader_script(data, len(data), o.ctypes.data, a.ctypes.data, r.ctypes.data, len(flat))
    out, k = [], 0
    for s in shape:
        if s is None:
            out.append(int(r[k]))
            k += 1
        else:
            out.append(bytes(int(x) for x in r[k:k + s]))
            k += s
    return out


def tree_decode(lengths, data: bytes, nsyms: int):
    l = np.array(lengths, dtype=np.uint32)
    s = np.zeros(nsyms, dtype=np.int32)
    lib().bro_oracle_tree_decode(l.ctypes.data, len(l), data, len(data), s.ctypes.data, nsyms)
    return s.tolist()
