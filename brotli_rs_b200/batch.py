"""Batch decode API: many independent Brotli streams per kernel launch.

Host-side mirror of the C ABI's bro_batch_decode / bro_batch_decode_host.  torch is used only as the owner of
device memory and streams; the decode itself is the CUDA kernel in libbrotli_b200.so.
"""
import ctypes

import numpy as np

from . import _lib


def pack_streams(streams, align=1):
    """Concatenate byte strings -> (uint8 array, uint64 offsets[n+1]).  `align` pads each start (kept in offsets
    as the *end of the previous stream*, so with align > 1 the offsets array describes padded slots and the true
    lengths must be carried separately; the decoder API uses align=1)."""
    assert align == 1
    lens = np.fromiter((len(s) for s in streams), dtype=np.uint64, count=len(streams))
    off = np.zeros(len(streams) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(streams), dtype=np.uint8) if len(streams) else np.zeros(0, dtype=np.uint8)
    return buf, off


def slot_offsets(capacities, align=16):
    """Output slot offsets for the given per-stream capacities, each slot start aligned to `align` bytes (aligned
    slots let the copy phase use full 16-byte stores from the first byte)."""
    caps = np.asarray(capacities, dtype=np.uint64)
    padded = (caps + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    off = np.zeros(len(caps) + 1, dtype=np.uint64)
    np.cumsum(padded, out=off[1:])
    return off


class BatchDecoder:
    """One GPU's decoder context (bro_ctx).  Not thread safe; use one per host thread / CUDA stream."""

    MODE_AUTO, MODE_WARP, MODE_TWOPHASE = 0, 1, 2

    def __init__(self, device=None, quirks=0, mode=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("brotli_rs_b200 needs a CUDA device: the decoder has no CPU path")
        self._lib = _lib.load_library()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)       # make sure the primary context exists before the library uses it
            h = ctypes.c_void_p()
            st = self._lib.bro_ctx_create(ctypes.byref(h), self.device.index)
        if st != 0:
            raise _lib.BroError(st, "bro_ctx_create failed with status %d" % st)
        self._ctx = h
        if quirks:
            self._lib.bro_ctx_set_quirks(self._ctx, quirks)
        if mode is not None:
            self.set_mode(mode)

    def set_mode(self, mode):
        """MODE_AUTO (default) / MODE_WARP / MODE_TWOPHASE -- see bro_ctx_set_mode in include/brotli_b200.h."""
        self._check(self._lib.bro_ctx_set_mode(self._ctx, int(mode)))

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.bro_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self._lib.bro_ctx_launch_count(self._ctx))

    def set_timing(self, on):
        """Record CUDA events around the kernels of every batch (bench.py's per-kernel roofline)."""
        self._check(self._lib.bro_ctx_set_timing(self._ctx, 1 if on else 0))

    def last_kernel_ms(self):
        """-> dict of kernel durations (ms) of the last batch; waits for it."""
        ms = (ctypes.c_float * 4)()
        self._check(self._lib.bro_ctx_last_kernel_ms(self._ctx, ms))
        return {"order": float(ms[0]), "parse": float(ms[1]), "copy": float(ms[2]), "fused": float(ms[3])}

    def last_batch_stats(self):
        """-> dict(copy_bytes, copy_records, retried_streams, gated_to_fused) of the last batch; synchronises the device."""
        st = (ctypes.c_uint64 * 4)()
        self._check(self._lib.bro_ctx_last_batch_stats(self._ctx, st))
        return {"copy_bytes": int(st[0]), "copy_records": int(st[1]), "retried_streams": int(st[2]), "gated_to_fused": int(st[3])}

    @property
    def num_warps(self):
        return int(self._lib.bro_ctx_num_warps(self._ctx))

    def _check(self, st):
        if st != 0:
            msg = self._lib.bro_ctx_last_cuda_error(self._ctx).decode() if st == _lib.CUDA_ERROR else None
            raise _lib.BroError(st, msg and "CUDA error: " + msg)

    def decode_device(self, d_in, d_in_off, d_out, d_out_off, d_out_len=None, d_status=None, stream=None):
        """bro_batch_decode on torch CUDA tensors (asynchronous on `stream` / the current stream).
        d_in uint8, d_in_off/d_out_off int64-or-uint64 [n+1], d_out uint8.  Returns (d_out_len, d_status)."""
        import torch
        n = d_in_off.numel() - 1
        assert d_out_off.numel() == n + 1
        for t in (d_in, d_in_off, d_out, d_out_off):
            assert t.is_cuda and t.is_contiguous()
        assert d_in_off.element_size() == 8 and d_out_off.element_size() == 8
        if d_out_len is None:
            d_out_len = torch.empty(n, dtype=torch.int64, device=d_in.device)
        if d_status is None:
            d_status = torch.empty(n, dtype=torch.int32, device=d_in.device)
        s = stream if stream is not None else torch.cuda.current_stream(d_in.device)
        # the buffer's size bounds the compressed bytes of the batch: no read-back of the offsets (bro_ctx_reserve)
        self._check(self._lib.bro_ctx_reserve(self._ctx, max(1, d_in.numel()), n))
        st = self._lib.bro_batch_decode(self._ctx, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(),
                                        d_out_off.data_ptr(), d_out_len.data_ptr(), d_status.data_ptr(), n,
                                        ctypes.c_void_p(s.cuda_stream))
        self._check(st)
        return d_out_len, d_status

    RESUME_DTYPE = np.dtype([("in_bits", "<u8"), ("pos", "<u4"), ("window", "<u4"), ("dist", "<u4", 4), ("p1", "<u4"),
                             ("p2", "<u4"), ("flags", "<u4"), ("reserved", "<u4")])      # bro_resume (include/brotli_b200.h)

    def decode_resume_device(self, d_in, d_in_off, d_out, d_out_off, d_resume, d_out_len=None, d_status=None, stream=None):
        """bro_batch_decode_resume on torch CUDA tensors: decode from and to the resume points in d_resume (uint8
        tensor of n * 48 bytes, RESUME_DTYPE records; zeros = start of stream).  Returns (d_out_len, d_status)."""
        import torch
        n = d_in_off.numel() - 1
        assert d_out_off.numel() == n + 1 and d_resume.numel() * d_resume.element_size() == n * self.RESUME_DTYPE.itemsize
        for t in (d_in, d_in_off, d_out, d_out_off, d_resume):
            assert t.is_cuda and t.is_contiguous()
        if d_out_len is None:
            d_out_len = torch.empty(n, dtype=torch.int64, device=d_in.device)
        if d_status is None:
            d_status = torch.empty(n, dtype=torch.int32, device=d_in.device)
        s = stream if stream is not None else torch.cuda.current_stream(d_in.device)
        self._check(self._lib.bro_batch_decode_resume(self._ctx, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(),
                                                      d_out_off.data_ptr(), d_out_len.data_ptr(), d_status.data_ptr(),
                                                      d_resume.data_ptr(), n, ctypes.c_void_p(s.cuda_stream)))
        return d_out_len, d_status

    def decode_host(self, in_buf, in_off, out_off, out=None):
        """bro_batch_decode_host on numpy arrays (or anything exposing a writable buffer, e.g. pinned torch tensors
        via .numpy()).  Returns (out, out_len, status)."""
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        n = len(in_off) - 1
        if out is None:
            out = np.empty(int(out_off[-1]), dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        st = self._lib.bro_batch_decode_host(self._ctx, in_buf.ctypes.data, in_off.ctypes.data, out.ctypes.data,
                                             out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data, n)
        self._check(st)
        return out, out_len, status

    def sizes_device(self, d_in, d_in_off, stream=None):
        """bro_batch_sizes on torch CUDA tensors: measure the streams without writing a byte (asynchronous).
        Returns (d_out_len, d_status); status 103 (SIZE_UNKNOWN) = only a real decode tells this stream's size."""
        import torch
        n = d_in_off.numel() - 1
        d_out_len = torch.empty(n, dtype=torch.int64, device=d_in.device)
        d_status = torch.empty(n, dtype=torch.int32, device=d_in.device)
        s = stream if stream is not None else torch.cuda.current_stream(d_in.device)
        self._check(self._lib.bro_batch_sizes(self._ctx, d_in.data_ptr(), d_in_off.data_ptr(), d_out_len.data_ptr(),
                                              d_status.data_ptr(), n, ctypes.c_void_p(s.cuda_stream)))
        return d_out_len, d_status

    def decode_unsized(self, streams):
        """bro_batch_decode_unsized_host: list of byte strings, NO size hints -> list of (status, bytes).  The library
        measures the streams, sizes the slots itself and returns one buffer."""
        in_buf, in_off = pack_streams(streams)
        n = len(streams)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        h_out = ctypes.c_void_p()
        st = self._lib.bro_batch_decode_unsized_host(self._ctx, in_buf.ctypes.data if len(in_buf) else None, in_off.ctypes.data, n,
                                                     ctypes.byref(h_out), out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data)
        self._check(st)
        try:
            total = int(out_off[-1])
            buf = (ctypes.c_uint8 * max(total, 1)).from_address(h_out.value) if h_out.value else None
            mv = memoryview(buf) if buf is not None else memoryview(b"")
            return [(int(status[i]), bytes(mv[int(out_off[i]): int(out_off[i]) + int(out_len[i])])) for i in range(n)]
        finally:
            if h_out.value:
                self._lib.bro_free(h_out)

    def decode_streams(self, streams, capacities):
        """Convenience: list of byte strings + per-stream output capacities -> list of (status, bytes)."""
        in_buf, in_off = pack_streams(streams)
        out_off = slot_offsets(capacities)
        out, out_len, status = self.decode_host(in_buf, in_off, out_off)
        res = []
        for i in range(len(streams)):
            b = int(out_off[i])
            n = min(int(out_len[i]), int(out_off[i + 1]) - b)      # (a failed stream may report a position beyond its slot)
            res.append((int(status[i]), out[b: b + n].tobytes()))
        return res


def mg_partition(in_off, out_off, ngpus):
    """bro_mg_partition: the contiguous ranges of streams bro_mg_decode_host gives the devices (n + 1 offsets each;
    -> ngpus + 1 cut points).  Pure host arithmetic: works without a GPU."""
    lib = _lib.load_library()
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    first = np.zeros(ngpus + 1, dtype=np.uint32)
    st = lib.bro_mg_partition(in_off.ctypes.data, out_off.ctypes.data, len(in_off) - 1, ngpus, first.ctypes.data)
    if st != 0:
        raise _lib.BroError(st)
    return first


class MultiGpuDecoder:
    """bro_mg_*: one batch over several GPUs of this box from ONE process (one context and one host thread per device;
    streams are independent, nothing crosses between GPUs).  The torch.distributed layer (shard.py, bench.py) is the
    one-process-per-GPU form of the same sharding."""

    def __init__(self, ngpus=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("brotli_rs_b200 needs a CUDA device: the decoder has no CPU path")
        self._lib = _lib.load_library()
        h = ctypes.c_void_p()
        st = self._lib.bro_mg_create(ctypes.byref(h), int(ngpus))
        if st != 0:
            raise _lib.BroError(st)
        self._mg = h

    @property
    def device_count(self):
        return int(self._lib.bro_mg_device_count(self._mg))

    def set_mode(self, mode):
        for k in range(self.device_count):
            self._lib.bro_ctx_set_mode(self._lib.bro_mg_ctx(self._mg, k), int(mode))

    def decode_host(self, in_buf, in_off, out_off, out=None):
        """bro_mg_decode_host: same arguments and results as BatchDecoder.decode_host"""
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        n = len(in_off) - 1
        if out is None:
            out = np.empty(int(out_off[-1]), dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        st = self._lib.bro_mg_decode_host(self._mg, in_buf.ctypes.data, in_off.ctypes.data, out.ctypes.data,
                                          out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data, n)
        if st != 0:
            msg = None
            if st == _lib.CUDA_ERROR:
                msg = "CUDA error: " + "; ".join(self._lib.bro_ctx_last_cuda_error(self._lib.bro_mg_ctx(self._mg, k)).decode()
                                                 for k in range(self.device_count))
            raise _lib.BroError(st, msg)
        return out, out_len, status

    def close(self):
        if getattr(self, "_mg", None):
            self._lib.bro_mg_destroy(self._mg)
            self._mg = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
