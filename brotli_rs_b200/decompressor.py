"""`Decompressor` -- the Python mirror of the reference's one public type, brotli::Decompressor<R: Read>
(src/lib.rs:378-410, 2173-2193), on top of the C ABI's Read-struct (bro_reader_*).

    with open("data/64x.compressed", "rb") as f:          # any object with .read(n)
        data = Decompressor(f).read()                      # == read_to_end

`Decompressor(r)` performs no I/O (like Decompressor::new); the first `read` drains `r`, decodes the stream on the
GPU and serves bytes from a host buffer.  An invalid stream raises BroError whose message is the reference's
error description (io::Error::new(InvalidData, description), src/lib.rs:2177).

`Decompressor(r, streaming=True)` (or streaming=<bytes asked of r at a time>) has the reference's memory behaviour
instead: input is read as it is needed and decoded meta-block by meta-block (bro_reader_new_streaming), the buffers are
bounded by the largest meta-block plus a window, and bytes decoded before an error are delivered before it is raised.
"""
import ctypes
import io

from . import _lib


class Decompressor(io.RawIOBase):
    def __init__(self, r, decoder=None, streaming=False):
        """r: a readable binary file-like object (or bytes).  decoder: an optional BatchDecoder whose context is
        shared; otherwise the reader creates its own context on the current device."""
        super().__init__()
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("brotli_rs_b200 needs a CUDA device: the decoder has no CPU path")
        self._lib = _lib.load_library()
        self._src = io.BytesIO(r) if isinstance(r, (bytes, bytearray, memoryview)) else r
        self._decoder = decoder

        def _cb(user, buf, cap):
            try:
                chunk = self._src.read(cap)
            except Exception:
                return -1
            if not chunk:
                return 0
            ctypes.memmove(buf, chunk, len(chunk))
            return len(chunk)

        self._cb = _lib.READ_CB(_cb)        # keep the trampoline alive as long as the reader
        ctx = decoder._ctx if decoder is not None else None
        if streaming:
            self._h = self._lib.bro_reader_new_streaming(ctx, self._cb, None, 0 if streaming is True else int(streaming))
        else:
            self._h = self._lib.bro_reader_new(ctx, self._cb, None)
        if not self._h:
            raise MemoryError("bro_reader_new failed")

    def readable(self):
        return True

    def readinto(self, b):
        """Read::read (src/lib.rs:2174-2192): fills `b` while data is available, 0 at end of stream."""
        mv = memoryview(b).cast("B")
        if len(mv) == 0:
            return 0
        buf = (ctypes.c_uint8 * len(mv)).from_buffer(mv)
        n = self._lib.bro_reader_read(self._h, buf, len(mv))
        if n < 0:
            msg = None
            if -n == _lib.CUDA_ERROR and self._decoder is not None:
                msg = "CUDA error: " + self._lib.bro_ctx_last_cuda_error(self._decoder._ctx).decode()
            raise _lib.BroError(-n, msg)
        return n

    @property
    def status(self):
        return self._lib.bro_reader_status(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bro_reader_free(self._h)
            self._h = None
        super().close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
