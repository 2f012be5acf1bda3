"""ctypes binding of libbrotli_b200.so (the C ABI of include/brotli_b200.h)."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK = 0
UNEXPECTED_EOF = 24
OUTPUT_TOO_SMALL = 100
CUDA_ERROR = 101
PANIC_UPPERCASE_ZERO = 102
SIZE_UNKNOWN = 103

# every symbol include/brotli_b200.h declares
ABI_SYMBOLS = [
    "bro_ctx_create", "bro_ctx_destroy", "bro_ctx_set_quirks", "bro_ctx_set_mode", "bro_ctx_last_cuda_error", "bro_ctx_launch_count",
    "bro_ctx_num_warps", "bro_ctx_reserve", "bro_ctx_set_timing", "bro_ctx_last_kernel_ms", "bro_ctx_last_batch_stats", "bro_batch_decode", "bro_batch_decode_host", "bro_batch_sizes", "bro_batch_decode_unsized_host", "bro_batch_decode_resume", "bro_reader_new_streaming", "bro_free", "bro_status_description",
    "bro_reader_new", "bro_reader_read", "bro_reader_status", "bro_reader_free",
    "bro_mg_create", "bro_mg_destroy", "bro_mg_device_count", "bro_mg_ctx", "bro_mg_partition", "bro_mg_decode_host",
]

READ_CB = ctypes.CFUNCTYPE(ctypes.c_ssize_t, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t)


class BroError(RuntimeError):
    """An invalid stream (status = the reference's DecompressorError number) or a CUDA failure."""

    def __init__(self, status, message=None):
        self.status = status
        super().__init__(message or status_description(status))


def library_path():
    # BRO_B200_LIB selects an experimental build of the same library (kernel tuning runs); default = the product build
    return os.environ.get("BRO_B200_LIB") or os.path.join(HERE, "lib", "libbrotli_b200.so")


def load_library():
    """Load the CUDA library; fail loudly if it has not been built (there is no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "libbrotli_b200.so is missing (%s). Build it with `python -m brotli_rs_b200.build`; "
            "brotli_rs_b200 has no CPU decode path." % path)
    L = ctypes.CDLL(path)
    vp, u32, u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64
    L.bro_ctx_create.restype = ctypes.c_int
    L.bro_ctx_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.bro_ctx_destroy.restype = None
    L.bro_ctx_destroy.argtypes = [vp]
    L.bro_ctx_set_quirks.restype = ctypes.c_int
    L.bro_ctx_set_quirks.argtypes = [vp, ctypes.c_int]
    L.bro_ctx_set_mode.restype = ctypes.c_int
    L.bro_ctx_set_mode.argtypes = [vp, ctypes.c_int]
    L.bro_ctx_last_cuda_error.restype = ctypes.c_char_p
    L.bro_ctx_last_cuda_error.argtypes = [vp]
    L.bro_ctx_launch_count.restype = u64
    L.bro_ctx_launch_count.argtypes = [vp]
    L.bro_ctx_num_warps.restype = u32
    L.bro_ctx_num_warps.argtypes = [vp]
    L.bro_ctx_set_timing.restype = ctypes.c_int
    L.bro_ctx_set_timing.argtypes = [vp, ctypes.c_int]
    L.bro_ctx_last_kernel_ms.restype = ctypes.c_int
    L.bro_ctx_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.bro_ctx_last_batch_stats.restype = ctypes.c_int
    L.bro_ctx_last_batch_stats.argtypes = [vp, ctypes.POINTER(u64)]
    L.bro_ctx_reserve.restype = ctypes.c_int
    L.bro_ctx_reserve.argtypes = [vp, u64, u32]
    L.bro_batch_decode.restype = ctypes.c_int
    L.bro_batch_decode.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, vp]
    L.bro_batch_decode_host.restype = ctypes.c_int
    L.bro_batch_decode_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32]
    L.bro_batch_sizes.restype = ctypes.c_int
    L.bro_batch_sizes.argtypes = [vp, vp, vp, vp, vp, u32, vp]
    L.bro_batch_decode_unsized_host.restype = ctypes.c_int
    L.bro_batch_decode_unsized_host.argtypes = [vp, vp, vp, u32, ctypes.POINTER(vp), vp, vp, vp]
    L.bro_free.restype = None
    L.bro_free.argtypes = [vp]
    L.bro_status_description.restype = ctypes.c_char_p
    L.bro_status_description.argtypes = [ctypes.c_int]
    L.bro_reader_new.restype = vp
    L.bro_reader_new.argtypes = [vp, READ_CB, vp]
    L.bro_reader_new_streaming.restype = vp
    L.bro_reader_new_streaming.argtypes = [vp, READ_CB, vp, ctypes.c_size_t]
    L.bro_batch_decode_resume.restype = ctypes.c_int
    L.bro_batch_decode_resume.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]
    L.bro_reader_read.restype = ctypes.c_ssize_t
    L.bro_reader_read.argtypes = [vp, vp, ctypes.c_size_t]
    L.bro_reader_status.restype = ctypes.c_int
    L.bro_reader_status.argtypes = [vp]
    L.bro_reader_free.restype = None
    L.bro_reader_free.argtypes = [vp]
    L.bro_mg_create.restype = ctypes.c_int
    L.bro_mg_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.bro_mg_destroy.restype = None
    L.bro_mg_destroy.argtypes = [vp]
    L.bro_mg_device_count.restype = ctypes.c_int
    L.bro_mg_device_count.argtypes = [vp]
    L.bro_mg_ctx.restype = vp
    L.bro_mg_ctx.argtypes = [vp, ctypes.c_int]
    L.bro_mg_partition.restype = ctypes.c_int
    L.bro_mg_partition.argtypes = [vp, vp, u32, ctypes.c_int, vp]
    L.bro_mg_decode_host.restype = ctypes.c_int
    L.bro_mg_decode_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32]
    _LIB = L
    return L


def status_description(status):
    return load_library().bro_status_description(int(status)).decode()
