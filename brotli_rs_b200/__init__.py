"""brotli_rs_b200 -- B200-native batched Brotli decoder behind the brotli::Decompressor Read-struct surface.

The decode path is a hand-written CUDA kernel for sm_100a reached through the C ABI of include/brotli_b200.h
(libbrotli_b200.so).  This package is the thin Python host side: a ctypes binding, a batch API over device
(torch) and host (numpy) buffers, the `Decompressor` mirror of the reference's public type, and the multi-GPU
stream sharder.  There is no CPU decode path: if the CUDA library is missing or no GPU is present, decode calls
raise.
"""
from ._lib import (BroError, OK, OUTPUT_TOO_SMALL, UNEXPECTED_EOF, library_path, load_library,
                   status_description)
from .batch import BatchDecoder, MultiGpuDecoder, mg_partition, pack_streams
from .decompressor import Decompressor
from .shard import gather_outputs, scatter_batch, shard_streams

__all__ = ["BatchDecoder", "MultiGpuDecoder", "mg_partition", "Decompressor", "BroError", "pack_streams", "shard_streams", "scatter_batch", "gather_outputs", "status_description",
           "load_library", "library_path", "OK", "OUTPUT_TOO_SMALL", "UNEXPECTED_EOF"]
