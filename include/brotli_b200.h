/*
 * brotli_b200.h -- C ABI of the B200-native batched Brotli decoder (libbrotli_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of ende76/brotli-rs: decoding a Brotli stream behind
 * `brotli::Decompressor<R: Read>`.  Every entry point is plain C (pointers and sizes, no CUDA or torch types)
 * so that the reference's host language can bind it over FFI; INTEGRATION.md shows the Rust shim.
 * Citations are relative to the reference repository.
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   --------------------------------------------
 *   Decompressor::new(r)            src/lib.rs:398-410    bro_reader_new
 *   <Decompressor as Read>::read    src/lib.rs:2173-2193  bro_reader_read
 *   Decompressor::decompress        src/lib.rs:1545-2170  bro_batch_decode / bro_batch_decode_host
 *                                                         (the whole state machine runs on the GPU, many
 *                                                          streams per call)
 *   DecompressorError + description src/lib.rs:294-357    int32 status + bro_status_description
 *   drop(Decompressor)                                    bro_reader_free
 *
 * There is no CPU fallback: every decode call runs the CUDA kernel or fails with BRO_ST_CudaError.
 */
#ifndef BROTLI_B200_H
#define BROTLI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status of one stream: 0 = OK, 1..24 = the reference's DecompressorError variants in enum order
 * (src/lib.rs:294-319), >= 100 = conditions the reference cannot express.  See csrc/bro_status.h. */
#define BRO_OK 0
#define BRO_UNEXPECTED_EOF 24
#define BRO_OUTPUT_TOO_SMALL 100      /* the stream decodes to more bytes than its output slot holds */
#define BRO_CUDA_ERROR 101            /* a CUDA call failed; bro_ctx_last_cuda_error() has the text */
#define BRO_PANIC_UPPERCASE_ZERO 102  /* the reference panics here (src/transformation/mod.rs:78) */
#define BRO_SIZE_UNKNOWN 103          /* bro_batch_sizes only: the size is known only after decoding the stream */
#define BRO_INVALID_ARGUMENT 104

typedef struct bro_ctx bro_ctx;       /* one GPU, one CUDA stream's worth of scratch; not thread safe */
typedef struct bro_reader bro_reader; /* the Read-struct: one compressed stream, served from a host buffer */

/* Quirk switch: 0 = bit-exact with the reference's code (default), 1 = follow the
 * Brotli specification where the reference deviates from it. */
#define BRO_QUIRKS_REFERENCE 0
#define BRO_QUIRKS_SPEC 1

/* Create a decoder context on CUDA device `device` (-1 = current device).  Allocates the static dictionary
 * image and the work counters; the table arenas and record buffers are allocated by the first batch that needs
 * them, sized by that batch (a context that only ever decodes single streams holds about 10 MB).  Returns BRO_OK
 * or BRO_CUDA_ERROR. */
int bro_ctx_create(bro_ctx** ctx, int device);
void bro_ctx_destroy(bro_ctx* ctx);
int bro_ctx_set_quirks(bro_ctx* ctx, int quirks);
/* Path selection.
 *   WARP:     one fused kernel, one warp per stream (entropy decode and copies interleaved per command).  Lowest
 *             latency for a few streams; also the retry pass of the two-phase path.
 *   TWOPHASE: phase one decodes the entropy-coded commands with one THREAD per stream (32 streams per warp) and turns
 *             every LZ77 copy into a record; phase two executes the records with one warp per stream at memory speed.
 *             Streams phase one cannot decode (literal context modelling, libbrotli quality >= 10) are re-run by the
 *             fused kernel inside the same call.
 *   AUTO (default): TWOPHASE for batches of at least 160 streams per SM (23,680 on a B200: the fused kernel would need
 *             several waves of its resident warps), WARP below -- and also for a large batch that is bound by its longest
 *             stream (decided on the device from the compressed sizes: the two-phase kernels return at once).
 * The environment variable BRO_B200_MODE=warp|twophase sets the initial mode. */
#define BRO_MODE_AUTO 0
#define BRO_MODE_WARP 1
#define BRO_MODE_TWOPHASE 2
int bro_ctx_set_mode(bro_ctx* ctx, int mode);
const char* bro_ctx_last_cuda_error(const bro_ctx* ctx);
/* Number of kernel launches issued through this context so far (bench.py's gpu_launches). */
uint64_t bro_ctx_launch_count(const bro_ctx* ctx);
/* Resident decoder warps per launch (one stream per warp at a time). */
uint32_t bro_ctx_num_warps(const bro_ctx* ctx);

/* Measurement support (bench.py's roofline): with timing on, every bro_batch_decode records CUDA events around its
 * kernels on the launching stream; bro_ctx_last_kernel_ms waits for the last batch and returns the durations in ms of
 * {ordering kernels, parse kernel, copy kernel, fused warp kernel} (0 for kernels the batch did not launch). */
int bro_ctx_set_timing(bro_ctx* ctx, int on);
int bro_ctx_last_kernel_ms(bro_ctx* ctx, float* ms4);
/* Counters of the last batch (synchronises the device): {bytes moved by copy records, copy records executed, streams
 * handed to the fused kernel's retry pass, 1 if AUTO's gate sent the whole batch to the fused kernel}. */
int bro_ctx_last_batch_stats(bro_ctx* ctx, uint64_t* stats4);

/* Optional: an upper bound on the compressed bytes (d_in_off[n] - d_in_off[0]) of the batches that follow.  The bound is
 * STICKY: it applies to every later batch of the context until it is changed (0 = forget).  The two-phase path sizes its
 * copy-record arena from it; without it bro_batch_decode reads the two end offsets back from the device (16 bytes, blocking on the stream) before it launches.  A bound that turns out too small
 * costs speed only: streams whose records do not fit are decoded by the fused kernel. */
int bro_ctx_reserve(bro_ctx* ctx, uint64_t total_in_bytes, uint32_t n_streams);

/* THE HOT PATH.  Decode n independent streams in one call; everything is device memory.
 *   d_in       concatenated compressed streams
 *   d_in_off   n+1 byte offsets into d_in (stream i = [d_in_off[i], d_in_off[i+1]))
 *   d_out      output buffer; stream i owns the slot [d_out_off[i], d_out_off[i+1])
 *   d_out_off  n+1 byte offsets into d_out (slot capacities; a slot never receives bytes past its end)
 *   d_out_len  n: bytes produced for stream i (meaningful for BRO_OK; for a failed stream it is where the attempt stopped, which
 *              may lie beyond a slot that was too small -- bytes that did not fit were not stored -- and the slot's bytes are not
 *              part of the contract)
 *   d_status   n: status of stream i (a bad stream never affects another)
 *   stream     cudaStream_t as void* (NULL = default stream).  The call is asynchronous.
 * d_in must be readable for 16 bytes past d_in_off[n] (stored meta-blocks are moved in 16-byte granules; the host-buffer
 * entry points pad their staging buffer accordingly).  The kernels run on the context's device whatever the caller's
 * current device is.
 * Replaces Decompressor::decompress (src/lib.rs:1545-2170) and all of its helpers. */
int bro_batch_decode(bro_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint8_t* d_out,
                     const uint64_t* d_out_off, uint64_t* d_out_len, int32_t* d_status, uint32_t n, void* stream);

/* Same, with HOST buffers (pinned memory recommended): copies the inputs to the device, launches, copies the
 * slots, lengths and statuses back, and synchronises.  This is the end-to-end path a host-language caller uses. */
int bro_batch_decode_host(bro_ctx* ctx, const uint8_t* h_in, const uint64_t* h_in_off, uint8_t* h_out,
                          const uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status, uint32_t n);

/* ---- resumable decode: the building block of streaming (incremental input, bounded memory) ----
 *
 * The reference decodes incrementally: its `State` and decoder fields survive between read() calls (src/lib.rs:245-291,
 * 378-394), input is pulled as needed and only a window of output is kept (src/ringbuffer/mod.rs:8-73).  Here a stream's
 * state between two calls is a RESUME POINT, written by the decoder in front of every meta-block header (stored and
 * metadata blocks included) and at the clean end of the stream.  All-zero = start of stream. */
typedef struct bro_resume {
    uint64_t in_bits;    /* bits consumed, counted from the first byte of the input given to the call that wrote it */
    uint32_t pos;        /* bytes in the output slot in front of the resume point, history included */
    uint32_t window;     /* (1 << WBITS) - 16 once BRO_RESUME_HEADER is set (src/lib.rs:1562) */
    uint32_t dist[4];    /* distance ring, last distance first (src/lib.rs:393) */
    uint32_t p1, p2;     /* the two bytes in front of the resume point (src/lib.rs:389) */
    uint32_t flags;      /* BRO_RESUME_* */
    uint32_t reserved;
} bro_resume;
#define BRO_RESUME_HEADER 1u  /* the stream header has been consumed */
#define BRO_RESUME_LAST 2u    /* the ISLAST meta-block has been decoded: only the end-of-stream checks remain */
#define BRO_RESUME_ENDED 4u   /* the stream ended cleanly */

/* bro_batch_decode_resume: bro_batch_decode from and to d_resume[i] (device memory, n resume points).  Stream i's input
 * [d_in_off[i], d_in_off[i+1]) starts at the byte that holds bit d_resume[i].in_bits; its slot must start with the
 * history -- the last min(window, bytes produced so far) bytes of output -- in [0, d_resume[i].pos).  The call decodes
 * whole meta-blocks while input and slot last.  d_status[i]: BRO_OK = the stream ended cleanly (d_out_len[i] = bytes in
 * the slot); BRO_UNEXPECTED_EOF / BRO_OUTPUT_TOO_SMALL = the input / the slot ended inside a meta-block: the bytes in front
 * of the NEW d_resume[i].pos are final, and the call may be repeated from it with more input / room (the caller drops
 * in_bits / 8 input bytes, keeps in_bits % 8, and moves the history to the front of the slot); any other status = the
 * stream is invalid.  (For a status other than BRO_OK d_out_len[i] is where the attempt stopped -- possibly beyond the slot:
 * bytes that did not fit were not stored -- and only d_resume[i].pos says what is final.)  Always the fused warp-per-stream
 * kernel.  Asynchronous on `stream`. */
int bro_batch_decode_resume(bro_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint8_t* d_out,
                            const uint64_t* d_out_off, uint64_t* d_out_len, int32_t* d_status, bro_resume* d_resume,
                            uint32_t n, void* stream);

/* ---- decoding without knowing the uncompressed sizes (a Brotli stream does not state its size: MLEN is per meta-block,
 * src/lib.rs:469-483; the reference grows a Vec as it goes, src/lib.rs:2183-2189) ----
 *
 * bro_batch_sizes: measure n streams without writing a byte (the entropy decode of the two-phase path, copies counted
 * but not made).  d_out_len[i] = decoded size and d_status[i] = BRO_OK or the stream's error -- final, the stream need
 * not be decoded again to learn it -- or BRO_SIZE_UNKNOWN for a stream whose size only a real decode tells (literal
 * context modelling, libbrotli quality >= 10).  Asynchronous on `stream`. */
int bro_batch_sizes(bro_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint64_t* d_out_len, int32_t* d_status,
                    uint32_t n, void* stream);

/* bro_batch_decode_unsized_host: bro_batch_decode_host for a caller who does not know the sizes.  Measures the streams,
 * gives every stream an exact slot (a geometrically growing one for BRO_SIZE_UNKNOWN streams, retried while they answer
 * BRO_OUTPUT_TOO_SMALL), decodes, and returns ONE buffer allocated by the library: *h_out (release with bro_free),
 * stream i at [h_out_off[i], h_out_off[i] + h_out_len[i]).  h_out_off (n + 1), h_out_len (n), h_status (n) are the caller's. */
int bro_batch_decode_unsized_host(bro_ctx* ctx, const uint8_t* h_in, const uint64_t* h_in_off, uint32_t n, uint8_t** h_out,
                                  uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status);
void bro_free(void* p);

/* The reference's error strings, byte-identical (src/lib.rs:331-354; typos included). */
const char* bro_status_description(int status);

/* ---- the Read-struct: brotli::Decompressor<R: Read> (src/lib.rs:378-410, 2173-2193) ----
 * `cb` trampolines to R::read: it fills up to `cap` bytes and returns the count, 0 at end of input, < 0 on an
 * I/O error (which the reference folds into UnexpectedEOF, src/bitreader/mod.rs:78-82).
 * Like Decompressor::new, bro_reader_new performs no I/O.  The first bro_reader_read drains `cb`, decodes the
 * stream on the GPU (one-stream batch; the output slot grows geometrically while the status is
 * BRO_OUTPUT_TOO_SMALL) and then serves bytes from a host buffer.  Returns the number of bytes written to
 * `buf` (0 = end of stream) or -(status) when the stream is invalid. */
/* ctx == NULL (what a drop-in Decompressor::new(r) passes): the reader uses a process-wide default context of the
 * current device, created by the first read of any such reader and shared by all of them (they take turns; it lives
 * as long as the process). */
typedef intptr_t (*bro_read_cb)(void* user, uint8_t* buf, size_t cap);
bro_reader* bro_reader_new(bro_ctx* ctx, bro_read_cb cb, void* user);
intptr_t bro_reader_read(bro_reader* r, uint8_t* buf, size_t len);
int bro_reader_status(const bro_reader* r);
void bro_reader_free(bro_reader* r);

/* The same Read-struct with the reference's memory behaviour (src/lib.rs:2174-2192: input is read as it is needed, the
 * decoder keeps a window of output): `cb` is asked for `in_chunk` bytes at a time (0 = 1 MiB), every step decodes the
 * whole meta-blocks the buffered input holds (bro_batch_decode_resume, one-stream batch) and hands their bytes out, and
 * the buffers are bounded by the largest meta-block plus one window, not by the stream.  bro_reader_read fills `buf`
 * while the stream has data; bytes decoded before an error are delivered first, then -(status). */
bro_reader* bro_reader_new_streaming(bro_ctx* ctx, bro_read_cb cb, void* user, size_t in_chunk);

/* ---- one batch on several GPUs of one box ----
 * Streams are independent, so a batch shards by stream with no exchange between GPUs: bro_mg_create makes one context
 * per device (devices 0 .. ngpus-1; ngpus <= 0 = every visible device); bro_mg_decode_host is bro_batch_decode_host over
 * all of them -- the batch is cut into one contiguous range of streams per device with (nearly) equal compressed + slot
 * bytes (bro_mg_partition: first[k] .. first[k+1] are device k's streams, first has ngpus + 1 entries), and every range
 * is decoded by its own host thread with its own context, copies and kernels of the devices overlapping.  Same
 * results, same argument meaning as bro_batch_decode_host.  bro_mg_ctx gives access to device k's context (mode,
 * quirks, counters). */
typedef struct bro_mg bro_mg;
int bro_mg_create(bro_mg** mg, int ngpus);
void bro_mg_destroy(bro_mg* mg);
int bro_mg_device_count(const bro_mg* mg);
bro_ctx* bro_mg_ctx(bro_mg* mg, int k);
int bro_mg_partition(const uint64_t* h_in_off, const uint64_t* h_out_off, uint32_t n, int ngpus, uint32_t* first);
int bro_mg_decode_host(bro_mg* mg, const uint8_t* h_in, const uint64_t* h_in_off, uint8_t* h_out,
                       const uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* BROTLI_B200_H */
