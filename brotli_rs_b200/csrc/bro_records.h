// bro_records.h -- the copy record shared by the two phases of the two-phase path and the host side.
#pragma once
#include <stdint.h>

// A copy record of the two-phase path: phase one (bro_parse.h, one thread per stream) decodes the entropy-coded
// commands and writes literals and dictionary words straight into the output slot; every LZ77 back-reference and
// every stored meta-block becomes one record, which phase two (bro_kernels_copy.cu, one warp per stream) executes.
struct alignas(16) BroRec {
    uint32_t dst;             // output position of the first byte, relative to the slot
    uint32_t len_kind;        // length (< 2^25) | kind << 28
    uint32_t a;               // BRO_REC_LZ: distance; BRO_REC_STORED: offset of the source bytes in the compressed stream
    uint32_t b;               // reserved
};
#define BRO_REC_KIND_SHIFT 28u
#define BRO_REC_LEN_MASK 0x0fffffffu
#define BRO_REC_LZ 0u
#define BRO_REC_STORED 1u
#define BRO_REC_PIECE_VECS 32u    /* a record moves at most this many aligned 16-byte vectors (+ < 16 ragged bytes at each end),
                                    unless it is a periodic fill (distance < length, distance shorter than a piece) */

// Stream i owns records [BRO_REC_BASE(in_off, i), BRO_REC_BASE(in_off, i + 1)) of the record arena: one record per 2
// compressed bytes + 32.  A stream that needs more is handed to the fused kernel (BRO_ST_RecordsFull).
#define BRO_REC_BASE(in_off, i) ((((in_off)[i] - (in_off)[0]) >> 1) + 32ull * (i))

// Resume point of a stream between two calls of the resumable decode (bro_batch_decode_resume; the reference keeps the
// equivalent in `State` + the decoder's fields, src/lib.rs:245-291, 378-394): the position of the next unread bit and
// everything that survives a meta-block boundary.  The decoder writes one at the stream header and before every
// meta-block header (stored and metadata blocks included); a call that runs out of input or output inside a meta-block
// is repeated from the last one with more input / room.  All-zero = start of stream.
struct BroResume {
    uint64_t in_bits;         // bits consumed, counted from the first byte of the input given to THIS call
    uint32_t pos;             // bytes in the output slot in front of the resume point (history included)
    uint32_t window;          // (1 << WBITS) - 16, valid once BRO_RESUME_HEADER is set
    uint32_t dist[4];         // distance ring, last distance first (src/lib.rs:393)
    uint32_t p1, p2;          // the two bytes before the resume point (src/lib.rs:389)
    uint32_t flags;
    uint32_t reserved;
};
#define BRO_RESUME_HEADER 1u  /* the stream header has been consumed */
#define BRO_RESUME_LAST 2u    /* the ISLAST meta-block has been decoded: only the end-of-stream checks remain */
#define BRO_RESUME_ENDED 4u   /* the stream ended cleanly */
