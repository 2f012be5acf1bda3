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
