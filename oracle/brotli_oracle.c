/*
 * brotli_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * A plain-C, single-threaded CPU restatement of the decode algorithm of ende76/brotli-rs v0.3.23
 * (the reference).  It exists to *check* the CUDA decoder: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  It is never linked into
 * libbrotli_b200.so and is not a fallback for anything.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle_*.py) against every golden vector the
 * reference holds for the path -- the 43 valid + 9 invalid data/ streams, the 33 stream tests of
 * tests/lib.rs, the doc-test, the 121 transform KATs, the 13 bit-reader KATs, the tree KATs and the IMTF
 * vectors -- and cross-checked on valid streams against the system libbrotlidec.
 * The reference itself (Rust) cannot be built in this image (no rustc/cargo), so corners that the
 * reference's own tests do not exercise follow the reference's source text and
 * are "parity unpinned".
 *
 * It deliberately keeps the reference's cost model, because it doubles as the CPU baseline "port":
 *   - bit-at-a-time reads with the (bit_pos, current_byte) state      src/bitreader/mod.rs:178-232
 *   - prefix codes as implicit-heap arrays of 2^(maxlen+1)-1 slots    src/huffman/tree/mod.rs:33-92
 *   - a ring-buffer window with a modulo per byte                      src/ringbuffer/mod.rs:50-73
 * Every function cites the reference lines it follows.  All citations are relative to /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../brotli_rs_b200/csrc/bro_tables_generated.h"

extern const uint8_t bro_dictionary_blob[]; /* oracle/dict_blob.c (.incbin of data/dictionary.bin) */

/* Status codes: 0 = OK, 1..24 = DecompressorError in enum order (src/lib.rs:294-319). */
enum {
    ST_OK = 0,
    ST_CodeLengthsChecksum = 1, ST_ExpectedEndOfStream, ST_ExceededExpectedBytes, ST_InvalidBlockCountCode,
    ST_InvalidBlockSwitchCommandCode, ST_InvalidLengthInStaticDictionary, ST_InvalidMSkipLen, ST_InvalidSymbol,
    ST_InvalidTransformId, ST_InvalidNonPositiveDistance, ST_LessThanTwoNonZeroCodeLengths, ST_NoCodeLength,
    ST_NonZeroFillBit, ST_NonZeroReservedBit, ST_NonZeroTrailerBit, ST_NonZeroTrailerNibble,
    ST_ParseErrorContextMap, ST_ParseErrorComplexPrefixCodeLengths, ST_ParseErrorDistanceCode,
    ST_ParseErrorInsertAndCopyLength, ST_ParseErrorInsertLiterals, ST_RingBufferError,
    ST_RunLengthExceededSizeOfContextMap, ST_UnexpectedEOF,
    /* things the reference cannot express as an error value */
    ST_OutputTooSmall = 100,       /* caller-provided slot too small (batch API only)            */
    ST_PanicUppercaseZero = 102,   /* reference hits unreachable!() src/transformation/mod.rs:78 */
    ST_OutOfMemory = 103
};

/* src/lib.rs:331-354 -- byte-identical strings, typos included. */
const char* bro_oracle_status_description(int st) {
    switch (st) {
    case ST_OK: return "OK";
    case ST_CodeLengthsChecksum: return "Code length check sum did not add up in complex prefix code";
    case ST_ExpectedEndOfStream: return "Expected end-of-stream, but stream did not end";
    case ST_ExceededExpectedBytes: return "More uncompressed bytes than expected in meta-block";
    case ST_InvalidBlockCountCode: return "Encountered invalid value for block count code";
    case ST_InvalidBlockSwitchCommandCode: return "Encountered invalid value for block switch command code";
    case ST_InvalidLengthInStaticDictionary: return "Encountered invalid length in reference to static dictionary";
    case ST_InvalidMSkipLen: return "Most significant byte of MSKIPLEN was zero";
    case ST_InvalidSymbol: return "Encountered invalid symbol in prefix code";
    case ST_InvalidTransformId: return "Encountered invalid transform id in reference to static dictionary";
    case ST_InvalidNonPositiveDistance: return "Encountered invalid non-positive distance";
    case ST_LessThanTwoNonZeroCodeLengths: return "Encountered invalid complex prefix code with less than two non-zero codelengths";
    case ST_NoCodeLength: return "Encountered invalid complex prefix code with all zero codelengths";
    case ST_NonZeroFillBit: return "Enocuntered non-zero fill bit";
    case ST_NonZeroReservedBit: return "Enocuntered non-zero reserved bit";
    case ST_NonZeroTrailerBit: return "Enocuntered non-zero bit trailing the stream";
    case ST_NonZeroTrailerNibble: return "Enocuntered non-zero nibble trailing";
    case ST_ParseErrorContextMap: return "Error parsing context map";
    case ST_ParseErrorComplexPrefixCodeLengths: return "Error parsing code lengths for complex prefix code";
    case ST_ParseErrorDistanceCode: return "Error parsing DistanceCode";
    case ST_ParseErrorInsertAndCopyLength: return "Error parsing Insert And Copy Length";
    case ST_ParseErrorInsertLiterals: return "Error parsing Insert Literals";
    case ST_RingBufferError: return "Error accessing distance ring buffer";
    case ST_RunLengthExceededSizeOfContextMap: return "Run length excceeded declared length of context map";
    case ST_UnexpectedEOF: return "Encountered unexpected EOF";
    case ST_OutputTooSmall: return "output slot too small";
    case ST_PanicUppercaseZero: return "reference panics: uppercase_first on a word starting with 0x00";
    case ST_OutOfMemory: return "oracle out of memory";
    default: return "unknown status";
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* BitReader -- src/bitreader/mod.rs:21-304.  Returns >= 0 value, BR_ERR generic failure, BR_EOF.     */
/* ------------------------------------------------------------------------------------------------ */
#define BR_ERR (-1)
#define BR_EOF (-2)

typedef struct {
    const uint8_t* p;
    size_t n, idx;
    uint8_t bit_pos;      /* src/bitreader/mod.rs:23 */
    int has_cur;          /* current_byte: Option<u8>, src/bitreader/mod.rs:24 */
    uint8_t cur;
} BitReader;

static void br_init(BitReader* br, const uint8_t* p, size_t n) {
    br->p = p; br->n = n; br->idx = 0; br->bit_pos = 0; br->has_cur = 0; br->cur = 0;
}

/* src/bitreader/mod.rs:178-203 (read_bit) and 207-232 (read_bit_as_usize): identical state updates. */
static int br_read_bit(BitReader* br) {
    if (br->has_cur) {
        uint8_t bit_pos = br->bit_pos;
        br->bit_pos = (uint8_t)((br->bit_pos + 1) % 8);
        if (br->bit_pos == 0) br->has_cur = 0;
        return (br->cur >> bit_pos) & 1;
    }
    if (br->idx >= br->n) return BR_ERR;
    br->cur = br->p[br->idx++];
    br->has_cur = 1;
    br->bit_pos = 1;
    return br->cur & 1;
}

/* src/bitreader/mod.rs:58-84 */
static int br_read_u8(BitReader* br) {
    int ok = br->idx < br->n;
    uint8_t nb = ok ? br->p[br->idx++] : 0;
    if (br->has_cur && ok) {
        uint8_t byte = br->cur;
        br->cur = nb;
        return (uint8_t)((byte >> br->bit_pos) | (uint8_t)(nb << (8 - br->bit_pos)));
    }
    if (!br->has_cur && ok) return nb;
    if (br->bit_pos == 0 && br->has_cur && !ok) { br->has_cur = 0; return br->cur; }
    return BR_EOF;
}

/* src/bitreader/mod.rs:140-158, 237-252, 272-288: n bits, least significant first. */
static int64_t br_read_bits(BitReader* br, unsigned n) {
    uint32_t v = 0;
    for (unsigned i = 0; i < n; i++) {
        int b = br_read_bit(br);
        if (b < 0) return BR_ERR;
        if (b) v |= (1u << i);
    }
    return (int64_t)v;
}

/* src/bitreader/mod.rs:88-135 */
static int br_read_nibble(BitReader* br) {
    if (br->bit_pos == 0 && !br->has_cur) {
        if (br->idx >= br->n) return BR_EOF;
        br->cur = br->p[br->idx++];
        br->bit_pos = 4;
        br->has_cur = 1;
        return br->cur & 0x0f;
    }
    if (br->bit_pos <= 3) {
        br->bit_pos = (uint8_t)(br->bit_pos + 4);
        return (br->cur >> (br->bit_pos - 4)) & 0x0f;
    }
    if (br->bit_pos == 4) {
        uint8_t byte = br->cur;
        br->bit_pos = 0;
        br->has_cur = 0;
        return (byte >> 4) & 0x0f;
    }
    {
        uint8_t bit_pos = br->bit_pos, byte = br->cur;
        if (br->idx >= br->n) return BR_EOF;
        uint8_t nb = br->p[br->idx++];
        br->bit_pos = (uint8_t)(br->bit_pos - 4);
        br->cur = nb;
        return ((byte >> bit_pos) | (uint8_t)(nb << (8 - bit_pos))) & 0x0f;
    }
}

/* src/bitreader/mod.rs:163-174 */
static int64_t br_read_nibbles(BitReader* br, unsigned n) {
    uint32_t v = 0;
    for (unsigned i = 0; i < n; i++) {
        int nb = br_read_nibble(br);
        if (nb < 0) return nb;
        v |= ((uint32_t)nb) << (4 * i);
    }
    return (int64_t)v;
}

/* src/bitreader/mod.rs:257-267 */
static int br_read_byte_tail(BitReader* br) {
    if (br->bit_pos == 0) return 0;
    return (int)br_read_bits(br, 8u - br->bit_pos);
}

/* ------------------------------------------------------------------------------------------------ */
/* Tree -- src/huffman/tree/mod.rs:7-92; slots hold the symbol or -1 for None.                        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t* buf;
    size_t buflen;
    size_t len;          /* number of inserts */
    int32_t last_symbol; /* -1 = None */
} Tree;

static void tree_free(Tree* t) { free(t->buf); t->buf = NULL; t->buflen = 0; t->len = 0; t->last_symbol = -1; }

/* src/huffman/tree/mod.rs:34-40 */
static int tree_with_max_depth(Tree* t, unsigned max_depth) {
    t->buflen = ((size_t)1 << (max_depth + 1)) - 1;
    t->buf = (int32_t*)malloc(t->buflen * sizeof(int32_t));
    if (!t->buf) return -1;
    for (size_t i = 0; i < t->buflen; i++) t->buf[i] = -1;
    t->len = 0;
    t->last_symbol = -1;
    return 0;
}

/* src/huffman/tree/mod.rs:50-61; code given as (value, length) with the first bit read most significant. */
static void tree_insert(Tree* t, uint32_t code, unsigned len, int32_t symbol) {
    t->len += 1;
    t->last_symbol = symbol;
    size_t at = (((size_t)1 << len) - 1) + code;
    if (at > t->buflen - 1) abort(); /* reference panics; unreachable for Kraft-valid lengths */
    t->buf[at] = symbol;
}

#define LOOKUP_NONE (-3)
/* src/huffman/tree/mod.rs:63-92: returns symbol, LOOKUP_NONE for Ok(None), BR_ERR for a failed bit read. */
static int32_t tree_lookup_symbol(const Tree* t, BitReader* br) {
    if (t->len == 0) return LOOKUP_NONE;
    if (t->len == 1) return t->last_symbol;
    size_t pseudo_code = 1;
    for (;;) {
        int b = br_read_bit(br);
        if (b < 0) return BR_ERR;
        pseudo_code = (pseudo_code << 1) + (size_t)b;
        size_t idx = pseudo_code - 1;
        if (idx > t->buflen - 1) return LOOKUP_NONE;
        if (t->buf[idx] >= 0) return t->buf[idx];
    }
}

/* src/huffman/mod.rs:19-43 (incl. the bl_count[0] quirk, SURVEY Q6: garbage high bits are masked off by
 * bit_string_from_code_and_length, src/huffman/mod.rs:3-11). */
static int codes_from_lengths_and_symbols(Tree* t, const unsigned* lengths, const uint16_t* symbols, size_t n) {
    unsigned max_length = 0;
    for (size_t i = 0; i < n; i++) if (lengths[i] > max_length) max_length = lengths[i];
    size_t bl_count[17] = {0}, next_code[17] = {0};
    for (size_t i = 0; i < n; i++) bl_count[lengths[i]] += 1;
    size_t code = 0;
    for (unsigned bits = 1; bits <= max_length; bits++) {
        code = (code + bl_count[bits - 1]) << 1;
        next_code[bits] = code;
    }
    if (tree_with_max_depth(t, max_length)) return -1;
    for (size_t i = 0; i < n; i++) {
        unsigned len = lengths[i];
        if (len > 0 || max_length == 0) {
            uint32_t masked = len ? (uint32_t)(next_code[len] & (((size_t)1 << len) - 1)) : 0;
            tree_insert(t, masked, len, symbols[i]);
            next_code[len] += 1;
        }
    }
    return 0;
}

/* src/huffman/mod.rs:45-49 */
static int codes_from_lengths(Tree* t, const unsigned* lengths, size_t n) {
    uint16_t* symbols = (uint16_t*)malloc(n * sizeof(uint16_t));
    if (!symbols) return -1;
    for (size_t i = 0; i < n; i++) symbols[i] = (uint16_t)i;
    int r = codes_from_lengths_and_symbols(t, lengths, symbols, n);
    free(symbols);
    return r;
}

/* Fixed trees of Header::new, src/lib.rs:86-135, written as (read-order bit string -> symbol) pairs and
 * inserted into the same heap layout (SURVEY appendix A lists the decoded codes). */
static int fixed_tree(Tree* t, unsigned depth, const char* const* codes, const int* syms, size_t n, size_t len, int last) {
    if (tree_with_max_depth(t, depth)) return -1;
    for (size_t i = 0; i < n; i++) {
        uint32_t c = 0; unsigned l = 0;
        for (const char* s = codes[i]; *s; s++, l++) c = (c << 1) | (uint32_t)(*s == '1');
        t->buf[(((size_t)1 << l) - 1) + c] = syms[i];
    }
    t->len = len;            /* from_raw_data(.., len, last_symbol): src/lib.rs:119,125,131 */
    t->last_symbol = last;
    return 0;
}

static int make_wbits_tree(Tree* t) {
    static const char* const c[] = {"0", "1000000", "1100", "1010", "1110", "1001", "1101", "1011", "1111",
                                    "1000010", "1000110", "1000001", "1000101", "1000011", "1000111"};
    static const int s[] = {16, 17, 18, 19, 20, 21, 22, 23, 24, 10, 11, 12, 13, 14, 15};
    return fixed_tree(t, 7, c, s, 15, 15, 24);
}
static int make_code_length_tree(Tree* t) {
    static const char* const c[] = {"00", "01", "10", "110", "1110", "1111"};
    static const int s[] = {0, 3, 4, 2, 1, 5};
    return fixed_tree(t, 4, c, s, 6, 6, 5);
}
static int make_bltype_tree(Tree* t) {
    static const char* const c[] = {"0", "1000", "1100", "1010", "1110", "1001", "1101", "1011", "1111"};
    static const int s[] = {1, 2, 3, 5, 9, 17, 33, 65, 129};
    return fixed_tree(t, 4, c, s, 9, 9, 129);
}

/* ------------------------------------------------------------------------------------------------ */
/* RingBuffer<u8> window -- src/ringbuffer/mod.rs:8-73                                               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { uint8_t* buf; size_t len, pos, cap; } Ring;

static void ring_push(Ring* r, uint8_t item) { /* src/ringbuffer/mod.rs:64-73 */
    if (r->len < r->cap) { r->buf[r->len] = item; r->pos = r->len; r->len += 1; }
    else { r->pos = (r->pos + 1) % r->len; r->buf[r->pos] = item; }
}
static int ring_slice_tail(const Ring* r, size_t n, uint8_t* out, size_t outlen) { /* :50-61 */
    size_t len = r->len;
    if (n >= len) return -1;
    for (size_t i = 0; i < outlen; i++) out[i] = r->buf[(r->pos + len - n + i) % len];
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Word transforms -- src/transformation/mod.rs:3-209                                                */
/* ------------------------------------------------------------------------------------------------ */
/* src/transformation/mod.rs:3-40 */
static size_t uppercase_all(const uint8_t* w, size_t l, uint8_t* v) {
    size_t i = 0, o = 0;
    while (i < l) {
        uint8_t c = w[i];
        if (c <= 96 || (c >= 123 && c <= 191)) { v[o++] = c; i += 1; }
        else if (c <= 122) { v[o++] = c ^ 32; i += 1; }
        else if (c <= 223) { v[o++] = c; if (i + 1 < l) v[o++] = w[i + 1] ^ 32; i += 2; }
        else { v[o++] = c; if (i + 1 < l) v[o++] = w[i + 1]; if (i + 2 < l) v[o++] = w[i + 2] ^ 5; i += 3; }
    }
    return o;
}

/* src/transformation/mod.rs:42-82; returns (size_t)-1 where the reference reaches unreachable!() (byte 0). */
static size_t uppercase_first(const uint8_t* w, size_t l, uint8_t* v, int quirk_spec) {
    if (l == 0) return 0;
    size_t o = 0, i;
    uint8_t c = w[0];
    if (c == 0) {
        if (!quirk_spec) return (size_t)-1;
        v[o++] = c; i = 1;
    } else if (c <= 96 || (c >= 123 && c <= 191)) { v[o++] = c; i = 1; }
    else if (c <= 122) { v[o++] = c ^ 32; i = 1; }
    else if (c <= 223) { v[o++] = c; if (1 < l) v[o++] = w[1] ^ 32; i = 2; }
    else { v[o++] = c; if (1 < l) v[o++] = w[1]; if (2 < l) v[o++] = w[2] ^ 5; i = 3; }
    for (; i < l; i++) v[o++] = w[i];
    return o;
}

/* src/transformation/mod.rs:84-209, table-driven from spec appendix B.  OmitFirstN keeps the last byte of
 * a word not longer than N (base_word[min(N, len-1)..], e.g. :89, :120) unless quirk_spec. */
static long transformation(unsigned id, const uint8_t* w, size_t l, uint8_t* out, int quirk_spec) {
    size_t o = 0;
    unsigned type = bro_xf_type[id];
    memcpy(out + o, bro_xf_strings + bro_xf_prefix_off[id], bro_xf_prefix_len[id]);
    o += bro_xf_prefix_len[id];
    if (type == 0) { memcpy(out + o, w, l); o += l; }
    else if (type == 1) {
        size_t k = uppercase_first(w, l, out + o, quirk_spec);
        if (k == (size_t)-1) return -1;
        o += k;
    } else if (type == 2) { o += uppercase_all(w, l, out + o); }
    else if (type <= 11) {
        size_t n = type - 2;
        size_t from = quirk_spec ? (n < l ? n : l) : (n < l - 1 ? n : l - 1);
        memcpy(out + o, w + from, l - from); o += l - from;
    } else {
        size_t n = type - 11;
        size_t keep = (l > n ? l : n) - n;
        memcpy(out + o, w, keep); o += keep;
    }
    memcpy(out + o, bro_xf_strings + bro_xf_suffix_off[id], bro_xf_suffix_len[id]);
    o += bro_xf_suffix_len[id];
    return (long)o;
}

/* ------------------------------------------------------------------------------------------------ */
/* Decompressor state -- src/lib.rs:378-394 plus the MetaBlock fields of src/lib.rs:139-166           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint8_t* data; size_t len, cap; int growable;
} Sink;

typedef struct {
    BitReader in;
    Tree wbits_codes, bit_lengths_code, bltype_codes;   /* Header, src/lib.rs:77-83 */
    size_t window_size;
    Ring window;
    size_t count_output;
    uint8_t lit_buf[2];   /* literal_buf: [0] = p1 (last), [1] = p2 */
    uint32_t dist_buf[4]; /* distance_buf: [0] = last */
    Sink out;
    int quirk_spec;       /* 0 = follow the reference (default), 1 = follow the specification on Q1/Q3/Q4 */
} Dec;

static int sink_push(Dec* d, uint8_t b) {
    Sink* s = &d->out;
    if (s->len == s->cap) {
        if (!s->growable) return ST_OutputTooSmall;
        size_t nc = s->cap ? s->cap * 2 : 4096;
        uint8_t* nd = (uint8_t*)realloc(s->data, nc);
        if (!nd) return ST_OutOfMemory;
        s->data = nd; s->cap = nc;
    }
    s->data[s->len++] = b;
    return 0;
}

static void litbuf_push(Dec* d, uint8_t b) { d->lit_buf[1] = d->lit_buf[0]; d->lit_buf[0] = b; }
static void distbuf_push(Dec* d, uint32_t v) {
    d->dist_buf[3] = d->dist_buf[2]; d->dist_buf[2] = d->dist_buf[1]; d->dist_buf[1] = d->dist_buf[0]; d->dist_buf[0] = v;
}

/* src/lib.rs:501-525 */
static int parse_n_bltypes(Dec* d, unsigned* out) {
    int32_t s = tree_lookup_symbol(&d->bltype_codes, &d->in);
    if (s < 0) return ST_UnexpectedEOF;
    unsigned extra_bits;
    switch (s) {
    case 1: case 2: extra_bits = 0; break;
    case 3: extra_bits = 1; break;
    case 5: extra_bits = 2; break;
    case 9: extra_bits = 3; break;
    case 17: extra_bits = 4; break;
    case 33: extra_bits = 5; break;
    case 65: extra_bits = 6; break;
    default: extra_bits = 7; break;
    }
    unsigned v = (unsigned)s;
    if (extra_bits > 0) {
        int64_t e = br_read_bits(&d->in, extra_bits);
        if (e < 0) return ST_UnexpectedEOF;
        v += (unsigned)e;
    }
    *out = v;
    return 0;
}

static void sort_u16(uint16_t* a, size_t n) { /* tiny insertion sort (n <= 4) */
    for (size_t i = 1; i < n; i++) { uint16_t k = a[i]; size_t j = i; while (j > 0 && a[j - 1] > k) { a[j] = a[j - 1]; j--; } a[j] = k; }
}

/* src/lib.rs:597-665 */
static int parse_simple_prefix_code(Dec* d, size_t alphabet_size, Tree* t) {
    unsigned bit_width = 0;
    { uint16_t a = (uint16_t)(alphabet_size - 1); while (a) { bit_width++; a >>= 1; } }
    int64_t v = br_read_bits(&d->in, 2);
    if (v < 0) return ST_UnexpectedEOF;
    size_t n_sym = (size_t)v + 1;
    uint16_t symbols[4];
    for (size_t i = 0; i < n_sym; i++) {
        int64_t s = br_read_bits(&d->in, bit_width);
        if (s < 0) return ST_UnexpectedEOF;
        if ((size_t)s >= alphabet_size) return ST_InvalidSymbol;
        symbols[i] = (uint16_t)s;
    }
    for (size_t i = 0; i + 1 < n_sym; i++)
        for (size_t j = i + 1; j < n_sym; j++)
            if (symbols[i] == symbols[j]) return ST_InvalidSymbol;
    int tree_select = -1;
    if (n_sym == 4) {
        int b = br_read_bit(&d->in);
        if (b < 0) return ST_UnexpectedEOF;
        tree_select = b;
    }
    unsigned lengths[4];
    switch (n_sym) {
    case 1: lengths[0] = 0; break;
    case 2: sort_u16(symbols, 2); lengths[0] = 1; lengths[1] = 1; break;
    case 3: sort_u16(symbols + 1, 2); lengths[0] = 1; lengths[1] = 2; lengths[2] = 2; break;
    default:
        if (!tree_select) { sort_u16(symbols, 4); lengths[0] = lengths[1] = lengths[2] = lengths[3] = 2; }
        else { sort_u16(symbols + 2, 2); lengths[0] = 1; lengths[1] = 2; lengths[2] = 3; lengths[3] = 3; }
        break;
    }
    return codes_from_lengths_and_symbols(t, lengths, symbols, n_sym) ? ST_OutOfMemory : 0;
}

/* src/lib.rs:667-875 */
static int parse_complex_prefix_code(Dec* d, unsigned h_skip, size_t alphabet_size, Tree* t) {
    unsigned code_lengths[18] = {0};
    size_t sum = 0, nonzero = 0;
    for (unsigned i = h_skip; i < 18; i++) {
        int32_t cl = tree_lookup_symbol(&d->bit_lengths_code, &d->in);
        if (cl == LOOKUP_NONE) return ST_ParseErrorComplexPrefixCodeLengths;
        if (cl < 0) return ST_UnexpectedEOF;
        code_lengths[i] = (unsigned)cl;
        if (cl > 0) {
            sum += 32u >> cl;
            nonzero += 1;
            if (sum == 32) break;
            if (sum > 32) return ST_CodeLengthsChecksum;
        }
    }
    if (nonzero == 0) return ST_NoCodeLength;
    if (nonzero >= 2 && sum < 32) return ST_CodeLengthsChecksum;

    /* src/lib.rs:719-720: reorder from transmission order 1,2,3,4,0,5,17,6,16,7,... to symbols 0..17 */
    static const unsigned order[18] = {4, 0, 1, 2, 3, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 8, 6};
    unsigned cl_sorted[18];
    uint16_t syms18[18];
    for (unsigned i = 0; i < 18; i++) { cl_sorted[i] = code_lengths[order[i]]; syms18[i] = (uint16_t)i; }
    Tree clc;
    if (codes_from_lengths_and_symbols(&clc, cl_sorted, syms18, 18)) return ST_OutOfMemory;

    unsigned* actual = (unsigned*)calloc(alphabet_size, sizeof(unsigned));
    if (!actual) { tree_free(&clc); return ST_OutOfMemory; }
    int st = 0;
    size_t sum2 = 0;
    int last_symbol = -1;       /* Option<u16> */
    long last_repeat = -1;      /* Option<usize> */
    unsigned last_non_zero = 8;
    size_t i = 0;
    while (i < alphabet_size) {
        int32_t c = tree_lookup_symbol(&clc, &d->in);
        if (c == BR_ERR) { st = ST_UnexpectedEOF; goto done; }
        if (c == LOOKUP_NONE) { st = ST_ParseErrorComplexPrefixCodeLengths; goto done; }
        if (c <= 15) {
            actual[i] = (unsigned)c;
            i += 1;
            last_symbol = c;
            last_repeat = -1;
            if (c > 0) {
                last_non_zero = (unsigned)c;
                sum2 += 32768u >> c;
                if (sum2 == 32768) break;
                else if (sum2 > 32768) { st = ST_CodeLengthsChecksum; goto done; }
            }
        } else if (c == 16) {
            int64_t eb = br_read_bits(&d->in, 2);
            if (eb < 0) { st = ST_UnexpectedEOF; goto done; }
            size_t extra = (size_t)eb, count, newrep;
            if (last_symbol == 16 && last_repeat >= 0) {
                newrep = 4 * ((size_t)last_repeat - 2) + extra + 3;
                if (i + newrep - (size_t)last_repeat > alphabet_size) { st = ST_ParseErrorComplexPrefixCodeLengths; goto done; }
                count = newrep - (size_t)last_repeat;
            } else {
                newrep = 3 + extra;
                if (i + newrep > alphabet_size) { st = ST_ParseErrorComplexPrefixCodeLengths; goto done; }
                count = newrep;
            }
            for (size_t k = 0; k < count; k++) { actual[i] = last_non_zero; i += 1; sum2 += 32768u >> last_non_zero; }
            last_repeat = (long)newrep;
            if (sum2 == 32768) break;             /* note: break leaves last_symbol unset, irrelevant */
            else if (sum2 > 32768) { st = ST_CodeLengthsChecksum; goto done; }
            last_symbol = 16;
        } else { /* 17 */
            int64_t eb = br_read_bits(&d->in, 3);
            if (eb < 0) { st = ST_UnexpectedEOF; goto done; }
            size_t extra = (size_t)eb;
            if (last_symbol == 17 && last_repeat >= 0) {
                size_t newrep = 8 * ((size_t)last_repeat - 2) + extra + 3;
                i += newrep - (size_t)last_repeat;
                last_repeat = (long)newrep;
            } else {
                size_t rep = 3 + extra;
                i += rep;
                last_repeat = (long)rep;
            }
            if (i > alphabet_size) { st = ST_ParseErrorComplexPrefixCodeLengths; goto done; }
            last_symbol = 17;
        }
    }
    {
        size_t nz = 0;
        for (size_t k = 0; k < alphabet_size; k++) if (actual[k] > 0) nz++;
        if (nz < 2) { st = ST_LessThanTwoNonZeroCodeLengths; goto done; }
    }
    if (codes_from_lengths(t, actual, alphabet_size)) st = ST_OutOfMemory;
done:
    free(actual);
    tree_free(&clc);
    return st;
}

/* src/lib.rs:589-595, 877-889 */
static int parse_prefix_code(Dec* d, size_t alphabet_size, Tree* t) {
    int64_t k = br_read_bits(&d->in, 2);
    if (k < 0) return ST_UnexpectedEOF;
    if (k == 1) return parse_simple_prefix_code(d, alphabet_size, t);
    return parse_complex_prefix_code(d, (unsigned)k, alphabet_size, t);
}

/* src/lib.rs:957-987 */
static int parse_block_count(Dec* d, const Tree* t, uint32_t* out) {
    int32_t s = tree_lookup_symbol(t, &d->in);
    if (s == BR_ERR || s == LOOKUP_NONE) return ST_UnexpectedEOF;
    if (s > 25) return ST_InvalidBlockCountCode;
    uint32_t base = bro_block_count[s] & 0xffff, extra = bro_block_count[s] >> 16;
    int64_t e = br_read_bits(&d->in, extra);
    if (e < 0) return ST_UnexpectedEOF;
    *out = base + (uint32_t)e;
    return 0;
}

/* src/lib.rs:1164-1177 */
static void inverse_move_to_front_transform(uint8_t* v, size_t n) {
    uint8_t mtf[256];
    for (unsigned i = 0; i < 256; i++) mtf[i] = (uint8_t)i;
    for (size_t k = 0; k < n; k++) {
        unsigned index = v[k];
        uint8_t value = mtf[index];
        v[k] = value;
        for (unsigned j = index; j >= 1; j--) mtf[j] = mtf[j - 1];
        mtf[0] = value;
    }
}

/* src/lib.rs:1070-1144 */
static int parse_context_map(Dec* d, unsigned n_trees, size_t len, uint8_t* c_map) {
    int b = br_read_bit(&d->in);
    if (b < 0) return ST_UnexpectedEOF;
    unsigned rlemax = 0;
    if (b) {
        int64_t v = br_read_bits(&d->in, 4);
        if (v < 0) return ST_UnexpectedEOF;
        rlemax = (unsigned)v + 1;
    }
    Tree t;
    int st = parse_prefix_code(d, rlemax + n_trees, &t);
    if (st) return st;
    size_t pushed = 0;
    while (pushed < len) {
        int32_t s = tree_lookup_symbol(&t, &d->in);
        if (s == BR_ERR) { st = ST_UnexpectedEOF; goto done; }
        if (s == LOOKUP_NONE) { st = ST_ParseErrorContextMap; goto done; }
        if (s > 0 && (unsigned)s <= rlemax) {
            int64_t e = br_read_bits(&d->in, (unsigned)s);
            if (e < 0) { st = ST_UnexpectedEOF; goto done; }
            uint32_t repeat = (1u << s) + (uint32_t)e;
            for (uint32_t k = 0; k < repeat; k++) {
                if (pushed + 1 > len) { st = ST_RunLengthExceededSizeOfContextMap; goto done; }
                c_map[pushed++] = 0;
            }
        } else {
            c_map[pushed++] = (uint8_t)(s == 0 ? 0 : (unsigned)s - rlemax);
        }
    }
    b = br_read_bit(&d->in);
    if (b < 0) { st = ST_UnexpectedEOF; goto done; }
    if (b) inverse_move_to_front_transform(c_map, len);
done:
    tree_free(&t);
    return st;
}

/* Block-switch state for one category (L, I, D): src/lib.rs:152-160 */
typedef struct {
    unsigned n_bltypes;
    Tree types, counts;
    int has_trees;
    unsigned btype, btype_prev;
    int has_blen; uint32_t blen;
} BlockCat;

/* src/lib.rs:1226-1250 */
static int parse_block_switch_command(Dec* d, BlockCat* c) {
    int32_t code = tree_lookup_symbol(&c->types, &d->in);
    if (code == LOOKUP_NONE) return ST_InvalidBlockSwitchCommandCode;
    if (code < 0) return ST_UnexpectedEOF;
    unsigned block_type = code == 0 ? c->btype_prev : code == 1 ? (c->btype + 1) % c->n_bltypes : (unsigned)code - 2;
    uint32_t count;
    int st = parse_block_count(d, &c->counts, &count);
    if (st) return st;
    c->btype_prev = c->btype;
    c->btype = block_type;
    c->blen = count - 1;
    return 0;
}

/* The blen bookkeeping shared by src/lib.rs:1182-1197, 1294-1306, 1377-1389 */
static int step_block(Dec* d, BlockCat* c) {
    if (!c->has_blen) return 0;
    if (c->blen == 0) return parse_block_switch_command(d, c);
    c->blen -= 1;
    return 0;
}

typedef struct {
    int is_last;
    uint32_t m_len;
    size_t count_output;
    BlockCat cat[3]; /* L, I, D */
    unsigned n_postfix, n_direct;
    uint8_t* context_modes;
    unsigned n_trees_l, n_trees_d;
    uint8_t* c_map_l; uint8_t* c_map_d;
    Tree* trees_l; Tree* trees_i; Tree* trees_d;
    size_t n_alloc_l, n_alloc_i, n_alloc_d;
} MetaBlock;

static void metablock_free(MetaBlock* m) {
    for (int k = 0; k < 3; k++) if (m->cat[k].has_trees) { tree_free(&m->cat[k].types); tree_free(&m->cat[k].counts); }
    free(m->context_modes); free(m->c_map_l); free(m->c_map_d);
    for (size_t i = 0; i < m->n_alloc_l; i++) tree_free(&m->trees_l[i]);
    for (size_t i = 0; i < m->n_alloc_i; i++) tree_free(&m->trees_i[i]);
    for (size_t i = 0; i < m->n_alloc_d; i++) tree_free(&m->trees_d[i]);
    free(m->trees_l); free(m->trees_i); free(m->trees_d);
}

static int emit(Dec* d, uint8_t b) { /* the per-byte output path of src/lib.rs:2057-2067 / 2110-2124 / 1718-1728 */
    int st = sink_push(d, b);
    if (st) return st;
    ring_push(&d->window, b);
    d->count_output += 1;
    return 0;
}

/* One compressed meta-block after MLEN/ISUNCOMPRESSED: src/lib.rs:1745-2141 */
static int decode_compressed_metablock(Dec* d, MetaBlock* m) {
    int st;
    /* NBLTYPES{L,I,D} + block type / count codes + first block count: src/lib.rs:1745-1885 */
    for (int k = 0; k < 3; k++) {
        BlockCat* c = &m->cat[k];
        c->btype = 0; c->btype_prev = 1; c->has_blen = 0; c->has_trees = 0;
        if ((st = parse_n_bltypes(d, &c->n_bltypes))) return st;
        if (c->n_bltypes >= 2) {
            memset(&c->types, 0, sizeof(Tree)); memset(&c->counts, 0, sizeof(Tree));
            c->has_trees = 1;
            if ((st = parse_prefix_code(d, c->n_bltypes + 2, &c->types))) return st;
            if ((st = parse_prefix_code(d, 26, &c->counts))) return st;
            if ((st = parse_block_count(d, &c->counts, &c->blen))) return st;
            c->has_blen = 1;
        }
    }
    /* NPOSTFIX, NDIRECT: src/lib.rs:548-560 */
    int64_t v = br_read_bits(&d->in, 2);
    if (v < 0) return ST_UnexpectedEOF;
    m->n_postfix = (unsigned)v;
    v = br_read_bits(&d->in, 4);
    if (v < 0) return ST_UnexpectedEOF;
    m->n_direct = (unsigned)((uint8_t)((unsigned)v << m->n_postfix)); /* u8 arithmetic: 15<<3 = 120 fits */
    /* context modes: src/lib.rs:562-573 */
    unsigned nbl = m->cat[0].n_bltypes;
    m->context_modes = (uint8_t*)calloc(nbl, 1);
    if (!m->context_modes) return ST_OutOfMemory;
    for (unsigned i = 0; i < nbl; i++) {
        v = br_read_bits(&d->in, 2);
        if (v < 0) return ST_UnexpectedEOF;
        m->context_modes[i] = (uint8_t)v;
    }
    /* NTREESL + cmap: src/lib.rs:1916-1943 */
    if ((st = parse_n_bltypes(d, &m->n_trees_l))) return st;
    m->c_map_l = (uint8_t*)calloc(64 * (size_t)nbl, 1);
    if (!m->c_map_l) return ST_OutOfMemory;
    if (m->n_trees_l >= 2 && (st = parse_context_map(d, m->n_trees_l, 64 * (size_t)nbl, m->c_map_l))) return st;
    /* NTREESD + cmap: src/lib.rs:1944-1973 */
    if ((st = parse_n_bltypes(d, &m->n_trees_d))) return st;
    unsigned nbd = m->cat[2].n_bltypes;
    m->c_map_d = (uint8_t*)calloc(4 * (size_t)nbd, 1);
    if (!m->c_map_d) return ST_OutOfMemory;
    if (m->n_trees_d >= 2 && (st = parse_context_map(d, m->n_trees_d, 4 * (size_t)nbd, m->c_map_d))) return st;
    /* prefix codes: src/lib.rs:1016-1068 */
    m->trees_l = (Tree*)calloc(m->n_trees_l, sizeof(Tree));
    m->trees_i = (Tree*)calloc(m->cat[1].n_bltypes, sizeof(Tree));
    m->trees_d = (Tree*)calloc(m->n_trees_d, sizeof(Tree));
    if (!m->trees_l || !m->trees_i || !m->trees_d) return ST_OutOfMemory;
    for (unsigned i = 0; i < m->n_trees_l; i++) { m->n_alloc_l = i + 1; if ((st = parse_prefix_code(d, 256, &m->trees_l[i]))) return st; }
    for (unsigned i = 0; i < m->cat[1].n_bltypes; i++) { m->n_alloc_i = i + 1; if ((st = parse_prefix_code(d, 704, &m->trees_i[i]))) return st; }
    size_t dist_alphabet = 16 + m->n_direct + ((size_t)48 << m->n_postfix);
    for (unsigned i = 0; i < m->n_trees_d; i++) { m->n_alloc_d = i + 1; if ((st = parse_prefix_code(d, dist_alphabet, &m->trees_d[i]))) return st; }

    uint8_t* scratch = NULL; size_t scratch_cap = 0;
    /* command loop: src/lib.rs:2003-2141 */
    for (;;) {
        /* parse_insert_and_copy_length: src/lib.rs:1179-1208 */
        if ((st = step_block(d, &m->cat[1]))) goto out;
        int32_t sym = tree_lookup_symbol(&m->trees_i[m->cat[1].btype], &d->in);
        if (sym == LOOKUP_NONE) { st = ST_ParseErrorInsertAndCopyLength; goto out; }
        if (sym < 0) { st = ST_UnexpectedEOF; goto out; }
        int implicit_zero = sym <= 127; /* src/lib.rs:2012-2015 */
        /* decode_insert_and_copy_length: src/lib.rs:1210-1224 */
        uint32_t insert_length = bro_ic_insert[sym] & 0xffff, copy_length = bro_ic_copy[sym] & 0xffff;
        int64_t e = br_read_bits(&d->in, bro_ic_insert[sym] >> 16);
        if (e < 0) { st = ST_UnexpectedEOF; goto out; }
        insert_length += (uint32_t)e;
        e = br_read_bits(&d->in, bro_ic_copy[sym] >> 16);
        if (e < 0) { st = ST_UnexpectedEOF; goto out; }
        copy_length += (uint32_t)e;
        /* src/lib.rs:2036-2039 */
        if ((size_t)m->m_len < m->count_output + insert_length) { st = ST_ExceededExpectedBytes; goto out; }
        /* parse_insert_literals: src/lib.rs:1286-1365 (literal_buf updated here, output in 2057-2067) */
        if (insert_length > scratch_cap) {
            free(scratch); scratch_cap = insert_length; scratch = (uint8_t*)malloc(scratch_cap);
            if (!scratch) { st = ST_OutOfMemory; goto out; }
        }
        for (uint32_t k = 0; k < insert_length; k++) {
            if ((st = step_block(d, &m->cat[0]))) goto out;
            unsigned btype = m->cat[0].btype;
            unsigned mode = m->context_modes[btype];
            unsigned p1 = d->lit_buf[0], p2 = d->lit_buf[1], cid;
            switch (mode) {
            case 0: cid = p1 & 0x3f; break;
            case 1: cid = p1 >> 2; break;
            case 2: cid = bro_lut0[p1] | bro_lut1[p2]; break;
            default: cid = ((unsigned)bro_lut2[p1] << 3) | bro_lut2[p2]; break;
            }
            unsigned index = m->c_map_l[btype * 64 + cid];
            /* the reference would panic on index >= NTREESL (Vec index); a context map symbol never exceeds
             * NTREESL-1 because the alphabet is RLEMAX+NTREESL and invalid symbols cannot be decoded, but IMTF
             * can permute only within 0..255 -- guard to stay memory-safe */
            if (index >= m->n_trees_l) { st = ST_ParseErrorInsertLiterals; goto out; }
            int32_t lit = tree_lookup_symbol(&m->trees_l[index], &d->in);
            if (lit == LOOKUP_NONE) { st = ST_ParseErrorInsertLiterals; goto out; }
            if (lit < 0) { st = ST_UnexpectedEOF; goto out; }
            scratch[k] = (uint8_t)lit;
            litbuf_push(d, (uint8_t)lit);
        }
        /* State::InsertLiterals: src/lib.rs:2048-2081 */
        for (uint32_t k = 0; k < insert_length; k++) {
            if ((st = emit(d, scratch[k]))) goto out;
            m->count_output += 1;
        }
        if ((size_t)m->m_len == m->count_output) { st = 0; goto out; }
        /* parse_distance_code: src/lib.rs:1367-1410 */
        uint32_t distance_code;
        if (implicit_zero) distance_code = 0;
        else {
            if ((st = step_block(d, &m->cat[2]))) goto out;
            unsigned cid = copy_length <= 4 ? copy_length - 2 : 3;
            unsigned index = m->c_map_d[m->cat[2].btype * 4 + cid];
            if (index >= m->n_trees_d) { st = ST_ParseErrorDistanceCode; goto out; }
            int32_t s = tree_lookup_symbol(&m->trees_d[index], &d->in);
            if (s == LOOKUP_NONE) { st = ST_ParseErrorDistanceCode; goto out; }
            if (s < 0) { st = ST_UnexpectedEOF; goto out; }
            distance_code = (uint32_t)s;
        }
        /* decode_distance: src/lib.rs:1412-1481 */
        uint32_t distance;
        if (distance_code <= 3) distance = d->dist_buf[distance_code];
        else if (distance_code <= 9) {
            int64_t sign = 2 * (int64_t)(distance_code % 2) - 1, dd = (distance_code - 2) >> 1;
            int64_t r = (int64_t)d->dist_buf[0] + sign * dd;
            if (r <= 0) { st = ST_InvalidNonPositiveDistance; goto out; }
            distance = (uint32_t)r;
        } else if (distance_code <= 15) {
            int64_t sign = 2 * (int64_t)(distance_code % 2) - 1, dd = (distance_code - 8) >> 1;
            int64_t r = (int64_t)d->dist_buf[1] + sign * dd;
            if (r <= 0) { st = ST_InvalidNonPositiveDistance; goto out; }
            distance = (uint32_t)r;
        } else if (distance_code <= 15 + m->n_direct) distance = distance_code - 15;
        else {
            uint32_t n_direct = m->n_direct, n_postfix = m->n_postfix;
            uint32_t ndistbits = 1 + ((distance_code - n_direct - 16) >> (n_postfix + 1));
            int64_t dextra = br_read_bits(&d->in, ndistbits);
            if (dextra < 0) { st = ST_UnexpectedEOF; goto out; }
            uint32_t hcode = (distance_code - n_direct - 16) >> n_postfix;
            uint32_t lcode = (distance_code - n_direct - 16) & ((1u << n_postfix) - 1);
            uint32_t offset = ((2 + (hcode & 1)) << ndistbits) - 4;
            distance = ((offset + (uint32_t)dextra) << n_postfix) + lcode + n_direct + 1;
        }
        size_t max_allowed = d->window_size < d->count_output ? d->window_size : d->count_output;
        if (distance_code > 0 && (size_t)distance <= max_allowed) distbuf_push(d, distance); /* src/lib.rs:1476-1478 */
        /* copy_literals: src/lib.rs:1483-1542 */
        uint8_t word_buf[64];
        uint8_t* copy = NULL; size_t copy_n = 0;
        if ((size_t)distance <= max_allowed) {
            if (copy_length > scratch_cap) {
                free(scratch); scratch_cap = copy_length; scratch = (uint8_t*)malloc(scratch_cap);
                if (!scratch) { st = ST_OutOfMemory; goto out; }
            }
            size_t l = distance < copy_length ? distance : copy_length;
            if (ring_slice_tail(&d->window, (size_t)distance - 1, scratch, copy_length)) { st = ST_RingBufferError; goto out; }
            for (size_t k = l; k < copy_length; k++) scratch[k] = scratch[k % l];
            copy = scratch; copy_n = copy_length;
        } else {
            if (copy_length < 4 || copy_length > 24) { st = ST_InvalidLengthInStaticDictionary; goto out; }
            size_t word_id = (size_t)distance - max_allowed - 1;
            unsigned bits = bro_dict_size_bits[copy_length];
            size_t index = word_id % ((size_t)1 << bits);
            size_t offset_from = bro_dict_offsets[copy_length] + index * copy_length;
            size_t transform_id = word_id >> bits;
            if (transform_id > 120) { st = ST_InvalidTransformId; goto out; }
            long n = transformation((unsigned)transform_id, bro_dictionary_blob + offset_from, copy_length, word_buf, d->quirk_spec);
            if (n < 0) { st = ST_PanicUppercaseZero; goto out; }
            copy = word_buf; copy_n = (size_t)n;
        }
        /* State::CopyLiterals: src/lib.rs:2102-2141 */
        if ((size_t)m->m_len < m->count_output + copy_n) { st = ST_ExceededExpectedBytes; goto out; }
        for (size_t k = 0; k < copy_n; k++) {
            litbuf_push(d, copy[k]);
            if ((st = emit(d, copy[k]))) goto out;
            m->count_output += 1;
        }
        if ((size_t)m->m_len == m->count_output) { st = 0; goto out; }
    }
out:
    free(scratch);
    return st;
}

/* The stream state machine, src/lib.rs:1545-2170, as straight-line code. */
static int decompress(Dec* d) {
    int st;
    /* parse_wbits: src/lib.rs:412-418, 1560-1568 */
    int32_t w = tree_lookup_symbol(&d->wbits_codes, &d->in);
    if (w < 0) return ST_UnexpectedEOF;
    d->window_size = ((size_t)1 << w) - 16;
    d->window.cap = d->window_size;
    d->window.buf = (uint8_t*)malloc(d->window_size);
    if (!d->window.buf) return ST_OutOfMemory;
    d->window.len = 0; d->window.pos = 0;

    for (;;) {
        MetaBlock m;
        memset(&m, 0, sizeof(m));
        /* ISLAST / ISLASTEMPTY: src/lib.rs:1572-1616 */
        int b = br_read_bit(&d->in);
        if (b < 0) return ST_UnexpectedEOF;
        m.is_last = b;
        if (m.is_last) {
            b = br_read_bit(&d->in);
            if (b < 0) return ST_UnexpectedEOF;
            if (b) break; /* -> StreamEnd */
        }
        /* MNIBBLES: src/lib.rs:434-440 */
        int64_t v = br_read_bits(&d->in, 2);
        if (v < 0) return ST_UnexpectedEOF;
        unsigned m_nibbles = v == 3 ? 0 : (unsigned)v + 4;
        if (m_nibbles == 0) {
            /* metadata block: src/lib.rs:1617-1683 */
            b = br_read_bit(&d->in);
            if (b < 0) return ST_UnexpectedEOF;
            if (b) return ST_NonZeroReservedBit;
            v = br_read_bits(&d->in, 2);
            if (v < 0) return ST_UnexpectedEOF;
            unsigned m_skip_bytes = (unsigned)v;
            if (m_skip_bytes == 0) {
                int t = br_read_byte_tail(&d->in);
                if (t < 0) return ST_UnexpectedEOF;
                if (t != 0) return ST_NonZeroFillBit;
            } else {
                /* parse_m_skip_len: src/lib.rs:449-467; every error becomes UnexpectedEOF at :1661-1664 (Q2) */
                uint8_t bytes[3];
                for (unsigned i = 0; i < m_skip_bytes; i++) {
                    int by = br_read_u8(&d->in);
                    if (by < 0) return ST_UnexpectedEOF;
                    bytes[i] = (uint8_t)by;
                }
                if (m_skip_bytes > 1 && bytes[m_skip_bytes - 1] == 0) return ST_UnexpectedEOF;
                uint32_t m_skip_len = 0;
                for (unsigned i = 0; i < m_skip_bytes; i++)
                    m_skip_len |= ((uint32_t)bytes[i]) << (d->quirk_spec ? 8 * i : i); /* sic: << i (Q1), :463 */
                m_skip_len += 1;
                int t = br_read_byte_tail(&d->in);
                if (t < 0) return ST_UnexpectedEOF;
                if (t != 0) return ST_NonZeroFillBit;
                for (uint32_t i = 0; i < m_skip_len; i++)
                    if (br_read_u8(&d->in) < 0) return ST_UnexpectedEOF;
            }
            if (m.is_last) break;
            continue;
        }
        /* MLEN: src/lib.rs:469-483 */
        v = br_read_nibbles(&d->in, m_nibbles);
        if (v < 0) return ST_UnexpectedEOF;
        if (m_nibbles > 4 && (((uint32_t)v) >> ((m_nibbles - 1) * 4)) == 0) return ST_NonZeroTrailerNibble;
        m.m_len = (uint32_t)v + 1;
        if (!m.is_last) {
            /* ISUNCOMPRESSED: src/lib.rs:1694-1734 */
            b = br_read_bit(&d->in);
            if (b < 0) return ST_UnexpectedEOF;
            if (b) {
                int t = br_read_byte_tail(&d->in);
                if (t < 0) return ST_UnexpectedEOF;
                if (t != 0) return ST_NonZeroFillBit;
                /* parse_mlen_literals reads all MLEN bytes before any is emitted: src/lib.rs:492-499 */
                uint8_t* lits = (uint8_t*)malloc(m.m_len);
                if (!lits) return ST_OutOfMemory;
                for (uint32_t i = 0; i < m.m_len; i++) {
                    int by = br_read_u8(&d->in);
                    if (by < 0) { free(lits); return ST_UnexpectedEOF; }
                    lits[i] = (uint8_t)by;
                }
                for (uint32_t i = 0; i < m.m_len; i++) {
                    if ((st = emit(d, lits[i]))) { free(lits); return st; }
                    litbuf_push(d, lits[i]);
                }
                free(lits);
                continue;
            }
        }
        st = decode_compressed_metablock(d, &m);
        metablock_free(&m);
        if (st) return st;
        if (m.is_last) break;
    }
    /* StreamEnd: src/lib.rs:2155-2167 */
    int t = br_read_byte_tail(&d->in);
    if (t < 0) return ST_UnexpectedEOF;
    if (t != 0) return ST_NonZeroTrailerBit;
    int by = br_read_u8(&d->in);
    if (by == BR_EOF) return ST_OK;
    if (by >= 0) return ST_ExpectedEndOfStream;
    return ST_UnexpectedEOF;
}

static int dec_run(const uint8_t* in, size_t in_len, Sink sink, int quirk_spec, Sink* sink_out) {
    Dec d;
    memset(&d, 0, sizeof(d));
    br_init(&d.in, in, in_len);
    d.lit_buf[0] = d.lit_buf[1] = 0;                                   /* src/lib.rs:407 */
    d.dist_buf[0] = 4; d.dist_buf[1] = 11; d.dist_buf[2] = 15; d.dist_buf[3] = 16; /* src/lib.rs:408 */
    d.out = sink;
    d.quirk_spec = quirk_spec;
    int st;
    if (make_wbits_tree(&d.wbits_codes) || make_code_length_tree(&d.bit_lengths_code) || make_bltype_tree(&d.bltype_codes))
        st = ST_OutOfMemory;
    else
        st = decompress(&d);
    tree_free(&d.wbits_codes); tree_free(&d.bit_lengths_code); tree_free(&d.bltype_codes);
    free(d.window.buf);
    *sink_out = d.out;
    return st;
}

/* ------------------------------------------------------------------------------------------------ */
/* Exported entry points (ctypes)                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* Decode one stream; *out is malloc'd (free with bro_oracle_free).  On error *out holds the bytes produced
 * before the error (not part of the parity contract).  quirks: 0 = reference, 1 = spec. */
int bro_oracle_decode(const uint8_t* in, size_t in_len, uint8_t** out, size_t* out_len, int quirks) {
    Sink s = {NULL, 0, 0, 1}, r;
    int st = dec_run(in, in_len, s, quirks, &r);
    *out = r.data; *out_len = r.len;
    return st;
}

void bro_oracle_free(void* p) { free(p); }

typedef struct {
    const uint8_t* in; const uint64_t* in_off; uint8_t* out; const uint64_t* out_off;
    uint64_t* out_len; int32_t* status; uint32_t n; volatile uint32_t* next; int quirks;
} BatchJob;

static void* batch_worker(void* arg) {
    BatchJob* j = (BatchJob*)arg;
    for (;;) {
        uint32_t i = __sync_fetch_and_add(j->next, 1);
        if (i >= j->n) break;
        Sink s = {j->out + j->out_off[i], 0, (size_t)(j->out_off[i + 1] - j->out_off[i]), 0}, r;
        j->status[i] = dec_run(j->in + j->in_off[i], (size_t)(j->in_off[i + 1] - j->in_off[i]), s, j->quirks, &r);
        j->out_len[i] = r.len;
    }
    return NULL;
}

/* Batch form with the same argument meaning as bro_batch_decode_host (include/brotli_b200.h); streams are
 * handed to `nthreads` host threads one at a time.  This is the CPU baseline leg. */
int bro_oracle_decode_batch(const uint8_t* in, const uint64_t* in_off, uint8_t* out, const uint64_t* out_off,
                            uint64_t* out_len, int32_t* status, uint32_t n, int nthreads, int quirks) {
    volatile uint32_t next = 0;
    BatchJob j = {in, in_off, out, out_off, out_len, status, n, &next, quirks};
    if (nthreads <= 1) { batch_worker(&j); return 0; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return -1;
    int started = 0;
    for (int t = 0; t < nthreads; t++) if (pthread_create(&th[t], NULL, batch_worker, &j) == 0) started++; else break;
    if (started == 0) batch_worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* ---- unit-test hooks for the KATs of the reference's private modules ---- */

/* transformation(id, word): returns output length, or -1 where the reference panics. */
long bro_oracle_transform(unsigned id, const uint8_t* word, size_t len, uint8_t* out, int quirks) {
    if (id > 120) return -2;
    return transformation(id, word, len, out, quirks);
}

void bro_oracle_imtf(uint8_t* v, size_t n) { inverse_move_to_front_transform(v, n); }

/* Bit-reader script: ops[i] in {0:u8, 1:bit, 2:bits(arg), 3:tail, 4:nibble, 5:nibbles(arg)}; results[i] gets
 * the value or a negative error. */
void bro_oracle_bitreader_script(const uint8_t* data, size_t n, const int* ops, const int* args, int64_t* results, size_t nops) {
    BitReader br;
    br_init(&br, data, n);
    for (size_t i = 0; i < nops; i++) {
        switch (ops[i]) {
        case 0: results[i] = br_read_u8(&br); break;
        case 1: results[i] = br_read_bit(&br); break;
        case 2: results[i] = br_read_bits(&br, (unsigned)args[i]); break;
        case 3: results[i] = br_read_byte_tail(&br); break;
        case 4: results[i] = br_read_nibble(&br); break;
        default: results[i] = br_read_nibbles(&br, (unsigned)args[i]); break;
        }
    }
}

/* Build a prefix tree from code lengths and decode `nsyms` symbols from `data` (tree KATs,
 * src/huffman/tree/mod.rs:96-212, and property tests against the CUDA table builder). */
int bro_oracle_tree_decode(const unsigned* lengths, size_t alphabet, const uint8_t* data, size_t n,
                           int32_t* syms, size_t nsyms) {
    Tree t;
    BitReader br;
    if (codes_from_lengths(&t, lengths, alphabet)) return -1;
    br_init(&br, data, n);
    for (size_t i = 0; i < nsyms; i++) syms[i] = tree_lookup_symbol(&t, &br);
    tree_free(&t);
    return 0;
}
